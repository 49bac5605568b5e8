"""CPU oracle for the retrieval hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import anything from this package, and only as the
checker (or the timed CPU baseline), never as the thing shipped.  The product
package (``multi_modal_retrieval_predict_project_b200``) never imports it and has
no CPU fallback: it raises if the CUDA library is missing.

What is here
------------
* ``ref_loader``  -- imports the UNMODIFIED reference modules from /root/reference
  through two namespace stubs (works only in the authoring container; the GPU box
  has no /root/reference).  Used by ``tests/golden/make_golden.py`` to generate the
  committed golden vectors and by the CPU tests (when the reference is present) to
  re-validate the restatement.
* ``search`` / ``rerank`` / ``metrics`` / ``gt`` -- numpy / pure-Python restatements
  of the reference algorithm, each function citing the reference file:line it
  follows.  These travel to the GPU box and are what the ``-m gpu`` parity tests
  compare the CUDA path with.

Parity pinning
--------------
The reference ships no tests, golden vectors or fixtures for this path
(SURVEY.md section 8c), so the restatement is pinned against OUTPUTS OF THE
REFERENCE ITSELF run in the authoring container: ``tests/golden/*.npz|json`` were
produced by ``tests/golden/make_golden.py`` executing the real
``Retrieval/retrieval.py``, ``Retrieval/reranker.py``,
``Helpers/retrieval_metrics.py``, ``Helpers/contructGT.py`` logic and
``sklearn.metrics.pairwise.cosine_similarity`` on seeded synthetic inputs.
``tests/test_oracle_golden.py`` checks every oracle function against them.
"""
