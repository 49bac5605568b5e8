"""Drop-in for the reference's ``Retrieval`` package.

The reference exports three names from here (``Retrieval/__init__.py:1-3``); the same three resolve to
the B200 implementations: the engine factory and abstract engine (``retrieval.py``: exact search on the
device behind the DLS engine's ``retrieve`` signature) and the label / knowledge-graph reranker
(``reranker.py``: device tables + rerank kernels).  ``B200RetrievalEngine`` is exported in addition for
callers that want the batched ``search`` entry point, ``save_blob`` / ``load_blob`` and ``from_arrays``;
``MultiGPURetrievalEngine`` is what ``make_retrieval_engine(..., devices=[...])`` returns.
"""
from . import reranker as _reranker
from . import retrieval as _retrieval

RetrievalEngine = _retrieval.RetrievalEngine
B200RetrievalEngine = _retrieval.B200RetrievalEngine
MultiGPURetrievalEngine = _retrieval.MultiGPURetrievalEngine
make_retrieval_engine = _retrieval.make_retrieval_engine
Reranker = _reranker.Reranker

__all__ = ("B200RetrievalEngine", "MultiGPURetrievalEngine", "Reranker", "RetrievalEngine", "make_retrieval_engine")
