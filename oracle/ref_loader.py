"""Import the unmodified reference modules (authoring container only).

TEST INFRASTRUCTURE.  The reference's ``Retrieval`` package imports
``KnowledgeGraph.label_attention`` and ``Helpers.config``; the package
``__init__`` files of ``KnowledgeGraph`` and ``Helpers`` pull in ``pydicom`` (absent
here) through ``DataHandler/tensorDICOM.py:3``.  Registering empty namespace
modules for those two packages (with ``__path__`` pointing at the reference
directories) lets the real files ``Retrieval/retrieval.py``,
``Retrieval/reranker.py``, ``KnowledgeGraph/label_attention.py``,
``Helpers/config.py`` and ``Helpers/retrieval_metrics.py`` load and run as-is
(SURVEY.md section 8c).  Nothing is written under /root/reference
(``sys.dont_write_bytecode``; callers must pass a writable ``fdb_path`` to the DLS
engine, reference ``Retrieval/retrieval.py:72-82``).
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MMR_REFERENCE_ROOT", "/root/reference")
_SRC = os.path.join(REFERENCE_ROOT, "src")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(_SRC, "Retrieval", "retrieval.py"))


_loaded = None


def load_reference():
    """Return a namespace with the reference's hot-path symbols.

    Attributes: ``make_retrieval_engine``, ``RetrievalEngine``, ``DLSRetrievalEngine``,
    ``Reranker``, ``metrics`` (the ``Helpers.retrieval_metrics`` module),
    ``LabelAttention``.
    """
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"reference not present under {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True
    for pkg in ("KnowledgeGraph", "Helpers"):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = [os.path.join(_SRC, pkg)]
            sys.modules[pkg] = m
    if _SRC not in sys.path:
        sys.path.insert(0, _SRC)
    import importlib

    retrieval = importlib.import_module("Retrieval.retrieval")
    reranker = importlib.import_module("Retrieval.reranker")
    metrics = importlib.import_module("Helpers.retrieval_metrics")
    la = importlib.import_module("KnowledgeGraph.label_attention")
    ns = types.SimpleNamespace(
        make_retrieval_engine=retrieval.make_retrieval_engine,
        RetrievalEngine=retrieval.RetrievalEngine,
        DLSRetrievalEngine=retrieval.DLSRetrievalEngine,
        Reranker=reranker.Reranker,
        metrics=metrics,
        LabelAttention=la.LabelAttention,
        retrieval_module=retrieval,
        reranker_module=reranker,
    )
    _loaded = ns
    return ns
