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
    """The full reference checkout is mounted (authoring container)."""
    return os.path.isfile(os.path.join(_SRC, "Retrieval", "retrieval.py"))


def staged_available() -> bool:
    """The hot-path files packed by ``oracle/stage_ref.py`` are present (the archive travels to the GPU box)."""
    from .stage_ref import ARCHIVE
    return os.path.isfile(ARCHIVE)


_loaded = None


def load_reference(allow_staged: bool = False):
    """Return a namespace with the reference's hot-path symbols.

    Attributes: ``make_retrieval_engine``, ``RetrievalEngine``, ``DLSRetrievalEngine``,
    ``Reranker``, ``metrics`` (the ``Helpers.retrieval_metrics`` module),
    ``LabelAttention``, ``root`` (where the files were loaded from).
    With ``allow_staged`` the copies under ``oracle/_ref`` are used when ``/root/reference`` is absent
    (the GPU box); they are the same files, byte for byte (``oracle/_ref/MANIFEST.json``).
    """
    global _loaded
    if _loaded is not None:
        return _loaded
    if reference_available():
        src = _SRC
    elif allow_staged and staged_available():
        from .stage_ref import staged_root
        src = os.path.join(staged_root(), "src")
    else:
        raise RuntimeError(f"reference not present under {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True
    for pkg in ("KnowledgeGraph", "Helpers"):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = [os.path.join(src, pkg)]
            sys.modules[pkg] = m
    if src not in sys.path:
        sys.path.insert(0, src)
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
        root=os.path.dirname(src),
    )
    _loaded = ns
    return ns
