"""Drop-in for the reference's ``Retrieval`` package (``Retrieval/__init__.py:1-3``)."""
from .retrieval import B200RetrievalEngine, RetrievalEngine, make_retrieval_engine
from .reranker import Reranker

__all__ = ["RetrievalEngine", "Reranker", "make_retrieval_engine", "B200RetrievalEngine"]
