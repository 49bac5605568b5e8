"""B200-native retrieval hot path (search -> top-K -> rerank -> metrics) behind the reference's
Python entry points.  Importing the package does not touch CUDA; the first call that needs the
device loads libmmr_b200.so and raises if it (or a GPU) is missing -- there is no CPU fallback."""
from . import _lib  # noqa: F401
from .Retrieval import (B200RetrievalEngine, MultiGPURetrievalEngine, Reranker, RetrievalEngine,  # noqa: F401
                        make_retrieval_engine)

__all__ = ["RetrievalEngine", "B200RetrievalEngine", "MultiGPURetrievalEngine", "Reranker", "make_retrieval_engine"]
