"""Drop-in for the retrieval-metric re-exports of the reference's ``Helpers`` package
(``Helpers/__init__.py:11,37-42``) and its embedding-dump merge (``Helpers/dumpEmbedding.py``)."""
from .dumpEmbedding import createDumpEmbedding, merged_engine
from .retrieval_metrics import (average_precision, evaluate_retrieval, mean_average_precision, mean_reciprocal_rank,
                                ndcg_at_k, metrics_from_rows, per_query_metrics, precision_at_k, recall_at_k)

__all__ = ["precision_at_k", "recall_at_k", "average_precision", "mean_average_precision", "mean_reciprocal_rank",
           "ndcg_at_k", "per_query_metrics", "evaluate_retrieval", "metrics_from_rows", "createDumpEmbedding", "merged_engine"]
