"""Oracle: retrieval metrics (TEST INFRASTRUCTURE; never imported by the product).

Pure-Python restatement of ``Helpers/retrieval_metrics.py`` of the reference, with the
exact operation order (fp64, sequential sums) so the device kernels can be compared
bit for bit.  ``recall_at_k`` follows the SECOND definition (``:74-79``), which shadows
the first (``:13-22``) at import time.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np


def precision_at_k(retrieved_ids, relevant_ids, k=5):
    """``retrieval_metrics.py:4-11``: hits in the first k (duplicates counted) / k."""
    rel = set(relevant_ids)
    return sum(1 for r in retrieved_ids[:k] if r in rel) / k


def recall_at_k(retrieved, relevant, k=5):
    """``retrieval_metrics.py:74-79``: |set(top-k) & set(rel)| / |set(rel)|; 0.0 if
    ``relevant`` is empty."""
    if len(relevant) == 0:
        return 0.0
    return len(set(retrieved[:k]) & set(relevant)) / len(set(relevant))


def average_precision(retrieved, relevant, k: Optional[int] = None) -> float:
    """``retrieval_metrics.py:24-38``: sum over hit ranks i of hits_i/i, divided by
    ``len(relevant)`` (the container as passed: a list keeps its duplicates)."""
    if k is None:
        k = len(retrieved)
    hits = 0
    score = 0.0
    for i, r in enumerate(retrieved[:k], start=1):
        if r in relevant:
            hits += 1
            score += hits / i
    return score / len(relevant) if relevant else 0.0


def mean_average_precision(all_retrieved, all_relevant, k: Optional[int] = None) -> float:
    """``retrieval_metrics.py:40-54``."""
    return float(np.mean([average_precision(a, b, k) for a, b in zip(all_retrieved, all_relevant)]))


def reciprocal_rank(retrieved, relevant) -> float:
    """inner loop of ``mean_reciprocal_rank`` (``retrieval_metrics.py:65-71``): first
    hit over the WHOLE retrieved list."""
    for i, r in enumerate(retrieved, start=1):
        if r in relevant:
            return 1.0 / i
    return 0.0


def mean_reciprocal_rank(all_retrieved, all_relevant) -> float:
    """``retrieval_metrics.py:56-72``."""
    return float(np.mean([reciprocal_rank(a, b) for a, b in zip(all_retrieved, all_relevant)]))


def ndcg_at_k(retrieved, relevant, k=5):
    """``retrieval_metrics.py:81-89``: binary gains, ``sum(score/np.log2(idx+2))``
    sequentially from integer 0, ideal = the SAME hit list sorted descending."""
    def dcg(scores):
        return sum(score / np.log2(idx + 2) for idx, score in enumerate(scores))
    scores = [1 if r in relevant else 0 for r in retrieved[:k]]
    ideal = sorted(scores, reverse=True)
    d = dcg(scores)
    i = dcg(ideal)
    return d / i if i > 0 else 0.0


def per_query_table(all_retrieved: Sequence[Sequence], all_relevant: Sequence[Sequence], k: int):
    """(Q,5) fp64 table [P@k, R@k, AP@k, RR, nDCG@k] -- the layout the device metrics
    kernel returns; aggregated by the callers with ``np.mean`` exactly as
    ``Evaluate/retrieval_eval.py:147-160`` does."""
    out = np.zeros((len(all_retrieved), 5), dtype=np.float64)
    for i, (ret, rel) in enumerate(zip(all_retrieved, all_relevant)):
        out[i, 0] = precision_at_k(ret, rel, k)
        out[i, 1] = recall_at_k(ret, rel, k)
        out[i, 2] = average_precision(ret, rel, k)
        out[i, 3] = reciprocal_rank(ret, rel)
        out[i, 4] = ndcg_at_k(ret, rel, k)
    return out
