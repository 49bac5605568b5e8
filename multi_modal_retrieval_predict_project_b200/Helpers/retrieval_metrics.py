"""B200-native mirror of the reference's ``Helpers/retrieval_metrics.py``.

Same free functions, same signatures (lists / sets of ids in, Python floats out), same quirks
(P@k divides by k; recall de-duplicates; AP divides by ``len(relevant)`` as passed; nDCG's ideal
ranking is the hit list itself sorted), but the per-query work runs in csrc/metrics.cu: ids are
interned to integers on the host (dict lookups -- the only host work), one warp per query does
the membership tests and replays the reference's fp64 arithmetic in order, so per-query values
are bit-identical.  Cross-query means are taken with ``np.mean`` exactly as the callers do
(``Evaluate/retrieval_eval.py:147-160``).  There is no host fallback, even for tiny inputs.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np

from .. import _lib

_P, _R, _AP, _RR, _NDCG = range(5)


def _intern(all_retrieved: Sequence[Sequence], all_relevant: Sequence[Iterable]):
    """String ids -> dense ints; CSR of sorted unique relevant ids + len(relevant) as passed."""
    table: Dict = {}
    nq = len(all_retrieved)
    k_ret = max((len(r) for r in all_retrieved), default=0)
    ret = -np.ones((nq, max(k_ret, 1)), dtype=np.int64)
    cnt = np.zeros(nq, dtype=np.int32)
    indptr = np.zeros(nq + 1, dtype=np.int64)
    list_len = np.zeros(nq, dtype=np.int64)
    rel_chunks: List[np.ndarray] = []
    for i, (r, rel) in enumerate(zip(all_retrieved, all_relevant)):
        cnt[i] = len(r)
        for j, x in enumerate(r):
            ret[i, j] = table.setdefault(x, len(table))
        list_len[i] = len(rel)
        u = np.unique(np.fromiter((table.setdefault(x, len(table)) for x in rel), dtype=np.int64, count=len(rel)))
        rel_chunks.append(u)
        indptr[i + 1] = indptr[i] + len(u)
    rel_sorted = np.concatenate(rel_chunks) if rel_chunks else np.zeros(0, dtype=np.int64)
    if rel_sorted.size == 0:
        rel_sorted = np.zeros(1, dtype=np.int64)
    return ret, cnt, indptr, np.ascontiguousarray(rel_sorted), list_len


def per_query_metrics(all_retrieved: Sequence[Sequence], all_relevant: Sequence[Iterable], k: Optional[int],
                      device=None) -> np.ndarray:
    """(Q, 5) fp64 table ``[P@k, Recall@k, AP@k, RR, nDCG@k]`` from one kernel launch.
    ``k=None`` means "all retrieved" (``average_precision``'s default)."""
    import torch
    nq = len(all_retrieved)
    if nq == 0:
        return np.zeros((0, 5), dtype=np.float64)
    ret, cnt, indptr, rel_sorted, list_len = _intern(all_retrieved, all_relevant)
    k_ret = ret.shape[1]
    kk = int(k) if k is not None else max(int(cnt.max()), 1)
    if kk < 1:
        raise ZeroDivisionError("division by zero")  # what the reference's `/ k` raises for k == 0
    tbl = np.log2(np.arange(2, max(kk, k_ret) + 2))  # np.log2(idx + 2), reference :83
    out = np.empty((nq, 5), dtype=np.float64)
    dev = _lib.require_cuda(device)
    lib = _lib.load()
    with torch.cuda.device(dev):
        _lib.check(lib.mmr_metrics(_lib.ptr(ret), _lib.ptr(cnt), nq, k_ret, _lib.ptr(indptr), _lib.ptr(rel_sorted),
                                   _lib.ptr(list_len), kk, _lib.ptr(tbl), _lib.ptr(out), dev,
                                   _lib.current_stream(dev)))
    return out


def precision_at_k(retrieved_ids, relevant_ids, k=5):
    """Precision@k = (# relevant in top-k) / k   (reference :4-11)."""
    return float(per_query_metrics([list(retrieved_ids)], [list(relevant_ids)], k)[0, _P])


def recall_at_k(retrieved, relevant, k=5):
    """|set(top-k) & set(relevant)| / |set(relevant)|, 0.0 for empty relevant (reference :74-79)."""
    return float(per_query_metrics([list(retrieved)], [list(relevant)], k)[0, _R])


def average_precision(retrieved: List[str], relevant, k: int = None) -> float:
    """reference :24-38."""
    return float(per_query_metrics([list(retrieved)], [list(relevant)], k)[0, _AP])


def mean_average_precision(all_retrieved, all_relevant, k: int = None) -> float:
    """reference :40-54."""
    rets = [list(r) for r in all_retrieved]
    rels = [list(r) for r in all_relevant]
    if k is None:  # each query uses its own len(retrieved)
        return float(np.mean([per_query_metrics([a], [b], None)[0, _AP] for a, b in zip(rets, rels)]))
    return float(np.mean(per_query_metrics(rets, rels, k)[:, _AP]))


def mean_reciprocal_rank(all_retrieved, all_relevant) -> float:
    """reference :56-72."""
    rets = [list(r) for r in all_retrieved]
    rels = [list(r) for r in all_relevant]
    return float(np.mean(per_query_metrics(rets, rels, 1)[:, _RR]))


def ndcg_at_k(retrieved, relevant, k=5):
    """reference :81-89."""
    return float(per_query_metrics([list(retrieved)], [list(relevant)], k)[0, _NDCG])


def evaluate_retrieval(all_retrieved, all_relevant, k: int = 5) -> Dict[str, float]:
    """One launch for what ``Evaluate/retrieval_eval.py:147-160`` / ``retrieval_eval_variants.py:117-122``
    compute with five Python list comprehensions: mean P@k, R@k, mAP@k, MRR, nDCG@k."""
    t = per_query_metrics([list(r) for r in all_retrieved], [list(r) for r in all_relevant], k)
    return {"P": float(np.mean(t[:, _P])), "R": float(np.mean(t[:, _R])), "mAP": float(np.mean(t[:, _AP])),
            "MRR": float(np.mean(t[:, _RR])), "nDCG": float(np.mean(t[:, _NDCG]))}


def metrics_from_rows(retrieved_rows, rel_indptr, rel_sorted, k: int, rel_list_len=None, ret_count=None):
    """Device-native batched metrics (BASELINE cfg5: evaluation without string ids or host lists).

    ``retrieved_rows`` (Q, K) int64 row ids as returned by ``engine.search`` (-1 padding allowed),
    relevance as CSR over row ids: ``rel_indptr`` (Q+1) int64, ``rel_sorted`` per-query sorted unique
    int64 ids, optional ``rel_list_len`` (Q) = len(relevant) as passed (AP denominator; default =
    unique count).  numpy arrays or CUDA tensors.  Returns the (Q, 5) fp64 table
    ``[P@k, Recall@k, AP@k, RR, nDCG@k]`` of the same type family as the input."""
    import torch
    on_device = hasattr(retrieved_rows, "is_cuda") and retrieved_rows.is_cuda
    nq, k_ret = int(retrieved_rows.shape[0]), int(retrieved_rows.shape[1])
    kk = int(k)
    if kk < 1:
        raise ZeroDivisionError("division by zero")
    tbl = np.log2(np.arange(2, max(kk, k_ret) + 2))
    if on_device:
        dev = retrieved_rows.device.index or 0
        out = torch.empty((nq, 5), dtype=torch.float64, device=retrieved_rows.device)
        retrieved_rows = retrieved_rows.contiguous()
    else:
        dev = _lib.require_cuda(None)
        out = np.empty((nq, 5), dtype=np.float64)
        retrieved_rows = np.ascontiguousarray(retrieved_rows, dtype=np.int64)
    lib = _lib.load()
    with torch.cuda.device(dev):
        _lib.check(lib.mmr_metrics(_lib.ptr(retrieved_rows), _lib.ptr(ret_count), nq, k_ret, _lib.ptr(rel_indptr),
                                   _lib.ptr(rel_sorted), _lib.ptr(rel_list_len), kk, _lib.ptr(tbl), _lib.ptr(out), dev,
                                   _lib.current_stream(dev)))
    return out
