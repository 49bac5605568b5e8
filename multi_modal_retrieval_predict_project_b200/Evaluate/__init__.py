"""Callers either side of the hot path ("next" rows of SURVEY.md section 8f), on the device:

* ``relevance_lists``  -- the relevance definition of ``create_gt`` (reference
  ``Helpers/contructGT.py:68-81``): gallery item j is relevant to query i iff their multi-hot label
  vectors share a positive (self excluded for test->test).  Label vectors become uint64 bit masks;
  the (Q, N) overlap matrix is one kernel (csrc/metrics.cu).
* ``compute_ranking_metrics`` -- ``Evaluate/retrieval_overlap.py:84-115`` (MRR over the full
  ranking, Hit@k, Recall@k with label-overlap relevance): exact search for the whole ranking on
  the device + the same relevance kernel.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np

from .. import _lib


def label_masks(vals: np.ndarray) -> np.ndarray:
    """(n, L) multi-hot (== 1) -> (n, ceil(L/64)) uint64 bit masks."""
    vals = np.asarray(vals)
    n, L = vals.shape
    words = max(1, (L + 63) // 64)
    out = np.zeros((n, words), dtype=np.uint64)
    on = vals == 1
    for c in range(L):
        out[on[:, c], c // 64] |= np.uint64(1) << np.uint64(c % 64)
    return out


def relevance_matrix(q_vals, g_vals, exclude_self: bool, device=None) -> np.ndarray:
    import torch
    qm, gm = label_masks(q_vals), label_masks(g_vals)
    # contructGT.py:71 uses a bitwise AND of the INTEGER label values then `.sum(axis=1) > 0`; for
    # 0/1 labels that is exactly "share a positive"
    out = np.empty((qm.shape[0], gm.shape[0]), dtype=np.uint8)
    dev = _lib.require_cuda(device)
    lib = _lib.load()
    with torch.cuda.device(dev):
        _lib.check(lib.mmr_label_relevance(_lib.ptr(qm), qm.shape[0], _lib.ptr(gm), gm.shape[0], qm.shape[1],
                                           1 if exclude_self else 0, _lib.ptr(out), dev, _lib.current_stream(dev)))
    return out


def relevance_lists(query_vals, query_ids: Sequence[str], gallery_vals, gallery_ids: Sequence[str],
                    exclude_self: bool, device=None) -> Dict[str, List[str]]:
    """``test_relevance.json`` / ``test_to_train_relevance.json`` content (dict qid -> list of ids)."""
    rel = relevance_matrix(np.asarray(query_vals).astype(int), np.asarray(gallery_vals).astype(int),
                           exclude_self, device)
    return {qid: [gallery_ids[j] for j in np.nonzero(rel[i])[0]] for i, qid in enumerate(query_ids)}


def compute_ranking_metrics(query_embs, gallery_embs, query_labels, gallery_labels, k: int = 1, device=None):
    """(MRR, Hit@k, mean Recall@k) of ``retrieval_overlap.py:84-115``.  The rank of the first relevant
    item over the FULL ranking comes from ``mmr_first_relevant_rank`` (two sweeps over the gallery on the
    device, any N), the top-k from the exact search; no (Q, N) matrix, no N-long argsorts."""
    import torch
    from ..Retrieval import B200RetrievalEngine
    g = np.ascontiguousarray(gallery_embs, dtype=np.float32)
    q = np.ascontiguousarray(query_embs, dtype=np.float32)
    n, nq = g.shape[0], q.shape[0]
    qm = label_masks(np.asarray(query_labels))
    gm = label_masks(np.asarray(gallery_labels))
    eng = B200RetrievalEngine.from_arrays(g, device=device)
    rank = np.empty(nq, dtype=np.int64)
    total = np.empty(nq, dtype=np.int64)
    lib = _lib.load()
    with torch.cuda.device(eng.device):
        _lib.check(lib.mmr_first_relevant_rank(eng._handle, _lib.ptr(q), nq, _lib.MMR_F32, _lib.ptr(qm), _lib.ptr(gm),
                                               qm.shape[1], _lib.ptr(rank), _lib.ptr(total),
                                               _lib.current_stream(eng.device)))
    if min(int(k), n) > _lib.MAX_K:
        eng.close()
        # (the reference accepts any k; the search kernels keep at most MMR_MAX_K candidates per query, and a
        # Recall@k computed over a shorter list would be silently wrong)
        raise NotImplementedError(f"compute_ranking_metrics: k = {k} exceeds the {_lib.MAX_K} results the search "
                                  "kernels return per query")
    kk = min(int(k), n)
    rows, _ = eng.search(q, kk)
    eng.close()
    valid = rows >= 0                                              # relevance of the k retrieved rows
    overlap = np.zeros(rows.shape, dtype=bool)
    safe_rows = np.where(valid, rows, 0)
    for w in range(qm.shape[1]):
        overlap |= (gm[safe_rows, w] & qm[:, w][:, None]) != 0
    rel_topk = (overlap & valid).sum(axis=1)
    rr = np.where(rank > 0, 1.0 / np.maximum(rank, 1), 0.0)
    hits = int(((rank > 0) & (rank <= k)).sum())
    recalls = np.where(total > 0, rel_topk / np.maximum(total, 1), 0.0)
    # np.mean over per-query python floats in the reference == pairwise mean over these arrays
    return np.mean(rr.tolist()), hits / nq, np.mean(recalls.tolist())


def compute_embedding_diversity(embeddings, device=None) -> float:
    """``1 - mean pairwise cosine`` of one result set (``retrieval_diversity_compute.py:171-182``)."""
    import torch
    if embeddings is None or len(embeddings) < 2:
        return 0.0
    e = np.ascontiguousarray(embeddings, dtype=np.float32)
    out = np.empty(1, dtype=np.float64)
    dev = _lib.require_cuda(device)
    lib = _lib.load()
    with torch.cuda.device(dev):
        _lib.check(lib.mmr_result_diversity(_lib.ptr(e), None, None, 1, e.shape[0], e.shape[1], 0, _lib.ptr(out), None,
                                            dev, _lib.current_stream(dev)))
    return float(out[0])


def result_diversity_batch(cand_embs=None, cand_masks=None, counts=None, device=None):
    """Batched diversity of B result sets: ``cand_embs`` (B, K, D) fp32 and/or ``cand_masks``
    (B, K, words) uint64 -> ``(emb_div (B,), label_div (B,))`` fp64 (either may be ``None``)."""
    import torch
    dev = _lib.require_cuda(device)
    lib = _lib.load()
    b = int((cand_embs if cand_embs is not None else cand_masks).shape[0])
    k = int((cand_embs if cand_embs is not None else cand_masks).shape[1])
    e = np.ascontiguousarray(cand_embs, dtype=np.float32) if cand_embs is not None else None
    m = np.ascontiguousarray(cand_masks, dtype=np.uint64) if cand_masks is not None else None
    c = np.ascontiguousarray(counts, dtype=np.int32) if counts is not None else None
    oe = np.empty(b, dtype=np.float64) if e is not None else None
    ol = np.empty(b, dtype=np.float64) if m is not None else None
    with torch.cuda.device(dev):
        _lib.check(lib.mmr_result_diversity(_lib.ptr(e), _lib.ptr(m), _lib.ptr(c), b, k, e.shape[2] if e is not None else 0,
                                            m.shape[2] if m is not None else 0, _lib.ptr(oe), _lib.ptr(ol), dev,
                                            _lib.current_stream(dev)))
    return oe, ol


def compute_label_diversity_from_labels(labels_list, device=None) -> float:
    """``unique labels / avg per-item label count`` (``retrieval_diversity_compute.py:184-194``)."""
    if not labels_list:
        return 0.0
    names = sorted({l for lab in labels_list for l in lab})
    if not names:
        return 0.0
    idx = {nm: i for i, nm in enumerate(names)}
    words = (len(names) + 63) // 64
    m = np.zeros((1, len(labels_list), words), dtype=np.uint64)
    for i, lab in enumerate(labels_list):
        for l in set(lab):
            m[0, i, idx[l] // 64] |= np.uint64(1) << np.uint64(idx[l] % 64)
    _, ol = result_diversity_batch(None, m, None, device)
    return float(ol[0])


def evaluate_label_ranking(embs, label_vals, topk=(1, 5, 10), device=None, return_table: bool = False):
    """Retrieval metrics of ``evaluate_label_attention`` (reference
    ``Trainner/train_label_attention.py:106-125``) for record embeddings ``(n, d)`` and their multi-hot
    labels ``(n, L)``: ``{"recall@k": ..., "mAP": ...}`` -- all-pairs cosine, full ranking per record
    (self kept with label 0), mean relevance of the top k and sklearn's average precision, averaged over
    the records.  One kernel (``mmr_label_ranking_eval``): no (n, n) matrix, no argsort, no Python loop."""
    import torch
    e = np.ascontiguousarray(embs, dtype=np.float32)
    n, d = e.shape
    norms = np.ascontiguousarray(np.linalg.norm(e, axis=1), dtype=np.float32)      # :108
    masks = label_masks(np.asarray(label_vals).astype(int))
    ks = np.ascontiguousarray(list(topk), dtype=np.int32)
    if len(ks) > 8:
        raise ValueError("at most 8 cut-offs")
    out = np.zeros((n, 1 + len(ks)), dtype=np.float64)
    dev = _lib.require_cuda(device)
    lib = _lib.load()
    with torch.cuda.device(dev):
        _lib.check(lib.mmr_label_ranking_eval(_lib.ptr(e), _lib.ptr(norms), n, d, _lib.ptr(masks), masks.shape[1],
                                              _lib.ptr(ks), len(ks), _lib.ptr(out), dev, _lib.current_stream(dev)))
    results = {f"recall@{int(k)}": float(np.mean(out[:, 1 + t])) for t, k in enumerate(ks)}
    results["mAP"] = float(np.mean(out[:, 0]))
    return (results, out) if return_table else results
