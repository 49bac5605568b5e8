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
    """(MRR, Hit@k, mean Recall@k) of ``retrieval_overlap.py:84-115`` with the full ranking from the
    exact GPU search (K = N) instead of a materialised (Q, N) matrix + N-long argsorts."""
    from ..Retrieval import B200RetrievalEngine
    g = np.ascontiguousarray(gallery_embs, dtype=np.float32)
    q = np.ascontiguousarray(query_embs, dtype=np.float32)
    n = g.shape[0]
    if n > _lib.MAX_K:
        raise NotImplementedError("full-ranking metrics need K = N <= 1024 in this round")
    eng = B200RetrievalEngine.from_arrays(g, device=device)
    rows, _ = eng.search(q, n)
    eng.close()
    rel = relevance_matrix(np.asarray(query_labels), np.asarray(gallery_labels), False, device).astype(bool)
    rr, recalls, hits = [], [], 0
    for i in range(q.shape[0]):
        ranked_rel = rel[i][rows[i]]
        pos = np.nonzero(ranked_rel)[0]
        rank = int(pos[0]) + 1 if pos.size else None
        rr.append(1.0 / rank if rank else 0.0)
        if rank and rank <= k:
            hits += 1
        total = int(rel[i].sum())
        recalls.append(int(ranked_rel[:k].sum()) / total if total > 0 else 0.0)
    return np.mean(rr), hits / q.shape[0], np.mean(recalls)
