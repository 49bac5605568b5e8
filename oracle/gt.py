"""Oracle: relevance ground truth + full-ranking metrics (TEST INFRASTRUCTURE).

Restates the relevance definition of ``create_gt`` (reference
``Helpers/contructGT.py:68-81``) and ``compute_ranking_metrics`` (reference
``Evaluate/retrieval_overlap.py:84-115``) on in-memory arrays.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np

from .search import cosine_similarity


def relevance_lists(query_vals: np.ndarray, query_ids: Sequence[str], gallery_vals: np.ndarray,
                    gallery_ids: Sequence[str], exclude_self: bool) -> Dict[str, List[str]]:
    """``contructGT.py:68-81``: gallery item j is relevant to query i iff their multi-hot
    label vectors share at least one positive; for test->test (``exclude_self``) the
    query's own position is dropped (``j != i``)."""
    query_vals = np.asarray(query_vals).astype(int)
    gallery_vals = np.asarray(gallery_vals).astype(int)
    out: Dict[str, List[str]] = {}
    for i, qid in enumerate(query_ids):
        shared = (gallery_vals & query_vals[i]).sum(axis=1) > 0
        out[qid] = [gallery_ids[j] for j, keep in enumerate(shared)
                    if keep and not (exclude_self and j == i)]
    return out


def compute_ranking_metrics(query_embs, gallery_embs, query_labels, gallery_labels, k=1):
    """``retrieval_overlap.py:84-115``: (MRR over the full ranking, Hit@k, mean Recall@k
    over ALL gallery items sharing a label).  Vectorised but arithmetically identical."""
    sim = cosine_similarity(query_embs, gallery_embs)
    n = sim.shape[1]
    ar = np.arange(n)
    rr, recalls, hits = [], [], 0
    ql = np.asarray(query_labels) == 1
    gl = np.asarray(gallery_labels) == 1
    for i in range(sim.shape[0]):
        idxs = np.lexsort((ar, -sim[i].astype(np.float64)))
        rel = (gl & ql[i]).any(axis=1)
        pos = np.nonzero(rel[idxs])[0]
        rank = int(pos[0]) + 1 if pos.size else None
        rr.append(1.0 / rank if rank else 0.0)
        if rank and rank <= k:
            hits += 1
        total = int(rel.sum())
        recalls.append(int(rel[idxs[:k]].sum()) / total if total > 0 else 0.0)
    return np.mean(rr), hits / sim.shape[0], np.mean(recalls)


def compute_embedding_diversity(embeddings) -> float:
    """``retrieval_diversity_compute.py:171-182``: 1 - mean pairwise cosine (rows normalised with the
    norm clipped at 1e-8; mean over the strict upper triangle; 0.0 for fewer than 2 items)."""
    if embeddings is None or len(embeddings) < 2:
        return 0.0
    norms = np.linalg.norm(embeddings, axis=1, keepdims=True).clip(min=1e-8)
    normed = embeddings / norms
    sim = np.dot(normed, normed.T)
    triu = np.triu_indices(sim.shape[0], k=1)
    return float(1.0 - float(np.mean(sim[triu])))


def compute_label_diversity_from_labels(labels_list) -> float:
    """``retrieval_diversity_compute.py:184-194``: |union of labels| / mean label count over the items
    that have labels; 0.0 when there are none."""
    if not labels_list:
        return 0.0
    all_labels = set(l for lab in labels_list for l in lab)
    sizes = [len(lab) for lab in labels_list if len(lab) > 0]
    if not sizes:
        return 0.0
    return float(len(all_labels) / float(np.mean(sizes)))


def average_precision_binary(labels, scores) -> float:
    """sklearn.metrics.average_precision_score for binary labels (the call at reference
    Trainner/train_label_attention.py:122), restated: thresholds at the distinct score values taken
    in decreasing order, precision P_n = tp/(tp+fp) and recall R_n = tp/|positives| at each, AP =
    sum_n (R_n - R_{n-1}) * P_n; 0.0 when there is no positive."""
    labels = np.asarray(labels).astype(np.int64)
    scores = np.asarray(scores)
    npos = int(labels.sum())
    if npos == 0:
        return 0.0
    order = np.argsort(-scores, kind="mergesort")
    s, y = scores[order], labels[order]
    last_of_threshold = np.r_[np.nonzero(np.diff(s))[0], len(s) - 1]
    tps = np.cumsum(y)[last_of_threshold].astype(np.float64)
    fps = (1 + last_of_threshold - tps).astype(np.float64)
    precision = tps / (tps + fps)
    recall = tps / npos
    return float(np.sum(np.diff(np.r_[0.0, recall]) * precision))


def label_ranking_eval(embs: np.ndarray, label_vals: np.ndarray, topk=(1, 5, 10)):
    """The retrieval metrics of evaluate_label_attention (reference
    Trainner/train_label_attention.py:106-125) given the record embeddings: all-pairs cosine, per
    record the relevance flags (labels share a positive, self = 0) in descending-similarity order,
    ``recall@k`` = their mean over the first k, AP over the full ranking; means over the records.
    Returns (results dict, per-record table (n, 1 + len(topk)) = [AP, recall@k...])."""
    all_embs = np.asarray(embs)
    norms = np.linalg.norm(all_embs, axis=1, keepdims=True)
    sims = all_embs @ all_embs.T / (norms @ norms.T)                      # :108-109
    vals = np.asarray(label_vals).astype(np.int64)
    n = all_embs.shape[0]
    table = np.zeros((n, 1 + len(topk)), dtype=np.float64)
    for i in range(n):
        labels = ((vals & vals[i]).sum(axis=1) > 0).astype(int)         # :114
        labels[i] = 0                                                     # :115
        idx = np.argsort(-sims[i])                                        # :118
        sorted_labels = labels[idx]
        table[i, 0] = average_precision_binary(labels, sims[i])           # :122
        for t, k in enumerate(topk):
            table[i, 1 + t] = sorted_labels[:k].mean()                    # :121
    results = {f"recall@{k}": float(np.mean(table[:, 1 + t])) for t, k in enumerate(topk)}
    results["mAP"] = float(np.mean(table[:, 0]))
    return results, table
