"""Oracle: exact cosine search + ranking, and the DLS walk (TEST INFRASTRUCTURE).

numpy restatement of the arithmetic the CUDA search kernels must reproduce.
Nothing in the product package imports this file.

Third-party arithmetic restated here (not under /root/reference):
``sklearn.metrics.pairwise.cosine_similarity`` (reference pins scikit-learn 1.7.0,
requirements.txt; 1.9.0 installed here, ``sklearn/metrics/pairwise.py``) =
``normalize(X)``, ``normalize(Y)`` -- row L2 norms ``sqrt(einsum('ij,ij->i'))`` in the
input dtype, zero norms replaced by 1 -- followed by one ``Xn @ Yn.T`` (BLAS sgemm for
fp32).  Call sites in the reference: ``Retrieval/retrieval.py:5,128`` and
``Evaluate/retrieval_overlap.py:15,85``.
"""
from __future__ import annotations

import heapq
from typing import List, Optional, Sequence, Tuple

import numpy as np


def l2_normalize_rows(x: np.ndarray) -> np.ndarray:
    """sklearn ``normalize(X, norm='l2')``: ``row_norms`` + ``_handle_zeros_in_scale``.

    Norms are computed in the array's own dtype (fp32 in -> fp32 norms); exact zeros
    are replaced by 1 so a zero row stays zero (no NaN).
    """
    x = np.array(x, copy=True)
    norms = np.sqrt(np.einsum("ij,ij->i", x, x))
    norms[norms == 0.0] = 1.0
    x /= norms[:, None]
    return x


def cosine_similarity(q: np.ndarray, g: Optional[np.ndarray] = None) -> np.ndarray:
    """(Q,D),(N,D) -> (Q,N) cosine matrix.

    Follows ``cosine_similarity(query_embs, gallery_embs)``
    (reference ``Evaluate/retrieval_overlap.py:85``) and ``cosine_similarity(self.embs)``
    (``Retrieval/retrieval.py:128``).
    """
    qn = l2_normalize_rows(np.asarray(q))
    gn = qn if g is None else l2_normalize_rows(np.asarray(g))
    return qn @ gn.T


def rank_descending(sim_row: np.ndarray) -> np.ndarray:
    """``np.argsort(sim[i])[::-1]`` (reference ``retrieval_overlap.py:90``,
    ``retrieval.py:134``).  Tie order is whatever numpy's unstable sort yields."""
    return np.argsort(sim_row)[::-1]


def exact_topk(q: np.ndarray, g: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """Exact search oracle (O-search).

    Returns ``(rows int64 (Q,k'), scores (Q,k'))`` with ``k' = min(k, N)``, best first,
    using the DETERMINISTIC tie rule of the CUDA path (score descending, then row
    ascending) so that results can be compared bit-for-bit outside exact-score ties;
    inside a tie group the reference's own order is unspecified (numpy unstable sort).
    """
    sim = cosine_similarity(q, g)
    n = sim.shape[1]
    kk = min(int(k), n)
    rows = np.empty((sim.shape[0], kk), dtype=np.int64)
    scores = np.empty((sim.shape[0], kk), dtype=sim.dtype)
    for i in range(sim.shape[0]):
        # every row scoring at least the kk-th largest value (all boundary ties included), then
        # lexsort (last key is primary): primary = -score (desc), secondary = row (asc).  Identical to
        # sorting the whole row, without the O(N log N) per query.
        if kk < n:
            thr = np.partition(sim[i], n - kk)[n - kk]
            idx = np.nonzero(sim[i] >= thr)[0]
        else:
            idx = np.arange(n)
        order = idx[np.lexsort((idx, -sim[i, idx].astype(np.float64)))[:kk]]
        rows[i] = order
        scores[i] = sim[i, order]
    return rows, scores


def exact_topk_f64(q: np.ndarray, g: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """fp64 version of :func:`exact_topk` (for the bf16 recall / epsilon check:
    inputs are the bf16-rounded values upcast, SURVEY.md section 7 "Hard parts")."""
    return exact_topk(np.asarray(q, dtype=np.float64), np.asarray(g, dtype=np.float64), k)


def to_bf16_round(x: np.ndarray) -> np.ndarray:
    """Round fp32 -> bf16 (round-to-nearest-even) and return the values as fp32.

    The bf16 configs use these rounded values AS the dataset.
    NaN/Inf are passed through.
    """
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    rounded = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    out = rounded.astype(np.uint32).view(np.float32).reshape(x.shape)
    bad = ~np.isfinite(x)
    if bad.any():
        out = out.copy()
        out[bad] = x[bad]
    return out


def to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """fp32 -> uint16 bf16 bit patterns (RNE), same rounding as :func:`to_bf16_round`."""
    return (to_bf16_round(x).view(np.uint32) >> 16).astype(np.uint16)


# --------------------------------------------------------------------------------------
# Link graph + DenseLinkSearch walk (reference Retrieval/retrieval.py:121-271)
# --------------------------------------------------------------------------------------

def build_link_graph(embs: np.ndarray, threshold: float, max_links: int) -> List[List[int]]:
    """``DLSRetrievalEngine._build_link_graph`` (reference ``retrieval.py:121-138``):
    all-pairs cosine, diagonal forced to -1, per-row descending order, keep neighbours
    with ``sim >= threshold``, truncate to ``max_links``.

    Deterministic tie rule (score desc, row asc) instead of numpy's unspecified one.
    """
    sim = cosine_similarity(embs)
    np.fill_diagonal(sim, -1)
    n = sim.shape[0]
    ar = np.arange(n)
    graph = []
    for i in range(n):
        order = np.lexsort((ar, -sim[i].astype(np.float64)))
        # descending order => the entries >= threshold are a prefix of ``order``
        graph.append([int(j) for j in order[:max_links] if sim[i, j] >= threshold])
    return graph


def dls_score(emb: np.ndarray, q: np.ndarray, q_norm: float) -> float:
    """Per-candidate cosine of the walk (reference ``retrieval.py:203-206,224-225``):
    ``emb @ q / (||emb|| * (||q|| + 1e-6) + 1e-12)``, fp32 arrays -> Python float."""
    return float(emb @ q / (np.linalg.norm(emb) * q_norm + 1e-12))


def dls_retrieve(
    embs: np.ndarray,
    link_graph: Sequence[Sequence[int]],
    query_emb: np.ndarray,
    K: int = 5,
    seed_size: int = 5,
    max_steps: int = 100,
    candidate_multiplier: int = 10,
    seed: Optional[int] = None,
) -> Tuple[List[int], List[float]]:
    """The greedy heap walk of ``DLSRetrievalEngine.retrieve`` up to the top-K cut
    (reference ``retrieval.py:178-244``), returning ROW indices instead of ids.

    Quirk kept on purpose: the node popped at every step is never pushed back, so it
    can never be returned (``retrieval.py:215`` vs ``:240``; SURVEY.md section 0
    finding 2).  ``seed`` must be given for reproducibility (the reference's
    ``query_id`` seeding uses Python's salted ``hash``, ``retrieval.py:193``).
    """
    q = np.asarray(query_emb).astype("float32").reshape(-1)
    N = embs.shape[0]
    if len(link_graph) != N:
        raise RuntimeError(
            f"Link graph size {len(link_graph)} != embeddings {N}. Rebuild or delete your pickle."
        )
    np.random.seed(seed)
    seeds = np.random.choice(N, size=min(seed_size, N), replace=False).tolist()
    visited = set(seeds)
    heap: list = []
    q_norm = np.linalg.norm(q) + 1e-6
    for idx in seeds:
        heapq.heappush(heap, (-dls_score(embs[idx], q, q_norm), idx))
    R = max(candidate_multiplier * K, seed_size)
    steps = 0
    while steps < max_steps and heap:
        _neg, best_idx = heapq.heappop(heap)
        improved = False
        for nbr in link_graph[best_idx]:
            if nbr < 0 or nbr >= N or nbr in visited:
                continue
            visited.add(nbr)
            heapq.heappush(heap, (-dls_score(embs[nbr], q, q_norm), nbr))
            improved = True
        if len(heap) > R:
            heap = heapq.nsmallest(R, heap)
            heapq.heapify(heap)
        if not improved:
            break
        steps += 1
    topk = heapq.nsmallest(K, heap)
    topk = sorted([(-neg, idx) for neg, idx in topk], reverse=True)
    return [idx for _, idx in topk], [sim for sim, _ in topk]


# --------------------------------------------------------------------------------------
# comparison helpers (tie-aware)
# --------------------------------------------------------------------------------------

def topk_matches(
    rows_a: np.ndarray,
    scores_a: np.ndarray,
    rows_b: np.ndarray,
    scores_b: np.ndarray,
    rtol: float = 1e-5,
    atol: float = 0.0,
) -> Tuple[bool, str]:
    """Compare two best-first top-K results for one query.

    ids must agree position by position except inside groups of (near-)equal score,
    where the two sides may order the same set differently, and at the K boundary,
    where a near-tie may swap an element in or out.  Scores must agree within
    ``rtol``/``atol`` position by position.
    """
    rows_a = np.asarray(rows_a); rows_b = np.asarray(rows_b)
    sa = np.asarray(scores_a, dtype=np.float64); sb = np.asarray(scores_b, dtype=np.float64)
    if rows_a.shape != rows_b.shape:
        return False, f"shape {rows_a.shape} vs {rows_b.shape}"
    tol = atol + rtol * np.maximum(np.abs(sa), np.abs(sb))
    if not np.all(np.abs(sa - sb) <= tol):
        i = int(np.argmax(np.abs(sa - sb) - tol))
        return False, f"score mismatch at rank {i}: {sa[i]!r} vs {sb[i]!r}"
    k = len(rows_a)
    bad = np.nonzero(rows_a != rows_b)[0]
    if bad.size == 0:
        return True, "exact"
    # every mismatching position must sit inside a near-tie group
    kth = min(sa[-1], sb[-1]) if k else 0.0
    kth_tol = atol + rtol * abs(kth)
    for i in bad:
        tol_i = 2 * (atol + rtol * abs(sa[i]))
        group = np.nonzero(np.abs(sa - sa[i]) <= tol_i)[0]
        set_a = set(rows_a[group].tolist()); set_b = set(rows_b[group].tolist())
        if set_a == set_b:
            continue
        # boundary swap: the differing elements must all score within tol of the K-th
        if abs(sa[i] - kth) <= 2 * kth_tol + tol_i:
            continue
        return False, f"id mismatch at rank {i}: {rows_a[i]} vs {rows_b[i]} (not a tie)"
    return True, "ties"
