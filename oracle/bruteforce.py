"""Oracle: brute-force exact search on a GPU with plain torch (TEST INFRASTRUCTURE).

The numpy restatement in ``oracle/search.py`` cannot finish the BASELINE.json full sizes (10M x 512
rows) in seconds, so the full-size parity tests and ``bench.py``'s ``parity_check`` use this torch
restatement of the SAME form -- ``cosine_similarity(Q, G)`` = row-normalise both operands, one matmul
(reference ``Evaluate/retrieval_overlap.py:85``; sklearn ``normalize`` + ``safe_sparse_dot``), then
``np.argsort(row)[::-1][:k]`` (``:90``) -- chunked over gallery rows, fp32 (TF32 off) or fp64.
It is pinned on ``oracle.search.exact_topk`` (``tests/test_oracle_golden.py``) and is only ever the
checker: nothing in the product package imports it, and no timed region runs it.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def bruteforce_topk(gallery, queries, k: int, row_offset: int = 0, chunk: int = 1 << 19, dtype=None,
                    margin: int = 16) -> Tuple[np.ndarray, np.ndarray]:
    """``gallery`` (N, D) and ``queries`` (B, D) torch tensors on one device (any float dtype; the values
    are used as they are, upcast to ``dtype``).  Returns numpy ``(rows int64 (B, k'), scores (B, k'))``,
    ``k' = min(k, N)``, ordered by score descending then row ascending (the CUDA path's tie rule)."""
    import torch
    dtype = dtype or torch.float32
    dev = gallery.device
    n = int(gallery.shape[0])
    b = int(queries.shape[0])
    kk = min(int(k), n)
    old_tf32 = torch.backends.cuda.matmul.allow_tf32
    old_prec = torch.get_float32_matmul_precision()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.set_float32_matmul_precision("highest")
    try:
        q = queries.to(dtype)
        qn = torch.sqrt((q * q).sum(dim=1, keepdim=True))        # row_norms: sqrt(einsum('ij,ij->i'))
        qn = torch.where(qn == 0, torch.ones_like(qn), qn)        # _handle_zeros_in_scale
        q = q / qn
        cand_s, cand_r = [], []
        for lo in range(0, n, chunk):
            hi = min(n, lo + chunk)
            g = gallery[lo:hi].to(dtype)
            gn = torch.sqrt((g * g).sum(dim=1, keepdim=True))
            gn = torch.where(gn == 0, torch.ones_like(gn), gn)
            sim = q @ (g / gn).T                                   # (B, rows)
            take = min(hi - lo, kk + margin)                       # margin: ties at a chunk's cut stay visible
            s, i = torch.topk(sim, take, dim=1)
            cand_s.append(s)
            cand_r.append(i + lo)
            del g, sim
        s = torch.cat(cand_s, dim=1).cpu().numpy()
        r = torch.cat(cand_r, dim=1).cpu().numpy().astype(np.int64)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old_tf32
        torch.set_float32_matmul_precision(old_prec)
    rows = np.empty((b, kk), dtype=np.int64)
    scores = np.empty((b, kk), dtype=s.dtype)
    for i in range(b):
        order = np.lexsort((r[i], -s[i].astype(np.float64)))[:kk]
        rows[i] = r[i, order] + row_offset
        scores[i] = s[i, order]
    return rows, scores


def check_topk(got_rows, got_scores, gallery, queries, k: int, row_offset: int = 0, rtol: float = 2e-5,
               atol: float = 1e-6, eps: float = 1e-5):
    """The full-size parity check of one result block: ``got_rows`` / ``got_scores`` (B, k) (torch or
    numpy, GLOBAL rows) against the fp32 brute force over the same stored values
    (``oracle.search.topk_matches``: ids position by position outside near-ties, scores within
    ``rtol``), plus exact recall@k against the fp64 ranking with the epsilon rule at the k boundary
    (a missing id must score within ``eps`` of the fp64 k-th score).  Returns ``(ok, detail)``."""
    from .search import topk_matches
    import torch
    gr = got_rows.cpu().numpy() if hasattr(got_rows, "cpu") else np.asarray(got_rows)
    gs = got_scores.cpu().numpy() if hasattr(got_scores, "cpu") else np.asarray(got_scores)
    want_r, want_s = bruteforce_topk(gallery, queries, k, row_offset=row_offset)
    r64, s64 = bruteforce_topk(gallery, queries, k, row_offset=row_offset, chunk=1 << 18, dtype=torch.float64)
    kk = want_r.shape[1]
    exact = 0
    for i in range(gr.shape[0]):
        ok, why = topk_matches(gr[i, :kk], gs[i, :kk], want_r[i], want_s[i], rtol=rtol, atol=atol)
        if not ok:
            return False, f"query {i}: {why}"
        exact += int(np.array_equal(gr[i, :kk], want_r[i]))
        have = set(gr[i, :kk].tolist())
        for j in np.nonzero(~np.isin(r64[i], gr[i, :kk]))[0]:
            if s64[i, j] - s64[i, -1] > eps:
                return False, f"query {i}: fp64 rank {int(j)} (row {int(r64[i, j])}) missing and not a boundary near-tie"
        if len(have) != kk:
            return False, f"query {i}: duplicate rows in the result"
    return True, f"{gr.shape[0]} queries, {exact} identical id lists, rest differ only inside near-ties"
