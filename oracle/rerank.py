"""Oracle: label/KG rerank (TEST INFRASTRUCTURE; never imported by the product).

Pure numpy/pandas restatement of ``Reranker`` (reference ``Retrieval/reranker.py``):
scores K candidates as ``alpha*minmax(cos) + beta*minmax(label Jaccard) +
gamma*minmax(KG cos)`` and sorts descending.  Each function cites the lines it follows.
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Set, Tuple

import numpy as np
import pandas as pd


def safe_cos(a, b) -> float:
    """reference ``reranker.py:135-142``: 0.0 if an operand is None or has zero norm,
    else ``dot/(||a||*||b||)`` in the arrays' dtype, returned as a Python float."""
    if a is None or b is None:
        return 0.0
    na = np.linalg.norm(a)
    nb = np.linalg.norm(b)
    if na == 0 or nb == 0:
        return 0.0
    return float(np.dot(a, b) / (na * nb))


def jaccard_sets(a: Set[str], b: Set[str]) -> float:
    """reference ``reranker.py:145-149``: |a&b|/|a|b|, 0.0 when both are empty."""
    if not a and not b:
        return 0.0
    inter = len(a & b)
    uni = len(a | b)
    return 0.0 if uni == 0 else inter / uni


def minmax_scale_list(x_list: Sequence[float]) -> List[float]:
    """reference ``reranker.py:152-159``: fp64 (x-lo)/(hi-lo); all zeros if hi==lo;
    NaN-aware min/max."""
    arr = np.array(x_list, dtype=float)
    if arr.size == 0:
        return arr.tolist()
    lo = float(np.nanmin(arr))
    hi = float(np.nanmax(arr))
    if hi - lo == 0:
        return [0.0] * len(arr)
    return ((arr - lo) / (hi - lo)).tolist()


def load_kg(kg_dir: Path) -> Dict[str, object]:
    """reference ``reranker.py:88-129``: ``node2id.json`` + the best / latest
    ``node_embeddings*.npy``; rows are divided by ``(||row|| + 1e-12)`` (``:120``)."""
    kg_dir = Path(kg_dir)
    node2id_path = kg_dir / "node2id.json"
    if not node2id_path.exists():
        raise FileNotFoundError(f"KG node2id.json not found at {node2id_path}")
    with open(node2id_path, "r", encoding="utf8") as f:
        node2id = json.load(f)
    best = sorted(kg_dir.glob("node_embeddings_best.npy"))
    if best:
        node_file = best[-1]
    else:
        files = sorted(kg_dir.glob("node_embeddings_epoch*.npy"))
        if not files:
            files = sorted(kg_dir.glob("node_embeddings*.npy"))
        if not files:
            raise FileNotFoundError("No .npy embeddings found in KG dir")
        node_file = files[-1]
    node_emb = np.load(node_file)
    node_emb = node_emb / (np.linalg.norm(node_emb, axis=1, keepdims=True) + 1e-12)
    return {"node2id": node2id, "node_emb": node_emb}


class OracleReranker:
    """Restatement of ``Reranker`` (reference ``reranker.py:18-333``).

    ``attn`` is an optional callable ``(n_labels, d) fp32 -> (d,)`` standing for the
    ``LabelAttention`` pooling (reference ``reranker.py:213-218``,
    ``KnowledgeGraph/label_attention.py:19-27``); ``None`` = mean pooling (``:220``),
    which is what the reference does when no checkpoint exists.
    ``record_kg_ids`` mirrors ``_try_precompute_record_kg`` (``:222-238``): when given,
    those records' KG vectors are precomputed into an fp64 table.
    """

    def __init__(self, kg_dir, labels_csv, alpha=0.6, beta=0.25, gamma=0.15, attn=None,
                 record_kg_ids: Optional[Sequence[str]] = None):
        self.alpha, self.beta, self.gamma = alpha, beta, gamma
        self.kg = load_kg(Path(kg_dir))
        self.labels_df = pd.read_csv(labels_csv, index_col="id")          # :47
        self.labels_df.index = self.labels_df.index.astype(str)           # :48
        self.attn = attn
        self.record_kg_vectors = None
        self.record_kg_id2idx = None
        if record_kg_ids is not None:
            ids = list(record_kg_ids)
            vecs = np.zeros((len(ids), self.kg["node_emb"].shape[1]), dtype=float)   # :231
            for i, rid in enumerate(ids):
                vecs[i] = self.get_record_kg_vec(str(rid))
            self.record_kg_vectors = vecs
            self.record_kg_id2idx = {str(r): i for i, r in enumerate(ids)}

    def get_record_label_set(self, rec_id) -> Set[str]:
        """reference ``reranker.py:161-179``."""
        if str(rec_id) not in self.labels_df.index:
            return set()
        row = self.labels_df.loc[str(rec_id)]
        labels = []
        for c, v in row.items():
            try:
                if int(v) == 1:
                    labels.append(c)
            except (ValueError, TypeError):
                continue
        return set(labels)

    def get_record_kg_vec(self, rec_id) -> np.ndarray:
        """reference ``reranker.py:181-220``."""
        node2id = self.kg["node2id"]
        node_emb = self.kg["node_emb"]
        k1 = f"report:{rec_id}"
        if k1 in node2id:
            return node_emb[node2id[k1]]
        if str(rec_id) in node2id:
            return node_emb[node2id[str(rec_id)]]
        labels = self.get_record_label_set(rec_id)
        if not labels:
            return np.zeros(node_emb.shape[1], dtype=float)
        vecs = []
        for lab in labels:
            for ck in (f"label:{lab}", lab, lab.lower(), lab.replace(" ", "_")):
                if ck in node2id:
                    vecs.append(node_emb[node2id[ck]])
                    break
        if not vecs:
            return np.zeros(node_emb.shape[1], dtype=float)
        label_embs = np.stack(vecs, axis=0)
        if self.attn is not None:
            return np.asarray(self.attn(label_embs.astype(np.float32)))
        return label_embs.mean(axis=0)

    def rerank(self, query_id, candidate_ids, candidate_embs=None, candidate_emb_lookup=None,
               topk=None, query_emb=None) -> List[Tuple[str, float, float, float, float]]:
        """reference ``reranker.py:240-333``."""
        N = len(candidate_ids)
        if candidate_embs is None:
            if candidate_emb_lookup is not None:
                zero = np.zeros(next(iter(candidate_emb_lookup.values())).shape, dtype=float)
                candidate_embs = np.vstack(
                    [candidate_emb_lookup.get(str(c), zero) for c in candidate_ids])
            else:
                raise ValueError("Please provide candidate_embs or candidate_emb_lookup.")
        if candidate_embs.shape[0] != N:
            raise ValueError("candidate_embs rows must match candidate_ids length")
        q_emb = None
        if candidate_emb_lookup is not None and str(query_id) in candidate_emb_lookup:
            q_emb = candidate_emb_lookup[str(query_id)]
        elif query_emb is not None:
            q_emb = query_emb
        else:
            for i, cid in enumerate(candidate_ids):
                if str(cid) == str(query_id):
                    q_emb = candidate_embs[i]
                    break
        if q_emb is None:
            raise ValueError("Query embedding not found.")
        emb_scores = [safe_cos(q_emb, candidate_embs[i]) for i in range(N)]
        q_labels = self.get_record_label_set(query_id)
        lab_scores = [jaccard_sets(q_labels, self.get_record_label_set(c)) for c in candidate_ids]

        def kg_of(rid):
            if self.record_kg_vectors is not None:
                idx = self.record_kg_id2idx.get(str(rid))
                if idx is not None:
                    return self.record_kg_vectors[idx]
            return self.get_record_kg_vec(str(rid))

        q_kg = kg_of(query_id)
        kg_scores = [safe_cos(q_kg, kg_of(c)) for c in candidate_ids]
        emb_n = np.array(minmax_scale_list(emb_scores))
        lab_n = np.array(minmax_scale_list(lab_scores))
        kg_n = np.array(minmax_scale_list(kg_scores))
        final = self.alpha * emb_n + self.beta * lab_n + self.gamma * kg_n
        # deterministic tie rule (final desc, candidate position asc); the reference's
        # ``np.argsort(final)[::-1]`` (:327) leaves tie order unspecified
        ranked = np.lexsort((np.arange(N), -final))
        if topk:
            ranked = ranked[:topk]
        return [(candidate_ids[i], float(final[i]), float(emb_n[i]), float(lab_n[i]),
                 float(kg_n[i])) for i in ranked]


def rerank_from_arrays(q_emb, cand_embs, q_mask, cand_masks, q_kg, cand_kg, alpha=0.6, beta=0.25, gamma=0.15,
                       topk=None, emb_scores=None):
    """The scoring half of ``Reranker.rerank`` (reference ``reranker.py:298-329``) on pre-resolved arrays --
    what the table-driven device path (``Reranker.from_tables``: integer label masks, KG rows) computes:
    ``emb_scores[i] = safe_cos(q_emb, cand_embs[i])`` (``:298``; or the supplied ``emb_scores``),
    Jaccard of the label sets encoded by the masks' bits (``:301-304``), ``safe_cos`` of the KG rows
    (``:307-319``), min-max each (``:322-324``), ``final = alpha*e + beta*l + gamma*g`` (``:325``), order by
    final descending (``:327``; ties by candidate position ascending).  Returns ``[(position, final, emb_n,
    lab_n, kg_n)]``."""
    n = len(cand_masks)

    def bits(m):
        m = int(m)
        return {i for i in range(m.bit_length()) if (m >> i) & 1}

    if emb_scores is None:
        emb_scores = [safe_cos(q_emb, cand_embs[i]) for i in range(n)]
    else:
        emb_scores = [float(x) for x in emb_scores]
    ql = bits(q_mask)
    lab_scores = [jaccard_sets(ql, bits(cand_masks[i])) for i in range(n)]
    kg_scores = [safe_cos(q_kg, cand_kg[i]) for i in range(n)]
    emb_n = np.array(minmax_scale_list(emb_scores))
    lab_n = np.array(minmax_scale_list(lab_scores))
    kg_n = np.array(minmax_scale_list(kg_scores))
    final = alpha * emb_n + beta * lab_n + gamma * kg_n
    ranked = np.lexsort((np.arange(n), -final))
    if topk:
        ranked = ranked[:topk]
    return [(int(i), float(final[i]), float(emb_n[i]), float(lab_n[i]), float(kg_n[i])) for i in ranked]


def reranked_lists_match(got_ids, got_final, cand_ids, want, rtol=1e-5, atol=2e-6):
    """Compare one query's device result (ids + combined scores, best first) with ``want`` from
    :func:`rerank_from_arrays` (positions into ``cand_ids``): finals within tolerance position by
    position; ids equal except for swaps between (near-)equal finals.  Returns ``(ok, why)``."""
    if len(got_ids) != len(want):
        return False, f"length {len(got_ids)} vs {len(want)}"
    for r, (gid, gf, w) in enumerate(zip(got_ids, got_final, want)):
        if abs(float(gf) - w[1]) > atol + rtol * abs(w[1]):
            return False, f"rank {r}: final {float(gf)!r} vs {w[1]!r}"
    want_ids = [int(cand_ids[w[0]]) for w in want]
    if sorted(int(x) for x in got_ids) != sorted(want_ids):
        # a swap at the cut (topk < candidates) between near-equal finals is allowed
        extra = set(int(x) for x in got_ids) ^ set(want_ids)
        fin_of = {int(cand_ids[w[0]]): w[1] for w in want}
        cut = want[-1][1]
        for e in extra:
            if e in fin_of and abs(fin_of[e] - cut) > atol + rtol * abs(cut):
                return False, f"id {e} is not a near-tie at the cut"
    for r, (gid, w) in enumerate(zip(got_ids, want)):
        if int(gid) != int(cand_ids[w[0]]):
            # the id found at rank r must have (near-)the same final as the expected one
            if abs(float(got_final[r]) - w[1]) > atol + rtol * abs(w[1]):
                return False, f"rank {r}: id {int(gid)} vs {int(cand_ids[w[0]])} (not a tie)"
    return True, "ok"
