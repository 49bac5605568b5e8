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
