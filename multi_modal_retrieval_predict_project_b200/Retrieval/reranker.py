"""B200-native mirror of the reference's ``Retrieval/reranker.py``.

Same constructor and ``rerank`` signature / return type / error behaviour as the reference
``Reranker`` (``reranker.py:29-37,240-248``), but the scoring runs in libmmr_b200.so:
label sets become 64-bit masks (one popcount instead of pandas ``.loc`` + Python loops),
record KG vectors become rows of a device table, and cosine / min-max / combine / ordering are
CUDA kernels (csrc/rerank.cu).  Table construction at ``__init__`` is host-side file parsing, done
once, exactly following ``_load_kg`` (:88-129), ``get_record_label_set`` (:161-179) and
``get_record_kg_vec`` (:181-220).
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import pandas as pd

from .. import _lib


class Reranker:
    """``alpha * minmax(cos) + beta * minmax(label Jaccard) + gamma * minmax(KG cos)``, descending.

    Extra (keyword-only) arguments over the reference: ``device``, ``label_attention`` (path to a
    ``label_attention_model.pt`` checkpoint; the reference looks for it under its own BASE_DIR,
    ``reranker.py:62-63``) and ``la_hidden_dim`` (``configs/config.yaml:47``).  ``preload_record_kg`` is accepted
    for signature compatibility: the record KG table is always built up front here (it IS the device table; fp32,
    the reference's optional precomputed table is fp64 -- the cosines are evaluated in fp32 either way).
    """

    def __init__(self, kg_dir: Optional[Path] = None, labels_csv: Optional[Path] = None, alpha: float = 0.6,
                 beta: float = 0.25, gamma: float = 0.15, preload_record_kg: bool = True, *, device=None,
                 label_attention: Optional[Path] = None, la_hidden_dim: int = 256):
        # the reference's defaults are relative to its checkout (reranker.py:11-15: BASE_DIR / "knowledge_graph",
        # BASE_DIR / "outputs" / "openi_labels_final.csv"); here BASE_DIR = $MMR_B200_BASE_DIR or the working
        # directory (the reference's scripts are run from the repository root).  Missing files raise
        # FileNotFoundError exactly like the reference (:103, :116, pandas for the CSV).
        import os
        base = Path(os.environ.get("MMR_B200_BASE_DIR", os.getcwd()))
        self.kg_dir = Path(kg_dir) if kg_dir else base / "knowledge_graph"
        self.labels_csv = Path(labels_csv) if labels_csv else base / "outputs" / "openi_labels_final.csv"
        self.alpha, self.beta, self.gamma = alpha, beta, gamma
        self.kg = self._load_kg(self.kg_dir)
        self.labels_df = pd.read_csv(self.labels_csv, index_col="id")
        self.labels_df.index = self.labels_df.index.astype(str)
        self.attn_params = self._load_label_attention(label_attention, la_hidden_dim)
        self.device = _lib.require_cuda(device)
        self._lib = _lib.load()
        self._tables = None
        self._build_tables()

    # ------------------------------------------------------------------ host-side table building
    @staticmethod
    def _load_kg(kg_dir: Path) -> Dict[str, Any]:
        node2id_path = kg_dir / "node2id.json"
        if not node2id_path.exists():
            raise FileNotFoundError(f"KG node2id.json not found at {node2id_path}")
        with open(node2id_path, "r", encoding="utf8") as f:
            node2id = json.load(f)
        best = sorted(kg_dir.glob("node_embeddings_best.npy"))
        if best:
            node_file = best[-1]
        else:
            files = sorted(kg_dir.glob("node_embeddings_epoch*.npy")) or sorted(kg_dir.glob("node_embeddings*.npy"))
            if not files:
                raise FileNotFoundError("No .npy embeddings found in KG dir")
            node_file = files[-1]
        node_emb = np.load(node_file)
        node_emb = node_emb / (np.linalg.norm(node_emb, axis=1, keepdims=True) + 1e-12)
        print(f"[Reranker] loaded KG embeddings from {node_file}")
        return {"node2id": node2id, "node_emb": node_emb}

    @staticmethod
    def _load_label_attention(path, hidden):
        """LabelAttention weights (reference KnowledgeGraph/label_attention.py:11-17) or None
        (mean pooling, reranker.py:84-86,219-220)."""
        if path is None or not Path(path).exists():
            print("[INFO] No LabelAttention model found – will fall back to mean pooling")
            return None
        import torch
        ckpt = torch.load(path, map_location="cpu")
        sd = ckpt.get("model_state", ckpt) if isinstance(ckpt, dict) else ckpt
        return {k: v.detach().float().numpy() for k, v in sd.items()}

    def _pool(self, label_embs: np.ndarray) -> np.ndarray:
        if self.attn_params is None:
            return label_embs.mean(axis=0)
        p = self.attn_params  # Linear -> Tanh -> Linear -> softmax -> weighted sum (label_attention.py:19-27)
        x = label_embs.astype(np.float32)
        h = np.tanh(x @ p["attn.0.weight"].T + p["attn.0.bias"])
        s = (h @ p["attn.2.weight"].T + p["attn.2.bias"]).reshape(-1)
        w = np.exp(s - s.max())
        w = (w / w.sum()).astype(np.float32)
        return (w[None, :] @ x).reshape(-1)

    def _label_bits(self) -> Tuple[np.ndarray, List[str]]:
        """(n_rec, words) uint64 masks: bit c set iff ``int(value) == 1`` in column c
        (reranker.py:173-178; non-numeric cells are skipped; duplicated ids yield an empty set
        because ``.loc`` then returns a frame and ``int(Series)`` raises)."""
        df = self.labels_df
        cols = list(df.columns)
        n = len(df)
        words = max(1, (len(cols) + 63) // 64)
        masks = np.zeros((n, words), dtype=np.uint64)
        for c, name in enumerate(cols):
            col = df[name]
            if pd.api.types.is_bool_dtype(col):
                on = col.to_numpy(dtype=bool)
            elif pd.api.types.is_numeric_dtype(col):
                v = col.to_numpy(dtype=np.float64)
                with np.errstate(invalid="ignore"):
                    on = np.trunc(v) == 1.0          # int(v) truncates; NaN -> ValueError -> skipped
            else:
                on = np.zeros(n, dtype=bool)
                for i, v in enumerate(col.to_numpy(dtype=object)):
                    try:
                        on[i] = int(v) == 1
                    except (ValueError, TypeError):
                        pass
            masks[on, c // 64] |= np.uint64(1) << np.uint64(c % 64)
        dup = df.index.duplicated(keep=False)
        masks[dup] = 0
        return masks, cols

    def _build_tables(self):
        node2id, node_emb = self.kg["node2id"], self.kg["node_emb"]
        rec_ids = [str(i) for i in self.labels_df.index]
        masks, cols = self._label_bits()
        rec2row: Dict[str, int] = {}
        for i, rid in enumerate(rec_ids):
            rec2row.setdefault(rid, i)
        # records that only exist in the KG (report:<id> nodes without a CSV row)
        extra = [k[len("report:"):] for k in node2id if k.startswith("report:") and k[len("report:"):] not in rec2row]
        for rid in extra:
            rec2row[rid] = len(rec_ids)
            rec_ids.append(rid)
        if extra:
            masks = np.vstack([masks, np.zeros((len(extra), masks.shape[1]), dtype=np.uint64)])
        d_kg = node_emb.shape[1]
        kg = np.zeros((len(rec_ids), d_kg), dtype=np.float32)
        node_f32 = node_emb.astype(np.float32, copy=False)
        label_node = []
        for lab in cols:  # candidate node keys of a label (reranker.py:203-207)
            hit = None
            for ck in (f"label:{lab}", lab, str(lab).lower(), str(lab).replace(" ", "_")):
                if ck in node2id:
                    hit = node2id[ck]
                    break
            label_node.append(hit)
        for i, rid in enumerate(rec_ids):
            k1 = f"report:{rid}"
            if k1 in node2id:
                kg[i] = node_f32[node2id[k1]]
            elif rid in node2id:
                kg[i] = node_f32[node2id[rid]]
            else:
                on = [c for c in range(len(cols)) if (int(masks[i, c // 64]) >> (c % 64)) & 1]
                vecs = [node_emb[label_node[c]] for c in on if label_node[c] is not None]
                if vecs:
                    kg[i] = self._pool(np.stack(vecs, axis=0))
        self.rec_ids, self.rec2row, self.label_columns = rec_ids, rec2row, cols
        self._d_kg = int(d_kg)
        self._masks_host, self._kg_host = np.ascontiguousarray(masks), np.ascontiguousarray(kg)
        import torch
        h = _lib.C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.mmr_rerank_tables_create(
                _lib.C.byref(h), _lib.ptr(self._masks_host), masks.shape[1], _lib.ptr(self._kg_host), d_kg,
                len(rec_ids), self.device, _lib.current_stream(self.device)))
        self._tables = h

    @classmethod
    def from_tables(cls, label_masks, kg_vecs, alpha: float = 0.6, beta: float = 0.25, gamma: float = 0.15,
                    device=None, rec_ids: Optional[Sequence[str]] = None) -> "Reranker":
        """Build from in-memory tables (synthetic / pre-processed data): ``label_masks`` (n_rec,) or
        (n_rec, words) uint64 (numpy) / int64 (torch, bit pattern), ``kg_vecs`` (n_rec, d_kg) fp32 --
        rows already L2-normalised the way ``_load_kg`` does.  numpy or CUDA torch tensors."""
        import torch
        self = cls.__new__(cls)
        self.alpha, self.beta, self.gamma = alpha, beta, gamma
        self.device = _lib.require_cuda(device)
        self._lib = _lib.load()
        self.attn_params = None
        n_rec = int(kg_vecs.shape[0])
        words = 1 if label_masks.ndim == 1 else int(label_masks.shape[1])
        if isinstance(label_masks, np.ndarray):
            label_masks = np.ascontiguousarray(label_masks, dtype=np.uint64)
            self._masks_host = label_masks.reshape(n_rec, words)
        else:
            label_masks = label_masks.contiguous()
            self._masks_host = None
        if isinstance(kg_vecs, np.ndarray):
            kg_vecs = np.ascontiguousarray(kg_vecs, dtype=np.float32)
            self._kg_host = kg_vecs
        else:
            kg_vecs = kg_vecs.float().contiguous()
            self._kg_host = None
        self.rec_ids = list(rec_ids) if rec_ids is not None else None
        self.rec2row = {str(r): i for i, r in enumerate(self.rec_ids)} if rec_ids is not None else {}
        self.label_columns = [f"bit{i}" for i in range(64 * words)]
        self._d_kg = int(kg_vecs.shape[1])
        h = _lib.C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.mmr_rerank_tables_create(
                _lib.C.byref(h), _lib.ptr(label_masks), words, _lib.ptr(kg_vecs), int(kg_vecs.shape[1]), n_rec,
                self.device, _lib.current_stream(self.device)))
            torch.cuda.current_stream(self.device).synchronize()
        self._tables = h
        return self

    def close(self):
        t, self._tables = self._tables, None
        if t is not None:
            self._lib.mmr_rerank_tables_destroy(t)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ reference helper surface
    def get_record_label_set(self, rec_id: str):
        row = self.rec2row.get(str(rec_id))
        if row is None:
            return set()
        m = self._masks_host[row]
        return {c for i, c in enumerate(self.label_columns) if (int(m[i // 64]) >> (i % 64)) & 1}

    def get_record_kg_vec(self, rec_id: str) -> np.ndarray:
        row = self.rec2row.get(str(rec_id))
        if row is None:
            return np.zeros(self._kg_host.shape[1], dtype=float)
        return self._kg_host[row]

    def _rec_rows(self, ids: Sequence) -> np.ndarray:
        return np.array([self.rec2row.get(str(i), -1) for i in ids], dtype=np.int64)

    # ------------------------------------------------------------------ rerank
    def rerank(self, query_id: str, candidate_ids: List[str], candidate_embs: Optional[np.ndarray] = None,
               candidate_emb_lookup: Optional[Dict[str, np.ndarray]] = None, topk: Optional[int] = None,
               query_emb: Optional[np.ndarray] = None) -> List[Tuple[str, float, float, float, float]]:
        """Signature and semantics of the reference ``Reranker.rerank`` (``reranker.py:240-333``)."""
        N = len(candidate_ids)
        if candidate_embs is None:
            if candidate_emb_lookup is not None:
                zero = np.zeros(next(iter(candidate_emb_lookup.values())).shape, dtype=float)
                candidate_embs = np.vstack([candidate_emb_lookup.get(str(c), zero) for c in candidate_ids])
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
            raise ValueError(
                "Query embedding not found. Provide candidate_emb_lookup keyed by query_id, "
                "or include the query_id in candidate_ids with matching candidate_embs, "
                "or pass query_emb explicitly.")
        if N == 0:
            return []
        cand = np.ascontiguousarray(candidate_embs, dtype=np.float32).reshape(1, N, -1)
        q = np.ascontiguousarray(np.asarray(q_emb, dtype=np.float32).reshape(1, -1))
        order, sc = self._call(None, q, cand, None, self._rec_rows([query_id]),
                               self._rec_rows(candidate_ids).reshape(1, N), None, topk or 0)
        return [(candidate_ids[int(j)], float(s[0]), float(s[1]), float(s[2]), float(s[3]))
                for j, s in zip(order[0], sc[0]) if j >= 0]

    def rerank_rows(self, engine, query_ids: Sequence, q_embs, cand_rows, cand_ids: Optional[Sequence[Sequence]],
                    topk: Optional[int] = None):
        """Batched rerank of search results: candidate embeddings are gathered on the device from
        ``engine``'s gallery by GLOBAL row id (no host copy of the K x D block).  ``cand_ids`` are
        the candidates' string ids (``None`` = use ``engine.ids``).  Returns, per query, the
        reference's list of ``(id, final, emb_n, lab_n, kg_n)``."""
        cand_rows = np.ascontiguousarray(cand_rows, dtype=np.int64)
        b, k = cand_rows.shape
        if cand_ids is None:
            cand_ids = [[engine.ids[int(r) - engine.row_offset] if r >= 0 else None for r in cand_rows[i]]
                        for i in range(b)]
        counts = np.array([len(c) for c in cand_ids], dtype=np.int32)
        crec = -np.ones((b, k), dtype=np.int64)
        for i in range(b):
            crec[i, : counts[i]] = self._rec_rows(cand_ids[i])
        q = np.ascontiguousarray(q_embs, dtype=np.float32).reshape(b, -1)
        order, sc = self._call(engine, q, None, cand_rows, self._rec_rows(query_ids), crec, counts, topk or 0)
        out = []
        for i in range(b):
            out.append([(cand_ids[i][int(j)], float(s[0]), float(s[1]), float(s[2]), float(s[3]))
                        for j, s in zip(order[i], sc[i]) if j >= 0])
        return out

    def rerank_device(self, engine, q_embs, cand_rows, q_rec, cand_rec, topk: int = 0, counts=None):
        """All-device batched rerank (bench / serving path): torch CUDA tensors in and out.
        ``q_embs`` (B, D) fp32, ``cand_rows`` (B, K) int64 global rows, ``q_rec`` (B) and
        ``cand_rec`` (B, K) int64 record-table rows, ``counts`` (B) int32 valid candidates per query (default K).
        Returns ``(order (B, keep) int32, scores (B, keep, 4) fp64)``."""
        import torch
        b, k = cand_rows.shape
        keep = topk if 0 < topk < k else k
        order = torch.empty((b, keep), dtype=torch.int32, device=cand_rows.device)
        sc = torch.empty((b, keep, 4), dtype=torch.float64, device=cand_rows.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.mmr_rerank(engine._handle, self._tables, _lib.ptr(q_embs), None,
                                            _lib.ptr(cand_rows), _lib.ptr(q_rec), _lib.ptr(cand_rec), _lib.ptr(counts), b, k,
                                            int(q_embs.shape[1]), self.alpha, self.beta, self.gamma, int(topk),
                                            _lib.ptr(order), _lib.ptr(sc), _lib.current_stream(self.device)))
        return order, sc

    def features_device(self, engine, q_embs, cand_rows, q_rec, cand_rec):
        """Raw (cos, Jaccard, KG cos) features (B, K, 3) fp64 on the device; the cosine of a candidate
        whose row is not in ``engine``'s shard is 0 (summed across ranks by the sharded path)."""
        import torch
        b, k = cand_rows.shape
        raw = torch.empty((b, k, 3), dtype=torch.float64, device=cand_rows.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.mmr_rerank_features(engine._handle, self._tables, _lib.ptr(q_embs), None,
                                                     _lib.ptr(cand_rows), _lib.ptr(q_rec), _lib.ptr(cand_rec), None,
                                                     b, k, int(q_embs.shape[1]), _lib.ptr(raw), None,
                                                     _lib.current_stream(self.device)))
        return raw

    def candidate_cosine_device(self, engine, q_embs, cand_rows, out=None):
        """fp32 cosine(q, gallery row) of (B, K) candidates given by GLOBAL row id, for the rows
        ``engine``'s shard owns (0 elsewhere) -- evaluated on the LOCAL top-K before the exchange."""
        import torch
        b, k = cand_rows.shape
        if out is None:
            out = torch.empty((b, k), dtype=torch.float32, device=cand_rows.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.mmr_candidate_cosine(engine._handle, _lib.ptr(q_embs), _lib.ptr(cand_rows), b, k,
                                                      int(q_embs.shape[1]), _lib.ptr(out), None,
                                                      _lib.current_stream(self.device)))
        return out

    def rerank_with_cos_device(self, emb_cos, q_rec, cand_rec, topk: int = 0, counts=None):
        """Rerank with the embedding cosines supplied (sharded path): label Jaccard + KG cosine from
        the replicated tables, then the same min-max / combine / ordering.  ``counts`` (B) int32: valid
        candidates per query (default K everywhere)."""
        import torch
        b, k = emb_cos.shape
        keep = topk if 0 < topk < k else k
        order = torch.empty((b, keep), dtype=torch.int32, device=emb_cos.device)
        sc = torch.empty((b, keep, 4), dtype=torch.float64, device=emb_cos.device)
        emb_cos, q_rec, cand_rec = emb_cos.contiguous(), q_rec.contiguous(), cand_rec.contiguous()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.mmr_rerank_with_cos(self._tables, _lib.ptr(emb_cos), _lib.ptr(q_rec),
                                                     _lib.ptr(cand_rec), _lib.ptr(counts), b, k, self.alpha, self.beta,
                                                     self.gamma, int(topk), _lib.ptr(order), _lib.ptr(sc),
                                                     self.device, _lib.current_stream(self.device)))
        return order, sc

    # the rerank's embedding feature on the batched device paths (ShardedSearcher.retrieve_reranked):
    # "search_score" = the score the candidate was found with (the same cosine, reranker.py:298, without a second
    # K x D gather); "recompute" = fp32 cosine of the query as passed against the stored row
    emb_feature = "search_score"

    def fused_tail_ok(self, k: int) -> bool:
        """Shapes the fused tail kernel covers (csrc/rerank_tail.cuh): k <= 128, KG dimension a multiple of 4
        and <= 512."""
        d_kg = getattr(self, "_d_kg", None)
        return 1 <= int(k) <= 128 and d_kg is not None and d_kg % 4 == 0 and d_kg <= 512

    def rerank_scored_device(self, rows, scores, q_rec, topk: int = 0, out=None, want_scores4: bool = False):
        """The fused tail of a search step (``mmr_rerank_scored``): rerank a search result ``(rows, scores)``
        (B, K) CUDA tensors whose embedding feature is the search score and whose record index is the global row
        id.  Returns ``(ids (B, keep) int64, combined scores (B, keep) fp64)`` (+ ``(B, keep, 4)`` score columns
        with ``want_scores4``) -- what ``retrieve(..., reranker=...)`` returns for a batch."""
        import torch
        b, k = rows.shape
        keep = topk if 0 < topk < k else k
        if out is None:
            ids = torch.empty((b, keep), dtype=torch.int64, device=rows.device)
            fin = torch.empty((b, keep), dtype=torch.float64, device=rows.device)
        else:
            ids, fin = out
        s4 = torch.empty((b, keep, 4), dtype=torch.float64, device=rows.device) if want_scores4 else None
        with torch.cuda.device(self.device):
            _lib.check(self._lib.mmr_rerank_scored(self._tables, _lib.ptr(rows), _lib.ptr(scores), _lib.ptr(q_rec), b, k,
                                                   self.alpha, self.beta, self.gamma, int(topk), _lib.ptr(ids),
                                                   _lib.ptr(fin), _lib.ptr(s4), self.device,
                                                   _lib.current_stream(self.device)))
        return (ids, fin, s4) if want_scores4 else (ids, fin)

    def combine_device(self, raw, topk: int = 0):
        import torch
        b, k, _ = raw.shape
        keep = topk if 0 < topk < k else k
        order = torch.empty((b, keep), dtype=torch.int32, device=raw.device)
        sc = torch.empty((b, keep, 4), dtype=torch.float64, device=raw.device)
        raw = raw.contiguous()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.mmr_rerank_combine(_lib.ptr(raw), None, b, k, self.alpha, self.beta, self.gamma,
                                                    int(topk), _lib.ptr(order), _lib.ptr(sc), self.device,
                                                    _lib.current_stream(self.device)))
        return order, sc

    def _call(self, engine, q, cand_emb, cand_rows, q_rec, cand_rec, counts, topk):
        import torch
        b = q.shape[0]
        k = cand_emb.shape[1] if cand_emb is not None else cand_rows.shape[1]
        keep = topk if 0 < topk < k else k
        order = np.empty((b, keep), dtype=np.int32)
        sc = np.empty((b, keep, 4), dtype=np.float64)
        q_rec = np.ascontiguousarray(q_rec, dtype=np.int64)        # keep alive across the call
        cand_rec = np.ascontiguousarray(cand_rec, dtype=np.int64)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.mmr_rerank(
                engine._handle if engine is not None else None, self._tables, _lib.ptr(q), _lib.ptr(cand_emb),
                _lib.ptr(cand_rows), _lib.ptr(q_rec), _lib.ptr(cand_rec),
                _lib.ptr(counts), b, k, q.shape[1], float(self.alpha), float(self.beta), float(self.gamma), int(topk),
                _lib.ptr(order), _lib.ptr(sc), _lib.current_stream(self.device)))
        return order, sc
