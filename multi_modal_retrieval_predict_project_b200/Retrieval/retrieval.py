"""B200-native mirror of the reference's ``Retrieval/retrieval.py``.

Same public names and signatures -- ``RetrievalEngine``, ``make_retrieval_engine``,
``engine.retrieve(query_emb, K, ...)`` -> ``(List[str], List[float])`` -- but the search is
EXACT brute-force cosine + fused top-K on the GPU (libmmr_b200.so), i.e. the arithmetic of
``cosine_similarity(Q, G)`` + ``np.argsort(row)[::-1][:K]`` (reference
``Evaluate/retrieval_overlap.py:85,90``; ``Retrieval/retrieval.py:128,134``) instead of the
approximate DenseLinkSearch walk (``Retrieval/retrieval.py:140-244``, which drops every node it
pops -- SURVEY.md section 0 finding 2).  There is no CPU fallback.
"""
from __future__ import annotations

import abc
import json
import os
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np

from .. import _lib


class _VirtualIds:
    """ids[i] == prefix + str(offset + i) without materialising 10^7 Python strings."""

    def __init__(self, n: int, prefix: str = "g", offset: int = 0):
        self.n, self.prefix, self.offset = int(n), prefix, int(offset)

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(self.n))]
        i = int(i)
        if i < 0:
            i += self.n
        if not 0 <= i < self.n:
            raise IndexError(i)
        return f"{self.prefix}{self.offset + i}"

    def __iter__(self):
        return (self[i] for i in range(self.n))


class _VirtualId2Idx:
    def __init__(self, ids: _VirtualIds):
        self._ids = ids

    def get(self, key, default=None):
        key = str(key)
        p = self._ids.prefix
        if not key.startswith(p):
            return default
        try:
            i = int(key[len(p):]) - self._ids.offset
        except ValueError:
            return default
        return i if 0 <= i < self._ids.n and self._ids[i] == key else default

    def __contains__(self, key):
        return self.get(key) is not None

    def __getitem__(self, key):
        i = self.get(key)
        if i is None:
            raise KeyError(key)
        return i


class RetrievalEngine(abc.ABC):
    """Abstract base: loads the gallery files once (reference ``retrieval.py:18-50``).

    ``embs`` fp32 ``(N, D)``, ``ids`` list of N ids, ``id2idx`` (last duplicate wins).
    """

    def __init__(self, features_path: str, ids_path: str):
        self.embs = np.load(features_path).astype("float32")
        with open(ids_path, "r") as f:
            self.ids = json.load(f)
        self.id2idx = {str(self.ids[i]): i for i in range(len(self.ids))}
        assert self.embs.shape[0] == len(self.ids), "embeddings count != ids count"

    @abc.abstractmethod
    def retrieve(self, query_emb: np.ndarray, K: int = 5, **kwargs) -> Tuple[List[str], List[float]]:
        """Given a query embedding (D,) or (1,D), return top-K IDs and their scores."""

    def get_embeddings_for_ids(self, ids: List[str]) -> np.ndarray:
        """Embeddings in the same order as ``ids`` (zeros if missing) -- reference ``:41-50``."""
        rows = []
        for _id in ids:
            idx = self.id2idx.get(str(_id), None)
            if idx is None:
                rows.append(np.zeros(self.embs.shape[1], dtype=self.embs.dtype))
            else:
                rows.append(self.embs[idx])
        return np.vstack(rows)


_DTYPES = {"float32": _lib.MMR_F32, "fp32": _lib.MMR_F32, "f32": _lib.MMR_F32,
           "bfloat16": _lib.MMR_BF16, "bf16": _lib.MMR_BF16}


def _is_tensor(x) -> bool:
    return hasattr(x, "data_ptr") and hasattr(x, "is_cuda")


class B200RetrievalEngine(RetrievalEngine):
    """Exact cosine top-K over a gallery (shard) resident in HBM.

    ``dtype``: storage type of the gallery on the device -- ``"float32"`` (default: the
    reference's own precision, scan kernel) or ``"bfloat16"`` (half the HBM bytes; tcgen05 GEMM
    for batched queries; queries are rounded to bf16 too).
    ``algo``: ``"auto"`` | ``"scan"`` | ``"gemm"``.
    The DLS keyword arguments of the reference factory (``link_threshold``, ``max_links``,
    ``fdb_path``, ``name``) are accepted and ignored.
    """

    def __init__(self, features_path: Optional[str] = None, ids_path: Optional[str] = None, *,
                 dtype: str = "float32", device=None, algo: str = "auto", embs=None, ids=None,
                 row_offset: int = 0, keep_host: bool = True, borrow: bool = False, **_ignored):
        if features_path is not None:
            super().__init__(features_path, ids_path)
            src = self.embs
        else:
            if embs is None:
                raise ValueError("provide features_path/ids_path or embs")
            src = embs
            n = int(src.shape[0])
            if ids is None:
                ids = _VirtualIds(n, "g", row_offset)
                self.id2idx = _VirtualId2Idx(ids)
            else:
                self.id2idx = {str(ids[i]): i for i in range(len(ids))}
            self.ids = ids
            assert n == len(self.ids), "embeddings count != ids count"
            if _is_tensor(src):
                self.embs = src.detach().float().cpu().numpy() if (keep_host and not src.is_cuda) else None
            else:
                src = np.ascontiguousarray(src)
                if src.dtype != np.float32:
                    src = src.astype("float32")
                self.embs = src if keep_host else None
        if dtype not in _DTYPES:
            raise ValueError(f"Unknown gallery dtype: {dtype}")
        if algo not in _lib.ALGOS:
            raise ValueError(f"Unknown search algorithm: {algo}")
        self.algo = algo
        self.dtype = "bfloat16" if _DTYPES[dtype] == _lib.MMR_BF16 else "float32"
        self.device = _lib.require_cuda(device)
        self.row_offset = int(row_offset)
        self._lib = _lib.load()
        self._handle = None
        self._borrowed = None
        import torch
        if _is_tensor(src):
            t = src.detach()
            if t.dtype == torch.bfloat16:
                dt_in = _lib.MMR_BF16
            else:
                t = t.float()
                dt_in = _lib.MMR_F32
            t = t.contiguous()
            n, d = int(t.shape[0]), int(t.shape[1])
            src_ptr = _lib.ptr(t)
            keepalive = t
        else:
            n, d = int(src.shape[0]), int(src.shape[1])
            dt_in = _lib.MMR_F32
            src_ptr = _lib.ptr(src)
            keepalive = src
        flags = 0
        if borrow:
            flags |= _lib.FLAG_BORROW
            self._borrowed = keepalive
        self.n, self.dim = n, d
        h = _lib.C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.mmr_index_create(_lib.C.byref(h), src_ptr, n, d, dt_in, _DTYPES[dtype],
                                                  self.row_offset, self.device, flags,
                                                  _lib.current_stream(self.device)))
        self._handle = h
        del keepalive

    # -- lifetime -----------------------------------------------------------------------------
    def close(self):
        h, self._handle = self._handle, None
        if h is not None and self._lib is not None:
            self._lib.mmr_index_destroy(h)
        self._borrowed = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @classmethod
    def from_arrays(cls, embs, ids=None, **kwargs) -> "B200RetrievalEngine":
        """Build from an in-memory gallery: numpy ``(N, D)`` or a (CUDA) torch tensor."""
        return cls(None, None, embs=embs, ids=ids, **kwargs)

    def hbm_bytes(self) -> int:
        v = _lib.C.c_int64()
        _lib.check(self._lib.mmr_index_info(self._handle, None, None, None, None, None, None, _lib.C.byref(v)))
        return v.value

    def profile(self, enable: bool):
        """Start/stop live CUDA-event timing of the dominant search kernel; returns
        ``(summed_ms, launches)`` accumulated since the previous call (mmr_index_profile)."""
        ms, cnt = _lib.C.c_double(), _lib.C.c_int32()
        _lib.check(self._lib.mmr_index_profile(self._handle, 1 if enable else 0, _lib.C.byref(ms), _lib.C.byref(cnt)))
        return ms.value, cnt.value

    def tune(self, variant: Optional[str] = None, parts: Optional[int] = None, pair: Optional[bool] = None):
        """Pin the GEMM kernel instantiation of this engine's searches (mmr_index_tune): ``variant``
        ``"auto"`` | ``"long"`` | ``"short"``, ``parts`` = gallery parts per query tile (0 = automatic),
        ``pair`` = use cta_group::2 CTA pairs.  All choices return the same results; the parity tests
        run each of them explicitly."""
        if variant is not None:
            _lib.check(self._lib.mmr_index_tune(self._handle, _lib.TUNE_GEMM_VARIANT, _lib.GEMM_VARIANTS[variant]))
        if parts is not None:
            _lib.check(self._lib.mmr_index_tune(self._handle, _lib.TUNE_GEMM_PARTS, int(parts)))
        if pair is not None:
            _lib.check(self._lib.mmr_index_tune(self._handle, _lib.TUNE_GEMM_PAIR, 0 if pair else 1))
        return self

    def last_plan(self) -> dict:
        """What the most recent search ran (mmr_index_last_plan): ``{"algo": "scan"|"gemm", "variant":
        "long"|"short"|None, "pair": bool, "parts": int, "tiles_per_part": int}``."""
        v = [_lib.C.c_int32() for _ in range(5)]
        _lib.check(self._lib.mmr_index_last_plan(self._handle, *[_lib.C.byref(x) for x in v]))
        algo = {0: None, _lib.ALGO_SCAN: "scan", _lib.ALGO_GEMM: "gemm"}[v[0].value]
        return {"algo": algo, "variant": {0: None, 1: "long", 2: "short"}[v[1].value], "pair": bool(v[2].value),
                "parts": v[3].value, "tiles_per_part": v[4].value}

    # -- batched search (the hot path) ----------------------------------------------------------
    def search(self, queries, K: int, exclude_rows=None, algo: Optional[str] = None, out_rows=None, out_scores=None):
        """Exact top-K for a batch.  ``queries``: numpy ``(B, D)`` / ``(D,)`` (host) or a torch
        tensor (CUDA tensors stay on the device: no host round trip).  Returns ``(rows, scores)``
        of shape ``(B, K)`` -- numpy for numpy input, CUDA tensors for CUDA input; rows are
        GLOBAL row ids (int64, -1 padding when the shard has fewer than K rows), best first.
        """
        import torch
        if self._handle is None:
            raise RuntimeError("engine is closed")
        K = int(K)
        if K < 1:
            raise ValueError("K must be >= 1")
        a = _lib.ALGOS[algo if algo is not None else self.algo]
        on_device = _is_tensor(queries) and queries.is_cuda
        if _is_tensor(queries):
            q = queries.detach()
            if q.dim() == 1:
                q = q.unsqueeze(0)
            if q.dtype == torch.bfloat16:
                qd = _lib.MMR_BF16
            else:
                q = q.float()
                qd = _lib.MMR_F32
            q = q.contiguous()
            if not on_device:
                q = q.numpy()
        else:
            q = np.asarray(queries)
            if q.ndim == 1:
                q = q.reshape(1, -1)
            q = np.ascontiguousarray(q, dtype=np.float32)
            qd = _lib.MMR_F32
        if q.shape[1] != self.dim:
            raise ValueError(f"query dimension {q.shape[1]} != gallery dimension {self.dim}")
        b = int(q.shape[0])
        ex = None
        if exclude_rows is not None:
            if _is_tensor(exclude_rows):
                ex = exclude_rows.detach().to(torch.int64).contiguous()
                if not ex.is_cuda:
                    ex = ex.numpy()
            else:
                ex = np.ascontiguousarray(exclude_rows, dtype=np.int64)
            if int(ex.shape[0]) != b:
                raise ValueError("exclude_rows must have one entry per query")
        if out_rows is not None:  # caller-provided (device) result buffers, e.g. views into an exchange blob
            rows, scores = out_rows, out_scores
            assert rows.is_contiguous() and scores.is_contiguous() and tuple(rows.shape) == (b, K) == tuple(scores.shape)
            assert rows.dtype == torch.int64 and scores.dtype == torch.float32
        elif on_device:
            rows = torch.empty((b, K), dtype=torch.int64, device=q.device)
            scores = torch.empty((b, K), dtype=torch.float32, device=q.device)
        else:
            rows = np.empty((b, K), dtype=np.int64)
            scores = np.empty((b, K), dtype=np.float32)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.mmr_search(self._handle, _lib.ptr(q), b, qd, K, a, _lib.ptr(ex),
                                            _lib.ptr(scores), _lib.ptr(rows), _lib.current_stream(self.device)))
        return rows, scores

    # -- reference-compatible single-query entry point -------------------------------------------
    def retrieve(self, query_emb, K: int = 5, seed_size: int = 5, max_steps: int = 100,
                 candidate_multiplier: int = 10, reranker=None, query_id=None,
                 rerank_topk: Optional[int] = None, seed: Optional[int] = None):
        """Signature of ``DLSRetrievalEngine.retrieve`` (reference ``retrieval.py:140-151``).

        ``seed_size``, ``max_steps``, ``candidate_multiplier`` and ``seed`` parameterise the
        reference's approximate walk and are ignored: the search is exact.  ``(D,)`` / ``(1, D)``
        input returns ``(List[str], List[float])``; ``(B, D)`` with B > 1 returns nested lists
        (the callers already accept those: ``web/app.py:431-437``).  With ``reranker`` and
        ``query_id`` the K candidates are reranked and ``rerank_topk or K`` combined scores are
        returned (``retrieval.py:257-269``).
        """
        q = query_emb
        if _is_tensor(q):
            single = q.dim() == 1 or q.shape[0] == 1
        else:
            q = np.asarray(q)
            single = q.ndim == 1 or q.shape[0] == 1
            q = q.astype("float32")
            if single:
                q = q.reshape(1, -1)
        rows, scores = self.search(q, K)
        if _is_tensor(rows):
            rows = rows.cpu().numpy()
            scores = scores.cpu().numpy()
        b = rows.shape[0]
        qids = [query_id] if (single or not isinstance(query_id, (list, tuple))) else list(query_id)
        if len(qids) != b:
            qids = (qids * b)[:b]
        out_ids, out_scores, local = [], [], []
        for i in range(b):
            valid = rows[i] >= 0
            r = rows[i][valid] - self.row_offset
            local.append(r)
            out_ids.append([self.ids[int(j)] for j in r])
            out_scores.append([float(s) for s in scores[i][valid]])
        todo = [i for i in range(b) if reranker is not None and qids[i] is not None]
        if todo and hasattr(reranker, "rerank_rows"):
            # ONE device call for the whole batch (the reference loops over queries in Python)
            qh = q.detach().float().cpu().numpy() if _is_tensor(q) else q
            kmax = max(len(local[i]) for i in todo)
            cand = -np.ones((len(todo), kmax), dtype=np.int64)
            q_embs = np.empty((len(todo), self.dim), dtype=np.float32)
            for t, i in enumerate(todo):
                cand[t, : len(local[i])] = local[i] + self.row_offset
                # the STORED gallery row replaces q when query_id is a gallery id (reference retrieval.py:251-254)
                q_embs[t] = self.get_embeddings_for_ids([qids[i]])[0] if str(qids[i]) in self.id2idx else qh[i]
            res = reranker.rerank_rows(self, [qids[i] for i in todo], q_embs, cand, [out_ids[i] for i in todo],
                                       rerank_topk or K)
            for t, i in enumerate(todo):
                out_ids[i], out_scores[i] = [x[0] for x in res[t]], [x[1] for x in res[t]]
        else:
            for i in todo:
                qv = q[i].detach().float().cpu().numpy() if _is_tensor(q) else q[i]
                out_ids[i], out_scores[i] = self._rerank(reranker, qids[i], qv, out_ids[i], local[i], rerank_topk or K)
        if single:
            return out_ids[0], out_scores[0]
        return out_ids, out_scores

    def _rerank(self, reranker, query_id, q_vec, ids, local_rows, topk):
        # query embedding for the rerank cosine: the STORED gallery row when query_id is a gallery
        # id, else the vector passed in (reference retrieval.py:251-254)
        if str(query_id) in self.id2idx:
            q_emb = self.get_embeddings_for_ids([query_id])[0]
        else:
            q_emb = q_vec
        if hasattr(reranker, "rerank_rows"):
            reranked = reranker.rerank_rows(self, [query_id], np.asarray(q_emb, dtype=np.float32).reshape(1, -1),
                                            np.asarray(local_rows, dtype=np.int64).reshape(1, -1) + self.row_offset,
                                            [ids], topk)[0]
        else:  # any object with the reference's Reranker.rerank interface
            cand_embs = self.get_embeddings_for_ids(ids)
            lookup = {str(rid): emb for rid, emb in zip(ids, cand_embs)}
            lookup[str(query_id)] = q_emb
            reranked = reranker.rerank(query_id=query_id, candidate_ids=ids, candidate_embs=cand_embs,
                                       candidate_emb_lookup=lookup, topk=topk)
        return [t[0] for t in reranked], [t[1] for t in reranked]

    def get_embeddings_for_ids(self, ids: List[str]) -> np.ndarray:
        if self.embs is not None:
            return super().get_embeddings_for_ids(ids)
        import torch
        rows = np.array([self.id2idx.get(str(i), -1) if self.id2idx.get(str(i)) is not None else -1 for i in ids],
                        dtype=np.int64)
        rows = np.where(rows >= 0, rows + self.row_offset, -1)
        out = np.empty((len(ids), self.dim), dtype=np.float32)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.mmr_index_get_rows(self._handle, _lib.ptr(rows), len(ids), _lib.ptr(out),
                                                    _lib.current_stream(self.device)))
        return out

    # -- "next" row f1: persisted device-format gallery ---------------------------------------------
    def save_blob(self, path: str, chunk: int = 1 << 18) -> str:
        """Write the gallery exactly as it is stored in HBM (bf16 indexes: the ROUNDED values as raw
        bf16 bits -- half the bytes of the ``.npy`` the reference writes, and no fp32 -> bf16 pass on
        reload) plus the ids, as one ``.npz``.  ``load_blob`` rebuilds a bit-identical index."""
        import json
        import torch
        parts = []
        for s in range(0, self.n, chunk):
            rows = np.arange(s, min(self.n, s + chunk), dtype=np.int64) + self.row_offset
            out = np.empty((len(rows), self.dim), dtype=np.float32)
            with torch.cuda.device(self.device):
                _lib.check(self._lib.mmr_index_get_rows(self._handle, _lib.ptr(rows), len(rows), _lib.ptr(out),
                                                        _lib.current_stream(self.device)))
            if self.dtype == "bfloat16":
                bits = out.view(np.uint32)
                assert not np.any(bits & np.uint32(0xFFFF)), "stored values are not bf16-representable"
                parts.append((bits >> np.uint32(16)).astype(np.uint16))
            else:
                parts.append(out)
        data = np.concatenate(parts, axis=0) if parts else np.zeros((0, self.dim), np.float32)
        ids = None if isinstance(self.ids, _VirtualIds) else json.dumps([str(i) for i in self.ids])
        meta = json.dumps({"dtype": self.dtype, "n": self.n, "dim": self.dim, "row_offset": self.row_offset})
        if not str(path).endswith(".npz"):
            path = str(path) + ".npz"
        np.savez(path, data=data, meta=np.array(meta), ids=np.array(ids if ids is not None else ""))
        return str(path)

    @classmethod
    def load_blob(cls, path: str, **kwargs) -> "B200RetrievalEngine":
        """Engine from a ``save_blob`` file: the stored values go to the device unchanged."""
        import json
        import torch
        z = np.load(path, allow_pickle=False)
        meta = json.loads(str(z["meta"]))
        ids_json = str(z["ids"])
        ids = json.loads(ids_json) if ids_json else None
        kwargs.setdefault("row_offset", int(meta["row_offset"]))
        kwargs.setdefault("keep_host", False)
        if meta["dtype"] == "bfloat16":
            t = torch.from_numpy(np.ascontiguousarray(z["data"]).view(np.int16)).view(torch.bfloat16)
            return cls.from_arrays(t, ids=ids, dtype="bfloat16", **kwargs)
        return cls.from_arrays(np.ascontiguousarray(z["data"], dtype=np.float32), ids=ids, dtype="float32", **kwargs)

    # -- "next" row: the link graph of the legacy engine, built on the GPU -------------------------
    def build_link_graph(self, threshold: float = 0.5, max_links: int = 10, batch: int = 4096) -> List[List[int]]:
        """``DLSRetrievalEngine._build_link_graph`` (reference ``retrieval.py:121-138``) as GPU
        all-pairs top-``max_links`` with the diagonal excluded and a score threshold -- without the
        N x N matrix the reference materialises."""
        if self.embs is None:
            raise RuntimeError("build_link_graph needs the host copy of the gallery (keep_host=True)")
        graph: List[List[int]] = []
        for s in range(0, self.n, batch):
            e = min(self.n, s + batch)
            ex = np.arange(s, e, dtype=np.int64) + self.row_offset
            rows, scores = self.search(self.embs[s:e], max_links, exclude_rows=ex)
            for i in range(e - s):
                keep = (rows[i] >= 0) & (scores[i] >= threshold)
                graph.append([int(j) - self.row_offset for j in rows[i][keep]])
        return graph


class MultiGPURetrievalEngine(RetrievalEngine):
    """The reference's single-process engine interface over SEVERAL GPUs of one box:
    ``make_retrieval_engine(fp, ip, method="b200", devices=[0, 1, ...])``.  The gallery is split into
    contiguous row shards, one ``B200RetrievalEngine`` per device; ``retrieve`` keeps
    the reference signature (string ids, ``reranker=``, ``query_id=``; ``Retrieval/retrieval.py:140-151``):
    every shard is searched on its own device and stream (the launches overlap), the per-shard lists are
    copied to the first device and merged there (``mmr_merge_topk``), and for a rerank every shard evaluates the
    fp32 cosine of the candidates IT owns (``mmr_candidate_cosine``), which the first device sums and feeds to
    the label / KG rerank.  A device may be listed more than once (several shards on one GPU)."""

    def __init__(self, features_path: Optional[str] = None, ids_path: Optional[str] = None, *, devices: Sequence[int],
                 dtype: str = "float32", algo: str = "auto", embs=None, ids=None, **_ignored):
        if features_path is not None:
            super().__init__(features_path, ids_path)
        else:
            self.embs = np.ascontiguousarray(embs, dtype=np.float32)
            self.ids = list(ids) if ids is not None else _VirtualIds(len(self.embs))
            self.id2idx = ({str(self.ids[i]): i for i in range(len(self.ids))} if ids is not None
                           else _VirtualId2Idx(self.ids))
        if not devices:
            raise ValueError("devices must name at least one GPU")
        from ..sharded import shard_bounds
        self.devices = [_lib.require_cuda(d) for d in devices]
        self.n, self.dim = int(self.embs.shape[0]), int(self.embs.shape[1])
        self.row_offset = 0
        self.dtype = "bfloat16" if _DTYPES[dtype] == _lib.MMR_BF16 else "float32"
        self.shards: List[B200RetrievalEngine] = []
        for r, dev in enumerate(self.devices):
            lo, hi = shard_bounds(self.n, len(self.devices), r)
            self.shards.append(B200RetrievalEngine.from_arrays(self.embs[lo:hi], dtype=dtype, device=dev, algo=algo,
                                                               row_offset=lo, keep_host=False))
        self.device = self.devices[0]

    def close(self):
        for s in self.shards:
            s.close()
        self.shards = []

    def search(self, queries, K: int):
        """Exact global top-K of a batch: ``(rows (B, K) int64, scores (B, K) fp32)`` CUDA tensors on the
        first device, best first, -1 / -inf padding when the gallery has fewer than K rows."""
        import torch
        from ..sharded import merge_topk
        q = queries.detach().float() if _is_tensor(queries) else torch.from_numpy(np.ascontiguousarray(queries, dtype=np.float32))
        if q.dim() == 1:
            q = q.unsqueeze(0)
        if int(q.shape[1]) != self.dim:
            raise ValueError(f"query dimension {q.shape[1]} != gallery dimension {self.dim}")
        parts = []
        for s in self.shards:                                    # asynchronous on every device
            with torch.cuda.device(s.device):
                qd = q.to(torch.device("cuda", s.device), non_blocking=True)
                parts.append(s.search(qd, K))
        dev0 = torch.device("cuda", self.device)
        for s in self.shards:
            torch.cuda.current_stream(s.device).synchronize()   # peer copies below read the shards' results
        rows = torch.stack([r.to(dev0) for r, _ in parts])
        scores = torch.stack([sc.to(dev0) for _, sc in parts])
        if len(parts) == 1:
            return rows[0], scores[0]
        with torch.cuda.device(self.device):
            return merge_topk(scores, rows, K)

    def _candidate_cosines(self, q_embs: np.ndarray, cand_rows: np.ndarray):
        """(B, K) fp32 cosine(q, stored row) of candidates given by GLOBAL row, each evaluated by the shard that
        owns the row (safe_cos, reranker.py:135-142); -1 rows give 0.  Returns a CUDA tensor on the first device."""
        import torch
        b, k = cand_rows.shape
        dev0 = torch.device("cuda", self.device)
        total = torch.zeros((b, k), dtype=torch.float32, device=dev0)
        outs = []
        for s in self.shards:
            with torch.cuda.device(s.device):
                d = torch.device("cuda", s.device)
                qd = torch.from_numpy(q_embs).to(d)
                rd = torch.from_numpy(cand_rows).to(d)
                out = torch.empty((b, k), dtype=torch.float32, device=d)
                _lib.check(s._lib.mmr_candidate_cosine(s._handle, _lib.ptr(qd), _lib.ptr(rd), b, k, self.dim,
                                                       _lib.ptr(out), None, _lib.current_stream(s.device)))
                outs.append((out, qd, rd))
        for (out, _qd, _rd), s in zip(outs, self.shards):
            torch.cuda.current_stream(s.device).synchronize()
            total += out.to(dev0)
        return total

    def retrieve(self, query_emb, K: int = 5, seed_size: int = 5, max_steps: int = 100,
                 candidate_multiplier: int = 10, reranker=None, query_id=None,
                 rerank_topk: Optional[int] = None, seed: Optional[int] = None):
        """Signature and return types of ``DLSRetrievalEngine.retrieve`` (reference ``retrieval.py:140-151``);
        see ``B200RetrievalEngine.retrieve``."""
        import torch
        q = query_emb.detach().float().cpu().numpy() if _is_tensor(query_emb) else np.asarray(query_emb, dtype=np.float32)
        single = q.ndim == 1 or q.shape[0] == 1
        q = np.ascontiguousarray(q.reshape(1, -1) if q.ndim == 1 else q, dtype=np.float32)
        rows_d, scores_d = self.search(q, K)
        rows, scores = rows_d.cpu().numpy(), scores_d.cpu().numpy()
        b = rows.shape[0]
        qids = [query_id] if (single or not isinstance(query_id, (list, tuple))) else list(query_id)
        if len(qids) != b:
            qids = (qids * b)[:b]
        out_ids = [[self.ids[int(j)] for j in rows[i][rows[i] >= 0]] for i in range(b)]
        out_scores = [[float(x) for x in scores[i][rows[i] >= 0]] for i in range(b)]
        todo = [i for i in range(b) if reranker is not None and qids[i] is not None]
        if todo:
            if not hasattr(reranker, "rerank_with_cos_device"):
                raise TypeError("the multi-GPU engine reranks with the B200 Reranker (device tables)")
            if reranker.device != self.device:
                raise ValueError("the Reranker's tables must live on the engine's first device")
            kmax = max(len(out_ids[i]) for i in todo)
            cand = -np.ones((len(todo), kmax), dtype=np.int64)
            crec = -np.ones((len(todo), kmax), dtype=np.int64)
            q_embs = np.empty((len(todo), self.dim), dtype=np.float32)
            counts = np.zeros(len(todo), dtype=np.int32)
            for t, i in enumerate(todo):
                n_i = len(out_ids[i])
                counts[t] = n_i
                cand[t, :n_i] = rows[i][:n_i]
                crec[t, :n_i] = reranker._rec_rows(out_ids[i])
                q_embs[t] = self.get_embeddings_for_ids([qids[i]])[0] if str(qids[i]) in self.id2idx else q[i]
            cos = self._candidate_cosines(q_embs, cand)
            dev0 = torch.device("cuda", self.device)
            keep = rerank_topk or K
            order, sc = reranker.rerank_with_cos_device(cos, torch.from_numpy(reranker._rec_rows([qids[i] for i in todo])).to(dev0),
                                                        torch.from_numpy(crec).to(dev0), keep,
                                                        counts=torch.from_numpy(counts).to(dev0))
            order, sc = order.cpu().numpy(), sc.cpu().numpy()
            for t, i in enumerate(todo):
                sel = [(int(j), float(s4[0])) for j, s4 in zip(order[t], sc[t]) if j >= 0]
                out_ids[i], out_scores[i] = [out_ids[i][j] for j, _ in sel], [f for _, f in sel]
        if single:
            return out_ids[0], out_scores[0]
        return out_ids, out_scores


def make_retrieval_engine(features_path: str, ids_path: str, method: str = "dls", **kwargs) -> RetrievalEngine:
    """Factory with the reference's signature (``retrieval.py:273-304``).

    ``"b200"`` / ``"exact"`` / ``"cuda"`` -> exact GPU engine (kwargs ``dtype``, ``device``,
    ``algo``); ``"bf16"`` is shorthand for ``dtype="bfloat16"``.  ``"dls"`` -- the only method the
    reference accepts -- also returns the exact engine: it is a strict improvement on the walk
    (every DLS result is drawn from the exact ranking) and its kwargs ``link_threshold``,
    ``max_links``, ``fdb_path``, ``name`` are accepted and ignored.  Anything else raises
    ``ValueError`` like the reference.  ``devices=[0, 1, ...]`` row-shards the gallery over several GPUs of
    the box behind the same ``retrieve`` signature (``MultiGPURetrievalEngine``).
    """
    method = method.lower()
    kwargs = dict(kwargs)
    if method in ("bf16", "b200-bf16"):
        kwargs["dtype"] = "bfloat16"
        method = "b200"
    if method in ("b200", "exact", "cuda", "dls"):
        devices = kwargs.pop("devices", None)
        if devices is not None and len(devices) > 0:        # row-sharded over several GPUs, same interface
            return MultiGPURetrievalEngine(features_path, ids_path, devices=list(devices), **kwargs)
        return B200RetrievalEngine(features_path, ids_path, **kwargs)
    raise ValueError(f"Unknown retrieval method: {method}")
