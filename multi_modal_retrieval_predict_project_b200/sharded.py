"""Row-sharded search across the GPUs of one NVSwitch box (one process per GPU).

Gallery rows are split into contiguous shards; every rank searches its own
shard (no data-path collective), then ONE exchange of the per-rank top-K ``(score fp32, global row
int64)`` lists over NVLink, followed by the on-device K-way merge (csrc/select.cu).

Two transports for the exchange:
  * ``PeerExchange`` (default for ``retrieve_reranked``): the lists are written straight into the
    owning rank's memory by the kernel that computes them (csrc/exchange.cu: CUDA-IPC peer mappings,
    NVLink stores, flag signalling) -- no collective library call on the data path;
  * NCCL ``all_gather_into_tensor`` (``search`` / ``search_rerank``, and the fallback when peer
    mappings are unavailable).
``torch.distributed`` is plumbing only (handle exchange, barriers).
"""
from __future__ import annotations

from typing import Optional, Tuple

from . import _lib


def shard_bounds(n_total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row block of ``rank``: ``[rank*ceil(N/G), min(N, (rank+1)*ceil(N/G)))``."""
    per = (n_total + world - 1) // world
    lo = min(n_total, rank * per)
    return lo, min(n_total, lo + per)


def weighted_shard_bounds(n_total: int, weights, rank: int, align: int = 256) -> Tuple[int, int]:
    """Contiguous row block of ``rank`` when the blocks are sized in proportion to ``weights`` (unequal devices,
    or ranks that also hold other data).  Interior cuts are multiples of ``align`` rows (the search kernel's
    tile) or the end of the gallery; every rank computes the same cuts from the same weights.  Global row = local row + block start, as
    with :func:`shard_bounds`.  bench.py keeps equal blocks: the kernel-time differences between the GPUs of one
    box (1-4 %) change sign between a calibration run and the timed run (measured at N = 2), so there is nothing
    stable to weight by."""
    w = [max(float(x), 0.0) for x in weights]
    tot = sum(w)
    if tot <= 0.0:
        return shard_bounds(n_total, len(w), rank)
    cuts, acc = [0], 0.0
    for x in w[:-1]:
        acc += x
        c = n_total if acc >= tot else int(round(n_total * acc / tot / align)) * align   # nothing left for the rest
        cuts.append(min(n_total, max(cuts[-1], c)))
    cuts.append(n_total)
    return cuts[rank], cuts[rank + 1]


def _valid_counts(rows):
    """(B,) int32 number of real candidates per query (-1 rows pad a result when the gallery has fewer than K rows):
    padding must stay out of the rerank's min-max scaling."""
    import torch
    return (rows >= 0).sum(dim=1).to(torch.int32).contiguous()


def merge_topk(scores, rows, k_out: int, want_src: bool = False):
    """K-way merge of ``(n_lists, B, K_in)`` CUDA tensors -> ``(B, k_out)`` best first (score desc,
    row asc).  With ``want_src`` also returns the flat source position ``list*K_in + j`` of every
    output slot (int32), so callers can gather payload carried next to the lists."""
    import torch
    assert scores.is_cuda and rows.is_cuda and scores.shape == rows.shape and scores.dim() == 3
    scores = scores.float().contiguous()
    rows = rows.to(torch.int64).contiguous()
    n_lists, b, k_in = scores.shape
    dev = scores.device.index or 0
    out_s = torch.empty((b, k_out), dtype=torch.float32, device=scores.device)
    out_r = torch.empty((b, k_out), dtype=torch.int64, device=scores.device)
    src = torch.empty((b, k_out), dtype=torch.int32, device=scores.device) if want_src else None
    lib = _lib.load()
    with torch.cuda.device(dev):
        _lib.check(lib.mmr_merge_topk(_lib.ptr(scores), _lib.ptr(rows), n_lists, b, k_in, k_out, _lib.ptr(out_s),
                                      _lib.ptr(out_r), _lib.ptr(src), dev, _lib.current_stream(dev)))
    if want_src:
        # the kernel reports positions in (list, j) order over a (list, b, k) layout: list*k_in + j
        return out_r, out_s, src
    return out_r, out_s


class PeerExchange:
    """This rank's exchange region + peer mappings of all other ranks' regions (csrc/exchange.cu).
    The CUDA IPC handles are exchanged once through ``torch.distributed`` (any backend)."""

    def __init__(self, device: int, b_max: int, k_max: int, group=None):
        import ctypes as C
        import torch
        import torch.distributed as dist
        self._lib = _lib.load()
        self.device = int(device)
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.b_max, self.k_max = int(b_max), int(k_max)
        self.step = 0
        # Every rank runs the same sequence of collectives whatever fails locally; ``self.error`` holds
        # the local failure (None = mapped) and the caller agrees on the outcome across ranks.
        self.error = None
        self._h = None
        h = C.c_void_p()
        mine = b""
        with torch.cuda.device(self.device):
            try:
                _lib.check(self._lib.mmr_exchange_create(C.byref(h), self.rank, self.world, self.b_max, self.k_max,
                                                         self.device))
                self._h = h
                buf = C.create_string_buffer(self._lib.mmr_exchange_handle_bytes())
                _lib.check(self._lib.mmr_exchange_handle(self._h, buf))
                mine = bytes(buf.raw)
            except Exception as e:  # noqa: BLE001
                self.error = str(e)
            handles = [None] * self.world
            dist.all_gather_object(handles, mine, group=group)
            if self.error is None and any(len(x) != len(mine) for x in handles):
                self.error = "a peer rank could not create its exchange region"
            if self.error is None:
                try:
                    _lib.check(self._lib.mmr_exchange_open(self._h, b"".join(handles)))
                except Exception as e:  # noqa: BLE001
                    self.error = str(e)

    def check(self):
        """Raise if a device-side wait of an earlier step gave up (timeout, abort, (b, k) mismatch)."""
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.check(self._lib.mmr_exchange_status(self._h, None))

    def abort(self):
        """Make every kernel that waits on this exchange -- here and on the peers -- give up at once."""
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.mmr_exchange_abort(self._h)

    def close(self, group=None, collective: bool = True):
        """Two-phase shutdown (collective: every rank calls it): unmap the peers' regions, barrier, then free
        this rank's region -- CUDA requires importers to close their mappings before the exporter frees.
        ``collective=False`` (interpreter exit, a peer is gone) skips the barrier."""
        if getattr(self, "_h", None) is None or not self._h.value:
            return
        import torch.distributed as dist
        self._lib.mmr_exchange_close_peers(self._h)
        if collective and self.world > 1 and dist.is_available() and dist.is_initialized():
            try:
                dist.barrier(group=group)
            except Exception:  # noqa: BLE001
                pass
        self._lib.mmr_exchange_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close(collective=False)
        except Exception:
            pass


class ShardedSearcher:
    """Per-rank shard + all-gather + merge.  ``engine`` holds this rank's rows (``row_offset`` set
    to the shard's first global row).  Works with any initialised ``torch.distributed`` process
    group whose backend supports CUDA tensors (NCCL); with world size 1 it is a pass-through."""

    def __init__(self, engine, group=None, merge=None, use_peer: bool = True):
        import torch.distributed as dist
        self.engine = engine
        self.group = group
        self.use_peer = use_peer
        self._merge = merge if merge is not None else merge_topk  # injectable for the CPU/gloo plumbing tests
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self._gs = self._gr = None

    def search(self, queries, K: int, algo: Optional[str] = None):
        import torch
        import torch.distributed as dist
        rows, scores = self.engine.search(queries, K, algo=algo)
        if self.world == 1:
            return rows, scores
        b = rows.shape[0]
        if self._gs is None or self._gs.shape[1:] != scores.shape:
            self._gs = torch.empty((self.world, b, K), dtype=scores.dtype, device=scores.device)
            self._gr = torch.empty((self.world, b, K), dtype=torch.int64, device=scores.device)
        # concatenated-along-dim-0 output form (accepted by both NCCL and gloo); rank-major == list-major
        dist.all_gather_into_tensor(self._gs.view(self.world * b, K), scores.contiguous(), group=self.group)
        dist.all_gather_into_tensor(self._gr.view(self.world * b, K), rows.contiguous(), group=self.group)
        return self._merge(self._gs, self._gr, K)

    # ------------------------------------------------------------------------------------------
    # one-collective step: local top-K + local cosines -> ONE all-gather of a per-rank blob
    # [scores f32 | cosines f32 | rows i64] -> strided merge straight out of the gathered buffer
    # ------------------------------------------------------------------------------------------
    def _local_blob(self, reranker, queries, K, algo):
        """Per-rank part of the step: search the shard, cosine of the local candidates, all-gather."""
        import torch
        import torch.distributed as dist
        eng = self.engine
        b = int(queries.shape[0])
        bk = b * K
        dev = queries.device
        if getattr(self, "_blob", None) is None or self._blob.numel() != bk * 16:
            self._blob = torch.empty(bk * 16, dtype=torch.uint8, device=dev)
            self._gblob = torch.empty(self.world * bk * 16, dtype=torch.uint8, device=dev)
        f = self._blob[: bk * 8].view(torch.float32)
        scores_v, cos_v = f[:bk].view(b, K), f[bk:].view(b, K)
        rows_v = self._blob[bk * 8:].view(torch.int64).view(b, K)
        eng.search(queries, K, algo=algo, out_rows=rows_v, out_scores=scores_v)
        if getattr(reranker, "emb_feature", "search_score") == "search_score":
            cos_v.copy_(scores_v)     # the rerank's embedding feature IS the search score (reranker.py:298)
        else:
            reranker.candidate_cosine_device(eng, queries, rows_v, out=cos_v)
        dist.all_gather_into_tensor(self._gblob, self._blob, group=self.group)
        return self._gblob

    def _merge_slice(self, gblob, b, K, q_lo, q_hi):
        """Merge queries [q_lo, q_hi) out of the gathered blobs -> (rows, scores, cosines)."""
        import torch
        bk = b * K
        dev = gblob.device
        n = q_hi - q_lo
        gf = gblob.view(torch.float32)            # per-rank stride: 4*bk floats
        gi = gblob.view(torch.int64)              # per-rank stride: 2*bk int64, rows start at +bk
        out_s = torch.empty((n, K), dtype=torch.float32, device=dev)
        out_r = torch.empty((n, K), dtype=torch.int64, device=dev)
        src = torch.empty((n, K), dtype=torch.int32, device=dev)
        cos = torch.empty((n, K), dtype=torch.float32, device=dev)
        lib = _lib.load()
        d = dev.index or 0
        off = q_lo * K
        with torch.cuda.device(d):
            st = _lib.current_stream(d)
            _lib.check(lib.mmr_merge_topk_strided(gf.data_ptr() + off * 4, gi.data_ptr() + (bk + off) * 8, self.world,
                                                  n, K, 4 * bk, 2 * bk, K, _lib.ptr(out_s), _lib.ptr(out_r),
                                                  _lib.ptr(src), d, st))
            _lib.check(lib.mmr_gather_payload(gf.data_ptr() + (bk + off) * 4, 4 * bk, _lib.ptr(src), n, K, K,
                                              _lib.ptr(cos), d, st))
        return out_r, out_s, cos

    def search_rerank(self, reranker, queries, K: int, q_rec, topk: int = 0, algo: Optional[str] = None):
        """The whole sharded step with ONE collective.  Every rank: local top-K (+ the embedding feature of
        its own K candidates), then a single NCCL all-gather of one blob per rank ``[scores f32 | cosines
        f32 | rows i64]``; the strided on-device merge reads the gathered blobs in place and reports where
        each winner came from, the cosines follow through that index, and label/KG features come from the
        replicated tables.  The record index of a candidate is its global row id.  Returns ``(rows,
        scores, order, rerank_scores)``."""
        eng = self.engine
        b = int(queries.shape[0])
        if self.world == 1:
            rows, scores = eng.search(queries, K, algo=algo)
            if getattr(reranker, "emb_feature", "search_score") == "search_score":
                order, sc = reranker.rerank_with_cos_device(scores, q_rec, rows, topk, counts=_valid_counts(rows))
            else:
                order, sc = reranker.rerank_device(eng, queries, rows, q_rec, rows, topk, counts=_valid_counts(rows))
            return rows, scores, order, sc
        gblob = self._local_blob(reranker, queries, K, algo)
        out_r, out_s, cos = self._merge_slice(gblob, b, K, 0, b)
        order, sc = reranker.rerank_with_cos_device(cos, q_rec, out_r, topk, counts=_valid_counts(out_r))
        return out_r, out_s, order, sc

    def retrieve_reranked(self, reranker, queries, K: int, q_rec, topk: int = 0, algo: Optional[str] = None,
                          buffers=None):
        """What ``retrieve(q, K, reranker=..., query_id=...)`` returns (Retrieval/retrieval.py:257-269)
        for a batch: ``(ids (B, keep) int64 in reranked order, combined scores (B, keep) fp64)``.  The record
        index of a candidate is its global row id; the rerank's embedding feature (reranker.py:298) is the
        search score unless ``reranker.emb_feature == "recompute"``.

        Single shard: search -> ONE fused kernel (features, min-max, combine, order; ``mmr_rerank_scored``).
        Sharded: every rank merges and reranks only ITS slice of the queries and every rank ends with the
        full result.  With the peer exchange (default) the search's selection kernel stores the local lists
        straight into the owner ranks' memory and one kernel per owner merges, reranks and publishes
        (csrc/exchange.cu); the returned tensors are views of this rank's exchange region, valid until the
        call after next.  Otherwise (``use_peer=False``, K > 128, fp32 index ...): one NCCL all-gather of the
        per-rank blobs and two of the result slices.

        ``buffers`` (single shard, fused tail): a dict the call fills with / reuses ``rows``, ``scores``, ``ids``,
        ``fin`` tensors, so that a serving loop allocates nothing per step."""
        import torch
        import torch.distributed as dist
        eng = self.engine
        b = int(queries.shape[0])
        keep = topk if 0 < topk < K else K
        dev = queries.device
        d = dev.index or 0
        lib = _lib.load()
        by_score = getattr(reranker, "emb_feature", "search_score") == "search_score"
        if self.world == 1:
            if buffers is not None and by_score and reranker.fused_tail_ok(K):
                if "rows" not in buffers or tuple(buffers["rows"].shape) != (b, K) or tuple(buffers["ids"].shape) != (b, keep):
                    buffers["rows"] = torch.empty((b, K), dtype=torch.int64, device=dev)
                    buffers["scores"] = torch.empty((b, K), dtype=torch.float32, device=dev)
                    buffers["ids"] = torch.empty((b, keep), dtype=torch.int64, device=dev)
                    buffers["fin"] = torch.empty((b, keep), dtype=torch.float64, device=dev)
                eng.search(queries, K, algo=algo, out_rows=buffers["rows"], out_scores=buffers["scores"])
                return reranker.rerank_scored_device(buffers["rows"], buffers["scores"], q_rec, topk,
                                                     out=(buffers["ids"], buffers["fin"]))
            rows, scores = eng.search(queries, K, algo=algo)
            if by_score and reranker.fused_tail_ok(K):
                return reranker.rerank_scored_device(rows, scores, q_rec, topk)
            if by_score:
                order, sc = reranker.rerank_with_cos_device(scores, q_rec, rows, topk, counts=_valid_counts(rows))
            else:
                order, sc = reranker.rerank_device(eng, queries, rows, q_rec, rows, topk, counts=_valid_counts(rows))
            ids = torch.empty((b, keep), dtype=torch.int64, device=dev)
            fin = torch.empty((b, keep), dtype=torch.float64, device=dev)
            with torch.cuda.device(d):
                _lib.check(lib.mmr_apply_order(_lib.ptr(rows), _lib.ptr(order), _lib.ptr(sc), b, K, keep, _lib.ptr(ids),
                                               _lib.ptr(fin), d, _lib.current_stream(d)))
            return ids, fin
        rank = dist.get_rank(self.group)
        per = (b + self.world - 1) // self.world
        q_lo, q_hi = min(b, rank * per), min(b, (rank + 1) * per)
        px = self._peer_exchange(b, K, queries, reranker) if by_score else None
        if px is not None:
            # ---- NVLink peer-memory path: no collective call between search and result ----
            import ctypes as C
            px.step += 1
            step = px.step
            q = queries.detach()
            qd = _lib.MMR_BF16 if q.dtype == torch.bfloat16 else _lib.MMR_F32
            if qd == _lib.MMR_F32:
                q = q.float()
            q = q.contiguous()
            a = _lib.ALGOS[algo if algo is not None else eng.algo]
            with torch.cuda.device(d):
                st = _lib.current_stream(d)
                _lib.check(lib.mmr_search_scatter(eng._handle, px._h, _lib.ptr(q), b, qd, K, a, step, st))
                p_ids, p_fin = C.c_void_p(), C.c_void_p()
                _lib.check(lib.mmr_exchange_rerank(px._h, reranker._tables, _lib.ptr(q_rec), b, K, reranker.alpha,
                                                   reranker.beta, reranker.gamma, int(topk), step, C.byref(p_ids),
                                                   C.byref(p_fin), st))
            # the full (b, keep) result lives in this rank's exchange region until step + 2
            return (_lib.as_cuda_tensor(p_ids.value, (b, keep), torch.int64, d),
                    _lib.as_cuda_tensor(p_fin.value, (b, keep), torch.float64, d))
        # ---- NCCL path ----
        gblob = self._local_blob(reranker, queries, K, algo)
        if getattr(self, "_fin", None) is None or self._fin[0].shape != (self.world * per, keep):
            self._fin = (torch.empty((self.world * per, keep), dtype=torch.int64, device=dev),
                         torch.empty((self.world * per, keep), dtype=torch.float64, device=dev),
                         torch.full((per, keep), -1, dtype=torch.int64, device=dev),
                         torch.zeros((per, keep), dtype=torch.float64, device=dev))
        all_ids, all_fin, my_ids, my_fin = self._fin
        if q_hi > q_lo:
            out_r, out_s, cos = self._merge_slice(gblob, b, K, q_lo, q_hi)
            if by_score and reranker.fused_tail_ok(K):
                reranker.rerank_scored_device(out_r, cos, q_rec[q_lo:q_hi], topk,
                                              out=(my_ids[: q_hi - q_lo], my_fin[: q_hi - q_lo]))
            else:
                order, sc = reranker.rerank_with_cos_device(cos, q_rec[q_lo:q_hi], out_r, topk, counts=_valid_counts(out_r))
                with torch.cuda.device(d):
                    _lib.check(lib.mmr_apply_order(_lib.ptr(out_r), _lib.ptr(order), _lib.ptr(sc), q_hi - q_lo, K, keep,
                                                   _lib.ptr(my_ids), _lib.ptr(my_fin), d, _lib.current_stream(d)))
        dist.all_gather_into_tensor(all_ids, my_ids, group=self.group)
        dist.all_gather_into_tensor(all_fin, my_fin, group=self.group)
        return all_ids[:b], all_fin[:b]

    def serve(self, reranker, host_batches, K: int, q_rec, topk: int = 0, algo: Optional[str] = None,
              to_host: bool = True):
        """Serving loop over HOST query batches: a generator that takes pinned host tensors ``(B, D)``
        fp32 and yields, per batch, the host ``(ids (B, keep) int64, scores (B, keep) fp64)`` pair
        that ``retrieve(..., reranker=...)`` returns.  Three streams keep the copies off the compute
        path: the host->device copy of batch i + 1 and the device->host copy of batch i - 1 overlap
        the search of batch i (results are yielded one batch late; the yielded pinned buffers are
        reused two batches later).  ``to_host=False`` (ranks that do not hand results to a caller)
        skips the device->host copy and yields the device tensors."""
        import torch
        dev = torch.device("cuda", self.engine.device)
        s_cmp = torch.cuda.current_stream(dev)
        # streams, staging buffers (device and pinned host) and events live on the searcher: a second serve() call
        # allocates nothing (a pinned allocation costs milliseconds)
        st = getattr(self, "_serve_state", None)
        if st is None:
            st = self._serve_state = {"s_in": torch.cuda.Stream(dev), "s_out": torch.cuda.Stream(dev),
                                      "qd": [None, None], "hout": [None, None], "bufs": [{}, {}]}
        s_in, s_out, qd, hout, bufs = st["s_in"], st["s_out"], st["qd"], st["hout"], st["bufs"]
        torch.cuda.current_stream(dev).synchronize()    # buffers of an earlier serve() call are free
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_cmp = [torch.cuda.Event() for _ in range(2)]
        ev_out = [torch.cuda.Event(blocking=True) for _ in range(2)]
        pending = None
        for i, batch in enumerate(host_batches):
            slot = i & 1
            if qd[slot] is None or qd[slot].shape != batch.shape:
                # allocated FROM THE COPY STREAM's pool: a block of the compute stream's pool may still be
                # read by kernels in flight (temporaries of the previous batch that Python has already
                # released), and the copy stream does not wait for those
                with torch.cuda.stream(s_in):
                    qd[slot] = torch.empty(batch.shape, dtype=torch.float32, device=dev)
                qd[slot].record_stream(s_cmp)
            if i >= 2:
                s_in.wait_event(ev_cmp[slot])          # the search of batch i - 2 has consumed qd[slot]
            with torch.cuda.stream(s_in):
                qd[slot].copy_(batch, non_blocking=True)
                ev_in[slot].record(s_in)
            s_cmp.wait_event(ev_in[slot])
            if i >= 2:
                s_cmp.wait_event(ev_out[slot])         # the results of batch i - 2 have left bufs[slot]
            ids, fin = self.retrieve_reranked(reranker, qd[slot], K, q_rec, topk=topk, algo=algo, buffers=bufs[slot])
            ev_cmp[slot].record(s_cmp)
            s_out.wait_event(ev_cmp[slot])
            if to_host:
                if hout[slot] is None or hout[slot][0].shape != ids.shape:
                    hout[slot] = (torch.empty(ids.shape, dtype=ids.dtype).pin_memory(),
                                  torch.empty(fin.shape, dtype=fin.dtype).pin_memory())
                with torch.cuda.stream(s_out):
                    hout[slot][0].copy_(ids, non_blocking=True)
                    hout[slot][1].copy_(fin, non_blocking=True)
                for t in (ids, fin):
                    try:
                        t.record_stream(s_out)
                    except Exception:
                        pass                            # views of the exchange region are not allocator-owned
                res = hout[slot]
            else:
                res = (ids, fin)
            ev_out[slot].record(s_out)
            if pending is not None:
                ev_out[pending[0]].synchronize()        # at most one batch in flight behind the caller
                yield pending[1]
            pending = (slot, res)
        if pending is not None:
            ev_out[pending[0]].synchronize()
            yield pending[1]

    def _peer_exchange(self, b: int, K: int, queries, reranker=None):
        """The NVLink peer-memory exchange for (b, K), created on first use (collective: every rank
        calls with the same sizes).  ``use_peer=False`` or MMR_B200_NO_PEER=1 selects the NCCL path; shapes
        the fused kernels do not cover (K > 128, more than 16 ranks, KG dimension > 512 or not a multiple of
        4) use NCCL as well."""
        import os
        if not self.use_peer or os.environ.get("MMR_B200_NO_PEER") == "1" or self.world > 16:
            return None
        if K > 128 or (reranker is not None and not reranker.fused_tail_ok(K)):
            return None
        px = getattr(self, "_px", None)
        if px is None or px.b_max < b or px.k_max < K:
            import sys
            import torch
            import torch.distributed as dist
            if px is not None:
                px.close(group=self.group)          # views returned by earlier calls are invalid from here on
            b_max, k_max = max(b, px.b_max if px else 0), max(K, px.k_max if px else 0)
            # creating the regions and mapping the peers is collective; if ANY rank cannot map its peers
            # (no P2P between the devices, IPC disabled) every rank falls back to the NCCL transport together
            px = PeerExchange(queries.device.index or 0, b_max, k_max, group=self.group)
            ok, err = (1, "") if px.error is None else (0, px.error)
            flag = torch.tensor([ok], device=queries.device, dtype=torch.int32)
            # (also the barrier: every rank has mapped every region before anybody stores into one)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
            if int(flag.item()) == 0:
                px.close(group=self.group)
                print(f"[mmr_b200] NVLink peer exchange unavailable ({err or 'a peer rank failed'}); "
                      "using the NCCL all-gather transport", file=sys.stderr)
                self.use_peer = False
                self._px = None
                return None
            self._px = px
        return px

    def close(self):
        """Release the peer exchange (collective at world > 1: unmap, barrier, free)."""
        px, self._px = getattr(self, "_px", None), None
        if px is not None:
            px.close(group=self.group)

    def rerank(self, reranker, q_embs, rows, q_rec, cand_rec, topk: int = 0):
        """Rerank merged global candidates: the label/KG tables are replicated, the candidate
        embedding cosine is computed by the rank that owns the row and summed across ranks
        (exactly one owner per candidate), then every rank combines."""
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return reranker.rerank_device(self.engine, q_embs, rows, q_rec, cand_rec, topk)
        raw = reranker.features_device(self.engine, q_embs, rows, q_rec, cand_rec)
        emb = raw[..., 0].contiguous()
        dist.all_reduce(emb, op=dist.ReduceOp.SUM, group=self.group)
        raw[..., 0] = emb
        return reranker.combine_device(raw, topk)
