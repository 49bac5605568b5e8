#!/usr/bin/env python
"""Benchmark of the retrieval hot path: exact cosine top-K search + label/KG rerank.

    python bench.py --gpus N --steps K --warmup W            # this repo (B200 kernels)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path on the host cores

Metric (BASELINE.json): queries/s for top-100 cosine + rerank over a 10M x 512 bf16 gallery.
One "step" = one batch of 4096 queries through search -> top-K -> rerank.  At N > 1 the gallery is
row-sharded over the ranks (strong scaling: total rows fixed), each rank searches its shard, the
per-rank lists are exchanged over NVLink peer memory, merged and reranked (split by query across the
ranks) and every rank ends with the full result.  Prints ONE JSON line on rank 0 with, besides the
contract's keys:

  roofline       dominant kernel (asked from the library: mmr_index_last_plan), live CUDA-event time,
                 achieved vs MEASURED_PEAKS.json, per-rank min/max at N > 1
  parity_check   computed OUTSIDE the timed region, at every N: 64 sampled queries of the timed batch vs
                 the brute-force oracle (oracle/bruteforce.py over the gathered shards), the reranked
                 result vs the oracle restatement of Reranker.rerank, NVLink-exchange result == NCCL
                 result, identical results on all ranks
  batch1_top10   (N = 1) p50 latency of single-query top-10 (HBM-scan regime, BASELINE cfg3)
  cfg2, cfg5     (N = 1) BASELINE configs[1] and [4]: 1M x 512 / batch 1024, and 10k queries x 1M with the
                 metrics computed on the device, each with its own parity check
  cfg4           (N >= 2) BASELINE configs[3]: 100M x 512 row-sharded, batch 4096, top-100
  cpu_baseline   (N = 1) the reference's CPU path on a bounded sample, host cores stated
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

if "reference" in sys.argv:
    # the reference arm is a CPU job: give BLAS every host core at every N (torchrun exports
    # OMP_NUM_THREADS=1 to its workers) -- must happen before numpy is imported
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

METRIC = "queries/sec top-100 cosine+rerank, 10M x 512 bf16 gallery"
UNIT = "queries/s"
SEED = 2709
CHUNK = 1 << 20  # rows per generation chunk (global alignment => shards see the same data for any N)
PARITY_QUERIES = 64


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=512)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--kg-dim", type=int, default=300)
    ap.add_argument("--algo", default="auto")
    ap.add_argument("--no-rerank", action="store_true")
    ap.add_argument("--latency-queries", type=int, default=200)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the (untimed) parity checks")
    ap.add_argument("--no-extra", action="store_true", help="headline only: skip cfg2 / cfg5 / cfg4")
    ap.add_argument("--cfg4-rows", type=int, default=100_000_000)
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks sampling (pynvml) during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake", 0x100: "display_clock"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons - {"gpu_idle"})}


# ------------------------------------------------------------------------------------------------
# synthetic data (device side; counter-style: chunk c is generated from seed SEED + c)
# ------------------------------------------------------------------------------------------------
def gen_rows(lo, hi, width, seed_base, device, dtype, normalize=False):
    import torch
    out = torch.empty((hi - lo, width), dtype=dtype, device=device)
    c0 = lo // CHUNK
    c1 = (hi - 1) // CHUNK if hi > lo else c0 - 1
    for c in range(c0, c1 + 1):
        g = torch.Generator(device=device)
        g.manual_seed(seed_base + c)
        blk = torch.randn((CHUNK, width), generator=g, device=device, dtype=torch.float32)
        if normalize:
            blk = blk / (blk.norm(dim=1, keepdim=True) + 1e-12)
        a, b_ = max(lo, c * CHUNK), min(hi, (c + 1) * CHUNK)
        out[a - lo:b_ - lo] = blk[a - c * CHUNK:b_ - c * CHUNK].to(dtype)
        del blk
    return out


def gen_masks(lo, hi, device, n_labels=43, p=0.08):
    import torch
    out = torch.empty((hi - lo,), dtype=torch.int64, device=device)
    c0 = lo // CHUNK
    c1 = (hi - 1) // CHUNK if hi > lo else c0 - 1
    w = (1 << torch.arange(n_labels, device=device, dtype=torch.int64))
    for c in range(c0, c1 + 1):
        g = torch.Generator(device=device)
        g.manual_seed(SEED + 500_000 + c)
        bits = (torch.rand((CHUNK, n_labels), generator=g, device=device) < p).to(torch.int64)
        m = (bits * w).sum(dim=1)
        a, b_ = max(lo, c * CHUNK), min(hi, (c + 1) * CHUNK)
        out[a - lo:b_ - lo] = m[a - c * CHUNK:b_ - c * CHUNK]
    return out


def gen_queries(b, dim, device, seed=SEED + 900_000):
    """bf16-rounded values held as fp32 (the dataset of the bf16 configs IS the rounded data)."""
    import torch
    gq = torch.Generator(device=device)
    gq.manual_seed(seed)
    return torch.randn((b, dim), generator=gq, device=device).to(torch.bfloat16).float()


# ------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the reference's exact path on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(args, steps, warmup):
    """The reference's CPU retrieval: ``cosine_similarity(Q, G)`` (scikit-learn, the very call of
    Evaluate/retrieval_overlap.py:85 and Retrieval/retrieval.py:128) + ``np.argsort(row)[::-1]`` per query
    (:90) + ``Reranker.rerank`` (Retrieval/reranker.py:240-333) on the top-k.  ``kind == "reference"``:
    the Reranker is the reference's own class, loaded unmodified from /root/reference or from the archive
    oracle/stage_ref.py packs at build time (oracle/_ref travels with the snapshot); ``kind == "port"``
    only if neither exists (oracle restatement).  Bounded sample: 1/10 of the gallery rows x 64 queries
    per step for the search (exact search cost is linear in rows => scaled x10, stated in `sample`) and
    the rerank of k-candidate lists timed on 4 queries per step."""
    import tempfile
    import numpy as np
    from multi_modal_retrieval_predict_project_b200 import synth
    from oracle import ref_loader
    from oracle import search as osr
    ncpu = os.cpu_count() or 1
    cores = ncpu
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        threadpool_limits(limits=ncpu)                              # explicit, whatever the launcher exported
        cores = max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        pass
    frac = 10
    n_s = max(1000, args.rows // frac)
    bq = 64 if steps + warmup <= 30 else 32
    rng = np.random.default_rng(SEED)
    g = osr.to_bf16_round(rng.standard_normal((n_s, args.dim), dtype=np.float32))
    q = osr.to_bf16_round(rng.standard_normal((bq, args.dim), dtype=np.float32))
    # rerank inputs for the sample (labels CSV + KG dir exactly as the reference reads them)
    tmp = tempfile.mkdtemp(prefix="mmr_cpu_")
    n_rec = 4000
    ids = synth.make_ids(n_rec)
    qids = synth.make_ids(8, "t")
    labels = synth.make_labels(n_rec + 8)
    csv = synth.write_labels_csv(os.path.join(tmp, "labels.csv"), ids + qids, labels)
    kg_dir = synth.write_kg(os.path.join(tmp, "kg"), ids + qids, synth.label_names(), d_kg=args.kg_dim)
    kind = "port"
    cos_fn = osr.cosine_similarity
    try:
        ref = ref_loader.load_reference(allow_staged=True)
        from sklearn.metrics.pairwise import cosine_similarity as cos_fn   # retrieval_overlap.py:15
        from pathlib import Path
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            rer = ref.Reranker(kg_dir=Path(kg_dir), labels_csv=Path(csv), preload_record_kg=False)
        kind = "reference"
    except Exception as e:  # noqa: BLE001
        print(f"[bench] reference files unavailable ({e}); timing the oracle port", file=sys.stderr)
        from oracle import rerank as orr
        rer = orr.OracleReranker(kg_dir, csv)

    def one_step():
        t0 = time.perf_counter()
        sim = cos_fn(q, g)                                              # retrieval_overlap.py:85
        top = [np.argsort(sim[i])[::-1][:args.k] for i in range(bq)]    # :90
        t_search = time.perf_counter() - t0
        t_rr = 0.0
        if not args.no_rerank:
            t1 = time.perf_counter()
            for i in range(4):                                          # 4 queries x k candidates
                cand = [int(j) % n_rec for j in top[i]]
                rer.rerank(qids[i], [ids[j] for j in cand], candidate_embs=g[top[i]], query_emb=q[i], topk=args.k)
            t_rr = (time.perf_counter() - t1) / 4
        # seconds per query on the FULL gallery: search scales with rows, rerank does not
        per_query = (t_search / bq) * (args.rows / n_s) + t_rr
        return per_query, t_search, t_rr

    for _ in range(max(1, warmup)):
        one_step()
    t_all = time.perf_counter()
    per = [one_step() for _ in range(max(1, steps))]
    step_wall_ms = 1e3 * (time.perf_counter() - t_all) / max(1, steps)
    pq = float(np.mean([p[0] for p in per]))
    sample = (f"per step: {bq} queries x {n_s}-row slice (1/{frac} of the gallery; search time scaled x{args.rows / n_s:.0f}) "
              f"through sklearn cosine_similarity + np.argsort, + {'the reference Reranker.rerank' if kind == 'reference' else 'the oracle port of Reranker.rerank'} "
              f"of {args.k} candidates timed on 4 queries; mean over {max(1, steps)} steps; BLAS threads={cores} of {ncpu} host cores")
    return {"value": 1.0 / pq, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
            "sample_step_wall_ms": step_wall_ms, "sample_queries_per_step": bq, "sample_rows": n_s,
            "search_s_per_query_full": float(np.mean([p[1] for p in per])) / bq * (args.rows / n_s),
            "rerank_s_per_query": float(np.mean([p[2] for p in per]))}


# ------------------------------------------------------------------------------------------------
# parity checks (untimed; oracle/ is used here only as the checker)
# ------------------------------------------------------------------------------------------------
def sample_queries(b, n=PARITY_QUERIES):
    """n query indices spread over the batch: every query tile and every lane position is hit."""
    import torch
    n = min(n, b)
    step = max(1, b // n)
    idx = torch.arange(n) * step + (torch.arange(n) % step)
    return idx.clamp_(max=b - 1)


def check_search(rows, scores, gallery, lo, q_sel, k, world, dist):
    """rows/scores (n_sel, k) GLOBAL top-k of the sampled queries (torch, this rank) vs the brute-force oracle
    over ALL shards: every rank brute-forces its own rows (fp32 and fp64), rank 0 merges and compares."""
    import numpy as np
    import torch
    from oracle import search as osr
    from oracle.bruteforce import bruteforce_topk
    r32, s32 = bruteforce_topk(gallery, q_sel, k, row_offset=lo)
    r64, s64 = bruteforce_topk(gallery, q_sel, k, row_offset=lo, chunk=1 << 18, dtype=torch.float64)
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, (r32, s32, r64, s64))
        r32, s32, r64, s64 = (np.concatenate([p[i] for p in parts], axis=1) for i in range(4))
    gr, gs = rows.cpu().numpy(), scores.cpu().numpy()
    exact = 0
    for i in range(gr.shape[0]):
        o = np.lexsort((r32[i], -s32[i].astype(np.float64)))[:k]
        ok, why = osr.topk_matches(gr[i], gs[i], r32[i, o], s32[i, o], rtol=2e-5, atol=1e-6)
        if not ok:
            return False, f"search, query {i}: {why}", 0
        exact += int(np.array_equal(gr[i], r32[i, o]))
        o64 = np.lexsort((r64[i], -s64[i]))[:k]
        kth = s64[i, o64[-1]]
        for j in o64:
            if r64[i, j] not in gr[i] and s64[i, j] - kth > 1e-5:
                return False, f"search, query {i}: row {int(r64[i, j])} of the fp64 top-{k} is missing (not a boundary tie)", 0
    return True, "ok", exact


def check_rerank(ids, fin, rows, q_sel, q_rec_sel, cand_emb, masks, kg, k, weights):
    """ids/fin (n_sel, keep): the device's reranked result; rows (n_sel, k) the candidates it was given;
    cand_emb (n_sel, k, d) their stored embeddings -> oracle restatement of Reranker.rerank's scoring."""
    from oracle import rerank as orr
    q = q_sel.cpu().numpy()
    ce = cand_emb.float().cpu().numpy()
    rows_h = rows.cpu().numpy()
    m_c = masks[rows].cpu().numpy()
    m_q = masks[q_rec_sel].cpu().numpy()
    kg_c = kg[rows.reshape(-1)].view(rows.shape[0], rows.shape[1], -1).cpu().numpy()
    kg_q = kg[q_rec_sel].cpu().numpy()
    ids_h, fin_h = ids.cpu().numpy(), fin.cpu().numpy()
    for i in range(rows_h.shape[0]):
        want = orr.rerank_from_arrays(q[i], ce[i], m_q[i], m_c[i], kg_q[i], kg_c[i], *weights, topk=ids_h.shape[1])
        ok, why = orr.reranked_lists_match(ids_h[i], fin_h[i], rows_h[i], want)
        if not ok:
            return False, f"rerank, query {i}: {why}"
    return True, "ok"


def main():
    args = parse_args()
    # stdout carries exactly ONE JSON line: everything else that libraries print there (e.g. NCCL's
    # version banner) is routed to stderr for the duration of the run
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(json_fd, (json.dumps(line) + "\n").encode())

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": f"{args.rows}x{args.dim} bf16 gallery, query batch {args.batch}, top-{args.k}"
                          + ("" if args.no_rerank else " + label/KG rerank (alpha,beta,gamma=0.6,0.25,0.15)"),
              "rows": args.rows, "dim": args.dim, "batch": args.batch, "k": args.k, "rerank": not args.no_rerank,
              "sharding": f"row-sharded x{world}" if world > 1 else "single shard",
              "result": "per query the reranked top-k row ids (int64) + combined scores (fp64); e2e = serving loop (ShardedSearcher.serve): every step copies its queries from pinned host memory and its results to the host on rank 0, copies overlap the neighbouring steps' search",
              "l2_policy": "inputs larger than L2: the gallery shard (>= 1.28 GB) is streamed from HBM every step"}

    if args.impl == "reference":
        if rank != 0:
            return
        t0 = time.perf_counter()
        cb = cpu_reference_run(args, args.steps, args.warmup)
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
                # a step of this arm = one bounded sample (cpu_baseline.sample); `value` is the extrapolated full-
                # gallery rate, ms_per_step the measured wall time of a sample step
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["sample_step_wall_ms"],
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config, "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0, "wall_s": time.perf_counter() - t0}
        emit(line)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    from multi_modal_retrieval_predict_project_b200 import B200RetrievalEngine, Reranker, _lib
    from multi_modal_retrieval_predict_project_b200.sharded import ShardedSearcher, shard_bounds

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_tf_burst = peaks.get("bf16_tflops", 1590.0)
    peak_hbm = peaks.get("hbm_gbs", 6650.0)
    weights = (0.6, 0.25, 0.15)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- data: this rank's gallery shard + replicated rerank tables ----------------
    lo, hi = shard_bounds(args.rows, world, rank)
    gallery = gen_rows(lo, hi, args.dim, SEED, dev, torch.bfloat16)
    engine = B200RetrievalEngine.from_arrays(gallery, dtype="bfloat16", device=local_rank, row_offset=lo, borrow=True,
                                             keep_host=False, algo=args.algo)
    searcher = ShardedSearcher(engine)
    b, k = args.batch, args.k
    q_dev = gen_queries(b, args.dim, dev)
    q_host = q_dev.cpu().pin_memory()
    reranker = masks = kg = q_rec = None
    if not args.no_rerank:
        n_rec = args.rows + b  # records: gallery rows then the query records
        masks = gen_masks(0, n_rec, dev)
        kg = gen_rows(0, n_rec, args.kg_dim, SEED + 700_000, dev, torch.float32, normalize=True)
        reranker = Reranker.from_tables(masks, kg, alpha=weights[0], beta=weights[1], gamma=weights[2], device=local_rank)
        q_rec = torch.arange(args.rows, args.rows + b, device=dev, dtype=torch.int64)
    torch.cuda.synchronize()

    def step(q):
        if reranker is None:
            return searcher.search(q, k)
        # what the reference's retrieve(..., reranker=...) returns: reranked ids + combined scores
        # (record index == global row); at N > 1 the merge + rerank are split across ranks by query
        return searcher.retrieve_reranked(reranker, q, k, q_rec, topk=k)

    for _ in range(max(args.warmup, 3)):
        out = step(q_dev)
    barrier()

    # ---------------- timed region: device-resident inputs -----------------------------------
    sampler = ClockSampler(local_rank)
    engine.profile(True)
    launches0 = lib.mmr_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    barrier()
    ev0.record()
    for _ in range(args.steps):
        out = step(q_dev)
    ev1.record()
    barrier()
    clocks = sampler.stop()
    launches = lib.mmr_launch_count() - launches0
    kern_ms, kern_n = engine.profile(False)
    plan = engine.last_plan()
    my_ms = ev0.elapsed_time(ev1)
    elapsed_ms = max_over_ranks(my_ms)
    value = b * args.steps / (elapsed_ms / 1e3)
    timed_out = tuple(t.clone() for t in out)      # the result of the last timed step (checked below)

    # ---------------- end to end: pinned host queries in, host results out, every step ---------
    # through the serving API (ShardedSearcher.serve): every step copies ITS queries from pinned host
    # memory and ITS results back to the host; copies of neighbouring steps overlap the search
    out_shapes = [(tuple(r.shape), r.dtype) for r in out]

    def host_batches(n):
        for _ in range(n):
            yield q_host                                 # this step's queries (pinned host memory)

    e2e_gaps = []

    def e2e_run(n):
        got = 0
        t_prev = time.perf_counter()
        for res in searcher.serve(reranker, host_batches(n), k, q_rec, topk=k, to_host=(rank == 0)):
            got += 1                                     # rank 0: res = host (ids, scores) of one step
            t_now = time.perf_counter()
            e2e_gaps.append(t_now - t_prev)              # host time between consecutive results
            t_prev = t_now
        assert got == n

    e2e_kern = None
    if reranker is not None:
        e2e_run(2)                                       # untimed warm-up of the copy path
        barrier()
        engine.profile(True)
        t0 = time.perf_counter()
        e2e_run(args.steps)
        barrier()
        e2e_s = time.perf_counter() - t0
        e2e_kern = engine.profile(False)
    else:
        out_host = [torch.empty(shp, dtype=dt).pin_memory() for shp, dt in out_shapes]
        qd = torch.empty_like(q_dev)
        done = torch.cuda.Event(blocking=True)

        def e2e_step():
            qd.copy_(q_host, non_blocking=True)
            res = step(qd)
            if rank == 0:
                for h, r in zip(out_host, res):
                    h.copy_(r, non_blocking=True)
            done.record()
            done.synchronize()

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        e2e_s = time.perf_counter() - t0
    d2h = sum(int(np.prod(shp)) * torch.empty((), dtype=dt).element_size() for shp, dt in out_shapes)
    e2e = {"value": b * args.steps / max_over_ranks(e2e_s), "unit": UNIT,
           "h2d_bytes_per_step": q_host.numel() * q_host.element_size(), "d2h_bytes_per_step": d2h}
    if e2e_kern is not None and e2e_kern[1] > 0:
        e2e["search_kernel_ms"] = e2e_kern[0] / e2e_kern[1]      # the dominant kernel inside the serving loop
    if len(e2e_gaps) >= args.steps > 2:
        g_ms = sorted(1e3 * x for x in e2e_gaps[-args.steps + 2:])      # the timed call, without its first results
        e2e["result_interval_ms"] = {"p50": g_ms[len(g_ms) // 2], "max": g_ms[-1]}

    # ---------------- roofline of the dominant kernel (which one: asked from the library) --------
    n_local = hi - lo
    kern_avg_ms = kern_ms / max(kern_n, 1)
    per_rank = torch.tensor([kern_avg_ms, my_ms / args.steps, float(clocks.get("sm_mhz") or 0),
                             float("sw_power_cap" in clocks.get("reasons", []))], device=dev, dtype=torch.float64)
    all_ranks = [torch.zeros_like(per_rank) for _ in range(world)]
    if world > 1:
        dist.all_gather(all_ranks, per_rank)
    else:
        all_ranks = [per_rank]
    kern_by_rank = [float(t[0]) for t in all_ranks]
    step_by_rank = [float(t[1]) for t in all_ranks]
    traffic_tbl = {}
    try:
        traffic_tbl = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        pass

    def traffic_for(kernel, rows, batch, kk):
        for t in traffic_tbl.get(kernel) or []:      # one entry per captured shape
            if (t["rows"], t["dim"], t["batch"], t["k"]) == (rows, args.dim, batch, kk):
                return t["bytes"]
        return None

    def gemm_roofline(flops, ms, pl, rows, batch, kk, in_long_step=True):
        ach = flops / (ms / 1e3) / 1e12 if ms > 0 else 0.0
        name = (f"gemm_topk_kernel<resident={'1' if args.dim <= 512 else '0'}, pair={int(pl['pair'])}, "
                f"probe={int(pl['variant'] == 'short')}> (tcgen05, {pl['parts']} parts x {pl['tiles_per_part']} tiles)")
        pk = peak_tf if in_long_step else peak_tf_burst
        return {"kernel": name, "bound": "tensor", "achieved": ach, "peak": pk, "unit": "TFLOP/s", "frac": ach / pk,
                "traffic": traffic_for("gemm_topk_kernel", rows, batch, kk),
                "peak_source": ("MEASURED_PEAKS.json " + ("bf16_tflops_sustained (kernel timed inside a long step)"
                                                         if in_long_step else "bf16_tflops (burst: short launches, no power cap)"))
                if peaks else "fallback", "peak_burst": peak_tf_burst, "frac_of_burst": ach / peak_tf_burst,
                "kernel_ms": ms, "algorithmic_flops_per_launch": flops}

    if plan["algo"] == "gemm":
        roofline = gemm_roofline(2.0 * b * n_local * args.dim, kern_avg_ms, plan, n_local, b, k)
    else:
        byts = n_local * args.dim * 2.0 + 4.0 * n_local
        groups = -(-b // 4)
        ach = byts * groups / (kern_avg_ms / 1e3) / 1e9 if kern_avg_ms > 0 else 0.0
        roofline = {"kernel": "scan_topk_kernel", "bound": "hbm", "achieved": ach, "peak": peak_hbm, "unit": "GB/s",
                    "frac": ach / peak_hbm, "traffic": None, "kernel_ms": kern_avg_ms,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback"}
    roofline["kernel_share_of_step"] = kern_ms / my_ms if my_ms else None
    if world > 1:
        roofline["kernel_ms_by_rank"] = {"min": min(kern_by_rank), "max": max(kern_by_rank)}
        roofline["step_ms_by_rank"] = {"min": min(step_by_rank), "max": max(step_by_rank)}
        roofline["outside_kernel_ms"] = elapsed_ms / args.steps - max(kern_by_rank)
        # rank skew of the same-sized shard launches is a clock difference between the GPUs when these differ
        roofline["by_rank"] = [{"rank": r, "kernel_ms": float(t[0]), "sm_mhz": int(t[2]), "sw_power_cap": bool(t[3])}
                               for r, t in enumerate(all_ranks)]

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config,
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline}

    # ---------------- parity of what was timed (untimed; every N) ---------------------------------
    if not args.no_parity:
        t0 = time.perf_counter()
        pc = {"queries": int(min(PARITY_QUERIES, b)), "ok": False}
        try:
            sel = sample_queries(b).to(dev)
            q_sel = q_dev[sel].contiguous()
            rows_all, scores_all = searcher.search(q_dev, k)                 # global top-k (NCCL all-gather + merge at N > 1)
            ok, why, exact = check_search(rows_all[sel], scores_all[sel], gallery, lo, q_sel, k, world, dist)
            pc.update({"search_vs_bruteforce_oracle": ok, "identical_id_lists": exact})
            detail = [] if ok else [why]
            if reranker is not None:
                ids_t, fin_t = timed_out
                # (a) the timed transport (NVLink peer exchange at N > 1) == the NCCL transport, bit for bit
                same = True
                if world > 1:
                    ids_n, fin_n = ShardedSearcher(engine, use_peer=False).retrieve_reranked(reranker, q_dev, k, q_rec, topk=k)
                    same = bool(torch.equal(ids_n, ids_t)) and bool(torch.equal(fin_n, fin_t))
                    pc["peer_exchange_equals_nccl"] = same
                    if not same:
                        detail.append("peer-exchange result differs from the NCCL all-gather result")
                    # (b) every rank holds the same (ids, scores): checksum min == max over ranks
                    cs = torch.stack([ids_t.sum(), fin_t.view(torch.int64).sum()]).to(torch.int64)
                    cmin, cmax = cs.clone(), cs.clone()
                    dist.all_reduce(cmin, op=dist.ReduceOp.MIN)
                    dist.all_reduce(cmax, op=dist.ReduceOp.MAX)
                    pc["all_ranks_identical"] = bool(torch.equal(cmin, cmax))
                    same = same and pc["all_ranks_identical"]
                    if not pc["all_ranks_identical"]:
                        detail.append("ranks hold different results")
                # (c) the reranked result vs the oracle restatement of Reranker.rerank's scoring, fed the
                # candidates the search returned and THEIR stored embeddings (gathered from the owning shards)
                cand = rows_all[sel]
                loc = cand - lo
                own = (loc >= 0) & (loc < n_local)
                ce = torch.zeros((cand.shape[0], k, args.dim), dtype=torch.float32, device=dev)
                ce[own] = gallery[loc[own]].float()
                if world > 1:
                    dist.all_reduce(ce, op=dist.ReduceOp.SUM)
                ok_r, why_r = check_rerank(ids_t[sel], fin_t[sel], cand, q_sel, q_rec[sel], ce, masks, kg, k, weights)
                pc["rerank_vs_oracle"] = ok_r
                if not ok_r:
                    detail.append(why_r)
                ok = ok and ok_r and same
            flag = torch.tensor([1 if ok else 0], device=dev)
            if world > 1:
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            pc["ok"] = bool(flag.item())
            if detail:
                pc["detail"] = detail
        except Exception as e:  # noqa: BLE001
            pc["detail"] = [f"{type(e).__name__}: {e}"]
            if world > 1:
                raise
        pc["seconds"] = round(time.perf_counter() - t0, 2)
        pc["how"] = ("64 queries of the timed batch spread over all query tiles: search ids/scores vs fp32 brute force "
                     "over the same bf16 values (oracle.search.topk_matches, rtol 2e-5) + fp64 recall with eps 1e-5; "
                     "reranked (ids, scores) of the last TIMED step vs oracle.rerank.rerank_from_arrays (1e-5)")
        line["parity_check"] = pc

    # ---------------- N = 1 extras: batch-1 latency, cfg2, cfg5, CPU baseline -----------------------
    if world == 1:
        nlat = args.latency_queries
        if nlat > 0:
            q1 = q_host[:1].numpy()
            for _ in range(5):
                engine.search(q1, 10, algo="scan")
            engine.profile(True)
            lat = []
            for i in range(nlat):
                qi = q_host[i % b:i % b + 1].numpy()
                t0 = time.perf_counter()
                engine.search(qi, 10, algo="scan")       # host in, host out: launch + H2D + D2H included
                lat.append(time.perf_counter() - t0)
            sms, sn = engine.profile(False)
            lat.sort()
            scan_ms = sms / max(sn, 1)
            byts = n_local * args.dim * 2.0 + 4.0 * n_local
            b1 = {"p50_ms": 1e3 * lat[len(lat) // 2], "p99_ms": 1e3 * lat[int(len(lat) * 0.99)],
                  "queries": nlat, "scan_kernel_ms": scan_ms,
                  "roofline": {"bound": "hbm", "achieved": byts / (scan_ms / 1e3) / 1e9,
                               "peak": peak_hbm, "unit": "GB/s",
                               "frac": byts / (scan_ms / 1e3) / 1e9 / peak_hbm,
                               "traffic": traffic_for("scan_topk_kernel", n_local, 1, 10),
                               "algorithmic_bytes_per_launch": byts}}
            if not args.no_parity:
                from oracle.bruteforce import check_topk
                r1, s1 = engine.search(q_dev[:8], 10, algo="scan")
                ok, why = check_topk(r1, s1, gallery, q_dev[:8], 10)
                b1["parity_check"] = {"queries": 8, "ok": bool(ok), "detail": why}
            line["batch1_top10"] = b1
        if not args.no_extra and args.rows >= 1_000_000:
            del engine, searcher
            line["cfg2"] = run_cfg2(args, dev, gallery, masks, kg, weights, peak_tf_burst, peaks, gemm_roofline)
            line["cfg5"] = run_cfg5(args, dev, gallery, peak_tf_burst)
        if rank == 0 and not args.no_cpu_baseline:
            del gallery
            line["cpu_baseline"] = cpu_reference_run(args, 2, 1)
    elif not args.no_extra and args.cfg4_rows > 0:
        # ---------------- N >= 2: BASELINE configs[3] (100M x 512 row-sharded, batch 4096, top-100) -------
        searcher.close()                                   # collective: unmap the peers, barrier, free
        del engine, searcher, reranker, masks, kg, gallery
        torch.cuda.empty_cache()
        line["cfg4"] = run_cfg4(args, dev, rank, world, local_rank, dist, peak_tf, gemm_roofline)
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def timed_loop(fn, steps, warmup):
    """CUDA-event time of `steps` calls after `warmup` (device-resident inputs), ms per call."""
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, out


def run_cfg2(args, dev, gallery, masks, kg, weights, peak_burst, peaks, gemm_roofline):
    """BASELINE configs[1]: 1M x 512 bf16 gallery (the first 1M rows of the synthetic gallery), query batch
    1024, top-100 (+ rerank), one GPU -- the tensor-core GEMM regime on a short launch."""
    import torch
    from multi_modal_retrieval_predict_project_b200 import B200RetrievalEngine, Reranker
    from multi_modal_retrieval_predict_project_b200.sharded import ShardedSearcher
    from oracle.bruteforce import check_topk
    n, b, k = 1_000_000, 1024, 100
    g = gallery[:n]
    eng = B200RetrievalEngine.from_arrays(g, dtype="bfloat16", device=dev.index, borrow=True, keep_host=False)
    s = ShardedSearcher(eng)
    q = gen_queries(b, args.dim, dev, seed=SEED + 900_001)
    out = {"workload": f"cfg2: {n}x{args.dim} bf16 gallery, query batch {b}, top-{k}"
                       + ("" if masks is None else " + label/KG rerank")}
    if masks is not None:
        rec = torch.cat([torch.arange(n, device=dev), torch.arange(args.rows, args.rows + b, device=dev)])
        rer = Reranker.from_tables(masks[rec].contiguous(), kg[rec].contiguous(), alpha=weights[0], beta=weights[1],
                                   gamma=weights[2], device=dev.index)
        q_rec = torch.arange(n, n + b, device=dev, dtype=torch.int64)
        fn = lambda: s.retrieve_reranked(rer, q, k, q_rec, topk=k)   # noqa: E731
    else:
        fn = lambda: s.search(q, k)                                  # noqa: E731
    eng.profile(True)
    ms, _ = timed_loop(fn, 20, 5)
    kms, kn = eng.profile(False)
    plan = eng.last_plan()
    out.update({"value": b / (ms / 1e3), "unit": UNIT, "ms_per_step": ms, "steps": 20, "warmup": 5,
                "roofline": gemm_roofline(2.0 * b * n * args.dim, kms / max(kn, 1), plan, n, b, k, in_long_step=False)})
    if not args.no_parity:
        sel = sample_queries(b).to(dev)
        rows, scores = eng.search(q, k)
        ok, why = check_topk(rows[sel], scores[sel], g, q[sel], k)
        out["parity_check"] = {"queries": int(sel.numel()), "ok": bool(ok), "detail": why}
    eng.close()
    return out


def run_cfg5(args, dev, gallery, peak_burst):
    """BASELINE configs[4]: on-device evaluation -- 10k queries x 1M gallery rows, top-100, then P@K / Recall@K /
    AP / RR / nDCG against synthetic relevance sets (CSR), all resident on the GPU."""
    import numpy as np
    import torch
    from multi_modal_retrieval_predict_project_b200 import B200RetrievalEngine
    from multi_modal_retrieval_predict_project_b200.Helpers import metrics_from_rows
    from oracle import metrics as om
    from oracle.bruteforce import check_topk
    n, nq, k, max_rel = 1_000_000, 10_000, 100, 200
    g = gallery[:n]
    eng = B200RetrievalEngine.from_arrays(g, dtype="bfloat16", device=dev.index, borrow=True, keep_host=False)
    q = gen_queries(nq, args.dim, dev, seed=SEED + 900_002)
    # synthetic relevance: per query a sorted unique set of U[1, max_rel] random gallery rows (seed 2709 + 1)
    # plus every 3rd of its true neighbours (phase i % 3), so that the metric values are not all zero
    rows0, _ = eng.search(q, k)
    rows0 = rows0.cpu().numpy()
    rng = np.random.default_rng(SEED + 1)
    sizes = rng.integers(1, max_rel + 1, size=nq)
    sets = [np.union1d(rng.choice(n, size=int(sizes[i]), replace=False), rows0[i, (i % 3)::3]) for i in range(nq)]
    indptr = np.zeros(nq + 1, dtype=np.int64)
    indptr[1:] = np.cumsum([len(x) for x in sets])
    rel = np.concatenate(sets).astype(np.int64)
    d_indptr, d_rel = torch.from_numpy(indptr).to(dev), torch.from_numpy(rel).to(dev)
    eng.profile(True)
    search_ms, (rows, scores) = timed_loop(lambda: eng.search(q, k), 5, 2)
    kms, kn = eng.profile(False)
    metrics_ms, tbl = timed_loop(lambda: metrics_from_rows(rows, d_indptr, d_rel, k), 5, 2)
    t = tbl.cpu().numpy()
    tf = 2.0 * nq * n * args.dim / (kms / max(kn, 1) / 1e3) / 1e12
    out = {"workload": f"cfg5: {nq} queries x {n}x{args.dim} bf16 gallery, top-{k} + P/R/AP/RR/nDCG on device "
                       f"(relevance: CSR of U[1,{max_rel}] sorted unique rows per query)",
           "search_ms": search_ms, "metrics_ms": metrics_ms, "value": nq / ((search_ms + metrics_ms) / 1e3), "unit": UNIT,
           "roofline": {"bound": "tensor", "achieved": tf, "peak": peak_burst, "unit": "TFLOP/s", "frac": tf / peak_burst,
                        "kernel_ms": kms / max(kn, 1), "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst)"},
           "P@k": float(np.mean(t[:, 0])), "R@k": float(np.mean(t[:, 1])), "mAP": float(np.mean(t[:, 2])),
           "MRR": float(np.mean(t[:, 3])), "nDCG": float(np.mean(t[:, 4]))}
    if not args.no_parity:
        nchk = PARITY_QUERIES
        r = rows[:nchk].cpu().numpy()
        rets = [[int(x) for x in r[i]] for i in range(nchk)]
        rels = [rel[indptr[i]:indptr[i + 1]].tolist() for i in range(nchk)]
        same = bool(np.array_equal(om.per_query_table(rets, rels, k), t[:nchk]))
        ok, why = check_topk(rows[:nchk], scores[:nchk], g, q[:nchk], k)
        out["parity_check"] = {"queries": nchk, "ok": bool(ok and same), "metrics_identical_to_oracle": same,
                               "search_vs_bruteforce_oracle": bool(ok), "detail": why}
    eng.close()
    return out


def run_cfg4(args, dev, rank, world, local_rank, dist, peak_tf, gemm_roofline):
    """BASELINE configs[3]: 100M x 512 bf16 gallery row-sharded over the ranks (generated per shard on the
    device), query batch 4096, top-100, per-rank search + NCCL all-gather of the lists + on-device merge."""
    import torch
    from multi_modal_retrieval_predict_project_b200 import B200RetrievalEngine
    from multi_modal_retrieval_predict_project_b200.sharded import ShardedSearcher, shard_bounds
    n, b, k = args.cfg4_rows, args.batch, args.k
    lo, hi = shard_bounds(n, world, rank)
    need = (hi - lo) * args.dim * 2 + (6 << 30)
    free, _ = torch.cuda.mem_get_info(dev)
    fits = torch.tensor([1 if free > need else 0], device=dev)
    dist.all_reduce(fits, op=dist.ReduceOp.MIN)
    if int(fits.item()) == 0:
        return {"skipped": f"a {hi - lo}-row shard needs {need >> 30} GiB, {free >> 30} GiB free"}
    g = gen_rows(lo, hi, args.dim, SEED, dev, torch.bfloat16)
    eng = B200RetrievalEngine.from_arrays(g, dtype="bfloat16", device=local_rank, row_offset=lo, borrow=True, keep_host=False)
    s = ShardedSearcher(eng)
    q = gen_queries(b, args.dim, dev)
    for _ in range(2):
        rows, scores = s.search(q, k)
    dist.barrier()
    torch.cuda.synchronize()
    eng.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 5
    e0.record()
    for _ in range(steps):
        rows, scores = s.search(q, k)
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    kms, kn = eng.profile(False)
    t = torch.tensor([e0.elapsed_time(e1) / steps, kms / max(kn, 1)], device=dev, dtype=torch.float64)
    tmax = t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax[0])
    plan = eng.last_plan()
    out = {"workload": f"cfg4: {n}x{args.dim} bf16 gallery row-sharded x{world} ({hi - lo} rows per GPU), query batch {b}, "
                       f"top-{k}, NCCL all-gather of the per-rank lists + on-device merge (no rerank)",
           "value": b / (ms / 1e3), "unit": UNIT, "ms_per_step": ms, "steps": steps, "warmup": 2, "scaling": "strong",
           "roofline": gemm_roofline(2.0 * b * (hi - lo) * args.dim, float(tmax[1]), plan, hi - lo, b, k)}
    if not args.no_parity:
        sel = sample_queries(b).to(dev)
        ok, why, exact = check_search(rows[sel], scores[sel], g, lo, q[sel].contiguous(), k, world, dist)
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        out["parity_check"] = {"queries": int(sel.numel()), "ok": bool(flag.item()), "detail": why,
                               "identical_id_lists": exact}
    eng.close()
    return out


if __name__ == "__main__":
    main()
