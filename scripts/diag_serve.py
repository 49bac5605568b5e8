"""Diagnostic: where the serving loop (ShardedSearcher.serve) spends its step on ONE GPU.

    python scripts/diag_serve.py [rows] [steps]

Modes: the device-resident loop (bench `value`), serve() with host batches and host results (bench `e2e`),
serve() without the device->host copy, serve() fed DEVICE batches (no PCIe traffic at all).  For each: wall ms
per step, the search kernel's CUDA-event average, and -- from events recorded around every retrieve_reranked call
on the compute stream -- the GPU time of a step and the idle gap between consecutive steps."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from multi_modal_retrieval_predict_project_b200 import B200RetrievalEngine, Reranker
from multi_modal_retrieval_predict_project_b200.sharded import ShardedSearcher

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
dim, b, k = 512, 4096, 100
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
g = bench.gen_rows(0, rows, dim, bench.SEED, dev, torch.bfloat16)
eng = B200RetrievalEngine.from_arrays(g, dtype="bfloat16", device=0, borrow=True, keep_host=False)
s = ShardedSearcher(eng)
q = bench.gen_queries(b, dim, dev)
qh = q.cpu().pin_memory()
masks = bench.gen_masks(0, rows + b, dev)
kg = bench.gen_rows(0, rows + b, 300, bench.SEED + 700000, dev, torch.float32, normalize=True)
rer = Reranker.from_tables(masks, kg, device=0)
q_rec = torch.arange(rows, rows + b, device=dev)

marks = []
inner = s.retrieve_reranked


def wrapped(*a, **kw):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = inner(*a, **kw)
    e1.record()
    marks.append((e0, e1))
    return out


s.retrieve_reranked = wrapped


def report(name, wall_s, n):
    torch.cuda.synchronize()
    kern_ms, kern_n = eng.profile(False)
    m = marks[-n:]
    busy = sorted(a.elapsed_time(b_) for a, b_ in m)
    gaps = sorted(m[i][1].elapsed_time(m[i + 1][0]) for i in range(len(m) - 1))
    span = m[0][0].elapsed_time(m[-1][1]) / n
    print(f"{name:34s} wall {wall_s / n * 1e3:7.3f} ms/step | gpu span {span:7.3f} | step on gpu p50 {busy[len(busy) // 2]:7.3f} "
          f"max {busy[-1]:7.3f} | gap p50 {gaps[len(gaps) // 2] * 1e3:6.1f} us max {gaps[-1] * 1e3:6.1f} us | "
          f"kernel {kern_ms / max(kern_n, 1):7.3f} ms", flush=True)


def run_value():
    for _ in range(5):
        s.retrieve_reranked(rer, q, k, q_rec, topk=k)
    torch.cuda.synchronize()
    eng.profile(True)
    t0 = time.perf_counter()
    for _ in range(steps):
        s.retrieve_reranked(rer, q, k, q_rec, topk=k)
    torch.cuda.synchronize()
    report("device-resident loop", time.perf_counter() - t0, steps)


def run_serve(name, batch, to_host):
    def batches(n):
        for _ in range(n):
            yield batch
    for _ in s.serve(rer, batches(3), k, q_rec, topk=k, to_host=to_host):
        pass
    torch.cuda.synchronize()
    eng.profile(True)
    t0 = time.perf_counter()
    for _ in s.serve(rer, batches(steps), k, q_rec, topk=k, to_host=to_host):
        pass
    torch.cuda.synchronize()
    report(name, time.perf_counter() - t0, steps)


for rep in range(2):
    run_value()
    run_serve("serve: host in, host out (e2e)", qh, True)
    run_serve("serve: host in, device out", qh, False)
    run_serve("serve: device in, host out", q, True)
    run_serve("serve: device in, device out", q, False)
