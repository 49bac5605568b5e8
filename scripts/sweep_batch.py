"""Regime sweep: per-batch search time of the scan and GEMM kernels over one gallery (device-resident
queries/results, CUDA events) -- the data behind the MMR_ALGO_AUTO threshold."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from multi_modal_retrieval_predict_project_b200 import B200RetrievalEngine

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=10_000_000)
ap.add_argument("--dim", type=int, default=512)
ap.add_argument("--k", type=int, nargs="+", default=[10, 100])
ap.add_argument("--batches", type=int, nargs="+", default=[1, 2, 4, 8, 16, 32, 64, 128, 256])
a = ap.parse_args()
dev = torch.device("cuda", 0)
g = bench.gen_rows(0, a.rows, a.dim, bench.SEED, dev, torch.bfloat16)
eng = B200RetrievalEngine.from_arrays(g, dtype="bfloat16", device=0, borrow=True, keep_host=False)
gbytes = a.rows * a.dim * 2 + 4 * a.rows
res = []
for k in a.k:
    for b in a.batches:
        q = torch.randn((b, a.dim), device=dev).to(torch.bfloat16).float()
        row = {"k": k, "b": b}
        for algo in ("scan", "gemm"):
            if algo == "scan" and b > 64:
                continue
            for _ in range(2):
                eng.search(q, k, algo=algo)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 5
            e0.record()
            for _ in range(n):
                eng.search(q, k, algo=algo)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            row[algo + "_ms"] = round(ms, 3)
            row[algo + "_gbs_single_pass"] = round(gbytes / (ms / 1e3) / 1e9, 1)
        res.append(row)
        print(json.dumps(row), flush=True)
