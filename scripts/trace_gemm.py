"""Diagnostic: per-CTA progress curves of one GEMM launch.  Needs a library built with
-DMMR_GEMM_TRACE (NVCC_EXTRA=-DMMR_GEMM_TRACE python -m multi_modal_retrieval_predict_project_b200.build --force).  Prints, for the
first and second wave, when tiles 0..39 and then every 32nd tile became ready (us since the first timestamp): min / median
/ max over CTAs."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
path = os.environ.setdefault("MMR_B200_GEMM_TRACE", "/tmp/gemm_trace.bin")
import torch
import bench
from multi_modal_retrieval_predict_project_b200 import B200RetrievalEngine
dev = torch.device("cuda", 0)
rows, b, k, dim = int(os.environ.get("ROWS", 1250000)), int(os.environ.get("BATCH", 4096)), 100, 512
g = bench.gen_rows(0, rows, dim, bench.SEED, dev, torch.bfloat16)
eng = B200RetrievalEngine.from_arrays(g, dtype="bfloat16", device=0, borrow=True, keep_host=False)
q = torch.randn((b, dim), device=dev).to(torch.bfloat16).float()
for _ in range(4):
    eng.search(q, k, algo="gemm")
torch.cuda.synchronize()
t = np.fromfile(path, dtype=np.uint64).reshape(-1, 64)
t = t[t[:, 1] > 0]                       # CTAs that recorded (leaders and peers both have group 0)
t0 = t[:, 1:][t[:, 1:] > 0].min()
first = (t[:, 1].astype(np.int64) - int(t0)) / 1e3
wave2 = first > np.median(first) + 200   # CTAs of the second wave start much later
for name, sel in (("wave 1", ~wave2), ("wave 2", wave2)):
    tt = t[sel]
    if len(tt) == 0:
        continue
    print(name, "CTAs:", len(tt))
    cols = [c for c in range(1, 64) if (tt[:, c] > 0).all() and (c <= 12 or c > 40)]
    for c in cols:
        us = (tt[:, c].astype(np.int64) - int(t0)) / 1e3
        print(f"  tile {(c - 1) if c <= 40 else 32 * (c - 40):5d}: min {us.min():9.1f}  med {np.median(us):9.1f}  max {us.max():9.1f}  spread {us.max() - us.min():7.1f} us")
