"""Diagnostic (1 GPU, -DMMR_DIAG -DMMR_GEMM_TRACE build loaded through MMR_B200_LIB): where the time of ONE GEMM
launch goes.  Prints the CUDA-event kernel time (10 launches) and, from the per-CTA trace of one more launch,
per wave: CTA entry, first accumulator tile ready, steady-state us / tile, last tile drained, final pass done --
min / median / max over CTAs, us since the first CTA entered.
  ROWS=1250000 BATCH=4096 MMR_B200_GEMM_DEBUG=<bits> python scripts/diag_gemm.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from multi_modal_retrieval_predict_project_b200 import B200RetrievalEngine
dev = torch.device("cuda", 0)
rows, b, k, dim = int(os.environ.get("ROWS", 1250000)), int(os.environ.get("BATCH", 4096)), 100, 512
variant = os.environ.get("VARIANT", "auto")
g = bench.gen_rows(0, rows, dim, bench.SEED, dev, torch.bfloat16)
eng = B200RetrievalEngine.from_arrays(g, dtype="bfloat16", device=0, borrow=True, keep_host=False)
eng.tune(variant=variant, parts=int(os.environ.get("PARTS", 0)))
q = bench.gen_queries(b, dim, dev)
path = os.environ.pop("MMR_B200_GEMM_TRACE", None)       # timing runs without the trace (it synchronises)
for _ in range(3):
    eng.search(q, k, algo="gemm")
torch.cuda.synchronize()
eng.profile(True)
for _ in range(10):
    eng.search(q, k, algo="gemm")
torch.cuda.synchronize()
ms, n = eng.profile(False)
plan = eng.last_plan()
tiles = plan["tiles_per_part"]
print({"rows": rows, "batch": b, "debug": os.environ.get("MMR_B200_GEMM_DEBUG", "0"), "plan": plan, "gemm_ms": round(ms / n, 3),
       "tflops": round(2.0 * b * rows * dim / (ms / n) / 1e9, 1)})
if path is None or os.environ.get("NO_TRACE"):
    sys.exit(0)
# one traced launch in a child process (the library reads the variable once)
import subprocess
code = f"""
import os, sys
sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})
import torch, bench
from multi_modal_retrieval_predict_project_b200 import B200RetrievalEngine
dev = torch.device('cuda', 0)
g = bench.gen_rows(0, {rows}, {dim}, bench.SEED, dev, torch.bfloat16)
eng = B200RetrievalEngine.from_arrays(g, dtype='bfloat16', device=0, borrow=True, keep_host=False)
eng.tune(variant={variant!r}, parts={int(os.environ.get('PARTS', 0))})
q = bench.gen_queries({b}, {dim}, dev)
for _ in range(4):
    eng.search(q, {k}, algo='gemm')
torch.cuda.synchronize()
"""
env = dict(os.environ, MMR_B200_GEMM_TRACE=path)
subprocess.run([sys.executable, "-c", code], env=env, check=True)
t = np.fromfile(path, dtype=np.uint64).reshape(-1, 64)
t = t[t[:, 1] > 0]
entry = (t[:, 0] & np.uint64(0xFFFFFFFFFFFF)).astype(np.int64)
smid = (t[:, 0] >> np.uint64(48)).astype(np.int64)
t0 = entry.min()
mask48 = (1 << 48) - 1
def rel(col):
    return ((t[:, col].astype(np.int64) & mask48) - (t0 & mask48)) / 1e3
e = (entry - t0) / 1e3
wave2 = e > 200
def stat(x):
    return f"min {x.min():8.1f}  med {np.median(x):8.1f}  max {x.max():8.1f}"
for name, sel in (("wave 1", ~wave2), ("wave 2", wave2)):
    if sel.sum() == 0:
        continue
    print(f"{name}: {int(sel.sum())} CTAs on {len(set(smid[sel].tolist()))} SMs")
    print("   entry             ", stat(e[sel]))
    print("   tile 0 ready      ", stat(rel(1)[sel]), "  (after entry:", stat(rel(1)[sel] - e[sel]), ")")
    print("   tile 8 ready      ", stat(rel(9)[sel]))
    last32 = 40 + ((tiles - 1) >> 5)
    if tiles > 64:
        per = (rel(last32)[sel] - rel(41)[sel]) / (32.0 * ((tiles - 1) >> 5) - 32.0)
        print("   steady us / tile  ", stat(per))
    print("   last tile drained ", stat(rel(63)[sel]), "  (since entry:", stat(rel(63)[sel] - e[sel]), ")")
    print("   final pass done   ", stat(rel(62)[sel]), "  (final pass:", stat(rel(62)[sel] - rel(63)[sel]), ")")
print("kernel span from the trace (us):", round(float(rel(62).max()), 1))
ch, sl, gr, ap = (t[:, c].astype(np.float64) for c in (58, 59, 60, 61))
print("append path (warp 0 of warpgroup 0 of every CTA): chunks %.0f, took the append path %.1f %%, hit groups per such chunk %.2f, "
      "entries appended by lane 0: %.0f" % (ch.mean(), 100.0 * sl.sum() / max(ch.sum(), 1), gr.sum() / max(sl.sum(), 1), ap.mean()))
