"""Diagnostic (1 GPU): GEMM kernel time vs gallery rows at a fixed batch (CUDA events through
mmr_index_profile).  With a -DMMR_DIAG build (MMR_B200_LIB=.../libmmr_b200_diag.so) MMR_B200_GEMM_DEBUG=1
disables the epilogue's score processing to expose the TMA + MMA pipeline alone."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from multi_modal_retrieval_predict_project_b200 import B200RetrievalEngine

dev = torch.device("cuda", 0)
b, k, dim = int(os.environ.get("BATCH", 4096)), 100, 512
q = torch.randn((b, dim), device=dev).to(torch.bfloat16).float()
for rows in [int(x) for x in os.environ.get("ROWS", "312500,625000,1250000,2500000,5000000").split(",")]:
    g = bench.gen_rows(0, rows, dim, bench.SEED, dev, torch.bfloat16)
    eng = B200RetrievalEngine.from_arrays(g, dtype="bfloat16", device=0, borrow=True, keep_host=False)
    for _ in range(3):
        eng.search(q, k, algo="gemm")
    torch.cuda.synchronize()
    eng.profile(True)
    for _ in range(10):
        eng.search(q, k, algo="gemm")
    torch.cuda.synchronize()
    ms, n = eng.profile(False)
    ms /= n
    print({"rows": rows, "batch": b, "gemm_ms": round(ms, 3), "tflops": round(2.0 * b * rows * dim / ms / 1e9, 1),
           "us_per_256row_tile_per_cta": round(ms * 1e3 / (rows / 256 / 9 * 2), 3)})
    del eng, g
