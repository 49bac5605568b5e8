"""Diagnostic (1 GPU): CUDA-event timing of the pieces of one search+rerank step after the GEMM:
ingest of the queries, select, rerank features, rerank combine -- and the fused tail kernel that replaces the
last two on the batched path.  ROWS/BATCH from the environment."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from multi_modal_retrieval_predict_project_b200 import B200RetrievalEngine, Reranker

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
rows, dim, k = int(os.environ.get("ROWS", 2_000_000)), 512, 100
b = int(os.environ.get("BATCH", 4096))
g = bench.gen_rows(0, rows, dim, bench.SEED, dev, torch.bfloat16)
eng = B200RetrievalEngine.from_arrays(g, dtype="bfloat16", device=0, borrow=True, keep_host=False)
q = torch.randn((b, dim), device=dev).to(torch.bfloat16).float()
masks = bench.gen_masks(0, rows + b, dev)
kg = bench.gen_rows(0, rows + b, 300, bench.SEED + 700000, dev, torch.float32, normalize=True)
rer = Reranker.from_tables(masks, kg, device=0)
del masks, kg
q_rec = torch.arange(rows, rows + b, device=dev)


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out


eng.profile(True)
t_search, (r, s) = timed(lambda: eng.search(q, k))
gemm_ms, gemm_n = eng.profile(False)
t_rerank, _ = timed(lambda: rer.rerank_device(eng, q, r, q_rec, r, k))
t_cos, cos = timed(lambda: rer.candidate_cosine_device(eng, q, r))
t_rr_cos, _ = timed(lambda: rer.rerank_with_cos_device(cos, q_rec, r, k))
t_fused, _ = timed(lambda: rer.rerank_scored_device(r, s, q_rec, k))          # the fused tail (mmr_rerank_scored)
print({"rows": rows, "batch": b, "search_ms": round(t_search, 3), "gemm_ms": round(gemm_ms / gemm_n, 3),
       "ingest+select_ms": round(t_search - gemm_ms / gemm_n, 3), "rerank(features+combine)_ms": round(t_rerank, 3),
       "candidate_cosine_ms": round(t_cos, 3), "rerank_with_cos_ms": round(t_rr_cos, 3),
       "fused_tail_ms": round(t_fused, 3)})
