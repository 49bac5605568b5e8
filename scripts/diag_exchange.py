"""Diagnostic (torchrun --nproc-per-node N): CUDA-event timing of the two calls of the sharded step over the NVLink
peer-memory exchange, per rank -- mmr_search_scatter (query prep + GEMM + selection kernel storing into the owners'
regions + signal) with the GEMM's own time inside it, and mmr_exchange_rerank (wait for the lists + merge + rerank +
publish, then the wait for every rank's results) -- next to the whole retrieve_reranked() call.  Prints min / max
over ranks and the bytes each rank moves over NVLink per step (counted from the layout, not from hardware counters:
ncu cannot wrap a multi-rank job)."""
import ctypes as C
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
from multi_modal_retrieval_predict_project_b200 import B200RetrievalEngine, Reranker, _lib
from multi_modal_retrieval_predict_project_b200.sharded import ShardedSearcher, shard_bounds

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
rows, dim, b, k = int(os.environ.get("ROWS", 10_000_000)), 512, int(os.environ.get("BATCH", 4096)), 100
lo, hi = shard_bounds(rows, world, rank)
g = bench.gen_rows(lo, hi, dim, bench.SEED, dev, torch.bfloat16)
eng = B200RetrievalEngine.from_arrays(g, dtype="bfloat16", device=lr, row_offset=lo, borrow=True, keep_host=False)
s = ShardedSearcher(eng)
q = bench.gen_queries(b, dim, dev)
masks = bench.gen_masks(0, rows + b, dev); kg = bench.gen_rows(0, rows + b, 300, bench.SEED + 700000, dev, torch.float32, normalize=True)
rer = Reranker.from_tables(masks, kg, device=lr); del masks, kg
q_rec = torch.arange(rows, rows + b, device=dev)
for _ in range(5):
    s.retrieve_reranked(rer, q, k, q_rec, topk=k)
torch.cuda.synchronize(); dist.barrier()
lib, px = _lib.load(), s._px
iters = 20
t_scatter = t_rerank = t_whole = 0.0
eng.profile(True)
for it in range(iters):
    dist.barrier(); torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    px.step += 1
    st = _lib.current_stream(lr)
    e[0].record()
    _lib.check(lib.mmr_search_scatter(eng._handle, px._h, _lib.ptr(q), b, _lib.MMR_F32, k, 0, px.step, st))
    e[1].record()
    p_ids, p_fin = C.c_void_p(), C.c_void_p()
    _lib.check(lib.mmr_exchange_rerank(px._h, rer._tables, _lib.ptr(q_rec), b, k, rer.alpha, rer.beta, rer.gamma, k, px.step,
                                       C.byref(p_ids), C.byref(p_fin), st))
    e[2].record()
    torch.cuda.synchronize()
    t_scatter += e[0].elapsed_time(e[1]); t_rerank += e[1].elapsed_time(e[2]); t_whole += e[0].elapsed_time(e[2])
gemm_ms, gemm_n = eng.profile(False)
mine = torch.tensor([t_scatter / iters, gemm_ms / max(gemm_n, 1), t_rerank / iters, t_whole / iters], device=dev, dtype=torch.float64)
allv = [torch.zeros_like(mine) for _ in range(world)]
dist.all_gather(allv, mine)
if rank == 0:
    m = torch.stack(allv).cpu()
    names = ["search_scatter_ms (prep + GEMM + select->peer stores + signal)", "  of which gemm_topk_kernel_ms",
             "exchange_rerank_ms (wait lists + merge + rerank + publish + wait results)", "whole step_ms"]
    per = (b + world - 1) // world
    kp = (k + 3) // 4 * 4
    print({"world": world, "rows": rows, "batch": b, "k": k,
           "nvlink_bytes_out_per_rank_per_step": {"lists": (b - per) * kp * 12, "results": per * k * 16 * (world - 1)}})
    for i, n in enumerate(names):
        print(f"{n}: min {m[:, i].min():.3f}  max {m[:, i].max():.3f}  mean {m[:, i].mean():.3f}")
    print("outside the GEMM (mean over ranks of step - gemm): %.3f ms" % float((m[:, 3] - m[:, 1]).mean()))
s.close()
dist.destroy_process_group()
