"""Diagnostic: CUDA-event timing of the phases of the sharded step (torchrun --nproc-per-node N)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
from multi_modal_retrieval_predict_project_b200 import B200RetrievalEngine, Reranker, _lib
from multi_modal_retrieval_predict_project_b200.sharded import ShardedSearcher, shard_bounds

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
rows, dim, b, k = int(os.environ.get("ROWS", 10_000_000)), 512, 4096, 100
lo, hi = shard_bounds(rows, world, rank)
g = bench.gen_rows(lo, hi, dim, bench.SEED, dev, torch.bfloat16)
eng = B200RetrievalEngine.from_arrays(g, dtype="bfloat16", device=lr, row_offset=lo, borrow=True, keep_host=False)
s = ShardedSearcher(eng)
q = torch.randn((b, dim), device=dev).to(torch.bfloat16).float()
masks = bench.gen_masks(0, rows + b, dev); kg = bench.gen_rows(0, rows + b, 300, bench.SEED + 700000, dev, torch.float32, normalize=True)
rer = Reranker.from_tables(masks, kg, device=lr); del masks, kg
q_rec = torch.arange(rows, rows + b, device=dev)
for _ in range(3): s.search_rerank(rer, q, k, q_rec, topk=k)
torch.cuda.synchronize()
names = ["search", "cosine", "allgather", "merge", "glue", "rerank"]
tot = {n: 0.0 for n in names}
lib = _lib.load()
for it in range(5):
    dist.barrier(); torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(7)]
    bk = b * k
    f = s._blob[: bk * 8].view(torch.float32)
    scores_v, cos_v = f[:bk].view(b, k), f[bk:].view(b, k)
    rows_v = s._blob[bk * 8:].view(torch.int64).view(b, k)
    ev[0].record()
    eng.search(q, k, out_rows=rows_v, out_scores=scores_v); ev[1].record()
    rer.candidate_cosine_device(eng, q, rows_v, out=cos_v); ev[2].record()
    dist.all_gather_into_tensor(s._gblob, s._blob); ev[3].record()
    gf = s._gblob.view(torch.float32); gi = s._gblob.view(torch.int64)
    out_s = torch.empty((b, k), dtype=torch.float32, device=dev); out_r = torch.empty((b, k), dtype=torch.int64, device=dev)
    src = torch.empty((b, k), dtype=torch.int32, device=dev)
    _lib.check(lib.mmr_merge_topk_strided(gf.data_ptr(), gi.data_ptr() + bk * 8, world, b, k, 4 * bk, 2 * bk, k, _lib.ptr(out_s), _lib.ptr(out_r), _lib.ptr(src), lr, _lib.current_stream(lr))); ev[4].record()
    srcl = src.long().clamp_(min=0)
    qoff = torch.arange(b, device=dev, dtype=torch.int64).unsqueeze(1) * k
    cos = gf[((srcl // k) * (4 * bk) + bk + qoff + (srcl % k)).reshape(-1)].view(b, k).contiguous(); ev[5].record()
    order, sc = rer.rerank_with_cos_device(cos, q_rec, out_r, k); ev[6].record()
    torch.cuda.synchronize()
    for i, n in enumerate(names): tot[n] += ev[i].elapsed_time(ev[i + 1])
if rank == 0: print({n: round(v / 5, 3) for n, v in tot.items()})
dist.destroy_process_group()
