"""Diagnostic: CUDA-event timing of the phases of the sharded step over the NCCL transport (torchrun --nproc-per-node N):
search | local cosine | blob all-gather | merge+payload of this rank's query slice | rerank |
apply order | result all-gathers, next to the whole retrieve_reranked() call and the replicated
search_rerank() it replaced."""
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
s = ShardedSearcher(eng, use_peer=False)   # phase split of the NCCL transport
sp = ShardedSearcher(eng)                    # NVLink peer-memory transport (whole call only)
q = torch.randn((b, dim), device=dev).to(torch.bfloat16).float()
masks = bench.gen_masks(0, rows + b, dev); kg = bench.gen_rows(0, rows + b, 300, bench.SEED + 700000, dev, torch.float32, normalize=True)
rer = Reranker.from_tables(masks, kg, device=lr); del masks, kg
q_rec = torch.arange(rows, rows + b, device=dev)
for _ in range(3):
    s.retrieve_reranked(rer, q, k, q_rec, topk=k); sp.retrieve_reranked(rer, q, k, q_rec, topk=k); s.search_rerank(rer, q, k, q_rec, topk=k)
torch.cuda.synchronize()
lib = _lib.load()
names = ["search", "cosine", "allgather_blob", "merge_slice", "rerank_slice", "apply_order", "allgather_result"]
tot = {n: 0.0 for n in names}
whole = {"retrieve_reranked(nccl)": 0.0, "retrieve_reranked(peer)": 0.0, "search_rerank(replicated)": 0.0}
per = (b + world - 1) // world
q_lo, q_hi = min(b, rank * per), min(b, (rank + 1) * per)
iters = 5
for it in range(iters):
    dist.barrier(); torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    bk = b * k
    f = s._blob[: bk * 8].view(torch.float32)
    scores_v, cos_v = f[:bk].view(b, k), f[bk:].view(b, k)
    rows_v = s._blob[bk * 8:].view(torch.int64).view(b, k)
    ev[0].record()
    eng.search(q, k, out_rows=rows_v, out_scores=scores_v); ev[1].record()
    rer.candidate_cosine_device(eng, q, rows_v, out=cos_v); ev[2].record()
    dist.all_gather_into_tensor(s._gblob, s._blob); ev[3].record()
    out_r, out_s, cos = s._merge_slice(s._gblob, b, k, q_lo, q_hi); ev[4].record()
    order, sc = rer.rerank_with_cos_device(cos, q_rec[q_lo:q_hi], out_r, k); ev[5].record()
    all_ids, all_fin, my_ids, my_fin = s._fin
    _lib.check(lib.mmr_apply_order(_lib.ptr(out_r), _lib.ptr(order), _lib.ptr(sc), q_hi - q_lo, k, k, _lib.ptr(my_ids),
                                   _lib.ptr(my_fin), lr, _lib.current_stream(lr))); ev[6].record()
    dist.all_gather_into_tensor(all_ids, my_ids); dist.all_gather_into_tensor(all_fin, my_fin); ev[7].record()
    torch.cuda.synchronize()
    for i, n in enumerate(names): tot[n] += ev[i].elapsed_time(ev[i + 1])
    for name, fn in (("retrieve_reranked(nccl)", lambda: s.retrieve_reranked(rer, q, k, q_rec, topk=k)),
                     ("retrieve_reranked(peer)", lambda: sp.retrieve_reranked(rer, q, k, q_rec, topk=k)),
                     ("search_rerank(replicated)", lambda: s.search_rerank(rer, q, k, q_rec, topk=k))):
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        whole[name] += e0.elapsed_time(e1)
if rank == 0:
    print({"world": world, "rows": rows, "batch": b})
    print({n: round(v / iters, 3) for n, v in tot.items()})
    print({n: round(v / iters, 3) for n, v in whole.items()})
dist.destroy_process_group()
