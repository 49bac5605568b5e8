"""Diagnostic: per-phase wall time of the sharded end-to-end step (torchrun --nproc-per-node N)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
from multi_modal_retrieval_predict_project_b200 import B200RetrievalEngine, Reranker
from multi_modal_retrieval_predict_project_b200.sharded import ShardedSearcher, shard_bounds

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if world > 1: dist.init_process_group("nccl", device_id=dev)
rows, dim, b, k = 10_000_000, 512, 4096, 100
lo, hi = shard_bounds(rows, world, rank)
g = bench.gen_rows(lo, hi, dim, bench.SEED, dev, torch.bfloat16)
eng = B200RetrievalEngine.from_arrays(g, dtype="bfloat16", device=lr, row_offset=lo, borrow=True, keep_host=False)
s = ShardedSearcher(eng)
q = torch.randn((b, dim), device=dev).to(torch.bfloat16).float(); qh = q.cpu().pin_memory(); qd = torch.empty_like(q)
masks = bench.gen_masks(0, rows + b, dev); kg = bench.gen_rows(0, rows + b, 300, bench.SEED + 700000, dev, torch.float32, normalize=True)
rer = Reranker.from_tables(masks, kg, device=lr); del masks, kg
q_rec = torch.arange(rows, rows + b, device=dev)
def sync(): torch.cuda.synchronize()
for _ in range(3): s.search_rerank(rer, q, k, q_rec, topk=k)
sync()
T = {}
def tick(name, t0):
    sync(); T[name] = T.get(name, 0) + time.perf_counter() - t0
for it in range(5):
    if world > 1: dist.barrier()
    sync()
    t0 = time.perf_counter(); qd.copy_(qh, non_blocking=True); tick("h2d", t0)
    t0 = time.perf_counter(); r, sc = eng.search(qd, k); tick("search", t0)
    t0 = time.perf_counter(); out = s.search_rerank(rer, qd, k, q_rec, topk=k); tick("search_rerank(all)", t0)
    t0 = time.perf_counter(); o = torch.gather(out[0], 1, out[2].long()).cpu(); tick("gather+d2h", t0)
if rank == 0: print({k_: round(v / 5 * 1e3, 3) for k_, v in T.items()})
# e2e loop as in bench
outh = [torch.empty((b, k), dtype=torch.int64).pin_memory(), torch.empty((b, k), dtype=torch.float64).pin_memory()]
def e2e():
    qd.copy_(qh, non_blocking=True)
    rows_, _s, order, sc = s.search_rerank(rer, qd, k, q_rec, topk=k)
    res = (torch.gather(rows_, 1, order.long()), sc[:, :, 0].contiguous())
    if rank == 0:
        for h, r_ in zip(outh, res): h.copy_(r_, non_blocking=True)
    torch.cuda.current_stream().synchronize()
e2e(); sync()
t0 = time.perf_counter()
for _ in range(5): e2e()
sync()
if rank == 0: print("e2e ms/step (spin sync)", (time.perf_counter() - t0) / 5 * 1e3)
ev = torch.cuda.Event(blocking=True)
def e2e_b():
    qd.copy_(qh, non_blocking=True)
    rows_, _s, order, sc = s.search_rerank(rer, qd, k, q_rec, topk=k)
    res = (torch.gather(rows_, 1, order.long()), sc[:, :, 0].contiguous())
    if rank == 0:
        for h, r_ in zip(outh, res): h.copy_(r_, non_blocking=True)
    ev.record(); ev.synchronize()
e2e_b(); sync()
t0 = time.perf_counter()
for _ in range(5): e2e_b()
sync()
if rank == 0: print("e2e ms/step (blocking event)", (time.perf_counter() - t0) / 5 * 1e3, "cpus", os.cpu_count(), len(os.sched_getaffinity(0)))
# phase split inside search_rerank with blocking sync
def bsync():
    ev.record(); ev.synchronize()
for it in range(3):
    if world > 1: dist.barrier()
    bsync(); t0 = time.perf_counter(); out = s.search_rerank(rer, qd, k, q_rec, topk=k); bsync()
    if rank == 0: print("search_rerank blocking-sync ms", (time.perf_counter() - t0) * 1e3)
if world > 1: dist.destroy_process_group()
