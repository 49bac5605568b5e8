"""The fused exchange kernels on ONE GPU (world = 1, through the C ABI), at the N = 8 shard shape -- for ncu captures
of select_fast_kernel<..., RemoteSink> and exchange_rerank_kernel (ncu cannot wrap a multi-rank job; with world = 1
the peer table points at this GPU's own region, so the stores are local, the code path is the same)."""
import ctypes as C
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from multi_modal_retrieval_predict_project_b200 import B200RetrievalEngine, Reranker, _lib

dev = torch.device("cuda", 0)
rows, dim, b, k = int(os.environ.get("ROWS", 1_250_000)), 512, int(os.environ.get("BATCH", 4096)), 100
g = bench.gen_rows(0, rows, dim, bench.SEED, dev, torch.bfloat16)
eng = B200RetrievalEngine.from_arrays(g, dtype="bfloat16", device=0, borrow=True, keep_host=False)
q = bench.gen_queries(b, dim, dev)
masks = bench.gen_masks(0, rows + b, dev)
kg = bench.gen_rows(0, rows + b, 300, bench.SEED + 700000, dev, torch.float32, normalize=True)
rer = Reranker.from_tables(masks, kg, device=0)
q_rec = torch.arange(rows, rows + b, device=dev)
lib = _lib.load()
ex = C.c_void_p()
_lib.check(lib.mmr_exchange_create(C.byref(ex), 0, 1, b, k, 0))
st = _lib.current_stream(0)
for step in range(1, 7):
    _lib.check(lib.mmr_search_scatter(eng._handle, ex, _lib.ptr(q), b, _lib.MMR_F32, k, 0, step, st))
    p_ids, p_fin = C.c_void_p(), C.c_void_p()
    _lib.check(lib.mmr_exchange_rerank(ex, rer._tables, _lib.ptr(q_rec), b, k, rer.alpha, rer.beta, rer.gamma, k, step,
                                       C.byref(p_ids), C.byref(p_fin), st))
torch.cuda.synchronize()
ids = _lib.as_cuda_tensor(p_ids.value, (b, k), torch.int64, 0)
rows_, scores_ = eng.search(q, k)
want_ids, _ = rer.rerank_scored_device(rows_, scores_, q_rec, k)
assert torch.equal(ids, want_ids)
print("exchange path (world 1) == single-shard path for", b, "queries")
_lib.check(lib.mmr_exchange_destroy(ex))
