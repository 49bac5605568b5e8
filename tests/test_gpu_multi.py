"""Multi-GPU check of the sharded path (world 2 / 4 / 8, whatever is visible): per-rank shard search, the
NVLink peer-memory exchange with the fused merge + rerank kernel, the NCCL all-gather transport, the serving
loop -- all equal to the single-GPU result, bit for bit; a missing peer is reported after the timeout."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from multi_modal_retrieval_predict_project_b200 import B200RetrievalEngine, Reranker, synth
    from multi_modal_retrieval_predict_project_b200.sharded import ShardedSearcher, weighted_shard_bounds
    from oracle import search as osr
    n, d, b, k = 40001, 256, 70, 50
    g = osr.to_bf16_round(synth.make_embeddings(n, d, seed=5))
    q = osr.to_bf16_round(synth.make_embeddings(b, d, seed=6))
    rng = np.random.default_rng(7)
    masks = rng.integers(0, 2 ** 43, size=n + b, dtype=np.uint64) & rng.integers(0, 2 ** 43, size=n + b, dtype=np.uint64)
    kg = rng.standard_normal((n + b, 64)).astype(np.float32)
    kg /= np.linalg.norm(kg, axis=1, keepdims=True) + 1e-12
    # unequal shards (sized by a per-GPU rate, as bench.py does at N > 1): global rows = local + offset either way
    lo, hi = weighted_shard_bounds(n, [1.0 + 0.15 * ((r * 5) % 3) for r in range(world)], rank)
    eng = B200RetrievalEngine.from_arrays(g[lo:hi], dtype="bfloat16", device=rank, row_offset=lo)
    rer = Reranker.from_tables(masks, kg, device=rank)
    s = ShardedSearcher(eng)
    qd = torch.from_numpy(q).cuda()
    rows, scores = s.search(qd, k)
    q_rec = torch.arange(n, n + b, device="cuda")
    order, sc = s.rerank(rer, qd, rows, q_rec, rows.clone(), topk=20)                 # recomputed cosines, all-reduced
    rows2, scores2, order2, sc2 = s.search_rerank(rer, qd, k, q_rec, topk=20)       # one-collective path (search scores)
    assert torch.equal(rows2, rows) and torch.equal(scores2, scores)
    want_ids, want_fin = torch.gather(rows2, 1, order2.long()), sc2[:, :, 0].contiguous()
    # query-split post-processing over NVLink peer memory (csrc/exchange.cu): several rounds so that
    # both buffer parities are used, then other batch sizes / K (ragged last slice, re-created region)
    for _ in range(3):
        ids3, fin3 = s.retrieve_reranked(rer, qd, k, q_rec, topk=20)
        torch.cuda.synchronize()
        assert getattr(s, "_px", None) is not None, "peer exchange was not used"
        assert torch.equal(ids3, want_ids) and torch.equal(fin3, want_fin)
    ids3, fin3 = ids3.clone(), fin3.clone()      # views of the exchange region: it is re-created below (larger K)
    s._px.check()
    # soak: 400 back-to-back steps with no host synchronisation in between (the double-buffered regions, the flag
    # protocol and the per-parity counters are reused 200 times each); every 50th result is kept and compared
    kept = []
    for it in range(400):
        a_ids, a_fin = s.retrieve_reranked(rer, qd, k, q_rec, topk=20)
        if it % 50 == 49:
            kept.append((a_ids.clone(), a_fin.clone()))
    torch.cuda.synchronize()
    s._px.check()
    assert all(torch.equal(i_, want_ids) and torch.equal(f_, want_fin) for i_, f_ in kept)
    s_nccl = ShardedSearcher(eng, use_peer=False)                                     # same step through NCCL
    ids4, fin4 = s_nccl.retrieve_reranked(rer, qd, k, q_rec, topk=20)
    torch.cuda.synchronize()
    assert torch.equal(ids4, want_ids) and torch.equal(fin4, want_fin)
    # serving loop (pipelined copies) == direct calls, on every rank
    batches = [torch.from_numpy(osr.to_bf16_round(synth.make_embeddings(b, d, seed=60 + i))).pin_memory() for i in range(4)]
    got = [(i_.clone(), f_.clone()) for i_, f_ in s.serve(rer, iter(batches), k, q_rec, topk=20)]
    assert len(got) == len(batches)
    for hb, (g_ids, g_fin) in zip(batches, got):
        w_ids, w_fin = s_nccl.retrieve_reranked(rer, hb.cuda(), k, q_rec, topk=20)
        torch.cuda.synchronize()
        assert torch.equal(g_ids, w_ids.cpu()) and torch.equal(g_fin, w_fin.cpu())
    for bb, kk in ((33, 50), (1, 10), (70, 64), (5, 128)):
        a_ids, a_fin = s.retrieve_reranked(rer, qd[:bb].contiguous(), kk, q_rec[:bb].contiguous(), topk=0)
        b_ids, b_fin = s_nccl.retrieve_reranked(rer, qd[:bb].contiguous(), kk, q_rec[:bb].contiguous(), topk=0)
        torch.cuda.synchronize()
        assert torch.equal(a_ids, b_ids) and torch.equal(a_fin, b_fin), (bb, kk)
    # K beyond the fused kernels: the searcher falls back to the all-gather transport by itself
    f_ids, f_fin = s.retrieve_reranked(rer, qd[:9].contiguous(), 200, q_rec[:9].contiguous(), topk=0)
    g_ids, g_fin = s_nccl.retrieve_reranked(rer, qd[:9].contiguous(), 200, q_rec[:9].contiguous(), topk=0)
    torch.cuda.synchronize()
    assert torch.equal(f_ids, g_ids) and torch.equal(f_fin, g_fin)
    # all rows on rank 0, EMPTY shards elsewhere (more GPUs than data): both transports, same answer
    lo_e, hi_e = weighted_shard_bounds(n, [1.0] + [0.0] * (world - 1), rank)
    assert (hi_e - lo_e) == (n if rank == 0 else 0)
    eng_e = B200RetrievalEngine.from_arrays(g[lo_e:hi_e], dtype="bfloat16", device=rank, row_offset=lo_e)
    s_e = ShardedSearcher(eng_e)
    e_ids, e_fin = s_e.retrieve_reranked(rer, qd, k, q_rec, topk=20)
    torch.cuda.synchronize()
    assert getattr(s_e, "_px", None) is not None
    s_e._px.check()
    assert torch.equal(e_ids, want_ids) and torch.equal(e_fin, want_fin)
    n_ids, n_fin = ShardedSearcher(eng_e, use_peer=False).retrieve_reranked(rer, qd, k, q_rec, topk=20)
    torch.cuda.synchronize()
    assert torch.equal(n_ids, want_ids) and torch.equal(n_fin, want_fin)
    s_e.close()
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), rows=rows.cpu().numpy(), scores=scores.cpu().numpy(),
             order=order.cpu().numpy(), sc=sc.cpu().numpy(), ids=ids3.cpu().numpy(), fin=fin3.cpu().numpy())
    if rank == 0:  # single-shard reference on the same device
        full = B200RetrievalEngine.from_arrays(g, dtype="bfloat16", device=0)
        r1, s1 = full.search(qd, k)
        o1, c1 = rer.rerank_device(full, qd, r1, q_rec, r1.clone(), topk=20)
        i1, f1 = rer.rerank_scored_device(r1, s1, q_rec, topk=20)
        np.savez(os.path.join(out_dir, "single.npz"), rows=r1.cpu().numpy(), scores=s1.cpu().numpy(),
                 order=o1.cpu().numpy(), sc=c1.cpu().numpy(), ids=i1.cpu().numpy(), fin=f1.cpu().numpy())
    dist.barrier()
    # a rank that never shows up: the others give up after the timeout and report it (no hung GPU)
    s2 = ShardedSearcher(eng)
    s2.retrieve_reranked(rer, qd, k, q_rec, topk=20)
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        from multi_modal_retrieval_predict_project_b200 import _lib
        _lib.check(_lib.load().mmr_exchange_set_timeout(s2._px._h, 300))
        s2.retrieve_reranked(rer, qd, k, q_rec, topk=20)            # rank 0 alone runs a step
        torch.cuda.synchronize()
        try:
            s2._px.check()
            raise AssertionError("the missing peer was not reported")
        except _lib.MMRError as e:
            assert "timed out" in str(e)
    dist.barrier()
    s2.close(); s.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_sharded_equals_single(tmp_path, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0)); port = sk.getsockname()[1]
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    single = np.load(tmp_path / "single.npz")
    for rank in range(world):
        z = np.load(tmp_path / f"r{rank}.npz")
        assert np.array_equal(z["rows"], single["rows"]) and np.array_equal(z["scores"], single["scores"])
        assert np.array_equal(z["order"], single["order"]) and np.allclose(z["sc"], single["sc"], rtol=0, atol=1e-12)
        # the sharded step (NVLink exchange, fused merge + rerank) == the single-GPU fused tail, bit for bit
        assert np.array_equal(z["ids"], single["ids"]) and np.array_equal(z["fin"], single["fin"])


def test_serve_loop_single_gpu():
    """ShardedSearcher.serve on one GPU (no process group): pinned host batches in, host results out,
    equal to the direct device call for every batch."""
    from multi_modal_retrieval_predict_project_b200 import B200RetrievalEngine, Reranker, synth
    from multi_modal_retrieval_predict_project_b200.sharded import ShardedSearcher
    from oracle import search as osr
    n, d, b, k = 20000, 128, 300, 40
    g = osr.to_bf16_round(synth.make_embeddings(n, d, seed=15))
    rng = np.random.default_rng(17)
    masks = rng.integers(0, 2 ** 43, size=n + b, dtype=np.uint64) & rng.integers(0, 2 ** 43, size=n + b, dtype=np.uint64)
    kg = rng.standard_normal((n + b, 32)).astype(np.float32)
    eng = B200RetrievalEngine.from_arrays(g, dtype="bfloat16", device=0)
    rer = Reranker.from_tables(masks, kg, device=0)
    s = ShardedSearcher(eng)
    q_rec = torch.arange(n, n + b, device="cuda")
    batches = [torch.from_numpy(osr.to_bf16_round(synth.make_embeddings(b, d, seed=30 + i))).pin_memory() for i in range(5)]
    got = [(i_.clone(), f_.clone()) for i_, f_ in s.serve(rer, iter(batches), k, q_rec, topk=10)]
    assert len(got) == 5
    for hb, (g_ids, g_fin) in zip(batches, got):
        w_ids, w_fin = s.retrieve_reranked(rer, hb.cuda(), k, q_rec, topk=10)
        torch.cuda.synchronize()
        assert g_ids.shape == (b, 10) and torch.equal(g_ids, w_ids.cpu()) and torch.equal(g_fin, w_fin.cpu())
