"""CPU, world_size 2, gloo: the host-side plumbing of the row-sharded search (shard bounds, global
row offsets, all-gather layout, merge inputs).  The per-shard search and the merge are replaced by
oracle stand-ins here (the CUDA kernels need a GPU; they are covered by the -m gpu tests, and
tests/test_gpu_search.py::test_merge_topk_matches_single_shard emulates the ranks on one device)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import search as osr


class _OracleShardEngine:
    """Stand-in with the engine's search() contract over one row shard (CPU tensors)."""

    def __init__(self, g, row_offset):
        self.g, self.row_offset = g, row_offset

    def search(self, queries, K, algo=None):
        rows, scores = osr.exact_topk(queries.numpy(), self.g, K)
        kk = rows.shape[1]
        r = -np.ones((rows.shape[0], K), np.int64); s = np.full((rows.shape[0], K), -np.inf, np.float32)
        r[:, :kk] = rows + self.row_offset; s[:, :kk] = scores
        return torch.from_numpy(r), torch.from_numpy(s)


def _cpu_merge(scores, rows, k_out):
    """Same contract as sharded.merge_topk: (n_lists, B, K) -> (B, k_out), score desc then row asc."""
    n_lists, b, k_in = scores.shape
    s = scores.permute(1, 0, 2).reshape(b, -1).numpy(); r = rows.permute(1, 0, 2).reshape(b, -1).numpy()
    out_r = -np.ones((b, k_out), np.int64); out_s = np.full((b, k_out), -np.inf, np.float32)
    for i in range(b):
        valid = r[i] >= 0
        order = np.lexsort((r[i][valid], -s[i][valid].astype(np.float64)))[:k_out]
        out_r[i, :len(order)] = r[i][valid][order]; out_s[i, :len(order)] = s[i][valid][order]
    return torch.from_numpy(out_r), torch.from_numpy(out_s)


def _worker(rank, world, port, n, d, b, k, out_dir, weights=None):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from multi_modal_retrieval_predict_project_b200 import synth
    from multi_modal_retrieval_predict_project_b200.sharded import ShardedSearcher, shard_bounds, weighted_shard_bounds
    g = synth.make_embeddings(n, d, seed=3)
    q = synth.make_embeddings(b, d, seed=4)
    lo, hi = shard_bounds(n, world, rank) if weights is None else weighted_shard_bounds(n, weights, rank, align=8)
    s = ShardedSearcher(_OracleShardEngine(g[lo:hi], lo), merge=_cpu_merge)
    assert s.world == world
    rows, scores = s.search(torch.from_numpy(q), k)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), rows=rows.numpy(), scores=scores.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n,k,weights", [(1001, 10, None), (37, 25, None), (1001, 10, [1.0, 2.5]), (37, 25, [0.0, 1.0])])
def test_sharded_search_plumbing_world2(tmp_path, n, k, weights):
    """world-2 gloo run of the sharded search plumbing: equal shards, unequal (weighted) shards, an EMPTY shard."""
    world, d, b = 2, 32, 9
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0)); port = sk.getsockname()[1]
    mp.spawn(_worker, args=(world, port, n, d, b, k, str(tmp_path), weights), nprocs=world, join=True)
    from multi_modal_retrieval_predict_project_b200 import synth
    g = synth.make_embeddings(n, d, seed=3); q = synth.make_embeddings(b, d, seed=4)
    want_rows, want_scores = osr.exact_topk(q, g, k)
    for rank in range(world):
        z = np.load(tmp_path / f"r{rank}.npz")
        kk = want_rows.shape[1]
        assert np.array_equal(z["rows"][:, :kk], want_rows)            # every rank ends with the global top-K
        assert np.allclose(z["scores"][:, :kk], want_scores, rtol=1e-6)
        assert np.all(z["rows"][:, kk:] == -1)
