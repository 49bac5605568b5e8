"""BASELINE cfg5: on-device evaluation -- Q queries x N gallery rows, top-100, then P@K / Recall@K /
AP / RR / nDCG against synthetic relevance sets, everything resident on the GPU.

    python scripts/eval_cfg5.py [--queries 10000] [--rows 1000000] [--k 100] [--check 64]

Prints one JSON line with the search and metric kernel times; ``--check n`` re-computes the first n
queries with the oracle (numpy) and asserts identical metric values."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from multi_modal_retrieval_predict_project_b200 import B200RetrievalEngine
from multi_modal_retrieval_predict_project_b200.Helpers import metrics_from_rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--queries", type=int, default=10000)
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=512)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--max-rel", type=int, default=200)
    ap.add_argument("--check", type=int, default=64)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    g = bench.gen_rows(0, a.rows, a.dim, bench.SEED, dev, torch.bfloat16)
    eng = B200RetrievalEngine.from_arrays(g, dtype="bfloat16", device=0, borrow=True, keep_host=False)
    gq = torch.Generator(device=dev); gq.manual_seed(bench.SEED + 900_000)
    q = torch.randn((a.queries, a.dim), generator=gq, device=dev).to(torch.bfloat16).float()
    # synthetic relevance: per query a sorted unique set of U[1, max_rel] gallery rows (seed 2709 + 1)
    rng = np.random.default_rng(bench.SEED + 1)
    sizes = rng.integers(1, a.max_rel + 1, size=a.queries)
    indptr = np.zeros(a.queries + 1, dtype=np.int64); indptr[1:] = np.cumsum(sizes)
    rel = np.concatenate([np.sort(rng.choice(a.rows, size=s, replace=False)) for s in sizes]).astype(np.int64)
    d_indptr, d_rel = torch.from_numpy(indptr).to(dev), torch.from_numpy(rel).to(dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    rows, _ = eng.search(q, a.k); metrics_from_rows(rows, d_indptr, d_rel, a.k)   # warm-up
    torch.cuda.synchronize()
    ev[0].record(); rows, scores = eng.search(q, a.k)
    ev[1].record(); tbl = metrics_from_rows(rows, d_indptr, d_rel, a.k)
    ev[2].record(); torch.cuda.synchronize()
    t = tbl.cpu().numpy()
    out = {"workload": f"cfg5: {a.queries} queries x {a.rows}x{a.dim} bf16 gallery, top-{a.k} + metrics on device",
           "search_ms": ev[0].elapsed_time(ev[1]), "metrics_ms": ev[1].elapsed_time(ev[2]),
           "tflops": 2.0 * a.queries * a.rows * a.dim / (ev[0].elapsed_time(ev[1]) / 1e3) / 1e12,
           "P@k": float(np.mean(t[:, 0])), "R@k": float(np.mean(t[:, 1])), "mAP": float(np.mean(t[:, 2])),
           "MRR": float(np.mean(t[:, 3])), "nDCG": float(np.mean(t[:, 4]))}
    if a.check > 0:
        from oracle import metrics as om
        r = rows[: a.check].cpu().numpy()
        rets = [[int(x) for x in r[i]] for i in range(a.check)]
        rels = [rel[indptr[i]:indptr[i + 1]].tolist() for i in range(a.check)]
        assert np.array_equal(om.per_query_table(rets, rels, a.k), t[: a.check]), "metrics differ from the oracle"
        out["checked_vs_oracle"] = a.check
    print(json.dumps(out))


if __name__ == "__main__":
    main()
