"""GPU, BASELINE.json full sizes (10M x 512 bf16): size-independent properties of the search path,
where the CPU oracle cannot finish in seconds.  Planted needles, sortedness, prefix consistency
(top-10 == head of top-100), agreement of the two independent kernels (HBM scan vs tcgen05 GEMM),
recomputation of the returned scores in fp32, shard + merge == single index (idempotence of the
merge), and the exclusion rule."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROWS, DIM = 10_000_000, 512


@pytest.fixture(scope="module")
def big():
    import bench
    from multi_modal_retrieval_predict_project_b200 import B200RetrievalEngine
    free, _ = torch.cuda.mem_get_info(0)
    if free < 40 << 30:
        pytest.skip("needs ~25 GB of free HBM")
    dev = torch.device("cuda", 0)
    g = bench.gen_rows(0, ROWS, DIM, bench.SEED, dev, torch.bfloat16)
    gq = torch.Generator(device=dev); gq.manual_seed(77)
    q = torch.randn((256, DIM), generator=gq, device=dev).to(torch.bfloat16)
    # planted needles: query i is (a positive multiple of) gallery row needle[i] for i < 32
    needles = torch.randint(0, ROWS, (32,), generator=gq, device=dev)
    q[:32] = (g[needles].float() * 3.0).to(torch.bfloat16)        # some positive multiple (rounded to bf16) ...
    g[needles] = q[:32]                                           # ... and the gallery row becomes exactly that vector
    eng = B200RetrievalEngine.from_arrays(g, dtype="bfloat16", device=0, borrow=True, keep_host=False)
    yield dict(g=g, q=q.float(), eng=eng, needles=needles)
    eng.close()


def _headline_queries(b=4096):
    import bench
    gq = torch.Generator(device="cuda"); gq.manual_seed(bench.SEED + 900_000)     # bench.py's query batch
    return torch.randn((b, DIM), generator=gq, device="cuda").to(torch.bfloat16).float()


@pytest.mark.parametrize("variant", ["auto", "long"])
def test_headline_shape_vs_bruteforce_oracle(big, variant):
    """The launch bench.py times at N = 1 -- 10M x 512 bf16, b = 4096, k = 100, AUTO dispatch = the CTA-pair
    instantiation gemm_topk_kernel<resident, pair, kProbe=1> on 9 parts x 4341 tiles -- compared with the
    oracle for 64 queries spread over all 32 query tiles: ids position by position outside near-ties and
    scores to 2e-5 against the fp32 brute force over the same bf16 values (oracle.search.topk_matches), exact
    recall@100 against the fp64 ranking with the epsilon rule.  The other instantiation (kProbe=0, round 1's
    choice for this launch) is forced on the same launch as well."""
    from oracle.bruteforce import check_topk
    eng, g = big["eng"], big["g"]
    q = _headline_queries()
    eng.tune(variant=variant)
    rows, scores = eng.search(q, 100)
    plan = eng.last_plan()
    eng.tune(variant="auto")
    assert plan["algo"] == "gemm" and plan["pair"] and plan["parts"] == 9 and plan["tiles_per_part"] == 4341, plan
    assert plan["variant"] == ("short" if variant == "auto" else "long"), plan
    sel = torch.arange(0, 4096, 64, device="cuda") + torch.arange(64, device="cuda") % 64   # one per 64, all lanes
    ok, detail = check_topk(rows[sel], scores[sel], g, q[sel], 100)
    assert ok, detail
    # every query: sorted, tie rule, in range, no duplicates inside a list
    assert bool((scores[:, :-1] >= scores[:, 1:]).all())
    tie = scores[:, :-1] == scores[:, 1:]
    assert bool((rows[:, :-1][tie] < rows[:, 1:][tie]).all())
    assert int(rows.min()) >= 0 and int(rows.max()) < ROWS
    assert bool((torch.sort(rows, dim=1).values.diff(dim=1) > 0).all())


def test_shard_shape_n8_vs_bruteforce_oracle(big):
    """The launches bench.py times at N = 8 (1.25M-row shard, b = 4096, k = 100: 64-tile pacing window) and
    at N = 2 (5M-row shard: 12-tile window), AUTO = <resident, pair, kProbe=1>.  64 sampled queries each
    against the brute-force oracle over the same rows."""
    from multi_modal_retrieval_predict_project_b200 import B200RetrievalEngine
    from oracle.bruteforce import check_topk
    g = big["g"]
    q = _headline_queries()
    sel = torch.arange(0, 4096, 64, device="cuda") + torch.arange(64, device="cuda") % 64
    for lo, hi, want in ((2_500_000, 3_750_000, "short"), (5_000_000, 10_000_000, "short")):
        shard = B200RetrievalEngine.from_arrays(g[lo:hi], dtype="bfloat16", device=0, borrow=True, keep_host=False,
                                                row_offset=lo)
        rows, scores = shard.search(q, 100)
        plan = shard.last_plan()
        assert plan["algo"] == "gemm" and plan["pair"] and plan["variant"] == want, plan
        ok, detail = check_topk(rows[sel], scores[sel], g[lo:hi], q[sel], 100, row_offset=lo)
        assert ok, (lo, hi, detail)
        shard.close()


def test_needles_sorted_and_prefix(big):
    eng, q = big["eng"], big["q"]
    r100, s100 = eng.search(q, 100)
    r10, s10 = eng.search(q, 10)
    assert torch.equal(r100[:32, 0], big["needles"]) and torch.allclose(s100[:32, 0], torch.ones(32, device="cuda"), atol=2e-6)
    assert bool((s100[:, :-1] >= s100[:, 1:]).all())                              # sorted descending
    tie = s100[:, :-1] == s100[:, 1:]
    assert bool((r100[:, :-1][tie] < r100[:, 1:][tie]).all())                     # ties: row ascending
    assert torch.equal(r10, r100[:, :10]) and torch.equal(s10, s100[:, :10])       # prefix consistency
    assert int(r100.min()) >= 0 and int(r100.max()) < ROWS
    assert all(len(set(row.tolist())) == 100 for row in r100[:16].cpu())           # no duplicates


def test_scan_and_gemm_kernels_agree_and_scores_recompute(big):
    eng, q, g = big["eng"], big["q"], big["g"]
    rg, sg = eng.search(q[:8], 100, algo="gemm")
    rs, ss = eng.search(q[:8], 100, algo="scan")
    # two independent kernels (different accumulation orders): same ids except near-ties, scores to 2e-5
    assert torch.allclose(sg, ss, rtol=2e-5, atol=1e-6)
    same = (rg == rs)
    assert float(same.float().mean()) > 0.98
    assert bool(((sg - ss).abs()[~same] < 1e-5).all())
    # recompute the cosine of the returned rows in fp32 with torch
    cand = g[rg.reshape(-1)].float().view(8, 100, DIM)
    qq = q[:8]
    cos = torch.einsum("bkd,bd->bk", cand, qq) / (cand.norm(dim=2) * qq.norm(dim=1, keepdim=True))
    assert torch.allclose(cos, sg, rtol=2e-5, atol=1e-6)
    # nothing outside the returned set beats the k-th score on a random sample of rows
    idx = torch.randint(0, ROWS, (200_000,), device="cuda")
    sample = g[idx].float()
    sc = (sample @ qq.T) / (sample.norm(dim=1, keepdim=True) * qq.norm(dim=1)[None, :])
    kth = sg[:, -1]
    viol = (sc > kth[None, :] + 1e-5)
    for j in range(8):
        assert set(idx[viol[:, j]].tolist()) <= set(rg[j].tolist())


def test_shard_merge_equals_single_and_exclusion(big):
    from multi_modal_retrieval_predict_project_b200 import B200RetrievalEngine
    from multi_modal_retrieval_predict_project_b200.sharded import merge_topk
    eng, q, g = big["eng"], big["q"], big["g"]
    r, s = eng.search(q[:64], 100)
    cut = 6_000_000
    a = B200RetrievalEngine.from_arrays(g[:cut], dtype="bfloat16", device=0, borrow=True, keep_host=False)
    b = B200RetrievalEngine.from_arrays(g[cut:], dtype="bfloat16", device=0, borrow=True, keep_host=False, row_offset=cut)
    ra, sa = a.search(q[:64], 100)
    rb, sb = b.search(q[:64], 100)
    mr, ms = merge_topk(torch.stack([sa, sb]), torch.stack([ra, rb]), 100)
    assert torch.equal(mr, r) and torch.equal(ms, s)
    # merging the merged list with itself changes nothing (idempotence)
    mr2, ms2 = merge_topk(torch.stack([ms, ms]), torch.stack([mr, torch.full_like(mr, -1)]), 100)
    assert torch.equal(mr2, r) and torch.equal(ms2, s)
    # excluding the best row shifts the rest up by one
    rx, sx = eng.search(q[:64], 99, exclude_rows=r[:, 0].contiguous())
    assert torch.equal(rx, r[:, 1:]) and torch.equal(sx, s[:, 1:])
    a.close(); b.close()


def test_cfg2_and_cfg5_full_size_vs_oracle(big):
    """BASELINE configs[1] and [4] at their full sizes on the first 1M rows of the gallery: cfg2 = 1024 queries,
    top-100 (64 sampled queries against the brute-force oracle); cfg5 = 10k queries, top-100, then P@k / Recall@k /
    AP / RR / nDCG on the device against CSR relevance sets -- the metric table of 64 queries must be IDENTICAL to
    the oracle's (fp64, bit for bit), and the searched rows of those queries must match the brute force."""
    import bench
    from multi_modal_retrieval_predict_project_b200 import B200RetrievalEngine
    from multi_modal_retrieval_predict_project_b200.Helpers import metrics_from_rows
    from oracle import metrics as om
    from oracle.bruteforce import check_topk
    n, k = 1_000_000, 100
    g = big["g"][:n]
    eng = B200RetrievalEngine.from_arrays(g, dtype="bfloat16", device=0, borrow=True, keep_host=False)
    # cfg2
    q2 = bench.gen_queries(1024, DIM, torch.device("cuda", 0), seed=bench.SEED + 900_001)
    rows2, scores2 = eng.search(q2, k)
    plan = eng.last_plan()
    assert plan["algo"] == "gemm" and plan["pair"] and plan["variant"] == "short", plan
    sel = torch.arange(0, 1024, 16, device="cuda") + torch.arange(64, device="cuda") % 16
    ok, detail = check_topk(rows2[sel], scores2[sel], g, q2[sel], k)
    assert ok, detail
    # cfg5
    nq = 10_000
    q5 = bench.gen_queries(nq, DIM, torch.device("cuda", 0), seed=bench.SEED + 900_002)
    rows5, scores5 = eng.search(q5, k)
    rows_h = rows5.cpu().numpy()
    rng = np.random.default_rng(bench.SEED + 1)
    sizes = rng.integers(1, 201, size=nq)
    sets = [np.union1d(rng.choice(n, size=int(sizes[i]), replace=False), rows_h[i, (i % 3)::3]) for i in range(nq)]
    indptr = np.zeros(nq + 1, dtype=np.int64)
    indptr[1:] = np.cumsum([len(x) for x in sets])
    rel = np.concatenate(sets).astype(np.int64)
    tbl = metrics_from_rows(rows5, torch.from_numpy(indptr).cuda(), torch.from_numpy(rel).cuda(), k).cpu().numpy()
    chk = np.arange(0, nq, nq // 64)[:64]
    want = om.per_query_table([[int(x) for x in rows_h[i]] for i in chk], [sets[i].tolist() for i in chk], k)
    assert np.array_equal(tbl[chk], want)
    assert 0.2 < tbl[:, 0].mean() < 0.5 and tbl[:, 4].mean() > 0.5          # the relevance sets are not degenerate
    ok, detail = check_topk(rows5[torch.from_numpy(chk).cuda()], scores5[torch.from_numpy(chk).cuda()], g,
                            q5[torch.from_numpy(chk).cuda()], k)
    assert ok, detail
    eng.close()
