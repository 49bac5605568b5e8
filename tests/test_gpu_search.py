"""GPU parity: the CUDA search path (through the C ABI) vs the oracle and the golden vectors."""
import ctypes as C

import numpy as np
import pytest

from oracle import search as osr
from tests._fixtures import load_search_case

pytestmark = pytest.mark.gpu


def _engine(g, **kw):
    from multi_modal_retrieval_predict_project_b200 import B200RetrievalEngine
    return B200RetrievalEngine.from_arrays(g, device=0, **kw)


def _check(rows, scores, want_rows, want_scores, rtol, atol=1e-7):
    assert rows.shape == want_rows.shape
    for i in range(rows.shape[0]):
        ok, why = osr.topk_matches(rows[i], scores[i], want_rows[i], want_scores[i], rtol=rtol, atol=atol)
        assert ok, (i, why)


@pytest.mark.parametrize("name", ["gauss", "clustered"])
def test_fp32_search_matches_reference_golden(name):
    """ids bit-exact except exact-score ties, scores within 1e-5 relative (north_star)."""
    c = load_search_case(name)
    eng = _engine(c["gallery"])
    k = c["order_top"].shape[1]
    rows, scores = eng.search(c["queries"], k)
    want_rows = c["order_top"].astype(np.int64)
    want_scores = np.take_along_axis(c["sim"], want_rows, axis=1)
    _check(rows, scores, want_rows, want_scores, rtol=1e-5)
    # zero query (row 1) scores exactly 0 everywhere; the zero gallery row (7) scores exactly 0
    assert np.all(scores[1] == 0.0)
    # exact duplicates (rows 3, 11, 12) must come out in ascending row order when they tie
    for i in range(rows.shape[0]):
        pos = {int(r): j for j, r in enumerate(rows[i])}
        if all(r in pos for r in (3, 11, 12)):
            assert pos[3] < pos[11] < pos[12]


def test_fp32_cfg1_openi_scale_vs_oracle():
    """BASELINE cfg1: 7.5k x 1024 fp32 gallery, 1.5k queries, top-10 (parity case)."""
    from multi_modal_retrieval_predict_project_b200 import synth
    g = synth.make_embeddings(7500, 1024, seed=synth.SEED)
    q = synth.make_embeddings(1500, 1024, seed=synth.SEED + 1)
    want_rows, want_scores = osr.exact_topk(q, g, 10)
    eng = _engine(g)
    rows, scores = eng.search(q, 10)
    _check(rows, scores, want_rows, want_scores, rtol=1e-5)
    assert (rows == want_rows).mean() > 0.999


@pytest.mark.parametrize("algo", ["scan", "gemm"])
@pytest.mark.parametrize("b,k", [(1, 10), (3, 100), (8, 7), (64, 100), (200, 10)])
def test_bf16_search_vs_oracle(algo, b, k):
    """bf16 storage: the oracle consumes the SAME bf16-rounded values upcast to fp32."""
    from multi_modal_retrieval_predict_project_b200 import synth
    n, d = 30000, 512
    g = osr.to_bf16_round(synth.make_embeddings(n, d, seed=11))
    q = osr.to_bf16_round(synth.make_embeddings(b, d, seed=12))
    eng = _engine(g, dtype="bfloat16")
    rows, scores = eng.search(q, k, algo=algo)
    want_rows, want_scores = osr.exact_topk(q, g, k)
    _check(rows, scores, want_rows, want_scores, rtol=2e-5, atol=1e-6)
    # exact recall@K against the fp64 ranking, epsilon rule at the K boundary
    r64, s64 = osr.exact_topk_f64(q, g, k)
    for i in range(b):
        missing = set(r64[i].tolist()) - set(rows[i].tolist())
        for m in missing:
            j = int(np.nonzero(r64[i] == m)[0][0])
            assert s64[i, j] - s64[i, -1] <= 1e-5, (i, m)


@pytest.mark.parametrize("algo", ["scan", "gemm"])
def test_unrounded_queries_are_rounded_to_bf16(algo):
    from multi_modal_retrieval_predict_project_b200 import synth
    g = synth.make_embeddings(5000, 256, seed=21)
    q = synth.make_embeddings(20, 256, seed=22)
    eng = _engine(g, dtype="bfloat16")
    rows, scores = eng.search(q, 10, algo=algo)
    want_rows, want_scores = osr.exact_topk(osr.to_bf16_round(q), osr.to_bf16_round(g), 10)
    _check(rows, scores, want_rows, want_scores, rtol=2e-5, atol=1e-6)


@pytest.mark.parametrize("dtype", ["float32", "bfloat16"])
@pytest.mark.parametrize("n,d,b,k", [(1, 64, 1, 5), (7, 96, 3, 10), (100, 100, 5, 100), (513, 40, 2, 1),
                                     (1000, 1024, 4, 17), (300, 130, 33, 300)])
def test_edge_shapes(dtype, n, d, b, k):
    """K > N (tail is -1/-inf), D not a multiple of 64, tiny galleries, ragged batch sizes."""
    from multi_modal_retrieval_predict_project_b200 import synth
    g = synth.make_embeddings(n, d, seed=31)
    q = synth.make_embeddings(b, d, seed=32)
    if dtype == "bfloat16":
        g, q = osr.to_bf16_round(g), osr.to_bf16_round(q)
    eng = _engine(g, dtype=dtype)
    rows, scores = eng.search(q, k)
    kk = min(k, n)
    want_rows, want_scores = osr.exact_topk(q, g, k)
    _check(rows[:, :kk], scores[:, :kk], want_rows, want_scores, rtol=2e-5, atol=1e-6)
    assert np.all(rows[:, kk:] == -1) and np.all(np.isneginf(scores[:, kk:]))


def test_exclude_rows_and_row_offset():
    from multi_modal_retrieval_predict_project_b200 import synth
    g = synth.make_embeddings(2000, 128, seed=41)
    eng = _engine(g, row_offset=5000)
    q = g[:50]
    rows, scores = eng.search(q, 5)
    assert np.array_equal(rows[:, 0], np.arange(50) + 5000)          # self match first
    rows2, scores2 = eng.search(q, 5, exclude_rows=np.arange(50) + 5000)
    assert not np.any(rows2 == (np.arange(50) + 5000)[:, None])
    assert np.array_equal(rows2[:, :4], rows[:, 1:])                  # rest shifts up by one
    # link graph == oracle link graph (Retrieval/retrieval.py:121-138 form)
    eng0 = _engine(g[:400])
    graph = eng0.build_link_graph(threshold=0.05, max_links=6)
    want = osr.build_link_graph(g[:400], 0.05, 6)
    sim = osr.cosine_similarity(g[:400])
    for i, (a, b_) in enumerate(zip(graph, want)):
        if a != b_:
            assert len(a) == len(b_) and np.allclose(sim[i, a], sim[i, b_], atol=2e-6), (i, a, b_)


def test_retrieve_signature_and_types(tmp_path):
    """The reference contract: retrieve(q (D,)|(1,D), K) -> (List[str], List[float]) best first."""
    from multi_modal_retrieval_predict_project_b200 import make_retrieval_engine, synth
    g = synth.make_embeddings(300, 64, seed=51)
    ids = synth.make_ids(300, "rec")
    fp, ip = synth.write_gallery(str(tmp_path), "train", g, ids)
    eng = make_retrieval_engine(fp, ip, method="DLS", link_threshold=0.3, max_links=8)   # kwargs ignored
    want_rows, want_scores = osr.exact_topk(g[5:6] * 2.0, g, 5)
    for q in (g[5] * 2.0, (g[5] * 2.0).reshape(1, -1), (g[5] * 2.0).astype(np.float64)):
        out_ids, out_scores = eng.retrieve(q, K=5, seed_size=5, max_steps=100, seed=2709)
        assert isinstance(out_ids, list) and isinstance(out_scores, list)
        assert all(isinstance(x, str) for x in out_ids) and all(isinstance(x, float) for x in out_scores)
        assert out_ids == [ids[int(r)] for r in want_rows[0]] and out_ids[0] == "rec5"
        assert np.allclose(out_scores, want_scores[0], rtol=1e-5)
    nested_ids, nested_scores = eng.retrieve(g[:3], K=4)
    assert len(nested_ids) == 3 and [x[0] for x in nested_ids] == ["rec0", "rec1", "rec2"]
    assert eng.retrieve(g[0], K=1000)[0].__len__() == 300                      # K > N -> N results
    assert np.array_equal(eng.get_embeddings_for_ids(["rec3", "nope"]), np.vstack([g[3], np.zeros(64, np.float32)]))
    with pytest.raises(ValueError, match="Unknown retrieval method"):
        make_retrieval_engine(fp, ip, method="faiss")
    with pytest.raises(ValueError):
        eng.search(np.zeros((1, 65), np.float32), 3)


def test_device_tensors_in_and_out_and_c_abi_direct():
    import torch
    from multi_modal_retrieval_predict_project_b200 import _lib, synth
    g = osr.to_bf16_round(synth.make_embeddings(10000, 512, seed=61))
    q = osr.to_bf16_round(synth.make_embeddings(4, 512, seed=62))
    gd = torch.from_numpy(g).cuda().to(torch.bfloat16)
    eng = _engine(gd, dtype="bfloat16", borrow=True, keep_host=False)
    rows, scores = eng.search(torch.from_numpy(q).cuda(), 10)
    assert rows.is_cuda and scores.is_cuda
    want_rows, want_scores = osr.exact_topk(q, g, 10)
    _check(rows.cpu().numpy(), scores.cpu().numpy(), want_rows, want_scores, rtol=2e-5, atol=1e-6)
    # virtual ids + device gather
    assert eng.retrieve(q[0], K=3)[0] == [f"g{int(r)}" for r in want_rows[0][:3]]
    assert np.array_equal(eng.get_embeddings_for_ids(["g7", "zzz"]), np.vstack([g[7], np.zeros(512, np.float32)]))
    # raw C ABI with host buffers
    lib = _lib.load()
    out_s = np.empty((4, 10), np.float32); out_r = np.empty((4, 10), np.int64)
    st = lib.mmr_search(eng._handle, q.ctypes.data, 4, _lib.MMR_F32, 10, _lib.ALGO_SCAN, None,
                        out_s.ctypes.data, out_r.ctypes.data, None)
    assert st == 0, lib.mmr_last_error()
    _check(out_r, out_s, want_rows, want_scores, rtol=2e-5, atol=1e-6)
    assert lib.mmr_search(eng._handle, q.ctypes.data, 4, _lib.MMR_F32, 0, 0, None, out_s.ctypes.data,
                          out_r.ctypes.data, None) == _lib.MMR_EINVAL
    assert b"k >= 1" in lib.mmr_last_error()


@pytest.mark.parametrize("algo", ["scan", "gemm"])
def test_merge_topk_matches_single_shard(algo):
    """Row shards + K-way merge == one index (the multi-GPU data path, emulated on one device)."""
    import torch
    from multi_modal_retrieval_predict_project_b200 import _lib, synth
    from multi_modal_retrieval_predict_project_b200.sharded import merge_topk
    g = osr.to_bf16_round(synth.make_embeddings(9001, 256, seed=71))
    q = osr.to_bf16_round(synth.make_embeddings(37, 256, seed=72))
    full = _engine(g, dtype="bfloat16")
    rows, scores = full.search(q, 50, algo=algo)
    bounds = [0, 2500, 2500 + 3000, 9001]
    parts_r, parts_s = [], []
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        e = _engine(g[lo:hi], dtype="bfloat16", row_offset=lo)
        r, s = e.search(torch.from_numpy(q).cuda(), 50, algo=algo)
        parts_r.append(r); parts_s.append(s)
    mr, ms, src = merge_topk(torch.stack(parts_s), torch.stack(parts_r), 50, want_src=True)
    assert np.array_equal(mr.cpu().numpy(), rows) and np.array_equal(ms.cpu().numpy(), scores)
    flat_r = torch.stack(parts_r).permute(1, 0, 2).reshape(37, -1)
    assert torch.equal(torch.gather(flat_r, 1, src.long()), mr)


@pytest.mark.parametrize("n,d,b,k", [(3000, 40, 20, 10), (5000, 100, 130, 33), (6000, 1024, 40, 100), (2000, 768, 5, 200)])
def test_gemm_other_dims(n, d, b, k):
    """GEMM path beyond the headline shape: padded D, the streamed-Q variant (D = 1024 does not fit the
    resident-Q shared-memory plan), K > 128 (larger candidate buffers), multiple query tiles."""
    from multi_modal_retrieval_predict_project_b200 import synth
    g = osr.to_bf16_round(synth.make_embeddings(n, d, seed=81))
    q = osr.to_bf16_round(synth.make_embeddings(b, d, seed=82))
    eng = _engine(g, dtype="bfloat16")
    rows, scores = eng.search(q, k, algo="gemm")
    want_rows, want_scores = osr.exact_topk(q, g, k)
    _check(rows, scores, want_rows, want_scores, rtol=2e-5, atol=1e-6)


@pytest.mark.parametrize("n,d,b,k", [(9000, 512, 300, 100), (9000, 1024, 300, 10), (100, 64, 257, 20),
                                     (129, 512, 385, 100), (70000, 512, 1500, 100)])
def test_gemm_cta_pairs(n, d, b, k):
    """cta_group::2 path (>= 2 query tiles): odd numbers of query tiles (the last pair runs a CTA with no
    valid query), the streamed-Q pair variant (D = 1024), galleries smaller than one 256-row tile (the
    peer CTA's half of the tile is entirely out of bounds) and a many-wave batch."""
    from multi_modal_retrieval_predict_project_b200 import synth
    g = osr.to_bf16_round(synth.make_embeddings(n, d, seed=83))
    q = osr.to_bf16_round(synth.make_embeddings(b, d, seed=84))
    eng = _engine(g, dtype="bfloat16")
    rows, scores = eng.search(q, k, algo="gemm")
    want_rows, want_scores = osr.exact_topk(q, g, k)
    _check(rows, scores, want_rows, want_scores, rtol=2e-5, atol=1e-6)


def test_gemm_duplicates_and_zero_rows():
    """Exact-score ties (duplicated gallery rows, reference Trainner/train.py:438-454 samples with
    replacement) resolve to ascending row id; zero rows / zero queries score exactly 0."""
    from multi_modal_retrieval_predict_project_b200 import synth
    g = osr.to_bf16_round(synth.make_embeddings(4096, 128, seed=91))
    g[100:400] = g[7]            # 300 copies of one row: more ties than k
    g[1000] = 0.0
    q = osr.to_bf16_round(synth.make_embeddings(17, 128, seed=92))
    q[3] = g[7]
    q[5] = 0.0
    eng = _engine(g, dtype="bfloat16")
    for algo in ("scan", "gemm"):
        rows, scores = eng.search(q, 50, algo=algo)
        assert rows[3, 0] == 7 and np.array_equal(rows[3, 1:50], np.arange(100, 149)), algo
        assert np.all(scores[5] == 0.0) and np.array_equal(rows[5], np.arange(50)), algo
        want_rows, want_scores = osr.exact_topk(q, g, 50)
        _check(rows, scores, want_rows, want_scores, rtol=2e-5, atol=1e-6)


def test_gemm_large_k_generic_compaction():
    """k > 256 uses 4096-entry candidate lists and the generic (re-reading) compaction path."""
    from multi_modal_retrieval_predict_project_b200 import synth
    g = osr.to_bf16_round(synth.make_embeddings(30000, 128, seed=101))
    q = osr.to_bf16_round(synth.make_embeddings(20, 128, seed=102))
    eng = _engine(g, dtype="bfloat16")
    for k in (300, 1000):
        rows, scores = eng.search(q, k, algo="gemm")
        want_rows, want_scores = osr.exact_topk(q, g, k)
        _check(rows, scores, want_rows, want_scores, rtol=2e-5, atol=1e-6)


# ---------------------------------------------------------------------------------------------------
# Both GEMM instantiations, pinned per engine (mmr_index_tune), on launches long enough (>= 200 gallery
# tiles per part) that lists fill and compact, the cross-list bound is refreshed many times and the final
# per-list pass has real work -- the regime of the headline launch (4341 tiles per part at 10M rows).
# ---------------------------------------------------------------------------------------------------
def _tuned_search(eng, q, k, variant, parts, pair=True, min_tiles=200):
    eng.tune(variant=variant, parts=parts, pair=pair)
    rows, scores = eng.search(q, k, algo="gemm")
    plan = eng.last_plan()
    assert plan["algo"] == "gemm" and plan["variant"] == variant, plan
    assert plan["tiles_per_part"] >= min_tiles, plan
    return rows, scores, plan


@pytest.mark.parametrize("variant", ["long", "short"])
@pytest.mark.parametrize("n,d,b,k,parts", [(120_000, 128, 300, 100, 2),    # 2 CTA pairs x 2 parts, odd tile count
                                           (150_001, 64, 100, 100, 1),     # single-CTA instantiation, ragged last tile
                                           (260_000, 512, 700, 100, 4),    # resident-Q pairs at d = 512, 3 pair units
                                           (215_000, 64, 300, 100, 1),     # > 768 tiles per part: the 12-tile pacing window
                                           (110_000, 1024, 260, 10, 2)])   # streamed-Q pairs
def test_gemm_variants_many_tiles_per_part(variant, n, d, b, k, parts):
    from multi_modal_retrieval_predict_project_b200 import synth
    g = osr.to_bf16_round(synth.make_embeddings(n, d, seed=131))
    q = osr.to_bf16_round(synth.make_embeddings(b, d, seed=132))
    eng = _engine(g, dtype="bfloat16", keep_host=False)
    rows, scores, plan = _tuned_search(eng, q, k, variant, parts)
    assert plan["pair"] == (b > 128) and plan["parts"] == parts
    want_rows, want_scores = osr.exact_topk(q, g, k)
    _check(rows, scores, want_rows, want_scores, rtol=2e-5, atol=1e-6)
    # the two instantiations agree bit for bit (same accumulation order, same ordering rule)
    other = "short" if variant == "long" else "long"
    rows2, scores2, _ = _tuned_search(eng, q, k, other, parts)
    assert np.array_equal(rows, rows2) and np.array_equal(scores, scores2)


@pytest.mark.parametrize("variant", ["long", "short"])
@pytest.mark.parametrize("pair", [True, False])
def test_gemm_variants_duplicates_across_compactions(variant, pair):
    """More exact-score ties than k, spread over gallery tiles that sit hundreds of tiles apart (so the
    copies reach a list before and after several compactions, and through both warpgroups' lists):
    the winners are the lowest row ids; zero rows / zero queries score exactly 0."""
    from multi_modal_retrieval_predict_project_b200 import synth
    n, d, b, k = 70_000, 128, 200, 50
    g = osr.to_bf16_round(synth.make_embeddings(n, d, seed=141))
    dup = np.r_[100:130, 20_000:20_030, 45_000:45_040, 69_950:69_990]
    g[dup] = g[7]
    g[1000] = 0.0
    q = osr.to_bf16_round(synth.make_embeddings(b, d, seed=142))
    q[3] = g[7]
    q[5] = 0.0
    q[130] = g[7] * 2.0            # second query tile
    eng = _engine(g, dtype="bfloat16", keep_host=False)
    rows, scores, plan = _tuned_search(eng, q, k, variant, 1, pair=pair)
    assert plan["pair"] == pair and plan["parts"] == 1
    for qi in (3, 130):
        assert rows[qi, 0] == 7 and np.array_equal(rows[qi, 1:50], np.sort(dup)[:49]), (variant, qi)
    assert np.all(scores[5] == 0.0) and np.array_equal(rows[5], np.arange(50))
    want_rows, want_scores = osr.exact_topk(q, g, k)
    _check(rows, scores, want_rows, want_scores, rtol=2e-5, atol=1e-6)


@pytest.mark.parametrize("variant", ["long", "short"])
def test_gemm_variants_large_k(variant):
    """k > 128 / > 256 (1024- and 4096-entry lists, generic compaction) on >= 200 tiles per part."""
    from multi_modal_retrieval_predict_project_b200 import synth
    g = osr.to_bf16_round(synth.make_embeddings(120_000, 128, seed=151))
    q = osr.to_bf16_round(synth.make_embeddings(20, 128, seed=152))
    eng = _engine(g, dtype="bfloat16", keep_host=False)
    for k in (200, 300, 1000):
        rows, scores, _ = _tuned_search(eng, q, k, variant, 2)
        want_rows, want_scores = osr.exact_topk(q, g, k)
        _check(rows, scores, want_rows, want_scores, rtol=2e-5, atol=1e-6)


def test_exclusion_at_max_k_and_tune_errors():
    """exclude_rows with k = MMR_MAX_K keeps k + 1 candidates inside the GEMM (ADVICE: it used to be
    rejected); unknown knobs / values are MMR_EINVAL."""
    from multi_modal_retrieval_predict_project_b200 import _lib, synth
    g = osr.to_bf16_round(synth.make_embeddings(5000, 64, seed=161))
    eng = _engine(g, dtype="bfloat16")
    q = g[:6]
    ex = np.arange(6, dtype=np.int64)
    rows, scores = eng.search(q, 1024, exclude_rows=ex, algo="gemm")
    sim = osr.cosine_similarity(q, g)
    sim[np.arange(6), ex] = -np.inf
    for i in range(6):
        order = np.lexsort((np.arange(5000), -sim[i].astype(np.float64)))[:1024]
        ok, why = osr.topk_matches(rows[i], scores[i], order, sim[i, order], rtol=2e-5, atol=1e-6)
        assert ok, (i, why)
    lib = _lib.load()
    assert lib.mmr_index_tune(eng._handle, 99, 0) == _lib.MMR_EINVAL
    assert lib.mmr_index_tune(eng._handle, _lib.TUNE_GEMM_VARIANT, 7) == _lib.MMR_EINVAL
    with pytest.raises(KeyError):
        eng.tune(variant="medium")


@pytest.mark.parametrize("n,parts,pair,k,exclude", [(9000, 12, 0, 128, False), (9000, 12, 1, 128, True),
                                                    (20000, 30, 1, 100, False), (3000, 16, 0, 127, True)])
def test_selection_over_many_full_lists(n, parts, pair, k, exclude):
    """The warp-per-query selection with more valid candidates than a warp holds: many gallery parts of a few
    hundred rows each leave up to k (k + 1 with an exclusion) entries per list and no published pruning bound, so
    the warp cuts its 512-key buffer back to the top-k several times per query (csrc/select.cu
    select_warp_kernel / warp_keep_topk).  Duplicated gallery rows put exact score ties across the cuts."""
    from multi_modal_retrieval_predict_project_b200 import synth
    g = osr.to_bf16_round(synth.make_embeddings(n, 64, seed=181 + parts, clustered=True))
    g[n // 2:n // 2 + 200] = g[:200]                      # exact duplicates in another part
    q = osr.to_bf16_round(synth.make_embeddings(70, 64, seed=182, clustered=True))
    q[:8] = g[:8]
    eng = _engine(g, dtype="bfloat16")
    try:
        eng.tune(parts=parts, pair=pair)
    except NotImplementedError:
        pytest.skip("plan not available")
    ex = np.arange(70, dtype=np.int64) if exclude else None
    rows, scores = eng.search(q, k, exclude_rows=ex, algo="gemm")
    plan = eng.last_plan()
    assert plan["algo"] == "gemm" and plan["parts"] >= 2
    sim = osr.cosine_similarity(q, g)
    if exclude:
        sim[np.arange(70), ex] = -np.inf
    for i in range(70):
        order = np.lexsort((np.arange(n), -sim[i].astype(np.float64)))[:k]
        ok, why = osr.topk_matches(rows[i], scores[i], order, sim[i, order], rtol=2e-5, atol=1e-6)
        assert ok, (i, why)
        assert len(set(rows[i].tolist())) == k


def test_empty_shard_returns_padding():
    """A shard without rows (more ranks than gallery tiles: sharded.shard_bounds / weighted_shard_bounds can hand a
    rank an empty block) answers every query with -1 / -inf padding, which the cross-shard merge ignores."""
    import torch
    from multi_modal_retrieval_predict_project_b200 import synth
    from multi_modal_retrieval_predict_project_b200.sharded import merge_topk
    eng = _engine(np.zeros((0, 64), dtype=np.float32), dtype="bfloat16", row_offset=1234)
    q = osr.to_bf16_round(synth.make_embeddings(5, 64, seed=191))
    rows, scores = eng.search(q, 7)
    assert rows.shape == (5, 7) and np.all(rows == -1) and np.all(np.isneginf(scores))
    qd = torch.from_numpy(q).cuda()
    r_d, s_d = eng.search(qd, 7)
    assert bool((r_d == -1).all()) and bool(torch.isneginf(s_d).all())
    g = osr.to_bf16_round(synth.make_embeddings(300, 64, seed=192))
    full = _engine(g, dtype="bfloat16")
    r_f, s_f = full.search(qd, 7)
    m_r, m_s = merge_topk(torch.stack([s_f, s_d]), torch.stack([r_f, r_d]), 7)
    assert torch.equal(m_r, r_f) and torch.equal(m_s, s_f)


def test_two_streams_one_handle_are_ordered_on_the_device():
    """Two streams (and two host threads) searching through ONE handle with device outputs: the handle's
    workspaces are shared, so the library orders the calls on the device (an event recorded at the end
    of each call, waited on by the next call's stream) -- results must equal the serial ones."""
    import threading
    import torch
    from multi_modal_retrieval_predict_project_b200 import synth
    g = osr.to_bf16_round(synth.make_embeddings(200_000, 128, seed=171))
    eng = _engine(g, dtype="bfloat16", keep_host=False)
    qs = [torch.from_numpy(osr.to_bf16_round(synth.make_embeddings(512, 128, seed=172 + i))).cuda() for i in range(2)]
    want = [tuple(t.clone() for t in eng.search(q, 20)) for q in qs]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    got = [[None] * 6, [None] * 6]

    def worker(i):
        with torch.cuda.stream(streams[i]):
            for it in range(6):
                got[i][it] = eng.search(qs[i], 20)

    th = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    torch.cuda.synchronize()
    for i in range(2):
        for r, s_ in got[i]:
            assert torch.equal(r, want[i][0]) and torch.equal(s_, want[i][1])


def test_merged_engine_from_split_dumps(tmp_path):
    """Helpers.merged_engine: train + val dumps -> one HBM-resident gallery, rows and ids in the order
    createDumpEmbedding writes them (reference Helpers/dumpEmbedding.py:28-39)."""
    import json
    from multi_modal_retrieval_predict_project_b200 import synth
    from multi_modal_retrieval_predict_project_b200.Helpers import merged_engine
    tr, va = synth.make_embeddings(300, 64, seed=111), synth.make_embeddings(120, 64, seed=112)
    np.save(tmp_path / "train_joint_embeddings.npy", tr)
    np.save(tmp_path / "val_joint_embeddings.npy", va)
    json.dump([f"t{i}" for i in range(300)], open(tmp_path / "train_ids.json", "w"))
    json.dump([f"v{i}" for i in range(120)], open(tmp_path / "val_ids.json", "w"))
    eng = merged_engine(tmp_path, device=0)
    assert eng.n == 420 and eng.ids[299] == "t299" and eng.ids[300] == "v0"
    ids, scores = eng.retrieve(va[5], K=3)
    assert ids[0] == "v5" and abs(scores[0] - 1.0) < 1e-5
    want_rows, want_scores = osr.exact_topk(va[5:6], np.concatenate([tr, va]), 3)
    assert ids == [eng.ids[int(r)] for r in want_rows[0]]


@pytest.mark.parametrize("dtype", ["bfloat16", "float32"])
def test_blob_round_trip_is_bit_identical(tmp_path, dtype):
    """save_blob / load_blob (persisted device-format gallery): the reloaded index
    returns bit-identical rows and scores; a bf16 blob holds the rounded values as raw bf16 bits."""
    from multi_modal_retrieval_predict_project_b200 import B200RetrievalEngine, synth
    g = synth.make_embeddings(3000, 100, seed=121)
    q = synth.make_embeddings(37, 100, seed=122)
    ids = [f"r{i}" for i in range(3000)]
    eng = B200RetrievalEngine.from_arrays(g, ids=ids, dtype=dtype, device=0)
    path = eng.save_blob(str(tmp_path / "gallery"))
    eng2 = B200RetrievalEngine.load_blob(path, device=0)
    assert eng2.n == 3000 and eng2.dim == 100 and eng2.dtype == dtype and eng2.ids[7] == "r7"
    for algo in (("scan", "gemm") if dtype == "bfloat16" else ("scan",)):
        r1, s1 = eng.search(q, 20, algo=algo)
        r2, s2 = eng2.search(q, 20, algo=algo)
        assert np.array_equal(r1, r2) and np.array_equal(s1, s2), algo
    z = np.load(path)
    assert z["data"].dtype == (np.uint16 if dtype == "bfloat16" else np.float32) and z["data"].shape == (3000, 100)
    if dtype == "bfloat16":
        assert np.array_equal(z["data"], (osr.to_bf16_round(g).view(np.uint32) >> 16).astype(np.uint16))
    ids_a, sc_a = eng.retrieve(g[5], K=3)
    ids_b, sc_b = eng2.retrieve(g[5], K=3)
    assert ids_a == ids_b and sc_a == sc_b
