"""GPU parity: rerank + metrics kernels (through the C ABI) vs the golden vectors and the oracle."""
import json

import numpy as np
import pytest

from oracle import gt as ogt
from oracle import metrics as om
from oracle import rerank as orr
from oracle import search as osr
from tests._fixtures import GOLDEN, assert_rerank_close, build_rerank_artifacts

pytestmark = pytest.mark.gpu

# fp32 cosines computed with a different summation order than numpy's, then min-max scaled
# (amplified by 1/(max-min)): tolerance on the fp64 combined values, stated per north_star (1e-5).
TOL = dict(rtol=1e-5, atol=5e-6)


def test_rerank_matches_reference_golden(tmp_path):
    from multi_modal_retrieval_predict_project_b200 import Reranker
    a = build_rerank_artifacts(str(tmp_path))
    for name, v in a["golden"]["variants"].items():
        al, be, ga = v["weights"]
        rer = Reranker(a["kg_dir"], a["csv"], al, be, ga, device=0)
        for qi, res in enumerate(v["results"]):
            cand = res["cand"]
            cand_ids = [a["ids"][j] for j in cand]
            embs = a["g"][cand]
            r1 = rer.rerank(a["qids"][qi], cand_ids, candidate_embs=embs, query_emb=a["qs"][qi], topk=10)
            assert_rerank_close(r1, [tuple(t) for t in res["qemb_top10"]], **TOL)
            lookup = {cid: a["g"][j] for cid, j in zip(cand_ids, cand)}
            lookup[a["qids"][qi]] = a["qs"][qi]
            r2 = rer.rerank(a["qids"][qi], cand_ids, candidate_embs=embs, candidate_emb_lookup=lookup)
            assert_rerank_close(r2, [tuple(t) for t in res["lookup_all"]], **TOL)
            r3 = rer.rerank(cand_ids[3], cand_ids, candidate_embs=embs, topk=5)
            assert_rerank_close(r3, [tuple(t) for t in res["gallery_qid_top5"]], **TOL)
            r4 = rer.rerank("not-a-record", cand_ids, candidate_embs=embs, query_emb=a["qs"][qi], topk=5)
            assert_rerank_close(r4, [tuple(t) for t in res["unknown_qid_top5"]], **TOL)
            assert all(isinstance(x, float) for t in r1 for x in t[1:]) and isinstance(r1[0][0], str)
        rer.close()


def test_rerank_errors_match_reference(tmp_path):
    from multi_modal_retrieval_predict_project_b200 import Reranker
    a = build_rerank_artifacts(str(tmp_path))
    rer = Reranker(a["kg_dir"], a["csv"], device=0)
    with pytest.raises(ValueError, match="Please provide candidate_embs"):
        rer.rerank("t0", ["g1", "g2"])
    with pytest.raises(ValueError, match="rows must match"):
        rer.rerank("t0", ["g1", "g2"], candidate_embs=a["g"][:3], query_emb=a["qs"][0])
    with pytest.raises(ValueError, match="Query embedding not found"):
        rer.rerank("t0", ["g1", "g2"], candidate_embs=a["g"][:2])
    with pytest.raises(FileNotFoundError):
        Reranker(str(tmp_path / "missing"), a["csv"], device=0)
    # the reference's default arguments resolve relative to its checkout (reranker.py:11-15, 38-39); here relative
    # to $MMR_B200_BASE_DIR / the working directory: missing files raise like the reference, present ones load
    import os
    import shutil
    with pytest.raises(FileNotFoundError, match="node2id.json"):
        Reranker(device=0)
    base = tmp_path / "checkout"
    shutil.copytree(a["kg_dir"], base / "knowledge_graph")
    os.makedirs(base / "outputs")
    shutil.copy(a["csv"], base / "outputs" / "openi_labels_final.csv")
    os.environ["MMR_B200_BASE_DIR"] = str(base)
    try:
        r0 = Reranker(device=0)
        assert r0.get_record_label_set("g0") == rer.get_record_label_set("g0")
        r0.close()
    finally:
        del os.environ["MMR_B200_BASE_DIR"]
    # label sets / kg vectors agree with the oracle's python-loop versions
    ora = orr.OracleReranker(a["kg_dir"], a["csv"])
    for rid in ["g0", "g5", "g6", "g13", "t2", "t3", "zzz"]:
        assert rer.get_record_label_set(rid) == ora.get_record_label_set(rid)
        assert np.allclose(rer.get_record_kg_vec(rid), ora.get_record_kg_vec(rid), atol=1e-6)


def test_engine_retrieve_with_reranker_vs_oracle(tmp_path):
    """retrieve(q, K, reranker=..., query_id=..., rerank_topk=...) (reference retrieval.py:246-269):
    exact candidates -> rerank; the stored gallery row replaces q when query_id is a gallery id."""
    from multi_modal_retrieval_predict_project_b200 import Reranker, make_retrieval_engine
    a = build_rerank_artifacts(str(tmp_path))
    eng = make_retrieval_engine(a["features_path"], a["ids_path"], method="b200")
    rer = Reranker(a["kg_dir"], a["csv"], device=0)
    ora = orr.OracleReranker(a["kg_dir"], a["csv"])
    for qi in range(len(a["qids"])):
        for qid, qvec in ((a["qids"][qi], a["qs"][qi]), (a["ids"][qi], a["qs"][qi])):
            ids, scores = eng.retrieve(qvec, K=10, reranker=rer, query_id=qid, rerank_topk=5, seed=1)
            wr, _ = osr.exact_topk(qvec[None], a["g"], 10)
            cand_ids = [a["ids"][int(j)] for j in wr[0]]
            q_emb = a["g"][a["ids"].index(qid)] if qid in a["ids"] else qvec
            want = ora.rerank(qid, cand_ids, candidate_embs=a["g"][wr[0]], query_emb=q_emb, topk=5)
            got = list(zip(ids, scores))
            assert len(got) == 5
            assert_rerank_close([(i, s, 0, 0, 0) for i, s in got], [(t[0], t[1], 0, 0, 0) for t in want], **TOL)
    # a foreign reranker object with the reference interface works too (here: the oracle)
    ids, scores = eng.retrieve(a["qs"][0], K=10, reranker=ora, query_id=a["qids"][0])
    ids2, scores2 = eng.retrieve(a["qs"][0], K=10, reranker=rer, query_id=a["qids"][0])
    assert_rerank_close([(i, s, 0, 0, 0) for i, s in zip(ids2, scores2)],
                        [(i, s, 0, 0, 0) for i, s in zip(ids, scores)], **TOL)


@pytest.mark.parametrize("dtype", ["float32", "bfloat16"])
def test_multi_device_engine_reference_signature(tmp_path, dtype):
    """make_retrieval_engine(fp, ip, method="b200", devices=[...]).retrieve(q, K, reranker=, query_id=) -- the
    reference signature (string ids, Retrieval/retrieval.py:140-151) over row shards on several devices -- returns
    what the single-GPU engine returns, with and without a rerank, for gallery-id and external query ids, single
    and batched queries.  Uses every visible GPU; on a one-GPU box the shards share the device (three shards on
    device 0: same code path, same merge and cross-shard cosine sum)."""
    import torch
    from multi_modal_retrieval_predict_project_b200 import MultiGPURetrievalEngine, Reranker, make_retrieval_engine
    a = build_rerank_artifacts(str(tmp_path))
    ndev = torch.cuda.device_count()
    devices = list(range(ndev)) if ndev >= 2 else [0, 0, 0]
    one = make_retrieval_engine(a["features_path"], a["ids_path"], method="b200", dtype=dtype, device=0)
    many = make_retrieval_engine(a["features_path"], a["ids_path"], method="b200", dtype=dtype, devices=devices)
    assert isinstance(many, MultiGPURetrievalEngine) and len(many.shards) == len(devices)
    assert many.ids == one.ids and np.array_equal(many.get_embeddings_for_ids(["g3", "zz"]), one.get_embeddings_for_ids(["g3", "zz"]))
    rer = Reranker(a["kg_dir"], a["csv"], device=0)
    # bf16: both engines run the same (vectorised) feature kernel => identical; fp32: the single engine's gather
    # kernel and the cross-shard path's KG kernel sum in different orders (1e-8 on the combined score)
    atol = 1e-12 if dtype == "bfloat16" else 1e-7
    for q, qid in ((a["qs"][0], a["qids"][0]), (a["qs"][1].reshape(1, -1), "g17"), (a["g"][5], "g5")):
        assert many.retrieve(q, K=7) == one.retrieve(q, K=7)
        ids_m, sc_m = many.retrieve(q, K=20, reranker=rer, query_id=qid, rerank_topk=8)
        ids_1, sc_1 = one.retrieve(q, K=20, reranker=rer, query_id=qid, rerank_topk=8)
        assert ids_m == ids_1 and len(ids_m) == 8 and np.allclose(sc_m, sc_1, rtol=0, atol=atol)
        assert all(isinstance(x, str) for x in ids_m) and all(isinstance(x, float) for x in sc_m)
    nested_m = many.retrieve(a["qs"][:5], K=10, reranker=rer, query_id=list(a["qids"][:5]))
    nested_1 = one.retrieve(a["qs"][:5], K=10, reranker=rer, query_id=list(a["qids"][:5]))
    assert nested_m[0] == nested_1[0] and np.allclose(np.array(nested_m[1]), np.array(nested_1[1]), rtol=0, atol=atol)
    assert len(many.retrieve(a["qs"][0], K=1000)[0]) == len(a["ids"])              # K > N -> N results
    # the batched result equals the per-query one (one device call for the whole batch)
    for i in range(5):
        ids_i, sc_i = one.retrieve(a["qs"][i], K=10, reranker=rer, query_id=a["qids"][i])
        assert ids_i == nested_1[0][i] and sc_i == nested_1[1][i]
    with pytest.raises(ValueError):
        many.search(np.zeros((1, 3), np.float32), 3)
    many.close(); one.close(); rer.close()


def test_rerank_tie_order_is_candidate_position_ascending():
    """Exact ties in the combined score (the rule of the reference's own la_only variant: alpha = 0, beta = 1,
    gamma = 0 -- discrete Jaccard values) come out in ASCENDING candidate position, in the unfused kernels and
    in the fused tail alike.  The reference's np.argsort(final)[::-1] (reranker.py:327) leaves tie order to
    numpy's unstable sort (for <= 16 candidates it happens to be descending position); INTEGRATION.md
    documents the deterministic rule as a deliberate divergence."""
    import torch
    from multi_modal_retrieval_predict_project_b200 import Reranker
    n, k = 64, 24
    masks = np.zeros(n + 1, dtype=np.uint64)
    masks[:n] = np.array([0b0011, 0b0110, 0b0011, 0b1000], dtype=np.uint64)[np.arange(n) % 4]
    masks[n] = 0b0011                                            # the query's labels
    kg = np.zeros((n + 1, 8), dtype=np.float32)
    rer = Reranker.from_tables(masks, kg, alpha=0.0, beta=1.0, gamma=0.0, device=0)
    rows = torch.arange(k, device="cuda").flip(0).reshape(1, k).contiguous()          # candidates 23, 22, ..., 0
    scores = torch.linspace(0.9, 0.1, k, device="cuda").reshape(1, k).contiguous()
    q_rec = torch.tensor([n], device="cuda")
    ids, fin, s4 = rer.rerank_scored_device(rows, scores, q_rec, 0, want_scores4=True)
    order, sc = rer.rerank_with_cos_device(scores, q_rec, rows, 0)
    torch.cuda.synchronize()
    fin_h, ids_h = fin.cpu().numpy()[0], ids.cpu().numpy()[0]
    assert np.all(np.diff(fin_h) <= 0) and set(np.unique(fin_h)) == {0.0, 1.0 / 3.0, 1.0}
    pos = {int(r): j for j, r in enumerate(rows.cpu().numpy()[0])}                  # candidate position of every id
    for v in np.unique(fin_h):
        tied = [pos[int(i)] for i in ids_h[fin_h == v]]
        assert tied == sorted(tied) and len(tied) > 1
    assert np.array_equal(order.cpu().numpy()[0], [pos[int(i)] for i in ids_h]) and torch.equal(sc, s4)


def test_batched_device_rerank_vs_oracle():
    import torch
    from multi_modal_retrieval_predict_project_b200 import B200RetrievalEngine, Reranker, synth
    n, d, b, k = 5000, 256, 40, 100
    g = osr.to_bf16_round(synth.make_embeddings(n, d, seed=5, clustered=True))
    q = osr.to_bf16_round(synth.make_embeddings(b, d, seed=6, clustered=True))
    rng = np.random.default_rng(7)
    bits = (rng.random((n + b, 43)) < 0.1)
    masks = (bits.astype(np.uint64) << np.arange(43, dtype=np.uint64)).sum(axis=1).astype(np.uint64)
    kg = rng.standard_normal((n + b, 300)).astype(np.float32)
    kg = kg / (np.linalg.norm(kg, axis=1, keepdims=True) + 1e-12)
    eng = B200RetrievalEngine.from_arrays(g, dtype="bfloat16", device=0)
    rer = Reranker.from_tables(masks, kg, alpha=0.4, beta=0.2, gamma=0.2, device=0)
    qd = torch.from_numpy(q).cuda()
    rows, _ = eng.search(qd, k)
    order, sc = rer.rerank_device(eng, qd, rows, torch.arange(n, n + b, device="cuda"), rows.clone(), topk=20)
    order, sc, rows = order.cpu().numpy(), sc.cpu().numpy(), rows.cpu().numpy()
    for i in range(b):
        cand = rows[i]
        emb = [orr.safe_cos(q[i], g[j]) for j in cand]
        lab = [orr.jaccard_sets(set(np.nonzero(bits[n + i])[0]), set(np.nonzero(bits[j])[0])) for j in cand]
        kgs = [orr.safe_cos(kg[n + i], kg[j]) for j in cand]
        e, l, kk_ = (np.array(orr.minmax_scale_list(x)) for x in (emb, lab, kgs))
        final = 0.4 * e + 0.2 * l + 0.2 * kk_
        want = np.lexsort((np.arange(k), -final))[:20]
        got_final = sc[i, :, 0]
        assert np.allclose(got_final, final[want], **TOL), i
        assert np.allclose(sc[i, :, 2], l[order[i]], rtol=0, atol=0)          # Jaccard is exact
        mism = order[i] != want
        assert np.all(np.abs(final[order[i]][mism] - final[want][mism]) < 1e-5)


def _tables(n, b, d_kg, seed, p=0.1):
    rng = np.random.default_rng(seed)
    bits = (rng.random((n + b, 43)) < p)
    masks = (bits.astype(np.uint64) << np.arange(43, dtype=np.uint64)).sum(axis=1).astype(np.uint64)
    kg = rng.standard_normal((n + b, d_kg)).astype(np.float32)
    kg = kg / (np.linalg.norm(kg, axis=1, keepdims=True) + 1e-12)
    kg[7] = 0.0                                   # a record without a KG vector
    return masks, kg


@pytest.mark.parametrize("n,d,b,k,topk,d_kg", [(5000, 256, 40, 100, 20, 300), (3000, 128, 70, 10, 0, 48),
                                               (60, 64, 5, 100, 0, 300), (2000, 64, 9, 128, 128, 512)])
def test_fused_tail_vs_oracle_and_unfused_kernels(n, d, b, k, topk, d_kg):
    """mmr_rerank_scored (one kernel: features + min-max + combine + order, embedding feature = the search
    score): (a) against the oracle restatement of Reranker.rerank's scoring (reranker.py:298-329) fed the
    candidates' stored embeddings -- the oracle recomputes safe_cos, the kernel reuses the search score, same
    quantity; (b) bit-identical to the unfused kernels (rerank_features with the same cosines -> combine ->
    apply_order); (c) K > N: padded candidates are excluded from the min-max and come out as id -1."""
    import torch
    from multi_modal_retrieval_predict_project_b200 import B200RetrievalEngine, Reranker, _lib, synth
    g = osr.to_bf16_round(synth.make_embeddings(n, d, seed=5, clustered=True))
    q = osr.to_bf16_round(synth.make_embeddings(b, d, seed=6, clustered=True))
    masks, kg = _tables(n, b, d_kg, 7)
    w = (0.4, 0.2, 0.2)
    eng = B200RetrievalEngine.from_arrays(g, dtype="bfloat16", device=0)
    rer = Reranker.from_tables(masks, kg, alpha=w[0], beta=w[1], gamma=w[2], device=0)
    assert rer.fused_tail_ok(k)
    qd = torch.from_numpy(q).cuda()
    q_rec = torch.arange(n, n + b, device="cuda")
    rows, scores = eng.search(qd, k)
    ids, fin, s4 = rer.rerank_scored_device(rows, scores, q_rec, topk, want_scores4=True)
    keep = topk if 0 < topk < k else k
    valid = min(k, n)
    rows_h, ids_h, fin_h = rows.cpu().numpy(), ids.cpu().numpy(), fin.cpu().numpy()
    for i in range(b):
        cand = rows_h[i, :valid]
        want = orr.rerank_from_arrays(q[i], g[cand], masks[n + i], masks[cand], kg[n + i], kg[cand], *w, topk=keep)
        ok, why = orr.reranked_lists_match(ids_h[i, :min(keep, valid)], fin_h[i, :min(keep, valid)], cand, want)
        assert ok, (i, why)
        assert np.all(ids_h[i, valid:] == -1) and np.all(fin_h[i, valid:] == 0.0)
    # (b) the unfused kernels given the same cosines
    if valid == k:
        order, sc = rer.rerank_with_cos_device(scores, q_rec, rows, topk)
        ids2 = torch.empty_like(ids); fin2 = torch.empty_like(fin)
        lib = _lib.load()
        _lib.check(lib.mmr_apply_order(_lib.ptr(rows), _lib.ptr(order), _lib.ptr(sc), b, k, keep, _lib.ptr(ids2),
                                       _lib.ptr(fin2), 0, _lib.current_stream(0)))
        torch.cuda.synchronize()
        assert torch.equal(ids, ids2) and torch.equal(fin, fin2) and torch.equal(s4, sc)
    # host buffers through the raw C ABI
    lib = _lib.load()
    o_ids = np.empty((b, keep), np.int64); o_fin = np.empty((b, keep), np.float64)
    scores_h, q_rec_h = scores.cpu().numpy(), q_rec.cpu().numpy()
    st = lib.mmr_rerank_scored(rer._tables, rows_h.ctypes.data, scores_h.ctypes.data, q_rec_h.ctypes.data, b, k,
                               w[0], w[1], w[2], topk, o_ids.ctypes.data, o_fin.ctypes.data, None, 0, None)
    assert st == 0, lib.mmr_last_error()
    assert np.array_equal(o_ids, ids_h) and np.array_equal(o_fin, fin_h)
    assert lib.mmr_rerank_scored(rer._tables, rows_h.ctypes.data, scores_h.ctypes.data, None, b, 200,
                                 w[0], w[1], w[2], 0, o_ids.ctypes.data, o_fin.ctypes.data, None, 0, None) == _lib.MMR_EUNSUP


def test_exchange_kernels_on_one_gpu_equal_the_single_shard_path():
    """The fused exchange (mmr_search_scatter: selection kernel storing into the owner's region;
    mmr_exchange_rerank: wait + merge + rerank + publish in one kernel) with world = 1, driven through the raw
    C ABI: same (ids, scores) as search -> mmr_rerank_scored, over several steps (both buffer parities) and a
    ragged K (padding of the lists to a multiple of 4).  Also: size / state errors, abort, and a (b, k)
    mismatch is reported instead of hanging."""
    import ctypes as C
    import torch
    from multi_modal_retrieval_predict_project_b200 import B200RetrievalEngine, Reranker, _lib, synth
    n, d, b = 30000, 128, 300
    g = osr.to_bf16_round(synth.make_embeddings(n, d, seed=25))
    masks, kg = _tables(n, b, 300, 27)
    eng = B200RetrievalEngine.from_arrays(g, dtype="bfloat16", device=0)
    rer = Reranker.from_tables(masks, kg, device=0)
    lib = _lib.load()
    ex = C.c_void_p()
    _lib.check(lib.mmr_exchange_create(C.byref(ex), 0, 1, b, 100, 0))
    q_rec = torch.arange(n, n + b, device="cuda")
    step = 0
    for k, bb, topk, algo in ((100, 300, 0, "gemm"), (10, 300, 5, "gemm"), (50, 37, 0, "gemm"), (7, 3, 0, "scan"),
                              (100, 2, 30, "scan")):
        qd = torch.from_numpy(osr.to_bf16_round(synth.make_embeddings(bb, d, seed=30 + step))).cuda()
        rows, scores = eng.search(qd, k, algo=algo)
        want_ids, want_fin = rer.rerank_scored_device(rows, scores, q_rec[:bb].contiguous(), topk)
        for _ in range(2):
            step += 1
            _lib.check(lib.mmr_search_scatter(eng._handle, ex, _lib.ptr(qd), bb, _lib.MMR_F32, k, _lib.ALGOS[algo], step,
                                              _lib.current_stream(0)))
            p_ids, p_fin = C.c_void_p(), C.c_void_p()
            _lib.check(lib.mmr_exchange_rerank(ex, rer._tables, _lib.ptr(q_rec), bb, k, rer.alpha, rer.beta, rer.gamma,
                                               topk, step, C.byref(p_ids), C.byref(p_fin), _lib.current_stream(0)))
            keep = topk if 0 < topk < k else k
            ids = _lib.as_cuda_tensor(p_ids.value, (bb, keep), torch.int64, 0)
            fin = _lib.as_cuda_tensor(p_fin.value, (bb, keep), torch.float64, 0)
            torch.cuda.synchronize()
            assert torch.equal(ids, want_ids) and torch.equal(fin, want_fin), (k, bb, topk, algo)
    # soak: 300 back-to-back steps of one shape, no host synchronisation in between
    rows, scores = eng.search(qd, k, algo=algo)
    want_ids, want_fin = rer.rerank_scored_device(rows, scores, q_rec[:bb].contiguous(), topk)
    for _ in range(300):
        step += 1
        _lib.check(lib.mmr_search_scatter(eng._handle, ex, _lib.ptr(qd), bb, _lib.MMR_F32, k, _lib.ALGOS[algo], step,
                                          _lib.current_stream(0)))
        _lib.check(lib.mmr_exchange_rerank(ex, rer._tables, _lib.ptr(q_rec), bb, k, rer.alpha, rer.beta, rer.gamma,
                                           topk, step, C.byref(p_ids), C.byref(p_fin), _lib.current_stream(0)))
    torch.cuda.synchronize()
    assert torch.equal(_lib.as_cuda_tensor(p_ids.value, (bb, keep), torch.int64, 0), want_ids)
    assert torch.equal(_lib.as_cuda_tensor(p_fin.value, (bb, keep), torch.float64, 0), want_fin)
    code = C.c_int32(-1)
    assert lib.mmr_exchange_status(ex, C.byref(code)) == 0 and code.value == 0
    assert lib.mmr_search_scatter(eng._handle, ex, _lib.ptr(qd), 301, _lib.MMR_F32, 10, 0, step + 1, None) == _lib.MMR_EINVAL
    assert lib.mmr_exchange_create(C.byref(C.c_void_p()), 0, 1, 10, 129, 0) == _lib.MMR_EUNSUP
    # a rerank whose (b, k) differs from what was scattered: reported, not hung
    step += 1
    _lib.check(lib.mmr_search_scatter(eng._handle, ex, _lib.ptr(qd), 2, _lib.MMR_F32, 100, 0, step, _lib.current_stream(0)))
    p_ids, p_fin = C.c_void_p(), C.c_void_p()
    _lib.check(lib.mmr_exchange_rerank(ex, rer._tables, _lib.ptr(q_rec), 2, 50, 0.6, 0.25, 0.15, 0, step, C.byref(p_ids),
                                       C.byref(p_fin), _lib.current_stream(0)))
    torch.cuda.synchronize()
    assert lib.mmr_exchange_status(ex, C.byref(code)) == _lib.MMR_ECUDA and code.value == 3
    assert b"different batch size or k" in lib.mmr_last_error()
    assert lib.mmr_search_scatter(eng._handle, ex, _lib.ptr(qd), 2, _lib.MMR_F32, 100, 0, step + 1, None) == _lib.MMR_ECUDA
    _lib.check(lib.mmr_exchange_destroy(ex))
    # a wait for a step nobody scattered gives up after the timeout (and at once after an abort)
    ex2 = C.c_void_p()
    _lib.check(lib.mmr_exchange_create(C.byref(ex2), 0, 1, 8, 10, 0))
    _lib.check(lib.mmr_exchange_set_timeout(ex2, 150))
    _lib.check(lib.mmr_exchange_rerank(ex2, rer._tables, _lib.ptr(q_rec), 4, 10, 0.6, 0.25, 0.15, 0, 1, C.byref(p_ids),
                                       C.byref(p_fin), _lib.current_stream(0)))
    torch.cuda.synchronize()
    assert lib.mmr_exchange_status(ex2, C.byref(code)) == _lib.MMR_ECUDA and code.value in (1, 2)
    assert torch.all(_lib.as_cuda_tensor(p_ids.value, (4, 10), torch.int64, 0) == -1)
    _lib.check(lib.mmr_exchange_abort(ex2))
    _lib.check(lib.mmr_exchange_destroy(ex2))


def test_metrics_match_reference_golden_bit_for_bit():
    from multi_modal_retrieval_predict_project_b200.Helpers import retrieval_metrics as m
    gold = json.load(open(GOLDEN / "metrics.json"))
    for c in gold["cases"]:
        ret, rel = c["retrieved"], c["relevant"]
        for k_s, w in c["by_k"].items():
            k = int(k_s)
            t = m.per_query_metrics([ret], [rel], k)[0]
            assert (t[0], t[1], t[2], t[4]) == (w["p"], w["r"], w["ap_list"], w["ndcg"]), (ret, rel, k)
            assert m.average_precision(ret, set(rel), k) == w["ap_set"]
        assert m.average_precision(ret, rel, None) == c["ap_none"]
        assert m.mean_reciprocal_rank([ret], [rel]) == c["rr"]
    rets = [c["retrieved"] for c in gold["cases"]]; rels = [c["relevant"] for c in gold["cases"]]
    for k_s, w in gold["agg"].items():
        k = int(k_s)
        assert float(np.mean([m.precision_at_k(r, l, k=k) for r, l in zip(rets[:5], rels[:5])])) == \
            float(np.mean([om.precision_at_k(r, l, k) for r, l in zip(rets[:5], rels[:5])]))
        assert m.mean_average_precision(rets, rels, k=k) == w["mAP"]
        assert m.mean_reciprocal_rank(rets, rels) == w["MRR"]
        e = m.evaluate_retrieval(rets, rels, k)
        assert (e["P"], e["R"], e["mAP"], e["MRR"], e["nDCG"]) == (w["P"], w["R"], w["mAP"], w["MRR"], w["nDCG"])
    with pytest.raises(ZeroDivisionError):
        m.precision_at_k(["a"], ["a"], k=0)


def test_metrics_large_vs_oracle_identical():
    """cfg1-style: 1500 queries, top-10 lists, ~1800 relevant ids each -- values identical."""
    from multi_modal_retrieval_predict_project_b200.Helpers import per_query_metrics
    rng = np.random.default_rng(17)
    universe = [f"g{i}" for i in range(7500)]
    rets, rels = [], []
    for i in range(1500):
        rets.append([universe[j] for j in rng.integers(0, 7500, size=10)])
        rels.append([universe[j] for j in rng.choice(7500, size=int(rng.integers(0, 3000)), replace=False)])
    for k in (5, 10):
        got = per_query_metrics(rets, rels, k)
        want = om.per_query_table(rets, rels, k)
        assert np.array_equal(got, want)


def test_label_relevance_matches_reference_gt():
    import torch
    from multi_modal_retrieval_predict_project_b200.Evaluate import label_masks, relevance_lists
    gold = json.load(open(GOLDEN / "gt.json"))
    z = np.load(GOLDEN / "gt_inputs.npz")
    te_ids = [f"te{i}" for i in range(z["test_labels"].shape[0])]
    tr_ids = [f"tr{i}" for i in range(z["train_labels"].shape[0])]
    assert relevance_lists(z["test_labels"], te_ids, z["test_labels"], te_ids, exclude_self=True) == gold["test_relevance"]
    assert relevance_lists(z["test_labels"], te_ids, z["train_labels"], tr_ids, exclude_self=False) == gold["test_to_train"]


def test_compute_ranking_metrics_matches_reference():
    from multi_modal_retrieval_predict_project_b200.Evaluate import compute_ranking_metrics
    gold = json.load(open(GOLDEN / "gt.json"))
    z = np.load(GOLDEN / "gt_inputs.npz")
    for k_s, w in gold["ranking"].items():
        mrr, hit, rec = compute_ranking_metrics(z["queries"], z["gallery"], z["test_labels"], z["train_labels"],
                                                k=int(k_s))
        assert np.isclose(mrr, w[0], rtol=1e-9) and hit == w[1] and np.isclose(rec, w[2], rtol=1e-9)


def test_device_native_metrics_cfg5_style():
    """cfg5 at reduced scale: search results (CUDA tensor of row ids) + CSR relevance on the device ->
    metrics identical to the oracle."""
    import torch
    from multi_modal_retrieval_predict_project_b200 import B200RetrievalEngine, synth
    from multi_modal_retrieval_predict_project_b200.Helpers import metrics_from_rows
    n, d, nq, k = 20000, 128, 300, 100
    g = osr.to_bf16_round(synth.make_embeddings(n, d, seed=33))
    q = osr.to_bf16_round(synth.make_embeddings(nq, d, seed=34))
    eng = B200RetrievalEngine.from_arrays(g, dtype="bfloat16", device=0)
    rows, _ = eng.search(torch.from_numpy(q).cuda(), k)
    rng = np.random.default_rng(35)
    sizes = rng.integers(0, 201, size=nq)
    indptr = np.zeros(nq + 1, np.int64); indptr[1:] = np.cumsum(sizes)
    rel = np.concatenate([np.sort(rng.choice(n, size=s, replace=False)) for s in sizes] + [np.zeros(0, np.int64)]).astype(np.int64)
    # make sure some relevant ids are actually retrieved
    r_host = rows.cpu().numpy()
    for i in range(0, nq, 3):
        if sizes[i] >= 3:
            seg = rel[indptr[i]:indptr[i + 1]]
            seg[:3] = r_host[i, [0, 7, 50]]
            rel[indptr[i]:indptr[i + 1]] = np.unique(seg).tolist() + [n + j for j in range(len(seg) - len(np.unique(seg)))]
            rel[indptr[i]:indptr[i + 1]] = np.sort(rel[indptr[i]:indptr[i + 1]])
    for kk in (10, 100):
        got = metrics_from_rows(rows, torch.from_numpy(indptr).cuda(), torch.from_numpy(rel).cuda(), kk).cpu().numpy()
        rets = [[int(x) for x in r_host[i]] for i in range(nq)]
        rels = [rel[indptr[i]:indptr[i + 1]].tolist() for i in range(nq)]
        assert np.array_equal(got, om.per_query_table(rets, rels, kk))
        assert got[:, 0].sum() > 0


def test_diversity_matches_reference_golden():
    from multi_modal_retrieval_predict_project_b200.Evaluate import (compute_embedding_diversity,
                                                                  compute_label_diversity_from_labels)
    gold = json.load(open(GOLDEN / "diversity.json"))
    for c in gold:
        e = np.array(c["emb"], dtype=np.float32).reshape(len(c["emb"]), -1) if c["emb"] else np.zeros((0, 4), np.float32)
        assert np.isclose(compute_embedding_diversity(e), c["emb_div"], rtol=0, atol=2e-6)
        assert compute_label_diversity_from_labels(c["labels"]) == c["label_div"]


def test_compute_ranking_metrics_rejects_k_beyond_the_kernel_limit():
    """k > MMR_MAX_K on a gallery larger than that used to compute Recall@k over the top 1024 only (ADVICE)."""
    from multi_modal_retrieval_predict_project_b200.Evaluate import compute_ranking_metrics
    rng = np.random.default_rng(0)
    g = rng.standard_normal((1500, 16)).astype(np.float32)
    lab = (rng.random((1500, 8)) < 0.2).astype(np.int64)
    with pytest.raises(NotImplementedError, match="exceeds"):
        compute_ranking_metrics(g[:4], g, lab[:4], lab, k=1200)
    compute_ranking_metrics(g[:4], g[:900], lab[:4], lab[:900], k=1200)      # k > N <= limit is fine


def test_compute_ranking_metrics_large_gallery_vs_oracle():
    """Full-ranking metrics beyond K = 1024: the rank of the first relevant item comes from the
    two-sweep device reduction, not from a top-K list."""
    from multi_modal_retrieval_predict_project_b200 import synth
    from multi_modal_retrieval_predict_project_b200.Evaluate import compute_ranking_metrics
    n, nq, d, L = 3000, 40, 96, 30
    g = synth.make_embeddings(n, d, seed=61, clustered=True)
    q = synth.make_embeddings(nq, d, seed=62, clustered=True)
    gl = synth.make_labels(n, L, p=0.004, seed=63)       # sparse labels: first relevant rank is often deep
    ql = synth.make_labels(nq, L, p=0.05, seed=64)
    ql[3] = 0                                            # a query with no labels: rank None, recall 0
    for k in (1, 10):
        got = compute_ranking_metrics(q, g, ql, gl, k=k)
        want = ogt.compute_ranking_metrics(q, g, ql, gl, k=k)
        assert np.isclose(got[0], want[0], rtol=1e-9) and got[1] == want[1] and np.isclose(got[2], want[2], rtol=1e-9)


def test_label_ranking_eval_vs_oracle():
    """Evaluate.evaluate_label_ranking (mmr_label_ranking_eval) == the restated evaluate_label_attention
    metrics (reference Trainner/train_label_attention.py:106-125): AP over the full ranking and
    mean relevance of the top k, per record and averaged."""
    from multi_modal_retrieval_predict_project_b200.Evaluate import evaluate_label_ranking
    from oracle import gt as ogt
    rng = np.random.default_rng(21)
    for n, d, L in ((300, 48, 14), (1000, 300, 43)):
        # integer-valued embeddings with amp^2 * d < 2^24: every fp32 dot product is exact in any summation
        # order, so the device and numpy/BLAS see bit-identical similarities (no near-tie reorderings)
        amp = int(np.sqrt(2.0 ** 24 / d)) // 2
        embs = rng.integers(-amp, amp + 1, size=(n, d)).astype(np.float32)
        vals = (rng.random((n, L)) < 0.08).astype(int)
        vals[3] = 0                                              # a record without labels: AP 0, recall 0
        got, table = evaluate_label_ranking(embs, vals, topk=(1, 5, 10), device=0, return_table=True)
        want, wtable = ogt.label_ranking_eval(embs, vals, topk=(1, 5, 10))
        assert np.allclose(table, wtable, rtol=0, atol=1e-9), np.abs(table - wtable).max()
        assert set(got) == set(want)
        for key in want:
            assert abs(got[key] - want[key]) < 1e-9, key
        assert table[3, 0] == 0.0


def test_cfg1_openi_scale_full_pipeline_vs_oracle(tmp_path):
    """BASELINE configs[0], the whole path at the reference's own scale and through its own entry points: a
    7.5k x 1024 fp32 gallery loaded from .npy / .json, 1.5k queries, make_retrieval_engine(...).retrieve(Q, K=10,
    reranker=, query_id=) (exact search + label / KG rerank, reranker.py weights 0.6 / 0.25 / 0.15), relevance by
    label overlap (contructGT.py:68-81) and P@k / Recall@k / mAP / MRR / nDCG at k = 5 and 10 (the evaluation loop
    of Evaluate/retrieval_eval_variants.py:96-122) -- against the oracle run on the same files: exact search
    (sklearn form) -> OracleReranker.rerank per query -> the oracle metric functions.  The ranked id lists must
    agree except for swaps between near-equal combined scores, and the metrics to 1e-9 wherever the lists are
    identical (bit-identical per query then)."""
    from multi_modal_retrieval_predict_project_b200 import Reranker, make_retrieval_engine, synth
    from multi_modal_retrieval_predict_project_b200.Evaluate import relevance_lists
    from multi_modal_retrieval_predict_project_b200.Helpers import per_query_metrics
    n, nq, d, k = 7500, 1500, 1024, 10
    g = synth.make_embeddings(n, d, seed=synth.SEED, clustered=True)
    q = synth.make_embeddings(nq, d, seed=synth.SEED + 1, clustered=True)
    ids, qids = synth.make_ids(n), synth.make_ids(nq, "t")
    labels = synth.make_labels(n + nq)
    names = synth.label_names()
    csv = synth.write_labels_csv(str(tmp_path / "labels.csv"), ids + qids, labels, names)
    kg_dir = synth.write_kg(str(tmp_path / "kg"), ids + qids, names)
    fp, ip = synth.write_gallery(str(tmp_path), "train", g, ids)
    # ---- device path, reference entry points ----
    eng = make_retrieval_engine(fp, ip, method="dls", link_threshold=0.5, max_links=10)     # fp32, exact
    rer = Reranker(kg_dir, csv, device=0)
    got_ids, got_scores = eng.retrieve(q, K=k, reranker=rer, query_id=qids, rerank_topk=k)
    rel = relevance_lists(labels[n:], qids, labels[:n], ids, exclude_self=False, device=0)
    # ---- oracle path ----
    want_rows, _ = osr.exact_topk(q, g, k)
    ora = orr.OracleReranker(kg_dir, csv)
    want_rel = ogt.relevance_lists(labels[n:], qids, labels[:n], ids, exclude_self=False)
    assert rel == want_rel
    want_ids = []
    same = np.zeros(nq, dtype=bool)
    for i in range(nq):
        cand = [ids[int(j)] for j in want_rows[i]]
        res = ora.rerank(qids[i], cand, candidate_embs=g[want_rows[i]], query_emb=q[i], topk=k)
        want_ids.append([t[0] for t in res])
        same[i] = got_ids[i] == want_ids[i]
        if not same[i]:       # only swaps between near-equal combined scores
            assert sorted(got_ids[i]) == sorted(want_ids[i]), i
            fin = {t[0]: t[1] for t in res}
            for a_, b_ in zip(got_ids[i], want_ids[i]):
                assert abs(fin[a_] - fin[b_]) < 1e-5, (i, a_, b_)
        assert np.allclose(got_scores[i], sorted((t[1] for t in res), reverse=True), rtol=1e-5, atol=5e-6), i
    assert same.mean() > 0.99
    rels = [rel[qid] for qid in qids]
    for kk in (5, 10):
        got_t = per_query_metrics(got_ids, rels, kk, device=0)
        want_t = om.per_query_table(want_ids, rels, kk)
        assert np.array_equal(got_t[same], want_t[same])
        assert np.allclose(got_t.mean(axis=0), want_t.mean(axis=0), rtol=0, atol=5e-3)     # near-tie swaps in < 1 % of the queries
    eng.close(); rer.close()


@pytest.mark.parametrize("emb_feature", ["search_score", "recompute"])
def test_batched_retrieve_unfused_paths_with_padding(emb_feature):
    """K beyond the fused tail (K = 200 > 128) on a gallery with fewer rows than K: the batched retrieve path
    falls back to the unfused kernels; the -1 padding of the search result stays out of the min-max scaling
    (valid-candidate counts are passed) and comes out as id -1.  Both embedding-feature modes, against the oracle
    restatement on the valid candidates."""
    import torch
    from multi_modal_retrieval_predict_project_b200 import B200RetrievalEngine, Reranker, synth
    from multi_modal_retrieval_predict_project_b200.sharded import ShardedSearcher
    n, d, b, k = 150, 64, 9, 200
    g = osr.to_bf16_round(synth.make_embeddings(n, d, seed=45, clustered=True))
    q = osr.to_bf16_round(synth.make_embeddings(b, d, seed=46, clustered=True))
    masks, kg = _tables(n, b, 48, 47)
    eng = B200RetrievalEngine.from_arrays(g, dtype="bfloat16", device=0)
    rer = Reranker.from_tables(masks, kg, device=0)
    rer.emb_feature = emb_feature
    qd = torch.from_numpy(q).cuda()
    q_rec = torch.arange(n, n + b, device="cuda")
    ids, fin = ShardedSearcher(eng).retrieve_reranked(rer, qd, k, q_rec, topk=0)
    rows, _ = eng.search(qd, k)
    torch.cuda.synchronize()
    ids_h, fin_h, rows_h = ids.cpu().numpy(), fin.cpu().numpy(), rows.cpu().numpy()
    assert ids_h.shape == (b, k) and np.all(ids_h[:, n:] == -1) and np.all(fin_h[:, n:] == 0.0)
    for i in range(b):
        cand = rows_h[i, :n]
        want = orr.rerank_from_arrays(q[i], g[cand], masks[n + i], masks[cand], kg[n + i], kg[cand], topk=n)
        ok, why = orr.reranked_lists_match(ids_h[i, :n], fin_h[i, :n], cand, want)
        assert ok, (emb_feature, i, why)
