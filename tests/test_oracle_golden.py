"""CPU: the oracle restatement against the golden vectors produced by the real reference."""
import json

import numpy as np
import pytest

from oracle import gt as ogt
from oracle import metrics as om
from oracle import rerank as orr
from oracle import search as osr
from tests._fixtures import GOLDEN, assert_rerank_close, build_rerank_artifacts, load_search_case


@pytest.mark.parametrize("name", ["gauss", "clustered"])
def test_cosine_and_ranking_match_reference(name):
    c = load_search_case(name)
    sim = osr.cosine_similarity(c["queries"], c["gallery"])
    # same arithmetic (normalise rows, one sgemm): allow BLAS blocking differences only
    assert np.allclose(sim, c["sim"], rtol=0, atol=2e-7)
    rows, scores = osr.exact_topk(c["queries"], c["gallery"], 50)
    for i in range(rows.shape[0]):
        ok, why = osr.topk_matches(rows[i], scores[i], c["order_top"][i], c["sim"][i][c["order_top"][i]],
                                   rtol=1e-6, atol=1e-7)
        assert ok, (i, why)
    # zero rows score exactly 0 (sklearn zero-norm rule), duplicates tie exactly
    assert np.all(sim[:, 7] == 0) and np.all(sim[1] == 0)
    assert np.array_equal(sim[:, 11], sim[:, 3]) and np.array_equal(sim[:, 12], sim[:, 3])


@pytest.mark.parametrize("name", ["gauss", "clustered"])
def test_link_graph_and_dls_walk_match_reference(name):
    c = load_search_case(name)
    g = c["gallery"]
    graph = osr.build_link_graph(g, float(c["link_threshold"]), int(c["max_links"]))
    sim = osr.cosine_similarity(g)
    np.fill_diagonal(sim, -1)
    for i, row in enumerate(graph):
        want = c["graph"][i][: c["graph_len"][i]].tolist()
        if row != want:  # only exact-score ties may be ordered differently
            assert len(row) == len(want)
            assert np.allclose(sim[i, row], sim[i, want], rtol=0, atol=1e-7), (i, row, want)
    ref_graph = [c["graph"][i][: c["graph_len"][i]].tolist() for i in range(g.shape[0])]
    for qi in range(c["dls_full_ids"].shape[0]):
        ids, sc = osr.dls_retrieve(g, ref_graph, c["queries"][qi], K=5, seed_size=g.shape[0], seed=1)
        want = [x for x in c["dls_full_ids"][qi].tolist() if x >= 0]
        assert ids == want
        assert np.allclose(sc, c["dls_full_scores"][qi][: len(sc)], rtol=1e-6)
        ids, sc = osr.dls_retrieve(g, ref_graph, c["queries"][qi], K=5, seed=2709)
        want = [x for x in c["dls_default_ids"][qi].tolist() if x >= 0]
        assert ids == want
        assert np.allclose(sc, c["dls_default_scores"][qi][: len(sc)], rtol=1e-6)


def test_dls_full_seed_is_exact_shifted_by_one():
    """SURVEY section 0 finding 2: retrieve(seed_size=N) == exact ranks 1..K (rank 0 dropped)."""
    c = load_search_case("gauss")
    for qi in (0, 2, 3):
        want = c["order_top"][qi][1:6].tolist()
        got = [x for x in c["dls_full_ids"][qi].tolist()]
        assert got == want


def test_rerank_matches_reference(tmp_path):
    a = build_rerank_artifacts(str(tmp_path))
    gold = a["golden"]
    # tolerance, not bit equality: for records without a ``report:`` node the reference mean-pools
    # label vectors in ``set`` iteration order (reranker.py:197-220), which depends on
    # PYTHONHASHSEED, so the reference itself is only reproducible to ~1 fp32 ulp there.
    tol = dict(rtol=1e-6, atol=1e-7)
    for name, v in gold["variants"].items():
        al, be, ga = v["weights"]
        rer = orr.OracleReranker(a["kg_dir"], a["csv"], al, be, ga)
        for qi, res in enumerate(v["results"]):
            cand = res["cand"]
            cand_ids = [a["ids"][j] for j in cand]
            embs = a["g"][cand]
            r1 = rer.rerank(a["qids"][qi], cand_ids, candidate_embs=embs, query_emb=a["qs"][qi], topk=10)
            assert_rerank_close(r1, [tuple(t) for t in res["qemb_top10"]], **tol)
            lookup = {cid: a["g"][j] for cid, j in zip(cand_ids, cand)}
            lookup[a["qids"][qi]] = a["qs"][qi]
            r2 = rer.rerank(a["qids"][qi], cand_ids, candidate_embs=embs, candidate_emb_lookup=lookup)
            assert_rerank_close(r2, [tuple(t) for t in res["lookup_all"]], **tol)
            r3 = rer.rerank(cand_ids[3], cand_ids, candidate_embs=embs, topk=5)
            assert_rerank_close(r3, [tuple(t) for t in res["gallery_qid_top5"]], **tol)
            r4 = rer.rerank("not-a-record", cand_ids, candidate_embs=embs, query_emb=a["qs"][qi], topk=5)
            assert_rerank_close(r4, [tuple(t) for t in res["unknown_qid_top5"]], **tol)


def test_metrics_match_reference_bit_for_bit():
    gold = json.load(open(GOLDEN / "metrics.json"))
    for c in gold["cases"]:
        ret, rel = c["retrieved"], c["relevant"]
        for k_s, w in c["by_k"].items():
            k = int(k_s)
            assert om.precision_at_k(ret, rel, k) == w["p"]
            assert om.recall_at_k(ret, rel, k) == w["r"]
            assert om.average_precision(ret, rel, k) == w["ap_list"]
            assert om.average_precision(ret, set(rel), k) == w["ap_set"]
            assert float(om.ndcg_at_k(ret, rel, k)) == w["ndcg"]
        assert om.average_precision(ret, rel, None) == c["ap_none"]
        assert om.reciprocal_rank(ret, rel) == c["rr"]
    rets = [c["retrieved"] for c in gold["cases"]]; rels = [c["relevant"] for c in gold["cases"]]
    for k_s, w in gold["agg"].items():
        t = om.per_query_table(rets, rels, int(k_s))
        assert float(np.mean(t[:, 0])) == w["P"] and float(np.mean(t[:, 1])) == w["R"]
        assert om.mean_average_precision(rets, rels, int(k_s)) == w["mAP"]
        assert om.mean_reciprocal_rank(rets, rels) == w["MRR"]
        assert float(np.mean(t[:, 4])) == w["nDCG"]


def test_gt_and_ranking_metrics_match_reference():
    gold = json.load(open(GOLDEN / "gt.json"))
    z = np.load(GOLDEN / "gt_inputs.npz")
    te_ids = [f"te{i}" for i in range(z["test_labels"].shape[0])]
    tr_ids = [f"tr{i}" for i in range(z["train_labels"].shape[0])]
    assert ogt.relevance_lists(z["test_labels"], te_ids, z["test_labels"], te_ids, True) == gold["test_relevance"]
    assert ogt.relevance_lists(z["test_labels"], te_ids, z["train_labels"], tr_ids, False) == gold["test_to_train"]
    for k_s, w in gold["ranking"].items():
        mrr, hit, rec = ogt.compute_ranking_metrics(z["queries"], z["gallery"], z["test_labels"],
                                                    z["train_labels"], k=int(k_s))
        assert np.isclose(mrr, w[0], rtol=1e-12) and hit == w[1] and np.isclose(rec, w[2], rtol=1e-12)


def test_bf16_rounding_helpers():
    import torch
    x = np.random.default_rng(0).standard_normal(10000).astype(np.float32) * 37.0
    want = torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy()
    assert np.array_equal(osr.to_bf16_round(x), want)
    bits = osr.to_bf16_bits(x)
    assert np.array_equal(bits, torch.from_numpy(x).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16))


@pytest.mark.reference
def test_oracle_against_live_reference(tmp_path):
    """When /root/reference is mounted, re-run the real code and compare again (fresh seed)."""
    from oracle.ref_loader import load_reference
    from sklearn.metrics.pairwise import cosine_similarity
    ref = load_reference()
    rng = np.random.default_rng(123)
    g = rng.standard_normal((257, 40)).astype(np.float32)
    q = rng.standard_normal((9, 40)).astype(np.float32)
    assert np.allclose(osr.cosine_similarity(q, g), cosine_similarity(q, g), rtol=0, atol=2e-7)
    ret = [["a", "b", "c", "a"], [], ["x"]]
    rel = [["b", "b", "z"], ["q"], []]
    for r, l in zip(ret, rel):
        for k in (1, 2, 5):
            assert om.precision_at_k(r, l, k) == ref.metrics.precision_at_k(r, l, k)
            assert om.recall_at_k(r, l, k) == ref.metrics.recall_at_k(r, l, k)
            assert om.average_precision(r, l, k) == ref.metrics.average_precision(r, l, k)
            assert om.ndcg_at_k(r, l, k) == ref.metrics.ndcg_at_k(r, l, k)
    assert om.mean_reciprocal_rank(ret, rel) == ref.metrics.mean_reciprocal_rank(ret, rel)


def test_diversity_matches_reference():
    gold = json.load(open(GOLDEN / "diversity.json"))
    for c in gold:
        e = np.array(c["emb"], dtype=np.float32).reshape(len(c["emb"]), -1) if c["emb"] else np.zeros((0, 4), np.float32)
        assert np.isclose(ogt.compute_embedding_diversity(e), c["emb_div"], rtol=0, atol=1e-7)
        assert ogt.compute_label_diversity_from_labels(c["labels"]) == c["label_div"]


def test_average_precision_restatement_matches_sklearn():
    """oracle.gt.average_precision_binary == sklearn.metrics.average_precision_score (the function the
    reference calls at Trainner/train_label_attention.py:122), incl. tied scores and the no-positive case;
    label_ranking_eval's table is consistent with it."""
    import warnings
    from sklearn.metrics import average_precision_score
    from oracle import gt as ogt
    rng = np.random.default_rng(11)
    for trial in range(30):
        n = int(rng.integers(2, 60))
        y = (rng.random(n) < 0.3).astype(int)
        s = rng.standard_normal(n).astype(np.float32)
        if trial % 3 == 0:
            s = np.round(s, 1)                                   # many ties
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want = float(average_precision_score(y, s)) if y.sum() > 0 else 0.0
        assert abs(ogt.average_precision_binary(y, s) - want) < 1e-12, trial
    embs = rng.standard_normal((40, 8)).astype(np.float32)
    vals = (rng.random((40, 6)) < 0.2).astype(int)
    res, table = ogt.label_ranking_eval(embs, vals, topk=(1, 5))
    assert set(res) == {"recall@1", "recall@5", "mAP"} and table.shape == (40, 3)
    assert abs(res["mAP"] - table[:, 0].mean()) < 1e-15 and np.all((table >= 0) & (table <= 1))


def test_label_ranking_eval_matches_reference_golden():
    """oracle.gt.label_ranking_eval == the reference's own evaluate_label_attention
    (Trainner/train_label_attention.py:95-131, executed by tests/golden/make_golden.py with the reference's
    LabelAttention module) on the recorded embeddings -- records with equal label sets have identical
    embeddings, so the fixture also pins the tie behaviour of the ``np.argsort(-sim_row)`` the reference uses."""
    import json
    from pathlib import Path
    from oracle import gt as ogt
    golden = Path(__file__).resolve().parent / "golden"
    z = np.load(golden / "label_ranking_inputs.npz")
    want = json.load(open(golden / "label_ranking.json"))
    got, table = ogt.label_ranking_eval(z["embs"], z["labels"], topk=(1, 5, 10))
    assert set(got) == set(want)
    for key, val in want.items():
        assert abs(got[key] - val) < 1e-12, (key, got[key], val)
    assert table.shape == (z["embs"].shape[0], 4)


def test_bruteforce_checker_is_pinned_on_the_numpy_oracle():
    """oracle/bruteforce.py (torch, chunked -- the checker of the full-size GPU parity tests and of
    bench.py's parity_check) returns what oracle.search.exact_topk returns, ties included."""
    import torch
    from oracle import search as osr
    from oracle.bruteforce import bruteforce_topk, check_topk
    rng = np.random.default_rng(5)
    g = osr.to_bf16_round(rng.standard_normal((3000, 48)).astype(np.float32))
    g[100:140] = g[7]
    g[9] = 0.0
    q = osr.to_bf16_round(rng.standard_normal((9, 48)).astype(np.float32))
    q[2] = g[7]
    q[4] = 0.0
    for k in (1, 10, 100, 5000):
        want_r, want_s = osr.exact_topk(q, g, k)
        for chunk in (512, 1 << 19):
            r, s = bruteforce_topk(torch.from_numpy(g), torch.from_numpy(q), k, chunk=chunk)
            assert r.shape == want_r.shape
            for i in range(9):
                ok, why = osr.topk_matches(r[i], s[i], want_r[i], want_s[i], rtol=1e-6, atol=1e-7)
                assert ok, (k, chunk, i, why)
            # (BLAS does not return bit-equal dots for equal rows at different positions, so the order inside
            # the 41-way tie is compared as a set by topk_matches, not position by position)
            assert set(r[2][: min(k, 41)].tolist()) <= set(np.r_[7, 100:140].tolist())
    r, s = bruteforce_topk(torch.from_numpy(g), torch.from_numpy(q), 10, row_offset=500)
    ok, detail = check_topk(r, s, torch.from_numpy(g), torch.from_numpy(q), 10, row_offset=500)
    assert ok, detail
    bad = r.copy(); bad[0, 0] = 1234 + 500
    assert not check_topk(bad, s, torch.from_numpy(g), torch.from_numpy(q), 10, row_offset=500)[0]


def test_array_rerank_checker_is_pinned_on_the_file_based_oracle(tmp_path):
    """oracle.rerank.rerank_from_arrays (the checker of bench.py's parity_check and of the fused-tail GPU tests:
    integer label masks + KG rows instead of the CSV / node2id files) returns what OracleReranker.rerank --
    itself pinned on the reference's golden vectors -- returns for the same records."""
    from oracle import rerank as orr
    from tests._fixtures import build_rerank_artifacts
    a = build_rerank_artifacts(str(tmp_path))
    ora = orr.OracleReranker(a["kg_dir"], a["csv"], 0.4, 0.2, 0.2)
    all_ids = a["ids"] + a["qids"]
    names = list(ora.labels_df.columns)
    mask_of = {rid: sum(1 << names.index(c) for c in ora.get_record_label_set(rid)) for rid in all_ids}
    kg_of = {rid: np.asarray(ora.get_record_kg_vec(rid), dtype=np.float32) for rid in all_ids}
    rng = np.random.default_rng(3)
    for qi in range(6):
        cand = rng.choice(len(a["ids"]), size=40, replace=False)
        cand_ids = [a["ids"][j] for j in cand]
        qid = a["qids"][qi]
        want = ora.rerank(qid, cand_ids, candidate_embs=a["g"][cand], query_emb=a["qs"][qi], topk=15)
        got = orr.rerank_from_arrays(a["qs"][qi], a["g"][cand], mask_of[qid], [mask_of[c] for c in cand_ids], kg_of[qid],
                                     [kg_of[c] for c in cand_ids], 0.4, 0.2, 0.2, topk=15)
        assert [cand_ids[g_[0]] for g_ in got] == [w[0] for w in want]
        assert np.allclose([g_[1:] for g_ in got], [w[1:] for w in want], rtol=0, atol=1e-6)
        ok, why = orr.reranked_lists_match([cand[g_[0]] for g_ in got], [g_[1] for g_ in got], cand, got)
        assert ok, why
    bad = list(got)
    bad[0], bad[5] = bad[5], bad[0]
    assert not orr.reranked_lists_match([cand[g_[0]] for g_ in bad], [g_[1] for g_ in bad], cand, got)[0]
