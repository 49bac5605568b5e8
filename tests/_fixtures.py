"""Shared builders for the test artifacts (same seeds/parameters as tests/golden/make_golden.py)."""
from __future__ import annotations

import json
import os
from pathlib import Path

import numpy as np

from multi_modal_retrieval_predict_project_b200 import synth

GOLDEN = Path(__file__).resolve().parent / "golden"


def load_search_case(name):
    return np.load(GOLDEN / f"search_{name}.npz")


def build_rerank_artifacts(tmp):
    """Recreate the labels CSV / KG dir / gallery files of make_golden.rerank_case in ``tmp``."""
    n, nq, d = 300, 12, 64
    g = synth.make_embeddings(n, d, seed=synth.SEED + 5, clustered=True)
    ids = synth.make_ids(n)
    qids = synth.make_ids(nq, prefix="t")
    rng = np.random.default_rng(synth.SEED + 6)
    qs = (g[rng.integers(0, n, size=nq)] + 0.4 * rng.standard_normal((nq, d), dtype=np.float32)).astype(np.float32)
    all_ids = ids + qids
    labels = synth.make_labels(n + nq, p=0.12, seed=synth.SEED + 7)
    labels[5] = 0
    labels[n + 2] = 0
    names = synth.label_names()
    csv = synth.write_labels_csv(os.path.join(tmp, "rr", "labels.csv"), all_ids, labels, names)
    kg_dir = synth.write_kg(os.path.join(tmp, "rr", "kg"), all_ids, names, d_kg=48, seed=synth.SEED + 8,
                            skip_every=7)
    fp, ip = synth.write_gallery(os.path.join(tmp, "rr"), "train", g, ids)
    stored = np.load(GOLDEN / "rerank_inputs.npz")
    assert np.array_equal(stored["gallery"], g) and np.array_equal(stored["queries"], qs)
    assert np.array_equal(stored["labels"], labels)
    golden = json.load(open(GOLDEN / "rerank.json"))
    return dict(g=g, qs=qs, ids=ids, qids=qids, labels=labels, names=names, csv=csv, kg_dir=kg_dir,
                features_path=fp, ids_path=ip, golden=golden)


def assert_rerank_close(got, want, rtol=1e-5, atol=2e-6):
    """Compare two rerank outputs [(id, final, emb_n, lab_n, kg_n)], tolerant to near-tie swaps."""
    assert len(got) == len(want), (len(got), len(want))
    gf = np.array([t[1] for t in got]); wf = np.array([t[1] for t in want])
    assert np.allclose(gf, wf, rtol=rtol, atol=atol), (gf, wf)
    want_by_id = {}
    for t in want:
        want_by_id.setdefault(t[0], []).append(t)
    for i, (g_, w_) in enumerate(zip(got, want)):
        if g_[0] == w_[0]:
            assert np.allclose(g_[1:], w_[1:], rtol=rtol, atol=atol), (i, g_, w_)
        else:
            # swap allowed only between (near-)equal finals
            assert abs(g_[1] - w_[1]) <= atol + rtol * abs(w_[1]), (i, g_, w_)
