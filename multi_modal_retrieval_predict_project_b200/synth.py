"""Seeded synthetic artifacts in the reference's on-disk formats.

The reference's data artifacts (``embeddings/*.npy|json``, ``outputs/openi_labels_final.csv``,
``knowledge_graph/{node2id.json,node_embeddings_best.npy}``, ``ground_truths/*.json``) are
git-ignored and absent (SURVEY.md section 0 finding 4), so tests, ``smoke()`` and
``bench.py`` use these generators.  Formats follow the producers:
``Trainner/train.py:731-733`` (fp32 ``(N,D)`` + JSON list of string ids),
``Retrieval/reranker.py:47-48`` (CSV with an ``id`` column, one 0/1 column per label),
``Retrieval/reranker.py:101-120`` (KG dir), ``Helpers/contructGT.py:95-99`` (relevance).
"""
from __future__ import annotations

import json
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

SEED = 2709  # reference configs/config.yaml:6
N_LABELS = 43  # 19 disease + 1 normal + 19 finding + 4 symptom groups (LabelData/labeledData.py)
KG_DIM = 300  # reference configs/config.yaml:23


def make_embeddings(n: int, d: int, seed: int = SEED, clustered: bool = False,
                    n_centroids: int = 64, noise: float = 0.6) -> np.ndarray:
    """iid N(0,1) fp32 rows, or 64 gaussian centroids + noise (non-degenerate rerank)."""
    rng = np.random.default_rng(seed)
    if not clustered:
        return rng.standard_normal((n, d), dtype=np.float32)
    cent = rng.standard_normal((n_centroids, d), dtype=np.float32)
    which = rng.integers(0, n_centroids, size=n)
    return (cent[which] + noise * rng.standard_normal((n, d), dtype=np.float32)).astype(np.float32)


def make_ids(n: int, prefix: str = "g") -> List[str]:
    return [f"{prefix}{i}" for i in range(n)]


def make_labels(n: int, n_labels: int = N_LABELS, p: float = 0.08, seed: int = SEED + 1) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return (rng.random((n, n_labels)) < p).astype(np.int64)


def label_names(n_labels: int = N_LABELS) -> List[str]:
    return [f"Label {i:02d}" for i in range(n_labels)]


def write_gallery(dirpath: str, stem: str, embs: np.ndarray, ids: Sequence[str]) -> Tuple[str, str]:
    os.makedirs(dirpath, exist_ok=True)
    fp = os.path.join(dirpath, f"{stem}_joint_embeddings.npy")
    ip = os.path.join(dirpath, f"{stem}_ids.json")
    np.save(fp, np.asarray(embs, dtype=np.float32))
    with open(ip, "w") as f:
        json.dump(list(ids), f)
    return fp, ip


def write_labels_csv(path: str, ids: Sequence[str], vals: np.ndarray,
                     names: Optional[Sequence[str]] = None, with_text_column: bool = True) -> str:
    """CSV with an ``id`` column, 0/1 label columns and (optionally) a free-text column
    that ``get_record_label_set`` must skip (reference ``reranker.py:173-178``)."""
    names = list(names) if names is not None else label_names(vals.shape[1])
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "w") as f:
        cols = ["id"] + names + (["report_text"] if with_text_column else [])
        f.write(",".join(cols) + "\n")
        for i, rid in enumerate(ids):
            row = [str(rid)] + [str(int(v)) for v in vals[i]]
            if with_text_column:
                row.append(f"findings for {rid}")
            f.write(",".join(row) + "\n")
    return path


def write_kg(dirpath: str, record_ids: Sequence[str], names: Sequence[str], d_kg: int = KG_DIM,
             seed: int = SEED + 2, skip_every: int = 0) -> str:
    """``node2id.json`` with ``report:<id>`` and ``label:<name>`` keys plus
    ``node_embeddings_best.npy`` (fp32, unnormalised -- the reranker normalises).
    ``skip_every`` > 0 leaves every n-th record without a ``report:`` node so the
    label-pooling fallback (reference ``reranker.py:196-220``) is exercised."""
    os.makedirs(dirpath, exist_ok=True)
    node2id: Dict[str, int] = {}
    for i, rid in enumerate(record_ids):
        if skip_every and i % skip_every == skip_every - 1:
            continue
        node2id[f"report:{rid}"] = len(node2id)
    for nm in names:
        node2id[f"label:{nm}"] = len(node2id)
    rng = np.random.default_rng(seed)
    emb = rng.standard_normal((len(node2id), d_kg), dtype=np.float32)
    np.save(os.path.join(dirpath, "node_embeddings_best.npy"), emb)
    with open(os.path.join(dirpath, "node2id.json"), "w", encoding="utf8") as f:
        json.dump(node2id, f)
    return dirpath


def random_relevance(query_ids: Sequence[str], gallery_ids: Sequence[str], max_rel: int = 200,
                     seed: int = SEED + 1) -> Dict[str, List[str]]:
    """cfg5-style relevance: per query a unique id set of size U[1, max_rel]."""
    rng = np.random.default_rng(seed)
    n = len(gallery_ids)
    out = {}
    for q in query_ids:
        m = int(rng.integers(1, min(max_rel, n) + 1))
        out[q] = [gallery_ids[j] for j in sorted(rng.choice(n, size=m, replace=False).tolist())]
    return out
