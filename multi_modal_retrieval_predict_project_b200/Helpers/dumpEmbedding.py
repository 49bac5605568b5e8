"""Drop-in for the reference's ``Helpers/dumpEmbedding.py``: merging the
train and validation embedding dumps into the gallery the retrieval engine serves.

``createDumpEmbedding`` keeps the reference's signature and writes the same two files
(``trainval_joint_embeddings.npy``, ``trainval_ids.json``; reference ``dumpEmbedding.py:28-39`` --
with the ids written next to the embeddings, where the reference's own readers look for them).
``merged_engine`` skips the round trip through the merged ``.npy``: the split files are read
(memory-mapped), concatenated once and ingested straight into an HBM-resident index.
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Sequence

import numpy as np


def _load_split(embeddings_dir: Path, split: str):
    emb = np.load(embeddings_dir / f"{split}_joint_embeddings.npy", mmap_mode="r")
    with open(embeddings_dir / f"{split}_ids.json") as f:
        ids = json.load(f)
    if emb.shape[0] != len(ids):
        raise RuntimeError(f"{split}: {emb.shape[0]} embeddings but {len(ids)} ids")
    return emb, ids


def createDumpEmbedding(base_dir, embeddings_dir):
    """Creates a merged dump of train and validation embeddings and IDs (same outputs as the
    reference).  ``base_dir`` is accepted for signature compatibility; ``embeddings_dir`` defaults to
    ``<base_dir>/embeddings``."""
    if not embeddings_dir:
        if not base_dir:
            raise ValueError("provide embeddings_dir (or base_dir containing an embeddings/ directory)")
        embeddings_dir = Path(base_dir) / "embeddings"
    embeddings_dir = Path(embeddings_dir)
    train_emb, train_ids = _load_split(embeddings_dir, "train")
    val_emb, val_ids = _load_split(embeddings_dir, "val")
    merged_emb = np.concatenate([train_emb, val_emb], axis=0)
    np.save(embeddings_dir / "trainval_joint_embeddings.npy", merged_emb)
    with open(embeddings_dir / "trainval_ids.json", "w") as fout:
        json.dump(list(train_ids) + list(val_ids), fout)
    print(f"Saved merged embeddings to: {embeddings_dir / 'trainval_joint_embeddings.npy'}")
    print(f"Saved merged IDs to:        {embeddings_dir / 'trainval_ids.json'}")


def merged_engine(embeddings_dir, splits: Sequence[str] = ("train", "val"), **engine_kwargs):
    """HBM-resident engine over the concatenation of the given splits (rows in split order, exactly
    the gallery ``createDumpEmbedding`` would write) without materialising the merged files."""
    from ..Retrieval import B200RetrievalEngine
    embeddings_dir = Path(embeddings_dir)
    embs, ids = [], []
    for split in splits:
        e, i = _load_split(embeddings_dir, split)
        embs.append(np.asarray(e, dtype=np.float32))
        ids.extend(i)
    return B200RetrievalEngine.from_arrays(np.concatenate(embs, axis=0), ids=ids, **engine_kwargs)
