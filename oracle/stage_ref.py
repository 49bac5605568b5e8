"""Stage the reference's own hot-path files under ``oracle/_ref/`` (TEST / BASELINE INFRASTRUCTURE).

    python -m oracle.stage_ref            # run by __graft_entry__.build() in the authoring container

``/root/reference`` is mounted read-only in the authoring container and does not exist on the GPU box,
but the snapshot that travels there carries everything under the repo that is not gpurun-ignored --
including git-ignored build outputs such as ``libmmr_b200.so`` and ``oracle/_ref/``.  This recipe copies
the handful of files the retrieval path consists of (verbatim, no edits) so that ``bench.py --impl
reference`` and the ``cpu_baseline`` leg time the REFERENCE'S OWN code on the box's host cores
(``cpu_baseline.kind == "reference"``) instead of the oracle port:

    src/Retrieval/{__init__,retrieval,reranker}.py   make_retrieval_engine / DLSRetrievalEngine / Reranker
    src/KnowledgeGraph/label_attention.py            imported by reranker.py:7
    src/Helpers/{config,retrieval_metrics}.py        imported by reranker.py:8 / the metric functions
    configs/config.yaml                              read by Reranker.__init__ (reranker.py:16,61)

``oracle/_ref/`` is listed in ``.gitignore`` (never committed: reference sources stay out of the
history) and NOT in ``.gpurunignore``.  The files are loaded through ``oracle.ref_loader`` with the same
two namespace stubs as the live reference.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = [
    "src/Retrieval/__init__.py",
    "src/Retrieval/retrieval.py",
    "src/Retrieval/reranker.py",
    "src/KnowledgeGraph/label_attention.py",
    "src/Helpers/config.py",
    "src/Helpers/retrieval_metrics.py",
    "configs/config.yaml",
]


ARCHIVE = os.path.join(DEST, "ref_hotpath.zip")


def stage(reference_root: str = "/root/reference", dest: str = DEST, quiet: bool = False) -> bool:
    """Pack the files; returns False when neither the reference nor an earlier archive is present."""
    import zipfile
    archive = os.path.join(dest, "ref_hotpath.zip")
    if not os.path.isfile(os.path.join(reference_root, FILES[1])):
        if not quiet:
            print(f"[stage_ref] {reference_root} not present: keeping whatever is under {dest}")
        return os.path.isfile(archive)
    os.makedirs(dest, exist_ok=True)
    manifest = {}
    tmp = archive + ".tmp"
    with zipfile.ZipFile(tmp, "w", zipfile.ZIP_DEFLATED) as z:
        for rel in FILES:
            with open(os.path.join(reference_root, rel), "rb") as f:
                data = f.read()
            manifest[rel] = hashlib.sha256(data).hexdigest()
            z.writestr(zipfile.ZipInfo(rel, date_time=(2020, 1, 1, 0, 0, 0)), data)   # reproducible archive
    os.replace(tmp, archive)
    with open(os.path.join(dest, "MANIFEST.json"), "w") as f:
        json.dump({"source": reference_root, "sha256": manifest}, f, indent=1)
    if not quiet:
        print(f"[stage_ref] packed {len(FILES)} reference files into {archive}")
    return True


_unpacked = None


def staged_root(dest: str = DEST):
    """Unpack the archive into a per-process temporary directory (once) and return it, or None when
    there is no archive.  Every file is checked against the manifest's sha256."""
    global _unpacked
    import atexit
    import tempfile
    import zipfile
    archive = os.path.join(dest, "ref_hotpath.zip")
    if _unpacked is not None:
        return _unpacked
    if not os.path.isfile(archive):
        return None
    with open(os.path.join(dest, "MANIFEST.json")) as f:
        want = json.load(f)["sha256"]
    root = tempfile.mkdtemp(prefix="mmr_ref_")
    atexit.register(shutil.rmtree, root, ignore_errors=True)
    with zipfile.ZipFile(archive) as z:
        for rel in FILES:
            data = z.read(rel)
            if hashlib.sha256(data).hexdigest() != want[rel]:
                raise RuntimeError(f"{archive}: {rel} does not match MANIFEST.json")
            dst = os.path.join(root, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            with open(dst, "wb") as f:
                f.write(data)
    _unpacked = root
    return root


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
