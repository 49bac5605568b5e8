"""Build libmmr_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m multi_modal_retrieval_predict_project_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.  nvcc cross-compiles
without a GPU.  No fast-math: the rerank/metrics kernels rely on IEEE division and unfused fp64.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libmmr_b200.so")
SOURCES = ["api.cu", "ingest.cu", "scan_topk.cu", "select.cu", "gemm_topk.cu", "rerank.cu", "metrics.cu", "eval.cu",
           "exchange.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libmmr_b200.so cannot be built (there is no CPU fallback)")


def _deps_mtime() -> float:
    paths = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    paths.append(os.path.join(HERE, "..", "include", "mmr_b200.h"))
    paths.append(os.path.abspath(__file__))
    return max(os.path.getmtime(p) for p in paths)


def needs_build() -> bool:
    return not os.path.exists(LIB) or os.path.getmtime(LIB) < _deps_mtime()


def build(force: bool = False, verbose: bool = False, defines=(), out: str = None) -> str:
    """Default: libmmr_b200.so.  ``defines`` / ``out`` build a VARIANT library next to it for A/B runs and
    diagnostics (e.g. ``defines=("MMR_DIAG",), out="libmmr_b200_diag.so"``; load it with MMR_B200_LIB=<path>):
    its objects go to their own directory and it is always rebuilt when sources changed."""
    global LIB, OBJ
    if out is not None or defines:
        tag = "_".join(d.replace("=", "") for d in defines) or "variant"
        lib = os.path.join(HERE, out or f"libmmr_b200_{tag}.so")
        saved = (LIB, OBJ)
        LIB, OBJ = lib, os.path.join(CSRC, "_obj_" + tag)
        extra = os.environ.get("NVCC_EXTRA", "")
        os.environ["NVCC_EXTRA"] = (extra + " " + " ".join("-D" + d for d in defines)).strip()
        try:
            return build(force=force, verbose=verbose)
        finally:
            LIB, OBJ = saved
            os.environ["NVCC_EXTRA"] = extra
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    headers_mtime = max(
        os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))
    )
    headers_mtime = max(headers_mtime, os.path.getmtime(os.path.join(HERE, "..", "include", "mmr_b200.h")),
                        os.path.getmtime(os.path.abspath(__file__)))

    def compile_one(src: str) -> str:
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        sp = os.path.join(CSRC, src)
        if (not force and os.path.exists(obj) and os.path.getmtime(obj) >= os.path.getmtime(sp)
                and os.path.getmtime(obj) >= headers_mtime):
            return obj
        cmd = [nvcc, *NVCC_FLAGS, *os.environ.get("NVCC_EXTRA", "").split(), "-c", sp, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with cf.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    tmp = LIB + ".tmp"
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp, *objs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    # python -m ...build [--force] [-v] [-DNAME[=VALUE] ...] [--out libname.so]
    defs = tuple(a[2:] for a in sys.argv[1:] if a.startswith("-D"))
    out_name = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs, out=out_name))
