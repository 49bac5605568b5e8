"""CPU: the C-ABI library loads and exports every symbol include/mmr_b200.h declares; host-side
logic (id interning, label bit masks, shard bounds); loud failure without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pandas as pd
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from multi_modal_retrieval_predict_project_b200 import _lib
    from multi_modal_retrieval_predict_project_b200.build import build
    build()
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    from multi_modal_retrieval_predict_project_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "mmr_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:int|int64_t|const char\*)\s+(mmr_\w+)\s*\(", hdr, flags=re.M))
    assert len(declared) >= 16
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name)
    assert lib.mmr_abi_version() == int(re.search(r"#define MMR_ABI_VERSION (\d+)", hdr).group(1))


def test_argument_errors_without_touching_the_gpu(lib):
    from multi_modal_retrieval_predict_project_b200 import _lib
    h = C.c_void_p()
    assert lib.mmr_index_create(C.byref(h), None, 10, 8, 0, 0, 0, 0, 0, None) == _lib.MMR_EINVAL
    assert b"emb is NULL" in lib.mmr_last_error()
    assert lib.mmr_search(None, None, 1, 0, 5, 0, None, None, None, None) == _lib.MMR_EINVAL
    assert lib.mmr_metrics(None, None, 3, 5, None, None, None, 0, None, None, 0, None) == _lib.MMR_EINVAL
    with pytest.raises(ValueError):
        _lib.check(_lib.MMR_EINVAL)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from multi_modal_retrieval_predict_project_b200 import B200RetrievalEngine
    from multi_modal_retrieval_predict_project_b200.Helpers import precision_at_k
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        B200RetrievalEngine.from_arrays(np.zeros((4, 8), np.float32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        precision_at_k(["a"], ["a"], 1)
    # and the raw ABI refuses as well
    from multi_modal_retrieval_predict_project_b200 import _lib
    lib = _lib.load()
    h = C.c_void_p()
    x = np.zeros((4, 8), np.float32)
    st = lib.mmr_index_create(C.byref(h), x.ctypes.data, 4, 8, 0, 0, 0, 0, 0, None)
    assert st in (_lib.MMR_ENODEV, _lib.MMR_ECUDA)


def test_metric_id_interning():
    from multi_modal_retrieval_predict_project_b200.Helpers.retrieval_metrics import _intern
    ret, cnt, indptr, rel, ll = _intern([["a", "b", "a"], [], ["z"]], [["b", "b", "q"], ["a"], []])
    assert cnt.tolist() == [3, 0, 1] and ll.tolist() == [3, 1, 0]
    assert indptr.tolist() == [0, 2, 3, 3]
    assert ret[0, 0] == ret[0, 2] != ret[0, 1] and ret[1, 0] == -1
    assert sorted(rel[0:2].tolist()) == rel[0:2].tolist() and ret[0, 1] in rel[0:2]


def test_label_bit_masks_follow_int_eq_1_rule():
    from multi_modal_retrieval_predict_project_b200.Retrieval.reranker import Reranker
    df = pd.DataFrame({"id": ["a", "b", "c", "c"], "L0": [1, 0, 1, 1], "L1": [1.0, 1.9, np.nan, 0.0],
                       "txt": ["x", "1", "no", "2"], "L3": [True, False, True, True]}).set_index("id")
    r = Reranker.__new__(Reranker)
    r.labels_df = df
    masks, cols = r._label_bits()
    assert cols == ["L0", "L1", "txt", "L3"]
    assert masks[:, 0].tolist() == [0b1011, 0b0110, 0, 0]     # duplicated id "c" -> empty set


def test_shard_bounds_cover_rows():
    from multi_modal_retrieval_predict_project_b200.sharded import shard_bounds
    for n, w in ((10, 3), (100_000_000, 8), (5, 8), (0, 2)):
        b = [shard_bounds(n, w, r) for r in range(w)]
        assert b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))


def test_weighted_shard_bounds():
    from multi_modal_retrieval_predict_project_b200.sharded import shard_bounds, weighted_shard_bounds
    for n, w in ((10_000_000, [1.0, 1.08]), (10_000_000, [3.4, 3.6, 3.5, 3.55, 3.5, 3.5, 3.6, 3.52]), (1000, [1, 1, 1]),
                 (5, [1.0] * 8), (0, [1, 2]), (777, [0, 1, 0])):
        b = [weighted_shard_bounds(n, w, r) for r in range(len(w))]
        assert b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] and b[i][0] <= b[i][1] for i in range(len(w) - 1))
        assert all(x[1] % 256 == 0 or x[1] == n for x in b[:-1])
        if n >= 1_000_000:      # proportional to within a tile
            tot = sum(w)
            assert all(abs((hi - lo) - n * wi / tot) <= 256 for (lo, hi), wi in zip(b, w))
    # equal weights = equal shards up to the tile alignment; no weights at all = the plain split
    assert weighted_shard_bounds(1 << 20, [2.0, 2.0], 0) == shard_bounds(1 << 20, 2, 0)
    assert weighted_shard_bounds(1001, [0.0, 0.0], 1) == shard_bounds(1001, 2, 1)
    # a rank with all the weight holds every row (no remainder of the tile alignment leaks to the next rank)
    assert [weighted_shard_bounds(40001, [1.0, 0.0, 0.0], r) for r in range(3)] == [(0, 40001), (40001, 40001), (40001, 40001)]


def test_create_dump_embedding_matches_reference(tmp_path):
    """createDumpEmbedding writes the reference's merged files (Helpers/dumpEmbedding.py:28-39); when the
    reference is mounted its own function is run on a copy of the inputs and the outputs compared."""
    import json
    from multi_modal_retrieval_predict_project_b200.Helpers import createDumpEmbedding
    rng = np.random.default_rng(5)
    d = tmp_path / "embeddings"
    d.mkdir()
    tr, va = rng.standard_normal((7, 16)).astype(np.float32), rng.standard_normal((3, 16)).astype(np.float32)
    np.save(d / "train_joint_embeddings.npy", tr)
    np.save(d / "val_joint_embeddings.npy", va)
    json.dump([f"t{i}" for i in range(7)], open(d / "train_ids.json", "w"))
    json.dump([f"v{i}" for i in range(3)], open(d / "val_ids.json", "w"))
    createDumpEmbedding(tmp_path, d)
    merged = np.load(d / "trainval_joint_embeddings.npy")
    assert merged.dtype == np.float32 and np.array_equal(merged, np.concatenate([tr, va]))
    assert json.load(open(d / "trainval_ids.json")) == [f"t{i}" for i in range(7)] + [f"v{i}" for i in range(3)]
    ref_file = "/root/reference/src/Helpers/dumpEmbedding.py"
    if os.path.exists(ref_file):
        import importlib.util
        import sys
        sys.dont_write_bytecode = True
        spec = importlib.util.spec_from_file_location("_ref_dump_embedding", ref_file)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        d2 = tmp_path / "ref_embeddings"
        d2.mkdir()
        for f in ("train_joint_embeddings.npy", "val_joint_embeddings.npy", "train_ids.json", "val_ids.json"):
            (d2 / f).write_bytes((d / f).read_bytes())
        mod.EMBEDDINGS_DIR = d2          # the reference writes the ids next to its module-level default dir
        mod.createDumpEmbedding(tmp_path, d2)
        assert np.array_equal(np.load(d2 / "trainval_joint_embeddings.npy"), merged)
        assert json.load(open(d2 / "trainval_ids.json")) == json.load(open(d / "trainval_ids.json"))


def test_label_attention_pooling_matches_reference_module(tmp_path):
    """Reranker._pool with a LabelAttention checkpoint (Linear -> Tanh -> Linear -> softmax ->
    weighted sum, reference KnowledgeGraph/label_attention.py:11-27) and the mean-pool fallback
    (Retrieval/reranker.py:84-86,219-220).  The torch module is the reference's own class when
    /root/reference is mounted, an identical restatement otherwise."""
    import torch
    import torch.nn as nn
    from multi_modal_retrieval_predict_project_b200.Retrieval.reranker import Reranker
    ref_file = "/root/reference/src/KnowledgeGraph/label_attention.py"
    if os.path.exists(ref_file):
        import importlib.util
        import sys
        sys.dont_write_bytecode = True
        spec = importlib.util.spec_from_file_location("_ref_label_attention", ref_file)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        LabelAttention = mod.LabelAttention
    else:
        class LabelAttention(nn.Module):
            def __init__(self, d_emb, hidden=64):
                super().__init__()
                self.attn = nn.Sequential(nn.Linear(d_emb, hidden), nn.Tanh(), nn.Linear(hidden, 1))

            def forward(self, label_embs, mask=None):
                w = torch.softmax(self.attn(label_embs).squeeze(-1), dim=1)
                return torch.bmm(w.unsqueeze(1), label_embs).squeeze(1), w
    torch.manual_seed(3)
    d, hidden = 24, 16
    model = LabelAttention(d, hidden).eval()
    ckpt = tmp_path / "label_attention_model.pt"
    torch.save({"model_state": model.state_dict()}, ckpt)
    rr = object.__new__(Reranker)                       # host-side pooling only: no tables, no GPU
    rr.attn_params = Reranker._load_label_attention(ckpt, hidden)
    x = torch.randn(5, d)
    with torch.no_grad():
        want, _ = model(x.unsqueeze(0))
    got = rr._pool(x.numpy())
    assert got.shape == (d,) and np.allclose(got, want.squeeze(0).numpy(), rtol=1e-5, atol=1e-6)
    rr.attn_params = Reranker._load_label_attention(tmp_path / "missing.pt", hidden)
    assert rr.attn_params is None and np.allclose(rr._pool(x.numpy()), x.numpy().mean(axis=0))


def test_header_is_plain_c_and_links(tmp_path, lib):
    """include/mmr_b200.h is a plain-C (C99) header -- no C++ / torch types in the signatures -- and a C
    program links against libmmr_b200.so and calls it (no GPU needed for mmr_abi_version)."""
    import shutil
    import subprocess
    from multi_modal_retrieval_predict_project_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    src = tmp_path / "t.c"
    src.write_text('#include "mmr_b200.h"\n'
                   'int main(void) { return mmr_abi_version() == MMR_ABI_VERSION ? 0 : 1; }\n')
    inc = os.path.join(ROOT, "include")
    libdir = os.path.dirname(_lib.LIB_PATH)
    obj, exe = tmp_path / "t.o", tmp_path / "t"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc, "-c", str(src), "-o", str(obj)],
                   check=True)
    subprocess.run(["gcc", str(obj), "-L", libdir, "-l:" + os.path.basename(_lib.LIB_PATH), "-Wl,-rpath," + libdir,
                    "-o", str(exe)], check=True)
    assert subprocess.run([str(exe)]).returncode == 0
