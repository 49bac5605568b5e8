"""CPU: the reference arm of bench.py runs without a GPU and prints exactly one JSON line with the
contract's keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_contract():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1", "--rows", "50000", "--batch", "64"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "queries/s" and d["value"] > 0
    from oracle import ref_loader
    have_ref = ref_loader.reference_available() or ref_loader.staged_available()
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"] and "model" not in d["config"]


def test_traffic_table_schema():
    """profiles/traffic.json (ncu dram__bytes per launch, read by bench.py's roofline.traffic): one LIST of captured
    shapes per kernel, each with the workload it was captured on; the headline shape must be present."""
    tbl = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    kernels = {k: v for k, v in tbl.items() if not k.startswith("_")}
    assert {"gemm_topk_kernel", "scan_topk_kernel"} <= set(kernels)
    for name, entries in kernels.items():
        assert isinstance(entries, list) and entries, name
        for e in entries:
            assert {"rows", "dim", "batch", "k", "bytes", "source"} <= set(e), (name, e)
            assert e["bytes"] >= e["rows"] * e["dim"] * 2      # at least one pass over the bf16 gallery
    assert any((e["rows"], e["dim"], e["batch"], e["k"]) == (10_000_000, 512, 4096, 100) for e in kernels["gemm_topk_kernel"])
