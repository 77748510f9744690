"""bench.py's contract, as far as it can be checked without a GPU: the reference arm (the CPU restatement of the reference's own
search path, oracle/) runs on a tiny workload and prints ONE JSON line with the agreed keys; our arm refuses to run without a CUDA
device instead of falling back to anything."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=600, env=e, cwd=ROOT)


def test_reference_arm_prints_one_contract_line():
    out = _run(["--impl", "reference", "--text-bytes", "150000", "--queries", "25000", "--steps", "2", "--warmup", "1"])
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "queries/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["metric"] == "fm_count_queries_per_s_len16_1GB_text" and d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0
    assert d["config"]["workload"].startswith("cfg2") and d["config"]["pattern_len"] == 16
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_other_workloads():
    """--workload cfg3 (count + sorted sa[sp..ep)) and cfg4 (ReTree._matchSA) through the oracle, scaled down"""
    for w, extra, unit in (("cfg3", ["--queries", "3000"], "queries/s"), ("cfg4", ["--queries", "200"], "regexes/s")):
        out = _run(["--impl", "reference", "--workload", w, "--text-bytes", "200000", "--steps", "1", "--warmup", "1"] + extra)
        assert out.returncode == 0, out.stderr[-2000:]
        d = json.loads([ln for ln in out.stdout.splitlines() if ln.strip()][-1])
        assert d["impl"] == "reference" and d["unit"] == unit and d["value"] > 0 and d["config"]["workload"].startswith(w)
        assert d["cpu_baseline"]["kind"] == "port"


def test_our_arm_needs_a_gpu_and_says_so():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    out = _run(["--text-bytes", "150000", "--queries", "25000", "--steps", "1", "--warmup", "3", "--no-cpu", "--regexes", "0"])
    assert out.returncode != 0
    assert "needs a CUDA device" in (out.stderr + out.stdout)
    assert not [ln for ln in out.stdout.splitlines() if ln.strip().startswith("{")]       # no number without the device
