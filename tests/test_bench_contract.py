"""bench.py's contract, as far as it can be checked without a GPU: the reference arm's JSON line
(run here on the small C1 workload) and the product arm's refusal to run without a CUDA device."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

BENCH = os.path.join(ROOT, "bench.py")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_reference_arm_line():
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ref_tier1_480x320x320")):
        pytest.skip("oracle/_ref not built")
    res = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--workload", "c1", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "Mrays/s" and line["unit"] == "Mrays/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 2 and line["warmup"] == 1
    assert line["value"] > 0 and line["ms_per_step"] > 0
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] == 1
    assert line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"].startswith("c1")


def test_product_arm_refuses_to_run_without_a_gpu():
    """No CPU fallback: without a CUDA device the product arm exits non-zero and prints no result line."""
    if _has_gpu():
        pytest.skip("a GPU is present")
    res = subprocess.run([sys.executable, BENCH, "--steps", "1", "--warmup", "1", "--no-cpu-baseline"],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode != 0
    assert "no CUDA device" in (res.stderr + res.stdout)
    assert not any(ln.startswith("{") for ln in res.stdout.splitlines())


def test_cpu_baseline_object_of_the_product_arm():
    """The cpu_baseline leg bench.py attaches to its own line (real reference, 1 thread, plus the
    oracle port on all cores), run here on C1."""
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ref_tier1_480x320x320")):
        pytest.skip("oracle/_ref not built")
    import argparse
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_module", BENCH)
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    cb = bench.cpu_baseline(argparse.Namespace(workload="c1", cpu_budget=5.0), 480, 320, 320)
    assert cb["kind"] == "reference" and cb["cores"] == 1 and cb["unit"] == "Mrays/s" and cb["value"] > 0
    assert "sample" in cb and cb["port_all_cores"]["kind"] == "port" and cb["port_all_cores"]["value"] > 0
