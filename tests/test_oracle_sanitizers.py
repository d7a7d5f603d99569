"""The oracle under ASan + UBSan: the reference reads outside its arrays on its own default scene
(quirk Q18; ASan flags alternative.cpp:476), the oracle defines those reads and must run clean — on the default
scene, on lights far outside the grid with entities straddling the view volume, and on a dense synthetic scene —
while producing the same frames as the optimised build."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import ROOT


def test_oracle_is_clean_under_asan_and_ubsan(oracle, tmp_path):
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    exe = tmp_path / "oracle_sanitize"
    build = subprocess.run([gcc, "-O1", "-g", "-std=gnu11", "-fopenmp", "-ffp-contract=off", "-fno-fast-math",
                            "-fsanitize=address,undefined", "-fno-sanitize-recover=all",
                            os.path.join(ROOT, "tests", "oracle_sanitize_main.c"), "-o", str(exe), "-lm"],
                           capture_output=True, text=True)
    if build.returncode != 0 and "sanitize" in build.stderr:
        pytest.skip("sanitizer runtimes not installed")
    assert build.returncode == 0, build.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1"))
    assert run.returncode == 0, run.stderr[-3000:]
    assert "ERROR" not in run.stderr and "runtime error" not in run.stderr, run.stderr[-3000:]
    got = dict(ln.split() for ln in run.stdout.splitlines())
    # the same frames from the optimised library (the one every other test uses)
    O = oracle
    c1 = O.render(480, 320, 320, O.scene_default(), O.light_default(), want_gbuf=False, want_texel=False)
    assert got["c1"] == "%016x" % O.fnv1a64(c1["rgba"])
    boxes, lights = O.scene_synthetic(640, 680, 680, n=3000)
    syn = O.render(640, 680, 680, boxes, lights, want_gbuf=False, want_texel=False)
    assert got["synthetic"] == "%016x" % O.fnv1a64(syn["rgba"])
    assert len(got["edges"]) == 16
