"""CPU property test of k_tile's three slab-test variants (exact ternaries, fminf/fmaxf, near/far
corners) against the oracle's occlusion predicate — the functions are extracted verbatim from
csrc/tile.cu and csrc/par_device.cuh and compiled for the host (tests/slab_property.cpp)."""
import os
import re
import subprocess

import pytest

from conftest import ROOT

CSRC = os.path.join(ROOT, "pixel-art-raytracer_b200", "csrc")


def _extract(text, start_pat, end_pat):
    a = re.search(start_pat, text, re.M)
    assert a, f"marker not found: {start_pat}"
    b = re.search(end_pat, text[a.start():], re.M)
    assert b, f"end marker not found: {end_pat}"
    return text[a.start():a.start() + b.end()]


@pytest.fixture(scope="module")
def harness(tmp_path_factory, oracle):
    td = tmp_path_factory.mktemp("slab")
    dev = open(os.path.join(CSRC, "par_device.cuh")).read()
    shade = open(os.path.join(CSRC, "tile.cu")).read()
    parts = [_extract(dev, r"^__device__ __forceinline__ float std_min", r"std_max\(float a, float b\) \{[^}]*\}"),
             _extract(shade, r"^__device__ __forceinline__ bool slab_hit_exact", r"^\}"),
             _extract(shade, r"^__device__ __forceinline__ bool slab_hit_fast", r"^\}"),
             _extract(shade, r"^__device__ __forceinline__ bool slab_hit_near_far", r"^\}")]
    (td / "slab_host.h").write_text("\n".join(parts) + "\n")
    exe = td / "slab_property"
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-O2", "-ffp-contract=off", "-I", str(td), os.path.join(ROOT, "tests", "slab_property.cpp"),
                    "-L", os.path.join(ROOT, "oracle"), "-loracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle"),
                    "-o", str(exe)], check=True)
    return str(exe)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_slab_variants_equal_the_reference_predicate(harness, seed):
    res = subprocess.run([harness, "3000000", str(seed)], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    f = res.stdout.split()
    zero_dir, hits = int(f[3]), int(f[5])
    assert zero_dir > 100000 and hits > 100000, res.stdout  # NaN cases and real hits are both exercised
