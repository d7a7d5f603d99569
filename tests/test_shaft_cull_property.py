"""CPU property test of the exact-output shaft cull (csrc/shaft.cuh) against the oracle's
occlusion predicate: a culled box is missed by EVERY ray origin inside the group bounds.
The GPU parity suite checks this end to end on a few hundred scenes; this runs millions of
adversarial (grazing) configurations on the host, with the reciprocal error of the device's
approximate division modelled (tests/shaft_property.cpp)."""
import os
import subprocess

import pytest

from conftest import ROOT

CSRC = os.path.join(ROOT, "pixel-art-raytracer_b200", "csrc")


@pytest.fixture(scope="module")
def harness(tmp_path_factory, oracle):
    td = tmp_path_factory.mktemp("shaft")
    src = open(os.path.join(CSRC, "shaft.cuh")).read()
    assert '#include "par_device.cuh"' in src
    (td / "shaft_host.h").write_text(src.replace('#include "par_device.cuh"', ""))
    exe = td / "shaft_property"
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    # -ffp-contract=off: no FMA contraction, like the device build (-fmad=false)
    subprocess.run([cxx, "-O2", "-ffp-contract=off", "-I", str(td), os.path.join(ROOT, "tests", "shaft_property.cpp"),
                    "-L", os.path.join(ROOT, "oracle"), "-loracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle"),
                    "-o", str(exe)], check=True)
    return str(exe)


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_culled_boxes_are_missed_by_every_origin(harness, seed):
    res = subprocess.run([harness, "1500000", str(seed)], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    culled, kept = int(res.stdout.split()[1]), int(res.stdout.split()[3])
    assert culled > 100000 and kept > 100000, res.stdout  # the test exercises both outcomes
