"""CPU property test of the shadow-walk restructuring (distinct bins per step = non-empty subsets of
the changed axes) against the reference's 7-probe walk, as sets of flat bin indices
(tests/walk_property.cpp)."""
import os
import subprocess

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = tmp_path_factory.mktemp("walk") / "walk_property"
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-O2", "-ffp-contract=off", os.path.join(ROOT, "tests", "walk_property.cpp"), "-o", str(exe)],
                   check=True)
    return str(exe)


@pytest.mark.parametrize("seed", [1, 2])
def test_probed_bin_sets_are_equal(harness, seed):
    res = subprocess.run([harness, "300000", str(seed)], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert int(res.stdout.split()[3]) > 1000000, res.stdout
