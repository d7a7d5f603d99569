"""Shared test plumbing.

`-m "not gpu"` : oracle vs the reference's golden vectors, host logic, C-ABI load/exports.
`-m gpu`       : parity of the CUDA path (through the C ABI) against the oracle, on a B200.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "pixel-art-raytracer_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")
    config.addinivalue_line("markers", "slow: long CPU test, enabled with PAR_SLOW=1")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure), compiled on demand with gcc."""
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def par():
    """The product binding; builds libpar_b200.so with nvcc when it is missing."""
    import par_b200
    if not os.path.exists(par_b200.LIB_PATH):
        import subprocess
        subprocess.run([os.path.join(PKG, "build_native.sh")], check=True)
    par_b200.lib()
    return par_b200


@pytest.fixture(scope="session")
def golden():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "reference_hashes.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_scenes():
    """Hashes of what the REAL reference renders for the scene files stored beside them
    (tests/golden/make_reference_scene_goldens.py)."""
    import json
    with open(os.path.join(ROOT, "tests", "golden", "reference_scenes.json")) as f:
        return json.load(f)


def decode_scene(entry, aabb_dtype, light_dtype):
    """(boxes, lights) of one reference_scenes.json entry as structured arrays."""
    import base64
    import numpy as np
    blob = base64.b64decode(entry["scene_b64"])
    n, nl = np.frombuffer(blob, np.int32, 2)
    boxes = np.frombuffer(blob, aabb_dtype, n, 8).copy()
    lights = np.frombuffer(blob, light_dtype, nl, 8 + 16 * n).copy()
    return boxes, lights


def sha256(arr):
    import hashlib
    import numpy as np
    return hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest()
