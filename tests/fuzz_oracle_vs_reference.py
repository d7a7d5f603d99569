#!/usr/bin/env python
"""Checker (not collected by pytest; needs /root/reference's headless builds in oracle/_ref):
randomized sweep of the oracle against the REAL reference through the scene hook of the tier-1
builds.  Complements the fixed scenes of tests/golden/reference_scenes.json.

    python tests/fuzz_oracle_vs_reference.py [n_scenes] [first_seed]
"""
import hashlib
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from oracle import oracle as O  # noqa: E402
from make_reference_scene_goldens import make_scene, scene_bytes  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")
VIEWS = [(480, 320, 320), (480, 320, 640), (640, 480, 200), (200, 40, 40), (40, 1000, 120)]


def main():
    n_scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    bad = crashed = 0
    for k in range(n_scenes):
        rng = np.random.default_rng(77000 + seed0 + k)
        W, H, L = VIEWS[int(rng.integers(0, len(VIEWS)))]
        n = int(rng.choice([20, 150, 600, 1500]))
        boxes, lights = make_scene(rng, W, H, L, n, 1, bool(rng.integers(0, 2)), bool(rng.integers(0, 2)),
                                   bool(rng.integers(0, 2)))
        with tempfile.TemporaryDirectory() as td:
            open(f"{td}/s.bin", "wb").write(scene_bytes(boxes, lights))
            env = dict(os.environ, PAR_REF_FRAMES="1", PAR_REF_SCENE=f"{td}/s.bin", PAR_REF_DUMP_PRE=f"{td}/pre.rgba",
                       PAR_REF_DUMP_GBUF=f"{td}/g.bin")
            r = subprocess.run([os.path.join(REF, f"ref_tier1_{W}x{H}x{L}")], env=env, stdout=subprocess.DEVNULL)
            if r.returncode != 0:
                crashed += 1
                print(f"seed {seed0 + k}: the reference crashed (rc {r.returncode}) at {W}x{H}x{L}, n={n}", flush=True)
                continue
            ref_rgba = open(f"{td}/pre.rgba", "rb").read()
            ref_gbuf = open(f"{td}/g.bin", "rb").read()
        o = O.render(W, H, L, boxes.view(O.AABB).reshape(-1), lights.view(O.LIGHT).reshape(-1))
        same = o["rgba"].tobytes() == ref_rgba and o["gbuf"].tobytes() == ref_gbuf
        if not same:
            bad += 1
            a = np.frombuffer(ref_rgba, np.uint32).reshape(H, W)
            d = np.argwhere(a != o["rgba"].view(np.uint32).reshape(H, W))
            print(f"seed {seed0 + k}: MISMATCH at {W}x{H}x{L}, n={n}: {len(d)} px differ, first {d[:3].tolist()}, "
                  f"gbuf equal: {o['gbuf'].tobytes() == ref_gbuf}, light {lights[0].tolist()}", flush=True)
    print(f"fuzz: {n_scenes - bad - crashed}/{n_scenes} scenes identical, {bad} mismatches, {crashed} reference crashes")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
