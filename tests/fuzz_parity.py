#!/usr/bin/env python
"""Checker (not collected by pytest; run by hand on a B200): randomized GPU-vs-oracle parity sweep (beyond the fixed seeds of tests/).
Exercises the exact-output optimisations (shared walks, de-dup, shaft cull, near/far slab path)
on many scene shapes: dense/sparse, cubes/ragged boxes, lights inside/outside/on surfaces.

    python tests/fuzz_parity.py [n_scenes] [first_seed]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pixel-art-raytracer_b200"))
import par_b200 as par  # noqa: E402
from oracle import oracle as O  # noqa: E402


def scene(rng, W, H, L):
    n = int(rng.choice([50, 400, 2000, 6000]))
    a = np.zeros(n, par.AABB)
    spread = rng.choice([1.0, 0.3])
    a["px"] = rng.integers(-30, int(W * spread) + 30, n)
    a["py"] = rng.integers(-30, int(rng.choice([60, 200, 400])), n)
    a["pz"] = rng.integers(-60, int(L * spread) + 60, n)
    if rng.random() < 0.5:
        a["ex"] = a["ey"] = a["ez"] = 20
    else:
        a["ex"] = rng.integers(0, 21, n)
        a["ey"] = rng.integers(0, 21, n)
        a["ez"] = rng.integers(0, 21, n)
    if rng.random() < 0.3:  # snap to a lattice: many exact coincidences (NaN / tie cases)
        for f in ("px", "py", "pz"):
            a[f] = (a[f] // 20) * 20
    nl = int(rng.choice([1, 2, 5, 16, 33]))
    l = np.zeros(nl, par.LIGHT)
    l["x"] = rng.integers(-200, W + 200, nl)
    l["y"] = rng.integers(-100, 500, nl)
    l["z"] = rng.integers(-200, L + 200, nl)
    if rng.random() < 0.5:  # put some lights exactly on box corners / faces
        k = rng.integers(0, n, nl)
        on = rng.random(nl) < 0.5
        l["x"] = np.where(on, a["px"][k] + rng.integers(0, 21, nl), l["x"])
        l["y"] = np.where(on, a["py"][k] + a["ey"][k], l["y"])
        l["z"] = np.where(on, a["pz"][k] + rng.integers(0, 21, nl), l["z"])
    ns = int(rng.choice([1, 3, 9]))
    atlas = np.zeros(ns, par.SPRITE)
    atlas["color"] = rng.integers(0, 4, (ns, 800))
    atlas["depth"] = rng.integers(0, 20, (ns, 800))
    nrm = rng.standard_normal((ns, 800, 3)).astype(np.float32)
    if rng.random() < 0.5:
        nrm = np.round(nrm)  # axis-ish normals with exact zeros
    atlas["normal"] = nrm
    ids = rng.integers(0, ns, n).astype(np.int32)
    return a, l, atlas, ids


def main():
    n_scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    bad = 0
    for k in range(n_scenes):
        rng = np.random.default_rng(seed0 + k)
        W, H, L = [(480, 320, 320), (640, 480, 480), (320, 640, 640), (960, 360, 360)][k % 4]
        a, l, atlas, ids = scene(rng, W, H, L)
        with par.Renderer(W, H, L) as r:
            r.set_atlas(atlas, par.default_palette())
            r.set_scene(a, ids)
            rgba, gbuf, _ = r.render(l, want_gbuf=True)
        ref = O.render(W, H, L, a.view(O.AABB), l.view(O.LIGHT), atlas=atlas.view(O.SPRITE), sprite_ids=ids)
        ok = gbuf.tobytes() == ref["gbuf"].tobytes() and np.array_equal(rgba.view(np.uint32), ref["rgba"].view(np.uint32))
        # production frames (no G-buffer checkpoint) on either build of the render kernel (tile.cu / tile_one_light.cu)
        for flags in ("256", "512"):
            os.environ["PAR_DEBUG_FLAGS"] = flags
            with par.Renderer(W, H, L) as r:
                r.set_atlas(atlas, par.default_palette())
                r.set_scene(a, ids)
                prod, _ = r.render(l)
            ok = ok and np.array_equal(prod.view(np.uint32), ref["rgba"].view(np.uint32))
        os.environ.pop("PAR_DEBUG_FLAGS", None)
        if not ok:
            bad += 1
            diff = np.argwhere(rgba.view(np.uint32) != ref["rgba"].view(np.uint32))
            print(f"seed {seed0 + k}: MISMATCH {len(diff)} px, first {diff[:3].tolist()}", flush=True)
    print(f"fuzz: {n_scenes - bad}/{n_scenes} scenes bit-exact")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
