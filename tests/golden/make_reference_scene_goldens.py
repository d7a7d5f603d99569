#!/usr/bin/env python
"""Generate tests/golden/reference_scenes.json by RUNNING THE REAL REFERENCE on arbitrary scenes.

The tier-1 builds of oracle/build_ref.sh carry a scene hook (par_stub_load_scene, called once
before the frame loop): with PAR_REF_SCENE set, the built-in scene and light are replaced by a
file's, inserted through the reference's own Entities::insert.  This script makes random scenes
(dense / sparse, cubes / ragged extents, lattice-snapped coordinates, the light on a box face or
free), renders them with the real reference at several view sizes — including view length !=
view height, which the default-scene goldens do not cover — and records hashes of the G-buffer,
the shaded frame before the debug overlay and the final frames.  The scenes themselves are stored
in the JSON (base64 of the scene file), so the tests do not depend on any RNG.

Lights are kept inside the view volume: a light bin outside the grid makes the reference read
outside its arrays (quirk Q18), which is undefined behaviour in the real binary and therefore
not something a golden vector can pin.

    python tests/golden/make_reference_scene_goldens.py
"""
import base64
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.path.join(ROOT, "oracle", "_ref")
OUT = os.path.join(ROOT, "tests", "golden", "reference_scenes.json")


def make_scene(rng, W, H, L, n, n_lights, ragged, lattice, lights_on_boxes, frames=1):
    b = np.zeros((n, 8), np.int16)
    b[:, 0] = rng.integers(-30, W + 30, n)
    b[:, 1] = rng.integers(-30, max(H // 2, 40), n)
    b[:, 2] = rng.integers(-60, L + 60, n)
    if ragged:
        b[:, 3] = rng.integers(0, 21, n)
        b[:, 4] = rng.integers(0, 21, n)
        b[:, 5] = rng.integers(0, 21, n)
    else:
        b[:, 3:6] = 20
    if lattice:  # many exact coincidences: depth ties, zero direction components
        b[:, 0:3] = (b[:, 0:3] // 20) * 20
    li = np.zeros((n_lights, 4), np.int16)
    for k in range(n_lights):
        li[k] = light_inside(rng, W, H, L, frames)
        if lights_on_boxes and k % 2 == 0:  # exactly on a box's top face
            for _ in range(50):
                e = int(rng.integers(0, n))
                x2, y2, z2 = int(b[e, 0]) + int(rng.integers(0, 21)), int(b[e, 1]) + int(b[e, 4]), int(b[e, 2]) + int(rng.integers(0, 21))
                if 5 <= x2 < W - 10 - 5 * frames and 5 <= z2 < L - 5 and 5 <= H - y2 - z2 <= H - 5:
                    li[k] = (x2, y2, z2, 10)
                    break
    return b, li


def light_inside(rng, W, H, L, frames=4):
    """A light whose bin (x/40, (H-y-z)/40, z/40) is inside the grid, with room for one 'o' key (x += 5) per frame."""
    x = int(rng.integers(5, max(W - 10 - 5 * frames, 6)))
    z = int(rng.integers(5, L - 5))
    row = int(rng.integers(5, H - 4))  # H - y - z
    return (x, H - row - z, z, 10)


def scene_bytes(boxes, lights):
    return np.array([len(boxes), len(lights)], np.int32).tobytes() + boxes.tobytes() + lights.tobytes()


def sha(b):
    return hashlib.sha256(b).hexdigest()


def run_reference(view, scene, frames=1, script=None):
    W, H, L = view
    exe = os.path.join(REF, f"ref_tier1_{W}x{H}x{L}")
    with tempfile.TemporaryDirectory() as td:
        env = dict(os.environ, PAR_REF_FRAMES=str(frames), PAR_REF_HASHES=f"{td}/h.txt",
                   PAR_REF_DUMP_PRE=f"{td}/pre.rgba", PAR_REF_DUMP_GBUF=f"{td}/gbuf.bin", PAR_REF_DUMP_FRAMES="0")
        if scene is not None:
            open(f"{td}/scene.bin", "wb").write(scene)
            env["PAR_REF_SCENE"] = f"{td}/scene.bin"
        if script:
            env["PAR_REF_SCRIPT"] = script
        subprocess.run([exe], env=env, check=True, stdout=subprocess.DEVNULL)
        return {"fnv1a64": [ln.split()[1] for ln in open(f"{td}/h.txt")],
                "frame0_pre_overlay_sha256": sha(open(f"{td}/pre.rgba", "rb").read()),
                "gbuf0_sha256": sha(open(f"{td}/gbuf.bin", "rb").read())}


# The reference shades with lights[0] only (alternative.cpp:712-733 index the vector with 0; more
# lights are this repo's SURVEY.md 8d extension and cannot be pinned by the reference), so every
# scene carries ONE light; variety comes from where it sits.
CASES = [  # view, n boxes, ragged, lattice, light on a box face, frames, script
    ((480, 320, 320), 400, False, False, False, 1, None),
    ((480, 320, 320), 1200, True, False, True, 1, None),
    ((480, 320, 320), 1000, True, True, True, 1, None),
    ((480, 320, 320), 800, False, True, False, 1, None),
    ((480, 320, 320), 60, False, True, False, 4, "D"),
    ((480, 320, 320), 1200, True, False, False, 3, "D"),
    ((480, 320, 640), 900, True, False, True, 1, None),
    ((480, 320, 640), 1200, False, True, False, 3, "D"),
    ((480, 320, 640), 700, True, True, False, 1, None),
    ((640, 480, 200), 1000, True, False, True, 1, None),
    ((640, 480, 200), 1200, False, False, False, 3, "C"),
    ((640, 480, 200), 500, True, True, True, 1, None),
    ((200, 40, 40), 150, True, True, True, 1, None),
    ((200, 40, 40), 40, False, False, False, 1, None),
    ((40, 1000, 120), 300, True, False, True, 1, None),
    ((40, 1000, 120), 200, False, True, False, 2, "D"),
    # long scripted runs: entity 0 walks through a ragged scene (bins change under it), the light moves too
    ((480, 320, 640), 1200, True, False, False, 120, "C"),
    ((640, 480, 200), 900, True, True, True, 20, "D"),
]


def main():
    gold = {"_how": "written by tests/golden/make_reference_scene_goldens.py: the REAL reference (tier-1 builds of "
                    "oracle/build_ref.sh, g++ -O3, x86-64) rendering scene files through its scene hook; "
                    "scene = base64 of int32 n_boxes, int32 n_lights, n_boxes x 8 int16, n_lights x 4 int16",
            "default_scene": {}, "scenes": []}
    # the built-in scene (its boxes come from the oracle's restatement, pinned by reference_hashes.json)
    # with the light moved inside the grid of each view: the built-in light sits one bin outside
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    default_boxes = O.scene_default().view(np.int16).reshape(-1, 8)
    for k, view in enumerate([(480, 320, 320), (480, 320, 640), (640, 480, 200), (200, 40, 40), (40, 1000, 120)]):
        print("default scene", view, flush=True)
        light = np.array([light_inside(np.random.default_rng(500 + k), *view)], np.int16)
        try:
            res = run_reference(view, scene_bytes(default_boxes, light), frames=3, script="D")
        except subprocess.CalledProcessError as e:
            print("   reference crashed:", e, flush=True)
            continue
        res.update({"view": list(view), "light": light[0].tolist(), "frames": 3, "script": "D"})
        gold["default_scene"]["%dx%dx%d" % view] = res
    for k, (view, n, ragged, lattice, on_boxes, frames, script) in enumerate(CASES):
        print("scene", k, view, n, flush=True)
        nl = 1
        rng = np.random.default_rng(1000 + k)
        boxes, lights = make_scene(rng, *view, n, nl, ragged, lattice, on_boxes, frames)
        blob = scene_bytes(boxes, lights)
        try:
            res = run_reference(view, blob, frames=frames, script=script)
        except subprocess.CalledProcessError as e:  # the real reference crashed: undefined behaviour, nothing to pin
            print("   reference crashed:", e, flush=True)
            continue
        res.update({"view": list(view), "n_boxes": n, "n_lights": nl, "frames": frames, "script": script,
                    "scene_b64": base64.b64encode(blob).decode()})
        gold["scenes"].append(res)
    json.dump(gold, open(OUT, "w"), indent=1, sort_keys=True)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    if not os.path.exists(os.path.join(REF, "ref_tier1_480x320x640")):
        sys.exit("oracle/_ref tier-1 builds missing: run oracle/build_ref.sh first")
    main()
