#!/usr/bin/env python
"""Generate tests/golden/workload_ops.json: the oracle's §8(d) counters and algorithmic
lane-op totals (SURVEY.md §8d weight table) for the bench workloads.  bench.py reads the
committed JSON for its roofline numerator (it may not run the oracle for that); the GPU
tests check that the device path renders the same scenes bit-exactly.

    python tests/golden/make_workload_ops.py c1 c4 c2 c3      # c5 / c5b take many minutes
"""
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "workload_ops.json")


def workloads():
    d, l = O.scene_default(), O.light_default()
    return {
        "c1": (480, 320, 320, lambda: (d, l)),
        "c4": (1920, 1080, 1080, lambda: (d, l)),
        "c2": (3840, 2160, 2160, lambda: (d, l)),
        "c3": (3840, 2160, 2160, lambda: O.scene_synthetic(3840, 2160, 2160)),
        "c5": (7680, 4320, 4320, lambda: O.scene_synthetic(7680, 4320, 4320)),
        "c5b": (7680, 4320, 4320, lambda: O.scene_synthetic(7680, 4320, 4320, n=40000)),
    }


def main():
    res = json.load(open(OUT)) if os.path.exists(OUT) else {}
    wl = workloads()
    for name in sys.argv[1:] or ["c1", "c4", "c2", "c3"]:
        W, H, L, make = wl[name]
        boxes, lights = make()
        t = time.time()
        r = O.render(W, H, L, boxes, lights, want_texel=True)
        res[name] = {"view": [W, H, L], "n_entities": int(len(boxes)), "n_lights": int(len(lights)),
                     "rays": W * H * (1 + len(lights)), "counters": r["counters"],
                     "algorithmic_ops": r["ops"], "frame_fnv1a64": "%016x" % O.fnv1a64(r["rgba"]),
                     "frame_sha256": hashlib.sha256(r["rgba"].tobytes()).hexdigest(),
                     "gbuf_fnv1a64": "%016x" % O.fnv1a64(r["gbuf"]),      # raw Pixel[] bytes (28 B each)
                     "texel_fnv1a64": "%016x" % O.fnv1a64(r["texel"]),    # int32 texel-index plane, -1 = miss
                     "oracle_seconds_container": round(time.time() - t, 2)}
        print(name, res[name]["algorithmic_ops"], res[name]["oracle_seconds_container"], flush=True)
        json.dump(res, open(OUT, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
