#!/usr/bin/env python
"""Generate tests/golden/reference_hashes.json by RUNNING THE REAL REFERENCE.

Runs the headless builds of Cons-Cat/Pixel-Art-Raytracer made by oracle/build_ref.sh
(tier 0 = unmodified source + SDL stub, tier 1 = sed-parameterised view size) and records
hashes of what they render.  These are the golden vectors that pin the oracle (the reference
ships none of its own — SURVEY.md §4).  Needs /root/reference, so it runs only in the build
container; the JSON it writes is committed and is what travels.

    python tests/golden/make_reference_goldens.py [--quick]   # --quick skips the 9-minute
                                                              # 1080p 240-frame script D run
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.path.join(ROOT, "oracle", "_ref")
OUT = os.path.join(ROOT, "tests", "golden", "reference_hashes.json")

PIXEL_BYTES = 28


def sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 22), b""):
            h.update(chunk)
    return h.hexdigest()


def frame_ms(ns):
    return {"mean": round(sum(ns) / len(ns) / 1e6, 2), "min": round(min(ns) / 1e6, 2),
            "max": round(max(ns) / 1e6, 2),
            "note": "alternative.cpp:689-772 as timed by the SDL stub in the build container, "
                    "1 thread; informational"}


def run(binary, frames=1, script=None, dump=True, dump_pre=False):
    with tempfile.TemporaryDirectory() as td:
        env = dict(os.environ, PAR_REF_FRAMES=str(frames), PAR_REF_HASHES=f"{td}/h.txt",
                   PAR_REF_TIMES=f"{td}/t.txt")
        if script:
            env["PAR_REF_SCRIPT"] = script
        if dump:
            env["PAR_REF_DUMP"] = f"{td}/frame.rgba"
        if dump_pre:
            env["PAR_REF_DUMP_PRE"] = f"{td}/pre.rgba"
            env["PAR_REF_DUMP_GBUF"] = f"{td}/gbuf.bin"
        subprocess.run([os.path.join(REF, binary)], env=env, check=True,
                       stdout=subprocess.DEVNULL)
        res = {"fnv1a64": [ln.split()[1] for ln in open(f"{td}/h.txt")],
               "hash_file_sha256": sha(f"{td}/h.txt"),
               "frame_ms_container": frame_ms([int(ln.split()[1]) for ln in open(f"{td}/t.txt")])}
        if dump:
            res["frame0_sha256"] = sha(f"{td}/frame.rgba")
        if dump_pre:
            res["frame0_pre_overlay_sha256"] = sha(f"{td}/pre.rgba")
            res["gbuf0_sha256"] = sha(f"{td}/gbuf.bin")
            import numpy as np
            g = np.fromfile(f"{td}/gbuf.bin", dtype=np.uint8).reshape(-1, PIXEL_BYTES)
            for name, off in (("entity", 24), ("y", 16), ("z", 20)):
                plane = np.ascontiguousarray(g[:, off:off + 4])
                res[f"gbuf0_{name}_sha256"] = hashlib.sha256(plane.tobytes()).hexdigest()
        return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    if not os.path.exists(os.path.join(REF, "ref_tier0")):
        sys.exit("oracle/_ref/ref_tier0 missing: run oracle/build_ref.sh first")
    gold = json.load(open(OUT)) if os.path.exists(OUT) else {}
    gold["_how"] = ("written by tests/golden/make_reference_goldens.py from the real reference "
                    "built by oracle/build_ref.sh; g++ -O3, x86-64")
    print("tier0 480x320 frame 0", flush=True)
    gold["tier0_480x320x320_frame0"] = run("ref_tier0")
    print("tier0 480x320 script C x240", flush=True)
    gold["tier0_480x320x320_scriptC_240"] = run("ref_tier0", frames=240, script="C", dump=False)
    print("tier1 480x320 frame 0 (+pre-overlay, gbuf)", flush=True)
    gold["tier1_480x320x320_frame0"] = run("ref_tier1_480x320x320", dump_pre=True)
    print("tier1 1920x1080 frame 0", flush=True)
    gold["tier1_1920x1080x1080_frame0"] = run("ref_tier1_1920x1080x1080", dump_pre=True)
    print("tier1 3840x2160 frame 0", flush=True)
    gold["tier1_3840x2160x2160_frame0"] = run("ref_tier1_3840x2160x2160", dump_pre=True)
    if not args.quick:
        print("tier1 1920x1080 script D x240 (about 9 minutes)", flush=True)
        gold["tier1_1920x1080x1080_scriptD_240"] = run("ref_tier1_1920x1080x1080", frames=240,
                                                      script="D", dump=False)
    json.dump(gold, open(OUT, "w"), indent=1, sort_keys=True)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
