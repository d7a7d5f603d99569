#!/usr/bin/env python
"""Developer tool: resource usage and SASS opcode histograms of every kernel in the built libpar_b200.so
-> profiles/r02_sass_summary.txt (+ the full SASS of the render kernels, gzipped).  No GPU needed."""
import collections
import gzip
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

SO = os.path.join(ROOT, "pixel-art-raytracer_b200", "par_b200", "libpar_b200.so")


def main():
    res = subprocess.run(["cuobjdump", "-res-usage", SO], capture_output=True, text=True, check=True).stdout
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    out = [f"# SASS summary of the shipped libpar_b200.so (sm_100a), build of source hash {bench.kernel_source_sha()}",
           "# cuobjdump -res-usage"]
    fn = None
    for ln in res.splitlines():
        m = re.match(r"\s*Function (\S+):", ln)
        if m:
            fn = m.group(1)
        elif fn and "REG:" in ln:
            out.append(f"{fn}: {ln.strip()}")
            fn = None
    cur, hist = None, collections.OrderedDict()
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = hist.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
        if m and cur is not None:
            cur[m.group(1)] += 1
    for fn, h in hist.items():
        n = sum(h.values())
        out.append("")
        out.append(f"## {fn}: {n} SASS instructions ({n * 16 // 1024} KB)")
        out.append("   " + ", ".join(f"{k} {v}" for k, v in h.most_common(24)))
        out.append("   " + ", ".join(f"{k} {h.get(k, 0)}" for k in ("FFMA", "FMUL", "FADD", "FMNMX", "MUFU", "BAR", "REDUX", "ATOMS",
                                                                   "LDGSTS", "UTMALDG"))
                   + "   (FFMA only inside IEEE division / reciprocal sequences: -fmad=false)")
    with open(os.path.join(ROOT, "profiles", "r02_sass_summary.txt"), "w") as f:
        f.write("\n".join(out) + "\n")
    keep, on = [], False
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            on = "k_tile" in m.group(1)
        if on:
            keep.append(ln)
    with gzip.open(os.path.join(ROOT, "profiles", "r02_tile_sm100a.sass.gz"), "wt") as f:
        f.write("\n".join(keep) + "\n")
    print("\n".join(out[:3]), f"\n... {len(hist)} kernels")


if __name__ == "__main__":
    main()
