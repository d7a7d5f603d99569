#!/usr/bin/env python
"""Developer tool: aggregate a k_shade ncu profile by kernel phase (markers found in shade.cu)."""
import re
import subprocess
import sys

rep = sys.argv[1]
src = open("pixel-art-raytracer_b200/csrc/shade.cu").read().splitlines()
marks = [("slab tests (device functions)", r"^__device__ __forceinline__ bool slab_hit_exact"),
         ("block scan / helpers", r"^// Block-wide exclusive scan"),
         ("prologue + tile load", r"^k_shade\("),
         ("find group", r"---- next group"),
         ("compact group", r"---- compact the group"),
         ("fetch precomputed lists", r"fast path: the walks of this group"),
         ("round setup (A/B)", r"// A\. describe the trial segments"),
         ("walk (C)", r"// C\. phase 1"),
         ("decide (D)", r"// D\. how many leading segments"),
         ("gather (E)", r"// E\. phase 2"),
         ("shade (F)", r"// F\. phase 3"),
         ("advance + store", r"// advance past the processed segments")]
starts = []
for name, pat in marks:
    for i, ln in enumerate(src, 1):
        if re.search(pat, ln):
            starts.append((i, name))
            break
starts.sort()


def phase(f, l):
    if f == "shaft.cuh":
        return "gather (E)"
    if f != "shade.cu":
        return "inlined: " + f
    name = "file header"
    for s, n in starts:
        if l >= s:
            name = n
    return name


out = subprocess.run([sys.executable, "tools/ncu_lines.py", rep, "shade", "k_shade", "--top", "2000", "--by", "inst"],
                     capture_output=True, text=True).stdout.splitlines()
print(out[0])
agg = {}
for ln in out[1:]:
    m = re.match(r"\s*([\d.]+)% inst\s+([\d.]+)% smp\s+lanes\s+([\d.]+)\s+(\S+):(\d+)", ln)
    if m:
        i, s, la, f, l = float(m.group(1)), float(m.group(2)), float(m.group(3)), m.group(4), int(m.group(5))
        a = agg.setdefault(phase(f, l), [0, 0, 0])
        a[0] += i
        a[1] += i * la
        a[2] += s
tt = sum(a[1] for a in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{v[0]:5.1f}% warp-inst {v[1] / tt * 100:5.1f}% thread-inst {v[2]:5.1f}% samples  lanes {v[1] / max(v[0], 1e-9):4.1f}  {k}")
