#!/usr/bin/env python
"""Developer tool: aggregate a k_tile ncu profile by kernel phase (markers found in tile.cu)."""
import re
import subprocess
import sys

rep = sys.argv[1]
src = open("pixel-art-raytracer_b200/csrc/tile.cu").read().splitlines()
marks = [("slab tests (device functions)", r"^__device__ __forceinline__ bool slab_hit_exact"),
         ("box store / octant / quantise helpers", r"^// Box -> shared list slot"),
         ("prologue", r"^k_tile\(const __grid_constant__"),
         ("P: column counts + scan", r"P\. primary rays ====="),
         ("P: entry gather", r"gather the chunk's entries"),
         ("P: per-pixel walk", r"// per-pixel walk: the entry list"),
         ("P: records", r"// records of this pass"),
         ("miss pixels + group keys", r"---- miss pixels"),
         ("G: find groups", r"---- smallest unprocessed group"),
         ("G: group bounds", r"---- pixels per group and bounds"),
         ("G: scatter into lists", r"exclusive scan of the group sizes"),
         ("R: round setup (A)", r"// A\. describe the trial segments"),
         ("R: walk (C)", r"// C\. walk\."),
         ("R: decide (D)", r"// D\. how many leading segments"),
         ("R: gather (E)", r"// E\. gather"),
         ("R: shade (F)", r"// F\. shade"),
         ("R: advance", r"// advance past the processed segments"),
         ("tail: 16-byte stores (+ peers)", r"---- 16-byte stores of the finished tile rows")]
starts = []
for name, pat in marks:
    for i, ln in enumerate(src, 1):
        if re.search(pat, ln):
            starts.append((i, name))
            break
starts.sort()


def phase(f, l):
    if f == "shaft.cuh":
        return "R: gather (E)"
    if f != "tile.cu":
        return "inlined: " + f
    name = "file header"
    for s, n in starts:
        if l >= s:
            name = n
    return name


# argv: report [mangled-name substring [cubin stem]] — e.g. k_tile_one_lightILb0 tile_one_light for the one-light build
out = subprocess.run([sys.executable, "tools/ncu_lines.py", rep, sys.argv[3] if len(sys.argv) > 3 else "tile", "k_tile", "--symbol",
                      sys.argv[2] if len(sys.argv) > 2 else "k_tileILb0", "--top", "2000", "--by", "inst"],
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
