#!/usr/bin/env python
"""Developer tool: ncu launch list (csv of `--metrics gpu__time_duration.sum`) -> per-kernel totals and the
shares of the step's own kernels (profiles/r02_bench_c2_launches_summary.txt).

    tools/launch_summary.py gpurun_out/r02_final_bench_c2_launches.csv > profiles/r02_bench_c2_launches_summary.txt
"""
import collections
import csv
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

rows = [r for r in csv.reader(ln for ln in open(sys.argv[1]) if ln.startswith('"'))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(",", "")) * (1000 if r[ui] == "us" else 1)
    a = agg.setdefault(r[ki], [0, 0.0])
    a[0] += 1
    a[1] += v
print("# ncu launch list of `python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-scale-8k --no-c4` (--metrics "
      f"gpu__time_duration.sum --clock-control none, first 400 launches), sources of hash {bench.kernel_source_sha()}")
print("# per-launch times are cold-cache and serialised: compare SHARES with bench.py's kernels_ms, not absolutes")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:5d} launches {t / 1e3:11.1f} us total {t / 1e3 / n:9.2f} us avg  {k[:70]}")
own = {k: v for k, v in agg.items() if "par::" in k and "k_clear_grid" not in k}
tot = sum(v[1] for v in own.values())
print()
print("share of the step's own kernels: " + ", ".join(f"{k.split('(')[0]} {v[1] / tot * 100:.1f}%" for k, v in own.items()))
