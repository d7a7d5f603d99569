#!/usr/bin/env python
"""Developer tool: turn `ncu --set full` captures of the render kernel into profiles/roofline_r02.json,
the file bench.py's `roofline` object quotes (executed warp instructions, DRAM traffic, issue-active).

    tools/roofline_from_ncu.py c2=gpurun_out/r02_final_tile_c2 c3=gpurun_out/r02_final_tile_c3

Each capture `<stem>.ncu-rep` must come with `<stem>.srcsha`, written ON THE GPU BOX by
`tools/profile_kernel.sh capture` from the sources the captured library was built from; the tool refuses
captures whose hash differs from the current tree, and bench.py refuses the file when the tree has moved on.
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

WANT = {
    "warp_instructions": "smsp__inst_executed.sum",
    "dram_bytes_read": "dram__bytes_read.sum",
    "dram_bytes_write": "dram__bytes_write.sum",
    "duration_us_under_ncu": "gpu__time_duration.sum",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "registers_per_thread": "launch__registers_per_thread",
    "alu_pipe_pct": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "fma_pipe_pct": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "lsu_pipe_pct": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
}
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1, "ms": 1e3, "ns": 1e-3, "s": 1e6}


def main():
    sha = bench.kernel_source_sha()
    out = {"_how": "tools/roofline_from_ncu.py over `ncu --set full --clock-control none` captures of k_tile<false> "
                   "(one launch each, tools/profile_kernel.sh capture); bench.py quotes these only while source_sha "
                   "equals the hash of the sources it runs on",
           "source_sha": sha, "kernels": {}}
    for arg in sys.argv[1:]:
        wl, stem = arg.split("=", 1)
        cap_sha = open(stem + ".srcsha").read().strip()
        if cap_sha != sha:
            raise SystemExit(f"{stem}: captured from sources {cap_sha}, the tree is at {sha} — re-capture")
        raw = subprocess.run(["ncu", "-i", stem + ".ncu-rep", "--page", "raw", "--csv"], check=True,
                             capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        names, units, vals = rows[0], rows[1], rows[2]
        col = {n: (u, v) for n, u, v in zip(names, units, vals)}
        rec = {"source": os.path.relpath(stem, ROOT) + ".ncu-rep", "kernel": col["Kernel Name"][1]}
        for key, metric in WANT.items():
            u, v = col[metric]
            x = float(v.replace(",", "")) * SCALE.get(u, 1)
            rec[key] = int(x) if key in ("warp_instructions", "dram_bytes_read", "dram_bytes_write", "registers_per_thread") else round(x, 3)
        out["kernels"][wl] = rec
    with open(bench.ROOFLINE_PROFILE, "w") as f:
        json.dump(out, f, indent=1)
        f.write("\n")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
