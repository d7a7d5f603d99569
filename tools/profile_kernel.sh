#!/usr/bin/env bash
# Developer helper: one ncu --set full capture of a kernel on a workload, then the per-line and
# per-phase tables.  Step 1 runs on the GPU box (under gpurun, ONE GPU, only after the same probe
# has exited 0 without ncu); step 2 runs anywhere (no GPU needed), against the SAME build.
#
#   gpurun --timeout 600 -- 'tools/profile_kernel.sh capture c2 k_tile r02_tile_c2'
#   tools/profile_kernel.sh report r02_tile_c3 tile k_tile k_tileILb0     # -> profiles/<tag>_lines.txt, _metrics.txt
#   tools/profile_kernel.sh report r02_tile_c2 tile_one_light k_tile k_tile_one_lightILb0   (one-light frames run that build)
set -euo pipefail
cmd=${1:?capture|report}
case "$cmd" in
capture)
    wl=${2:?workload}; kernel=${3:?kernel name}; tag=${4:?tag}
    python tools/probe_gpu.py "$wl" > "gpurun_out/${tag}_probe.log" 2>&1   # must pass on its own first
    python -c "import bench; print(bench.kernel_source_sha())" > "gpurun_out/${tag}.srcsha"   # keys the capture (bench.py)
    ncu --set full --clock-control none --import-source on -k "regex:${kernel}" -c 1 -f \
        -o "gpurun_out/${tag}" python tools/probe_gpu.py "$wl" > "gpurun_out/${tag}_ncu.log" 2>&1
    ls -la "gpurun_out/${tag}.ncu-rep"
    ;;
report)
    tag=${2:?tag}; stem=${3:?cubin stem, e.g. tile}; kernel=${4:?kernel name}   # [5: mangled-name substring, e.g. k_tileILb0]
    rep="gpurun_out/${tag}.ncu-rep"
    mkdir -p profiles
    {
        echo "# ${tag} — ${kernel}, ncu --set full"
        python tools/ncu_phases.py "$rep" "${5:-$kernel}" "$stem" 2>/dev/null || true
        echo
        python tools/ncu_lines.py "$rep" "$stem" "$kernel" --symbol "${5:-$kernel}" --top 40 --by inst
    } > "profiles/${tag}_lines.txt"
    ncu -i "$rep" --page raw --csv -k "regex:${kernel}" > "gpurun_out/${tag}_raw.csv" 2>/dev/null
    python - "$tag" <<'PY'
import csv, sys
rows = list(csv.reader(open(f"gpurun_out/{sys.argv[1]}_raw.csv")))
want = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu", "sm__inst_executed_pipe_fma", "sm__inst_executed_pipe_lsu",
        "smsp__average_warps_issue_stalled", "launch__registers_per_thread", "launch__occupancy_limit")
if len(rows) >= 3:
    names, units, vals = rows[0], rows[1], rows[2]
    with open(f"profiles/{sys.argv[1]}_metrics.txt", "w") as f:
        for n, u, v in zip(names, units, vals):
            if any(n.startswith(w) for w in want):
                f.write(f"{n} [{u}] = {v}\n")
PY
    echo "wrote profiles/${tag}_lines.txt profiles/${tag}_metrics.txt"
    ;;
*) echo "usage: $0 capture <workload> <kernel> <tag> | report <tag> <cubin stem> <kernel>"; exit 2 ;;
esac
