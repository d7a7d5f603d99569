#!/usr/bin/env bash
# Developer helper: everything a round's final 1-GPU gpurun call collects, most important first, every
# step under its own timeout.   gpurun --timeout 1500 -- 'tools/final_run.sh r02_final'
#   1. the GPU parity suite   2. smoke()   3. the bench line   4. one ncu --set full capture of k_tile per
#   workload (c2, c3, c5b; each only after the same probe exited 0 without ncu)   5. the ncu launch list
tag=${1:-final}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py > gpurun_out/${tag}_bench_line.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
head -c 600 gpurun_out/${tag}_bench_line.json; echo
timeout 200 python tests/fuzz_parity.py ${FUZZ:-150} 5000 > gpurun_out/${tag}_fuzz.log 2>&1; echo "fuzz rc=$?"; tail -2 gpurun_out/${tag}_fuzz.log
for wl in c2 c3 c5b; do
    timeout 400 tools/profile_kernel.sh capture $wl k_tile ${tag}_tile_$wl > gpurun_out/${tag}_capture_$wl.log 2>&1; echo "capture $wl rc=$?"
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_bench_c2_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-scale-8k --no-c4 > gpurun_out/${tag}_launches_ncu.log 2>&1; echo "launch list rc=$?"
ls -la gpurun_out | tail -30
