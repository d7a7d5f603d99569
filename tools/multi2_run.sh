#!/usr/bin/env bash
# Developer helper for a 2-GPU box (gpurun --gpus 2): the multi-GPU parity tests, then bench.py at N = 2 as the
# driver runs it, with the one-light kernel configuration pinned on (PAR_DEBUG_FLAGS=512; the automatic choice
# takes the general one at 2592 tiles per rank) and with a forced stripe_split of 2.
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02_n2_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r02_n2_pytest.log
STEPS=100 EXTRA="--no-c4 --steps-8k 5" tools/scale_run.sh r02_n2_auto 2
PAR_DEBUG_FLAGS=512 STEPS=100 EXTRA="--no-c4 --no-scale-8k" tools/scale_run.sh r02_n2_one_light 2
PAR_BENCH_STRIPE_SPLIT=2 STEPS=100 EXTRA="--no-c4 --steps-8k 5" tools/scale_run.sh r02_n2_split2 2
tail -3 gpurun_out/r02_n2_*err_2.log
