#!/usr/bin/env bash
# Developer helper: the 1/2/4/8-GPU scaling series of bench.py on one box (run under gpurun --gpus 8).
set -u
out=gpurun_out/scale_${1:-c2}.jsonl
: > "$out"
for n in 1 2 4 8; do
  if [ "$n" = 1 ]; then
    python bench.py --gpus 1 --steps ${STEPS:-20} --warmup 5 --workload ${1:-c2} --no-cpu-baseline 2>gpurun_out/scale_err.log | tail -1 >> "$out"
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) \
      bench.py --gpus $n --steps ${STEPS:-20} --warmup 5 --workload ${1:-c2} 2>gpurun_out/scale_err.log | tail -1 >> "$out"
  fi
done
python - "$out" <<'PY'
import json, sys
rows = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")]
base = rows[0]["value"] if rows else 1
for r in rows:
    print(f"N={r['n_gpus']}  value {r['value']:>10.1f} Mrays/s  {r['ms_per_step']:.3f} ms/step  speedup {r['value']/base:.2f}  "
          f"eff {r['value']/base/r['n_gpus']:.2f}  e2e {r['e2e']['value']:.1f}  k_shade(rank0) {r['kernels_ms']['k_shade']}")
PY
