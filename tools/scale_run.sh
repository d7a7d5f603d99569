#!/usr/bin/env bash
# Developer helper: a scaling series of bench.py on one box (run under gpurun --gpus 8).
#   tools/scale_run.sh <workload> [N ...]      (default N: 1 2 4 8)
set -u
wl=${1:-c2}
shift || true
ns=${*:-1 2 4 8}
out=gpurun_out/scale_${wl}.jsonl
: > "$out"
for n in $ns; do
  if [ "$n" = 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps ${STEPS:-20} --warmup 5 --workload $wl --no-cpu-baseline 2>gpurun_out/scale_err_$n.log | tail -1 >> "$out"
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) \
      bench.py --gpus $n --steps ${STEPS:-20} --warmup 5 --workload $wl ${EXCHANGE:+--exchange $EXCHANGE} 2>gpurun_out/scale_err_$n.log | tail -1 >> "$out"
  fi
done
python - "$out" <<'PY'
import json, sys
rows = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")]
base = rows[0]["value"] / rows[0]["n_gpus"] if rows else 1
for r in rows:
    print(f"N={r['n_gpus']}  value {r['value']:>10.1f} Mrays/s  {r['ms_per_step']:.3f} ms/step  x{r['value']/base:.2f} (vs first row per GPU)  "
          f"e2e {r['e2e']['value']:.1f} ({r['e2e']['ms_per_step']} ms)  k_shade(rank0) {r['kernels_ms']['k_shade']}")
PY
