#!/usr/bin/env bash
# Developer helper: a scaling series of bench.py on one box (run under gpurun --gpus 8).
#   tools/scale_run.sh <tag> [N ...]      (default N: 1 2 4 8; EXCHANGE=root|peer|nccl, STEPS=..., EXTRA="--no-c4 ...")
set -u
tag=${1:-scale}
shift || true
ns=${*:-1 2 4 8}
out=gpurun_out/${tag}.jsonl
: > "$out"
for n in $ns; do
  if [ "$n" = 1 ]; then
    timeout 400 python bench.py --gpus 1 --steps ${STEPS:-200} --warmup 5 --no-cpu-baseline ${EXTRA:-} 2>gpurun_out/${tag}_err_$n.log | tail -1 >> "$out"
  else
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) \
      bench.py --gpus $n --steps ${STEPS:-200} --warmup 5 ${EXCHANGE:+--exchange $EXCHANGE} ${EXTRA:-} 2>gpurun_out/${tag}_err_$n.log | tail -1 >> "$out"
  fi
done
python - "$out" <<'PY'
import json, sys
rows = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")]
base = {}
for r in rows:
    n = r["n_gpus"]
    k = r["kernels_ms"]
    base.setdefault("c2", r["ms_per_step"] * n)
    print(f"N={n}  C2 {r['ms_per_step']:.4f} ms/step  x{base['c2'] / r['ms_per_step'] / 1:.2f} eff {base['c2'] / r['ms_per_step'] / n * (1 if rows[0]['n_gpus'] == 1 else 1):.2f}  "
          f"loader {k['scene_loader']} k_tile {k['k_tile_min']}..{k['k_tile_max']} step-kernel {k['step_minus_render_kernel_ms']}  "
          f"e2e {r['e2e']['ms_per_step']} ms  check {r['frame_check']['host_frame_equals_oracle']}/{r['frame_check']['device_frames_equal_oracle']}")
    for w, s in (r.get("scale_8k") or {}).items():
        base.setdefault(w, s["ms_per_step"] * n)
        print(f"       {w} {s['ms_per_step']:.4f} ms/step  x{base[w] / s['ms_per_step']:.2f} eff {base[w] / s['ms_per_step'] / n:.2f}  "
              f"k_tile {s['render_kernel_ms_min']}..{s['render_kernel_ms_max']}  check {s['frame_check']['device_frames_equal_oracle']}")
PY
