#!/usr/bin/env bash
# Developer A/B on one box: alternate the library variants built by tools/ab_build.sh (plus "main" = the
# in-tree library) over the probe workloads, REPS times each, and print render-kernel / step times.
#   tools/ab_run.sh "base gmap" "c2 c3" 3
variants=${1:?variants}; cfgs=${2:-c2 c3}; reps=${3:-3}
for r in $(seq $reps); do
  for v in $variants; do
    lib=pixel-art-raytracer_b200/build/variants/$v/libpar_b200.so
    [ "$v" = main ] && lib=pixel-art-raytracer_b200/par_b200/libpar_b200.so
    PAR_B200_LIB=$lib timeout 300 python tools/probe_gpu.py $cfgs 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print('$v', d['config'][:3], 'render', round(d['ms_render'],4), 'step', d['resident_step_ms'])
"
  done
done
