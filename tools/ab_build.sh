# Developer A/B of compile-time variants on one GPU box: builds the library twice and probes each.
#   VARIANTS="-DPAR_TRIM=0 -DPAR_TRIM=1" CFGS="c2 c3" tools/ab_build.sh
for v in ${VARIANTS}; do
  touch pixel-art-raytracer_b200/csrc/shade.cu
  flags="${v//,/ }"   # commas separate several flags of one variant
  PAR_NVCC_EXTRA="$flags" pixel-art-raytracer_b200/build_native.sh > /tmp/ab_build.log 2>&1 || { echo "BUILD FAILED for $flags"; tail -5 /tmp/ab_build.log; continue; }
  echo "== $flags"
  python tools/probe_gpu.py ${CFGS:-c2 c3 c5} | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['config'], 'prim', round(d['ms_primary'],3), 'shade', round(d['ms_shade'],3), 'total', round(d['ms_total'],3))
"
done
