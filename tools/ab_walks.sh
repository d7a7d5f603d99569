# Developer A/B of the separate walk kernel on one box: flags 8 = walks.cu feeds k_shade, 0 = k_shade walks itself
for f in ${FLAGS:-8 0}; do echo "== PAR_DEBUG_FLAGS=$f"; PAR_DEBUG_FLAGS=$f timeout 300 python tools/probe_gpu.py ${CFGS:-c2 c3 c5} 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print(d['config'], 'build', round(d['ms_grid_build'],3), 'prim', round(d['ms_primary'],3), 'walks', round(d['ms_walks'],3), 'shade', round(d['ms_shade'],3), 'total', round(d['ms_total'],3))
"; done
