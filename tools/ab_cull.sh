# Developer A/B of the shaft cull on one box: flags 0 = measured bounds, 2 = analytic bounds, 1 = no cull
for f in ${FLAGS:-0 2 1}; do echo "== PAR_DEBUG_FLAGS=$f"; PAR_DEBUG_FLAGS=$f PAR_PHASES=1 timeout 300 python tools/probe_gpu.py ${CFGS:-c2 c3} 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print(d['config'], 'prim', round(d['ms_primary'],3), 'shade', round(d['ms_shade'],3), d.get('phases_pct'), d.get('lists'))
"; done
