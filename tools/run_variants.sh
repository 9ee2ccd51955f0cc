#!/bin/bash
# On the GPU box: time every variant with the short bench.
cd "$(dirname "$0")/.."
for lib in visrtx_b200/variants/libdvr_*.so; do
  echo "== $lib"
  DVR_B200_LIB=$PWD/$lib python bench.py --steps 40 --warmup 5 --extra ${EXTRA:-0} --no-cpu-baseline "$@" 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('   fps %.1f  kernel_ms %.4f  frac %.3f  e2e %.1f'%(d['value'],r['kernel_ms'],r['frac'],d['e2e']['value']))
        v=d['extra'].get('variants')
        if v:
            for k,x in v.items(): print('      %s fps %.1f gsamples/s %.1f'%(k,x['fps'],x['gsamples_per_s']))
    elif 'Error' in l or 'error' in l: print(l.strip())
"
done
