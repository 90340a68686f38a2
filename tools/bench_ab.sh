#!/bin/bash
# long-run (power-capped) A/B of variant libraries on the bench workload: headline mode and the TF32 mode
for rep in 1 2; do
for lib in gpurun_dbg/lib_*.so; do
  PINN_B200_LIB=/root/repo/$lib python bench.py --steps 5 --warmup 3 --other-modes tf32 --other-steps 5 --no-cpu-baseline --no-eager-baseline --lbfgs-iters 0 --real-shapes 0 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('$(basename $lib)', 'x3', round(d['value']), d['clocks']['sm_mhz'], 'tf32', round(d['modes']['tf32']['value']), d['modes']['tf32']['clocks']['sm_mhz'])
"
done
done
