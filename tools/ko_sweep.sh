#!/bin/bash
# times every variant library under gpurun_dbg/ (knock-out experiments: results of KO_* builds are deliberately wrong)
for lib in gpurun_dbg/lib_*.so; do
  for p in tf32 tf32x3; do
    echo -n "$(basename $lib) $p: "
    PINN_B200_LIB=/root/repo/$lib timeout 120 python tools/time_eval.py --cfg wide_nswe --n 1212416 --precision $p --iters 3 2>&1 | tail -1
  done
done
