#!/bin/bash
for lib in gpurun_dbg/lib_*.so; do
  for c in "wide_nswe 262144" "cmb_h 12514" "txyz 9600" "wide_cont 262144"; do
    set -- $c
    echo -n "$(basename $lib) $1: "
    PINN_B200_LIB=/root/repo/$lib timeout 120 python tools/time_eval.py --cfg $1 --n $2 --precision fp32 --iters 5 2>&1 | tail -1 | cut -c1-110
  done
done
