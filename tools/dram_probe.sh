#!/bin/bash
# DRAM bytes of one launch per variant library (ncu, two metrics only)
for lib in gpurun_dbg/lib_*.so; do
  for p in tf32 tf32x3; do
    PINN_B200_LIB=/root/repo/$lib python tools/time_eval.py --cfg wide_nswe --n 1048576 --precision $p --iters 1 > /dev/null 2>&1 &&
    PINN_B200_LIB=/root/repo/$lib ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:jet_ -s 2 -c 1 python tools/time_eval.py --cfg wide_nswe --n 1048576 --precision $p --iters 1 2>&1 | grep -E "dram__|gpu__time|hit_rate" | sed "s|^|$(basename $lib) $p |"
  done
done
