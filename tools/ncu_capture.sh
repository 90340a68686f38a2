set -e
for p in tf32x3 tf32; do python tools/time_eval.py --cfg wide_nswe --n 1048576 --precision $p --iters 1 > gpurun_out/r2e_${p}_plain.log 2>&1; done
for p in tf32x3 tf32; do
  ncu --set full --clock-control none --import-source on -k regex:jet_ -s 2 -c 1 -o gpurun_out/r2e_${p}_prof -f python tools/time_eval.py --cfg wide_nswe --n 1048576 --precision $p --iters 1 > gpurun_out/r2e_${p}_ncu.log 2>&1
done
