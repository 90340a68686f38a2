#!/bin/bash
# Builds libpinn_b200.so in-tree for sm_100a (cross-compiles without a GPU).
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
out="${PINN_OUT:-$here/../libpinn_b200.so}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 ${PINN_EXTRA_FLAGS:-}
       -Xcompiler -fPIC -Xptxas -v --threads 4)
"$NVCC" "${FLAGS[@]}" -shared -o "$out" "$here/cabi.cu" "$here/jet_fp32.cu" "$here/jet_tc.cu" "$here/jet3.cu" "$here/optim.cu" "$here/lbfgs_dev.cu" "$here/data.cu" "$here/probe.cu" "$@" \
  2> "$here/../ptxas.log" || { cat "$here/../ptxas.log" >&2; exit 1; }
echo "built $out"
