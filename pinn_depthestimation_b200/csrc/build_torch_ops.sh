#!/bin/bash
# Builds libpinn_b200_torch.so in-tree: the TORCH_LIBRARY(pinn_b200) shims over libpinn_b200.so (host C++ only).
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
pkg="$here/.."
PY="${PYTHON:-python}"
INC=$($PY -c "from torch.utils.cpp_extension import include_paths; print(' '.join('-I' + p for p in include_paths()))")
LIB=$($PY -c "import os, torch; print(os.path.join(os.path.dirname(torch.__file__), 'lib'))")
ABI=$($PY -c "import torch; print(int(torch._C._GLIBCXX_USE_CXX11_ABI))")
g++ -O2 -std=c++17 -fPIC -shared -D_GLIBCXX_USE_CXX11_ABI=$ABI $INC -I/usr/local/cuda/include \
    "$here/torch_ops.cpp" -o "$pkg/libpinn_b200_torch.so" \
    -L"$pkg" -l:libpinn_b200.so -L"$LIB" -lc10 -lc10_cuda -ltorch_cpu -ltorch -Wl,-rpath,'$ORIGIN' -Wl,-rpath,"$LIB"
echo "built $pkg/libpinn_b200_torch.so"
