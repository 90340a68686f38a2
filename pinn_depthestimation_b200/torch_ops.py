"""`torch.ops.pinn_b200.*`: the C ABI registered as torch custom ops (csrc/torch_ops.cpp, TORCH_LIBRARY(pinn_b200)).

    from pinn_depthestimation_b200 import torch_ops
    ops = torch_ops.load()                      # torch.ops.pinn_b200; RuntimeError if the library is not built
    sums = ops.jet_loss(torch_ops.desc_tensor(spec), params, inputs, targets, None, n, n, workspace, grad, None, 0)

The ops take the CURRENT torch CUDA stream, call the same extern "C" entry points that `_cabi` binds with ctypes, and
raise RuntimeError with pinn_last_error() on failure.  The Python facades keep using the ctypes binding (no compiled
dependency on a torch version); this library is the binding for callers that live in torch's dispatcher (C++ extensions,
TorchScript-free C++ drivers, custom autograd code).  No CPU fallback: CPU tensors are rejected by every op.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libpinn_b200_torch.so")
_loaded = False


def load():
    global _loaded
    if not _loaded:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with pinn_depthestimation_b200/csrc/build_torch_ops.sh "
                               "(or __graft_entry__.build())")
        torch.ops.load_library(LIB_PATH)
        _loaded = True
    return torch.ops.pinn_b200


def desc_tensor(spec) -> torch.Tensor:
    """PassSpec -> the CPU uint8 tensor holding the bytes of its pinn_desc_t."""
    d = spec.to_desc()
    buf = (C.c_uint8 * C.sizeof(d)).from_buffer_copy(d)
    return torch.frombuffer(bytearray(buf), dtype=torch.uint8).clone()
