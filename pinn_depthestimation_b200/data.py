"""The reference's one-off data preparation, done on the device.

What the reference's __main__ blocks do with numpy before anything reaches the GPU -- per-variable min/max
normalisation to [-1,1] (operations.py:4-30), hstack of the variable columns to `[N,d]`, removal of the rows that hold
a NaN (train_newmethod.py:226-255; train.py:203-276) -- as ONE C-ABI call (`pinn_assemble_points`: normalise + NaN
filter + order-preserving compaction) plus `pinn_nan_minmax` for the data-derived ranges.  The raw columns are
uploaded once; the `[N,d]` / `[N,n_true]` tensors the trainer keeps resident are produced where they will live.

    inputs, trues = assemble_points({'x': X, 'y': Y}, {'U': U, 'V': V}, config)          # train_newmethod form
    inputs, _     = assemble_points({'x': X, 'y': Y}, {}, config, drop_nan_inputs=True)  # train.py residual grid
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Mapping, Optional, Tuple

import numpy as np
import torch

from . import _cabi


def _col(a, device):
    t = a if isinstance(a, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(a))
    return t.to(device=device, dtype=torch.float32).reshape(-1).contiguous()


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def min_max(column: torch.Tensor, key: str, config) -> Tuple[float, float]:
    """operations.get_min_max (operations.py:16-30): x / y ranges come from config['data_test'], every other
    variable from the data itself, NaNs ignored -- computed on the device."""
    if key == 'x':
        return float(config['data_test']['x_min']), float(config['data_test']['x_max'])
    if key == 'y':
        return float(config['data_test']['y_min']), float(config['data_test']['y_max'])
    out = torch.empty(2, dtype=torch.float32, device=column.device)
    with torch.cuda.device(column.device):
        _cabi.check(_cabi.lib().pinn_nan_minmax(_cabi.ptr(column), column.numel(), _cabi.ptr(out),
                                                _stream(column.device)), "pinn_nan_minmax")
    lo, hi = out.tolist()
    return float(lo), float(hi)


def assemble_points(input_columns: Mapping[str, object], true_columns: Mapping[str, object], config,
                    ranges: Optional[Dict[str, Tuple[float, float]]] = None, drop_nan_trues: bool = True,
                    drop_nan_inputs: bool = False, device="cuda"):
    """-> (inputs [N_kept, d] normalised, trues [N_kept, n_true] or None, ranges).  Column order = mapping order
    (the reference iterates config['data']['inputs'] / ['trues'] in file order)."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("assemble_points runs on the GPU: this path has no CPU fallback")
    names_in, names_true = list(input_columns), list(true_columns)
    cols = [_col(input_columns[k], dev) for k in names_in] + [_col(true_columns[k], dev) for k in names_true]
    n = cols[0].numel()
    if any(c.numel() != n for c in cols):
        raise ValueError("all columns must have the same number of rows")
    ranges = dict(ranges or {})
    for k, c in zip(names_in, cols):
        if k not in ranges:
            ranges[k] = min_max(c, k, config)
    lo = (C.c_float * len(names_in))(*[ranges[k][0] for k in names_in])
    hi = (C.c_float * len(names_in))(*[ranges[k][1] for k in names_in])
    ptrs = (C.c_void_p * len(cols))(*[c.data_ptr() for c in cols])
    inputs = torch.empty(n, len(names_in), dtype=torch.float32, device=dev)
    trues = torch.empty(n, max(1, len(names_true)), dtype=torch.float32, device=dev)
    kept = torch.zeros(1, dtype=torch.int64, device=dev)
    scratch = torch.empty((n + 255) // 256 + 1, dtype=torch.int32, device=dev)
    policy = (1 if (drop_nan_trues and names_true) else 0) | (2 if drop_nan_inputs else 0)
    with torch.cuda.device(dev):
        _cabi.check(_cabi.lib().pinn_assemble_points(
            ptrs, len(names_in), len(names_true), lo, hi, policy, n, _cabi.ptr(inputs), _cabi.ptr(trues),
            _cabi.ptr(kept), _cabi.ptr(scratch), _stream(dev)), "pinn_assemble_points")
    k = int(kept.item())          # one host sync, once per run
    return inputs[:k], (trues[:k, :len(names_true)].contiguous() if names_true else None), ranges
