"""ctypes binding of libpinn_b200.so (the C ABI declared in include/pinn_b200.h).

The library is built in-tree by csrc/build.sh (nvcc, sm_100a).  There is deliberately no fallback:
if the shared object is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os

MAX_LINEAR, MAX_OUT, MAX_DIRS, NSUMS = 128, 8, 3, 16

ACT = {"tanh": 0, "leaky_relu": 1}
RES_NONE, RES_CONT_ONLY, RES_CONT_FTEMP, RES_NSWE, RES_WAVE_AVG, RES_EXTERNAL, RES_BOUSSINESQ, RES_BOUSS_SIMPLE = range(8)
PREC = {"fp32": 0, "tf32": 1, "tf32x3": 2}
FLAG_ACCUMULATE, FLAG_SKIP_PACK = 1, 2
SUM_FC, SUM_FX, SUM_FY, SUM_COND, SUM_MASKCNT, SUM_TARGET0, SUM_NPOINTS = 0, 1, 2, 3, 4, 5, 13


class Desc(C.Structure):
    _fields_ = [
        ("n_linear", C.c_int32),
        ("widths", C.c_int32 * (MAX_LINEAR + 1)),
        ("activation", C.c_int32),
        ("residual_kind", C.c_int32),
        ("n_dirs", C.c_int32),
        ("dir_cols", C.c_int32 * MAX_DIRS),
        ("field_cols", C.c_int32 * MAX_OUT),
        ("mask_col", C.c_int32),
        ("cond_threshold", C.c_float),
        ("cond_value", C.c_float),
        ("n_targets", C.c_int32),
        ("target_cols", C.c_int32 * MAX_OUT),
        ("target_w", C.c_float * MAX_OUT),
        ("w_fid", C.c_float),
        ("w_res", C.c_float),
        ("precision", C.c_int32),
    ]


class EvalArgs(C.Structure):
    _fields_ = [
        ("params", C.c_void_p),
        ("inputs", C.c_void_p),
        ("targets", C.c_void_p),
        ("n_points", C.c_int64),
        ("n_res_global", C.c_int64),
        ("n_fid_global", C.c_int64),
        ("mask_count", C.c_void_p),
        ("seed_out", C.c_void_p),
        ("seed_dout", C.c_void_p * MAX_DIRS),
        ("grad", C.c_void_p),
        ("sums", C.c_void_p),
        ("out", C.c_void_p),
        ("dout", C.c_void_p * MAX_DIRS),
        ("workspace", C.c_void_p),
        ("workspace_bytes", C.c_size_t),
        ("flags", C.c_int32),
    ]


class LbfgsCfg(C.Structure):
    _fields_ = [("lr", C.c_double), ("tolerance_grad", C.c_double), ("tolerance_change", C.c_double),
                ("max_iter", C.c_int32), ("max_eval", C.c_int32), ("history_size", C.c_int32)]


class LbfgsStatus(C.Structure):
    _fields_ = [("code", C.c_int32), ("n_iter", C.c_int32), ("current_evals", C.c_int32), ("n_iter_total", C.c_int32),
                ("func_evals_total", C.c_int32), ("phase", C.c_int32), ("history_used", C.c_int32), ("pad", C.c_int32),
                ("t", C.c_double), ("loss", C.c_double), ("first_loss", C.c_double), ("gtd", C.c_double),
                ("d_norm", C.c_double)]


# PINN_B200_LIB selects another build of the same library (e.g. the -DPINN_TC_DEBUG cycle-counter build)
LIB_PATH = os.environ.get("PINN_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libpinn_b200.so")

# every symbol include/pinn_b200.h declares: name -> (restype, argtypes)
_P, _I32, _I64, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_float
SYMBOLS = {
    "pinn_version": (C.c_char_p, []),
    "pinn_last_error": (C.c_char_p, []),
    "pinn_param_count": (C.c_int, [C.POINTER(Desc), C.POINTER(C.c_int64)]),
    "pinn_workspace_bytes": (C.c_int, [C.POINTER(Desc), _I64, C.POINTER(C.c_size_t)]),
    "pinn_workspace_bytes_ex": (C.c_int, [C.POINTER(Desc), _I64, _I32, C.POINTER(C.c_size_t)]),
    "pinn_jet_loss_fwd": (C.c_int, [C.POINTER(Desc), C.POINTER(EvalArgs), _P]),
    "pinn_jet_loss_fwdbwd": (C.c_int, [C.POINTER(Desc), C.POINTER(EvalArgs), _P]),
    "pinn_mask_count": (C.c_int, [C.POINTER(Desc), _P, _I64, _P, _P]),
    "pinn_loss_finalize": (C.c_int, [C.POINTER(Desc), _P, _P, _I64, _I64, _P, _P, _P]),
    "pinn_lbfgs_direction": (C.c_int, [_P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _I64, _P, _P]),
    "pinn_lbfgs_workspace_bytes": (C.c_int, [_I64, _I32, C.POINTER(C.c_size_t)]),
    "pinn_lbfgs_begin": (C.c_int, [_P, _I64, C.POINTER(LbfgsCfg), _I32, _P]),
    "pinn_lbfgs_advance": (C.c_int, [_P, _I64, _I32, _P, _P, _P, _P, _P]),
    "pinn_lbfgs_direction_probe": (C.c_int, [_P, _I64, _I32, _P, C.POINTER(C.c_double), _P]),
    "pinn_vec_stats": (C.c_int, [_P, _P, _I64, _P, _P]),
    "pinn_axpy": (C.c_int, [_F, _P, _P, _I64, _P]),
    "pinn_adam_step": (C.c_int, [_P, _P, _P, _P, _I64, _F, _F, _F, _F, _F, _I64, _P]),
    "pinn_nan_minmax": (C.c_int, [_P, _I64, _P, _P]),
    "pinn_assemble_points": (C.c_int, [C.POINTER(C.c_void_p), _I32, _I32, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                       _I32, _I64, _P, _P, _P, _P, _P]),
    "pinn_fma_probe": (C.c_int, [_P, _I32, _I32, C.POINTER(C.c_double), _P]),
}

_lib = None


def lib():
    """The loaded library.  Raises RuntimeError (never falls back) if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with pinn_depthestimation_b200/csrc/build.sh "
                "(or __graft_entry__.build()). There is no CPU fallback for this path.")
        h = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(h, name)   # AttributeError here = header/library mismatch
            fn.restype, fn.argtypes = res, args
        _lib = h
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().pinn_last_error().decode(errors="replace")
        raise RuntimeError(f"pinn_b200 {what} failed (code {rc}): {msg}")


def ptr(t):
    """device pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())
