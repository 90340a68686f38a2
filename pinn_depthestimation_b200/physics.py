"""Drop-in for the reference's physics.py (physics.py:6-120): same function names and argument
order, each returning a 0-d differentiable loss tensor on the inputs' device.

    compute_gradient(pred, var)
    continuity_only(x, y, h, U, V)            continuity_ftemp(x, y, h, U, V)
    Navier_Stokes(t, x, y, h, z, u, v)        physics_equation(x, y, h, U, V, eta_mean, Hrms, k)

The residual functions receive [N,1] *views* of a DNN output (`predictions[:, i:i+1]`) and the
individual input columns, exactly as pinn.loss_func passes them (train_newmethod.py:136-156,
train.py:144-154).  They recognise which `pinn_depthestimation_b200.dnn.DNN.forward` call produced
the views, and run ONE fused kernel (jet forward + residual + reverse sweep, pinn_jet_loss_fwdbwd)
instead of the reference's one autograd sweep per compute_gradient call.  The returned scalar
carries a custom backward that hands the already-computed weight gradient to autograd, so
`loss.backward()` in the caller keeps working unchanged.

If the arguments cannot be traced to a DNN.forward call (e.g. the user did arithmetic on them
first) a RuntimeError explains why -- nothing silently falls back to another implementation.
`compute_gradient` itself is the reference's autograd call; the DNN facade makes it work by
answering d out / d x with forward jets.
"""
from __future__ import annotations

import os

import torch

from . import dnn as _dnn
from .fused import JetLoss
from .spec import DIR_ORDER, FIELD_ORDER, PassSpec

_CACHE = {}


def compute_gradient(pred, var):
    """physics.py:6-15 -- per-point d pred / d var, differentiable."""
    return torch.autograd.grad(pred, var, grad_outputs=torch.ones_like(pred),
                               retain_graph=True, create_graph=True)[0]


def _column_of_output(t, out):
    o = out.shape[1]
    if t.dim() != 2 or t.shape != (out.shape[0], 1) or t.stride(0) != o:
        raise RuntimeError("residual arguments must be [N,1] column views of the DNN output "
                           "(predictions[:, i:i+1])")
    return (t.storage_offset() - out.storage_offset()) % o


def _column_of_input(v, inputs):
    """Which column of the [N,d] tensor fed to DNN.forward is `v`?"""
    n, d = inputs.shape
    if v.shape not in ((n, 1), (n,)):
        raise RuntimeError("input-column argument has the wrong shape")
    # (1) a view of the input tensor
    if v.untyped_storage().data_ptr() == inputs.untyped_storage().data_ptr():
        return (v.storage_offset() - inputs.storage_offset()) % d
    # (2) one of the tensors torch.cat-ed into it (train_newmethod.py:123-124)
    fn = inputs.grad_fn
    if fn is not None and fn.name().startswith("CatBackward"):
        for c, (nxt, _) in enumerate(fn.next_functions):
            if nxt is None:
                continue
            if v.grad_fn is not None and nxt is v.grad_fn:
                return c
            if v.grad_fn is None and getattr(nxt, "variable", None) is v:
                return c
    # (3) compare values (one host sync; cached per tensor version)
    key = ("col", v.data_ptr(), v._version, inputs.data_ptr(), inputs._version)
    if key not in _CACHE:
        if len(_CACHE) > 64:
            _CACHE.clear()
        hit = [c for c in range(d) if torch.equal(v.reshape(-1).to(inputs.dtype), inputs[:, c])]
        if len(hit) != 1:
            raise RuntimeError("cannot tell which input column this tensor is")
        _CACHE[key] = hit[0]
    return _CACHE[key]


class _ResidualFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, jl, *params):
        flat = module.flat_params()
        grad = torch.empty_like(flat)
        parts = jl.loss_and_grad(flat, grad)
        ctx.module, ctx.grad = module, grad
        return parts[1].clone()

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        return (None, None, *ctx.module.split_flat(ctx.grad * g))


def _fused(kind, dir_args, field_args):
    names_d, names_f = DIR_ORDER[kind], FIELD_ORDER[kind]
    if len(field_args) == 1 and field_args[0].dim() == 2 and field_args[0].shape[1] >= len(names_f):
        # physics_functions.* receive the whole DNN output and slice the columns 0..3 themselves (output[:, i:i+1])
        out_t = field_args[0]
        field_args = tuple(out_t[:, i:i + 1] for i in range(len(names_f)))
    src = _dnn.provenance(field_args[0])
    if src is None:
        raise RuntimeError(
            f"{kind}: the prediction columns do not come from a pinn_depthestimation_b200 DNN.forward "
            "call that is still alive; the fused residual needs `predictions[:, i:i+1]` views")
    module, inputs, out = src
    fields = {n: _column_of_output(t, out) for n, t in zip(names_f, field_args)}
    dirs = {n: _column_of_input(t, inputs) for n, t in zip(names_d, dir_args)}
    if len(set(fields.values())) != len(fields) or len(set(dirs.values())) != len(dirs):
        raise RuntimeError(f"{kind}: arguments must be distinct columns")
    xin = inputs.detach()
    if xin.dtype != torch.float32 or not xin.is_contiguous():
        xin = xin.to(torch.float32).contiguous()
    # one fused pass per (module, residual, column mapping, N, device), kept on the module; the reference's loss_func
    # builds a new `torch.cat` of its input columns on every call, so the pass is re-pointed, not rebuilt
    prec = os.environ.get("PINN_B200_PRECISION", "fp32")   # tf32 / tf32x3: 256-wide nets on the tensor cores
    key = (kind, tuple(sorted(fields.items())), tuple(sorted(dirs.items())), tuple(xin.shape), str(xin.device), prec)
    cache = module._runner.jet_losses
    jl = cache.get(key)
    if jl is None:
        while len(cache) >= 4:
            cache.popitem(last=False)
        spec = PassSpec(layers=module.layer_sizes, activation=module.activation_name, kind=kind,
                        dirs=dirs, fields=fields, w_fid=0.0, w_res=1.0, precision=prec)
        jl = JetLoss(spec, xin, None)
        cache[key] = jl
    else:
        cache.move_to_end(key)
        jl.res.rebind(xin)
    return _ResidualFunction.apply(module, jl, *module.parameters())


def continuity_only(x, y, h, U, V):
    """physics.py:18-33."""
    return _fused("continuity_only", (x, y), (h, U, V))


def continuity_ftemp(x, y, h, U, V):
    """physics.py:37-47."""
    return _fused("continuity_ftemp", (x, y), (h, U, V))


def Navier_Stokes(t, x, y, h, z, u, v):
    """physics.py:50-88."""
    return _fused("Navier_Stokes", (t, x, y), (h, z, u, v))


def physics_equation(x, y, h, U, V, eta_mean, Hrms, k):
    """physics.py:91-120."""
    return _fused("physics_equation", (x, y), (h, U, V, eta_mean, Hrms, k))
