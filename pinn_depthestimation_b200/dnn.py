"""Drop-in for the reference's dnn.py (dnn.py:5-55): same class name, constructor signature,
submodule names (`layers.layer_{i}`, `activation_{i}`, `dropout_{i}`), parameter order, state_dict
keys and pickling behaviour -- but forward/backward run on the sm_100a kernels through the C ABI.

    DNN(layers: list[int], dropout_rate: float, init_type: 'xavier' | 'kaiming')

* `forward(x)` launches the value-only fused forward (pinn_jet_loss_fwd, residual kind NONE).
* its autograd backward launches the fused forward+reverse with caller-supplied output seeds
  (PINN_RES_EXTERNAL) and hands each nn.Parameter its slice of the flat gradient; that node is separate
  from the one that answers d out / d x, so input-only derivatives never pay for a weight gradient.
* d out / d x (what physics.compute_gradient asks autograd for) is answered with forward jets, and
  is itself differentiable w.r.t. the weights (again through PINN_RES_EXTERNAL), so the reference's
  unmodified physics.py also runs on top of this module (3 input directions per jet launch).
* CPU tensors are rejected: there is no CPU fallback.

Parameters are real nn.Parameters whose storage is a view into one flat fp32 buffer in
nn.Module.parameters() order, which is exactly the flat vector the C ABI and the L-BFGS kernels use.
"""
from __future__ import annotations

import ctypes as C
import weakref
from collections import OrderedDict

import torch
import torch.nn as nn

from . import _cabi
from .spec import PassSpec

_RECENT = OrderedDict()   # output storage ptr -> (module, inputs, out)  (provenance for physics.*)
_RECENT_MAX = 8


def _remember(out, module, inputs):
    # weak references only: the registry must not keep [N,o] outputs alive
    _RECENT[out.untyped_storage().data_ptr()] = (weakref.ref(module), weakref.ref(inputs),
                                                 weakref.ref(out))
    while len(_RECENT) > _RECENT_MAX:
        _RECENT.popitem(last=False)


def provenance(t):
    """(module, inputs, out) of the DNN.forward call that produced the storage `t` views, or None."""
    rec = _RECENT.get(t.untyped_storage().data_ptr())
    if rec is None:
        return None
    module, inputs, out = (r() for r in rec)
    if module is None or inputs is None or out is None:
        return None
    return module, inputs, out


class _Runner:
    """One descriptor + workspace per (kind, directions, point count, device, forward-only?) of one module.  The key
    does NOT contain the input pointer: the reference's loss_func feeds a freshly `torch.cat`-ed tensor on every call
    (train_newmethod.py:123-124), so the pass is re-pointed at the new storage instead of being rebuilt."""

    def __init__(self, module):
        self.module = module
        self.cache = OrderedDict()
        self.jet_losses = OrderedDict()   # physics.py facade: fused residual passes of this module

    def get(self, kind, ext_dirs, inputs, want_grad=True):
        from .fused import _Pass
        key = (kind, tuple(ext_dirs), tuple(inputs.shape), str(inputs.device), bool(want_grad))
        p = self.cache.get(key)
        if p is None:
            while len(self.cache) >= 12:
                self.cache.popitem(last=False)
            spec = PassSpec(layers=self.module.layer_sizes, activation=self.module.activation_name,
                            kind=kind, ext_dirs=list(ext_dirs))
            p = _Pass(spec, inputs, None, want_grad=want_grad)
            self.cache[key] = p
        else:
            self.cache.move_to_end(key)
            p.rebind(inputs)
        return p


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _check_input(module, x):
    if not x.is_cuda:
        raise RuntimeError("pinn_b200 DNN.forward needs CUDA tensors: this path has no CPU fallback")
    if x.dim() != 2 or x.shape[1] != module.layer_sizes[0]:
        raise ValueError(f"expected input [N,{module.layer_sizes[0]}], got {tuple(x.shape)}")
    if module.training and module.dropout_rate != 0.0:
        raise NotImplementedError(
            "dropout_rate != 0 in training mode is not supported by the fused path "
            "(every reference config uses 0.0; see SURVEY.md 3.3)")


class _JetFunction(torch.autograd.Function):
    """(out, d out/d x_c for c in dirs) with a weight gradient from external seeds (<= 3 dirs)."""

    @staticmethod
    def forward(ctx, module, x, dirs, *params):
        o = module.layer_sizes[-1]
        xin = x.detach().to(torch.float32).contiguous()
        flat = module.flat_params()
        ps = module._runner.get("external", dirs, xin, want_grad=False)
        n = xin.shape[0]
        out = torch.empty(n, o, device=xin.device)
        douts = [torch.empty(n, o, device=xin.device) for _ in dirs]
        a = ps.args(flat, None, 1, 1, 0, out=out, douts=douts)
        with torch.cuda.device(xin.device):
            _cabi.check(_cabi.lib().pinn_jet_loss_fwd(C.byref(ps.desc), C.byref(a), _stream(xin.device)),
                        "pinn_jet_loss_fwd")
        ctx.module, ctx.xin, ctx.dirs = module, xin, dirs
        ctx.set_materialize_grads(False)
        return (out, *douts)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_out, *g_douts):
        module, xin, dirs = ctx.module, ctx.xin, ctx.dirs
        flat = module.flat_params()
        ps = module._runner.get("external", dirs, xin)
        grad = torch.empty_like(flat)
        f32 = lambda g: None if g is None else g.to(torch.float32).contiguous()
        a = ps.args(flat, grad, 1, 1, 0, seed_out=f32(g_out), seed_douts=[f32(g) for g in g_douts])
        with torch.cuda.device(xin.device):
            _cabi.check(_cabi.lib().pinn_jet_loss_fwdbwd(C.byref(ps.desc), C.byref(a), _stream(xin.device)),
                        "pinn_jet_loss_fwdbwd")
        return (None, None, None, *module.split_flat(grad))


class _MLPValue(torch.autograd.Function):
    """Value-only forward; its backward answers d out / d x with forward jets (differentiable w.r.t. the weights, so
    physics.compute_gradient's create_graph=True works).  The weights are deliberately NOT inputs of this node: their
    gradient hangs on the separate _MLPParamGrad node, so an `autograd.grad(pred, var)` w.r.t. the inputs only
    (physics.py:6-15; 13 of them per Navier_Stokes call) never launches a weight-gradient sweep it would throw away."""

    @staticmethod
    def forward(ctx, module, x):
        xin = x.detach().to(torch.float32).contiguous()
        flat = module.flat_params()
        ps = module._runner.get("none", (), xin, want_grad=False)
        out = torch.empty(xin.shape[0], module.layer_sizes[-1], device=xin.device)
        a = ps.args(flat, None, 1, 1, 0, out=out)
        with torch.cuda.device(xin.device):
            _cabi.check(_cabi.lib().pinn_jet_loss_fwd(C.byref(ps.desc), C.byref(a), _stream(xin.device)),
                        "pinn_jet_loss_fwd")
        ctx.module, ctx.x = module, x
        return out

    @staticmethod
    def backward(ctx, g_out):
        module, x = ctx.module, ctx.x
        if not ctx.needs_input_grad[1]:
            return None, None
        # vjp w.r.t. the inputs = sum_n g_out[:,n] * d out_n/d x_c, written with differentiable torch ops on the jet
        # outputs so that create_graph=True works
        d = module.layer_sizes[0]
        params = tuple(module.parameters())
        with torch.enable_grad():
            cols = []
            for c0 in range(0, d, _cabi.MAX_DIRS):      # <= 3 jet directions per launch
                dirs = tuple(range(c0, min(d, c0 + _cabi.MAX_DIRS)))
                jets = _JetFunction.apply(module, x, dirs, *params)
                cols += [(g_out * dj).sum(dim=1) for dj in jets[1:]]
            gx = torch.stack(cols, dim=1).to(x.dtype)
        return None, gx


class _MLPParamGrad(torch.autograd.Function):
    """Zero-valued [N,o] tensor added to the network output; its backward is the weight gradient of the value path
    (fused forward + reverse sweep with the caller's d loss / d out as external seeds)."""

    @staticmethod
    def forward(ctx, module, x, *params):
        ctx.module = module
        ctx.xin = x.detach().to(torch.float32).contiguous()
        return ctx.xin.new_zeros(ctx.xin.shape[0], module.layer_sizes[-1])

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_out):
        module, xin = ctx.module, ctx.xin
        flat = module.flat_params()
        ps = module._runner.get("external", (), xin)
        grad = torch.empty_like(flat)
        a = ps.args(flat, grad, 1, 1, 0, seed_out=g_out.detach().to(torch.float32).contiguous())
        with torch.cuda.device(xin.device):
            _cabi.check(_cabi.lib().pinn_jet_loss_fwdbwd(C.byref(ps.desc), C.byref(a),
                                                         _stream(xin.device)), "pinn_jet_loss_fwdbwd")
        return (None, None, *module.split_flat(grad))


class DNN(nn.Module):

    def __init__(self, layers, dropout_rate, init_type):
        super().__init__()
        if init_type == 'xavier':
            self.activation = nn.Tanh()
        elif init_type == 'kaiming':
            self.activation = nn.LeakyReLU(negative_slope=0.01)
        else:
            raise ValueError(f"Invalid init_type: {init_type}. Use 'kaiming' or 'xavier'.")
        self.layer_sizes = [int(v) for v in layers]
        self.dropout_rate = float(dropout_rate)
        self.init_type = init_type
        mods = []
        n_lin = len(self.layer_sizes) - 1
        for i in range(n_lin):
            lin = nn.Linear(self.layer_sizes[i], self.layer_sizes[i + 1])
            if init_type == 'kaiming':
                nn.init.kaiming_uniform_(lin.weight, nonlinearity='leaky_relu')
            else:
                nn.init.xavier_uniform_(lin.weight)
            if i < n_lin - 1:            # the last bias keeps nn.Linear's default init (dnn.py:33)
                nn.init.zeros_(lin.bias)
            mods.append((f'layer_{i}', lin))
            if i < n_lin - 1:
                mods.append((f'activation_{i}', self.activation))
                mods.append((f'dropout_{i}', nn.Dropout(dropout_rate)))
        self.layers = nn.Sequential(OrderedDict(mods))
        self._flat = None
        self._runner = _Runner(self)

    # -- pickling (torch.save(model.dnn) in train_newmethod.py:184,270): drop device caches ------
    def __getstate__(self):
        st = self.__dict__.copy()
        st["_flat"] = None
        st["_runner"] = None
        return st

    def __setstate__(self, st):
        self.__dict__.update(st)
        # whole-module pickles written by the reference's own dnn.DNN carry only nn.Module state:
        # recover the constructor facts from the submodules
        lins = [m for m in self.layers if isinstance(m, nn.Linear)]
        if "layer_sizes" not in st:
            self.layer_sizes = [lins[0].in_features] + [m.out_features for m in lins]
        if "init_type" not in st:
            self.init_type = 'kaiming' if isinstance(self.activation, nn.LeakyReLU) else 'xavier'
        if "dropout_rate" not in st:
            drops = [m.p for m in self.layers if isinstance(m, nn.Dropout)]
            self.dropout_rate = float(drops[0]) if drops else 0.0
        self._flat = None
        self._runner = _Runner(self)

    @property
    def activation_name(self):
        return "tanh" if self.init_type == 'xavier' else "leaky_relu"

    # -- flat parameter vector -------------------------------------------------------------------
    def flat_params(self) -> torch.Tensor:
        """The flat fp32 vector [P] backing every parameter (re-established after .to()/load)."""
        ps = list(self.parameters())
        flat = self._flat
        ok = flat is not None and flat.device == ps[0].device
        if ok:
            o, base = 0, flat.data_ptr()
            for p in ps:
                if p.data_ptr() != base + 4 * o or p.dtype != torch.float32 or not p.is_contiguous():
                    ok = False
                    break
                o += p.numel()
        if not ok:
            # adopts the buffer when an optimiser has already flattened the parameters (lbfgs.flatten_params is idempotent)
            from .lbfgs import flatten_params
            for p in ps:
                if p.dtype != torch.float32:
                    p.data = p.data.to(torch.float32)
            self._flat = flatten_params(ps, allow_cpu=True)
        return self._flat

    def split_flat(self, vec):
        out, o = [], 0
        for p in self.parameters():
            out.append(vec[o:o + p.numel()].view_as(p))
            o += p.numel()
        return out

    def forward(self, x):
        _check_input(self, x)
        out = _MLPValue.apply(self, x)
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            out = out + _MLPParamGrad.apply(self, x, *self.parameters())
        _remember(out, self, x)
        return out
