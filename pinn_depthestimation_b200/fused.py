"""One loss evaluation (+ flat weight gradient) of the reference's pinn.loss_func, as a pair of
C-ABI launches on the current CUDA stream.

`JetLoss` is the fused fast path that bench.py, the trainer and the physics/dnn facades share:

    jl = JetLoss(spec_res, inputs, targets)              # train_newmethod.py form (one point set)
    jl = JetLoss(spec_res, inputs_res, None,             # train.py form (two point sets)
                 fid=(spec_fid, inputs_fid, targets_fid))
    parts = jl.loss_and_grad(flat_params, flat_grad)     # device tensor [fidelity, residual, total, 0]

Nothing here synchronises with the host.  With torch.distributed initialised and `group` given,
the points passed in are this rank's shard; raw sums and gradient are all-reduced (one collective
over a single fp32 buffer) before the means are formed.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _cabi
from .spec import PassSpec


def _dev_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: this path has no CPU fallback")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    return t.contiguous()


class _Pass:
    """Buffers + descriptor of one pass over one point set."""

    def __init__(self, spec: PassSpec, inputs: torch.Tensor, targets: Optional[torch.Tensor],
                 want_grad: bool = True):
        self.spec = spec
        self.desc = spec.to_desc()
        self.want_grad = bool(want_grad)
        self.inputs = _dev_f32(inputs, "inputs")
        if self.inputs.dim() != 2 or self.inputs.shape[1] != spec.layers[0]:
            raise ValueError(f"inputs must be [N,{spec.layers[0]}], got {tuple(self.inputs.shape)}")
        self.n = int(self.inputs.shape[0])
        self._mask_key = None
        self.targets = None
        if spec.target_cols:
            if targets is None:
                raise ValueError("spec has target_cols but no targets were given")
            self.targets = _dev_f32(targets, "targets")
            if tuple(self.targets.shape) != (self.n, len(spec.target_cols)):
                raise ValueError("targets must be [N, len(target_cols)]")
        dev = self.inputs.device
        lib = _cabi.lib()
        nbytes = C.c_size_t(0)
        with torch.cuda.device(dev):
            _cabi.check(lib.pinn_workspace_bytes_ex(C.byref(self.desc), self.n, 1 if self.want_grad else 0,
                                                    C.byref(nbytes)), "pinn_workspace_bytes_ex")
        self.workspace = torch.empty(nbytes.value + 256, dtype=torch.uint8, device=dev)
        off = (-self.workspace.data_ptr()) % 256
        self._ws_ptr = self.workspace.data_ptr() + off
        self._ws_bytes = nbytes.value
        self.sums = torch.zeros(_cabi.NSUMS, dtype=torch.float64, device=dev)
        self.mask_count = None
        if spec.kind == "continuity_only":
            self.mask_count = torch.zeros(1, dtype=torch.float32, device=dev)
            self._count_mask()

    def _count_mask(self):
        """#{x < 25.5} of the current inputs (physics.py:27) -> self.mask_count (this rank's points only)."""
        dev = self.inputs.device
        st = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            _cabi.check(_cabi.lib().pinn_mask_count(C.byref(self.desc), _cabi.ptr(self.inputs), self.n,
                                                    _cabi.ptr(self.mask_count), C.c_void_p(st)),
                        "pinn_mask_count")
        self._mask_key = (self.inputs.data_ptr(), self.inputs._version)

    def rebind(self, inputs: torch.Tensor, targets: Optional[torch.Tensor] = None):
        """Point this pass at another [N,d] input tensor of the same shape (the reference's loss_func builds a fresh
        `torch.cat` of its input columns on every call): descriptor, workspace and sums are kept, only the pointers
        change; the mask count is redone when the tensor (or its version) differs from the one it was counted on."""
        inputs = _dev_f32(inputs, "inputs")
        if tuple(inputs.shape) != tuple(self.inputs.shape) or inputs.device != self.inputs.device:
            raise ValueError("rebind: inputs must keep their shape and device")
        self.inputs = inputs
        if targets is not None:
            targets = _dev_f32(targets, "targets")
            if self.targets is None or tuple(targets.shape) != tuple(self.targets.shape):
                raise ValueError("rebind: targets must keep their shape")
            self.targets = targets
        if self.mask_count is not None and self._mask_key != (inputs.data_ptr(), inputs._version):
            self._count_mask()

    def args(self, params, grad, n_res_global, n_fid_global, flags, out=None, douts=None,
             seed_out=None, seed_douts=None, shard=None):
        """shard = (first point, count): evaluate only that slice of the point set (sums and gradient add up over
        slices because they are sums with the GLOBAL divisors)."""
        a = _cabi.EvalArgs()
        a.params = params.data_ptr()
        lo, cnt = shard if shard is not None else (0, self.n)
        a.inputs = self.inputs.data_ptr() + 4 * lo * self.inputs.shape[1] if self.n else None
        a.targets = (self.targets.data_ptr() + 4 * lo * self.targets.shape[1]
                     if self.targets is not None and self.n else None)
        a.n_points = cnt
        a.n_res_global = n_res_global
        a.n_fid_global = n_fid_global
        a.mask_count = self.mask_count.data_ptr() if self.mask_count is not None else None
        a.seed_out = seed_out.data_ptr() if seed_out is not None else None
        for j in range(_cabi.MAX_DIRS):
            sd = seed_douts[j] if seed_douts is not None and j < len(seed_douts) else None
            a.seed_dout[j] = sd.data_ptr() if sd is not None else None
            do = douts[j] if douts is not None and j < len(douts) else None
            a.dout[j] = do.data_ptr() if do is not None else None
        a.grad = grad.data_ptr() if grad is not None else None
        a.sums = self.sums.data_ptr()
        a.out = out.data_ptr() if out is not None else None
        a.workspace = self._ws_ptr
        a.workspace_bytes = self._ws_bytes
        a.flags = flags
        return a


# Points per launch above which a loss+gradient evaluation is split into several launches whose gradients are added
# outside the kernel: inside one launch a gradient element is the FP32 sum (atomics) of one contribution per tile, and
# the rounding of ~10^6 such additions shows (tools/accum_noise.py: 2.1e-4 norm-wise for the FP32 kernel's 16-point
# tiles at 16.8M points, 1.8e-5 for the tensor-core kernel's 64- / 32-point tile pairs, which contract in TMEM first).
SHARD_POINTS = {"fp32": 1 << 21, "tf32": 1 << 24, "tf32x3": 1 << 24}


class JetLoss:
    def __init__(self, spec: PassSpec, inputs: torch.Tensor, targets: Optional[torch.Tensor] = None,
                 fid: Optional[Tuple[PassSpec, torch.Tensor, torch.Tensor]] = None, group=None,
                 shard_points: Optional[int] = None):
        self.shard_points = int(shard_points) if shard_points else SHARD_POINTS.get(spec.precision, 1 << 21)
        self._shard_tmp = None
        self.res = _Pass(spec, inputs, targets)
        self.fid = _Pass(*fid) if fid is not None else None
        if self.fid is not None and self.fid.spec.layers != spec.layers:
            raise ValueError("fidelity and residual passes must share the network")
        self.device = self.res.inputs.device
        self.n_params = spec.n_params
        self.parts = torch.zeros(4, dtype=torch.float32, device=self.device)
        self.group = group
        self.world = 1
        self.n_res_global = self.res.n
        self.n_fid_global = self.fid.n if self.fid is not None else self.res.n
        if group is not None:
            import torch.distributed as dist
            from . import dist as pdist
            self.world = dist.get_world_size(group)
            self.n_res_global, self.n_fid_global = pdist.global_counts(
                self.res.n, self.fid.n if self.fid is not None else self.res.n, self.device, group)
            if self.res.mask_count is not None:
                dist.all_reduce(self.res.mask_count, group=group)
            # one fp32 buffer [P + 2*NSUMS] per evaluation
            self._coll = torch.zeros(pdist.collective_numel(self.n_params), dtype=torch.float32,
                                     device=self.device)

    # ------------------------------------------------------------------
    def _shards(self, p: _Pass):
        k = (p.n + self.shard_points - 1) // self.shard_points
        m = (p.n + k - 1) // k if k > 0 else 0
        return [(i * m, min(m, p.n - i * m)) for i in range(k) if p.n - i * m > 0]

    @property
    def launches_per_eval(self) -> int:
        """kernel launches of this library per loss+gradient evaluation (pack + jet kernel per pass and shard, finalize)"""
        passes = [q for q in (self.fid, self.res) if q is not None]
        if not any(q.n > self.shard_points for q in passes):
            return 2 * len(passes) + 1
        return sum(1 + len(self._shards(q)) for q in passes) + 1

    def _launch_sharded(self, params, grad):
        """loss+gradient in slices of at most `shard_points` points: every slice accumulates into a zeroed scratch vector
        that is then added to `grad`, so no FP32 running sum sees more than one slice's tiles."""
        lib = _cabi.lib()
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        if self._shard_tmp is None or self._shard_tmp.numel() != grad.numel():
            self._shard_tmp = torch.empty_like(grad)
        tmp = self._shard_tmp
        grad.zero_()
        with torch.cuda.device(self.device):
            for p in [q for q in (self.fid, self.res) if q is not None]:
                p.sums.zero_()
                for i, sh in enumerate(self._shards(p)):
                    tmp.zero_()
                    flags = _cabi.FLAG_ACCUMULATE | (_cabi.FLAG_SKIP_PACK if i > 0 else 0)
                    a = p.args(params, tmp, self.n_res_global, self.n_fid_global, flags, shard=sh)
                    _cabi.check(lib.pinn_jet_loss_fwdbwd(C.byref(p.desc), C.byref(a), st), "sharded pass")
                    grad.add_(tmp)

    def _launch(self, params, grad, want_grad, out=None, douts=None):
        if want_grad and out is None and douts is None and (
                self.res.n > self.shard_points or (self.fid is not None and self.fid.n > self.shard_points)):
            return self._launch_sharded(params, grad)
        lib = _cabi.lib()
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        fn = lib.pinn_jet_loss_fwdbwd if want_grad else lib.pinn_jet_loss_fwd
        with torch.cuda.device(self.device):
            flags = 0
            if self.fid is not None:
                a = self.fid.args(params, grad, self.n_res_global, self.n_fid_global, 0)
                _cabi.check(fn(C.byref(self.fid.desc), C.byref(a), st), "fidelity pass")
                flags = _cabi.FLAG_ACCUMULATE     # gradient adds; sums go to the pass's own buffer
                self.res.sums.zero_()
            a = self.res.args(params, grad, self.n_res_global, self.n_fid_global, flags, out, douts)
            _cabi.check(fn(C.byref(self.res.desc), C.byref(a), st), "residual pass")

    def _finalize(self):
        lib = _cabi.lib()
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        # the descriptor that owns the loss weights and target list
        d = self.fid.desc if self.fid is not None else self.res.desc
        dres = self.res.desc
        with torch.cuda.device(self.device):
            if self.fid is None:
                _cabi.check(lib.pinn_loss_finalize(
                    C.byref(dres), _cabi.ptr(self.res.sums), None, self.n_fid_global,
                    self.n_res_global, _cabi.ptr(self.res.mask_count), _cabi.ptr(self.parts), st),
                    "pinn_loss_finalize")
            else:
                # residual desc has no targets of its own in the two-pass form: finalize with a
                # merged view (targets from the fidelity desc, residual kind from the residual desc)
                m = _cabi.Desc.from_buffer_copy(dres)
                m.n_targets = d.n_targets
                for i in range(_cabi.MAX_OUT):
                    m.target_cols[i] = d.target_cols[i]
                    m.target_w[i] = d.target_w[i]
                m.w_fid = d.w_fid
                _cabi.check(lib.pinn_loss_finalize(
                    C.byref(m), _cabi.ptr(self.res.sums), _cabi.ptr(self.fid.sums),
                    self.n_fid_global, self.n_res_global, _cabi.ptr(self.res.mask_count),
                    _cabi.ptr(self.parts), st), "pinn_loss_finalize")

    def _allreduce(self, grad):
        from . import dist as pdist
        pdist.all_reduce_eval(self._coll, grad, self.res.sums,
                              self.fid.sums if self.fid is not None else None, self.group)

    # ------------------------------------------------------------------
    def loss_and_grad(self, params: torch.Tensor, grad: torch.Tensor) -> torch.Tensor:
        """params, grad: flat fp32 CUDA vectors [P].  Returns the device tensor
        [fidelity, residual, total, 0] (train_newmethod.py:133,156,159) -- no host sync."""
        params = _dev_f32(params, "params")
        if grad.dtype != torch.float32 or not grad.is_cuda or not grad.is_contiguous():
            raise TypeError("grad must be a contiguous float32 CUDA tensor")
        if params.numel() != self.n_params or grad.numel() != self.n_params:
            raise ValueError(f"params/grad must have {self.n_params} elements")
        self._launch(params, grad, True)
        if self.group is not None and self.world > 1:
            self._allreduce(grad)
        self._finalize()
        return self.parts

    def loss(self, params: torch.Tensor, out=None, douts=None) -> torch.Tensor:
        params = _dev_f32(params, "params")
        self._launch(params, None, False, out, douts)
        if self.group is not None and self.world > 1:
            self._allreduce(None)
        self._finalize()
        return self.parts
