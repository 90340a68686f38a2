"""L-BFGS with the `torch.optim.LBFGS` contract, vector work on the GPU through the C ABI.

Reference call sites: train_newmethod.py:108-117 (construction from config['lbfgs_optimizer']),
train_newmethod.py:204-209 (`optimizer_LBFGS.step(closure)`, once, max_iter=50000).  The algorithm
(and every termination test) is torch/optim/lbfgs.py:12-209 (cubic interpolation, strong-Wolfe
bracket + zoom) and :333-537 (`step`), restated branch for branch:

* the two-loop recursion (lbfgs.py:432-447) is ONE thread-block-cluster kernel over ring buffers
  `hist_s/hist_y [history_size, P]` (pinn_lbfgs_direction) instead of 2m host-synchronising dots;
* g.d, sum|g|, max|g|, max|d|, y.s, y.y come from one deterministic cluster reduction
  (pinn_vec_stats); `params = x0 + t*d` is one launch on the flat parameter vector;
* with line_search_fn='strong_wolfe' (every reference config) the whole optimiser state machine between two
  closure evaluations -- strong-Wolfe bracket / zoom transition, the gradient clones, termination tests,
  curvature-pair update, two-loop recursion, step length, next trial point -- is ONE cluster kernel
  (pinn_lbfgs_advance, csrc/lbfgs_dev.cu): the host launches the closure, then that kernel, and reads one small
  status block (one stream synchronisation per evaluation);
* the same logic is kept here as host code (`strong_wolfe`, `cubic_interpolate`, `LBFGS._step_host`): it serves
  line_search_fn=None, is unit-tested on CPU against torch's own private functions, and is the yardstick the
  device state machine is compared with (tests/test_gpu_optim.py).

`step(closure)` accepts the reference's closure (zero_grad -> loss_func -> backward -> return
loss).  A closure object that also has `flat_loss_and_grad(flat_params, flat_grad) -> parts`
(see trainer.py) is evaluated without touching per-parameter `.grad` tensors.
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _cabi


# ------------------------------------------------------------------------------------------------
# host-side scalar logic (pure Python floats; unit-tested on CPU against torch's own functions)
# ------------------------------------------------------------------------------------------------
def _ieee_div(a, b):
    """a / b with IEEE semantics (inf / nan instead of ZeroDivisionError), like torch's tensor maths."""
    try:
        return a / b
    except ZeroDivisionError:
        if a == 0 or a != a:
            return math.nan
        return math.copysign(math.inf, a) * math.copysign(1.0, b)


def cubic_interpolate(x1, f1, g1, x2, f2, g2, bounds=None):
    """Minimiser of the cubic through (x1,f1,g1), (x2,f2,g2) clipped to bounds (lbfgs.py:12-37).
    A collapsed bracket (x1 == x2) yields inf/nan intermediates and falls through to bisection or the
    clipping, exactly as torch's 0-d tensor arithmetic does."""
    if bounds is not None:
        xmin_bound, xmax_bound = bounds
    else:
        xmin_bound, xmax_bound = (x1, x2) if x1 <= x2 else (x2, x1)
    d1 = g1 + g2 - 3 * _ieee_div(f1 - f2, x1 - x2)
    d2_square = d1 * d1 - g1 * g2
    if d2_square >= 0:
        d2 = math.sqrt(d2_square)
        if x1 <= x2:
            min_pos = x2 - (x2 - x1) * _ieee_div(g2 + d2 - d1, g2 - g1 + 2 * d2)
        else:
            min_pos = x1 - (x1 - x2) * _ieee_div(g1 + d2 - d1, g1 - g2 + 2 * d2)
        return min(max(min_pos, xmin_bound), xmax_bound)
    return (xmin_bound + xmax_bound) / 2.0


def strong_wolfe(evaluate, clone, t, f, g, gtd, d_norm, c1=1e-4, c2=0.9, tolerance_change=1e-9,
                 max_ls=25):
    """Strong-Wolfe line search (lbfgs.py:40-209).

    evaluate(t) -> (f_new, g_new, gtd_new): loss, gradient handle and g_new.d at x + t*d; the
    gradient handle may be overwritten by the next evaluate call, so `clone(handle)` is used wherever
    torch clones.  Returns (f, g_handle, t, n_evals).
    """
    g = clone(g)
    f_new, g_new, gtd_new = evaluate(t)
    ls_func_evals = 1
    t_prev, f_prev, g_prev, gtd_prev = 0.0, f, g, gtd
    done = False
    ls_iter = 0
    bracket = bracket_f = bracket_g = bracket_gtd = None
    while ls_iter < max_ls:
        if f_new > (f + c1 * t * gtd) or (ls_iter > 1 and f_new >= f_prev):
            bracket, bracket_f = [t_prev, t], [f_prev, f_new]
            bracket_g, bracket_gtd = [g_prev, clone(g_new)], [gtd_prev, gtd_new]
            break
        if abs(gtd_new) <= -c2 * gtd:
            bracket, bracket_f, bracket_g = [t], [f_new], [g_new]
            done = True
            break
        if gtd_new >= 0:
            bracket, bracket_f = [t_prev, t], [f_prev, f_new]
            bracket_g, bracket_gtd = [g_prev, clone(g_new)], [gtd_prev, gtd_new]
            break
        min_step = t + 0.01 * (t - t_prev)
        max_step = t * 10
        tmp = t
        t = cubic_interpolate(t_prev, f_prev, gtd_prev, t, f_new, gtd_new, bounds=(min_step, max_step))
        t_prev, f_prev, g_prev, gtd_prev = tmp, f_new, clone(g_new), gtd_new
        f_new, g_new, gtd_new = evaluate(t)
        ls_func_evals += 1
        ls_iter += 1
    if ls_iter == max_ls or bracket is None:
        bracket, bracket_f, bracket_g = [0.0, t], [f, f_new], [g, g_new]
        bracket_gtd = [gtd, gtd_new]
    insuf_progress = False
    low_pos, high_pos = (0, 1) if bracket_f[0] <= bracket_f[-1] else (1, 0)
    while not done and ls_iter < max_ls:
        if abs(bracket[1] - bracket[0]) * d_norm < tolerance_change:
            break
        t = cubic_interpolate(bracket[0], bracket_f[0], bracket_gtd[0],
                              bracket[1], bracket_f[1], bracket_gtd[1])
        eps = 0.1 * (max(bracket) - min(bracket))
        if min(max(bracket) - t, t - min(bracket)) < eps:
            if insuf_progress or t >= max(bracket) or t <= min(bracket):
                if abs(t - max(bracket)) < abs(t - min(bracket)):
                    t = max(bracket) - eps
                else:
                    t = min(bracket) + eps
                insuf_progress = False
            else:
                insuf_progress = True
        else:
            insuf_progress = False
        f_new, g_new, gtd_new = evaluate(t)
        ls_func_evals += 1
        ls_iter += 1
        if f_new > (f + c1 * t * gtd) or f_new >= bracket_f[low_pos]:
            bracket[high_pos], bracket_f[high_pos] = t, f_new
            bracket_g[high_pos], bracket_gtd[high_pos] = clone(g_new), gtd_new
            low_pos, high_pos = (0, 1) if bracket_f[0] <= bracket_f[1] else (1, 0)
        else:
            if abs(gtd_new) <= -c2 * gtd:
                done = True
            elif gtd_new * (bracket[high_pos] - bracket[low_pos]) >= 0:
                bracket[high_pos], bracket_f[high_pos] = bracket[low_pos], bracket_f[low_pos]
                bracket_g[high_pos], bracket_gtd[high_pos] = bracket_g[low_pos], bracket_gtd[low_pos]
            bracket[low_pos], bracket_f[low_pos] = t, f_new
            bracket_g[low_pos], bracket_gtd[low_pos] = clone(g_new), gtd_new
    if len(bracket) == 1:
        low_pos = 0
    return bracket_f[low_pos], bracket_g[low_pos], bracket[low_pos], ls_func_evals


# ------------------------------------------------------------------------------------------------
# flat parameter plumbing shared by LBFGS and FusedAdam
# ------------------------------------------------------------------------------------------------
def flatten_params(params, allow_cpu=False):
    """Make every parameter a view into ONE flat fp32 buffer (parameters() order) and return it.
    Idempotent: parameters that already are consecutive views of one buffer are left alone, so the DNN facade,
    LBFGS, FusedAdam and LBFGSBOptimizer all end up sharing the same flat vector whichever comes first."""
    params = list(params)
    if not params:
        raise ValueError("optimizer got an empty parameter list")
    for p in params:
        if not p.is_cuda and not allow_cpu:
            raise RuntimeError("pinn_b200 optimisers need CUDA parameters: no CPU fallback")
        if p.dtype != torch.float32:
            raise TypeError("pinn_b200 optimisers need float32 parameters")
    total = sum(p.numel() for p in params)
    base = params[0].data_ptr()
    consecutive, o = True, 0
    for p in params:
        if p.data_ptr() != base + 4 * o or not p.is_contiguous():
            consecutive = False
            break
        o += p.numel()
    if consecutive:
        st = params[0].untyped_storage()
        off = (base - st.data_ptr()) // 4
        if st.nbytes() >= (off + total) * 4:
            flat = torch.empty(0, dtype=torch.float32, device=params[0].device)
            flat.set_(st, off, (total,), (1,))
            return flat
    flat = torch.cat([p.detach().reshape(-1) for p in params]).contiguous()
    o = 0
    for p in params:
        p.data = flat[o:o + p.numel()].view_as(p)
        o += p.numel()
    return flat


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


class _Vec:
    """Device vector helpers bound to one device (all launches on the current stream)."""

    def __init__(self, device):
        self.device = device
        self.lib = _cabi.lib()
        self.stats = torch.zeros(8, dtype=torch.float32, device=device)
        self.host = torch.zeros(8, dtype=torch.float32).pin_memory()

    def read_stats(self, a, b, extra=None):
        """-> [a.b, sum|a|, max|a|, max|b|, a.a, b.b, extra] as Python floats (ONE host sync)."""
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.pinn_vec_stats(_cabi.ptr(a), _cabi.ptr(b), a.numel(),
                                                _cabi.ptr(self.stats), _stream(self.device)),
                        "pinn_vec_stats")
        if extra is not None:
            self.stats[6:7].copy_(extra.detach().reshape(1).to(torch.float32))
        self.host.copy_(self.stats, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return [float(v) for v in self.host.tolist()]

    def axpy_into(self, out, x0, alpha, d):
        """out = x0 + alpha*d (one copy + one pinn_axpy launch)."""
        out.copy_(x0)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.pinn_axpy(float(alpha), _cabi.ptr(d), _cabi.ptr(out), out.numel(),
                                           _stream(self.device)), "pinn_axpy")

    def axpy(self, alpha, x, y):
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.pinn_axpy(float(alpha), _cabi.ptr(x), _cabi.ptr(y), y.numel(),
                                           _stream(self.device)), "pinn_axpy")


class LBFGS(torch.optim.Optimizer):
    """Same constructor and `step(closure)` behaviour as torch.optim.LBFGS (lbfgs.py:218-537)."""

    def __init__(self, params, lr=1, max_iter=20, max_eval=None, tolerance_grad=1e-7,
                 tolerance_change=1e-9, history_size=100, line_search_fn=None):
        if not 0.0 <= lr:
            raise ValueError(f"Invalid learning rate: {lr}")
        if max_eval is None:
            max_eval = max_iter * 5 // 4
        defaults = dict(lr=lr, max_iter=max_iter, max_eval=max_eval, tolerance_grad=tolerance_grad,
                        tolerance_change=tolerance_change, history_size=history_size,
                        line_search_fn=line_search_fn)
        super().__init__(params, defaults)
        if len(self.param_groups) != 1:
            raise ValueError("LBFGS doesn't support per-parameter options (parameter groups)")
        self._params = self.param_groups[0]["params"]
        self._flat = None
        self._bufs = None
        self._dev = None
        self.device_line_search = True    # False: keep the scalar logic on the host (tests compare the two)

    # ---- buffers --------------------------------------------------------------------------------
    def _setup(self):
        flat = flatten_params(self._params)
        if self._flat is None or self._flat.data_ptr() != flat.data_ptr():
            self._flat = flat
            P, dev = flat.numel(), flat.device
            m = int(self.param_groups[0]["history_size"])
            z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
            # ring of history_size+1 slots: the candidate (s,y) pair is written into the free slot and
            # only becomes part of the history if y.s > 1e-10 (lbfgs.py:408)
            self._bufs = dict(g=z(P), d=z(P), prev_g=z(P), x0=z(P), S=z(m + 1, P), Y=z(m + 1, P),
                              rho=z(m + 1), h_diag=torch.ones(1, dtype=torch.float32, device=dev),
                              scratch=z(2 * m + 64), head=0, used=0)
            self._vec = _Vec(dev)
        return self._flat

    def _evaluate(self, closure):
        """closure -> (device loss scalar); leaves the flat gradient in self._bufs['g']."""
        g = self._bufs["g"]
        fl = getattr(closure, "flat_loss_and_grad", None)
        if fl is not None:
            parts = fl(self._flat, g)
            return parts[2]
        with torch.enable_grad():
            loss = closure()
        views = []
        for p in self._params:
            views.append(p.grad.reshape(-1) if p.grad is not None else p.new_zeros(p.numel()))
        torch.cat(views, 0, out=g)
        return loss

    def _direction(self):
        b = self._bufs
        m = b["S"].shape[0]
        with torch.cuda.device(self._flat.device):
            _cabi.check(self._vec.lib.pinn_lbfgs_direction(
                _cabi.ptr(b["S"]), _cabi.ptr(b["Y"]), _cabi.ptr(b["rho"]), _cabi.ptr(b["h_diag"]),
                _cabi.ptr(b["g"]), _cabi.ptr(b["d"]), m, b["used"], b["head"], self._flat.numel(),
                _cabi.ptr(b["scratch"]), _stream(self._flat.device)), "pinn_lbfgs_direction")

    # ---- device-resident state machine -------------------------------------------------------------
    def _setup_device(self):
        flat = flatten_params(self._params)
        m = int(self.param_groups[0]["history_size"])
        if self._dev is None or self._dev["flat_ptr"] != flat.data_ptr() or self._dev["m"] != m:
            lib = _cabi.lib()
            nbytes = C.c_size_t(0)
            _cabi.check(lib.pinn_lbfgs_workspace_bytes(flat.numel(), m, C.byref(nbytes)), "pinn_lbfgs_workspace_bytes")
            ws = torch.empty(nbytes.value + 256, dtype=torch.uint8, device=flat.device)
            off = (-ws.data_ptr()) % 256
            self._dev = dict(flat_ptr=flat.data_ptr(), m=m, ws=ws, ws_ptr=ws.data_ptr() + off, fresh=True,
                             g=torch.zeros_like(flat),
                             status=torch.zeros(C.sizeof(_cabi.LbfgsStatus), dtype=torch.uint8).pin_memory())
        self._flat = flat
        return flat

    def _evaluate_into(self, closure, g):
        """closure -> device loss scalar [1]; the flat gradient lands in g."""
        fl = getattr(closure, "flat_loss_and_grad", None)
        if fl is not None:
            return fl(self._flat, g)[2:3]
        with torch.enable_grad():
            loss = closure()
        views = [p.grad.reshape(-1) if p.grad is not None else p.new_zeros(p.numel()) for p in self._params]
        torch.cat(views, 0, out=g)
        return loss.detach().reshape(1).to(torch.float32)

    def _step_device(self, closure):
        group = self.param_groups[0]
        flat = self._setup_device()
        dv, lib, dev = self._dev, _cabi.lib(), flat.device
        state = self.state[self._params[0]]
        state.setdefault("func_evals", 0)
        state.setdefault("n_iter", 0)
        cfg = _cabi.LbfgsCfg(float(group["lr"]), float(group["tolerance_grad"]), float(group["tolerance_change"]),
                             int(group["max_iter"]), int(group["max_eval"]), dv["m"])
        with torch.cuda.device(dev):
            _cabi.check(lib.pinn_lbfgs_begin(C.c_void_p(dv["ws_ptr"]), flat.numel(), C.byref(cfg),
                                             1 if dv["fresh"] else 0, _stream(dev)), "pinn_lbfgs_begin")
        dv["fresh"] = False
        stream = torch.cuda.current_stream(dev)
        status = _cabi.LbfgsStatus.from_buffer(dv["status"].numpy())
        orig_loss = None
        while True:
            loss = self._evaluate_into(closure, dv["g"])
            if orig_loss is None:
                orig_loss = loss.clone()
            if loss.dtype != torch.float32 or not loss.is_cuda:
                loss = loss.to(device=dev, dtype=torch.float32)
            with torch.cuda.device(dev):
                _cabi.check(lib.pinn_lbfgs_advance(C.c_void_p(dv["ws_ptr"]), flat.numel(), dv["m"], _cabi.ptr(flat),
                                                   _cabi.ptr(dv["g"]), _cabi.ptr(loss),
                                                   C.c_void_p(dv["status"].data_ptr()), _stream(dev)),
                            "pinn_lbfgs_advance")
            stream.synchronize()                      # the one host round trip per evaluation
            if status.code != 1:
                break
        if status.code != 2:
            raise RuntimeError(f"pinn_lbfgs_advance returned status {status.code}")
        state["n_iter"] = int(status.n_iter_total)
        state["func_evals"] = int(status.func_evals_total)
        state["t"], state["loss"] = float(status.t), float(status.loss)
        state["history_used"] = int(status.history_used)
        return orig_loss.reshape(())

    @torch.no_grad()
    def step(self, closure):
        if self.param_groups[0]["line_search_fn"] == "strong_wolfe" and self.device_line_search:
            return self._step_device(closure)
        return self._step_host(closure)

    @torch.no_grad()
    def _step_host(self, closure):
        group = self.param_groups[0]
        lr = float(group["lr"])
        max_iter, max_eval = group["max_iter"], group["max_eval"]
        tolerance_grad, tolerance_change = group["tolerance_grad"], group["tolerance_change"]
        line_search_fn, history_size = group["line_search_fn"], int(group["history_size"])
        if line_search_fn is not None and line_search_fn != "strong_wolfe":
            raise RuntimeError("only 'strong_wolfe' is supported")

        flat = self._setup()
        b, vec = self._bufs, self._vec
        state = self.state[self._params[0]]
        state.setdefault("func_evals", 0)
        state.setdefault("n_iter", 0)

        orig_loss = self._evaluate(closure).detach().clone()   # (a flat closure returns a view of a reused device buffer)
        st = vec.read_stats(b["g"], b["d"], extra=orig_loss)
        loss = st[6]
        current_evals = 1
        state["func_evals"] += 1
        if st[2] <= tolerance_grad:          # max|g|
            return orig_loss

        t = state.get("t")
        prev_loss = state.get("prev_loss")
        have_prev = state.get("have_prev", False)
        g, d = b["g"], b["d"]
        n_iter = 0
        while n_iter < max_iter:
            n_iter += 1
            state["n_iter"] += 1
            if state["n_iter"] == 1:
                b["used"], b["head"] = 0, 0
                b["h_diag"].fill_(1.0)
            else:
                # memory update (lbfgs.py:404-421): y = g - prev_g, s = t*d into the next ring slot
                cap = history_size + 1
                slot = (b["head"] + b["used"]) % cap
                torch.sub(g, b["prev_g"], out=b["Y"][slot])
                torch.mul(d, t, out=b["S"][slot])
                ys_st = vec.read_stats(b["Y"][slot], b["S"][slot])
                ys, yy = ys_st[0], ys_st[4]
                if ys > 1e-10:
                    if b["used"] == history_size:
                        b["head"] = (b["head"] + 1) % cap            # drop the oldest pair
                    else:
                        b["used"] += 1
                    b["rho"][slot:slot + 1].fill_(1.0 / ys)
                    b["h_diag"].fill_(ys / yy)
            self._direction()                # d = -H g  (two-loop recursion, one cluster kernel)
            b["prev_g"].copy_(g)
            have_prev = True
            prev_loss = loss
            st = vec.read_stats(g, d)
            gtd, g_l1, d_norm = st[0], st[1], st[3]
            if state["n_iter"] == 1:
                t = min(1.0, 1.0 / g_l1) * lr
            else:
                t = lr
            if gtd > -tolerance_change:
                break
            ls_func_evals = 0
            if line_search_fn is not None:
                b["x0"].copy_(flat)

                def evaluate(tt):
                    vec.axpy_into(flat, b["x0"], tt, d)
                    lt = self._evaluate(closure)
                    s2 = vec.read_stats(g, d, extra=lt)
                    return s2[6], g, s2[0]

                loss, g_best, t, ls_func_evals = strong_wolfe(
                    evaluate, lambda h: h.clone(), t, loss, g, gtd, d_norm,
                    max_ls=max_eval - current_evals)   # torch passes only max_ls: zoom keeps its 1e-9 default
                if g_best.data_ptr() != g.data_ptr():
                    g.copy_(g_best)
                vec.axpy_into(flat, b["x0"], t, d)
                gmax = vec.read_stats(g, d)[2]
                opt_cond = gmax <= tolerance_grad
            else:
                vec.axpy(t, d, flat)
                opt_cond = False
                if n_iter != max_iter:
                    lt = self._evaluate(closure)
                    s2 = vec.read_stats(g, d, extra=lt)
                    loss, opt_cond = s2[6], s2[2] <= tolerance_grad
                    ls_func_evals = 1
            current_evals += ls_func_evals
            state["func_evals"] += ls_func_evals
            if n_iter == max_iter:
                break
            if current_evals >= max_eval:
                break
            if opt_cond:
                break
            if d_norm * abs(t) <= tolerance_change:
                break
            if abs(loss - prev_loss) < tolerance_change:
                break
        state["t"] = t
        state["prev_loss"] = prev_loss
        state["have_prev"] = have_prev
        state["loss"] = loss
        return orig_loss


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam semantics (train_newmethod.py:95-98) as one launch on the flat vector;
    works with torch.optim.lr_scheduler.StepLR (train_newmethod.py:101-105)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        if len(self.param_groups) != 1:
            raise ValueError("FusedAdam supports a single parameter group")
        self._params = self.param_groups[0]["params"]
        self._flat = None

    def _setup(self):
        flat = flatten_params(self._params)
        if self._flat is None or self._flat.data_ptr() != flat.data_ptr():
            self._flat = flat
            self._m = torch.zeros_like(flat)
            self._v = torch.zeros_like(flat)
            self._g = torch.zeros_like(flat)
            self._step = 0
        return flat

    @torch.no_grad()
    def step(self, closure=None, flat_grad=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        flat = self._setup()
        if flat_grad is None:
            views = [p.grad.reshape(-1) if p.grad is not None else p.new_zeros(p.numel())
                     for p in self._params]
            torch.cat(views, 0, out=self._g)
            flat_grad = self._g
        g = self.param_groups[0]
        self._step += 1
        with torch.cuda.device(flat.device):
            _cabi.check(_cabi.lib().pinn_adam_step(
                _cabi.ptr(flat), _cabi.ptr(flat_grad), _cabi.ptr(self._m), _cabi.ptr(self._v),
                flat.numel(), float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]),
                float(g["eps"]), float(g["weight_decay"]), self._step, _stream(flat.device)),
                "pinn_adam_step")
        return loss
