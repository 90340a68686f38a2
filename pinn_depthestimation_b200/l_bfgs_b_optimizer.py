"""Torch-backed re-offer of the reference's historical `l_bfgs_b_optimizer.LBFGSBOptimizer`
(only shipped as __pycache__/l_bfgs_b_optimizer.cpython-38.pyc; reconstructed in SURVEY.md 2.3).

Interface kept: `LBFGSBOptimizer(model, inputs, outputs, loss_function, options=None)`,
`.shapes_and_sizes`, `.options` (maxiter/maxfun/maxcor/maxls/ftol), `.function_for_scipy(x) ->
(loss, flat_grad)`, `.optimize()` driving scipy.optimize.minimize(method='L-BFGS-B', jac=True).

`loss_function(model, inputs, outputs)` returns a differentiable scalar (its backward runs the
fused reverse kernel when `model` is a pinn_depthestimation_b200 DNN), or is an object with
`flat_loss_and_grad(flat_params, flat_grad) -> parts` (trainer.FusedClosure), in which case no
per-parameter gradients are materialised.  The SciPy driver is host code by definition; every loss /
gradient evaluation runs on the GPU.
"""
from __future__ import annotations

import numpy as np
import torch

from .lbfgs import flatten_params


class LBFGSBOptimizer:
    def __init__(self, model, inputs, outputs, loss_function, options=None):
        self.model = model
        self.inputs = inputs
        self.outputs = outputs
        self.loss_function = loss_function
        self.options = options or {'maxiter': 50000, 'maxfun': 50000, 'maxcor': 50, 'maxls': 50,
                                   'ftol': 1.0 * np.finfo(float).eps}
        self.shapes_and_sizes = [(tuple(p.shape), p.numel()) for p in model.parameters()]
        self._flat = flatten_params(model.parameters())
        self._grad = torch.zeros_like(self._flat)

    def _set_weights(self, x):
        x = np.asarray(x, dtype=np.float64)
        if x.size != self._flat.numel():
            raise ValueError(f"expected {self._flat.numel()} weights, got {x.size}")
        with torch.no_grad():
            self._flat.copy_(torch.from_numpy(x.astype(np.float32)).to(self._flat.device))

    def function_for_scipy(self, x):
        self._set_weights(x)
        fl = getattr(self.loss_function, "flat_loss_and_grad", None)
        if fl is not None:
            parts = fl(self._flat, self._grad)
            loss = parts[2]
        else:
            for p in self.model.parameters():
                p.grad = None
            with torch.enable_grad():
                loss = self.loss_function(self.model, self.inputs, self.outputs)
            loss.backward()
            views = [p.grad.reshape(-1) if p.grad is not None else p.new_zeros(p.numel())
                     for p in self.model.parameters()]
            torch.cat(views, 0, out=self._grad)
        return (float(loss.detach().cpu().double()),
                self._grad.detach().cpu().numpy().astype(np.float64))

    def optimize(self):
        from scipy.optimize import minimize
        x0 = self._flat.detach().cpu().numpy().astype(np.float64)
        result = minimize(fun=self.function_for_scipy, x0=x0, jac=True, method='L-BFGS-B',
                          options=self.options)
        self._set_weights(result.x)
        return result
