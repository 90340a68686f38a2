"""GPU tests of the historical `l_bfgs_b_optimizer.LBFGSBOptimizer` interface (reconstructed
from __pycache__/l_bfgs_b_optimizer.cpython-38.pyc): flat x <-> weights, `function_for_scipy(x) -> (loss, grad)`,
`optimize()` driving SciPy's L-BFGS-B -- checked against the torch-autograd restatement of the reference path."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import autograd_port as ap
from oracle import jet_oracle as jo
from tests import cases

pytestmark = pytest.mark.gpu
DROPIN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dropin")


def _setup(name="cmb_h_small"):
    if DROPIN not in sys.path:
        sys.path.insert(0, DROPIN)
    import dnn
    import physics
    from l_bfgs_b_optimizer import LBFGSBOptimizer
    dev = torch.device("cuda:0")
    case, z = cases.load(name)
    flat, X, T, _, _ = cases.data(case, np.float32)
    model = dnn.DNN(case["layers"], 0.0, "xavier").to(dev)
    x = torch.from_numpy(X[:, 0:1]).to(dev).requires_grad_(True)
    y = torch.from_numpy(X[:, 1:2]).to(dev).requires_grad_(True)
    Tt = torch.from_numpy(T).to(dev)

    def loss_function(mdl, inputs, outputs):          # the body of pinn.loss_func (train_newmethod.py:120-159)
        xx, yy = inputs
        pred = mdl(torch.cat([xx, yy], dim=-1))
        fid = sum(torch.nn.functional.mse_loss(pred[:, i:i + 1], outputs[:, i:i + 1]) for i in range(2))
        res = physics.continuity_only(xx, yy, pred[:, 2:3], pred[:, 0:1], pred[:, 1:2])
        return fid + res

    opt = LBFGSBOptimizer(model, (x, y), Tt, loss_function)
    return opt, model, case, z, flat, X, T


def test_interface_shape_and_defaults():
    opt, model, case, *_ = _setup()
    L = case["layers"]
    assert opt.shapes_and_sizes == [s for i in range(len(L) - 1)
                                    for s in (((L[i + 1], L[i]), L[i] * L[i + 1]), ((L[i + 1],), L[i + 1]))]
    assert set(opt.options) == {"maxiter", "maxfun", "maxcor", "maxls", "ftol"}
    assert opt.options["maxcor"] == 50 and opt.options["maxls"] == 50 and opt.options["ftol"] == np.finfo(float).eps


def test_function_for_scipy_matches_reference_loss_and_gradient():
    opt, model, case, z, flat, X, T = _setup()
    loss, grad = opt.function_for_scipy(flat.astype(np.float64))
    assert isinstance(loss, float) and grad.dtype == np.float64 and grad.shape == flat.shape
    assert abs(loss - z["loss64"]) <= 1e-5 * abs(z["loss64"])
    assert cases.golden_grad_check(z, grad) <= 1e-4
    # x really is installed as the weights (set_weights, source line 22), at another point too
    x2 = flat.astype(np.float64) * 0.9 + 0.01
    loss2, grad2 = opt.function_for_scipy(x2)
    sres, _ = cases.specs(case)
    tf = torch.from_numpy
    r = ap.loss_and_grad(sres, tf(x2), tf(X.astype(np.float64)), tf(T.astype(np.float64)))
    assert abs(loss2 - r["loss"].item()) <= 1e-5 * abs(r["loss"].item())
    assert np.linalg.norm(grad2 - r["grad"].numpy()) <= 1e-4 * np.linalg.norm(r["grad"].numpy())
    got = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).cpu().numpy()
    assert np.allclose(got, x2.astype(np.float32))


def test_optimize_drives_scipy_lbfgsb_downhill_and_installs_the_result():
    opt, model, case, z, flat, X, T = _setup()
    opt._set_weights(flat)
    opt.options = dict(opt.options, maxiter=15, maxfun=25)
    f0, _ = opt.function_for_scipy(flat.astype(np.float64))
    res = opt.optimize()
    assert res.nit >= 1 and res.fun < f0
    got = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).cpu().numpy()
    assert np.allclose(got, res.x.astype(np.float32))
    # the same SciPy driver on the reference algorithm (CPU) follows the same path: same iterate count, same loss
    from scipy.optimize import minimize
    sres, _ = cases.specs(case)
    tf = torch.from_numpy

    def ref_fun(xv):
        r = ap.loss_and_grad(sres, tf(xv.astype(np.float32)), tf(X), tf(T))
        return float(r["loss"]), r["grad"].numpy().astype(np.float64)
    ref = minimize(fun=ref_fun, x0=flat.astype(np.float64), jac=True, method="L-BFGS-B", options=opt.options)
    assert abs(res.fun - ref.fun) <= 1e-3 * abs(ref.fun)
    assert abs(res.nit - ref.nit) <= 2


def test_flat_closure_fast_path():
    """loss_function may be an object with flat_loss_and_grad (trainer.FusedClosure): no per-parameter .grad."""
    if DROPIN not in sys.path:
        sys.path.insert(0, DROPIN)
    from l_bfgs_b_optimizer import LBFGSBOptimizer
    from pinn_depthestimation_b200 import PassSpec
    from pinn_depthestimation_b200.dnn import DNN
    from pinn_depthestimation_b200.fused import JetLoss
    from tests.gpu_util import pass_specs
    dev = torch.device("cuda:0")
    case, z = cases.load("txyz")
    flat, X, T, _, _ = cases.data(case, np.float32)
    spec, _ = pass_specs(case)
    jl = JetLoss(spec, torch.from_numpy(X).to(dev), torch.from_numpy(T).to(dev))

    class Closure:
        def flat_loss_and_grad(self, fp, fg):
            return jl.loss_and_grad(fp, fg)
    model = DNN(case["layers"], 0.0, "xavier").to(dev)
    opt = LBFGSBOptimizer(model, None, None, Closure())
    loss, grad = opt.function_for_scipy(flat.astype(np.float64))
    assert abs(loss - z["loss64"]) <= 1e-5 * abs(z["loss64"])
    assert cases.golden_grad_check(z, grad) <= 1e-4
    assert all(p.grad is None for p in model.parameters())
