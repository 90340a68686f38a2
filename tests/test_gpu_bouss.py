"""GPU tests of the historical physics_functions residuals: `Boussinesq` on the third-order
Taylor-jet kernel (csrc/jet3.cu) and `Boussinesq_simple` on the first-order kernels, against golden vectors made by
running the decompiled bytecode (oracle/boussinesq_oracle.py) under torch autograd in float64.  FP32 tolerances of
north_star: 1e-5 on loss / residual, 1e-4 norm-wise on the weight gradient."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import jet_oracle as jo
from tests import cases

pytestmark = pytest.mark.gpu
LOSS_RTOL, GRAD_RTOL = 1e-5, 1e-4
DROPIN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dropin")


@pytest.mark.parametrize("name", cases.BOUSS)
def test_loss_and_gradient_match_the_decompiled_reference(name):
    from tests.gpu_util import run_case
    case, z = cases.load(name)
    parts, grad, _, _ = run_case(case)
    el = abs(parts[2] - z["loss64"]) / abs(z["loss64"])
    eg = cases.golden_grad_check(z, grad)
    print(f"{name}: loss rel {el:.2e} (fid {abs(parts[0] - z['fidelity64']) / abs(z['fidelity64']):.1e}, "
          f"res {abs(parts[1] - z['residual64']) / abs(z['residual64']):.1e}), grad {eg:.2e}")
    assert abs(parts[0] - z["fidelity64"]) <= LOSS_RTOL * abs(z["fidelity64"])
    assert abs(parts[1] - z["residual64"]) <= LOSS_RTOL * abs(z["residual64"])
    assert el <= LOSS_RTOL
    assert np.all(np.isfinite(grad))
    assert eg <= GRAD_RTOL


@pytest.mark.parametrize("n", [1, 7, 9, 23])
def test_third_order_kernel_ragged_tiles_and_forward_only(n):
    """tile tails of the 8-point tiles; the forward-only entry gives the same loss and the network outputs."""
    import torch as th
    from oracle import boussinesq_oracle as bo
    from pinn_depthestimation_b200.fused import JetLoss
    from tests.gpu_util import pass_specs
    case, _ = cases.load("bouss")
    flat, X, T, _, _ = cases.data(case, np.float32)
    spec = dict(layers=case["layers"], activation="tanh", kind=case["kind"], dirs=case["dirs"], target_cols=case["target_cols"])
    r = bo.loss_and_grad(spec, th.from_numpy(flat.astype(np.float64)), th.from_numpy(X[:n].astype(np.float64)),
                         th.from_numpy(T[:n].astype(np.float64)))
    dev = th.device("cuda:0")
    ps, _ = pass_specs(case)
    jl = JetLoss(ps, th.from_numpy(X[:n]).to(dev), th.from_numpy(T[:n]).to(dev))
    p = th.from_numpy(flat).to(dev)
    g = th.full_like(p, float("nan"))
    parts = jl.loss_and_grad(p, g).cpu().numpy().astype(np.float64)
    assert abs(parts[2] - r["loss"].item()) <= LOSS_RTOL * abs(r["loss"].item())
    gr = r["grad"].numpy()
    assert np.linalg.norm(g.cpu().numpy() - gr) <= GRAD_RTOL * np.linalg.norm(gr)
    out = th.empty(n, 4, device=dev)
    parts_f = jl.loss(p, out=out).cpu().numpy().astype(np.float64)
    assert abs(parts_f[2] - parts[2]) <= 1e-6 * abs(parts[2])
    assert np.abs(out.cpu().numpy() - r["out"].numpy()).max() <= 1e-5 * np.abs(r["out"].numpy()).max()
    assert jl.res.sums[13].item() == n


def test_boussinesq_simple_runs_on_the_tensor_core_modes_too():
    """a first-order residual: the shared epilogue serves the FP32, TF32 and split-operand kernels alike."""
    from pinn_depthestimation_b200 import PassSpec
    from pinn_depthestimation_b200.fused import JetLoss
    dev = torch.device("cuda:0")
    layers = [3] + [256] * 3 + [4]
    kw = dict(kind="Boussinesq_simple", dirs={"t": 0, "x": 1, "y": 2}, fields={"h": 0, "z": 1, "u": 2, "v": 3}, target_cols=[0, 1, 2, 3])
    flat = torch.from_numpy(jo.make_params(layers, 1234, "tanh", np.float32)).to(dev)
    X, T = jo.make_points(300, 3, 4, seed=3)
    Xd, Td = torch.from_numpy(X).to(dev), torch.from_numpy(T).to(dev)
    res = {}
    for prec in ("fp32", "tf32x3", "tf32"):
        jl = JetLoss(PassSpec(layers=layers, precision=prec, **kw), Xd, Td)
        g = torch.empty_like(flat)
        parts = jl.loss_and_grad(flat, g).clone()
        torch.cuda.synchronize()
        res[prec] = (parts[2].item(), g)
    for prec, lt, gt in (("tf32x3", 3e-6, 3e-5), ("tf32", 5e-3, 5e-3)):
        assert abs(res[prec][0] - res["fp32"][0]) <= lt * abs(res["fp32"][0])
        assert ((res[prec][1] - res["fp32"][1]).norm() / res["fp32"][1].norm()).item() <= gt


def test_physics_functions_dropin_on_the_dnn_facade():
    """`from physics_functions import Boussinesq` exactly as the historical code called it: the whole DNN output plus the
    t, x, y columns; loss.backward() delivers the reference's gradient."""
    if DROPIN not in sys.path:
        sys.path.insert(0, DROPIN)
    import dnn
    import physics_functions as pf
    dev = torch.device("cuda:0")
    for name in ("bouss", "bouss_simple"):
        case, z = cases.load(name)
        flat, X, T, _, _ = cases.data(case, np.float32)
        m = dnn.DNN(case["layers"], 0.0, "xavier").to(dev)
        with torch.no_grad():
            m.flat_params().copy_(torch.from_numpy(flat).to(dev))
        t, x, y = (torch.from_numpy(X[:, i:i + 1]).to(dev).requires_grad_(True) for i in range(3))
        out = m(torch.cat([t, x, y], dim=-1))
        fn = pf.Boussinesq if name == "bouss" else pf.Boussinesq_simple
        res = fn(out, t, x, y, dev)
        fid = sum(torch.mean((torch.from_numpy(T[:, i:i + 1]).to(dev) - out[:, i:i + 1]) ** 2) for i in range(4))
        loss = fid + res
        loss.backward()
        assert abs(res.item() - z["residual64"]) <= LOSS_RTOL * abs(z["residual64"])
        assert abs(loss.item() - z["loss64"]) <= LOSS_RTOL * abs(z["loss64"])
        grad = torch.cat([p.grad.reshape(-1) for p in m.parameters()]).cpu().numpy().astype(np.float64)
        assert cases.golden_grad_check(z, grad) <= GRAD_RTOL
