"""GPU tests of the drop-in boundary: `dnn.DNN` and `physics.*` used exactly the way
pinn.loss_func uses them (train_newmethod.py:120-159, train.py:128-157), compared with the golden
vectors from the real reference."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import jet_oracle as jo
from tests import cases

pytestmark = pytest.mark.gpu

DROPIN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dropin")


def _dropin():
    if DROPIN not in sys.path:
        sys.path.insert(0, DROPIN)
    import dnn
    import physics
    return dnn, physics


def _model(case, dev):
    dnn, _ = _dropin()
    init = "xavier" if case["activation"] == "tanh" else "kaiming"
    m = dnn.DNN(case["layers"], 0.0, init).to(dev)
    flat = jo.make_params(case["layers"], 1234, case["activation"], np.float32)
    with torch.no_grad():
        o = 0
        for p in m.parameters():
            p.copy_(torch.from_numpy(flat[o:o + p.numel()]).view_as(p))
            o += p.numel()
    m.train()
    return m


def _flat_grad(m):
    return torch.cat([p.grad.reshape(-1) for p in m.parameters()]).cpu().numpy().astype(np.float64)


def test_state_dict_keys_and_forward_match_oracle():
    dev = torch.device("cuda:0")
    case, _ = cases.load("cmb_h_small")
    m = _model(case, dev)
    keys = list(m.state_dict().keys())
    assert keys[:2] == ["layers.layer_0.weight", "layers.layer_0.bias"]
    assert len(keys) == 2 * (len(case["layers"]) - 1)
    flat, X, _, _, _ = cases.data(case, np.float32)
    out = m(torch.from_numpy(X).to(dev)).detach().cpu().numpy()
    ref = jo.mlp_forward(case["layers"], flat.astype(np.float64), X.astype(np.float64))
    assert np.abs(out - ref).max() <= 1e-5 * np.abs(ref).max()


def test_invalid_init_type_raises_like_reference():
    dnn, _ = _dropin()
    with pytest.raises(ValueError):
        dnn.DNN([2, 4, 1], 0.0, "he")


@pytest.mark.parametrize("name", ["cmb_h_small", "ftemp_small", "txyz", "leaky", "ragged"])
def test_loss_func_single_pass_form(name):
    """The body of train_newmethod.py:120-159 on top of the drop-in modules."""
    _, physics = _dropin()
    dev = torch.device("cuda:0")
    case, z = cases.load(name)
    m = _model(case, dev)
    _, X, T, _, _ = cases.data(case, np.float32)
    d = X.shape[1]
    col_of = {c: n for n, c in case["dirs"].items()}
    cols = [torch.tensor(X[:, c:c + 1].astype(np.float64), requires_grad=(c in col_of)).float().to(dev)
            for c in range(d)]
    named = {col_of[c]: cols[c] for c in col_of}
    pred = m(torch.cat(cols, dim=-1))
    tw = case.get("target_w", [1.0] * len(case["target_cols"]))
    fid = 0
    for i, c in enumerate(case["target_cols"]):
        fid = fid + tw[i] * torch.nn.functional.mse_loss(pred[:, c:c + 1], torch.from_numpy(T[:, i:i + 1]).to(dev))
    f = {n: pred[:, c:c + 1] for n, c in case["fields"].items()}
    if case["kind"] == jo.CONT_ONLY:
        res = physics.continuity_only(named["x"], named["y"], f["h"], f["U"], f["V"])
    elif case["kind"] == jo.CONT_FTEMP:
        res = physics.continuity_ftemp(named["x"], named["y"], f["h"], f["U"], f["V"])
    else:
        res = physics.Navier_Stokes(named["t"], named["x"], named["y"], f["h"], f["z"], f["u"], f["v"])
    loss = case.get("w_fid", 1.0) * fid + case.get("w_res", 1.0) * res
    for p in m.parameters():
        p.grad = None
    loss.backward()
    assert abs(loss.item() - z["loss64"]) <= 1e-5 * abs(z["loss64"])
    assert abs(res.item() - z["residual64"]) <= 1e-5 * abs(z["residual64"])
    assert cases.golden_grad_check(z, _flat_grad(m)) <= 1e-4


def test_loss_func_two_pass_form_physics_equation():
    """The body of train.py:128-157 (fidelity forward + residual forward, config_CMB shape)."""
    _, physics = _dropin()
    dev = torch.device("cuda:0")
    case, z = cases.load("cmb")
    m = _model(case, dev)
    _, X, _, Xf, Tf = cases.data(case, np.float32)
    pf = m(torch.from_numpy(Xf).to(dev))
    fid = 0
    for i, c in enumerate(case["target_cols"]):
        fid = fid + case["target_w"][i] * torch.mean((torch.from_numpy(Tf[:, i:i + 1]).to(dev) - pf[:, c:c + 1]) ** 2)
    cols = [torch.tensor(X[:, c:c + 1].astype(np.float64), requires_grad=True).float().to(dev) for c in range(2)]
    pred = m(torch.cat(cols, dim=-1))
    f = {n: pred[:, c:c + 1] for n, c in case["fields"].items()}
    res = physics.physics_equation(cols[0], cols[1], f["h"], f["U"], f["V"], f["eta_mean"], f["Hrms"], f["k"])
    loss = 1.0 * fid + 1.0 * res
    loss.backward()
    assert abs(loss.item() - z["loss64"]) <= 1e-5 * abs(z["loss64"])
    assert cases.golden_grad_check(z, _flat_grad(m)) <= 1e-4


def test_generic_compute_gradient_on_jets():
    """physics.compute_gradient (autograd.grad with create_graph) through DNN: d out/d x comes from
    forward jets and is differentiable w.r.t. the weights, so a residual written with plain torch
    ops in the reference's style gives the reference's loss and gradient."""
    _, physics = _dropin()
    dev = torch.device("cuda:0")
    case, z = cases.load("ftemp_small")
    m = _model(case, dev)
    _, X, T, _, _ = cases.data(case, np.float32)
    x = torch.from_numpy(X[:, 0:1]).to(dev).requires_grad_(True)
    y = torch.from_numpy(X[:, 1:2]).to(dev).requires_grad_(True)
    pred = m(torch.cat([x, y], dim=-1))
    U, V, h = pred[:, 0:1], pred[:, 1:2], pred[:, 2:3]
    fc = physics.compute_gradient(h * U, x) + physics.compute_gradient(h * V, y)
    res = torch.mean(fc ** 2)
    fid = sum(torch.nn.functional.mse_loss(pred[:, i:i + 1], torch.from_numpy(T[:, i:i + 1]).to(dev)) for i in range(2))
    loss = case["w_fid"] * fid + case["w_res"] * res
    loss.backward()
    assert abs(loss.item() - z["loss64"]) <= 1e-5 * abs(z["loss64"])
    assert cases.golden_grad_check(z, _flat_grad(m)) <= 1e-4


def test_whole_module_checkpoint_round_trip(tmp_path):
    """torch.save(self.dnn) / torch.load as in train_newmethod.py:184 and test_newmethod.py:35-42."""
    dev = torch.device("cuda:0")
    case, _ = cases.load("cmb_h_small")
    m = _model(case, dev)
    path = tmp_path / "model.pth"
    torch.save(m, path)
    m2 = torch.load(path, map_location=dev, weights_only=False)
    _, X, _, _, _ = cases.data(case, np.float32)
    xt = torch.from_numpy(X).to(dev)
    assert torch.equal(m(xt), m2(xt))


def test_cpu_tensor_is_rejected_not_silently_computed():
    dev = torch.device("cuda:0")
    case, _ = cases.load("cmb_h_small")
    m = _model(case, dev)
    with pytest.raises(RuntimeError):
        m(torch.zeros(4, 2))
