"""Pins the oracle (oracle/jet_oracle.py, numpy, graph-free) and the autograd port
(oracle/autograd_port.py) against golden vectors produced by the real reference's dnn.py + physics.py
(oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import autograd_port as ap
from oracle import jet_oracle as jo
from tests import cases


def _run_oracle(case, dtype):
    sres, sfid = cases.specs(case)
    flat, X, T, Xf, Tf = cases.data(case, dtype)
    if sfid is None:
        return jo.loss_and_grad(sres, flat, X, T)
    return jo.two_pass_loss_and_grad(sfid, sres, flat, Xf, Tf, X)


@pytest.mark.parametrize("name", cases.ALL)
def test_jet_oracle_fp64_matches_reference(name):
    case, z = cases.load(name)
    r = _run_oracle(case, np.float64)
    # float64 vs float64: only summation-order noise is allowed
    assert abs(r["loss"] - z["loss64"]) <= 1e-12 * abs(z["loss64"])
    assert abs(r["fidelity"] - z["fidelity64"]) <= 1e-12 * max(abs(z["fidelity64"]), 1e-30)
    assert abs(r["residual"] - z["residual64"]) <= 1e-12 * abs(z["residual64"])
    assert cases.golden_grad_check(z, r["grad"]) <= 1e-11


@pytest.mark.parametrize("name", cases.SMALL)
def test_jet_oracle_fp32_within_north_star_tolerance(name):
    case, z = cases.load(name)
    r = _run_oracle(case, np.float32)
    assert abs(r["loss"] - z["loss64"]) <= 1e-5 * abs(z["loss64"])      # north_star: 1e-5
    assert cases.golden_grad_check(z, r["grad"]) <= 1e-4                 # north_star: 1e-4


@pytest.mark.parametrize("name", ["cmb_h_small", "ftemp_small", "txyz", "leaky", "ragged", "cmb"])
def test_autograd_port_matches_reference(name):
    case, z = cases.load(name)
    sres, sfid = cases.specs(case)
    flat, X, T, Xf, Tf = cases.data(case, np.float64)
    tf = torch.from_numpy
    if sfid is None:
        r = ap.loss_and_grad(sres, tf(flat), tf(X.astype(np.float64)), tf(T.astype(np.float64)))
        loss, grad = r["loss"].item(), r["grad"].numpy()
    else:
        a = ap.loss_and_grad(sfid, tf(flat), tf(Xf.astype(np.float64)), tf(Tf.astype(np.float64)))
        b = ap.loss_and_grad(sres, tf(flat), tf(X.astype(np.float64)), None)
        loss = (sfid["w_fid"] * a["fidelity"] + sres["w_res"] * b["residual"]).item()
        grad = (a["grad"] + b["grad"]).numpy()
    assert abs(loss - z["loss64"]) <= 1e-12 * abs(z["loss64"])
    assert cases.golden_grad_check(z, grad) <= 1e-11


def test_sharded_sums_reassemble_global_mean():
    """a shard evaluated with n_global / global mask count contributes its exact share."""
    case, z = cases.load("cmb_h_small")
    sres, _ = cases.specs(case)
    flat, X, T, _, _ = cases.data(case, np.float64)
    n = X.shape[0]
    cut = 113
    cnt = float((X[:, 0] < 25.5).sum())
    a = jo.loss_and_grad(sres, flat, X[:cut], T[:cut], n_global=n, mask_count=cnt)
    b = jo.loss_and_grad(sres, flat, X[cut:], T[cut:], n_global=n, mask_count=cnt)
    assert abs((a["loss"] + b["loss"]) - z["loss64"]) <= 1e-12 * abs(z["loss64"])
    assert cases.golden_grad_check(z, a["grad"] + b["grad"]) <= 1e-11


def test_empty_mask_gives_nan_loss_like_reference():
    """physics.py:27-28: torch.mean over an empty selection is NaN; the gradient stays finite."""
    case, _ = cases.load("cmb_h_small")
    sres, _ = cases.specs(case)
    flat, X, T, _, _ = cases.data(case, np.float64)
    X = X.copy()
    X[:, 0] += 100.0     # no point satisfies x < 25.5
    r = jo.loss_and_grad(sres, flat, X, T)
    assert np.isnan(r["loss"]) and np.all(np.isfinite(r["grad"]))
    q = ap.loss_and_grad(sres, torch.from_numpy(flat), torch.from_numpy(X.astype(np.float64)),
                         torch.from_numpy(T.astype(np.float64)))
    assert torch.isnan(q["loss"]) and torch.isfinite(q["grad"]).all()
    np.testing.assert_allclose(r["grad"], q["grad"].numpy(), rtol=0, atol=1e-12 * np.abs(r["grad"]).max())


@pytest.mark.parametrize("name", cases.NAN)
def test_oracles_reproduce_the_reference_nan(name):
    """physics_equation with k == 0: sinh(2kh) = 0, the zero-E stress terms are 0 * (0/0) and the reference's loss
    is NaN (physics.py:106-108; golden made by running the real reference).  Both oracles must say NaN too."""
    case, z = cases.load(name)
    assert np.isnan(z["loss64"]) and np.isnan(z["loss32"]) and np.isnan(z["residual64"])
    assert np.isfinite(z["fidelity64"])
    sres, sfid = cases.specs(case)
    flat, X, T, Xf, Tf = cases.data(case, np.float64)
    with np.errstate(all="ignore"):
        r = jo.two_pass_loss_and_grad(sfid, sres, flat, Xf, Tf, X)
    assert np.isnan(r["loss"]) and np.isnan(r["residual"])
    assert abs(r["fidelity"] - z["fidelity64"]) <= 1e-12 * abs(z["fidelity64"])
    tf = torch.from_numpy
    b = ap.loss_and_grad(sres, tf(flat), tf(X.astype(np.float64)), None)
    assert torch.isnan(b["residual"])


@pytest.mark.parametrize("name", cases.BOUSS)
def test_decompiled_boussinesq_oracle_reproduces_its_golden(name):
    """The historical physics_functions residuals: golden = oracle/boussinesq_oracle.py (the decompiled bytecode) in float64;
    this guards the fixture against drift of the oracle and checks float32 against north_star's bounds."""
    from oracle import boussinesq_oracle as bo
    case, z = cases.load(name)
    flat, X, T, _, _ = cases.data(case, np.float32)
    spec = dict(layers=case["layers"], activation=case["activation"], kind=case["kind"], dirs=case["dirs"],
                target_cols=case["target_cols"], w_fid=case.get("w_fid", 1.0), w_res=case.get("w_res", 1.0))
    tf = torch.from_numpy
    r = bo.loss_and_grad(spec, tf(flat.astype(np.float64)), tf(X.astype(np.float64)), tf(T.astype(np.float64)))
    assert abs(r["loss"].item() - z["loss64"]) <= 1e-12 * abs(z["loss64"])
    assert cases.golden_grad_check(z, r["grad"].numpy()) <= 1e-11
    if name != "bouss_wide":
        r32 = bo.loss_and_grad(spec, tf(flat), tf(X), tf(T))
        assert abs(r32["loss"].item() - z["loss64"]) <= 1e-5 * abs(z["loss64"])
        assert cases.golden_grad_check(z, r32["grad"].numpy().astype(np.float64)) <= 1e-4
