"""GPU parity: the CUDA path (through the C ABI) against the golden vectors from the real reference
and against the numpy oracle on the same seeded inputs.

Tolerances are BASELINE.json's north_star: relative error <= 1e-5 on residuals and loss,
<= 1e-4 (norm-wise) on the weight gradient in FP32.
"""
import numpy as np
import pytest
import torch

from oracle import jet_oracle as jo
from tests import cases

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-5
GRAD_RTOL = 1e-4


def _oracle(case, n_override=None):
    sres, sfid = cases.specs(case)
    flat, X, T, Xf, Tf = cases.data(case, np.float64)
    if n_override is not None:
        X, T = X[:n_override], (T[:n_override] if T is not None else None)
    if sfid is None:
        return jo.loss_and_grad(sres, flat, X, T)
    return jo.two_pass_loss_and_grad(sfid, sres, flat, Xf, Tf, X)


@pytest.mark.parametrize("name", cases.ALL)
def test_loss_and_gradient_match_reference_golden(name):
    from tests.gpu_util import run_case
    case, z = cases.load(name)
    parts, grad, _, _ = run_case(case)
    assert abs(parts[2] - z["loss64"]) <= LOSS_RTOL * abs(z["loss64"]), (parts, z["loss64"])
    assert abs(parts[0] - z["fidelity64"]) <= LOSS_RTOL * max(abs(z["fidelity64"]), 1e-12)
    assert abs(parts[1] - z["residual64"]) <= LOSS_RTOL * abs(z["residual64"])
    assert np.all(np.isfinite(grad))
    assert cases.golden_grad_check(z, grad) <= GRAD_RTOL


@pytest.mark.parametrize("name", ["cmb_h_small", "txyz", "ragged", "leaky", "cmb"])
@pytest.mark.parametrize("n", [1, 3, 16, 17, 33])
def test_ragged_tiles_match_oracle(name, n):
    """tile tails: N not a multiple of the tile size, N smaller than one tile."""
    from tests.gpu_util import run_case
    case, _ = cases.load(name)
    parts, grad, _, _ = run_case(case, n_override=n)
    r = _oracle(case, n_override=n)
    assert abs(parts[2] - r["loss"]) <= LOSS_RTOL * abs(r["loss"])
    assert np.linalg.norm(grad - r["grad"]) <= GRAD_RTOL * np.linalg.norm(r["grad"])


def test_forward_only_outputs_and_jets_match_oracle():
    from pinn_depthestimation_b200.fused import JetLoss
    from tests.gpu_util import pass_specs
    case, _ = cases.load("txyz")
    sres, _ = pass_specs(case)
    flat, X, T, _, _ = cases.data(case, np.float32)
    dev = torch.device("cuda:0")
    jl = JetLoss(sres, torch.from_numpy(X).to(dev), torch.from_numpy(T).to(dev))
    n, o = X.shape[0], case["layers"][-1]
    out = torch.empty(n, o, device=dev)
    douts = [torch.empty(n, o, device=dev) for _ in range(3)]
    parts = jl.loss(torch.from_numpy(flat).to(dev), out=out, douts=douts).cpu().numpy()
    ospec, _ = cases.specs(case)
    r = jo.loss_and_grad(ospec, flat.astype(np.float64), X.astype(np.float64), T.astype(np.float64))
    assert abs(parts[2] - r["loss"]) <= LOSS_RTOL * abs(r["loss"])
    scale = np.abs(r["out"]).max()
    assert np.abs(out.cpu().numpy() - r["out"]).max() <= 1e-5 * scale
    for j in range(3):
        sj = np.abs(r["douts"][j]).max()
        assert np.abs(douts[j].cpu().numpy() - r["douts"][j]).max() <= 1e-5 * sj


def test_external_seeds_reverse_matches_oracle():
    """PINN_RES_EXTERNAL: caller-supplied d loss/d(out jets) -> weight gradient (autograd facade)."""
    import ctypes as C
    from pinn_depthestimation_b200 import PassSpec, _cabi
    from pinn_depthestimation_b200.fused import _Pass
    layers = [3, 24, 24, 24, 2]
    flat = jo.make_params(layers, 7, "tanh", np.float32)
    X, _ = jo.make_points(77, 3, 0, seed=5)
    rs = np.random.RandomState(3)
    so = rs.standard_normal((77, 2)).astype(np.float32)
    sd = [rs.standard_normal((77, 2)).astype(np.float32) for _ in range(2)]
    dev = torch.device("cuda:0")
    spec = PassSpec(layers=layers, kind="external", ext_dirs=[2, 0])
    ps = _Pass(spec, torch.from_numpy(X).to(dev), None)
    params = torch.from_numpy(flat).to(dev)
    grad = torch.empty_like(params)
    tso = torch.from_numpy(so).to(dev)
    tsd = [torch.from_numpy(a).to(dev) for a in sd]
    a = ps.args(params, grad, 1, 1, 0, seed_out=tso, seed_douts=tsd)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _cabi.check(_cabi.lib().pinn_jet_loss_fwdbwd(C.byref(ps.desc), C.byref(a), st))
    torch.cuda.synchronize()
    f64 = flat.astype(np.float64)
    _, _, tape = jo.jet_forward(layers, f64, X.astype(np.float64), [2, 0])
    ref = jo.jet_reverse(layers, f64, tape, so.astype(np.float64), [s.astype(np.float64) for s in sd])
    g = grad.cpu().numpy().astype(np.float64)
    assert np.linalg.norm(g - ref) <= GRAD_RTOL * np.linalg.norm(ref)


def test_empty_mask_is_nan_loss_finite_grad():
    """physics.py:27-28 with no point below the threshold: NaN loss (mean of nothing), finite grad."""
    from pinn_depthestimation_b200.fused import JetLoss
    from tests.gpu_util import pass_specs
    case, _ = cases.load("cmb_h_small")
    sres, _ = pass_specs(case)
    flat, X, T, _, _ = cases.data(case, np.float32)
    X = X.copy()
    X[:, 0] += 100.0
    dev = torch.device("cuda:0")
    jl = JetLoss(sres, torch.from_numpy(X).to(dev), torch.from_numpy(T).to(dev))
    params = torch.from_numpy(flat).to(dev)
    grad = torch.empty_like(params)
    parts = jl.loss_and_grad(params, grad).cpu().numpy()
    assert np.isnan(parts[2]) and np.isnan(parts[1]) and np.isfinite(parts[0])
    g = grad.cpu().numpy().astype(np.float64)
    ospec, _ = cases.specs(case)
    r = jo.loss_and_grad(ospec, flat.astype(np.float64), X.astype(np.float64), T.astype(np.float64))
    assert np.all(np.isfinite(g))
    assert np.linalg.norm(g - r["grad"]) <= GRAD_RTOL * np.linalg.norm(r["grad"])


def test_sums_are_additive_over_shards_at_scale():
    """Size-independent property at a size the oracle cannot reach: evaluating two shards with the
    global divisors reproduces the single-pass loss and gradient (what multi-GPU relies on)."""
    from pinn_depthestimation_b200 import PassSpec
    from pinn_depthestimation_b200.fused import JetLoss
    layers = [4] + [256] * 8 + [4]
    n = 1 << 17
    dev = torch.device("cuda:0")
    flat = torch.from_numpy(jo.make_params(layers, 1234, "tanh", np.float32)).to(dev)
    g = torch.Generator(device="cpu").manual_seed(1234)
    X = (torch.rand(n, 4, generator=g) * 2 - 1).to(dev)
    T = (0.05 * torch.randn(n, 4, generator=g)).to(dev)
    spec = PassSpec(layers=layers, kind="Navier_Stokes", dirs={"t": 0, "x": 1, "y": 2},
                    fields={"h": 0, "z": 1, "u": 2, "v": 3}, target_cols=[0, 1, 2, 3])
    full = JetLoss(spec, X, T)
    g_full = torch.empty_like(flat)
    p_full = full.loss_and_grad(flat, g_full).clone()
    cut = 50001
    acc = torch.zeros_like(flat)
    sums = torch.zeros(16, dtype=torch.float64, device=dev)
    for lo, hi in ((0, cut), (cut, n)):
        sh = JetLoss(spec, X[lo:hi].contiguous(), T[lo:hi].contiguous())
        sh.n_res_global = sh.n_fid_global = n
        gi = torch.empty_like(flat)
        sh.loss_and_grad(flat, gi)
        acc += gi
        sums += sh.res.sums
    loss = (sums[0] + sums[1] + sums[2]) / n + sums[5:9].sum() / n
    assert abs(loss.item() - p_full[2].item()) <= 2e-6 * abs(p_full[2].item())
    rel = (acc - g_full).norm() / g_full.norm()
    assert rel.item() <= 1e-5
    assert sums[13].item() == n


@pytest.mark.parametrize("name", cases.NAN)
def test_physics_equation_nan_poison_like_the_reference(name):
    """k == 0 => sinh(2kh) = 0 => the zero-E radiation-stress terms are 0 * (0/0): the reference's loss (and its
    gradient) are NaN (physics.py:106-108; golden from the real reference).  The fused epilogue must not hide that."""
    from tests.gpu_util import run_case
    case, z = cases.load(name)
    assert np.isnan(z["loss64"])
    parts, grad, _, _ = run_case(case)
    assert np.isnan(parts[2]) and np.isnan(parts[1])
    assert abs(parts[0] - z["fidelity64"]) <= LOSS_RTOL * abs(z["fidelity64"])   # the data misfit is unaffected
    assert np.isnan(grad).any()
