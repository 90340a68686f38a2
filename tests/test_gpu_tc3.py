"""GPU tests of the split-operand tensor-core mode (precision="tf32x3": every operand carried as a TF32 hi + lo pair,
tcgen05 kind::tf32, FP32 accumulate, tanhf) -- held to north_star's FP32 tolerances: relative error <= 1e-5 on
loss / residual / misfit and <= 1e-4 norm-wise on the weight gradient against the reference's float64 golden."""
import numpy as np
import pytest
import torch

from oracle import jet_oracle as jo
from tests import cases

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-5
GRAD_RTOL = 1e-4


@pytest.mark.parametrize("name", ["wide_nswe", "wide_cont", "wide_wave", "wide_ftemp"])
def test_tf32x3_matches_reference_golden_at_fp32_tolerance(name):
    from tests.gpu_util import run_case
    case, z = cases.load(name)
    parts, grad, _, _ = run_case(case, precision="tf32x3")
    el = abs(parts[2] - z["loss64"]) / abs(z["loss64"])
    eg = cases.golden_grad_check(z, grad)
    print(f"{name}: tf32x3 loss rel {el:.2e}, grad {eg:.2e}")
    assert el <= LOSS_RTOL
    assert abs(parts[0] - z["fidelity64"]) <= LOSS_RTOL * abs(z["fidelity64"])
    assert abs(parts[1] - z["residual64"]) <= LOSS_RTOL * abs(z["residual64"])
    assert np.all(np.isfinite(grad))
    assert eg <= GRAD_RTOL


@pytest.mark.parametrize("n", [1, 15, 17, 33, 100])
def test_tf32x3_ragged_tiles(n):
    from tests.gpu_util import run_case
    case, _ = cases.load("wide_nswe")
    parts, grad, _, _ = run_case(case, precision="tf32x3", n_override=n)
    sres, _ = cases.specs(case)
    flat, X, T, _, _ = cases.data(case, np.float64)
    r = jo.loss_and_grad(sres, flat, X[:n], T[:n])
    assert abs(parts[2] - r["loss"]) <= LOSS_RTOL * abs(r["loss"])
    assert np.linalg.norm(grad - r["grad"]) <= GRAD_RTOL * np.linalg.norm(r["grad"])


def _run(prec, layers, kind, kw, X, T, flat):
    from pinn_depthestimation_b200 import PassSpec
    from pinn_depthestimation_b200.fused import JetLoss
    jl = JetLoss(PassSpec(layers=layers, kind=kind, precision=prec, **kw), X, T)
    grad = torch.empty_like(flat)
    parts = jl.loss_and_grad(flat, grad).clone()
    torch.cuda.synchronize()
    return parts, grad, jl


_NSWE = dict(dirs={"t": 0, "x": 1, "y": 2}, fields={"h": 0, "z": 1, "u": 2, "v": 3}, target_cols=[0, 1, 2, 3])
_CONT = dict(dirs={"x": 0, "y": 1}, fields={"U": 0, "V": 1, "h": 2}, target_cols=[0, 1])


@pytest.mark.parametrize("kind", ["Navier_Stokes", "continuity_only"])
def test_tf32x3_many_tiles_per_cta_agrees_with_fp32_kernel(kind):
    """~55 tiles per persistent CTA (every mbarrier phase wraps many times); the FP32 FMA kernel is the yardstick."""
    dev = torch.device("cuda:0")
    n = (1 << 17) + 77
    layers, kw = ([4] + [256] * 8 + [4], _NSWE) if kind == "Navier_Stokes" else ([2] + [256] * 3 + [3], _CONT)
    flat = torch.from_numpy(jo.make_params(layers, 1234, "tanh", np.float32)).to(dev)
    g = torch.Generator(device="cpu").manual_seed(7)
    X = (torch.rand(n, layers[0], generator=g) * 2 - 1).to(dev)
    T = (0.05 * torch.randn(n, len(kw["target_cols"]), generator=g)).to(dev)
    p32, g32, _ = _run("fp32", layers, kind, kw, X, T, flat)
    p3, g3, jl = _run("tf32x3", layers, kind, kw, X, T, flat)
    el = abs(p3[2].item() - p32[2].item()) / abs(p32[2].item())
    eg = ((g3 - g32).norm() / g32.norm()).item()
    print(f"{kind}: tf32x3 vs fp32 kernel: loss {el:.2e} grad {eg:.2e}")
    assert torch.isfinite(g3).all()
    assert el <= 3e-6 and eg <= 3e-5
    assert jl.res.sums[13].item() == n
    g2 = torch.empty_like(g3)
    p2 = jl.loss_and_grad(flat, g2)          # same workspace again: no state leaks between launches
    assert abs(p2[2].item() - p3[2].item()) <= 1e-6 * abs(p3[2].item())
    assert ((g2 - g3).norm() / g3.norm()).item() <= 1e-5


@pytest.mark.parametrize("hidden", [2, 5])
def test_tf32x3_other_depths_agree_with_fp32_kernel(hidden):
    dev = torch.device("cuda:0")
    layers = [4] + [256] * hidden + [4]
    flat = torch.from_numpy(jo.make_params(layers, 1234, "tanh", np.float32)).to(dev)
    g = torch.Generator(device="cpu").manual_seed(11)
    n = 16 * 7 + 5                                   # odd number of tiles: one padding tile in the last pair
    X = (torch.rand(n, 4, generator=g) * 2 - 1).to(dev)
    T = (0.05 * torch.randn(n, 4, generator=g)).to(dev)
    p32, g32, _ = _run("fp32", layers, "Navier_Stokes", _NSWE, X, T, flat)
    p3, g3, _ = _run("tf32x3", layers, "Navier_Stokes", _NSWE, X, T, flat)
    assert torch.isfinite(g3).all()
    assert abs(p3[2].item() - p32[2].item()) <= 3e-6 * abs(p32[2].item())
    assert ((g3 - g32).norm() / g32.norm()).item() <= 3e-5


def test_tf32x3_forward_only_loss_and_outputs():
    from pinn_depthestimation_b200.fused import JetLoss
    from tests.gpu_util import pass_specs
    dev = torch.device("cuda:0")
    case, z = cases.load("wide_nswe")
    spec, _ = pass_specs(case, "tf32x3")
    flat, X, T, _, _ = cases.data(case, np.float32)
    jl = JetLoss(spec, torch.from_numpy(X).to(dev), torch.from_numpy(T).to(dev))
    out = torch.empty(X.shape[0], 4, device=dev)
    parts = jl.loss(torch.from_numpy(flat).to(dev), out=out).cpu().numpy()
    assert abs(parts[2] - z["loss64"]) <= LOSS_RTOL * abs(z["loss64"])
    assert np.abs(out.cpu().numpy()[:16] - z["out64_head"]).max() <= 1e-5 * np.abs(z["out64_head"]).max()


def test_tf32x3_shards_add_up_at_bench_scale():
    """2^19 points of the bench net: two ragged shards evaluated with the global divisors add up to the single pass
    (what the multi-GPU all-reduce relies on)."""
    from pinn_depthestimation_b200 import PassSpec
    from pinn_depthestimation_b200.fused import JetLoss
    layers = [4] + [256] * 8 + [4]
    n = 1 << 19
    dev = torch.device("cuda:0")
    flat = torch.from_numpy(jo.make_params(layers, 1234, "tanh", np.float32)).to(dev)
    g = torch.Generator(device="cpu").manual_seed(99)
    X = (torch.rand(n, 4, generator=g) * 2 - 1).to(dev)
    T = (0.05 * torch.randn(n, 4, generator=g)).to(dev)
    kw = dict(layers=layers, kind="Navier_Stokes", precision="tf32x3", **_NSWE)

    def run(x, t, n_global=None):
        jl = JetLoss(PassSpec(**kw), x, t)
        if n_global is not None:
            jl.n_res_global = jl.n_fid_global = n_global
        gr = torch.empty_like(flat)
        parts = jl.loss_and_grad(flat, gr).clone()
        torch.cuda.synchronize()
        return parts, gr, jl.res.sums.clone()

    p_full, g_full, s_full = run(X, T)
    assert s_full[13].item() == n and torch.isfinite(g_full).all()
    cut = 133337
    acc = torch.zeros_like(flat)
    sums = torch.zeros(16, dtype=torch.float64, device=dev)
    for lo, hi in ((0, cut), (cut, n)):
        _, gi, si = run(X[lo:hi].contiguous(), T[lo:hi].contiguous(), n_global=n)
        acc += gi
        sums += si
    loss = (sums[0] + sums[1] + sums[2]) / n + sums[5:9].sum() / n
    assert abs(loss.item() - p_full[2].item()) <= 2e-6 * abs(p_full[2].item())
    assert ((acc - g_full).norm() / g_full.norm()).item() <= 1e-5
    assert sums[13].item() == n
