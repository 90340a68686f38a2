"""The TORCH_LIBRARY(pinn_b200) registration: the library loads and exposes every op without a GPU; on a
GPU the ops give the ctypes binding's numbers (same entry points underneath)."""
import numpy as np
import pytest
import torch

from tests import cases

OPS = ["workspace_bytes", "jet_loss", "mask_count", "loss_finalize", "lbfgs_direction", "vec_stats", "axpy_", "adam_step_"]


def test_library_loads_and_registers_every_op():
    from pinn_depthestimation_b200 import torch_ops
    ops = torch_ops.load()
    for name in OPS:
        assert hasattr(ops, name), name
    # schema sanity: the descriptor is checked before anything touches a device
    with pytest.raises(RuntimeError, match="desc must be"):
        ops.workspace_bytes(torch.zeros(3, dtype=torch.uint8), 10, True)


@pytest.mark.gpu
def test_ops_match_the_ctypes_binding_on_a_golden_case():
    from pinn_depthestimation_b200 import torch_ops
    from tests.gpu_util import pass_specs, run_case
    ops = torch_ops.load()
    dev = torch.device("cuda:0")
    case, z = cases.load("cmb_h_small")
    spec, _ = pass_specs(case)
    flat, X, T, _, _ = cases.data(case, np.float32)
    desc = torch_ops.desc_tensor(spec)
    Xd, Td, p = torch.from_numpy(X).to(dev), torch.from_numpy(T).to(dev), torch.from_numpy(flat).to(dev)
    n = X.shape[0]
    ws = torch.empty(ops.workspace_bytes(desc, n, True) + 256, dtype=torch.uint8, device=dev)
    mc = ops.mask_count(desc, Xd)
    assert mc.item() == n                                     # x is normalised: the mask is all-true
    grad = torch.full_like(p, float("nan"))
    sums = ops.jet_loss(desc, p, Xd, Td, mc, n, n, ws, grad, None, 0)
    parts = ops.loss_finalize(desc, sums, None, n, n, mc).cpu().numpy().astype(np.float64)
    assert abs(parts[2] - z["loss64"]) <= 1e-5 * abs(z["loss64"])
    assert cases.golden_grad_check(z, grad.cpu().numpy()) <= 1e-4
    parts_c, grad_c, _, _ = run_case(case)                    # the ctypes path
    assert abs(parts[2] - parts_c[2]) <= 1e-6 * abs(parts_c[2])
    # forward-only with outputs
    out = torch.empty(n, 3, device=dev)
    ops.jet_loss(desc, p, Xd, Td, mc, n, n, ws, None, out, 0)
    assert np.abs(out.cpu().numpy()[:16] - z["out64_head"]).max() <= 1e-5 * np.abs(z["out64_head"]).max()
    # vector ops
    a, b = torch.randn(1000, device=dev), torch.randn(1000, device=dev)
    st = ops.vec_stats(a, b).cpu()
    assert abs(st[0].item() - (a @ b).item()) <= 1e-3 and abs(st[2].item() - a.abs().max().item()) == 0
    y = b.clone()
    ops.axpy_(0.5, a, y)
    assert torch.allclose(y, b + 0.5 * a)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.vec_stats(a.cpu(), None)
