"""Sharded evaluation (fused.JetLoss above SHARD_POINTS points per launch): slices of the point set evaluated with the
GLOBAL divisors and added outside the kernel must reproduce the one-launch answer -- for both trainer forms (one point
set: train_newmethod.py:120-159; fidelity + residual point sets: train.py:128-157), the continuity_only mask count
(physics.py:27) and both FP32-tolerance kernels."""
import numpy as np
import pytest
import torch

from tests import cases

pytestmark = pytest.mark.gpu


def _run(case, precision, shard_points):
    from pinn_depthestimation_b200.fused import JetLoss
    from tests.gpu_util import pass_specs
    dev = torch.device("cuda:0")
    sres, sfid = pass_specs(case, precision)
    flat, X, T, Xf, Tf = cases.data(case, np.float32)
    t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    if sfid is None:
        jl = JetLoss(sres, t(X), t(T), shard_points=shard_points)
    else:
        jl = JetLoss(sres, t(X), None, fid=(sfid, t(Xf), t(Tf)), shard_points=shard_points)
    p = t(flat)
    g = torch.full_like(p, float("nan"))
    parts = jl.loss_and_grad(p, g).cpu().numpy().astype(np.float64)
    return parts, g.cpu().numpy().astype(np.float64), jl


@pytest.mark.parametrize("name,precision", [("txyz", "fp32"), ("cmb_h_small", "fp32"), ("cmb", "fp32"), ("config_json", "fp32"),
                                            ("wide_nswe", "tf32x3"), ("wide_cont", "tf32x3")])
def test_sharded_evaluation_reproduces_one_launch(name, precision):
    case, _ = cases.load(name)
    parts1, g1, jl1 = _run(case, precision, None)
    n = jl1.res.n
    assert n <= jl1.shard_points                      # default: one launch
    partsS, gS, jlS = _run(case, precision, max(7, n // 5 + 1))
    assert len(jlS._shards(jlS.res)) >= 4
    assert jlS.launches_per_eval > jl1.launches_per_eval
    assert np.allclose(partsS[:3], parts1[:3], rtol=2e-6, atol=0)
    assert np.linalg.norm(gS - g1) <= 5e-6 * np.linalg.norm(g1)


def test_shards_cover_the_point_set_exactly():
    from pinn_depthestimation_b200.fused import JetLoss
    case, _ = cases.load("txyz")
    _, _, jl = _run(case, "fp32", 1000)
    for sp in (1, 7, 1000, jl.res.n - 1, jl.res.n, jl.res.n + 1):
        jl.shard_points = sp
        sh = jl._shards(jl.res)
        assert sh[0][0] == 0 and sum(c for _, c in sh) == jl.res.n
        assert all(a + c == b for (a, c), (b, _) in zip(sh, sh[1:]))
        assert max(c for _, c in sh) <= sp
