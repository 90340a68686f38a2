"""CPU tests (-m "not gpu"): host-side logic and the C-ABI surface.  No compute calls on a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch
from torch.optim import lbfgs as tl

from pinn_depthestimation_b200 import PassSpec, _cabi
from pinn_depthestimation_b200.lbfgs import cubic_interpolate, strong_wolfe

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_symbol_declared_in_the_header():
    hdr = open(os.path.join(ROOT, "include", "pinn_b200.h")).read()
    body = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)      # drop comments
    declared = set(re.findall(r"\b(pinn_[a-z0-9_]+)\s*\(", body))
    assert declared, "no prototypes found"
    assert declared == set(_cabi.SYMBOLS), declared ^ set(_cabi.SYMBOLS)
    lib = _cabi.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.pinn_version()


def test_desc_struct_layout_matches_header_sizes():
    # pinn_desc_t: 1 + 129 + 3 ints, 3 + 8 ints, int, 2 floats, int, 8 ints, 8 floats, 2 floats, int
    assert C.sizeof(_cabi.Desc) == 4 * (1 + 129 + 1 + 1 + 1 + 3 + 8 + 1 + 2 + 1 + 8 + 8 + 2 + 1)
    assert _cabi.EvalArgs.workspace_bytes.size == C.sizeof(C.c_size_t)


def test_param_count_and_validation_without_gpu():
    lib = _cabi.lib()
    s = PassSpec(layers=[2] + [20] * 100 + [3], kind="continuity_only", dirs={"x": 0, "y": 1},
                 fields={"h": 2, "U": 0, "V": 1}, target_cols=[0, 1])
    n = C.c_int64()
    assert lib.pinn_param_count(C.byref(s.to_desc()), C.byref(n)) == 0 and n.value == 41703 == s.n_params
    bad = PassSpec(layers=[2, 300, 3])          # wider than PINN_MAX_WIDTH
    assert lib.pinn_param_count(C.byref(bad.to_desc()), C.byref(n)) == 2
    assert b"width" in lib.pinn_last_error()
    d = s.to_desc()
    d.field_cols[1] = d.field_cols[0]           # duplicate field column
    assert lib.pinn_param_count(C.byref(d), C.byref(n)) == 1
    with pytest.raises(ValueError):
        PassSpec(layers=[2, 3], kind="nope").to_desc()


def test_compute_entry_points_fail_loudly_without_a_device():
    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    lib = _cabi.lib()
    s = PassSpec(layers=[2, 8, 3])
    nbytes = C.c_size_t()
    rc = lib.pinn_workspace_bytes(C.byref(s.to_desc()), 10, C.byref(nbytes))
    assert rc == 4 and b"CUDA error" in lib.pinn_last_error()      # PINN_E_CUDA, no fallback
    from pinn_depthestimation_b200.fused import JetLoss
    with pytest.raises(RuntimeError):
        JetLoss(s, torch.zeros(4, 2), None)


@pytest.mark.parametrize("seed", range(20))
def test_cubic_interpolate_matches_torch(seed):
    rs = np.random.RandomState(seed)
    x1, x2 = sorted(rs.uniform(0, 2, 2))
    if seed % 3 == 0:
        x1, x2 = x2, x1
    f1, f2, g1, g2 = rs.standard_normal(4)
    tt = lambda v: torch.tensor(v, dtype=torch.float64)
    ref = tl._cubic_interpolate(x1, f1, tt(g1), x2, f2, tt(g2))
    mine = cubic_interpolate(x1, f1, g1, x2, f2, g2)
    assert abs(float(ref) - mine) <= 1e-12 * max(1.0, abs(mine))
    ref = tl._cubic_interpolate(x1, f1, tt(g1), x2, f2, tt(g2), bounds=(0.3, 0.9))
    assert abs(float(ref) - cubic_interpolate(x1, f1, g1, x2, f2, g2, bounds=(0.3, 0.9))) <= 1e-12


def _objective(kind, dim, seed):
    rs = np.random.RandomState(seed)
    A = rs.standard_normal((dim, dim))
    A = A @ A.T + 0.1 * np.eye(dim)
    b = rs.standard_normal(dim)

    def fg(x):
        if kind == "quad":
            return 0.5 * x @ A @ x - b @ x, A @ x - b
        r = np.concatenate([10 * (x[1:] - x[:-1] ** 2), 1 - x[:-1]])       # Rosenbrock-like
        Jt = np.zeros((dim, 2 * (dim - 1)))
        for i in range(dim - 1):
            Jt[i + 1, i] += 10
            Jt[i, i] += -20 * x[i]
            Jt[i, dim - 1 + i] += -1
        return 0.5 * r @ r, Jt @ r
    return fg


@pytest.mark.parametrize("kind", ["quad", "rosen"])
@pytest.mark.parametrize("seed", range(8))
def test_strong_wolfe_follows_torch_branch_for_branch(kind, seed):
    dim = 6
    fg = _objective(kind, dim, seed)
    rs = np.random.RandomState(100 + seed)
    x0 = rs.standard_normal(dim)
    f0, g0 = fg(x0)
    d = -g0 * rs.uniform(0.5, 2.0, dim)
    t0 = float(rs.choice([1e-3, 0.1, 1.0, 5.0]))
    gtd = float(g0 @ d)

    def obj(x, t, dd):
        f, g = fg(x0 + float(t) * dd.numpy())
        return float(f), torch.from_numpy(g.copy())
    rf, rg, rt, rn = tl._strong_wolfe(obj, None, t0, torch.from_numpy(d), float(f0),
                                      torch.from_numpy(g0.copy()), torch.tensor(gtd, dtype=torch.float64), max_ls=25)

    def evaluate(t):
        f, g = fg(x0 + t * d)
        return float(f), g.copy(), float(g @ d)
    mf, mg, mt, mn = strong_wolfe(evaluate, lambda h: h.copy(), t0, float(f0), g0.copy(), gtd,
                                  float(np.abs(d).max()), max_ls=25)
    assert mn == rn
    assert abs(mt - float(rt)) <= 1e-9 * max(1.0, abs(mt))
    assert abs(mf - rf) <= 1e-9 * max(1.0, abs(mf))
    np.testing.assert_allclose(mg, rg.numpy(), rtol=1e-9, atol=1e-12)


def test_cubic_interpolate_collapsed_bracket_behaves_like_torch():
    """x1 == x2: torch's tensor arithmetic gives inf/nan and the result degenerates; no exception."""
    tt = lambda v: torch.tensor(v, dtype=torch.float64)
    for (x1, f1, g1, x2, f2, g2) in [(0.5, 1.0, -1.0, 0.5, 1.0, 0.3), (0.5, 1.0, -1.0, 0.5, 2.0, 0.3),
                                     (0.2, 3.0, 0.0, 0.2, 1.0, 0.0)]:
        ref = tl._cubic_interpolate(tt(x1), f1, tt(g1), tt(x2), f2, tt(g2))
        mine = cubic_interpolate(x1, f1, g1, x2, f2, g2)
        assert (mine != mine and float(ref) != float(ref)) or abs(float(ref) - mine) <= 1e-12
