"""Oracle for the reference's historical fully-nonlinear Boussinesq residual -- TEST INFRASTRUCTURE ONLY.

The reference ships `physics_functions` only as stale bytecode (__pycache__/physics_functions.cpython-38.pyc, source
/home/reza/Projects/AI/main/physics_functions.py, not in the tree).  The two functions below are the
decompilation of that bytecode (tools/pyc38_decompile.py reads the 3.8 marshal stream by hand and replays the bytecode
on a symbolic stack), re-typed statement by statement with the source line of each `def` kept: `Boussinesq_simple`
(:18) and `Boussinesq` (:55).  They are run under torch autograd exactly as the reference would run them
(`compute_gradient` = autograd.grad(..., create_graph=True), :5-15), which needs input derivatives up to THIRD order --
the only place in the reference where the "second-order residual terms" of BASELINE.json configs[2] exist.

Parity status: pinned to the decompiled bytecode only -- the module cannot be imported under Python 3.12, the reference
has no caller for it and no fixture; tests/golden/bouss*.npz are produced by THIS file in float64 (oracle/make_golden.py).
"""
from __future__ import annotations

import torch

from . import autograd_port as ap


def compute_gradient(pred, var):
    """physics_functions.py:5-15 (same body as physics.py:6-15)."""
    return torch.autograd.grad(pred, var, grad_outputs=torch.ones_like(pred), retain_graph=True, create_graph=True)[0]


def Boussinesq_simple(output, t, x, y, device=None):
    """physics_functions.py:18-52."""
    h_pred, z_pred, u_pred, v_pred = output[:, 0:1], output[:, 1:2], output[:, 2:3], output[:, 3:4]
    u_t, u_x, u_y = compute_gradient(u_pred, t), compute_gradient(u_pred, x), compute_gradient(u_pred, y)
    v_t, v_x, v_y = compute_gradient(v_pred, t), compute_gradient(v_pred, x), compute_gradient(v_pred, y)
    z_t, z_x, z_y = compute_gradient(z_pred, t), compute_gradient(z_pred, x), compute_gradient(z_pred, y)
    hu = h_pred * u_pred
    hv = h_pred * v_pred
    hu_x = compute_gradient(hu, x)
    hv_y = compute_gradient(hv, y)
    f_cont = z_t + hu_x + hv_y
    f_momx = u_t + u_pred * u_x + v_pred * u_y + 9.81 * z_x
    f_momy = v_t + u_pred * v_x + v_pred * v_y + 9.81 * z_y
    return torch.mean(f_cont ** 2) + torch.mean(f_momx ** 2) + torch.mean(f_momy ** 2)


def Boussinesq(output, t, x, y, device=None):
    """physics_functions.py:55-130."""
    h_pred, z_pred, u_pred, v_pred = output[:, 0:1], output[:, 1:2], output[:, 2:3], output[:, 3:4]
    u_t, u_x, u_y = compute_gradient(u_pred, t), compute_gradient(u_pred, x), compute_gradient(u_pred, y)
    v_t, v_x, v_y = compute_gradient(v_pred, t), compute_gradient(v_pred, x), compute_gradient(v_pred, y)
    z_t, z_x, z_y = compute_gradient(z_pred, t), compute_gradient(z_pred, x), compute_gradient(z_pred, y)
    hu = h_pred * u_pred
    hv = h_pred * v_pred
    hu_x = compute_gradient(hu, x)
    hv_y = compute_gradient(hv, y)
    A = hu_x + hv_y
    B = u_x + v_y
    A_t, A_x, A_y = compute_gradient(A, t), compute_gradient(A, x), compute_gradient(A, y)
    B_t, B_x, B_y = compute_gradient(B, t), compute_gradient(B, x), compute_gradient(B, y)
    z_alpha = (-0.53) * h_pred + 0.47 * z_pred
    z_alpha_x = compute_gradient(z_alpha, x)
    z_alpha_y = compute_gradient(z_alpha, y)
    temp1 = (z_alpha ** 2) / 2 - 0.16666666666666666 * (h_pred ** 2 - h_pred * z_pred + z_pred ** 2)
    temp2 = z_alpha + 0.5 * (h_pred - z_pred)
    u_2 = temp1 * B_x + temp2 * A_x
    v_2 = temp1 * B_y + temp2 * A_y
    u_surface = u_pred + u_2
    v_surface = v_pred + v_2
    H = h_pred + z_pred
    Hu_surface = H * u_surface
    Hv_surface = H * v_surface
    Hu_x = compute_gradient(Hu_surface, x)
    Hv_y = compute_gradient(Hv_surface, y)
    V1Ax = ((z_alpha ** 2) / 2) * B_x + z_alpha * A_x
    V1Ax_t = compute_gradient(V1Ax, t)
    V1Ay = ((z_alpha ** 2) / 2) * B_y + z_alpha * A_y
    V1Ay_t = compute_gradient(V1Ay, t)
    V1B = ((z_pred ** 2) / 2) * B_t + z_pred * A_t
    V1Bx = compute_gradient(V1B, x)
    V1By = compute_gradient(V1B, y)
    V1x = V1Ax_t - V1Bx
    V1y = V1Ay_t - V1By
    V2 = ((z_alpha - z_pred) * (u_pred * A_x + v_pred * A_y)
          + (0.5 * (z_alpha ** 2 - z_pred ** 2)) * (u_pred * B_x + v_pred * B_y)
          + 0.5 * (A + z_pred * B) ** 2)
    V2x = compute_gradient(V2, x)
    V2y = compute_gradient(V2, y)
    omega0 = v_x - u_y
    omega2 = z_alpha_x * (A_y + z_alpha * B_y) - z_alpha_y * (A_x + z_alpha * B_x)
    V3x = (-omega0) * v_2 - omega2 * v_pred
    V3y = omega0 * u_2 + omega2 * u_pred
    f_cont = z_t + Hu_x + Hv_y
    f_momx = u_t + u_pred * u_x + v_pred * u_y + 9.81 * z_x + V1x + V2x + V3x
    f_momy = v_t + u_pred * v_x + v_pred * v_y + 9.81 * z_y + V1y + V2y + V3y
    return torch.mean(f_cont ** 2) + torch.mean(f_momx ** 2) + torch.mean(f_momy ** 2)


KINDS = {"Boussinesq": Boussinesq, "Boussinesq_simple": Boussinesq_simple}


def loss_and_grad(spec, flat, X, targets=None):
    """Same contract as autograd_port.loss_and_grad for kind in KINDS: outputs (h, z, u, v) are columns 0..3 of the
    network output (the functions slice `output` themselves), dirs = {'t','x','y'} -> input columns."""
    layers, kind = spec["layers"], spec["kind"]
    p = flat.detach().clone().requires_grad_(True)
    cols, by = [], {}
    col_of = {c: n for n, c in spec["dirs"].items()}
    for c in range(X.shape[1]):
        tcol = X[:, c:c + 1].detach().clone()
        if c in col_of:
            tcol.requires_grad_(True)
            by[col_of[c]] = tcol
        cols.append(tcol)
    out = ap.mlp(p, layers, torch.cat(cols, dim=-1), spec.get("activation", "tanh"))
    fid = torch.zeros((), dtype=flat.dtype)
    if targets is not None:
        tw = spec.get("target_w", [1.0] * len(spec["target_cols"]))
        for i, c in enumerate(spec["target_cols"]):
            fid = fid + tw[i] * torch.mean((targets[:, i:i + 1] - out[:, c:c + 1]) ** 2)
    res = KINDS[kind](out, by["t"], by["x"], by["y"])
    loss = spec.get("w_fid", 1.0) * fid + spec.get("w_res", 1.0) * res
    loss.backward()
    return {"loss": loss.detach(), "fidelity": fid.detach(), "residual": res.detach(), "grad": p.grad.detach(),
            "out": out.detach()}
