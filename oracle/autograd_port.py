"""Torch-autograd restatement of the reference's CPU path -- TEST / BASELINE INFRASTRUCTURE ONLY.

This is the *same algorithm* the reference runs (reverse-mode autograd with create_graph=True for
every input derivative, then loss.backward()), re-typed from its description so that it can travel
to the GPU box where /root/reference does not exist.  It is what bench.py times as the CPU
baseline (`cpu_baseline.kind == "port"`, `--impl reference`).  It is validated against the real
reference in tests/test_oracle_golden.py through the committed golden vectors.

Follows: dnn.py:31-38,54-55 (MLP), physics.py:6-15 (input derivative), physics.py:18-120
(residuals), train_newmethod.py:120-159 and train.py:128-157 (loss assembly).
"""
from __future__ import annotations

import torch

from . import jet_oracle as jo


def mlp(flat, layers, x, activation="tanh"):
    """Functional MLP over the flat parameter vector (parameter order of dnn.py:31-34)."""
    o = 0
    n_lin = len(layers) - 1
    a = x
    for i in range(n_lin):
        fin, fout = layers[i], layers[i + 1]
        W = flat[o:o + fin * fout].view(fout, fin)
        o += fin * fout
        b = flat[o:o + fout]
        o += fout
        a = torch.addmm(b, a, W.t())
        if i < n_lin - 1:
            a = torch.tanh(a) if activation == "tanh" else torch.nn.functional.leaky_relu(a, 0.01)
    return a


def ddx(pred, var):
    """per-point d pred / d var, differentiable (physics.py:6-15)."""
    return torch.autograd.grad(pred, var, grad_outputs=torch.ones_like(pred),
                               retain_graph=True, create_graph=True)[0]


def residual(kind, cols, f):
    """cols: dict dir -> [N,1] leaf column; f: dict field -> [N,1] prediction slice."""
    if kind in (jo.CONT_ONLY, jo.CONT_FTEMP):
        x, y = cols["x"], cols["y"]
        fc = ddx(f["h"] * f["U"], x) + ddx(f["h"] * f["V"], y)
        loss = torch.mean(fc ** 2)
        if kind == jo.CONT_ONLY:
            sel = torch.where(x < 25.5)
            loss = loss + torch.mean((f["h"][sel] - 0.75) ** 2)
        return loss
    if kind == jo.NSWE:
        t, x, y = cols["t"], cols["x"], cols["y"]
        h, z, u, v = f["h"], f["z"], f["u"], f["v"]
        u_t, u_x, u_y = ddx(u, t), ddx(u, x), ddx(u, y)
        v_t, v_x, v_y = ddx(v, t), ddx(v, x), ddx(v, y)
        z_t, z_x, z_y = ddx(z, t), ddx(z, x), ddx(z, y)
        H = h + z
        H_x, H_y = ddx(H, x), ddx(H, y)
        Hu_x, Hv_y = ddx(H * u, x), ddx(H * v, y)
        cb = 3.0 / 16.0 * 9.81 * 0.78 ** 2
        fc = z_t + Hu_x + Hv_y
        fx = u_t + u * u_x + v * u_y + 9.81 * z_x + 0 + cb * H_x * H
        fy = v_t + u * v_x + v * v_y + 9.81 * z_y + 0 + cb * H_y * H
        return torch.mean(fc ** 2) + torch.mean(fx ** 2) + torch.mean(fy ** 2)
    if kind == jo.WAVE_AVG:
        x, y = cols["x"], cols["y"]
        h, U, V, eta, Hrms, k = (f[n] for n in ("h", "U", "V", "eta_mean", "Hrms", "k"))
        U_x, U_y, V_x, V_y = ddx(U, x), ddx(U, y), ddx(V, x), ddx(V, y)
        e_x, e_y = ddx(eta, x), ddx(eta, y)
        rho, Cd, g = 1025, 0.002, 9.81
        tbx, tby = rho * Cd * U * abs(U), rho * Cd * V * abs(V)
        E = 1 / 8 ** rho * g * Hrms ** 2          # == 0.0 * Hrms^2, kept as written (physics.py:106)
        Sxx = E * (2 * k * h / torch.sinh(2 * k * h) + 0.5)
        Syy = E * (1 * k * h / torch.sinh(2 * k * h) + 0.0)
        Sxx_x, Syy_y = ddx(Sxx, x), ddx(Syy, y)
        inv = 1 / (rho * (eta + h))
        fc = U_x + V_y
        fx = U * U_x + V * U_y + g * e_x + inv * (Sxx_x + 0) + inv * tbx
        fy = U * V_x + V * V_y + g * e_y + inv * (0 + Syy_y) + inv * tby
        return torch.mean(fc ** 2) + torch.mean(fx ** 2) + torch.mean(fy ** 2)
    raise ValueError(kind)


def loss_and_grad(spec, flat, X, targets=None):
    """Same contract as jet_oracle.loss_and_grad, computed the reference's way (autograd)."""
    layers = spec["layers"]
    kind = spec.get("kind", jo.NONE)
    act = spec.get("activation", "tanh")
    p = flat.detach().clone().requires_grad_(True)
    d = X.shape[1]
    dirs = spec.get("dirs", {}) if kind != jo.NONE else {}
    col_of = {c: n for n, c in dirs.items()}
    cols, by_name = [], {}
    for c in range(d):
        t = X[:, c:c + 1].detach().clone()
        if c in col_of:
            t.requires_grad_(True)
            by_name[col_of[c]] = t
        cols.append(t)
    out = mlp(p, layers, torch.cat(cols, dim=-1), act)
    fid = torch.zeros((), dtype=flat.dtype)
    if targets is not None:
        tw = spec.get("target_w", [1.0] * len(spec["target_cols"]))
        for i, c in enumerate(spec["target_cols"]):
            fid = fid + tw[i] * torch.mean((targets[:, i:i + 1] - out[:, c:c + 1]) ** 2)
    res = torch.zeros((), dtype=flat.dtype)
    if kind != jo.NONE:
        f = {n: out[:, c:c + 1] for n, c in spec["fields"].items()}
        if "mask_col" in spec and kind == jo.CONT_ONLY and spec["mask_col"] != dirs["x"]:
            raise NotImplementedError("mask column other than x")
        res = residual(kind, by_name, f)
    loss = spec.get("w_fid", 1.0) * fid + spec.get("w_res", 1.0) * res
    loss.backward()
    return {"loss": loss.detach(), "fidelity": fid.detach(), "residual": res.detach(),
            "grad": p.grad.detach(), "out": out.detach()}
