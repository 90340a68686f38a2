"""Generate tests/golden/*.npz by running the REAL reference (dnn.py + physics.py imported from
/root/reference, unmodified) on deterministic weights and points.

Run in the build container only (the GPU box has no /root/reference):

    python -m oracle.make_golden

Each fixture stores the case description, the reference's loss parts and flat weight gradient in
float64 and float32, and -- for the large nets -- a strided subsample of the gradient plus
projections on fixed random vectors instead of the full vector.  Weights / inputs / targets are
NOT stored: they are re-created bit-exactly by oracle.jet_oracle.make_params / make_points
(numpy legacy RandomState streams).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

from . import jet_oracle as jo

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name -> case.  The shapes are the reference's configs at CPU-friendly N.
CASES = {
    # train_newmethod.py + config_CMB_h.json: [2]+[20]*100+[3], outputs (U,V,h), continuity_only
    "cmb_h": dict(layers=[2] + [20] * 100 + [3], activation="tanh", kind=jo.CONT_ONLY,
                  dirs={"x": 0, "y": 1}, fields={"U": 0, "V": 1, "h": 2},
                  target_cols=[0, 1], n=777, form="single"),
    # same, small/shallow (fast, exercises every code path in seconds)
    "cmb_h_small": dict(layers=[2] + [20] * 6 + [3], activation="tanh", kind=jo.CONT_ONLY,
                        dirs={"x": 0, "y": 1}, fields={"U": 0, "V": 1, "h": 2},
                        target_cols=[0, 1], n=301, form="single"),
    "ftemp_small": dict(layers=[2] + [12] * 3 + [3], activation="tanh", kind=jo.CONT_FTEMP,
                        dirs={"x": 0, "y": 1}, fields={"U": 0, "V": 1, "h": 2},
                        target_cols=[0, 1], n=130, form="single", w_fid=0.7, w_res=1.3),
    # train.py + config_CMB.json: [2]+[10]*10+[6], physics_equation, 12 fidelity pts + 243 grid pts
    "cmb": dict(layers=[2] + [10] * 10 + [6], activation="tanh", kind=jo.WAVE_AVG,
                dirs={"x": 0, "y": 1},
                fields={"h": 0, "U": 1, "V": 2, "eta_mean": 3, "Hrms": 4, "k": 5},
                target_cols=[0, 1, 2, 3, 4, 5], target_w=[1.0, 0.5, 2.0, 1.0, 0.25, 1.5],
                n=243, n_fid=12, form="two_pass"),
    # config.json shape: [5]+[20]*100+[4], (t,x,y,u,v) inputs, Navier_Stokes on (h,z,u,v)
    "config_json": dict(layers=[5] + [20] * 100 + [4], activation="tanh", kind=jo.NSWE,
                        dirs={"t": 0, "x": 1, "y": 2}, fields={"h": 0, "z": 1, "u": 2, "v": 3},
                        target_cols=[0, 1, 2, 3], n=333, n_fid=96, form="two_pass"),
    # config_txyz.json shape: [4]+[20]*20+[4], z input not differentiated
    "txyz": dict(layers=[4] + [20] * 20 + [4], activation="tanh", kind=jo.NSWE,
                 dirs={"t": 0, "x": 1, "y": 2}, fields={"h": 0, "z": 1, "u": 2, "v": 3},
                 target_cols=[0, 1, 2, 3], n=500, form="single"),
    # BASELINE.json configs[4]: width-256 x 8 hidden layers
    "wide_nswe": dict(layers=[4] + [256] * 8 + [4], activation="tanh", kind=jo.NSWE,
                      dirs={"t": 0, "x": 1, "y": 2}, fields={"h": 0, "z": 1, "u": 2, "v": 3},
                      target_cols=[0, 1, 2, 3], n=1000, form="single"),
    "wide_cont": dict(layers=[2] + [256] * 8 + [3], activation="tanh", kind=jo.CONT_ONLY,
                      dirs={"x": 0, "y": 1}, fields={"U": 0, "V": 1, "h": 2},
                      target_cols=[0, 1], n=1000, form="single"),
    # init_type='kaiming' => LeakyReLU(0.01) (dnn.py:20-21)
    "leaky": dict(layers=[2] + [16] * 4 + [3], activation="leaky_relu", kind=jo.CONT_FTEMP,
                  dirs={"x": 0, "y": 1}, fields={"U": 0, "V": 1, "h": 2},
                  target_cols=[0, 1], n=200, form="single"),
    # 256-wide nets for the residuals the round-1 tensor-core tests did not cover: physics_equation (divisions by eta+h,
    # sinh; the last bias keeps the depth h and eta positive like the data, so 1/(eta+h) stays off its pole) and
    # continuity_ftemp
    "wide_wave": dict(layers=[2] + [256] * 4 + [6], activation="tanh", kind=jo.WAVE_AVG,
                      dirs={"x": 0, "y": 1},
                      fields={"h": 0, "U": 1, "V": 2, "eta_mean": 3, "Hrms": 4, "k": 5},
                      target_cols=[0, 1, 2, 3, 4, 5], target_w=[1.0, 0.5, 2.0, 1.0, 0.25, 1.5],
                      last_bias=[2.0, 0.05, -0.03, 0.6, 0.4, 0.9], n=500, form="single"),
    "wide_ftemp": dict(layers=[2] + [256] * 3 + [3], activation="tanh", kind=jo.CONT_FTEMP,
                       dirs={"x": 0, "y": 1}, fields={"U": 0, "V": 1, "h": 2},
                       target_cols=[0, 1], n=500, form="single", w_fid=0.7, w_res=1.3),
    # physics_equation with the wavenumber output k identically 0: sinh(2kh) = 0, so the (zero-E) radiation-stress terms
    # are 0 * (0/0) = NaN and the loss is NaN exactly as in the reference (physics.py:106-108)
    "cmb_nan": dict(layers=[2] + [10] * 3 + [6], activation="tanh", kind=jo.WAVE_AVG,
                    dirs={"x": 0, "y": 1},
                    fields={"h": 0, "U": 1, "V": 2, "eta_mean": 3, "Hrms": 4, "k": 5},
                    target_cols=[0, 1, 2, 3, 4, 5], zero_out_cols=[5], n=64, n_fid=12, form="two_pass"),
    # the historical physics_functions residuals (bytecode only in the reference; the "reference run" for these three is
    # oracle/boussinesq_oracle.py = the decompiled functions re-typed, under torch autograd: see run_reference)
    "bouss": dict(layers=[3] + [20] * 6 + [4], activation="tanh", kind="Boussinesq",
                  dirs={"t": 0, "x": 1, "y": 2}, fields={"h": 0, "z": 1, "u": 2, "v": 3},
                  target_cols=[0, 1, 2, 3], n=90, form="single"),
    "bouss_wide": dict(layers=[4] + [64] * 3 + [4], activation="tanh", kind="Boussinesq",
                       dirs={"t": 0, "x": 1, "y": 2}, fields={"h": 0, "z": 1, "u": 2, "v": 3},
                       target_cols=[2, 3], n=41, form="single", w_fid=0.5, w_res=2.0),
    "bouss_simple": dict(layers=[3] + [16] * 4 + [4], activation="tanh", kind="Boussinesq_simple",
                         dirs={"t": 0, "x": 1, "y": 2}, fields={"h": 0, "z": 1, "u": 2, "v": 3},
                         target_cols=[0, 1, 2, 3], n=150, form="single"),
    # odd widths (not multiples of 4) and a ragged tile
    "ragged": dict(layers=[3, 7, 13, 5, 4], activation="tanh", kind=jo.NSWE,
                   dirs={"t": 0, "x": 1, "y": 2}, fields={"h": 3, "z": 0, "u": 2, "v": 1},
                   target_cols=[2, 0], n=67, form="single"),
}


def _ref_modules():
    sys.path.insert(0, REF)
    import dnn as ref_dnn          # noqa: E402  (the reference's own file)
    import physics as ref_physics  # noqa: E402
    sys.path.pop(0)
    return ref_dnn, ref_physics


def _load_flat(model, flat):
    o = 0
    with torch.no_grad():
        for p in model.parameters():
            n = p.numel()
            p.copy_(torch.from_numpy(flat[o:o + n]).view_as(p))
            o += n
    assert o == flat.size


def _flat_grad(model):
    return torch.cat([p.grad.reshape(-1) for p in model.parameters()]).detach().numpy().copy()


def run_decompiled(case, dtype):
    """Boussinesq / Boussinesq_simple: the reference module exists only as Python-3.8 bytecode and cannot be imported here;
    the decompiled functions (oracle/boussinesq_oracle.py) are run under torch autograd instead."""
    from . import boussinesq_oracle as bo
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    flat = jo.make_case_params(case, dtype)
    d, nt = case["layers"][0], len(case["target_cols"])
    X, T = jo.make_points(case["n"], d, nt, seed=1234)
    spec = dict(layers=case["layers"], activation=case["activation"], kind=case["kind"], dirs=case["dirs"],
                target_cols=case["target_cols"], w_fid=case.get("w_fid", 1.0), w_res=case.get("w_res", 1.0))
    r = bo.loss_and_grad(spec, torch.from_numpy(flat).to(tdt), torch.from_numpy(X.astype(dtype)).to(tdt),
                         torch.from_numpy(T.astype(dtype)).to(tdt))
    return dict(loss=float(r["loss"]), fidelity=float(r["fidelity"]), residual=float(r["residual"]),
                grad=r["grad"].numpy().copy(), out=r["out"].numpy().copy())


def run_reference(case, dtype):
    """Evaluate loss + gradient exactly the way the reference's loss_func does."""
    if case["kind"] in ("Boussinesq", "Boussinesq_simple"):
        return run_decompiled(case, dtype)
    ref_dnn, ref_physics = _ref_modules()
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    layers = case["layers"]
    init = "xavier" if case["activation"] == "tanh" else "kaiming"
    model = ref_dnn.DNN(layers, 0.0, init).to(tdt)
    flat = jo.make_case_params(case, dtype)
    _load_flat(model, flat)
    model.train()
    d = layers[0]
    nt = len(case["target_cols"])
    X, _ = jo.make_points(case["n"], d, 0, seed=1234)
    w_fid, w_res = case.get("w_fid", 1.0), case.get("w_res", 1.0)
    tw = case.get("target_w", [1.0] * nt)

    # input columns built like train_newmethod.py:77-79 (float64 leaf -> cast)
    col_of = {c: n for n, c in case["dirs"].items()}
    cols, named = [], {}
    for c in range(d):
        t = torch.tensor(X[:, c:c + 1].astype(np.float64), requires_grad=(c in col_of)).to(tdt)
        cols.append(t)
        if c in col_of:
            named[col_of[c]] = t
    pred = model(torch.cat(cols, dim=-1))
    f = {n: pred[:, c:c + 1] for n, c in case["fields"].items()}
    kind = case["kind"]
    if kind == jo.CONT_ONLY:
        res = ref_physics.continuity_only(named["x"], named["y"], f["h"], f["U"], f["V"])
    elif kind == jo.CONT_FTEMP:
        res = ref_physics.continuity_ftemp(named["x"], named["y"], f["h"], f["U"], f["V"])
    elif kind == jo.NSWE:
        res = ref_physics.Navier_Stokes(named["t"], named["x"], named["y"],
                                        f["h"], f["z"], f["u"], f["v"])
    else:
        res = ref_physics.physics_equation(named["x"], named["y"], f["h"], f["U"], f["V"],
                                           f["eta_mean"], f["Hrms"], f["k"])
    if case["form"] == "single":
        _, T = jo.make_points(case["n"], d, nt, seed=1234)
        fid = 0
        for i, c in enumerate(case["target_cols"]):
            # train_newmethod.py:129-133
            fid = fid + tw[i] * torch.nn.functional.mse_loss(
                pred[:, c:c + 1], torch.tensor(T[:, i:i + 1].astype(np.float64)).to(tdt))
        out_pred = pred
    else:
        Xf, Tf = jo.make_points(case["n_fid"], d, nt, seed=4321)
        pf = model(torch.tensor(Xf.astype(np.float64)).to(tdt))
        fid = 0
        for i, c in enumerate(case["target_cols"]):
            # train.py:136-141
            fid = fid + tw[i] * torch.mean(
                (torch.tensor(Tf[:, i:i + 1].astype(np.float64)).to(tdt) - pf[:, c:c + 1]) ** 2)
        out_pred = pred
    loss = w_fid * fid + w_res * res
    model.zero_grad()
    loss.backward()
    return dict(loss=loss.item(), fidelity=float(fid), residual=res.item(),
                grad=_flat_grad(model), out=out_pred.detach().numpy().copy())


def pack(name, case):
    r64 = run_reference(case, np.float64)
    r32 = run_reference(case, np.float32)
    P = r64["grad"].size
    rs = np.random.RandomState(99)
    probes = rs.standard_normal((4, P))
    rec = dict(
        case=json.dumps(case),
        loss64=r64["loss"], fidelity64=r64["fidelity"], residual64=r64["residual"],
        loss32=r32["loss"], fidelity32=r32["fidelity"], residual32=r32["residual"],
        grad_norm64=float(np.linalg.norm(r64["grad"])),
        grad_proj64=probes @ r64["grad"],
        grad_proj32=probes @ r32["grad"].astype(np.float64),
        out64_head=r64["out"][:16].copy(),
    )
    if P <= 50000:
        rec["grad64"] = r64["grad"]
        rec["grad32"] = r32["grad"]
    else:
        rec["grad_stride"] = 61
        rec["grad64_sub"] = r64["grad"][::61].copy()
        rec["grad32_sub"] = r32["grad"][::61].copy()
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **rec)
    noise = np.linalg.norm(r32["grad"] - r64["grad"]) / np.linalg.norm(r64["grad"])
    print(f"{name:14s} P={P:7d} loss64={r64['loss']:.9e} loss32={r32['loss']:.9e} "
          f"|g|={rec['grad_norm64']:.4e} ref fp32-vs-fp64 grad rel-L2={noise:.2e}")


def main():
    torch.set_num_threads(8)
    only = sys.argv[1:]          # python -m oracle.make_golden [case ...]: regenerate just these
    for name, case in CASES.items():
        if not only or name in only:
            pack(name, case)


if __name__ == "__main__":
    main()
