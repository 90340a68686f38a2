"""Shared helpers: load a golden case (tests/golden/*.npz, made by oracle/make_golden.py from the
real reference) and rebuild its deterministic weights / points."""
import json
import os

import numpy as np

from oracle import jet_oracle as jo

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAN = ["cmb_nan"]            # the reference's loss is NaN by construction (physics.py:106-108 with k == 0)
# the historical physics_functions residuals: golden = the decompiled bytecode re-typed (oracle/boussinesq_oracle.py)
BOUSS = ["bouss", "bouss_wide", "bouss_simple"]
ALL = sorted(f[:-4] for f in os.listdir(GOLDEN)
             if f.endswith(".npz") and f[:-4] not in NAN + BOUSS and not f.startswith(("curve_", "ref_")))
SMALL = [n for n in ALL if not n.startswith("wide")]


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    case = json.loads(str(z["case"]))
    return case, z


def specs(case):
    """-> (spec_res, spec_fid or None). For 'single' form one spec carries both terms."""
    base = dict(layers=case["layers"], activation=case["activation"],
                w_fid=case.get("w_fid", 1.0), w_res=case.get("w_res", 1.0))
    nt = len(case["target_cols"])
    tw = case.get("target_w", [1.0] * nt)
    res = dict(base, kind=case["kind"], dirs=case["dirs"], fields=case["fields"],
               target_cols=case["target_cols"], target_w=tw)
    if case["form"] == "single":
        return res, None
    fid = dict(base, kind=jo.NONE, target_cols=case["target_cols"], target_w=tw)
    return res, fid


def data(case, dtype=np.float32):
    d = case["layers"][0]
    nt = len(case["target_cols"])
    flat = jo.make_case_params(case, dtype)
    if case["form"] == "single":
        X, T = jo.make_points(case["n"], d, nt, seed=1234)
        return flat, X, T, None, None
    X, _ = jo.make_points(case["n"], d, 0, seed=1234)
    Xf, Tf = jo.make_points(case["n_fid"], d, nt, seed=4321)
    return flat, X, None, Xf, Tf


def golden_grad_check(z, grad, tag="64"):
    """relative L2 error of `grad` (full vector) against the golden (full or strided)."""
    grad = np.asarray(grad, dtype=np.float64)
    if f"grad{tag}" in z.files:
        ref = z[f"grad{tag}"].astype(np.float64)
        return np.linalg.norm(grad - ref) / np.linalg.norm(ref)
    st = int(z["grad_stride"])
    ref = z[f"grad{tag}_sub"].astype(np.float64)
    e_sub = np.linalg.norm(grad[::st] - ref) / np.linalg.norm(ref)
    rs = np.random.RandomState(99)
    probes = rs.standard_normal((4, grad.size))
    proj = probes @ grad
    e_proj = np.max(np.abs(proj - z[f"grad_proj{tag}"])) / float(z["grad_norm64"])
    e_norm = abs(np.linalg.norm(grad) - float(z["grad_norm64"])) / float(z["grad_norm64"])
    return max(e_sub, e_proj, e_norm)
