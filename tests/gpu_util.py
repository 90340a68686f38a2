"""Helpers for the -m gpu tests: build PassSpecs from golden cases and run the CUDA path."""
import numpy as np
import torch

from pinn_depthestimation_b200 import PassSpec
from pinn_depthestimation_b200.fused import JetLoss
from tests import cases


def pass_specs(case, precision="fp32"):
    nt = len(case["target_cols"])
    tw = case.get("target_w", [1.0] * nt)
    common = dict(layers=case["layers"], activation=case["activation"],
                  w_fid=case.get("w_fid", 1.0), w_res=case.get("w_res", 1.0), precision=precision)
    if case["form"] == "single":
        return PassSpec(kind=case["kind"], dirs=case["dirs"], fields=case["fields"],
                        target_cols=case["target_cols"], target_w=tw, **common), None
    res = PassSpec(kind=case["kind"], dirs=case["dirs"], fields=case["fields"], **common)
    fid = PassSpec(kind="none", target_cols=case["target_cols"], target_w=tw, **common)
    return res, fid


def run_case(case, precision="fp32", n_override=None):
    dev = torch.device("cuda:0")
    sres, sfid = pass_specs(case, precision)
    flat, X, T, Xf, Tf = cases.data(case, np.float32)
    if n_override is not None:
        X, T = X[:n_override], (T[:n_override] if T is not None else None)
    t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    if sfid is None:
        jl = JetLoss(sres, t(X), t(T))
    else:
        jl = JetLoss(sres, t(X), None, fid=(sfid, t(Xf), t(Tf)))
    params = t(flat)
    grad = torch.full_like(params, float("nan"))   # must be overwritten, not accumulated into
    parts = jl.loss_and_grad(params, grad)
    torch.cuda.synchronize()
    return parts.cpu().numpy().astype(np.float64), grad.cpu().numpy().astype(np.float64), jl, params
