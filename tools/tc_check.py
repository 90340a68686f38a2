"""Diagnostic: tensor-core (tf32) path vs the fp64 oracle, per parameter block."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import jet_oracle as jo
from tests import cases
from tests.gpu_util import pass_specs
from pinn_depthestimation_b200.fused import JetLoss

name = sys.argv[1] if len(sys.argv) > 1 else "wide_nswe"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
case, z = cases.load(name)
dev = torch.device("cuda:0")
flat, X, T, _, _ = cases.data(case, np.float32)
X, T = X[:n], T[:n]
ospec, _ = cases.specs(case)
ref = jo.loss_and_grad(ospec, flat.astype(np.float64), X.astype(np.float64), T.astype(np.float64))
offs, total = jo.layer_offsets(case["layers"])
for prec in (sys.argv[3].split(",") if len(sys.argv) > 3 else ("fp32", "tf32")):
    spec, _ = pass_specs(case, prec)
    jl = JetLoss(spec, torch.from_numpy(X).to(dev), torch.from_numpy(T).to(dev))
    p = torch.from_numpy(flat).to(dev)
    g = torch.full_like(p, float("nan"))
    parts = jl.loss_and_grad(p, g)
    torch.cuda.synchronize()
    parts = parts.cpu().numpy().astype(np.float64)
    gg = g.cpu().numpy().astype(np.float64)
    print(f"[{prec}] loss {parts[2]:.8e} (ref {ref['loss']:.8e}) rel {abs(parts[2]-ref['loss'])/abs(ref['loss']):.2e}  "
          f"fid rel {abs(parts[0]-ref['fidelity'])/abs(ref['fidelity']):.2e}  res rel {abs(parts[1]-ref['residual'])/abs(ref['residual']):.2e}")
    print(f"[{prec}] grad rel-L2 {np.linalg.norm(gg-ref['grad'])/np.linalg.norm(ref['grad']):.3e}  finite={np.isfinite(gg).all()}")
    for l, (ow, ob) in enumerate(offs):
        nb = case["layers"][l + 1]
        ew = np.linalg.norm(gg[ow:ob] - ref["grad"][ow:ob]) / max(np.linalg.norm(ref["grad"][ow:ob]), 1e-30)
        eb = np.linalg.norm(gg[ob:ob + nb] - ref["grad"][ob:ob + nb]) / max(np.linalg.norm(ref["grad"][ob:ob + nb]), 1e-30)
        print(f"    layer {l}: dW rel {ew:.3e}   db rel {eb:.3e}")
    if prec != "fp32":
        l = len(offs) // 2
        ow, ob = offs[l]
        Hh = case["layers"][l]
        G = gg[ow:ob].reshape(Hh, Hh); R = ref["grad"][ow:ob].reshape(Hh, Hh)
        print("    |G|", np.linalg.norm(G), "|R|", np.linalg.norm(R), "max|G|", np.abs(G).max())
        def corr(a, b): return float((a * b).sum() / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-30))
        print("    corr(G,R)", corr(G, R), "corr(G,R^T)", corr(G, R.T))
        print("    G[0,:8]", G[0, :8]); print("    R[0,:8]", R[0, :8])
        print("    G[:8,0]", G[:8, 0]); print("    R[:8,0]", R[:8, 0])
        nz = np.abs(G) > 0
        print("    nonzero frac", nz.mean(), "rows with nz", nz.any(1).sum(), "cols with nz", nz.any(0).sum())
