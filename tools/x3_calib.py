"""Calibration of the accumulator-truncation compensation of the split-operand (tf32x3) mode (jet_tc.cu: x3_comp).

Sweeps PINN_X3_KAPPA and prints, per test net, the SIGNED relative loss error, the gradient-norm ratio minus one
(the systematic scale error), the angular gradient error (what no scale factor can remove) and the total rel-L2
gradient error, all against the float64 oracle.  The right kappa zeroes the scale error for every depth at once."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import jet_oracle as jo
from pinn_depthestimation_b200 import PassSpec
from pinn_depthestimation_b200.fused import JetLoss

NSWE = dict(kind="Navier_Stokes", dirs={"t": 0, "x": 1, "y": 2}, fields={"h": 0, "z": 1, "u": 2, "v": 3}, target_cols=[0, 1, 2, 3])
CONT = dict(kind="continuity_only", dirs={"x": 0, "y": 1}, fields={"U": 0, "V": 1, "h": 2}, target_cols=[0, 1])
NETS = [("nswe 256x8", [4] + [256] * 8 + [4], NSWE, 1000), ("nswe 256x5", [4] + [256] * 5 + [4], NSWE, 600),
        ("nswe 256x2", [4] + [256] * 2 + [4], NSWE, 600), ("cont 256x8", [2] + [256] * 8 + [3], CONT, 1000),
        ("cont 256x3", [2] + [256] * 3 + [3], CONT, 600)]
KAPPAS = [float(v) for v in (sys.argv[1].split(",") if len(sys.argv) > 1 else "0,2e-8,4e-8,6e-8,8e-8".split(","))]
dev = torch.device("cuda:0")
refs = {}
for name, layers, kw, n in NETS:
    flat = jo.make_params(layers, 1234, "tanh", np.float32)
    X, T = jo.make_points(n, layers[0], len(kw["target_cols"]), seed=1234)
    okw = dict(kw)
    okw["kind"] = {"Navier_Stokes": jo.NSWE, "continuity_only": jo.CONT_ONLY}[kw["kind"]]
    refs[name] = (flat, X, T, jo.loss_and_grad(dict(layers=layers, **okw), flat.astype(np.float64), X.astype(np.float64), T.astype(np.float64)))
for kappa in KAPPAS:
    os.environ["PINN_X3_KAPPA"] = repr(kappa)
    for name, layers, kw, n in NETS:
        flat, X, T, ref = refs[name]
        jl = JetLoss(PassSpec(layers=layers, precision="tf32x3", **kw), torch.from_numpy(X).to(dev), torch.from_numpy(T).to(dev))
        p = torch.from_numpy(flat).to(dev)
        g = torch.empty_like(p)
        parts = jl.loss_and_grad(p, g).cpu().numpy().astype(np.float64)
        gg = g.cpu().numpy().astype(np.float64)
        R = ref["grad"]
        scale = np.linalg.norm(gg) / np.linalg.norm(R) - 1
        cos = float(gg @ R / (np.linalg.norm(gg) * np.linalg.norm(R)))
        ang = np.sqrt(max(0.0, 2 * (1 - cos)))
        print(f"kappa {kappa:8.2e}  {name:11s} loss {(parts[2] - ref['loss']) / ref['loss']:+.2e}  fid {(parts[0] - ref['fidelity']) / ref['fidelity']:+.2e}  "
              f"res {(parts[1] - ref['residual']) / ref['residual']:+.2e}  |g| ratio-1 {scale:+.2e}  angular {ang:.2e}  rel-L2 {np.linalg.norm(gg - R) / np.linalg.norm(R):.2e}")
