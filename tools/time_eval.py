"""Quick device timing of one loss+grad evaluation (CUDA events), for development."""
import argparse
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import jet_oracle as jo
from pinn_depthestimation_b200 import PassSpec
from pinn_depthestimation_b200.fused import JetLoss

CFG = {
    "wide_nswe": dict(layers=[4] + [256] * 8 + [4], kind="Navier_Stokes", dirs={"t": 0, "x": 1, "y": 2},
                      fields={"h": 0, "z": 1, "u": 2, "v": 3}, target_cols=[0, 1, 2, 3]),
    "wide_cont": dict(layers=[2] + [256] * 8 + [3], kind="continuity_only", dirs={"x": 0, "y": 1},
                      fields={"U": 0, "V": 1, "h": 2}, target_cols=[0, 1]),
    "cmb_h": dict(layers=[2] + [20] * 100 + [3], kind="continuity_only", dirs={"x": 0, "y": 1},
                  fields={"U": 0, "V": 1, "h": 2}, target_cols=[0, 1]),
    "bouss": dict(layers=[3] + [20] * 20 + [4], kind="Boussinesq", dirs={"t": 0, "x": 1, "y": 2},
                  fields={"h": 0, "z": 1, "u": 2, "v": 3}, target_cols=[0, 1, 2, 3]),
    "bouss100": dict(layers=[3] + [20] * 100 + [4], kind="Boussinesq", dirs={"t": 0, "x": 1, "y": 2},
                     fields={"h": 0, "z": 1, "u": 2, "v": 3}, target_cols=[0, 1, 2, 3]),
    "bouss64": dict(layers=[3] + [64] * 8 + [4], kind="Boussinesq", dirs={"t": 0, "x": 1, "y": 2},
                    fields={"h": 0, "z": 1, "u": 2, "v": 3}, target_cols=[0, 1, 2, 3]),
    "txyz": dict(layers=[4] + [20] * 20 + [4], kind="Navier_Stokes", dirs={"t": 0, "x": 1, "y": 2},
                 fields={"h": 0, "z": 1, "u": 2, "v": 3}, target_cols=[0, 1, 2, 3]),
}

ap = argparse.ArgumentParser()
ap.add_argument("--cfg", default="wide_nswe")
ap.add_argument("--n", type=int, default=1 << 20)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--precision", default="fp32")
a = ap.parse_args()
c = CFG[a.cfg]
dev = torch.device("cuda:0")
spec = PassSpec(precision=a.precision, **c)
d, nt = c["layers"][0], len(c["target_cols"])
g = torch.Generator().manual_seed(1234)
X = (torch.rand(a.n, d, generator=g) * 2 - 1).to(dev)
T = (0.05 * torch.randn(a.n, nt, generator=g)).to(dev)
p = torch.from_numpy(jo.make_params(c["layers"], 1234)).to(dev)
gr = torch.empty_like(p)
jl = JetLoss(spec, X, T)
for _ in range(2):
    jl.loss_and_grad(p, gr)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    parts = jl.loss_and_grad(p, gr)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
k = len(c["dirs"])
sig = sum(c["layers"][i] * c["layers"][i + 1] for i in range(len(c["layers"]) - 1))
jets = 20 if c["kind"] == "Boussinesq" else 1 + k          # third-order Taylor jets in (t, x, y): 20 coefficients
flop = 6 * jets * sig * a.n
print(f"{a.cfg} N={a.n} {a.precision}: {ms:.3f} ms/eval  {a.n / ms * 1e3:.4e} pts/s  "
      f"{flop / ms / 1e9:.2f} TFLOP/s  loss={parts.cpu().numpy()}")
