"""Soak test of the tensor-core path: many back-to-back evaluations must reproduce the first one (races in the
mbarrier / slice / relay protocol would show up as a drifting gradient or a trapped launch)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import jet_oracle as jo
from pinn_depthestimation_b200 import PassSpec
from pinn_depthestimation_b200.fused import JetLoss

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 300
layers = [4] + [256] * 8 + [4]
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(5)
X = (torch.rand(n, 4, generator=g) * 2 - 1).to(dev)
T = (0.05 * torch.randn(n, 4, generator=g)).to(dev)
p = torch.from_numpy(jo.make_params(layers, 1234)).to(dev)
spec = PassSpec(layers=layers, kind="Navier_Stokes", dirs={"t": 0, "x": 1, "y": 2}, fields={"h": 0, "z": 1, "u": 2, "v": 3},
                target_cols=[0, 1, 2, 3], precision=(sys.argv[3] if len(sys.argv) > 3 else "tf32"))
jl = JetLoss(spec, X, T)
g0 = torch.empty_like(p)
parts0 = jl.loss_and_grad(p, g0).clone()
torch.cuda.synchronize()
worst = 0.0
gi = torch.empty_like(p)
t0 = time.time()
for i in range(iters):
    parts = jl.loss_and_grad(p, gi)
    rel = ((gi - g0).norm() / g0.norm()).item()
    dl = abs(parts[2].item() - parts0[2].item()) / abs(parts0[2].item())
    worst = max(worst, rel, dl)
    assert rel < 1e-5 and dl < 1e-6, (i, rel, dl)
print(f"soak ok: {iters} evaluations of {n} points in {time.time() - t0:.1f} s, worst deviation {worst:.2e}")
