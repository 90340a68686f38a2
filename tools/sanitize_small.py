"""Tiny workloads of the round-2 kernels for `compute-sanitizer --tool memcheck` (one tool per gpurun call):
third-order jet kernel, device L-BFGS (advance / dots / combine / trial), data-path kernels, FP32 jet kernel."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import jet_oracle as jo
from pinn_depthestimation_b200 import PassSpec, data as pdata
from pinn_depthestimation_b200.fused import JetLoss
from pinn_depthestimation_b200.lbfgs import LBFGS

dev = torch.device("cuda:0")
TXY = dict(dirs={"t": 0, "x": 1, "y": 2}, fields={"h": 0, "z": 1, "u": 2, "v": 3}, target_cols=[0, 1, 2, 3])
for layers, kind, n in (([3, 12, 12, 4], "Boussinesq", 13), ([4, 64, 33, 4], "Boussinesq", 9), ([3, 20, 20, 20, 4], "Boussinesq_simple", 37),
                        ([4, 7, 13, 5, 4], "Navier_Stokes", 67)):
    flat = torch.from_numpy(jo.make_params(layers, 1234, "tanh", np.float32)).to(dev)
    X, T = jo.make_points(n, layers[0], 4, seed=1)
    jl = JetLoss(PassSpec(layers=layers, kind=kind, **TXY), torch.from_numpy(X).to(dev), torch.from_numpy(T).to(dev))
    g = torch.empty_like(flat)
    parts = jl.loss_and_grad(flat, g)
    out = torch.empty(n, 4, device=dev)
    jl.loss(flat, out=out)
    torch.cuda.synchronize()
    print(kind, layers, parts.cpu().numpy())
    if kind == "Boussinesq_simple":
        q = torch.nn.Parameter(flat.clone())
        opt = LBFGS([q], lr=1, max_iter=9, max_eval=12, history_size=4, tolerance_grad=0, tolerance_change=0, line_search_fn="strong_wolfe")

        class C:
            def flat_loss_and_grad(self, fp, fg):
                return jl.loss_and_grad(fp, fg)
        opt.step(C())
        opt.step(C())
        torch.cuda.synchronize()
        print("lbfgs", opt.state[q]["n_iter"], opt.state[q]["func_evals"], opt.state[q]["loss"])
cfg = {"data_test": {"x_min": 25.0, "x_max": 33.0, "y_min": -13.0, "y_max": 13.0}}
rs = np.random.RandomState(0)
for n in (1, 300, 1025):
    cols = {"x": rs.uniform(25, 33, n).astype(np.float32), "y": rs.uniform(-13, 13, n).astype(np.float32), "t": rs.rand(n).astype(np.float32)}
    tr = {"U": rs.randn(n).astype(np.float32)}
    tr["U"][::5] = np.nan
    Xa, Ta, _ = pdata.assemble_points(cols, tr, cfg)
    torch.cuda.synchronize()
    print("assemble", n, tuple(Xa.shape))
print("SANITIZE_SMALL_DONE")
