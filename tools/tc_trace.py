"""Event trace of one tile pair of the tensor-core kernel (library built with -DPINN_TC_TRACE=<tile pair index>)."""
import argparse, ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import jet_oracle as jo
from pinn_depthestimation_b200 import PassSpec, _cabi
from pinn_depthestimation_b200.fused import JetLoss

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=606208)
ap.add_argument("--precision", default="tf32")
ap.add_argument("--out", default="gpurun_out/tc_trace.txt")
a = ap.parse_args()
c = dict(layers=[4] + [256] * 8 + [4], kind="Navier_Stokes", dirs={"t": 0, "x": 1, "y": 2},
         fields={"h": 0, "z": 1, "u": 2, "v": 3}, target_cols=[0, 1, 2, 3])
dev = torch.device("cuda:0")
spec = PassSpec(precision=a.precision, **c)
g = torch.Generator().manual_seed(1234)
X = (torch.rand(a.n, 4, generator=g) * 2 - 1).to(dev)
T = (0.05 * torch.randn(a.n, 4, generator=g)).to(dev)
p = torch.from_numpy(jo.make_params(c["layers"], 1234)).to(dev)
gr = torch.empty_like(p)
jl = JetLoss(spec, X, T)
lib = _cabi.lib()
buf = (ctypes.c_longlong * 16384)()
lib.pinn_debug_trace.argtypes = [ctypes.POINTER(ctypes.c_longlong), ctypes.c_int]
for rep in range(3):
    jl.loss_and_grad(p, gr)
    torch.cuda.synchronize()
    n = lib.pinn_debug_trace(buf, 8192)
ev = sorted((buf[2 * i + 1], buf[2 * i]) for i in range(n) if buf[2 * i] != 0)
t0 = ev[0][0]
with open(a.out, "w") as f:
    for t, tag in ev:
        f.write(f"{t - t0:8d} {tag}\n")
print(len(ev), "events ->", a.out)
