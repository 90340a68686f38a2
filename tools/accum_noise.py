"""Run-to-run and kernel-to-kernel gradient differences at bench scale (accumulation-order noise of the FP32 atomics)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import jet_oracle as jo
from pinn_depthestimation_b200 import PassSpec
from pinn_depthestimation_b200.fused import JetLoss

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
c = dict(layers=[4] + [256] * 8 + [4], kind="Navier_Stokes", dirs={"t": 0, "x": 1, "y": 2},
         fields={"h": 0, "z": 1, "u": 2, "v": 3}, target_cols=[0, 1, 2, 3])
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1234)
X = (torch.rand(n, 4, generator=g) * 2 - 1).to(dev)
T = (0.05 * torch.randn(n, 4, generator=g)).to(dev)
p = torch.from_numpy(jo.make_params(c["layers"], 1234)).to(dev)
grads = {}
for prec in ("tf32x3", "fp32"):
    jl = JetLoss(PassSpec(precision=prec, **c), X, T)
    for rep in range(2):
        gr = torch.empty_like(p)
        parts = jl.loss_and_grad(p, gr)
        torch.cuda.synchronize()
        grads[(prec, rep)] = gr.double()
        print(prec, rep, parts.cpu().numpy())
    del jl
rel = lambda a, b: float((a - b).norm() / b.norm())
print("x3 run0 vs run1   ", rel(grads[("tf32x3", 0)], grads[("tf32x3", 1)]))
print("fp32 run0 vs run1 ", rel(grads[("fp32", 0)], grads[("fp32", 1)]))
print("x3 vs fp32        ", rel(grads[("tf32x3", 0)], grads[("fp32", 0)]))
# chunked evaluation: 16 shards of n/16 points summed in float64 (each shard's accumulation is 16x shorter)
for prec in ("tf32x3", "fp32"):
    acc = torch.zeros_like(p, dtype=torch.float64)
    m = n // 16
    for k in range(16):
        jl = JetLoss(PassSpec(precision=prec, **c), X[k * m:(k + 1) * m], T[k * m:(k + 1) * m])
        gr = torch.empty_like(p)
        jl.loss_and_grad(p, gr)
        acc += gr.double() / 16
        del jl
    grads[(prec, "sh")] = acc
print("x3 sharded vs fp32 sharded", rel(grads[("tf32x3", "sh")], grads[("fp32", "sh")]))
print("x3 one pass vs x3 sharded ", rel(grads[("tf32x3", 0)], grads[("tf32x3", "sh")]))
print("fp32 one pass vs fp32 sharded", rel(grads[("fp32", 0)], grads[("fp32", "sh")]))
