"""Training-curve equivalence at the reference's own config shape (config_CMB_h.json:
[2]+[20]*100+[3], continuity_only, N=12,514 as in data_at50k.mat) on synthetic data:

  reference algorithm  = oracle/autograd_port.py (dnn.py+physics.py via torch autograd) + torch.optim.Adam
                         + StepLR + torch.optim.LBFGS on CPU      (train_newmethod.py:95-117,194-209)
  B200 path            = fused jet kernel + FusedAdam + StepLR + device-backed LBFGS

Same initial weights, same data, same hyper-parameters.  Writes profiles/r1_curve_equivalence.{csv,json}.
data_at50k.mat itself cannot be replayed (it holds only predictions).
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import autograd_port as ap
from oracle import jet_oracle as jo
from pinn_depthestimation_b200 import PassSpec
from pinn_depthestimation_b200.fused import JetLoss
from pinn_depthestimation_b200.lbfgs import LBFGS, FusedAdam

ap_ = argparse.ArgumentParser()
ap_.add_argument("--n", type=int, default=12514)
ap_.add_argument("--hidden", type=int, default=100)
ap_.add_argument("--adam", type=int, default=200)
ap_.add_argument("--lbfgs", type=int, default=40)
ap_.add_argument("--out", default="profiles/r1_curve_equivalence")
a = ap_.parse_args()

layers = [2] + [20] * a.hidden + [3]
ospec = dict(layers=layers, activation="tanh", kind=jo.CONT_ONLY, dirs={"x": 0, "y": 1},
             fields={"U": 0, "V": 1, "h": 2}, target_cols=[0, 1])
flat0 = jo.make_params(layers, 1234)
# smooth synthetic "currents" so that the fit is non-trivial
X, _ = jo.make_points(a.n, 2, 0, seed=1234)
T = np.stack([-0.045 + 0.04 * np.sin(2.0 * X[:, 0]) * np.cos(1.5 * X[:, 1]),
              0.03 * np.cos(1.0 * X[:, 0] + 0.5) * np.sin(2.5 * X[:, 1])], axis=1).astype(np.float32)
lb_kw = dict(lr=1, max_iter=a.lbfgs, max_eval=a.lbfgs * 5 // 4, history_size=100, tolerance_grad=1e-5,
             tolerance_change=1e-7, line_search_fn="strong_wolfe")

# ---------------- reference algorithm on CPU ----------------
torch.set_num_threads(os.cpu_count() or 1)
p = torch.nn.Parameter(torch.from_numpy(flat0.copy()))
Xt, Tt = torch.from_numpy(X), torch.from_numpy(T)
ref = []
t0 = time.perf_counter()
adam = torch.optim.Adam([p], lr=1e-4)
sched = torch.optim.lr_scheduler.StepLR(adam, step_size=10000, gamma=0.8)
for _ in range(a.adam):
    r = ap.loss_and_grad(ospec, p.detach(), Xt, Tt)
    p.grad = r["grad"]
    ref.append(float(r["loss"]))
    adam.step(); sched.step()
t_ref_adam = time.perf_counter() - t0
# L-BFGS phase: restarted from the initial weights (after 200 Adam steps this problem is already below
# tolerance_grad and both optimisers return after one evaluation)
p = torch.nn.Parameter(torch.from_numpy(flat0.copy()))
opt = torch.optim.LBFGS([p], **lb_kw)


def closure():
    opt.zero_grad()
    r = ap.loss_and_grad(ospec, p.detach(), Xt, Tt)
    p.grad = r["grad"]
    ref.append(float(r["loss"]))
    return r["loss"]


t0 = time.perf_counter()
opt.step(closure)
t_ref_lbfgs = time.perf_counter() - t0
ref_state = dict(opt.state[p])

# ---------------- B200 path ----------------
dev = torch.device("cuda:0")
spec = PassSpec(layers=layers, kind="continuity_only", dirs={"x": 0, "y": 1}, fields={"U": 0, "V": 1, "h": 2},
                target_cols=[0, 1])
jl = JetLoss(spec, torch.from_numpy(X).to(dev), torch.from_numpy(T).to(dev))
q = torch.nn.Parameter(torch.from_numpy(flat0.copy()).to(dev))
g = torch.empty_like(q)
n_tot = a.adam + 4 * a.lbfgs
mine = torch.zeros(n_tot, device=dev)
k = 0
fa = FusedAdam([q], lr=1e-4)
fs = torch.optim.lr_scheduler.StepLR(fa, step_size=10000, gamma=0.8)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(a.adam):
    parts = jl.loss_and_grad(q.detach(), g)
    mine[k] = parts[2]; k += 1
    fa.step(flat_grad=g); fs.step()
torch.cuda.synchronize(); t_b200_adam = time.perf_counter() - t0
q = torch.nn.Parameter(torch.from_numpy(flat0.copy()).to(dev))
lb = LBFGS([q], **lb_kw)


class C:
    def flat_loss_and_grad(self, fp, fg):
        global k
        parts = jl.loss_and_grad(fp, fg)
        mine[k] = parts[2]; k += 1
        return parts


t0 = time.perf_counter()
lb.step(C())
torch.cuda.synchronize(); t_b200_lbfgs = time.perf_counter() - t0
mine = mine[:k].cpu().numpy()
st = lb.state[q]

n = min(len(ref), len(mine))
ref_a, mine_a = np.array(ref[:n]), mine[:n]
rel = np.abs(mine_a - ref_a) / np.abs(ref_a)
summary = {
    "config": {"layers": f"[2]+[20]*{a.hidden}+[3]", "n_points": a.n, "residual": "continuity_only",
               "adam_iters": a.adam, "adam_lr": 1e-4, "lbfgs": {k_: str(v) for k_, v in lb_kw.items()}},
    "evaluations_compared": int(n), "ref_evaluations": len(ref), "b200_evaluations": int(len(mine)),
    "adam_phase_max_rel_diff": float(rel[:a.adam].max()),
    "lbfgs_phase_max_rel_diff": float(rel[a.adam:].max()) if n > a.adam else None,
    "loss_start": float(ref_a[0]), "ref_loss_end": float(ref[-1]), "b200_loss_end": float(mine[-1]),
    "ref": {"n_iter": ref_state["n_iter"], "func_evals": ref_state["func_evals"],
            "adam_s": t_ref_adam, "lbfgs_s": t_ref_lbfgs, "lbfgs_it_per_s": ref_state["n_iter"] / t_ref_lbfgs,
            "device": f"cpu x{os.cpu_count()}"},
    "b200": {"n_iter": st["n_iter"], "func_evals": st["func_evals"], "adam_s": t_b200_adam,
             "lbfgs_s": t_b200_lbfgs, "lbfgs_it_per_s": st["n_iter"] / t_b200_lbfgs,
             "adam_it_per_s": a.adam / t_b200_adam},
}
os.makedirs(os.path.dirname(a.out), exist_ok=True)
with open(a.out + ".json", "w") as f:
    json.dump(summary, f, indent=1)
with open(a.out + ".csv", "w") as f:
    f.write("evaluation,phase,reference_loss,b200_loss,rel_diff\n")
    for i in range(n):
        f.write(f"{i + 1},{'adam' if i < a.adam else 'lbfgs'},{ref_a[i]:.8e},{mine_a[i]:.8e},{rel[i]:.3e}\n")
print(json.dumps(summary))
