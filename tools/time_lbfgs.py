"""Steady-state L-BFGS iteration rate at the config_CMB_h shape (development timing)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import jet_oracle as jo
from pinn_depthestimation_b200 import PassSpec
from pinn_depthestimation_b200.fused import JetLoss
from pinn_depthestimation_b200.lbfgs import LBFGS

layers = [2] + [20] * 100 + [3]
n = 12514
dev = torch.device("cuda:0")
X, _ = jo.make_points(n, 2, 0, seed=1234)
rs = np.random.RandomState(0)
T = (0.3 * np.sin(3 * X) + 0.05 * rs.standard_normal(X.shape)).astype(np.float32)   # noisy: never converges
spec = PassSpec(layers=layers, kind="continuity_only", dirs={"x": 0, "y": 1}, fields={"U": 0, "V": 1, "h": 2}, target_cols=[0, 1])
jl = JetLoss(spec, torch.from_numpy(X).to(dev), torch.from_numpy(T).to(dev))
q = torch.nn.Parameter(torch.from_numpy(jo.make_params(layers, 1234)).to(dev))
nev = [0]
class C:
    def flat_loss_and_grad(self, fp, fg):
        nev[0] += 1
        return jl.loss_and_grad(fp, fg)
for iters in (20, 200):
    lb = LBFGS([q], lr=1, max_iter=iters, max_eval=iters * 5 // 4, history_size=100, tolerance_grad=0, tolerance_change=0, line_search_fn="strong_wolfe")
    torch.cuda.synchronize(); t0 = time.perf_counter(); nev[0] = 0
    lb.step(C())
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    st = lb.state[q]
    print(f"max_iter={iters}: n_iter={st['n_iter']} evals={st['func_evals']} {dt*1e3:.1f} ms -> {st['n_iter']/dt:.1f} it/s, {dt/st['func_evals']*1e3:.2f} ms per evaluation, loss {st['loss']:.4e}")
import cProfile, pstats
lb = LBFGS([q], lr=1, max_iter=100, max_eval=125, history_size=100, tolerance_grad=0, tolerance_change=0, line_search_fn="strong_wolfe")
pr = cProfile.Profile(); pr.enable(); lb.step(C()); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
