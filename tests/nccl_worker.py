"""Worker of tests/test_gpu_nccl.py: one rank per GPU (torchrun).  Each rank evaluates its contiguous shard of a golden
case through the C ABI with the global divisors; gradient and raw sums go through the ONE all-reduce of the product's
own code (fused.JetLoss(group=...) -> dist.all_reduce_eval); every rank must hold the reference's loss and gradient,
and a short L-BFGS run on the sharded evaluation must take identical branches on every rank."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np   # noqa: E402
import torch   # noqa: E402
import torch.distributed as dist   # noqa: E402

from pinn_depthestimation_b200 import dist as pdist   # noqa: E402
from pinn_depthestimation_b200.fused import JetLoss   # noqa: E402
from pinn_depthestimation_b200.lbfgs import LBFGS   # noqa: E402
from tests import cases   # noqa: E402
from tests.gpu_util import pass_specs   # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    report = {}
    for name, prec, lt, gt in (("cmb_h_small", "fp32", 1e-5, 1e-4), ("wide_nswe", "fp32", 1e-5, 1e-4),
                               ("wide_nswe", "tf32x3", 1e-5, 1e-4), ("wide_cont", "tf32", 5e-3, 5e-3)):
        case, z = cases.load(name)
        spec, _ = pass_specs(case, prec)
        flat, X, T, _, _ = cases.data(case, np.float32)
        lo, hi = pdist.shard_bounds(X.shape[0], rank, world)
        jl = JetLoss(spec, torch.from_numpy(X[lo:hi]).to(dev), torch.from_numpy(T[lo:hi]).to(dev), group=dist.group.WORLD)
        assert jl.n_res_global == X.shape[0]
        p = torch.from_numpy(flat).to(dev)
        g = torch.full_like(p, float("nan"))
        parts = jl.loss_and_grad(p, g).cpu().numpy().astype(np.float64)
        gg = g.cpu().numpy().astype(np.float64)
        el = abs(parts[2] - z["loss64"]) / abs(z["loss64"])
        eg = cases.golden_grad_check(z, gg)
        assert el <= lt and eg <= gt, (name, prec, rank, el, eg)
        # every rank holds bit-identical results (all decisions downstream are replicated)
        both = [torch.empty_like(g) for _ in range(world)]
        dist.all_gather(both, g)
        assert all(torch.equal(both[0], b) for b in both[1:]), "gradient differs between ranks"
        report[f"{name}/{prec}"] = {"loss_rel": el, "grad_rel": eg}
        if name == "cmb_h_small":
            q = torch.nn.Parameter(p.clone())
            opt = LBFGS([q], lr=1, max_iter=12, max_eval=15, history_size=100, tolerance_grad=1e-9,
                        tolerance_change=1e-12, line_search_fn="strong_wolfe")

            class C:
                def flat_loss_and_grad(self, fp, fg):
                    return jl.loss_and_grad(fp, fg)
            opt.step(C())
            st = opt.state[q]
            mine = torch.tensor([st["n_iter"], st["func_evals"]], device=dev, dtype=torch.float64)
            allv = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allv, mine)
            assert all(torch.equal(allv[0], v) for v in allv[1:]), "L-BFGS took different branches on different ranks"
            qs = [torch.empty_like(q.data) for _ in range(world)]
            dist.all_gather(qs, q.data)
            assert all(torch.equal(qs[0], v) for v in qs[1:]), "weights diverged between ranks"
            report["lbfgs"] = {"n_iter": st["n_iter"], "func_evals": st["func_evals"], "loss": st["loss"]}
    if rank == 0:
        print("NCCL_WORKER_OK " + json.dumps({"world": world, **report}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
