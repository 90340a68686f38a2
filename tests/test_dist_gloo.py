"""world_size-2 gloo test (CPU) of the multi-GPU protocol: each rank evaluates its contiguous shard
with the GLOBAL divisors (here with the numpy oracle standing in for the kernel), the single fp32
collective [grad | sums] is all-reduced with the product's own pack/unpack code, and every rank
must arrive at the reference's single-process loss and gradient."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, name, out):
    sys.path.insert(0, ROOT)
    from oracle import jet_oracle as jo
    from pinn_depthestimation_b200 import dist as pdist
    from tests import cases
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        case, z = cases.load(name)
        sres, _ = cases.specs(case)
        flat, X, T, _, _ = cases.data(case, np.float64)
        n = X.shape[0]
        lo, hi = pdist.shard_bounds(n, rank, world)
        n_res, n_fid = pdist.global_counts(hi - lo, hi - lo, torch.device("cpu"))
        assert n_res == n and n_fid == n
        cnt = torch.tensor([float((X[lo:hi, 0] < 25.5).sum())])
        dist.all_reduce(cnt)
        r = jo.loss_and_grad(sres, flat, X[lo:hi], T[lo:hi], n_global=n, mask_count=float(cnt))
        # the kernel's raw sums for this shard: sum fc^2, masked sum (h-0.75)^2, per-target SSE
        sums = torch.zeros(pdist.NSUMS, dtype=torch.float64)
        sums[0] = float(r["terms"]["continuity"]) * n
        sums[3] = float(r["terms"]["condition"]) * float(cnt)
        for i, c in enumerate(sres["target_cols"]):
            sums[5 + i] = float(((r["out"][:, c] - T[lo:hi, i]) ** 2).sum())
        sums[13] = hi - lo
        P = flat.size
        grad = torch.from_numpy(r["grad"].astype(np.float32))
        buf = torch.zeros(pdist.collective_numel(P), dtype=torch.float32)
        pdist.all_reduce_eval(buf, grad, sums, None)
        # finalize like pinn_loss_finalize
        fid = sum(sums[5 + i].item() / n for i in range(len(sres["target_cols"])))
        res = sums[0].item() / n + sums[3].item() / float(cnt)
        loss = fid + res
        assert sums[13].item() == n
        assert abs(loss - z["loss64"]) <= 2e-6 * abs(z["loss64"])          # fp32 collective
        assert cases.golden_grad_check(z, grad.numpy()) <= 1e-5
        out.put((rank, loss, float(grad.double().norm())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["cmb_h_small"])
def test_two_rank_sharded_evaluation_matches_single_process_reference(name):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    got = sorted(out.get(timeout=10) for _ in range(2))
    # every rank holds bit-identical reduced values -> identical optimiser decisions
    assert got[0][1:] == got[1][1:]


def test_shard_bounds_cover_everything_once():
    from pinn_depthestimation_b200.dist import shard_bounds
    for n in (0, 1, 7, 16, 12514, 16777216):
        for w in (1, 2, 3, 4, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1
