"""Multi-GPU plumbing for the sharded evaluation: contiguous point shards, one
all-reduce of a single fp32 buffer [gradient | residual sums | fidelity sums] per evaluation.

Pure tensor code with no device assumptions, so the protocol is testable with gloo on CPU
(tests/test_dist_gloo.py); on GPUs the same functions run over NCCL.
"""
from __future__ import annotations

import torch

NSUMS = 16


def shard_bounds(n_points: int, rank: int, world: int):
    """Contiguous, balanced shard [lo, hi) of rank `rank`."""
    return rank * n_points // world, (rank + 1) * n_points // world


def collective_numel(n_params: int) -> int:
    return n_params + 2 * NSUMS


def pack_collective(buf, grad, sums_res, sums_fid=None):
    """[P] grad (or None), [16] double sums -> the fp32 collective buffer."""
    P = buf.numel() - 2 * NSUMS
    if grad is not None:
        buf[:P].copy_(grad)
    else:
        buf[:P].zero_()
    buf[P:P + NSUMS].copy_(sums_res)
    if sums_fid is not None:
        buf[P + NSUMS:].copy_(sums_fid)
    else:
        buf[P + NSUMS:].zero_()
    return buf


def unpack_collective(buf, grad, sums_res, sums_fid=None):
    P = buf.numel() - 2 * NSUMS
    if grad is not None:
        grad.copy_(buf[:P])
    sums_res.copy_(buf[P:P + NSUMS])
    if sums_fid is not None:
        sums_fid.copy_(buf[P + NSUMS:])


def all_reduce_eval(buf, grad, sums_res, sums_fid=None, group=None):
    """The one collective of an evaluation."""
    import torch.distributed as dist
    pack_collective(buf, grad, sums_res, sums_fid)
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    unpack_collective(buf, grad, sums_res, sums_fid)


def global_counts(n_res_local: int, n_fid_local: int, device, group=None):
    import torch.distributed as dist
    cnt = torch.tensor([n_res_local, n_fid_local], dtype=torch.int64, device=device)
    dist.all_reduce(cnt, group=group)
    return int(cnt[0]), int(cnt[1])
