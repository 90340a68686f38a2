"""Runnable counterpart of the reference's `python train_newmethod.py` (train_newmethod.py:212-270)
on the fused B200 trainer:

    python -m pinn_depthestimation_b200.train_main --config config_CMB_h.json [--data file.mat]
           [--log-dir DIR] [--precision fp32|tf32] [--synthetic N]

Unlike the reference the config name is an argument (the reference hard-codes it per script and
reads it at import, train_newmethod.py:35-36).  `--synthetic N` replaces the .mat file (which the
reference repository does not ship) by N smooth synthetic points so that the script can be tried.
Under torchrun every rank trains on its contiguous shard of the points.
"""
from __future__ import annotations

import argparse
import datetime
import json
import os
import time

import numpy as np
import torch

from . import operations as op
from .dist import shard_bounds
from .trainer import pinn


def load_mat_columns(path, names):
    from scipy.io import loadmat
    return {k: loadmat(path, variable_names=k)[k] for k in names}


def build_arrays(config, data_file):
    """train_newmethod.py:226-255: normalised inputs hstacked in config order, trues hstacked,
    rows with NaN trues dropped."""
    input_vars = list(config['data']['inputs'].keys())
    true_vars = list(config['data']['trues'])
    raw = load_mat_columns(data_file, input_vars + true_vars)
    cols = []
    for key in input_vars:
        lo, hi = op.get_min_max(raw[key], key, config)[key]
        cols.append(op.normalize(raw[key], lo, hi).reshape(-1, 1))
    data_input = np.hstack(cols)
    data_true = np.hstack([raw[k].reshape(-1, 1) for k in true_vars])
    keep = ~np.isnan(data_true).any(axis=1)
    return data_input[keep], data_true[keep]


def synthetic_arrays(config, n, seed=1234):
    d = config['layers']['input_features']
    nt = len(config['data']['trues'])
    rs = np.random.RandomState(seed)
    X = rs.uniform(-1, 1, size=(n, d))
    T = np.stack([0.04 * np.sin(2.0 * X[:, 0] + i) * np.cos(1.5 * X[:, min(1, d - 1)] - i)
                  for i in range(nt)], axis=1)
    return X.astype(np.float32), T.astype(np.float32)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True)
    ap.add_argument("--data", default=None)
    ap.add_argument("--synthetic", type=int, default=0)
    ap.add_argument("--log-dir", default=None)
    ap.add_argument("--precision", default="fp32", choices=["fp32", "tf32"])
    ap.add_argument("--residual", default="continuity_only")
    args = ap.parse_args(argv)
    with open(args.config) as f:
        config = json.load(f)
    np.random.seed(1234)
    torch.manual_seed(1234)                      # weights are drawn on the CPU
    torch.cuda.manual_seed_all(1234)

    group, rank, world = None, 0, 1
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        group, rank, world = dist.group.WORLD, dist.get_rank(), dist.get_world_size()

    if args.synthetic:
        X, T = synthetic_arrays(config, args.synthetic)
    else:
        X, T = build_arrays(config, args.data or config['data']['file'])
    lo, hi = shard_bounds(X.shape[0], rank, world)
    log_dir = args.log_dir
    if log_dir is None and rank == 0:
        log_dir = os.path.join("..", "log", datetime.datetime.now().strftime("%Y%m%d_%H%M"))
    if rank == 0:
        os.makedirs(log_dir, exist_ok=True)
    model = pinn(config, X[lo:hi], T[lo:hi], residual=args.residual, log_dir=log_dir if rank == 0 else None,
                 group=group, precision=args.precision)
    if group is not None:       # identical initial weights on every rank
        import torch.distributed as dist
        dist.broadcast(model.flat, src=0, group=group)
    start = time.time()
    model.train()
    torch.cuda.synchronize()
    print('Training time: %.4f' % (time.time() - start))
    if rank == 0:
        torch.save(model.dnn, os.path.join(log_dir, 'model.pth'))
    return model


if __name__ == "__main__":
    main()
