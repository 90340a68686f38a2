"""Runnable counterparts of the reference's two training scripts on the fused B200 trainer:

    python -m pinn_depthestimation_b200.train_main --config config_CMB_h.json                    (train_newmethod.py:212-270)
    python -m pinn_depthestimation_b200.train_main --form train --config config_CMB.json          (train.py:203-288)
           [--data file.mat | --fid-csv f.csv --res-mat r.mat] [--synthetic N] [--log-dir DIR]
           [--precision fp32|tf32|tf32x3] [--residual NAME] [--adam-it N] [--lbfgs-it N]

Unlike the reference the config name is an argument (the reference hard-codes it per script and reads it at import,
train_newmethod.py:35-36, train.py:35-36) and the normalise / hstack / NaN-row filter of the __main__ blocks runs on
the device (data.assemble_points).  `--form train` is the repaired `train.py.__main__` (as shipped it
imports matplotlib, calls operations.get_min_max with the old 2-argument signature and requires keys that config.json /
config_txyz.json do not have): fidelity points from the CSV + residual grid from the .mat, per-output loss weights
(default 1), dropout / init type (default 0.0 / 'xavier'), and the residual function chosen from the outputs the config
names -- (h, U, V, eta_mean, Hrms, k) -> physics_equation, (h, z, u, v) -> Navier_Stokes.  `--synthetic N` replaces the data
files (which the reference repository does not ship) so that every config can be run end to end.
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

from . import data as pdata
from . import operations as op
from .dist import shard_bounds
from .trainer import pinn


def load_mat_columns(path, names):
    from scipy.io import loadmat
    return {k: loadmat(path, variable_names=k)[k] for k in names}


# ------------------------------------------------------------------------------------------------
# train_newmethod.py form
# ------------------------------------------------------------------------------------------------
def build_arrays(config, data_file, device="cuda"):
    """train_newmethod.py:226-255 on the device: normalised inputs hstacked in config order, trues hstacked, rows with
    NaN trues dropped -> (inputs [N,d], trues [N,n_true]) CUDA tensors."""
    input_vars = list(config['data']['inputs'].keys())
    true_vars = list(config['data']['trues'])
    raw = load_mat_columns(data_file, input_vars + true_vars)
    X, T, _ = pdata.assemble_points({k: raw[k] for k in input_vars}, {k: raw[k] for k in true_vars}, config,
                                    device=device)
    return X, T


def synthetic_arrays(config, n, seed=1234):
    d = config['layers']['input_features']
    nt = len(config['data']['trues'])
    rs = np.random.RandomState(seed)
    X = rs.uniform(-1, 1, size=(n, d))
    T = np.stack([0.04 * np.sin(2.0 * X[:, 0] + i) * np.cos(1.5 * X[:, min(1, d - 1)] - i)
                  for i in range(nt)], axis=1)
    return X.astype(np.float32), T.astype(np.float32)


# ------------------------------------------------------------------------------------------------
# train.py form
# ------------------------------------------------------------------------------------------------
def residual_for_outputs(outputs):
    """Which physics.py function fits the outputs a config names (train.py:17 hard-codes physics_equation, which only
    fits config_CMB.json; config.json / config_txyz.json describe the (h, z, u, v) system of Navier_Stokes)."""
    names = list(outputs)
    if set(names) >= {"h", "U", "V", "eta_mean", "Hrms", "k"}:
        return "physics_equation"
    if set(names) >= {"h", "z", "u", "v"}:
        return "Navier_Stokes"
    if set(names) >= {"h", "U", "V"}:
        return "continuity_only"
    raise ValueError(f"no residual function of physics.py takes the outputs {names}")


def normalized_config(config):
    """Fill in what train.py reads but config.json / config_txyz.json lack (train.py:59,62,95,209)."""
    cfg = json.loads(json.dumps(config))
    cfg['layers'].setdefault('dropout_rate', 0.0)
    cfg['layers'].setdefault('init_type', 'xavier')
    for k in ('max_it', 'max_evaluation'):
        if k in cfg['lbfgs_optimizer']:
            cfg['lbfgs_optimizer'][k] = int(cfg['lbfgs_optimizer'][k])     # 5.00e4 / 6.25e4 are floats in config.json
    for k in cfg['data_fidelity']['outputs']:
        cfg['loss'].setdefault(f'weight_{k}_loss', 1)
    return cfg


def build_train_form_arrays(config, fid_csv, res_mat, device="cuda", seed=1234):
    """train.py:203-276: fidelity points from the CSV (rounded to 3 decimals, `training_points` rows drawn without
    replacement), residual grid from the .mat decimated by interval_x / interval_y, both normalised with the ranges of
    the fidelity inputs (x, y from config['data_test']), NaN rows of the residual grid dropped."""
    import pandas as pd
    from scipy.io import loadmat
    fin = list(config['data_fidelity']['inputs'])
    fout = list(config['data_fidelity']['outputs'])
    data = pd.read_csv(fid_csv).round(3)
    fid_in = {k: data[k].to_numpy() for k in fin}
    fid_true = {k: data[k].to_numpy() for k in fout}
    rs = np.random.RandomState(seed)                      # np.random.seed(1234) at train.py:22
    n_training = int(config['data_fidelity']['training_points'])
    idx = rs.choice(len(data), n_training, replace=False)
    Xf, Tf, ranges = pdata.assemble_points({k: fid_in[k][idx] for k in fin}, {k: fid_true[k][idx] for k in fout},
                                           config, drop_nan_trues=False, device=device,
                                           ranges={k: tuple(float(v) for v in op.get_min_max(fid_in[k], k, config)[k])
                                                   for k in fin})
    rin = list(config['data_residual']['inputs'].keys())
    ix, iy = int(config['data_residual']['interval_x']), int(config['data_residual']['interval_y'])
    cols = {}
    for k in rin:
        g = loadmat(res_mat, variable_names=k)[k][::ix, ::iy]
        cols[k] = np.transpose(g.reshape(-1, g.shape[1])).reshape(-1)       # train.py:262-265
    Xr, _, _ = pdata.assemble_points(cols, {}, config, ranges={k: ranges[k] for k in rin if k in ranges},
                                     drop_nan_inputs=True, device=device)
    return Xf, Tf, Xr


def synthetic_train_form_arrays(config, n_res, seed=1234):
    """Stand-in for the absent FUNWAVE / G1a files: `training_points` fidelity rows and n_res residual rows."""
    d = config['layers']['input_features']
    fout = list(config['data_fidelity']['outputs'])
    rs = np.random.RandomState(seed)
    nf = int(config['data_fidelity']['training_points'])
    Xf = rs.uniform(-1, 1, size=(nf, d)).astype(np.float32)
    base = {"h": 0.75, "eta_mean": 0.05, "Hrms": 0.3, "k": 0.8}
    Tf = np.stack([base.get(k, 0.0) + 0.04 * np.sin(2.0 * Xf[:, 0] + i) * np.cos(1.5 * Xf[:, 1] - i)
                   for i, k in enumerate(fout)], axis=1).astype(np.float32)
    Xr = rs.uniform(-1, 1, size=(n_res, d)).astype(np.float32)
    return Xf, Tf, Xr


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True)
    ap.add_argument("--form", default="newmethod", choices=["newmethod", "train"])
    ap.add_argument("--data", default=None)
    ap.add_argument("--fid-csv", default=None)
    ap.add_argument("--res-mat", default=None)
    ap.add_argument("--synthetic", type=int, default=0)
    ap.add_argument("--log-dir", default=None)
    ap.add_argument("--precision", default="fp32", choices=["fp32", "tf32", "tf32x3"])
    ap.add_argument("--residual", default=None)
    ap.add_argument("--adam-it", type=int, default=None, help="override adam_optimizer.max_it")
    ap.add_argument("--lbfgs-it", type=int, default=None, help="override lbfgs_optimizer.max_it")
    args = ap.parse_args(argv)
    with open(args.config) as f:
        config = json.load(f)
    if args.adam_it is not None:
        config['adam_optimizer']['max_it'] = args.adam_it
    if args.lbfgs_it is not None:
        config['lbfgs_optimizer']['max_it'] = args.lbfgs_it
        config['lbfgs_optimizer']['max_evaluation'] = args.lbfgs_it * 5 // 4
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

    log_dir = args.log_dir
    if log_dir is None and rank == 0:
        log_dir = os.path.join("..", "log", datetime.datetime.now().strftime("%Y%m%d_%H%M"))
    if rank == 0:
        os.makedirs(log_dir, exist_ok=True)
    common = dict(log_dir=log_dir if rank == 0 else None, group=group, precision=args.precision,
                  dump_path=os.path.join(log_dir, 'data_at50k.mat') if rank == 0 else 'data_at50k.mat')
    if args.form == "newmethod":
        if args.synthetic:
            X, T = synthetic_arrays(config, args.synthetic)
        else:
            X, T = build_arrays(config, args.data or config['data']['file'])
        lo, hi = shard_bounds(X.shape[0], rank, world)
        model = pinn(config, X[lo:hi], T[lo:hi], residual=args.residual or "continuity_only", **common)
    else:
        config = normalized_config(config)
        residual = args.residual or residual_for_outputs(config['data_residual']['outputs'])
        if args.synthetic:
            Xf, Tf, Xr = synthetic_train_form_arrays(config, args.synthetic)
        else:
            fid_csv = args.fid_csv or config['data_fidelity'].get('file') or config['data_fidelity'].get('dir')
            Xf, Tf, Xr = build_train_form_arrays(config, fid_csv, args.res_mat or config['data_residual']['file'])
        lo, hi = shard_bounds(Xr.shape[0], rank, world)
        fl, fh = shard_bounds(Xf.shape[0], rank, world)
        model = pinn(config, Xr[lo:hi], None, residual=residual, fid_input=Xf[fl:fh], fid_true=Tf[fl:fh], **common)
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
