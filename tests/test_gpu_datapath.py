"""GPU tests of the widened rows of SURVEY.md 8f: the on-device data path (normalise / nanmin-nanmax / hstack / NaN-row
filter of train_newmethod.py:226-255 and train.py:203-276 with operations.py:4-30), the repaired train.py-form entry point
for config_CMB.json / config.json / config_txyz.json, the prediction dump of train_newmethod.py:141-153 and the grid
inference of test_newmethod.py:56-72."""
import json
import os
import sys

import numpy as np
import pytest
import torch

from oracle import jet_oracle as jo

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = {"data_test": {"x_min": 25.0, "x_max": 33.0, "y_min": -13.0, "y_max": 13.0, "nx": 81, "ny": 261,
                     "inputs": {"x": {"requires_grad": ["true"]}, "y": {"requires_grad": ["true"]}},
                     "outputs": ["U", "V", "h"]}}


def _numpy_reference(cols_in, cols_true, names_in, drop_true=True, drop_in=False):
    """The reference's host code, re-typed: operations.get_min_max + normalize, hstack, NaN-row mask."""
    from pinn_depthestimation_b200 import operations as op
    X = []
    for k in names_in:
        lo, hi = op.get_min_max(cols_in[k], k, CFG)[k]
        X.append(op.normalize(cols_in[k].astype(np.float64), lo, hi).reshape(-1, 1))
    X = np.hstack(X)
    T = np.hstack([cols_true[k].reshape(-1, 1) for k in cols_true]) if cols_true else None
    keep = np.ones(X.shape[0], dtype=bool)
    if drop_true and T is not None:
        keep &= ~np.isnan(T).any(axis=1)
    if drop_in:
        keep &= ~np.isnan(X).any(axis=1)
    return X[keep], (T[keep] if T is not None else None)


@pytest.mark.parametrize("n", [1, 255, 256, 257, 5000, 300001])
def test_assemble_points_matches_the_reference_host_code(n):
    from pinn_depthestimation_b200 import data as pdata
    rs = np.random.RandomState(n)
    cols_in = {"x": rs.uniform(25, 33, n).astype(np.float32), "y": rs.uniform(-13, 13, n).astype(np.float32),
               "t": rs.uniform(3, 9, n).astype(np.float32)}
    cols_true = {"U": rs.standard_normal(n).astype(np.float32), "V": rs.standard_normal(n).astype(np.float32)}
    cols_true["U"][rs.rand(n) < 0.1] = np.nan
    cols_true["V"][rs.rand(n) < 0.05] = np.nan
    cols_in["t"][rs.rand(n) < 0.07] = np.nan          # nanmin / nanmax must ignore these; rows stay (policy: trues only)
    X, T, ranges = pdata.assemble_points(cols_in, cols_true, CFG)
    Xr, Tr = _numpy_reference(cols_in, cols_true, ["x", "y", "t"])
    assert X.shape == Xr.shape and T.shape == Tr.shape
    assert ranges["x"] == (25.0, 33.0) and ranges["y"] == (-13.0, 13.0)
    assert ranges["t"] == (float(np.nanmin(cols_in["t"])), float(np.nanmax(cols_in["t"]))) or np.isnan(cols_in["t"]).all()
    np.testing.assert_allclose(X.cpu().numpy(), Xr, rtol=0, atol=2e-6, equal_nan=True)
    assert np.array_equal(T.cpu().numpy(), Tr.astype(np.float32))      # order preserved, values untouched


def test_assemble_points_input_nan_policy_and_degenerate_range():
    from pinn_depthestimation_b200 import data as pdata
    n = 1000
    rs = np.random.RandomState(3)
    cols_in = {"x": rs.uniform(25, 33, n).astype(np.float32), "y": rs.uniform(-13, 13, n).astype(np.float32),
               "z": np.full(n, 4.0, np.float32)}                      # max == min -> all zeros (operations.py:5-6)
    cols_in["x"][::7] = np.nan
    X, T, _ = pdata.assemble_points(cols_in, {}, CFG, drop_nan_inputs=True)
    Xr, _ = _numpy_reference(cols_in, {}, ["x", "y", "z"], drop_in=True)
    assert T is None and X.shape == Xr.shape == (n - len(range(0, n, 7)), 3)
    np.testing.assert_allclose(X.cpu().numpy(), Xr, rtol=0, atol=2e-6)
    assert (X[:, 2] == 0).all()
    with pytest.raises(RuntimeError):
        pdata.assemble_points(cols_in, {}, CFG, device="cpu")


def _cmbh_config(adam=30, lbfgs=6):
    from tests import configs
    cfg = configs.cmb_h(hidden_layers=6)
    cfg["adam_optimizer"]["max_it"] = adam
    cfg["lbfgs_optimizer"]["max_it"] = lbfgs
    cfg["lbfgs_optimizer"]["max_evaluation"] = lbfgs * 5 // 4
    return cfg


def test_prediction_dump_and_grid_inference_round_trip(tmp_path):
    """train_newmethod.py:141-153 (the .mat dump fires in the evaluation that starts with iter == dump_at) and
    test_newmethod.py:35-72 (load the whole-module checkpoint, forward on the 81 x 261 grid)."""
    from scipy.io import loadmat
    from pinn_depthestimation_b200 import inference
    from pinn_depthestimation_b200.trainer import pinn
    cfg = _cmbh_config()
    n = 2000
    X, _ = jo.make_points(n, 2, 0, seed=1234)
    T = np.stack([0.04 * np.sin(2 * X[:, 0]), 0.03 * np.cos(X[:, 1])], axis=1).astype(np.float32)
    dump = str(tmp_path / "data_at50k.mat")
    model = pinn(cfg, X, T, device="cuda:0", log_dir=str(tmp_path), dump_at=cfg["adam_optimizer"]["max_it"], dump_path=dump)
    model.train()
    assert model.iter > cfg["adam_optimizer"]["max_it"]
    mat = loadmat(dump)
    assert sorted(k for k in mat if k.startswith("pred_")) == ["pred_U", "pred_V", "pred_h"]
    for k in ("pred_U", "pred_V", "pred_h"):
        assert mat[k].shape == (n, 1) and mat[k].dtype == np.float32      # the layout of the reference's data_at50k.mat
    # log.txt: header + one line per evaluation, %.5e (train_newmethod.py:164-175)
    lines = open(tmp_path / "log.txt").read().splitlines()
    assert lines[0] == "Epoch, Fidelity Loss, Residual Loss, Total Loss" and len(lines) == model.iter + 1
    # checkpoint -> inference on the data_test grid
    ck = str(tmp_path / "model.pth")
    torch.save(model.dnn, ck)
    cfg["data_test"]["model"] = ck
    tester = inference.pinn(ck, cfg)
    grid = inference.grid_inputs(cfg)
    assert grid.shape == (81 * 261, 2) and grid.min() == -1.0 and grid.max() == 1.0
    pred = tester.test(grid)
    assert pred.shape == (81 * 261, 3)
    flat = model.flat.detach().cpu().numpy().astype(np.float64)
    ref = jo.mlp_forward(model.layers, flat, grid.astype(np.float64))
    assert np.abs(pred - ref).max() <= 1e-5 * np.abs(ref).max()
    # test.py:91-104: the optional test-time physics step (one L-BFGS iteration on the residual over the grid)
    cfg2 = dict(cfg, perform_optimization=True)
    tester2 = inference.pinn(ck, cfg2)
    pred2 = tester2.test(grid)
    assert pred2.shape == pred.shape and np.isfinite(pred2).all()
    assert np.abs(pred2 - pred).max() > 0                      # the step moved the weights
    st2 = tester2.optimizer_LBFGS.state[tester2.optimizer_LBFGS._params[0]]
    assert st2["n_iter"] == 1 and st2["func_evals"] <= 3       # max_iter=1, max_eval=2 as in test.py:47-49 (max_ls bounds line-search ITERATIONS: 1 + 1 + 1 evaluations at most, like torch)
    out = str(tmp_path / "pred.mat")
    inference.export_mat(out, pred, cfg["data_test"]["outputs"])
    assert loadmat(out)["pred_h"].shape == (81 * 261, 1)
    # the dump holds the predictions of the weights at that evaluation: compare with a forward at the dump point is not
    # possible afterwards, but the columns must be the network's three outputs in (trues, unknowns) order
    assert np.isfinite(mat["pred_h"]).all()


@pytest.mark.parametrize("name,residual", [("cmb", "physics_equation"), ("config_json", "Navier_Stokes"),
                                           ("config_txyz", "Navier_Stokes")])
def test_train_form_entry_point_runs_every_legacy_config(tmp_path, name, residual):
    """The repaired train.py.__main__ (train.py:203-288) on configs with the schema and values of the reference's
    config_CMB.json / config.json / config_txyz.json (tests/configs.py), schedules shortened on the command line,
    synthetic stand-ins for the absent data files."""
    from pinn_depthestimation_b200 import train_main
    from tests import configs
    raw = getattr(configs, name)()
    cfg_path = str(tmp_path / f"{name}.json")
    json.dump(raw, open(cfg_path, "w"))
    cfg = train_main.normalized_config(raw)
    assert train_main.residual_for_outputs(cfg["data_residual"]["outputs"]) == residual
    model = train_main.main(["--form", "train", "--config", cfg_path, "--synthetic", "600", "--adam-it", "20",
                             "--lbfgs-it", "4", "--log-dir", str(tmp_path)])
    assert model.jl.fid is not None and model.jl.res.spec.kind == residual
    h = model.history
    assert len(h) >= 21 and all(np.isfinite(v[3]) for v in h)
    assert h[19][3] < h[0][3]                              # 20 Adam steps made progress
    assert os.path.exists(tmp_path / "model.pth") and os.path.exists(tmp_path / "log.txt")


def test_train_form_data_preparation_from_files(tmp_path):
    """train.py:203-276 on real files: fidelity CSV (rounded to 3 decimals, 12 rows drawn), residual .mat grid decimated by
    interval_x / interval_y, transposed flatten, NaN rows dropped -- against the same steps re-typed in numpy."""
    import pandas as pd
    from scipy.io import savemat
    from pinn_depthestimation_b200 import operations as op
    from pinn_depthestimation_b200 import train_main
    from tests import configs
    cfg = configs.cmb()
    rs = np.random.RandomState(5)
    nrow = 300
    outs = cfg["data_fidelity"]["outputs"]
    df = pd.DataFrame({"x": rs.uniform(25, 33, nrow), "y": rs.uniform(-13, 13, nrow),
                       **{k: rs.standard_normal(nrow) for k in outs}})
    csv = str(tmp_path / "input_fid.csv")
    df.to_csv(csv, index=False)
    xs, ys = np.meshgrid(np.linspace(25, 33, 81), np.linspace(-13, 13, 261), indexing="ij")
    xs = xs.copy()
    xs[40, 100] = np.nan                                   # one grid node without data
    mat = str(tmp_path / "input_res.mat")
    savemat(mat, {"x": xs, "y": ys})
    Xf, Tf, Xr = train_main.build_train_form_arrays(cfg, csv, mat)
    # numpy restatement
    d3 = df.round(3)
    idx = np.random.RandomState(1234).choice(nrow, 12, replace=False)
    Xf_ref = np.stack([op.normalize(d3["x"].to_numpy()[idx], 25.0, 33.0), op.normalize(d3["y"].to_numpy()[idx], -13.0, 13.0)], 1)
    np.testing.assert_allclose(Xf.cpu().numpy(), Xf_ref, atol=2e-6)
    np.testing.assert_allclose(Tf.cpu().numpy(), np.stack([d3[k].to_numpy()[idx] for k in outs], 1).astype(np.float32))
    cols = []
    for g, (lo, hi) in ((xs, (25.0, 33.0)), (ys, (-13.0, 13.0))):
        g = op.normalize(g[::10, ::10], lo, hi)
        cols.append(np.transpose(g.reshape(-1, g.shape[1])).reshape(-1, 1))
    R = np.hstack(cols)
    R = R[~np.isnan(R).any(axis=1)]
    assert Xr.shape == R.shape == (9 * 27 - 1, 2)
    np.testing.assert_allclose(Xr.cpu().numpy(), R, atol=2e-6)
