"""Config dictionaries with the schema the reference's trainers consume, one per shipped config
file -- only the keys that are read (`layers`, `adam_optimizer`, `lbfgs_optimizer`, `loss`, `data*`), values as
shipped.  Written out here (not copied) so the tests do not depend on /root/reference at run time."""
import copy

_OPT = {
    "adam_optimizer": {"max_it": 50000, "learning_rate": 1e-4, "scheduler_step_size": 10000, "scheduler_gamma": 0.8},
    "lbfgs_optimizer": {"max_it": 50000, "learning_rate": 1, "max_evaluation": 6.25e4, "history_size": 100,
                        "tolerance_grad": 1e-5, "tolerance_change": 1e-7, "line_search_fn": "strong_wolfe"},
}
_XY = {"x": {"requires_grad": ["true"]}, "y": {"requires_grad": ["true"]}}
_GRID = {"nx": 81, "ny": 261, "dx": 0.1, "dy": 0.1, "x_min": 25.0, "x_max": 33.0, "y_min": -13.0, "y_max": 13.0}


def cmb_h(hidden_layers=100):
    """config_CMB_h.json: train_newmethod.py, continuity_only, outputs (U, V | h)."""
    return copy.deepcopy({
        "layers": {"input_features": 2, "hidden_layers": hidden_layers, "hidden_width": 20, "output_features": 3,
                   "dropout_rate": 0.0, "init_type": "xavier"},
        **_OPT, "loss": {"weight_fid_loss": 1, "weight_res_loss": 1},
        "data": {"file": "../data/G1a/processed/data_60percent.mat", "inputs": _XY, "trues": ["U", "V"], "unknowns": ["h"]},
        "data_test": {"model": "model.pth", "inputs": _XY, "outputs": ["U", "V", "h"], **_GRID},
    })


def cmb():
    """config_CMB.json: train.py, physics_equation, 12 fidelity points + decimated residual grid."""
    outs = ["h", "U", "V", "eta_mean", "Hrms", "k"]
    return copy.deepcopy({
        "layers": {"input_features": 2, "hidden_layers": 10, "hidden_width": 10, "output_features": 6,
                   "dropout_rate": 0.0, "init_type": "xavier"},
        **_OPT, "loss": {**{f"weight_{k}_loss": 1 for k in outs}, "weight_fid_loss": 1, "weight_res_loss": 1},
        "data_fidelity": {"file": "input_fid.csv", "inputs": ["x", "y"], "outputs": outs, "training_points": 12},
        "data_residual": {"file": "input_res.mat", "inputs": _XY, "outputs": outs, "snapshots": [1],
                          "interval_x": 10, "interval_y": 10},
        "data_test": {"model": "model.pth", "inputs": _XY, "outputs": outs, **_GRID},
    })


def legacy_txy(inputs, hidden_layers):
    """config.json ((t,x,y,u,v) inputs, 100 layers) / config_txyz.json ((t,x,y,z) inputs, 20 layers): the (h,z,u,v)
    system; no dropout_rate / init_type / per-output weights; float iteration counts; Adam disabled (max_it 0)."""
    opt = copy.deepcopy(_OPT)
    opt["adam_optimizer"]["max_it"] = 0
    opt["lbfgs_optimizer"]["max_it"] = 5.00e4
    return {
        "layers": {"input_features": len(inputs), "hidden_layers": hidden_layers, "hidden_width": 20, "output_features": 4},
        **opt, "loss": {"weight_fid_loss": 1, "weight_res_loss": 100000},
        "data_fidelity": {"dir": "../data/beach2d_irr.csv", "inputs": list(inputs), "outputs": ["h", "z", "u", "v"],
                          "training_points": 9600},
        "data_residual": {"inputs": {k: {"file": k, "requires_grad": ["true" if k in "txy" else "false"]} for k in inputs},
                          "outputs": {"h": {"file": "dep.out"}, "z": {"file": "eta"}, "u": {"file": "u"}, "v": {"file": "v"}}},
    }


def config_json():
    return legacy_txy(["t", "x", "y", "u", "v"], 100)


def config_txyz():
    return legacy_txy(["t", "x", "y", "z"], 20)
