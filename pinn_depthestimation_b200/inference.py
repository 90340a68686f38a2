"""Inference on the fused kernels: counterpart of the reference's test_newmethod.py (test_newmethod.py:15-121; test.py is the
same flow for the older configs).  `class pinn(model_path, config)` loads a whole-module checkpoint written by
`torch.save(model.dnn)` -- by this package or by the reference itself (train_newmethod.py:184,270) -- and `test()` runs the
value-only fused forward (pinn_jet_loss_fwd, residual kind NONE) over the evaluation grid.

    python -m pinn_depthestimation_b200.inference --config config_CMB_h.json [--model model.pth] [--file grid.mat] [--out pred.mat]

Differences from the reference, all deliberate: `test()` RETURNS the predictions (the reference computes them and drops them,
test_newmethod.py:69-72); `torch.load(..., weights_only=False)` (the reference's bare `torch.load` refuses whole-module
pickles on torch >= 2.6 and exits); when the grid file is absent (the reference repository does not ship
its data) the grid is rebuilt from config['data_test'] (nx, ny, x/y ranges); predictions can be exported to .mat with the
`pred_<name>` keys of train_newmethod.py:141-153.
"""
from __future__ import annotations

import argparse
import json
import sys

import numpy as np
import torch

from . import dnn as _dnn
from . import operations as op


class pinn:
    def __init__(self, model_path, config, device=None):
        self.config = config
        self.model_path = model_path
        self.device = torch.device(device) if device is not None else self.get_device()
        self.model = self.load_model()
        self.test_input_vars = config['data_test']['inputs']

    def get_device(self):
        if not torch.cuda.is_available():
            raise RuntimeError("pinn_b200 inference needs a CUDA device: this path has no CPU fallback")
        return torch.device('cuda')

    def load_model(self):
        # the pickle names its class `dnn.DNN` (reference) or `pinn_depthestimation_b200.dnn.DNN`; make both resolve here
        sys.modules.setdefault('dnn', _dnn)
        model = torch.load(self.model_path, map_location=self.device, weights_only=False)
        if not isinstance(model, _dnn.DNN):
            raise TypeError(f"{self.model_path} does not hold a DNN module (got {type(model).__name__}); if it was "
                            "unpickled through the reference's own dnn.py, put dropin/ first on sys.path")
        model.eval()
        return model

    def init_optimizers(self):
        """test.py:44-54 / test_newmethod.py:44-54: the one-iteration L-BFGS of the optional test-time physics step."""
        from .lbfgs import LBFGS
        l = self.config['lbfgs_optimizer']
        self.optimizer_LBFGS = LBFGS(self.model.parameters(), lr=l['learning_rate'], max_iter=1, max_eval=2, history_size=10,
                                     tolerance_grad=l['tolerance_grad'], tolerance_change=l['tolerance_change'],
                                     line_search_fn=l['line_search_fn'])

    def physics_step(self, x, residual=None):
        """test.py:91-104 (`perform_optimization`): one `optimizer_LBFGS.step(closure)` on the PDE residual over the test
        points (no data term), as one fused evaluation per closure call."""
        from .fused import JetLoss
        from .spec import DIR_ORDER, FIELD_ORDER, PassSpec
        from .train_main import residual_for_outputs
        in_names = list(self.test_input_vars.keys())
        out_names = list(self.config['data_test']['outputs'])
        kind = residual or residual_for_outputs(out_names)
        spec = PassSpec(layers=self.model.layer_sizes, activation=self.model.activation_name, kind=kind,
                        dirs={n: in_names.index(n) for n in DIR_ORDER[kind]},
                        fields={n: out_names.index(n) for n in FIELD_ORDER[kind]}, w_fid=0.0, w_res=1.0)
        jl = JetLoss(spec, x, None)
        self.init_optimizers()
        self.model.train()

        class _Closure:
            def flat_loss_and_grad(self, fp, fg):
                return jl.loss_and_grad(fp, fg)
        loss = self.optimizer_LBFGS.step(_Closure())
        self.model.eval()
        return float(loss)

    def test(self, test_input_data, perform_optimization=None, residual=None):
        """[N,d] normalised inputs (numpy or tensor) -> [N,o] predictions as a numpy array (test_newmethod.py:56-72).
        With config['perform_optimization'] (test.py:91) the test-time physics step runs first."""
        x = torch.as_tensor(np.ascontiguousarray(test_input_data) if not isinstance(test_input_data, torch.Tensor)
                            else test_input_data).float().to(self.device).contiguous()
        if perform_optimization is None:
            perform_optimization = bool(self.config.get('perform_optimization', False))
        if perform_optimization:
            self.physics_loss_before = self.physics_step(x, residual)
        with torch.no_grad():
            self.test_prediction_data = self.model(x).detach().cpu().numpy()
        return self.test_prediction_data


def grid_inputs(config, file=None):
    """Evaluation grid as the [nx*ny, d] normalised input array of test_newmethod.py:91-115.  From `file` (.mat with one 2-D
    array per input variable) when given; otherwise rebuilt from config['data_test'] (x_min..x_max in nx points etc.)."""
    dt = config['data_test']
    inputs = list(dt['inputs'].keys())
    cols = []
    if file is not None:
        from scipy.io import loadmat
        for key in inputs:
            data = loadmat(file, variable_names=key)[key]
            lo, hi = op.get_min_max(data, key, config)[key]
            cols.append(op.normalize(data, lo, hi).reshape(-1, 1))
    else:
        if inputs != ['x', 'y']:
            raise ValueError("the synthetic grid is defined for (x, y) inputs; pass --file for other input sets")
        xs = np.linspace(dt['x_min'], dt['x_max'], int(dt['nx']))
        ys = np.linspace(dt['y_min'], dt['y_max'], int(dt['ny']))
        X, Y = np.meshgrid(xs, ys, indexing='ij')          # [nx, ny] like the reference's .mat arrays
        cols = [op.normalize(X, dt['x_min'], dt['x_max']).reshape(-1, 1),
                op.normalize(Y, dt['y_min'], dt['y_max']).reshape(-1, 1)]
    return np.hstack(cols).astype(np.float32)


def export_mat(path, predictions, names):
    """{'pred_<name>': [N,1] float32}: the layout of the reference's data_at50k.mat (train_newmethod.py:141-153)."""
    import scipy.io as sio
    sio.savemat(path, {f'pred_{k}': np.ascontiguousarray(predictions[:, i:i + 1], dtype=np.float32)
                       for i, k in enumerate(names)})


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True)
    ap.add_argument("--model", default=None, help="checkpoint (default: config['data_test']['model'])")
    ap.add_argument("--file", default=None, help="grid .mat (default: rebuilt from config['data_test'])")
    ap.add_argument("--out", default=None, help="write predictions to this .mat")
    args = ap.parse_args(argv)
    with open(args.config) as f:
        config = json.load(f)
    tester = pinn(args.model or config['data_test']['model'], config)
    pred = tester.test(grid_inputs(config, args.file))
    print('Prediction done!', pred.shape)
    if args.out:
        export_mat(args.out, pred, list(config['data_test']['outputs']))
    return pred


if __name__ == "__main__":
    main()
