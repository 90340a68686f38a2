"""Checkpoint compatibility: whole-module pickles written by the REAL reference's dnn.DNN
(tests/golden/ref_module_*.pth, made by oracle/make_ref_pickles.py with `torch.save(model)` as in
train_newmethod.py:184,270) load through the drop-in DNN.__setstate__ once `dnn` resolves to dropin/dnn.py --
which is what test_newmethod.py:35-42 / test.py:37 do with `torch.load(path)`."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import jet_oracle as jo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "dropin")
CASES = [("xavier", "tanh", [2, 8, 8, 3]), ("kaiming", "leaky_relu", [3, 6, 5, 4])]


def _load(init, device="cpu"):
    if DROPIN not in sys.path:
        sys.path.insert(0, DROPIN)
    import dnn   # noqa: F401  (the pickle names its class as dnn.DNN)
    path = os.path.join(ROOT, "tests", "golden", f"ref_module_{init}.pth")
    return torch.load(path, map_location=device, weights_only=False)


@pytest.mark.parametrize("init,act,layers", CASES)
def test_reference_written_pickle_loads_into_the_dropin_class(init, act, layers):
    from pinn_depthestimation_b200.dnn import DNN
    m = _load(init)
    assert isinstance(m, DNN)
    # constructor facts recovered from the submodules (the reference's pickle carries only nn.Module state)
    assert m.layer_sizes == layers and m.init_type == init and m.dropout_rate == 0.0
    assert m.activation_name == act
    keys = list(m.state_dict().keys())
    assert keys == [f"layers.layer_{i}.{w}" for i in range(len(layers) - 1) for w in ("weight", "bias")]
    flat = jo.make_params(layers, 1234, act, np.float32)
    got = torch.cat([p.detach().reshape(-1) for p in m.parameters()]).numpy()
    assert np.array_equal(got, flat)
    # and it pickles again in the reference's format (train_newmethod.py:270)
    import io
    buf = io.BytesIO()
    torch.save(m, buf)
    buf.seek(0)
    m2 = torch.load(buf, weights_only=False)
    assert m2.layer_sizes == layers


@pytest.mark.gpu
@pytest.mark.parametrize("init,act,layers", CASES)
def test_reference_written_pickle_runs_on_the_fused_kernels(init, act, layers):
    dev = torch.device("cuda:0")
    m = _load(init, dev).to(dev)
    m.eval()
    X, _ = jo.make_points(77, layers[0], 0, seed=5)
    out = m(torch.from_numpy(X).to(dev)).detach().cpu().numpy()
    flat = jo.make_params(layers, 1234, act, np.float32)
    ref = jo.mlp_forward(layers, flat.astype(np.float64), X.astype(np.float64), act)
    assert np.abs(out - ref).max() <= 1e-5 * np.abs(ref).max()
