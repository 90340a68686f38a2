"""`from dnn import DNN` -> B200-native DNN (drop-in for the reference's dnn.py)."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from pinn_depthestimation_b200.dnn import DNN  # noqa: E402,F401

# whole-module checkpoints (torch.save(model.dnn), train_newmethod.py:184,270) must stay loadable by
# anything that has a module called `dnn` with a class `DNN` -- including the reference itself.
DNN.__module__ = "dnn"
