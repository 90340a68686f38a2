"""`import operations as op` -> drop-in for the reference's operations.py."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from pinn_depthestimation_b200.operations import denormalize, get_min_max, normalize  # noqa: E402,F401
