"""`from physics_functions import Boussinesq` -> B200-native drop-in for the reference's historical physics_functions
module (bytecode only in the reference tree; see pinn_depthestimation_b200/physics_functions.py)."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from pinn_depthestimation_b200.physics_functions import (  # noqa: E402,F401
    Boussinesq, Boussinesq_simple, compute_gradient)
