"""`from physics import ...` -> B200-native residuals (drop-in for the reference's physics.py)."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from pinn_depthestimation_b200.physics import (  # noqa: E402,F401
    Navier_Stokes, compute_gradient, continuity_ftemp, continuity_only, physics_equation)
