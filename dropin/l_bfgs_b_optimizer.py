"""`from l_bfgs_b_optimizer import LBFGSBOptimizer` -> torch/B200-backed re-offer of the
reference's historical module (see pinn_depthestimation_b200/l_bfgs_b_optimizer.py)."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from pinn_depthestimation_b200.l_bfgs_b_optimizer import LBFGSBOptimizer  # noqa: E402,F401
