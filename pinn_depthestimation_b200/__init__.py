"""pinn_depthestimation_b200 -- B200-native PINN training hot path behind the reference's own
dnn.DNN / physics.* / LBFGS interfaces.  See DESIGN.md and include/pinn_b200.h."""
from .spec import PassSpec, param_count  # noqa: F401

__all__ = ["PassSpec", "param_count"]
