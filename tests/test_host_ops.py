"""CPU tests of the host-side data helpers against the reference semantics (operations.py:4-30)."""
import numpy as np

from pinn_depthestimation_b200 import operations as op

CFG = {"data_test": {"x_min": 25.0, "x_max": 33.0, "y_min": -13.0, "y_max": 13.0}}


def test_normalize_round_trip_and_degenerate_range():
    x = np.linspace(25.0, 33.0, 9)
    n = op.normalize(x, 25.0, 33.0)
    assert n.min() == -1.0 and n.max() == 1.0
    np.testing.assert_allclose(op.denormalize(n, 25.0, 33.0), x)
    assert np.all(op.normalize(x, 3.0, 3.0) == 0)


def test_get_min_max_uses_config_for_xy_and_data_otherwise():
    assert op.get_min_max(None, 'x', CFG) == {'x': (25.0, 33.0)}
    assert op.get_min_max(None, 'y', CFG) == {'y': (-13.0, 13.0)}
    d = np.array([[1.0], [np.nan], [-2.5]])
    assert op.get_min_max(d, 't', CFG) == {'t': (-2.5, 1.0)}
    assert op.get_min_max({'t': d}, 't', CFG) == {'t': (-2.5, 1.0)}


def test_the_condition_mask_is_all_true_on_normalised_x():
    """x is normalised to [-1,1] before physics.continuity_only compares it with 25.5."""
    x = op.normalize(np.linspace(25.0, 33.0, 81), 25.0, 33.0)
    assert np.all(x < 25.5)
