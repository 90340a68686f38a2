"""Drop-in for the reference's operations.py (operations.py:4-30): min/max normalisation of the
inputs to [-1, 1].  One-off host-side numpy work done before the tensors reach the GPU (SURVEY.md 2
row 12); kept so that scripts written against the reference keep importing `operations`.

Note that every derivative in the PDE residuals is taken with respect to these NORMALISED inputs
(no chain-rule rescaling anywhere in the reference), which is why `x < 25.5` in
physics.continuity_only is always true.
"""
import numpy as np


def normalize(data, data_min, data_max):
    """2 (x - min) / (max - min) - 1; all zeros when max == min (operations.py:4-7)."""
    if data_max == data_min:
        return np.zeros_like(data)
    return 2 * (data - data_min) / (data_max - data_min) - 1


def denormalize(data, data_min, data_max):
    """Inverse of normalize (operations.py:10-13)."""
    if data_max == data_min:
        return np.zeros_like(data_min)
    return (data + 1) / 2 * (data_max - data_min) + data_min


def get_min_max(data, key, config):
    """{key: (min, max)}: x / y ranges come from config['data_test'], every other variable from the
    data itself ignoring NaNs (operations.py:16-30).  `data` may be the array itself or a mapping
    holding it under `key` (the reference passes the array and then indexes it with the key, which
    only works for .mat record arrays; both call styles are accepted here)."""
    if key == 'x':
        return {key: (config['data_test']['x_min'], config['data_test']['x_max'])}
    if key == 'y':
        return {key: (config['data_test']['y_min'], config['data_test']['y_max'])}
    try:
        arr = data[key]
    except (IndexError, KeyError, TypeError, ValueError):
        arr = data
    return {key: (np.nanmin(arr), np.nanmax(arr))}
