"""Drop-in for the reference's historical `physics_functions` module, which the reference ships only as stale bytecode
(__pycache__/physics_functions.cpython-38.pyc, source /home/reza/Projects/AI/main/physics_functions.py; decompiled with
tools/pyc38_decompile.py):

    compute_gradient(pred, var)                       :5-15
    Boussinesq_simple(output, t, x, y, device)        :18-52   first-order residual (continuity with h u, h v; momentum)
    Boussinesq(output, t, x, y, device)               :55-130  fully nonlinear: nested compute_gradient three levels deep

`output` is the [N,4] DNN output (h, z, u, v in columns 0..3), t / x / y the differentiated input columns.  Both
residuals return a 0-d differentiable loss, evaluated by ONE fused launch: `Boussinesq_simple` by the first-order jet
kernels, `Boussinesq` by the third-order Taylor-jet kernel (csrc/jet3.cu: 20 Taylor coefficients per feature, the
residual in polynomial arithmetic, the weight gradient by the reverse of the same program) -- instead of the several
hundred autograd sweeps the nested calls cost the reference.
"""
from __future__ import annotations

from .physics import _fused, compute_gradient  # noqa: F401


def Boussinesq_simple(output, t, x, y, device=None):
    """physics_functions.py:18-52."""
    return _fused("Boussinesq_simple", (t, x, y), (output,))


def Boussinesq(output, t, x, y, device=None):
    """physics_functions.py:55-130."""
    return _fused("Boussinesq", (t, x, y), (output,))
