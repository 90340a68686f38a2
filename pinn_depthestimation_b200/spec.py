"""Description of one loss-evaluation pass, in the reference's own vocabulary.

`kind` is the name of the reference's physics.py function; `dirs` / `fields` name the physics
function's arguments.  `PassSpec.to_desc()` fills the C `pinn_desc_t`.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

from . import _cabi

KINDS = {
    "none": _cabi.RES_NONE,
    "continuity_only": _cabi.RES_CONT_ONLY,     # physics.py:18-33
    "continuity_ftemp": _cabi.RES_CONT_FTEMP,   # physics.py:37-47
    "Navier_Stokes": _cabi.RES_NSWE,            # physics.py:50-88
    "physics_equation": _cabi.RES_WAVE_AVG,     # physics.py:91-120
    "external": _cabi.RES_EXTERNAL,
    # the historical physics_functions module (bytecode only; decompiled)
    "Boussinesq": _cabi.RES_BOUSSINESQ,            # physics_functions.py:55-130 (third-order input derivatives)
    "Boussinesq_simple": _cabi.RES_BOUSS_SIMPLE,   # physics_functions.py:18-52
}
# argument order of the physics functions = order of dir_cols / field_cols in the C desc
DIR_ORDER = {
    "continuity_only": ("x", "y"), "continuity_ftemp": ("x", "y"),
    "Navier_Stokes": ("t", "x", "y"), "physics_equation": ("x", "y"),
    "Boussinesq": ("t", "x", "y"), "Boussinesq_simple": ("t", "x", "y"),
}
FIELD_ORDER = {
    "continuity_only": ("h", "U", "V"), "continuity_ftemp": ("h", "U", "V"),
    "Navier_Stokes": ("h", "z", "u", "v"),
    "physics_equation": ("h", "U", "V", "eta_mean", "Hrms", "k"),
    "Boussinesq": ("h", "z", "u", "v"), "Boussinesq_simple": ("h", "z", "u", "v"),
}


def param_count(layers):
    return sum(layers[i] * layers[i + 1] + layers[i + 1] for i in range(len(layers) - 1))


@dataclass
class PassSpec:
    layers: List[int]
    activation: str = "tanh"                 # 'tanh' (init_type 'xavier') | 'leaky_relu' ('kaiming')
    kind: str = "none"
    dirs: Dict[str, int] = field(default_factory=dict)     # direction name -> input column
    fields: Dict[str, int] = field(default_factory=dict)   # field name -> output column
    mask_col: Optional[int] = None           # column passed as `x` to continuity_only
    cond_threshold: float = 25.5             # physics.py:27
    cond_value: float = 0.75                 # physics.py:28
    target_cols: List[int] = field(default_factory=list)
    target_w: Optional[List[float]] = None
    w_fid: float = 1.0
    w_res: float = 1.0
    precision: str = "fp32"
    ext_dirs: List[int] = field(default_factory=list)      # kind == 'external': diff. columns

    @property
    def n_params(self):
        return param_count(self.layers)

    def to_desc(self) -> "_cabi.Desc":
        if self.kind not in KINDS:
            raise ValueError(f"unknown residual kind {self.kind!r}")
        if self.activation not in _cabi.ACT:
            raise ValueError(f"unknown activation {self.activation!r}")
        if len(self.layers) - 1 > _cabi.MAX_LINEAR:
            raise ValueError("too many layers")
        d = _cabi.Desc()
        d.n_linear = len(self.layers) - 1
        for i, w in enumerate(self.layers):
            d.widths[i] = int(w)
        d.activation = _cabi.ACT[self.activation]
        d.residual_kind = KINDS[self.kind]
        if self.kind in DIR_ORDER:
            names = DIR_ORDER[self.kind]
            d.n_dirs = len(names)
            for j, n in enumerate(names):
                d.dir_cols[j] = int(self.dirs[n])
            for f, n in enumerate(FIELD_ORDER[self.kind]):
                d.field_cols[f] = int(self.fields[n])
        elif self.kind == "external":
            d.n_dirs = len(self.ext_dirs)
            for j, c in enumerate(self.ext_dirs):
                d.dir_cols[j] = int(c)
        mc = self.mask_col if self.mask_col is not None else self.dirs.get("x", 0)
        d.mask_col = int(mc)
        d.cond_threshold = float(self.cond_threshold)
        d.cond_value = float(self.cond_value)
        d.n_targets = len(self.target_cols)
        tw = self.target_w if self.target_w is not None else [1.0] * len(self.target_cols)
        for i, c in enumerate(self.target_cols):
            d.target_cols[i] = int(c)
            d.target_w[i] = float(tw[i])
        d.w_fid = float(self.w_fid)
        d.w_res = float(self.w_res)
        d.precision = _cabi.PREC[self.precision]
        return d
