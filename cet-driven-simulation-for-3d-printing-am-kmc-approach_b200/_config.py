"""Physical constants and kernel parameter blocks.

The reference keeps every constant as a module global of `constants.py` and imports them by
name (kmc_event_rates.py:3-7, kmc_simulation.py:197-201, thermal_solver.py:3).  The drop-in does
the same: if a `constants` module is importable (the user's own, sitting beside main.py) its
values are used, so edits to constants.py keep working.  Otherwise the reference's shipped
defaults below apply (values cited from constants.py).
"""
from __future__ import annotations

import importlib

from . import _lib

# constants.py line numbers in comments
DEFAULTS = {
    "LATTICE_SIZE": 30,        # :54
    "VOXEL_SIZE": 5e-6,        # :55
    "N_STEPS": 20000,          # :56
    "METRIC_UPDATE_STEP": 200,  # :57
    "N_SEEDS": 20,             # :59
    "K_T": 8.617333262e-5,     # :64
    "T_MELT": 3695,            # :65
    "T_SUB": 2800,             # :66
    "ATOMIC_SPACING_W": 2.74e-10,  # :68
    "NU": 1e13,                # :71
    "NU_DEP": 2e13,            # :72
    "E_B_W": 3.8, "E_DIFF_W": 0.35,                          # :75-76
    "E_B_RE": 4.2, "E_DIFF_RE": 0.50, "IMPURITY_RE": 0.10,   # :81-83
    "E_B_C": 3.2, "E_DIFF_C": 0.30,                          # :85-86
    "MAX_IMP_FRACTION": 1.0,   # :91
    "ANISOTROPY_FACTOR": 0.25,  # :104
    "CET_EQ_THRESHOLD": 0.50, "CET_AR_THRESHOLD": 3.0,       # :109-110
    "CET_CHECK_INTERVAL": 100,  # :115
    "DELTA_T_C": 10,           # :125
    "I0": 5e13, "K_NUC": 500, "BETA_IMP_NUC": 0.4,           # :131-133
    "DEFECT_PROB": 3e-3, "DEFECT_PROB_BASE": 0.12,           # :139-140
    "RATE_THRESHOLD": 1e-30,   # :146
    "RANDOM_SEED": 42,         # :148
    "DEFECT_ID": 4,            # :43
}


class _Constants:
    """Attribute view: user's `constants` module first, shipped defaults second."""

    def __init__(self):
        self._mod = None
        try:
            mod = importlib.import_module("constants")
            if hasattr(mod, "T_MELT") and hasattr(mod, "K_T"):
                self._mod = mod
        except Exception:
            self._mod = None

    def __getattr__(self, name):
        mod = object.__getattribute__(self, "_mod")
        if mod is not None and hasattr(mod, name):
            return getattr(mod, name)
        try:
            return DEFAULTS[name]
        except KeyError:
            raise AttributeError(name) from None

    @property
    def source(self):
        return getattr(self._mod, "__file__", None) or "built-in defaults"


constants = _Constants()

# thermal_solver.py:6-13
K = 173.0
RHO = 19300.0
CP = 132.0
ALPHA = K / (RHO * CP)
DEFAULT_BEAM_RADIUS = 50e-6
DEFAULT_ABSORPTIVITY = 0.35


def rate_params(impurity_c=0.0, states_w=1, states_re=2, states_c=3, overrides=None) -> _lib.RateParams:
    """cet_rate_params from constants.py (kmc_event_rates.py:164-173 + the module globals the
    jitted code closes over)."""
    c = constants
    get = (lambda n: overrides[n] if overrides and n in overrides else getattr(c, n))
    p = _lib.RateParams()
    p.nu, p.nu_dep = get("NU"), get("NU_DEP")
    p.E_b[:] = [get("E_B_W"), get("E_B_RE"), get("E_B_C")]
    p.E_diff[:] = [get("E_DIFF_W"), get("E_DIFF_RE"), get("E_DIFF_C")]
    p.kT, p.T_melt, p.i0, p.delta_T_c = get("K_T"), get("T_MELT"), get("I0"), get("DELTA_T_C")
    p.k_nuc, p.beta_imp_nuc = get("K_NUC"), get("BETA_IMP_NUC")
    p.max_imp_fraction = get("MAX_IMP_FRACTION")
    p.rate_threshold, p.anisotropy = get("RATE_THRESHOLD"), get("ANISOTROPY_FACTOR")
    p.impurity_re, p.impurity_c = get("IMPURITY_RE"), float(impurity_c)
    p.states_w, p.states_re, p.states_c = int(states_w), int(states_re), int(states_c)
    p.defect_id = 4                                   # kmc_event_rates.py:80 hard-codes 4
    return p


def thermal_params(dt=1e-6, nan_to_num=False, t_floor=None) -> _lib.ThermalParams:
    """cet_thermal_params with Python's exact doubles (thermal_solver.py:114-117).  t_floor: lower clip
    and NaN fill of a run whose substrate temperature is not constants.T_SUB (the G-R sweep): the
    reference clips to the constant, which would lift a colder substrate to T_SUB at the first step and
    change the gradient the run reports."""
    c = constants
    tp = _lib.ThermalParams()
    tp.dt_alpha = dt * ALPHA
    tp.inv_dx2 = 1.0 / (c.VOXEL_SIZE * c.VOXEL_SIZE)
    tp.lo = float(c.T_SUB if t_floor is None else min(t_floor, c.T_SUB))
    tp.hi = c.T_MELT * 1.1
    tp.nan_value = tp.lo                              # kmc_simulation.py:249
    tp.nan_to_num = 1 if nan_to_num else 0
    return tp


def thermal_full_params(dt) -> _lib.ThermalFullParams:
    c = constants
    p = _lib.ThermalFullParams()
    p.dt = float(dt)
    p.alpha = ALPHA
    p.inv_dx2 = 1.0 / (c.VOXEL_SIZE * c.VOXEL_SIZE)
    p.rho_cp = RHO * CP
    p.latent_over_cp = 200e3 / CP                     # thermal_solver.py:102
    p.lo = float(c.T_SUB)
    p.hi = c.T_MELT * 1.1
    return p
