"""Drop-in for the reference's thermal_solver.py: same public names and signatures
(thermal_solver.py:6-13,15,36,107), the stencil runs on the GPU through libcetkmc.

    update_temperature_cet(T, state, dt=1e-6)   -> new (n0,n1,n2) float64 array
    update_temperature(T, state, prev_state, dt, laser_pos, laser_power, beam_radius, absorptivity)
    build_temperature_field(L=None)

Inputs are caller-owned NumPy arrays and are not modified; a new array is returned, as in the
reference.  Device contexts are cached per shape.  For a lattice that stays resident in HBM
across steps use `cetkmc.Context.thermal_cet` directly (that is what run_kmc does).
"""
from __future__ import annotations

import numpy as np

from . import _lib
from ._config import (ALPHA, CP, DEFAULT_ABSORPTIVITY, DEFAULT_BEAM_RADIUS, K, RHO,  # noqa: F401
                      constants, thermal_full_params, thermal_params)

_contexts = {}


def _context(shape, device=0):
    key = (tuple(int(x) for x in shape), device)
    ctx = _contexts.get(key)
    if ctx is None:
        if len(_contexts) >= 4:                      # keep HBM use bounded
            _contexts.pop(next(iter(_contexts))).close()
        ctx = _lib.Context(shape=key[0], device=device)
        _contexts[key] = ctx
    return ctx


def release():
    """Free the cached device contexts."""
    while _contexts:
        _contexts.popitem()[1].close()


def build_temperature_field(L: int = None) -> np.ndarray:
    """thermal_solver.py:15-34 — T[i,:,:] = T_SUB + (T_MELT-T_SUB)/(L-1) * i, filled on the GPU."""
    if L is None:
        L = constants.LATTICE_SIZE
    g = (constants.T_MELT - constants.T_SUB) / (L - 1) if L > 1 else 0.0
    ctx = _context((L, L, L))
    ctx.fill_gradient(float(constants.T_SUB), float(g))
    return ctx.download(T=True)["T"]


def update_temperature_cet(T: np.ndarray, state: np.ndarray, dt: float = 1e-6) -> np.ndarray:
    """thermal_solver.py:107-117 — `state` is accepted and unused, as in the reference."""
    T = np.asarray(T)
    if T.ndim != 3:
        raise ValueError("T must be a 3-D array")
    ctx = _context(T.shape)
    ctx.upload(T=T)
    ctx.thermal_cet(thermal_params(dt))
    return ctx.download(T=True)["T"]


def laser_source_top(L, laser_pos, laser_power, beam_radius, absorptivity):
    """I_surface / VOXEL_SIZE on the top plane (thermal_solver.py:80-94).  Host NumPy: L^2 values,
    evaluated with the reference's expression so the doubles are identical (note the reference
    uses j0 for both in-plane offsets and ignores i0)."""
    _i0, j0 = laser_pos
    ax = np.arange(L, dtype=np.float64)
    JJ, KK = np.meshgrid(ax, ax, indexing="ij")
    r_m = np.sqrt((JJ - j0) ** 2 + (KK - j0) ** 2) * constants.VOXEL_SIZE
    area_norm = np.pi * beam_radius * beam_radius
    I_surface = (laser_power * absorptivity / area_norm) * np.exp(-(r_m ** 2) / (beam_radius ** 2))
    return I_surface / constants.VOXEL_SIZE


def update_temperature(T: np.ndarray, state: np.ndarray, prev_state: np.ndarray, dt: float,
                       laser_pos: tuple, laser_power: float,
                       beam_radius: float = DEFAULT_BEAM_RADIUS,
                       absorptivity: float = DEFAULT_ABSORPTIVITY) -> np.ndarray:
    """thermal_solver.py:36-105 — stencil + Gaussian top-plane source + latent heat of the
    sites that solidified between prev_state and state."""
    L = T.shape[0]
    assert T.shape == (L, L, L)
    assert state.shape == T.shape and prev_state.shape == T.shape
    ctx = _context(T.shape)
    ctx.upload(state=state, T=T)
    ctx.upload_prev_state(prev_state)
    ctx.thermal_full(thermal_full_params(dt), laser_source_top(L, laser_pos, laser_power, beam_radius, absorptivity))
    return ctx.download(T=True)["T"]
