"""Drop-in for the reference's defects.py: the defect mask is drawn on the GPU
(csrc/defects.cu) from the same NumPy global stream, one draw per carbon site in C order
(defects.py:18), so a run driven through either implementation consumes the stream identically.

    track_defects(state, atom_type, L, T=None)                          defects.py:4-19
    get_defect_density(defects, voxel_size=5e-6)                        defects.py:21-23
    introduce_defects(state, atom_type, T=None, apply_to_state=False)   defects.py:25-31

`refresh_resident(ctx)` is what `run_kmc` calls every METRIC_UPDATE_STEP steps: it redraws the
mask of the lattice already resident in HBM without moving atom_type / T / the mask across PCIe.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from ._config import constants as K

CARBON_ID = 3              # defects.py:8 hard-codes STATES['C']
E_MIGRATION = 0.3          # defects.py:14


def refresh_resident(ctx, with_T=True, apply_to_state=False):
    """Redraw the defect mask of the lattice resident in `ctx` from NumPy's global stream.
    Returns (n_carbon, n_defects)."""
    n_c = int(ctx.counts()[CARBON_ID])
    if n_c == 0:                                   # defects.py:9 — no draw when there is no carbon site
        return ctx.defects_refresh(draws=np.zeros(0), prob_base=K.DEFECT_PROB_BASE, e_mig=E_MIGRATION, kT=K.K_T,
                                   T_default=K.T_SUB, carbon_id=CARBON_ID, defect_id=K.DEFECT_ID,
                                   apply_to_state=apply_to_state)
    draws = np.random.random(n_c)                  # defects.py:18
    return ctx.defects_refresh(draws=draws, prob_base=K.DEFECT_PROB_BASE, e_mig=E_MIGRATION if with_T else 0.0,
                               kT=K.K_T, T_default=K.T_SUB, carbon_id=CARBON_ID, defect_id=K.DEFECT_ID,
                               apply_to_state=apply_to_state)


def track_defects(state, atom_type, L, T=None, device=0):
    """defects.py:4-19.  The mask follows atom_type == 3, as in the reference."""
    if L == 0:
        return np.zeros((0, 0, 0), dtype=int)
    atom_type = np.asarray(atom_type)
    ctx = _lib.Context(shape=atom_type.shape, device=device)
    try:
        at = np.where((atom_type >= 0) & (atom_type <= 15), atom_type, 0).astype(np.int64)
        ctx.upload(state=at, T=np.ones(atom_type.shape) if T is None else np.asarray(T, dtype=np.float64))
        refresh_resident(ctx, with_T=T is not None)          # T=None: prob = DEFECT_PROB_BASE (defects.py:16)
        packed = ctx.download_packed()
    finally:
        ctx.close()
    return (packed >> 4).astype(int)


def get_defect_density(defects, voxel_size=5e-6):
    """defects.py:21-23"""
    volume = defects.size * (voxel_size ** 3)
    return np.sum(defects) / volume if volume > 0 else 0.0


def introduce_defects(state, atom_type, T=None, apply_to_state=False, voxel_size=5e-6, device=0):
    """defects.py:25-31"""
    L = state.shape[0]
    mask = track_defects(state, atom_type, L, T, device=device)
    if apply_to_state:
        state[mask == 1] = 4
    return mask, get_defect_density(mask, voxel_size)
