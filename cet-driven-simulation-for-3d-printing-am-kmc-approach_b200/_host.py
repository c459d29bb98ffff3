"""Host-side set-up used by `run_kmc` when the user's own modules (`lattice_init`, `defects`) are
not importable: the initial lattice and the initial defect mask.  They are out of the hot path
(SURVEY §2 rows 5-6): O(L^3) NumPy work that runs once per run, before anything is uploaded.

Each function reproduces the observable behaviour of the reference function it names (same
global-RNG draws in the same order, same return types) so that a run driven through either
implementation yields the same trajectory and the same metrics.csv.  Observables (grain
clustering, metrics rows) have no host implementation in the product: csrc/grains.cu + metrics.py.
"""
from __future__ import annotations

import numpy as np

from ._config import constants as K

# kmc_event_rates.py:29-36 — neighbour offsets, reference order
NEIGHBOR_OFFSETS = np.array(
    [(1, 1, 0), (1, -1, 0), (-1, 1, 0), (-1, -1, 0), (0, 1, 1), (0, 1, -1), (0, -1, 1), (0, -1, -1),
     (2, 0, 0), (-2, 0, 0), (0, 2, 0), (0, -2, 0), (0, 0, 2), (0, 0, -2)], dtype=np.int64)


def initialize_lattice(lattice_size=None, n_seeds=None, T_sub=None, T_melt=None, random_seed=42,
                       impurity_c=0.0, verbose=False):
    """lattice_init.py:10-59 — seeds NumPy's global stream and consumes it identically."""
    L = K.LATTICE_SIZE if lattice_size is None else lattice_size
    n_seeds = K.N_SEEDS if n_seeds is None else n_seeds
    T_sub = K.T_SUB if T_sub is None else T_sub
    T_melt = K.T_MELT if T_melt is None else T_melt
    np.random.seed(random_seed)
    state = np.zeros((L, L, L), dtype=int)
    atom_type = np.zeros((L, L, L), dtype=int)
    theta = np.zeros((L, L, L))
    phi = np.zeros((L, L, L))
    grad = (T_melt - T_sub) / L
    T = np.ascontiguousarray(np.broadcast_to(T_sub + grad * np.arange(L), (L, L, L)))
    for flat in np.random.choice(L * L, n_seeds, replace=False):
        x, y = int(flat) // L, int(flat) % L
        r = np.random.random()
        if r < K.IMPURITY_RE:
            atom = 2
        elif r < K.IMPURITY_RE + impurity_c:
            atom = 3
        else:
            atom = 1
        state[x, y, 0] = atom
        atom_type[x, y, 0] = atom
        theta[x, y, 0] = np.random.uniform(0, np.pi)
        phi[x, y, 0] = np.random.uniform(0, 2 * np.pi)
    return state, theta, phi, T, atom_type


def introduce_defects(state, atom_type, T=None, apply_to_state=False, voxel_size=5e-6):
    """defects.py:4-31 — Bernoulli defect mask on carbon sites (one global-stream draw each)."""
    L = state.shape[0]
    mask = np.zeros((L, L, L), dtype=int)
    carbon = (atom_type == 3)
    n_c = int(carbon.sum())
    if n_c:
        if T is not None:
            tv = T[carbon]
            with np.errstate(divide="ignore", invalid="ignore"):
                tv = np.where(tv > 0, tv, K.T_SUB)
                prob = K.DEFECT_PROB_BASE * np.exp(-0.3 / (K.K_T * tv))
        else:
            prob = np.full(n_c, K.DEFECT_PROB_BASE)
        prob = np.clip(prob, 0.0, 1.0)
        mask[carbon] = (np.random.random(n_c) < prob).astype(int)
    if apply_to_state:
        state[mask == 1] = 4
    volume = mask.size * (voxel_size ** 3)
    density = np.sum(mask) / volume if volume > 0 else 0.0
    return mask, density
