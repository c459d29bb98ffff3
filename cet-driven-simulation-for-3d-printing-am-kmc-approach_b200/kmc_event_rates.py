"""Drop-in for the reference's kmc_event_rates.py (kmc_event_rates.py:10,26,43,162).

`get_event_rates` / `compute_row_events` return the reference's list of
`(bytes, (i,j,k), float, (i,j,k), int)` tuples in the reference's order; the rates are evaluated
by the CUDA rate kernel and the list is materialised only for compatibility / validation —
`run_kmc` never builds it.  `get_bcc_neighbors` and `compute_misorientation` stay scalar host
functions because utils.py:45,54 calls them per neighbour from a Python DFS.
"""
from __future__ import annotations

import math

import numpy as np

from . import _lib
from ._config import constants, rate_params
from ._host import NEIGHBOR_OFFSETS

_species_rng = np.random.RandomState(constants.RANDOM_SEED)
_contexts = {}


def seed_species(seed):
    """Seed the stream that picks the deposited species (kmc_event_rates.py:65).  The reference
    draws it from Numba's private generator, which np.random.seed() does not reach; here it is an
    explicit MT19937 stream (identical to Numba's when that is seeded with the same value)."""
    global _species_rng
    _species_rng = np.random.RandomState(seed)


def species_rng():
    return _species_rng


def compute_misorientation(theta1, phi1, theta2, phi2):
    """kmc_event_rates.py:10-23 — angle between the two orientation unit vectors."""
    s1, s2 = math.sin(theta1), math.sin(theta2)
    dot = (s1 * math.cos(phi1)) * (s2 * math.cos(phi2)) + (s1 * math.sin(phi1)) * (s2 * math.sin(phi2)) \
        + math.cos(theta1) * math.cos(theta2)
    return math.acos(max(min(dot, 1.0), -1.0))


def get_bcc_neighbors(i, j, k, L):
    """kmc_event_rates.py:26-40 — in-bounds subset of the 14 offsets, order kept, (n,3) int64."""
    nb = NEIGHBOR_OFFSETS + np.array([i, j, k], dtype=np.int64)
    return nb[np.all((nb >= 0) & (nb < L), axis=1)]


def _context(L, device=0):
    ctx = _contexts.get((L, device))
    if ctx is None:
        if len(_contexts) >= 2:
            _contexts.pop(next(iter(_contexts))).close()
        ctx = _lib.Context(L=L, device=device)
        _contexts[(L, device)] = ctx
    return ctx


def release():
    while _contexts:
        _contexts.popitem()[1].close()


def events_soa(state, orientation_theta, orientation_phi, T, defects_mask, L, params, draw_species=True):
    """Upload the fields, evaluate, and return the event list as a structure of arrays."""
    ctx = _context(int(L))
    ctx.set_rate_params(params)
    ctx.upload(state=state, theta=orientation_theta, phi=orientation_phi, T=T, defects=defects_mask)
    _n, n_dep = ctx.events_count()
    draws = _species_rng.random_sample(n_dep) if (draw_species and n_dep) else None
    return ctx.events_export(draws)


def _tuples(ev, L):
    LL = L * L
    pos, tgt = ev["pos"], ev["target"]
    pi, pj, pk = (pos // LL).tolist(), ((pos // L) % L).tolist(), (pos % L).tolist()
    has = tgt >= 0
    ti = np.where(has, tgt // LL, -1).tolist()
    tj = np.where(has, (tgt // L) % L, -1).tolist()
    tk = np.where(has, tgt % L, -1).tolist()
    names = np.array(_lib.EV_NAMES, dtype=object)[ev["type"]].tolist()
    # nested zips build the position / target tuples in C: the list of ~3 events per site dominates the call
    return list(zip(names, zip(pi, pj, pk), ev["rate"].tolist(), zip(ti, tj, tk), ev["atom"].tolist()))


def get_event_rates(state, orientation_theta, orientation_phi, T, atom_type, defects_mask, L,
                    states_w, states_re, states_c, step=0, debug_step=1000, impurity_c=0.0):
    """kmc_event_rates.py:162-176.  atom_type, step and debug_step are accepted and unused, as in
    the reference."""
    params = rate_params(impurity_c, states_w, states_re, states_c)
    return _tuples(events_soa(state, orientation_theta, orientation_phi, T, defects_mask, L, params), int(L))


def compute_row_events(i, state, orientation_theta, orientation_phi, T, atom_type, defects, nu, nu_dep,
                       E_b, E_diff, kT, T_melt, I0, delta_T_c, top_layer, L, states_w, states_re,
                       states_c, impurity_c):
    """kmc_event_rates.py:43-160 — events of plane i.  The explicit physics arguments override
    constants.py, as they do in the reference's jitted function."""
    if int(top_layer) != int(L) - 1:
        raise ValueError("compute_row_events: top_layer must be L-1 (the only value the reference passes)")
    ov = dict(NU=nu, NU_DEP=nu_dep, E_B_W=E_b[0], E_B_RE=E_b[1], E_B_C=E_b[2], E_DIFF_W=E_diff[0],
              E_DIFF_RE=E_diff[1], E_DIFF_C=E_diff[2], K_T=kT, T_MELT=T_melt, I0=I0, DELTA_T_C=delta_T_c)
    params = rate_params(impurity_c, states_w, states_re, states_c, overrides=ov)
    L = int(L)
    ev = events_soa(state, orientation_theta, orientation_phi, T, defects, L, params,
                    draw_species=(int(i) == L - 1))
    sel = (ev["pos"] // (L * L)) == int(i)
    return _tuples({k: (v[sel] if isinstance(v, np.ndarray) else v) for k, v in ev.items()}, L)
