"""Synthetic lattices of the benchmark shapes.  Product-side generators used by
bench.py and the examples; they never touch the oracle."""
from __future__ import annotations

import numpy as np


def half_grown(L, seed=1234, grain=8, planes=None, fill=0.5, n0=None):
    """'Half-grown' lattice: salt-and-pepper solid (W/Re/C/defect = .85/.10/.04/.01, `fill` of the
    sites) below a wavy front k < L/2 + 8 sin(2 pi i / L), empty above; orientations constant over
    grain^3 blocks; linear gradient T = 2800 + 895 k / L; defect flag on 5 % of the C sites.

    Returns (packed uint8 [state | defects << 4], theta, phi, T) for planes [p0, p1) of axis 0
    (default: all).  Every plane is generated from its own seeded stream, so a slab of a
    distributed lattice equals the same planes of the full one."""
    n0 = L if n0 is None else n0          # planes along axis 0 (stacked slabs when > L)
    p0, p1 = (0, n0) if planes is None else planes
    n = p1 - p0
    packed = np.empty((n, L, L), dtype=np.uint8)
    theta = np.empty((n, L, L), dtype=np.float64)
    phi = np.empty((n, L, L), dtype=np.float64)
    k = np.arange(L)
    T = np.ascontiguousarray(np.broadcast_to(2800.0 + 895.0 * k / L, (n, L, L)))
    g = (L + grain - 1) // grain
    for i in range(p0, p1):
        rng = np.random.default_rng([seed, i])
        front = L / 2 + 8 * np.sin(2 * np.pi * i / L)
        u = rng.random((L, L))
        sp = rng.choice(np.array([1, 2, 3, 4], dtype=np.uint8), size=(L, L), p=[.85, .10, .04, .01])
        st = np.where((u < fill) & (k[None, :] < front), sp, 0).astype(np.uint8)
        dfl = ((st == 3) & (rng.random((L, L)) < 0.05)).astype(np.uint8)
        packed[i - p0] = st | (dfl << 4)
        grng = np.random.default_rng([seed, 1 << 20, i // grain])     # one stream per grain layer
        tb = grng.uniform(0, np.pi, (g, g)); pb = grng.uniform(0, 2 * np.pi, (g, g))
        solid = (st >= 1) & (st <= 3)
        theta[i - p0] = np.where(solid, np.repeat(np.repeat(tb, grain, 0), grain, 1)[:L, :L], 0.0)
        phi[i - p0] = np.where(solid, np.repeat(np.repeat(pb, grain, 0), grain, 1)[:L, :L], 0.0)
    return packed, theta, phi, T


def unpack(packed):
    """(state int64, defects int64) from the packed byte."""
    return (packed & 0x0F).astype(np.int64), (packed >> 4).astype(np.int64)
