"""Drop-in for the observables the KMC driver takes from the reference's metrics.py / utils.py
(SURVEY §8f row N1): grain clustering runs on the GPU (csrc/grains.cu, a union-find connected-
component labelling over the 14-offset neighbourhood) instead of the pure-Python DFS of
utils.py:28-84, which costs ~1 s at 30^3 and is run twice per metrics row
(kmc_simulation.py:341-377).

    compute_metrics(state, theta, phi, defects=None, ...)   metrics.py:41-96
    compute_CET(state, theta, phi)                          metrics.py:99-101
    detect_CET_transition(metrics_dict)                     metrics.py:103-105
    get_clusters(state, theta, phi, theta_threshold=0.5)    utils.py:69-84
    calculate_aspect_ratio(cluster)                         utils.py:104-111

`compute_metrics(..., ctx=<cetkmc.Context>)` clusters the lattice already resident in HBM (what
`run_kmc` does every METRIC_UPDATE_STEP steps) and needs no upload.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from ._config import constants as K


def _percentile_of_counts(values, counts, q):
    """np.percentile(np.repeat(values, counts), q) (method 'linear') without materialising the
    repeat; values ascending."""
    n = int(np.sum(counts))
    if n == 0:
        return 0.0
    cum = np.cumsum(counts)
    pos = (q / 100.0) * (n - 1)
    lo, hi = int(np.floor(pos)), int(np.ceil(pos))
    a = float(values[np.searchsorted(cum, lo, side="right")])
    b = float(values[np.searchsorted(cum, hi, side="right")])
    t = pos - lo
    return a + (b - a) * t if t < 0.5 else b - (b - a) * (1 - t)


def grain_aspect_ratios(g):
    """utils.py:104-111 per grain: longest / max(shortest, 1) of the bounding-box dimensions."""
    dims = (g["box_hi"].astype(np.int64) - g["box_lo"].astype(np.int64)) + 1
    if dims.shape[0] == 0:
        return np.zeros(0)
    return dims.max(axis=1).astype(np.float64) / np.maximum(dims.min(axis=1), 1).astype(np.float64)


def metrics_from_grains(g, n_sites, defects=None, voxel_size=None, rng_seed=None):
    """metrics.py:41-96 from the per-grain statistics of Context.grains()."""
    voxel_size = K.VOXEL_SIZE if voxel_size is None else voxel_size
    n = int(g["n"])
    if n == 0:
        return {"AspectRatio": 0.0, "EquiaxedFraction": 0.0, "NucleationDensity": 0.0,
                "AvgGrainSize": 0.0, "GrainCount": 0, "DefectDensity": 0.0,
                "Frac_W": 0.0, "Frac_Re": 0.0, "Frac_C": 0.0,
                "C_boundary_frac": 0.0, "Re_boundary_frac": 0.0,
                "Defect_voxel_count": 0, "Defect_voxel_frac": 0.0,
                "Grain_d50_um": 0.0, "Grain_d90_um": 0.0,
                "VOXEL_SIZE_m": voxel_size, "RANDOM_SEED": rng_seed}
    ar = grain_aspect_ratios(g)
    sizes = g["size"].astype(np.int64)
    volume = n_sites * (voxel_size ** 3)
    def_count = np.sum(defects) if defects is not None else 0
    # metrics.py:43,76 hands the `visited` label volume (not the sizes) to equivalent_diameter_um:
    # value 0 on the empty sites, grain number q (1..n) on the sizes[q-1] voxels of grain q
    vals = np.arange(0, n + 1, dtype=np.float64)
    cnts = np.concatenate(([n_sites - int(sizes.sum())], sizes))
    diam = ((6.0 * (vals * (voxel_size ** 3)) / np.pi) ** (1.0 / 3.0)) * 1e6
    return {
        "AspectRatio": np.mean(ar.tolist()),
        "EquiaxedFraction": np.mean(ar < K.CET_AR_THRESHOLD),
        "NucleationDensity": n / volume if volume > 0 else 0.0,
        "AvgGrainSize": np.mean(sizes.tolist()) * voxel_size * 1e6,
        "GrainCount": n,
        "DefectDensity": def_count / volume if volume > 0 else 0.0,
        "Frac_W": 0.0, "Frac_Re": 0.0, "Frac_C": 0.0,
        "C_boundary_frac": 0.0, "Re_boundary_frac": 0.0,
        "Defect_voxel_count": def_count,
        "Defect_voxel_frac": def_count / n_sites if n_sites > 0 else 0.0,
        "Grain_d50_um": _percentile_of_counts(diam, cnts, 50.0),
        "Grain_d90_um": _percentile_of_counts(diam, cnts, 90.0),
        "VOXEL_SIZE_m": voxel_size, "RANDOM_SEED": rng_seed,
    }


def _resident(state, theta, phi, device=0):
    state = np.asarray(state)
    if state.ndim != 3 or len(set(state.shape)) != 1:
        raise ValueError("grain clustering needs a cubic (L, L, L) lattice")
    ctx = _lib.Context(L=state.shape[0], device=device)
    ctx.upload(state=state, theta=theta, phi=np.zeros_like(np.asarray(theta, dtype=np.float64)) if phi is None else phi)
    return ctx


def compute_metrics(state, theta, phi, defects=None, voxel_size=None, W_mask=None, Re_mask=None,
                    C_mask=None, grain_ids=None, rng_seed=None, ctx=None, device=0):
    """metrics.py:41-96.  With `ctx` the lattice resident in that context is clustered (state /
    theta / phi are then only used for `state.size`)."""
    own = ctx is None
    if own:
        ctx = _resident(state, theta, phi, device)
    try:
        g = ctx.grains(0.5)
    finally:
        if own:
            ctx.close()
    n_sites = int(np.prod(ctx.owned_shape)) if state is None else int(np.asarray(state).size)
    m = metrics_from_grains(g, n_sites, defects=defects, voxel_size=voxel_size, rng_seed=rng_seed)
    if m["GrainCount"]:
        for key, mask in (("Frac_W", W_mask), ("Frac_Re", Re_mask), ("Frac_C", C_mask)):
            if mask is not None:
                m[key] = np.count_nonzero(mask) / n_sites if n_sites > 0 else 0.0
    return m


def detect_CET_transition(metrics_dict):
    """metrics.py:103-105"""
    return (metrics_dict["AspectRatio"] < K.CET_AR_THRESHOLD and
            metrics_dict["EquiaxedFraction"] > K.CET_EQ_THRESHOLD)


def compute_CET(state, theta, phi, voxel_size=None, ctx=None):
    """metrics.py:99-101"""
    m = compute_metrics(state, theta, phi, voxel_size=voxel_size, ctx=ctx)
    return "Equiaxed" if detect_CET_transition(m) else "Columnar"


def get_clusters(state, orientation_theta, orientation_phi=None, theta_threshold=0.5, device=0):
    """utils.py:69-84: (clusters, visited).  Clusters come in the reference's order (raster order
    of their first voxel) and `visited` is the reference's label volume; inside a cluster the
    voxels are listed in raster order (the reference lists them in DFS order; its consumers —
    len() and the bounding box — do not depend on it).  orientation_phi=None selects the
    reference's |theta1 - theta2| criterion (utils.py:49-50), which only the unused
    detect_CET_transition of utils.py:117 takes."""
    state = np.asarray(state)
    if state.size == 0:
        return [], np.array([])
    ctx = _resident(state, orientation_theta, orientation_phi, device)
    try:
        g = ctx.grains(theta_threshold, labels=True, theta_only=orientation_phi is None)
    finally:
        ctx.close()
    visited, n = g["labels"], g["n"]
    flat = visited.ravel()
    occ = np.flatnonzero(flat)
    order = np.argsort(flat[occ], kind="stable")
    coords = np.stack(np.unravel_index(occ[order], visited.shape), axis=1)
    bounds = np.searchsorted(flat[occ][order], np.arange(1, n + 2))
    clusters = [[tuple(map(int, v)) for v in coords[bounds[q]:bounds[q + 1]]] for q in range(n)]
    return clusters, visited


def calculate_aspect_ratio(cluster):
    """utils.py:104-111"""
    coords = np.array(cluster)
    dims = coords.max(axis=0) - coords.min(axis=0) + 1
    return float(np.max(dims)) / float(max(np.min(dims), 1))
