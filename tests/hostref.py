"""Test infrastructure: NumPy / SciPy restatement of the reference's observables (utils.py:28-84,
104-111; metrics.py:41-105) — graph connected components instead of the pure-Python DFS, fast enough
for the level-3 ensembles.  It is checked against the reference itself in tests/test_dropin_surface.py
and used as the host-side checker of the GPU clustering; nothing in the product imports it."""
import numpy as np

from cetkmc._config import constants as K
from cetkmc._host import NEIGHBOR_OFFSETS


def _unit_vectors(theta, phi):
    st = np.sin(theta)
    return st * np.cos(phi), st * np.sin(phi), np.cos(theta)


def label_grains(state, theta, phi=None, theta_threshold=0.5):
    """Grain labels = connected components of the graph whose edges join occupied sites that are
    neighbours (14-offset set) with misorientation < theta_threshold (utils.py:28-84).  The
    reference grows them by DFS; components of a symmetric edge relation do not depend on the
    visiting order.  Labels are numbered 1.. in raster order of each grain's first voxel, as the
    reference's DFS discovers them.  Returns (labels int32 (L,L,L), n_grains)."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components

    shape = state.shape
    n = state.size
    occ = state != 0
    idx = np.arange(n, dtype=np.int64).reshape(shape)
    if phi is not None:
        vx, vy, vz = _unit_vectors(theta, phi)
    rows, cols = [], []
    for off in NEIGHBOR_OFFSETS[[0, 1, 4, 5, 8, 10, 12]]:     # one of each +/- pair
        sl_a, sl_b = [], []
        for d, ext in zip(off, shape):
            d = int(d)
            if d >= 0:
                sl_a.append(slice(0, ext - d)); sl_b.append(slice(d, ext))
            else:
                sl_a.append(slice(-d, ext)); sl_b.append(slice(0, ext + d))
        sl_a, sl_b = tuple(sl_a), tuple(sl_b)
        both = occ[sl_a] & occ[sl_b]
        if not both.any():
            continue
        if phi is None:
            mis = np.abs(theta[sl_a] - theta[sl_b])
        else:
            dot = vx[sl_a] * vx[sl_b] + vy[sl_a] * vy[sl_b] + vz[sl_a] * vz[sl_b]
            mis = np.arccos(np.maximum(np.minimum(dot, 1.0), -1.0))
        edge = both & (mis < theta_threshold)
        rows.append(idx[sl_a][edge]); cols.append(idx[sl_b][edge])
    if rows:
        r = np.concatenate(rows); c = np.concatenate(cols)
    else:
        r = c = np.zeros(0, dtype=np.int64)
    graph = coo_matrix((np.ones(r.size, dtype=np.int8), (r, c)), shape=(n, n))
    _, comp = connected_components(graph, directed=False)
    comp = comp.reshape(shape)
    labels = np.zeros(shape, dtype=np.int32)
    occ_flat = occ.ravel()
    comp_occ = comp.ravel()[occ_flat]
    # renumber by first raster occurrence
    uniq, first = np.unique(comp_occ, return_index=True)
    order = np.argsort(first, kind="stable")
    remap = np.empty(uniq.size, dtype=np.int32)
    remap[order] = np.arange(1, uniq.size + 1, dtype=np.int32)
    labels.ravel()[occ_flat] = remap[np.searchsorted(uniq, comp_occ)]
    return labels, int(uniq.size)


def grain_statistics(labels, n_grains):
    """Per-grain voxel count and bounding-box aspect ratio (utils.py:104-111:
    longest / max(shortest, 1) of the box dimensions)."""
    if n_grains == 0:
        return np.zeros(0, dtype=np.int64), np.zeros(0)
    occ = labels > 0
    lab = labels[occ].astype(np.int64) - 1
    sizes = np.bincount(lab, minlength=n_grains)
    coords = np.nonzero(occ)
    dims = []
    for ax in range(3):
        lo = np.full(n_grains, np.iinfo(np.int64).max)
        hi = np.full(n_grains, -1)
        np.minimum.at(lo, lab, coords[ax])
        np.maximum.at(hi, lab, coords[ax])
        dims.append(hi - lo + 1)
    dims = np.stack(dims, axis=1)
    ar = dims.max(axis=1).astype(np.float64) / np.maximum(dims.min(axis=1), 1).astype(np.float64)
    return sizes, ar


def compute_metrics(state, theta, phi, defects=None, voxel_size=None, rng_seed=None):
    """metrics.py:41-96 (the fields run_kmc consumes; mask/grain-id options the driver never
    passes are omitted)."""
    voxel_size = K.VOXEL_SIZE if voxel_size is None else voxel_size
    labels, n = label_grains(state, theta, phi, theta_threshold=0.5)
    if n == 0:
        return {"AspectRatio": 0.0, "EquiaxedFraction": 0.0, "NucleationDensity": 0.0,
                "AvgGrainSize": 0.0, "GrainCount": 0, "DefectDensity": 0.0,
                "Defect_voxel_count": 0, "Defect_voxel_frac": 0.0,
                "Grain_d50_um": 0.0, "Grain_d90_um": 0.0,
                "VOXEL_SIZE_m": voxel_size, "RANDOM_SEED": rng_seed}
    sizes, ar = grain_statistics(labels, n)
    volume = state.size * (voxel_size ** 3)
    def_count = np.sum(defects) if defects is not None else 0
    # metrics.py:43,76 hands the label volume to equivalent_diameter_um
    diam = ((6.0 * (labels * (voxel_size ** 3)) / np.pi) ** (1.0 / 3.0)) * 1e6
    return {
        "AspectRatio": np.mean(ar.tolist()),
        "EquiaxedFraction": np.mean(ar < K.CET_AR_THRESHOLD),
        "NucleationDensity": n / volume if volume > 0 else 0.0,
        "AvgGrainSize": np.mean(sizes.tolist()) * voxel_size * 1e6,
        "GrainCount": n,
        "DefectDensity": def_count / volume if volume > 0 else 0.0,
        "Defect_voxel_count": def_count,
        "Defect_voxel_frac": def_count / state.size if state.size > 0 else 0.0,
        "Grain_d50_um": np.median(diam), "Grain_d90_um": np.percentile(diam, 90),
        "VOXEL_SIZE_m": voxel_size, "RANDOM_SEED": rng_seed,
    }


def detect_CET_transition(m):
    """metrics.py:103-105"""
    return bool(m["AspectRatio"] < K.CET_AR_THRESHOLD and m["EquiaxedFraction"] > K.CET_EQ_THRESHOLD)


def cet_class(m):
    """metrics.py:99-101 applied to an already computed metrics dict."""
    return "Equiaxed" if detect_CET_transition(m) else "Columnar"
