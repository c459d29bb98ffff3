"""CPU: the C oracle (oracle/oracle.c) against the fixtures generated from the reference
(oracle/gen_golden.py).  Bit-exact: same libm, same evaluation order, no FMA contraction."""
import ast
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden

RATE_CASES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "rates_*.npz")))
TRAJ_CASES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "traj_*.npz")))


def test_fixtures_present():
    assert len(RATE_CASES) >= 5 and len(TRAJ_CASES) >= 3


@pytest.mark.parametrize("name", RATE_CASES)
def test_event_list_matches_reference(oracle, name):
    g = golden(name)
    L = g["state"].shape[0]
    p = oracle.make_params(float(g["impurity_c"]))
    draws = np.random.RandomState(int(g["species_seed"])).random_sample(L * L)
    ev = oracle.event_rates(g["state"], g["theta"], g["phi"], g["T"], g["defects"], L, p, draws)
    assert ev["type"].size == g["ev_type"].size
    np.testing.assert_array_equal(ev["type"], g["ev_type"])
    np.testing.assert_array_equal(ev["pos"], g["ev_pos"])
    np.testing.assert_array_equal(ev["target"], g["ev_target"])
    np.testing.assert_array_equal(ev["atom"], g["ev_atom"])
    np.testing.assert_array_equal(ev["rate"], g["ev_rate"])          # bit-exact
    assert oracle.pysum(ev["rate"]) == float(g["total_pysum"])


def test_thermal_cet_sequences(oracle):
    for name, steps in (("thermal_cet_random.npz", (1, 2, 5, 40)), ("thermal_cet_gradient12.npz", (1, 3, 10, 60))):
        g = golden(name)
        T = g["T0"]
        for n in range(1, max(steps) + 1):
            T = oracle.thermal_cet(T)
            if n in steps:
                np.testing.assert_array_equal(T, g[f"T_after_{n}"])


def test_thermal_full(oracle):
    g = golden("thermal_full10.npz")
    T1 = oracle.thermal_full(g["T0"], g["state"], g["prev_state"], 1e-7, (3, 4.5), 200.0)
    np.testing.assert_array_equal(T1, g["T1"])
    T2 = oracle.thermal_full(g["T0"], g["state"], g["prev_state"], 1e-9, (0, 2), 50.0, beam_radius=20e-6,
                             absorptivity=0.5)
    np.testing.assert_array_equal(T2, g["T2"])


@pytest.mark.parametrize("name", TRAJ_CASES)
def test_trajectory_matches_reference(oracle, name):
    g = golden(name)
    kw = ast.literal_eval(str(g["kwargs"]))
    state, atom_type, total_time, theta, phi, info = oracle.run_kmc(
        L=kw["L"], n_steps=kw["n_steps"], temp=kw["temp"], defect_fraction=kw["defect_fraction"],
        n_seeds=kw["n_seeds"], impurity_c=kw["impurity_c"], seed=42, species_seed=42)
    np.testing.assert_array_equal(state, g["state"])
    np.testing.assert_array_equal(atom_type, g["atom_type"])
    np.testing.assert_array_equal(theta, g["theta"])
    np.testing.assert_array_equal(phi, g["phi"])
    assert total_time == float(g["total_time"])
    assert info["nucleation_count"] == int(g["csv_NucleationCount"][-1])


def test_neighbors_and_misorientation(oracle):
    nb = oracle.bcc_neighbors(0, 0, 0, 5)
    assert nb.tolist() == [[1, 1, 0], [0, 1, 1], [2, 0, 0], [0, 2, 0], [0, 0, 2]]
    assert oracle.bcc_neighbors(2, 2, 2, 5).shape == (14, 3)
    assert oracle.bcc_neighbors(0, 0, 0, 1).shape == (0, 3)
    assert oracle.misorientation(0.3, 1.0, 0.3, 1.0) < 1e-7
    assert abs(oracle.misorientation(0.0, 0.0, np.pi / 2, 0.0) - np.pi / 2) < 1e-15


@pytest.mark.parametrize("name", ["grains_grown12.npz", "grains_grown16.npz", "grains_half14.npz"])
def test_grain_clustering_golden(oracle, name):
    """utils.get_clusters / metrics.compute_metrics restatement against the reference's output."""
    g = golden(name)
    st, th, ph = g["state"].astype(np.int64), g["theta"], g["phi"]
    clusters, visited = oracle.get_clusters(st, th, ph, 0.5)
    np.testing.assert_array_equal(visited, g["visited"])
    np.testing.assert_array_equal([len(c) for c in clusters], g["sizes"])
    np.testing.assert_array_equal([oracle.calculate_aspect_ratio(c) for c in clusters], g["aspect"])
    m = oracle.compute_metrics(st, th, ph, defects=g["defects"].astype(np.int64))
    for k in ("AspectRatio", "EquiaxedFraction", "NucleationDensity", "AvgGrainSize", "GrainCount", "DefectDensity",
              "Grain_d50_um", "Grain_d90_um"):
        assert m[k] == g[f"m_{k}"], k


@pytest.mark.parametrize("name", ["defects_c20.npz", "defects_c35.npz"])
def test_defect_mask_golden(oracle, name):
    """defects.py:4-31 restatement against the reference's own masks (NumPy global stream seeded)."""
    g = golden(name)
    st, T, seed = g["state"].astype(np.int64), g["T"], int(g["seed"])
    rs = np.random.RandomState(seed)
    np.testing.assert_array_equal(oracle.track_defects(st, T, rs), g["mask"])
    assert rs.random_sample() == float(g["next_draw"])                    # one draw per carbon site, no more
    assert g["mask"].sum() / (g["mask"].size * (5e-6) ** 3) == float(g["density"])
    u = np.random.RandomState(seed).random_sample(int((st == 3).sum()))
    np.testing.assert_array_equal(g["mask_noT"][st == 3], (u < 0.12).astype(np.int8))      # T=None: flat probability


@pytest.mark.parametrize("name", ["grains_grown12.npz", "grains_grown16.npz", "grains_half14.npz"])
def test_grain_clustering_c_restatement_golden(oracle, name):
    """oracle.c's DFS (used for lattices the pure-Python one is too slow for) against the reference."""
    g = golden(name)
    f = oracle.clusters_fast(g["state"].astype(np.int64), g["theta"], g["phi"], 0.5)
    np.testing.assert_array_equal(f["visited"], g["visited"])
    np.testing.assert_array_equal(f["size"], g["sizes"])
    dims = f["box_hi"] - f["box_lo"] + 1
    np.testing.assert_array_equal(dims.max(1) / np.maximum(dims.min(1), 1), g["aspect"])
