"""CPU: host-side logic of the observables / on-disk contracts (cetkmc/metrics.py, campaign.py) that
needs no GPU: the floating-point expressions of metrics.py:41-96 formed from per-grain integers, the
histogram percentile that reproduces the label-volume quirk of Grain_d50/d90, snapshot round trips."""
import json
import os

import numpy as np
import pytest

from conftest import golden


def _grain_dict(clusters, shape):
    """What Context.grains() returns, built from the oracle's DFS clusters."""
    L1, L2 = shape[1], shape[2]
    first = np.array([min((i * L1 + j) * L2 + k for i, j, k in c) for c in clusters], dtype=np.int32)
    order = np.argsort(first)
    cl = [clusters[q] for q in order]
    return dict(n=len(cl), root=first[order], size=np.array([len(c) for c in cl], dtype=np.int32),
                box_lo=np.array([np.min(c, axis=0) for c in cl], dtype=np.int32).reshape(-1, 3),
                box_hi=np.array([np.max(c, axis=0) for c in cl], dtype=np.int32).reshape(-1, 3))


@pytest.mark.parametrize("name", ["grains_grown12.npz", "grains_grown16.npz", "grains_half14.npz"])
def test_metrics_from_grain_integers_match_reference(oracle, name):
    from cetkmc import metrics as M
    g = golden(name)
    st, th, ph = g["state"].astype(np.int64), g["theta"], g["phi"]
    clusters, _ = oracle.get_clusters(st, th, ph, 0.5)
    gd = _grain_dict(clusters, st.shape)
    np.testing.assert_array_equal(gd["root"], g["first"])          # the DFS discovers grains in raster order of their first voxel
    m = M.metrics_from_grains(gd, st.size, defects=g["defects"].astype(np.int64))
    for k in ("AspectRatio", "EquiaxedFraction", "NucleationDensity", "AvgGrainSize", "GrainCount", "DefectDensity"):
        assert m[k] == g[f"m_{k}"], k
    np.testing.assert_allclose([m["Grain_d50_um"], m["Grain_d90_um"]], [g["m_Grain_d50_um"], g["m_Grain_d90_um"]], rtol=1e-12)
    assert M.detect_CET_transition(m) == (str(g["cet"]) == "Equiaxed")
    np.testing.assert_array_equal(M.grain_aspect_ratios(gd), g["aspect"])


def test_metrics_empty_lattice():
    from cetkmc import metrics as M
    m = M.metrics_from_grains(dict(n=0, root=np.zeros(0, np.int32), size=np.zeros(0, np.int32),
                                   box_lo=np.zeros((0, 3), np.int32), box_hi=np.zeros((0, 3), np.int32)), 216)
    assert m["GrainCount"] == 0 and m["AspectRatio"] == 0.0 and m["Grain_d90_um"] == 0.0      # metrics.py:44-55


def test_percentile_of_counts_equals_numpy():
    from cetkmc.metrics import _percentile_of_counts
    rng = np.random.default_rng(3)
    for _ in range(50):
        n = int(rng.integers(1, 12))
        vals = np.sort(rng.random(n)) * 10
        cnts = rng.integers(0, 7, n)
        if cnts.sum() == 0:
            cnts[0] = 1
        full = np.repeat(vals, cnts)
        for q in (0.0, 12.5, 50.0, 90.0, 100.0):
            assert _percentile_of_counts(vals, cnts, q) == pytest.approx(np.percentile(full, q), rel=1e-14, abs=0)


def test_snapshot_round_trip_and_rng_json(tmp_path):
    from cetkmc import campaign
    rng = np.random.default_rng(1)
    L = 5
    st = rng.integers(0, 5, (L, L, L)); th = rng.random((L, L, L)); ph = rng.random((L, L, L)); T = 3000 + rng.random((L, L, L))
    prefix = str(tmp_path / "snap")
    campaign.save_lattice(st, th, ph, T, st, prefix=prefix)
    assert sorted(os.listdir(tmp_path)) == sorted(f"snap_{n}.npy" for n in
                                                  ("state", "orientation_theta", "orientation_phi", "temperature", "atom_type"))
    for a, b in zip((st, th, ph, T, st), campaign.load_lattice(prefix)):
        assert np.array_equal(a, b)
    np.random.seed(11); np.random.random(7)
    j = json.loads(json.dumps(campaign._rng_to_json(np.random.get_state())))
    want = np.random.random(5)
    np.random.seed(0)
    np.random.set_state(campaign._rng_from_json(j))
    assert np.array_equal(np.random.random(5), want)
    G, R, R_phys, gr = campaign._gr(30, 2800, 2e13)                                        # kmc_simulation.py:236-239
    assert G == (3695 - 2800) / (30 * 5e-6) and R == 2e13 * 2.74e-10 / 5e-6 and gr == G / R_phys


def test_gr_sweep_partitions_cases_over_ranks_and_merges(tmp_path, monkeypatch):
    """Config 5 is replicas only: case q belongs to rank q % world; the per-rank maps merge into one
    cet_map.csv ordered by case.  The GPU run of a case is stubbed (host logic only)."""
    import csv
    from cetkmc import campaign
    monkeypatch.chdir(tmp_path)
    calls = []

    def fake_run(L, n_sweeps, temp, nu_dep, output_prefix, device, **kw):
        calls.append((temp, nu_dep, device))
        os.makedirs(f"outputs/{output_prefix}", exist_ok=True)
        with open(f"outputs/{output_prefix}/metrics.csv", "w", newline="") as fh:
            w = csv.DictWriter(fh, fieldnames=["Step", "Time", "AspectRatio", "EquiaxedFraction", "GrainCount", "AvgGrainSize",
                                               "NucleationCount", "CET_Class", "CET_Detected"])
            w.writeheader()
            w.writerow(dict(Step=n_sweeps - 1, Time=1e-9, AspectRatio=temp / 1000, EquiaxedFraction=0.5, GrainCount=7,
                            AvgGrainSize=1.0, NucleationCount=3, CET_Class="Columnar", CET_Detected=False))

    monkeypatch.setattr(campaign, "run_cet_sublattice", fake_run)
    temps, rates = [2600, 2800, 3000], [2e12, 2e13]
    r0 = campaign.run_gr_sweep(temps, rates, L=8, n_sweeps=5, rank=0, world=2, devices=[0])
    r1 = campaign.run_gr_sweep(temps, rates, L=8, n_sweeps=5, rank=1, world=2, devices=[1])
    assert [r["case"] for r in r0] == [0, 2, 4] and [r["case"] for r in r1] == [1, 3, 5]
    assert len(calls) == 6 and {c[2] for c in calls[:3]} == {0} and {c[2] for c in calls[3:]} == {1}
    merged = campaign.merge_cet_map()
    assert [int(r["case"]) for r in merged] == list(range(6))
    assert [float(r["T_sub"]) for r in merged] == [2600, 2600, 2800, 2800, 3000, 3000]
    G = (3695 - 2600) / (8 * 5e-6)
    assert float(merged[0]["G"]) == G and float(merged[0]["G_over_R"]) == G / (2e12 * 2.74e-10 / 5e-6)
    assert os.path.exists("outputs/gr_sweep/cet_map.csv")
