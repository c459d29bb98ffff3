"""GPU: the large-lattice drivers and on-disk contracts (cetkmc/campaign.py)."""
import csv
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

COLUMNS = ["Step", "Time", "AspectRatio", "EquiaxedFraction", "NucleationDensity", "DefectDensity", "AvgGrainSize",
           "GrainCount", "W_Count", "Re_Count", "C_Count", "NucleationCount", "G_over_R", "G_phys", "R_phys",
           "G_over_R_phys", "CET_Class", "CET_Detected"]          # kmc_simulation.py:359-378


def _rows(path):
    with open(path) as fh:
        return list(csv.DictReader(fh))


def test_run_cet_sublattice_csv_and_resume(cet, oracle, tmp_path, monkeypatch):
    from cetkmc import campaign
    monkeypatch.chdir(tmp_path)
    kw = dict(L=20, temp=2800, defect_fraction=3e-3, n_seeds=8, impurity_c=0.2, metrics_every=100,
              events_per_sweep=0.004 * 20 ** 3, verbose=False)
    a = campaign.run_cet_sublattice(n_sweeps=401, output_prefix="straight", **kw)
    rows = _rows("outputs/straight/metrics.csv")
    assert list(rows[0].keys()) == COLUMNS                       # the reference's 18 columns, in order
    assert [int(r["Step"]) for r in rows] == [0, 100, 200, 300, 400]
    st = a[0]
    assert int(rows[-1]["W_Count"]) == int((st == 1).sum()) and int(rows[-1]["C_Count"]) == int((st == 3).sum())
    # the last row's observables are those of the returned lattice (oracle DFS restatement)
    m = oracle.compute_metrics(a[0], a[3], a[4])
    assert int(rows[-1]["GrainCount"]) == m["GrainCount"] and float(rows[-1]["AspectRatio"]) == m["AspectRatio"]
    assert float(rows[-1]["Time"]) == a[2] and a[2] > 0
    # stop after 201 sweeps with a checkpoint, resume to 401: identical trajectory and rows
    campaign.run_cet_sublattice(n_sweeps=201, output_prefix="part", checkpoint_every=1, **kw)
    b = campaign.run_cet_sublattice(n_sweeps=401, output_prefix="part", resume_from="outputs/part/checkpoint", **kw)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert _rows("outputs/part/metrics.csv") == rows


def test_snapshot_files_follow_the_reference_layout(cet, tmp_path):
    from cetkmc import campaign
    rng = np.random.default_rng(0)
    L = 6
    st = rng.integers(0, 4, (L, L, L)); th = rng.random((L, L, L)); ph = rng.random((L, L, L))
    T = 3000 + rng.random((L, L, L))
    prefix = str(tmp_path / "init")
    campaign.save_lattice(st, th, ph, T, st, prefix=prefix)
    for name in ("state", "orientation_theta", "orientation_phi", "temperature", "atom_type"):   # lattice_init.py:98-105
        assert os.path.exists(f"{prefix}_{name}.npy")
    back = campaign.load_lattice(prefix)
    for x, y in zip((st, th, ph, T, st), back):
        assert np.array_equal(x, y)
    assert back[0].dtype == np.int64 and back[1].dtype == np.float64


def test_impurity_prefix_writes_the_name_plot_cet_globs(cet, tmp_path, monkeypatch):
    from cetkmc import kmc_simulation as ks
    monkeypatch.chdir(tmp_path)
    ks.run_kmc(L=8, n_steps=3, n_seeds=3, impurity_c=0.1, output_prefix="impurity_c_10")
    assert os.path.exists("outputs/impurity_c_10/metrics.csv")
    assert _rows("outputs/impurity_c_10/metrics_10.csv") == _rows("outputs/impurity_c_10/metrics.csv")   # plot_cet.py:26


def test_melt_pool_and_gr_sweep(cet, tmp_path, monkeypatch):
    from cetkmc import campaign
    monkeypatch.chdir(tmp_path)
    out = campaign.run_cet_sublattice(L=16, n_sweeps=61, n_seeds=6, metrics_every=30, output_prefix="pool", verbose=False,
                                      laser=dict(power=200.0, speed=0.5, start=(8.0, 2.0), dt=1e-7))
    assert out[0].shape == (16, 16, 16) and len(_rows("outputs/pool/metrics.csv")) == 3
    rows = campaign.run_gr_sweep([2800, 3200], [2e13, 2e14], L=12, n_sweeps=41, n_seeds=5, metrics_every=20)
    assert len(rows) == 4 and len({(r["G"], r["R"]) for r in rows}) == 4
    merged = campaign.merge_cet_map()
    assert [int(r["case"]) for r in merged] == [0, 1, 2, 3]
    assert set(merged[0]) >= {"G", "R", "G_over_R", "AspectRatio", "EquiaxedFraction", "CET_Class"}
    # a case is one rank's work: rank 1 of 2 runs cases 1 and 3 only
    sub = campaign.run_gr_sweep([2800, 3200], [2e13, 2e14], L=12, n_sweeps=21, n_seeds=5, metrics_every=20,
                                output_root="gr2", rank=1, world=2)
    assert [r["case"] for r in sub] == [1, 3]
