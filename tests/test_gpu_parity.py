"""GPU: the CUDA path (through the C-ABI) against the oracle and the reference fixtures.

Parity levels (BASELINE.json north_star):
  1. rate arrays / event lists: event set, order, targets, species exact; rates rtol 1e-12
     (fp64; the only differences are ulps of exp/sin/cos/acos between CUDA libdevice and libm);
  2. selection + state updates bit-exact under injected draws;
  3. trajectory observables (tests/test_gpu_sweep.py).
The thermal stencil is bit-exact (no transcendental functions, fixed operation order)."""
import ast
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden

pytestmark = pytest.mark.gpu
RTOL = 1e-12
RATE_CASES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "rates_*.npz")))
TRAJ_CASES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "traj_*.npz")))


# ---------------------------------------------------------------- thermal (bit-exact)
def test_thermal_cet_golden_sequences(cet):
    from cetkmc import thermal_solver as ts
    for name, steps in (("thermal_cet_random.npz", (1, 2, 5, 40)), ("thermal_cet_gradient12.npz", (1, 3, 10, 60))):
        g = golden(name)
        T = g["T0"]
        T_in = T.copy()
        for n in range(1, max(steps) + 1):
            T = ts.update_temperature_cet(T, None, dt=1e-6)
            if n in steps:
                np.testing.assert_array_equal(T, g[f"T_after_{n}"])
        np.testing.assert_array_equal(T_in, g["T0"])          # caller's array untouched
        assert T.dtype == np.float64 and T.flags.c_contiguous


@pytest.mark.parametrize("shape", [(64, 64, 64), (5, 130, 33), (1, 1, 7), (2, 3, 1), (40, 17, 257), (6, 10, 2),
                                   (3, 5, 66), (9, 11, 130), (33, 20, 64), (4, 4, 4)])
def test_thermal_cet_resident_vs_oracle(cet, oracle, shape):
    from cetkmc._config import thermal_params
    rng = np.random.default_rng(3)
    T0 = 2800 + 900 * rng.random(shape)
    ctx = cet.Context(shape=shape)
    ctx.upload(T=T0)
    T = T0
    for n in range(50):
        ctx.thermal_cet(thermal_params(1e-6))
        T = oracle.thermal_cet(T)
        if n in (0, 1, 7, 49):
            np.testing.assert_array_equal(ctx.download(T=True)["T"], T)
    ctx.close()


def test_thermal_nan_to_num(cet, oracle):
    from cetkmc._config import thermal_params
    rng = np.random.default_rng(4)
    T0 = 2800 + 900 * rng.random((12, 12, 12))
    T0[3, 4, 5] = np.nan; T0[0, 0, 0] = np.inf; T0[11, 11, 11] = -np.inf
    ctx = cet.Context(L=12)
    ctx.upload(T=T0)
    ctx.thermal_cet(thermal_params(1e-6, nan_to_num=True))
    np.testing.assert_array_equal(ctx.download(T=True)["T"], oracle.thermal_cet(T0, nan_to_num=True))
    ctx.close()


def test_thermal_full_golden(cet):
    from cetkmc import thermal_solver as ts
    g = golden("thermal_full10.npz")
    st, pv = g["state"].astype(np.int64), g["prev_state"].astype(np.int64)
    np.testing.assert_array_equal(ts.update_temperature(g["T0"], st, pv, 1e-7, (3, 4.5), 200.0), g["T1"])
    np.testing.assert_array_equal(
        ts.update_temperature(g["T0"], st, pv, 1e-9, (0, 2), 50.0, beam_radius=20e-6, absorptivity=0.5), g["T2"])


def test_build_temperature_field(cet):
    from cetkmc import thermal_solver as ts
    T = ts.build_temperature_field(9)
    want = np.repeat(np.repeat((2800 + ((3695 - 2800) / 8) * np.arange(9.0))[:, None, None], 9, 1), 9, 2)
    np.testing.assert_array_equal(T, want)


# ---------------------------------------------------------------- level 1: rates
@pytest.mark.parametrize("name", RATE_CASES)
def test_get_event_rates_golden(cet, name):
    from cetkmc import kmc_event_rates as ker
    g = golden(name)
    L = g["state"].shape[0]
    st = g["state"].astype(np.int64)
    ker.seed_species(int(g["species_seed"]))
    ev = ker.get_event_rates(st, g["theta"], g["phi"], g["T"], st.copy(), g["defects"].astype(np.int64), L, 1, 2, 3,
                             step=0, debug_step=1000, impurity_c=float(g["impurity_c"]))
    assert isinstance(ev, list) and len(ev) == g["ev_type"].size
    e0 = ev[0]
    assert isinstance(e0[0], bytes) and isinstance(e0[1], tuple) and isinstance(e0[2], float) \
        and isinstance(e0[3], tuple) and isinstance(e0[4], int)
    names = (b"dep", b"diff", b"nuc", b"att")
    LL = L * L
    assert [names.index(e[0]) for e in ev] == g["ev_type"].tolist()
    assert [(e[1][0] * L + e[1][1]) * L + e[1][2] for e in ev] == g["ev_pos"].tolist()
    assert [(-1 if e[3][0] < 0 else (e[3][0] * L + e[3][1]) * L + e[3][2]) for e in ev] == g["ev_target"].tolist()
    assert [e[4] for e in ev] == g["ev_atom"].tolist()
    np.testing.assert_allclose([e[2] for e in ev], g["ev_rate"], rtol=RTOL, atol=0.0)
    assert LL > 0


def test_compute_row_events_plane(cet):
    from cetkmc import kmc_event_rates as ker
    g = golden("rates_half10_t0.npz")
    L = 10
    st = g["state"].astype(np.int64)
    for i in (0, 4, L - 1):
        ker.seed_species(int(g["species_seed"]))
        ev = ker.compute_row_events(i, st, g["theta"], g["phi"], g["T"], st, g["defects"].astype(np.int64),
                                    1e13, 2e13, np.array([3.8, 4.2, 3.2]), np.array([0.35, 0.50, 0.30]),
                                    8.617333262e-5, 3695, 5e13, 10, L - 1, L, 1, 2, 3, float(g["impurity_c"]))
        sel = (g["ev_pos"] // (L * L)) == i
        assert len(ev) == int(sel.sum())
        np.testing.assert_allclose([e[2] for e in ev], g["ev_rate"][sel], rtol=RTOL)
        assert [e[4] for e in ev] == g["ev_atom"][sel].tolist()


@pytest.mark.parametrize("L,ups", [(48, 0), (33, 3), (64, 40)])
def test_site_rates_and_total_vs_oracle(cet, oracle, L, ups):
    from cetkmc._config import rate_params
    st, th, ph, T, df = oracle.half_grown_lattice(L, seed=L, T_updates=ups)
    ctx = cet.Context(L=L)
    ctx.set_rate_params(rate_params(0.1))
    ctx.upload(state=st, theta=th, phi=ph, T=T, defects=df)
    ctx.rates_build()
    sr, dr = ctx.rates_download()
    total, n_dep = ctx.rates_total()
    o_sr, o_dep, n_ev, _ = oracle.site_rates(st, th, ph, T, df, L, oracle.make_params(0.1))
    np.testing.assert_allclose(sr, o_sr, rtol=RTOL, atol=0.0)
    assert np.array_equal(np.isnan(dr), np.isnan(o_dep))
    np.testing.assert_allclose(np.nan_to_num(dr), np.nan_to_num(o_dep), rtol=RTOL, atol=0.0)
    assert n_dep == int((~np.isnan(o_dep)).sum())
    ev = oracle.event_rates(st, th, ph, T, df, L, oracle.make_params(0.1))
    assert abs(total - oracle.pysum(ev["rate"])) <= 1e-11 * total
    n, nd = ctx.events_count()
    assert n == n_ev and nd == n_dep
    ctx.close()


# ---------------------------------------------------------------- level 2: exact BKL, injected draws
@pytest.mark.parametrize("L", [36, 66])
def test_site_rates_dense_interfaces(cet, oracle, L):
    """Rows of solid separated by empty rows: every occupied site owns 12 diffusion events and
    every empty one up to 8 attachment events, more pairs per 32 sites than the warp's pair slots
    hold (the two-half-tile path of rate_tile.cuh); general orientations on the empty sites too."""
    from cetkmc._config import rate_params
    rng = np.random.default_rng(L)
    i, j = np.arange(L)[:, None, None], np.arange(L)[None, :, None]
    st = np.where((i % 3 == 0) & (j % 3 == 0), rng.choice(np.array([1, 2, 3, 4]), size=(L, L, L), p=[.7, .15, .1, .05]), 0)
    st = np.ascontiguousarray(np.broadcast_to(st, (L, L, L))).astype(np.int64)
    th = rng.uniform(0, np.pi, (L, L, L)); ph = rng.uniform(0, 2 * np.pi, (L, L, L))
    T = 2800 + 895.0 * rng.random((L, L, L))
    df = ((st == 3) & (rng.random((L, L, L)) < 0.3)).astype(np.int64)
    ctx = cet.Context(L=L)
    ctx.set_rate_params(rate_params(0.1))
    ctx.upload(state=st, theta=th, phi=ph, T=T, defects=df)
    ctx.rates_build()
    sr, dr = ctx.rates_download()
    o_sr, o_dep, n_ev, _ = oracle.site_rates(st, th, ph, T, df, L, oracle.make_params(0.1))
    np.testing.assert_allclose(sr, o_sr, rtol=RTOL, atol=0.0)
    assert np.array_equal(np.isnan(dr), np.isnan(o_dep))
    # atol: libdevice's exp flushes results in the last subnormal binades (~1e-323) to zero, libm does not
    np.testing.assert_allclose(np.nan_to_num(dr), np.nan_to_num(o_dep), rtol=RTOL, atol=1e-300)
    # the per-event list (site_events code path) agrees with the dense sums
    ev = oracle.event_rates(st, th, ph, T, df, L, oracle.make_params(0.1))
    assert ctx.events_count()[0] == ev["rate"].size


# the last case is BASELINE.json configs[1]: the 128^3 lattice with injected draws (few steps: the
# oracle rebuilds all 2.1e6 rates for every executed event, ~0.5 s per step)
@pytest.mark.parametrize("L,steps,defect_fraction,c,ups", [(12, 400, 0.02, 0.1, 0), (16, 300, 0.0, 0.0, 2),
                                                              (20, 250, 3e-3, 0.2, 30), (128, 16, 3e-3, 0.1, 0)])
def test_kmc_run_bit_exact_vs_oracle(cet, oracle, L, steps, defect_fraction, c, ups):
    from cetkmc._config import rate_params, thermal_params
    st, th, ph, T, df = oracle.half_grown_lattice(L, seed=100 + L, T_updates=ups, grain=4)
    draws = oracle.DrawStreams(seed=L, n_py=3 * steps, n_np=2 * steps, n_sp=steps * L * L)
    o = [a.copy() for a in (st, st, th, ph, T)]
    ores = oracle.kmc_run(o[0], o[1], o[2], o[3], o[4], df, L, oracle.make_params(c), 0, steps, defect_fraction,
                          draws.py, draws.np, draws.sp)
    ctx = cet.Context(L=L)
    ctx.set_rate_params(rate_params(c))
    ctx.upload(state=st, theta=th, phi=ph, T=T, defects=df)
    res = ctx.kmc_run(0, steps, defect_fraction, thermal_params(1e-6, nan_to_num=True), 20, draws.py, draws.np,
                      draws.sp, log=True)
    got = ctx.download(state=True, atom_type=True, theta=True, phi=True, T=True)
    ctx.close()
    assert res["steps_done"] == ores["steps_done"] == steps and not res["terminated"] and not res["starved"]
    for k in ("type", "pos", "target", "atom"):
        np.testing.assert_array_equal(res["log_" + k], ores["log_" + k], err_msg=k)
    np.testing.assert_allclose(res["log_rate"], ores["log_rate"], rtol=RTOL)
    np.testing.assert_allclose(res["log_total"], ores["log_total"], rtol=1e-11)
    assert (res["py_used"], res["np_used"], res["sp_used"]) == (ores["py_used"], ores["np_used"], ores["sp_used"])
    assert res["nucleation_count"] == ores["nucleation_count"]
    assert res["total_time"] == ores["total_time"]
    np.testing.assert_array_equal(got["state"], o[0])
    np.testing.assert_array_equal(got["atom_type"], o[1])
    np.testing.assert_array_equal(got["theta"], o[2])
    np.testing.assert_array_equal(got["phi"], o[3])
    np.testing.assert_array_equal(got["T"], o[4])


def test_kmc_run_incremental_hierarchy_equals_rebuild(cet, oracle):
    """After N incremental steps the resident rate sums equal a from-scratch rebuild bit for bit."""
    from cetkmc._config import rate_params
    L, steps = 14, 200
    st, th, ph, T, df = oracle.half_grown_lattice(L, seed=9, grain=4)
    draws = oracle.DrawStreams(seed=1, n_py=2 * steps, n_np=2 * steps, n_sp=steps * L * L)
    ctx = cet.Context(L=L)
    ctx.set_rate_params(rate_params(0.0))
    ctx.upload(state=st, theta=th, phi=ph, T=T, defects=df)
    res = ctx.kmc_run(0, steps, 0.0, None, 0, draws.py, draws.np, draws.sp)
    assert res["steps_done"] == steps
    sr1, dr1 = ctx.rates_download()
    t1 = ctx.rates_total()
    ctx.rates_build()
    sr2, dr2 = ctx.rates_download()
    t2 = ctx.rates_total()
    ctx.close()
    np.testing.assert_array_equal(sr1, sr2)
    np.testing.assert_array_equal(dr1, dr2)
    assert t1 == t2


def test_kmc_run_starved_and_terminated(cet, oracle):
    from cetkmc._config import rate_params
    L = 8
    st, th, ph, T, df = oracle.half_grown_lattice(L, seed=2, grain=4)
    ctx = cet.Context(L=L)
    ctx.set_rate_params(rate_params(0.0))
    ctx.upload(state=st, theta=th, phi=ph, T=T, defects=df)
    d = oracle.DrawStreams(seed=3, n_py=2 * 5, n_np=2 * 50, n_sp=50 * L * L)
    res = ctx.kmc_run(0, 50, 0.0, None, 0, d.py, d.np, d.sp)
    assert res["starved"] == 1 and res["steps_done"] == 5 and res["py_used"] == 10
    # a lattice with no possible event terminates at once (kmc_simulation.py:260-262)
    full = np.full((L, L, L), 4, dtype=np.int64)
    ctx.upload(state=full, theta=np.zeros((L, L, L)), phi=np.zeros((L, L, L)), T=T, defects=np.zeros_like(full))
    d = oracle.DrawStreams(seed=3, n_py=20, n_np=20, n_sp=0)
    res = ctx.kmc_run(0, 10, 0.0, None, 0, d.py, d.np, None)
    assert res["terminated"] == 1 and res["steps_done"] == 0
    ctx.close()


# ---------------------------------------------------------------- run_kmc drop-in vs the reference's own runs
@pytest.mark.parametrize("name", TRAJ_CASES)
def test_run_kmc_reproduces_reference_trajectory(cet, name, tmp_path, monkeypatch):
    from cetkmc import kmc_simulation as ks
    g = golden(name)
    kw = ast.literal_eval(str(g["kwargs"]))
    monkeypatch.chdir(tmp_path)
    state, atom_type, total_time, theta, phi = ks.run_kmc(output_prefix="t", **kw)
    assert state.dtype == np.int64 and atom_type.dtype == np.int64 and theta.dtype == np.float64
    np.testing.assert_array_equal(state, g["state"])
    np.testing.assert_array_equal(atom_type, g["atom_type"])
    np.testing.assert_array_equal(theta, g["theta"])
    np.testing.assert_array_equal(phi, g["phi"])
    assert total_time == float(g["total_time"])
    import pandas as pd
    df = pd.read_csv(tmp_path / "outputs" / "t" / "metrics.csv")
    assert list(df.columns) == ["Step", "Time", "AspectRatio", "EquiaxedFraction", "NucleationDensity",
                                "DefectDensity", "AvgGrainSize", "GrainCount", "W_Count", "Re_Count", "C_Count",
                                "NucleationCount", "G_over_R", "G_phys", "R_phys", "G_over_R_phys", "CET_Class",
                                "CET_Detected"]
    for col in df.columns:
        want = g[f"csv_{col}"]
        if col == "CET_Class":
            assert df[col].tolist() == want.tolist()
        else:
            np.testing.assert_allclose(df[col].to_numpy().astype(float), want.astype(float), rtol=1e-12, err_msg=col)
