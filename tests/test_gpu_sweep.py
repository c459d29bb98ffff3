"""GPU: synchronous-sublattice sweeps (csrc/sweep.cu, csrc/sweep_tile.cu) — invariants, determinism
and level-3 parity (trajectory observables against the serial oracle within statistical bounds).
Rate-maintenance variants (Context.debug_flags): COMPACT (default) = stamped sites refreshed by list-driven
gathers from the compact tile state (class-sorted kernel of rates_refresh.cu; PAIR_COMPACT = the pair-compacting
kernel of rates.cu), dense rebuilds by the class-sorted TMA tile kernel (rates_dense.cu) when
L % 16 == 0 (DENSE_UNSORTED: its pass B walks the sites by class only, not by class and pair count); TILE = the refresh's tile kernel for the
refresh and the rebuild (TMA staging), VECTOR / SCALAR its other staging modes, SERIAL its per-lane pair
loop; DENSE_COMPACT = dense rebuilds by the compact gather kernel; GATHER = refresh + rebuild of the first design
(neighbour-class cache + unit vectors)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _sweep_params(cet, seed, L, eps=0.02, p_max=0.25, defect_fraction=0.0, thermal_every=0):
    sp = cet._lib.SweepParams()
    sp.seed, sp.events_per_sweep, sp.p_max = seed, eps * L ** 3, p_max
    sp.defect_fraction, sp.thermal_every = defect_fraction, thermal_every
    return sp


COMPACT, GATHER, TILE, DENSE_COMPACT, DENSE_UNSORTED, PAIR_COMPACT = 0, 2, 32, 65536, 131072, 262144
SCALAR, SERIAL, VECTOR = TILE | 1, TILE | 4, TILE | 16
TMA = TILE
FUSED = COMPACT


def _setup(cet, L, seed=3, c=0.1, flags=FUSED):
    from cetkmc import _synth
    from cetkmc._config import rate_params
    packed, th, ph, T = _synth.half_grown(L, seed=seed, grain=4)
    ctx = cet.Context(L=L)
    ctx.debug_flags(flags)
    ctx.set_rate_params(rate_params(c))
    st, df = _synth.unpack(packed)
    ctx.upload(state=st, theta=th, phi=ph, T=T, defects=df)
    return ctx, st, th, ph, T, df


@pytest.mark.parametrize("L,flags", [(40, COMPACT), (64, TMA), (40, GATHER)])
def test_sweep_is_deterministic_and_consistent(cet, L, flags):
    outs = []
    for _ in range(2):
        ctx, st, th, ph, T, df = _setup(cet, L, flags=flags)
        res = ctx.sweep_run(12, _sweep_params(cet, 11, L, defect_fraction=0.01), None)
        outs.append((res, ctx.download(state=True, theta=True, phi=True), ctx.counts()))
        if flags == GATHER:
            assert ctx.nst_mismatches() == 0      # the incrementally patched neighbour cache equals a fresh gather
        ctx.close()
    (r1, f1, c1), (r2, f2, c2) = outs
    assert r1 == r2 and r1["events_applied"] > 0 and r1["events_applied"] <= r1["events_fired"]
    assert not r1["overflow"] and not r1["terminated"]
    for k in f1:
        np.testing.assert_array_equal(f1[k], f2[k])
    state, theta, phi = f1["state"], f1["theta"], f1["phi"]
    assert state.min() >= 0 and state.max() <= 4
    # empty and defect sites carry no orientation (kmc_simulation.py:289-301,323-327)
    assert np.all(theta[state == 0] == 0.0) and np.all(phi[state == 0] == 0.0)
    assert np.all(theta[state == 4] == 0.0)
    assert np.all((theta >= 0) & (theta <= np.pi)) and np.all((phi >= 0) & (phi <= 2 * np.pi))
    assert c1.sum() == L ** 3 and np.array_equal(c1, c2)


@pytest.mark.parametrize("flags", [FUSED, GATHER])
def test_first_sweep_is_primed(cet, oracle, flags):
    """A fresh clock (tau = 0) is primed by a pass that only measures the total rate, so the first
    counted sweep is a real one: it fires about events_per_sweep events, advances time by
    tau = events_per_sweep / R_total, and the total it reports equals the oracle's total rate of the
    initial lattice.  The expected number of fired sites is sum(1 - exp(-R tau))."""
    L = 24
    ctx, st, th, ph, T, df = _setup(cet, L, flags=flags)
    sp = _sweep_params(cet, 1, L, eps=0.01, p_max=0.1)
    res = ctx.sweep_run(1, sp, None)
    ev = oracle.event_rates(st, th, ph, T, df, L, oracle.make_params(0.1))
    total = oracle.pysum(ev["rate"])
    o_sr0, o_dep0, _, _ = oracle.site_rates(st, th, ph, T, df, L, oracle.make_params(0.1))
    R = o_sr0.copy(); R[L - 1] += np.nan_to_num(o_dep0)
    tau = min(sp.events_per_sweep / total, -np.log1p(-sp.p_max) / R.max())
    assert abs(res["time"] - tau) <= 1e-11 * tau and res["sweep_index"] == 1
    expect = float(np.sum(-np.expm1(-R * tau)))
    assert abs(res["events_fired"] - expect) <= 5 * np.sqrt(expect) + 1
    assert 0 < res["events_applied"] <= res["events_fired"]
    assert abs(res["last_total_rate"] - total) <= 1e-11 * total
    assert abs(res["last_max_rate"] - R.max()) <= 1e-12 * R.max()
    assert res["last_tau"] > 0
    ctx.close()


@pytest.mark.parametrize("L", [64, 80, 50])
def test_refresh_variants_agree(cet, L):
    """Every refresh variant — TMA tile kernel, cooperative tile loads, per-lane pair loop, and the
    gather kernels of the first design — runs the same trajectory bit for bit (the stream / pick /
    apply kernels are shared and every variant's rates equal the per-event code): lattice, counters
    and resident rates."""
    from cetkmc._config import thermal_params
    outs = []
    for flags in (COMPACT, COMPACT | 1048576, COMPACT | 2097152, PAIR_COMPACT, DENSE_UNSORTED, DENSE_COMPACT, TILE, SCALAR, VECTOR, SERIAL, SERIAL | 16, GATHER, COMPACT | 8):
        ctx, st, th, ph, T, df = _setup(cet, L, flags=flags)
        res = ctx.sweep_run(7, _sweep_params(cet, 5, L, eps=0.01, p_max=0.2, defect_fraction=0.01, thermal_every=3),
                            thermal_params(1e-6, nan_to_num=True))
        outs.append((res, ctx.download(state=True, theta=True, phi=True, T=True), ctx.rates_download()))
        ctx.close()
    r1, f1, (s1, d1) = outs[0]
    assert r1["events_applied"] > 100
    for n, (r2, f2, (s2, d2)) in enumerate(outs[1:], 1):
        assert r1 == r2, f"variant #{n}"
        for k in f1:
            np.testing.assert_array_equal(f1[k], f2[k], err_msg=f"variant #{n} {k}")
        np.testing.assert_array_equal(s1, s2, err_msg=f"variant #{n}")
        np.testing.assert_array_equal(d1, d2, err_msg=f"variant #{n}")


def test_oriented_empty_sites_take_the_gather_path(cet):
    """Empty sites that carry an orientation (never produced by the reference, but legal input)
    break the invariant the fused kernel's pair operands rely on: the sweep detects it on the device
    and runs the general gather kernels, whose rates stay equal to a dense rebuild."""
    L = 32
    ctx, st, th, ph, T, df = _setup(cet, L)
    rng = np.random.default_rng(0)
    th2 = np.where(st == 0, rng.uniform(0, np.pi, st.shape), th)
    ctx.upload(theta=th2)
    res = ctx.sweep_run(4, _sweep_params(cet, 2, L, eps=0.01, p_max=0.2), None)
    assert res["events_applied"] > 50
    assert ctx.nst_mismatches() == 0          # only the gather path maintains (and validates) the cache
    sr1, dr1 = ctx.rates_download()
    ctx.rates_build()
    sr2, dr2 = ctx.rates_download()
    ctx.close()
    np.testing.assert_array_equal(sr1, sr2)
    np.testing.assert_array_equal(dr1, dr2)


def test_diffusion_only_conserves_atoms(cet):
    """Attachment, nucleation and deposition switched off (enormous bond energies make attachment
    rates underflow, I0 = 0, a 1 K top plane): only diffusion of bond-free atoms can fire, which
    must conserve the species counts and the multiset of orientations."""
    from cetkmc import _synth
    from cetkmc._config import rate_params
    L = 32
    packed, th, ph, T = _synth.half_grown(L, seed=8, grain=4, fill=0.08)
    st, df = _synth.unpack(packed)
    T = T.copy()
    T[:] = 3000.0
    T[L - 1] = 1.0                                   # deposition rate underflows to exactly 0
    ctx = cet.Context(L=L)
    rp = rate_params(0.0, overrides=dict(I0=0.0, E_B_W=1e6, E_B_RE=1e6, E_B_C=1e6))
    ctx.set_rate_params(rp)
    ctx.upload(state=st, theta=th, phi=ph, T=T, defects=df)
    before = ctx.counts()
    res = ctx.sweep_run(10, _sweep_params(cet, 5, L, eps=0.01), None)
    after = ctx.counts()
    f = ctx.download(state=True, theta=True)
    ctx.close()
    np.testing.assert_array_equal(before, after)
    assert res["events_applied"] > 0
    assert not np.array_equal(f["state"], st)
    np.testing.assert_array_equal(np.sort(f["theta"][f["state"] > 0]), np.sort(th[st > 0]))


LEVEL3_K = 4.0          # tolerance of the level-3 comparison, in standard errors of the difference of the two ensemble means
L3 = dict(L=30, n_seeds=16, c=0.1, defect_fraction=3e-3, checkpoints=(500, 1000, 1500, 2000))
L3_NAMES = ("EquiaxedFraction", "GrainCount", "AspectRatio", "occupied", "Re fraction")


def _l3_observe(m, state):
    occ = int((state != 0).sum())
    return [m["EquiaxedFraction"], m["GrainCount"], m["AspectRatio"], occ, (state == 2).sum() / max(occ, 1)]


@pytest.fixture(scope="module")
def l3_oracle(oracle):
    """The serial side of level 3: oracle.kmc_run (kmc_simulation.py:246-332, thermal every 20 steps,
    defect injection) on 16 seeded main.py-default lattices; observables by the host restatement."""
    import hostref
    L, c = L3["L"], L3["c"]
    n_events = L3["checkpoints"][-1]
    obs, cet_pos, lattices = [], [], []
    for seed in range(L3["n_seeds"]):
        st, th, ph, T, at = oracle.initialize_lattice(L, n_seeds=20, random_seed=seed, impurity_c=c)
        lattices.append((st, th, ph, T))
        df = np.zeros_like(st)
        d = oracle.DrawStreams(seed=seed, n_py=3 * n_events, n_np=2 * n_events, n_sp=n_events * L * L)
        o = [a.copy() for a in (st, at, th, ph, T)]
        pos, first, done = [0, 0, 0], -1, 0
        for cp in L3["checkpoints"]:
            r = oracle.kmc_run(o[0], o[1], o[2], o[3], o[4], df, L, oracle.make_params(c), done, cp - done,
                               L3["defect_fraction"], d.py[pos[0]:], d.np[pos[1]:], d.sp[pos[2]:],
                               thermal=oracle.make_thermal_params(), log=False)
            assert r["steps_done"] == cp - done
            pos = [pos[0] + r["py_used"], pos[1] + r["np_used"], pos[2] + r["sp_used"]]
            done = cp
            m = hostref.compute_metrics(o[0], o[2], o[3])
            if first < 0 and hostref.detect_CET_transition(m):
                first = cp
        obs.append(_l3_observe(m, o[0])); cet_pos.append(first)
    return lattices, np.array(obs, dtype=float), np.array(cet_pos, dtype=float)


def _l3_sublattice(cet, lattices, events_per_sweep):
    """The sublattice side: same lattices, one thermal update per 20 EXECUTED events, observables from
    the GPU clustering (Context.grains)."""
    from cetkmc import metrics as M
    from cetkmc._config import rate_params, thermal_params
    L, c = L3["L"], L3["c"]
    n_events = L3["checkpoints"][-1]
    obs, cet_pos, over = [], [], []
    for seed, (st, th, ph, T) in enumerate(lattices):
        ctx = cet.Context(L=L)
        ctx.set_rate_params(rate_params(c))
        ctx.upload(state=st, theta=th, phi=ph, T=T, defects=np.zeros_like(st))
        sp = _sweep_params(cet, 1000 + seed, L, eps=events_per_sweep / L ** 3, p_max=0.1,
                           defect_fraction=L3["defect_fraction"], thermal_every=0)
        tp = thermal_params(1e-6, nan_to_num=True)
        applied, first, n_thermal = 0, -1, 0
        for cp in L3["checkpoints"]:
            while applied < cp:
                while n_thermal <= applied // 20:            # steps 0, 20, 40, ... of the reference (kmc_simulation.py:248)
                    ctx.thermal_cet(tp)
                    n_thermal += 1
                applied += ctx.sweep_run(1, sp, None)["events_applied"]
            m = M.metrics_from_grains(ctx.grains(0.5), L ** 3)
            if first < 0 and M.detect_CET_transition(m):
                first = cp
        f = ctx.download(state=True)
        ctx.close()
        obs.append(_l3_observe(m, f["state"])); cet_pos.append(first); over.append(applied - n_events)
    return np.array(obs, dtype=float), np.array(cet_pos, dtype=float), np.array(over, dtype=float)


def test_level3_observables_vs_serial_oracle(cet, l3_oracle):
    """Level-3 parity at the main.py default (BASELINE configs[0]: L = 30, n_seeds = 20,
    defect_fraction = 3e-3, thermal update every 20 events, kmc_simulation.py:248): over N = 16 seeds
    the sublattice path and the serial oracle, started from the same initial lattices and compared at
    the same number of executed events, agree on the north-star observables
        equiaxed fraction, grain count, mean grain aspect ratio, CET position
    (CET position = the first checkpoint at which detect_CET_transition holds, metrics.py:103-105,
    checked every 500 events; -1 = never) and on the occupied count and Re fraction, within
    LEVEL3_K standard errors of the difference of the means (no additive slack).

    Matching conventions, stated because the two algorithms are not step-for-step comparable:
      * events: 4 events per sweep (0.015 % of the sites).  The synchronous sweep is a first-order
        approximation in the sweep interval; its bias against the serial algorithm at larger sweeps is
        measured, not assumed away (test_level3_bias_at_the_benchmark_setting);
      * thermal: the stencil is driven from the host once per 20 EXECUTED events like the reference's
        `step % 20 == 0` — not once per sweep: when the unstable stencil has produced sites above T_MELT
         the p_max cap shortens the sweeps to a few events each, and a per-sweep cadence
        would advance the temperature field several times faster per event than the reference does;
      * time: the reference's clock is dt = max(-ln u / R, 1e-12) with R ~ 1e17..1e18 s^-1, so the floor
        always wins and Time == 1e-12 s x events — equal event counts ARE equal reference
        times; the sublattice path's own clock (sum of tau) is physical and not comparable."""
    lattices, obs_o, cet_o = l3_oracle
    obs_g, cet_g, over = _l3_sublattice(cet, lattices, 4.0)
    n = L3["n_seeds"]
    assert over.max() < 20                                   # the last sweep overshoots by a few events at most
    mo, mg = obs_o.mean(0), obs_g.mean(0)
    se = np.sqrt(obs_o.var(0, ddof=1) / n + obs_g.var(0, ddof=1) / n)
    report = {k: (round(a, 4), round(b, 4), round(e, 4)) for k, a, b, e in zip(L3_NAMES, mo, mg, se)}
    assert np.all(np.abs(mo - mg) <= LEVEL3_K * se + 1e-12), report
    se_c = np.sqrt(cet_o.var(ddof=1) / n + cet_g.var(ddof=1) / n)
    assert abs(cet_o.mean() - cet_g.mean()) <= LEVEL3_K * se_c + 1e-12, (cet_o, cet_g)


def test_level3_bias_at_the_benchmark_setting(cet, l3_oracle):
    """The throughput benchmark fires 0.5 % of the sites per sweep (bench.py).  At that setting the
    synchronous sweeps are a coarser approximation of the serial algorithm: this test MEASURES the
    bias of the same observables on the same ensemble and bounds it — relative deviation of the
    ensemble means <= 8 % for the counts, <= 0.03 absolute for the equiaxed fraction and the aspect
    ratio, CET position unchanged (profiles/ records the measured values)."""
    lattices, obs_o, cet_o = l3_oracle
    obs_g, cet_g, over = _l3_sublattice(cet, lattices, 0.005 * L3["L"] ** 3)
    mo, mg = obs_o.mean(0), obs_g.mean(0)
    scale = L3["checkpoints"][-1] / (L3["checkpoints"][-1] + over.mean())      # the last sweep overshoots by ~half a sweep
    report = {k: (round(a, 4), round(b, 4)) for k, a, b in zip(L3_NAMES, mo, mg)}
    print("level-3 bias at 0.5 % per sweep:", report, "overshoot", over.mean())
    assert abs(mg[0] - mo[0]) <= 0.03 and abs(mg[2] - mo[2]) <= 0.03, report
    assert abs(mg[1] * scale / mo[1] - 1) <= 0.08 and abs(mg[3] * scale / mo[3] - 1) <= 0.08, report
    assert cet_o.mean() == cet_g.mean(), (cet_o, cet_g)


@pytest.mark.parametrize("L,flags", [(36, COMPACT), (64, COMPACT), (64, PAIR_COMPACT), (96, TMA), (96, SERIAL), (36, GATHER)])
def test_resident_rates_equal_rebuild_after_sweeps(cet, L, flags):
    """Neighbour-rate refresh invariant: after N sweeps (thermal steps and defect injection
    included) the resident rate sums equal a dense rebuild by the gather kernel of rates.cu bit for
    bit — for the fused tile kernel this also pins its class-code / pair-operand arithmetic and its
    incrementally maintained tile state to the per-event code."""
    from cetkmc._config import thermal_params
    ctx, st, th, ph, T, df = _setup(cet, L, flags=flags)
    res = ctx.sweep_run(9, _sweep_params(cet, 21, L, eps=0.01, p_max=0.2, defect_fraction=0.02, thermal_every=4),
                        thermal_params(1e-6, nan_to_num=True))
    assert res["events_applied"] > 100 and res["sites_refreshed"] > res["events_applied"]
    sr1, dr1 = ctx.rates_download()          # resident arrays as the sweeps left them
    if flags == GATHER:
        assert ctx.nst_mismatches() == 0
    ctx.rates_build()                        # dense rebuild from the lattice
    sr2, dr2 = ctx.rates_download()
    ctx.close()
    np.testing.assert_array_equal(sr1, sr2)
    np.testing.assert_array_equal(dr1, dr2)


@pytest.mark.parametrize("L,flags", [(64, COMPACT), (80, COMPACT), (80, DENSE_UNSORTED), (96, DENSE_COMPACT)])
def test_dense_rebuild_edge_cases(cet, oracle, L, flags):
    """The dense rebuild kernels on inputs that leave fast_exp's range and exercise every clamp: cold blocks
    (T below 1 K and a few kelvin: Arrhenius arguments far beyond -700, the out-of-line path), sites at and
    above T_melt (no nucleation; the max(T_melt - T, 1) clamps), a steep temperature step along k (grad_z),
    partial tiles at the lattice faces (L = 80), the top plane's deposition rates.  The sweep's resident rates
    after a rebuild with (practically) no event must equal the per-event gather kernel bit for bit and the
    oracle within 1e-12."""
    from cetkmc import _synth
    from cetkmc._config import rate_params
    packed, th, ph, T = _synth.half_grown(L, seed=9, grain=4)
    T = T.copy()
    T[3:9, 5:20, 10:30] = 0.25
    T[20:26, :, 4:12] = 7.0
    T[30:40, 10:30, :] = 3695.0
    T[41:44, :, L // 2 - 3:] = 3900.0
    T[50:60, 20:40, L // 2:] += 600.0
    T[L - 1, 0:8, 0:8] = 0.5
    ctx = cet.Context(L=L)
    ctx.debug_flags(flags)
    ctx.set_rate_params(rate_params(0.1))
    st, df = _synth.unpack(packed)
    ctx.upload(state=st, theta=th, phi=ph, T=T, defects=df)
    res = ctx.sweep_run(1, _sweep_params(cet, 3, L, eps=1e-12, p_max=1e-9), None)
    assert res["events_applied"] == 0
    sr1, dr1 = ctx.rates_download()
    ctx.rates_build()
    sr2, dr2 = ctx.rates_download()
    ctx.close()
    np.testing.assert_array_equal(sr1, sr2)
    np.testing.assert_array_equal(dr1, dr2)
    o_sr, o_dep, _, _ = oracle.site_rates(st, th, ph, T, df, L, oracle.make_params(0.1))
    np.testing.assert_allclose(sr1, o_sr, rtol=1e-12, atol=0.0)
    np.testing.assert_allclose(dr1, o_dep, rtol=1e-12, atol=0.0, equal_nan=True)
