"""Generate tests/golden/*.npz from the UNMODIFIED reference (test infrastructure).

Runs only where /root/reference exists (the build container).  The fixtures pin both the CPU
oracle (tests -m "not gpu") and the CUDA path (tests -m gpu) to outputs of the reference's own
code on seeded inputs:
    rates_*.npz       get_event_rates (kmc_event_rates.py:162) event lists
    thermal_*.npz     update_temperature_cet / update_temperature (thermal_solver.py:36,107)
    traj_*.npz        run_kmc (kmc_simulation.py:203) final arrays + metrics.csv rows
    grains_*.npz      utils.get_clusters / metrics.compute_metrics (utils.py:69, metrics.py:41)
    defects_*.npz     defects.introduce_defects (defects.py:25)
Usage:  python oracle/gen_golden.py [--grains-only | --defects-only | --traj NAME]
"""
import io
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import oracle as O          # noqa: E402  (input generators only)
import refharness           # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
TYPE_CODE = {b"dep": 0, b"diff": 1, b"nuc": 2, b"att": 3}


def events_to_soa(events, L):
    n = len(events)
    ty = np.zeros(n, np.uint8); pos = np.zeros(n, np.int32); rate = np.zeros(n, np.float64)
    tgt = np.zeros(n, np.int32); atom = np.zeros(n, np.int8)
    for q, (t, p, r, g, a) in enumerate(events):
        ty[q] = TYPE_CODE[t]; pos[q] = (p[0] * L + p[1]) * L + p[2]; rate[q] = r
        tgt[q] = -1 if g[0] < 0 else (g[0] * L + g[1]) * L + g[2]; atom[q] = a
    return dict(ev_type=ty, ev_pos=pos, ev_rate=rate, ev_target=tgt, ev_atom=atom)


def rates_case(ref, name, state, theta, phi, T, defects, impurity_c, nb_seed_value):
    L = state.shape[0]
    ref["nb_seed"](nb_seed_value)
    ev = ref["kmc_event_rates"].get_event_rates(state, theta, phi, T, state.copy(), defects, L, 1, 2, 3,
                                                step=0, debug_step=1000, impurity_c=impurity_c)
    d = events_to_soa(ev, L)
    d.update(state=state.astype(np.int8), theta=theta, phi=phi, T=T, defects=defects.astype(np.int8),
             impurity_c=np.float64(impurity_c), species_seed=np.int64(nb_seed_value),
             total_pysum=np.float64(sum(e[2] for e in ev)))
    np.savez_compressed(os.path.join(OUT, f"rates_{name}.npz"), **d)
    print(f"rates_{name}: L={L} events={len(ev)}")


def grains_cases(ref):
    """utils.get_clusters / metrics.compute_metrics (utils.py:69, metrics.py:41) on seeded lattices."""
    ut, me = ref["utils"], ref["metrics"]
    for name, L, seed, grain, fill, jitter in (("grown12", 12, 3, 4, 0.7, 0.15), ("grown16", 16, 8, 5, 0.45, 0.3),
                                               ("half14", 14, 21, 4, None, None)):
        if fill is None:
            st, th, ph, _, df = O.half_grown_lattice(L, seed=seed, grain=grain)
        else:
            st, th, ph = O.grown_lattice(L, seed=seed, grain=grain, fill=fill, jitter=jitter)
            df = (st == 4).astype(np.int64)
        clusters, visited = ut.get_clusters(st, th, ph, theta_threshold=0.5)
        m = me.compute_metrics(st, th, ph, defects=df)
        ars = np.array([ut.calculate_aspect_ratio(c) for c in clusters])
        keys = ("AspectRatio", "EquiaxedFraction", "NucleationDensity", "AvgGrainSize", "GrainCount", "DefectDensity",
                "Grain_d50_um", "Grain_d90_um")
        np.savez_compressed(os.path.join(OUT, f"grains_{name}.npz"), state=st.astype(np.int8), theta=th, phi=ph,
                            defects=df.astype(np.int8), visited=np.asarray(visited, dtype=np.int32),
                            sizes=np.array([len(c) for c in clusters], dtype=np.int32), aspect=ars,
                            first=np.array([(c[0][0] * L + c[0][1]) * L + c[0][2] for c in clusters], dtype=np.int32),
                            cet=np.array(me.compute_CET(st, th, ph)),
                            **{f"m_{k}": np.float64(m[k]) for k in keys})
        print(f"grains_{name}: L={L} grains={len(clusters)} AR={m['AspectRatio']:.4f} eq={m['EquiaxedFraction']:.3f}")


def defects_cases(ref):
    """defects.introduce_defects (defects.py:25-31) with NumPy's global stream seeded."""
    de = ref["defects"]
    for name, L, seed in (("c20", 14, 4), ("c35", 20, 9)):
        rng = np.random.default_rng(seed)
        st = rng.choice(np.array([0, 1, 2, 3, 4]), size=(L, L, L), p=[.3, .25, .1, .3, .05]).astype(np.int64)
        T = 2500 + 1500 * rng.random((L, L, L))
        T[rng.random((L, L, L)) < 0.03] = -5.0                     # defects.py:13: falls back to T_SUB
        np.random.seed(seed)
        mask, density = de.introduce_defects(st.copy(), st, T, apply_to_state=False)
        nxt = np.random.random()                                     # stream position after the call
        np.random.seed(seed)
        mask_noT, _ = de.introduce_defects(st.copy(), st, None)
        np.savez_compressed(os.path.join(OUT, f"defects_{name}.npz"), state=st.astype(np.int8), T=T, seed=np.int64(seed),
                            mask=mask.astype(np.int8), density=np.float64(density), next_draw=np.float64(nxt),
                            mask_noT=mask_noT.astype(np.int8))
        print(f"defects_{name}: L={L} carbon={(st == 3).sum()} masked={mask.sum()} / {mask_noT.sum()} (no T)")


# run_kmc trajectories.  The L30 cases are the main.py default (main.py:33-64: L=30,
# defect_fraction=DEFECT_PROB=3e-3, n_seeds=20, one run per carbon level), shortened from 20 000 to
# 2 001 steps (~3 min of reference time each); generate them with --traj NAME (they run in parallel).
TRAJ_CASES = {
    "L10_c01_def": dict(L=10, n_steps=401, temp=2800, defect_fraction=3e-3, n_seeds=5, impurity_c=0.1),
    "L8_c00": dict(L=8, n_steps=250, temp=2800, defect_fraction=0.0, n_seeds=4, impurity_c=0.0),
    "L12_c02_def": dict(L=12, n_steps=601, temp=2800, defect_fraction=0.02, n_seeds=8, impurity_c=0.2),
    "L30_c00": dict(L=30, n_steps=2001, temp=2800, defect_fraction=3e-3, n_seeds=20, impurity_c=0.0),
    "L30_c01": dict(L=30, n_steps=2001, temp=2800, defect_fraction=3e-3, n_seeds=20, impurity_c=0.1),
    "L30_c02": dict(L=30, n_steps=2001, temp=2800, defect_fraction=3e-3, n_seeds=20, impurity_c=0.2),
}


def traj_case(ref, name):
    import pandas as pd
    km = ref["kmc_simulation"]
    kw = TRAJ_CASES[name]
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            ref["nb_seed"](42)
            stdout, sys.stdout = sys.stdout, io.StringIO()
            try:
                state, atom_type, total_time, theta, phi = km.run_kmc(output_prefix="g", **kw)
            finally:
                sys.stdout = stdout
            df = pd.read_csv(os.path.join("outputs", "g", "metrics.csv"))
        finally:
            os.chdir(cwd)
    cols = {f"csv_{c}": df[c].to_numpy() for c in df.columns if c not in ("CET_Class",)}
    cols["csv_CET_Class"] = np.array(df["CET_Class"].tolist())
    np.savez_compressed(os.path.join(OUT, f"traj_{name}.npz"), state=state.astype(np.int8),
                        atom_type=atom_type.astype(np.int8), theta=theta, phi=phi,
                        total_time=np.float64(total_time), kwargs=np.array(repr(kw)), **cols)
    print(f"traj_{name}: occupied={int((state != 0).sum())} rows={len(df)} time={total_time:.3e}")


def main():
    ref = refharness.load()
    os.makedirs(OUT, exist_ok=True)
    if "--grains-only" in sys.argv:
        grains_cases(ref)
        return
    if "--defects-only" in sys.argv:
        defects_cases(ref)
        return
    if "--traj" in sys.argv:
        traj_case(ref, sys.argv[sys.argv.index("--traj") + 1])
        return
    li, ts, km = ref["lattice_init"], ref["thermal_solver"], ref["kmc_simulation"]

    # ---- rates -----------------------------------------------------------------------------
    st, th, ph, T, at = li.initialize_lattice(lattice_size=8, n_seeds=6, T_sub=2800, random_seed=42, impurity_c=0.0)
    rates_case(ref, "fresh8", st.astype(np.int64), th, ph, T, np.zeros_like(st, dtype=np.int64), 0.0, 42)
    for name, L, seed, ups, c in (("half10_t0", 10, 1234, 0, 0.1), ("half10_t3", 10, 77, 3, 0.0),
                                  ("half12_t40", 12, 5, 40, 0.2)):
        st, th, ph, T, df = O.half_grown_lattice(L, seed=seed, T_updates=0, grain=4)
        for _ in range(ups):
            T = ts.update_temperature_cet(T, st, dt=1e-6)
        rates_case(ref, name, st, th, ph, T, df, c, 7 + ups)
    # general inputs: empty sites that carry an orientation, T below 1 K and NaN-free extremes
    rng = np.random.default_rng(99)
    st, th, ph, T, df = O.half_grown_lattice(9, seed=3, grain=3)
    th = rng.uniform(0, np.pi, th.shape); ph = rng.uniform(0, 2 * np.pi, ph.shape)
    T = T.copy(); T[0, 0, :] = 0.5; T[-1, :, 0] = 3690.0; T[-1, :, 1] = 5000.0
    rates_case(ref, "general9", st, th, ph, T, df, 0.05, 11)

    # ---- thermal ---------------------------------------------------------------------------
    rng = np.random.default_rng(7)
    T0 = 2800 + 900 * rng.random((9, 11, 13))
    seq, T = {}, T0
    for n in range(1, 41):
        T = ts.update_temperature_cet(T, None, dt=1e-6)
        if n in (1, 2, 5, 40):
            seq[f"T_after_{n}"] = T
    np.savez_compressed(os.path.join(OUT, "thermal_cet_random.npz"), T0=T0, **seq)
    _, _, _, T0, _ = li.initialize_lattice(lattice_size=12, n_seeds=3)
    seq, T = {}, T0
    for n in range(1, 61):
        T = ts.update_temperature_cet(T, None, dt=1e-6)
        if n in (1, 3, 10, 60):
            seq[f"T_after_{n}"] = T
    np.savez_compressed(os.path.join(OUT, "thermal_cet_gradient12.npz"), T0=T0, **seq)
    st, th, ph, T0, df = O.half_grown_lattice(10, seed=21, grain=4)
    prev = st.copy(); prev[rng.random(st.shape) < 0.2] = 0
    Tn = ts.update_temperature(T0, st, prev, 1e-7, (3, 4.5), 200.0)
    Tn2 = ts.update_temperature(T0, st, prev, 1e-9, (0, 2), 50.0, beam_radius=20e-6, absorptivity=0.5)
    np.savez_compressed(os.path.join(OUT, "thermal_full10.npz"), T0=T0, state=st.astype(np.int8),
                        prev_state=prev.astype(np.int8), T1=Tn, T2=Tn2)
    print("thermal fixtures written")

    grains_cases(ref)
    defects_cases(ref)

    for name in TRAJ_CASES:
        if not name.startswith("L30"):
            traj_case(ref, name)


if __name__ == "__main__":
    main()
