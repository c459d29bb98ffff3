"""Diagnostic: event-type breakdown of the serial oracle vs the sublattice sweeps at equal event counts."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import oracle as O
import cetkmc
from cetkmc._config import rate_params, thermal_params
O.build()
L, c, dfrac, n_events = 30, 0.1, 3e-3, 2000
eps = float(sys.argv[1]) if len(sys.argv) > 1 else 20.0
n_seeds = int(sys.argv[2]) if len(sys.argv) > 2 else 6
acc = []
for seed in range(n_seeds):
    st, th, ph, T, at = O.initialize_lattice(L, n_seeds=20, random_seed=seed, impurity_c=c)
    df = np.zeros_like(st)
    d = O.DrawStreams(seed=seed, n_py=3 * n_events, n_np=2 * n_events, n_sp=n_events * L * L)
    o = [a.copy() for a in (st, at, th, ph, T)]
    r = O.kmc_run(o[0], o[1], o[2], o[3], o[4], df, L, O.make_params(c), 0, n_events, dfrac, d.py, d.np, d.sp,
                  thermal=O.make_thermal_params(), log=True)
    ty = np.bincount(r["log_type"], minlength=4)
    ctx = cetkmc.Context(L=L)
    ctx.set_rate_params(rate_params(c))
    ctx.upload(state=st, theta=th, phi=ph, T=T, defects=df)
    sp = cetkmc._lib.SweepParams()
    sp.seed, sp.events_per_sweep, sp.p_max, sp.defect_fraction, sp.thermal_every = 1000 + seed, eps, 0.1, dfrac, 0
    tp = thermal_params(1e-6, nan_to_num=True)
    applied = fired = nuc = sweeps = nth = 0
    while applied < n_events:
        while nth <= applied // 20:
            ctx.thermal_cet(tp); nth += 1
        res = ctx.sweep_run(1, sp, None)
        applied += res["events_applied"]; fired += res["events_fired"]; nuc += res["nucleation_count"]; sweeps += 1
    f = ctx.download(state=True, T=True)
    ctx.close()
    occ_o, occ_g = int((o[0] != 0).sum()), int((f["state"] != 0).sum())
    print(f"seed {seed}: oracle dep/diff/nuc/att = {ty.tolist()} occ {occ_o} | sweeps {sweeps} fired {fired} applied {applied} "
          f"nuc {nuc} occ {occ_g} diff~{applied - (occ_g - 20)} | T>Tmelt sites oracle {(o[4] > 3695).sum()} gpu {(f['T'] > 3695).sum()} "
          f"Tdiff {np.abs(o[4] - f['T']).max():.3g}")
    acc.append([ty[1], applied - (occ_g - 20), ty[2], nuc, ty[3] + ty[0], occ_g - 20 - nuc, applied - n_events])
a = np.array(acc, float)
d = a[:, 1::2][:, :3] - a[:, 0::2][:, :3]
print("mean oracle diff/nuc/att+dep:", a[:, 0].mean(), a[:, 2].mean(), a[:, 4].mean(), " gpu:", a[:, 1].mean(), a[:, 3].mean(), a[:, 5].mean(),
      " overshoot", a[:, 6].mean())
print("gpu - oracle (diff, nuc, att+dep):", d.mean(0), "+-", d.std(0, ddof=1) / np.sqrt(len(a)))
