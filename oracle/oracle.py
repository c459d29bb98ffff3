"""ctypes front-end of the CPU oracle (oracle.c) + NumPy restatements of the host
set-up code the hot path depends on.

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package never
imports this module.

Parity status: "unpinned" by the reference's own tests (it has none); pinned
instead against the reference itself (tests/test_oracle_vs_reference.py, run in
the build container where /root/reference exists) and against the fixtures that
oracle/gen_golden.py generated from the reference (tests/golden/).
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

# ---------------------------------------------------------------------------
# constants.py values the hot path consumes (file:line of the reference)
# ---------------------------------------------------------------------------
CONSTANTS = dict(
    LATTICE_SIZE=30,            # constants.py:54
    VOXEL_SIZE=5e-6,            # constants.py:55
    N_STEPS=20000,              # constants.py:56
    METRIC_UPDATE_STEP=200,     # constants.py:57
    N_SEEDS=20,                 # constants.py:59
    K_T=8.617333262e-5,         # constants.py:64
    T_MELT=3695,                # constants.py:65
    T_SUB=2800,                 # constants.py:66
    ATOMIC_SPACING_W=2.74e-10,  # constants.py:68
    NU=1e13,                    # constants.py:71
    NU_DEP=2e13,                # constants.py:72
    E_B_W=3.8, E_DIFF_W=0.35,   # constants.py:75-76
    E_B_RE=4.2, E_DIFF_RE=0.50, IMPURITY_RE=0.10,   # constants.py:81-83
    E_B_C=3.2, E_DIFF_C=0.30,   # constants.py:85-86
    MAX_IMP_FRACTION=1.0,       # constants.py:91
    ANISOTROPY_FACTOR=0.25,     # constants.py:104
    CET_EQ_THRESHOLD=0.50, CET_AR_THRESHOLD=3.0,    # constants.py:109-110
    DELTA_T_C=10,               # constants.py:125
    I0=5e13, K_NUC=500, BETA_IMP_NUC=0.4,           # constants.py:131-133
    DEFECT_PROB=3e-3, DEFECT_PROB_BASE=0.12,        # constants.py:139-140
    RATE_THRESHOLD=1e-30,       # constants.py:146
    RANDOM_SEED=42,             # constants.py:148
    DEFECT_ID=4,                # constants.py:43
)
# thermal_solver.py:6-9
K_COND, RHO, CP = 173.0, 19300.0, 132.0
ALPHA = K_COND / (RHO * CP)

EV_NAMES = (b"dep", b"diff", b"nuc", b"att")


class Params(C.Structure):
    _fields_ = [("nu", C.c_double), ("nu_dep", C.c_double),
                ("E_b", C.c_double * 3), ("E_diff", C.c_double * 3),
                ("kT", C.c_double), ("T_melt", C.c_double), ("i0", C.c_double),
                ("delta_T_c", C.c_double), ("k_nuc", C.c_double), ("beta_imp_nuc", C.c_double),
                ("max_imp_fraction", C.c_double), ("rate_threshold", C.c_double),
                ("anisotropy", C.c_double), ("impurity_re", C.c_double),
                ("impurity_c", C.c_double),
                ("states_w", C.c_int32), ("states_re", C.c_int32), ("states_c", C.c_int32),
                ("pad_", C.c_int32)]


class ThermalParams(C.Structure):
    _fields_ = [("dt_alpha", C.c_double), ("inv_dx2", C.c_double), ("lo", C.c_double),
                ("hi", C.c_double), ("nan_value", C.c_double), ("every", C.c_int64)]


def make_params(impurity_c: float = 0.0, consts: Optional[dict] = None) -> Params:
    k = dict(CONSTANTS)
    if consts:
        k.update(consts)
    p = Params()
    p.nu, p.nu_dep = k["NU"], k["NU_DEP"]
    p.E_b[:] = [k["E_B_W"], k["E_B_RE"], k["E_B_C"]]
    p.E_diff[:] = [k["E_DIFF_W"], k["E_DIFF_RE"], k["E_DIFF_C"]]
    p.kT, p.T_melt, p.i0, p.delta_T_c = k["K_T"], k["T_MELT"], k["I0"], k["DELTA_T_C"]
    p.k_nuc, p.beta_imp_nuc, p.max_imp_fraction = k["K_NUC"], k["BETA_IMP_NUC"], k["MAX_IMP_FRACTION"]
    p.rate_threshold, p.anisotropy = k["RATE_THRESHOLD"], k["ANISOTROPY_FACTOR"]
    p.impurity_re, p.impurity_c = k["IMPURITY_RE"], impurity_c
    p.states_w, p.states_re, p.states_c = 1, 2, 3      # kmc_simulation.py:255
    return p


def make_thermal_params(dt: float = 1e-6, consts: Optional[dict] = None) -> ThermalParams:
    k = dict(CONSTANTS)
    if consts:
        k.update(consts)
    tp = ThermalParams()
    tp.dt_alpha = dt * ALPHA                                   # thermal_solver.py:116 (dt*ALPHA first)
    tp.inv_dx2 = 1.0 / (k["VOXEL_SIZE"] * k["VOXEL_SIZE"])     # thermal_solver.py:114
    tp.lo, tp.hi = float(k["T_SUB"]), k["T_MELT"] * 1.1        # thermal_solver.py:117
    tp.nan_value = float(k["T_SUB"])                           # kmc_simulation.py:249
    tp.every = 20                                              # kmc_simulation.py:248
    return tp


_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.isfile(_LIB_PATH) or \
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "oracle.c")):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int64)
    vp = C.c_void_p
    L.oracle_bcc_neighbors.argtypes = [C.c_int64] * 4 + [vp]
    L.oracle_bcc_neighbors.restype = C.c_int
    L.oracle_misorientation.argtypes = [C.c_double] * 4
    L.oracle_misorientation.restype = C.c_double
    L.oracle_event_rates.argtypes = [vp, vp, vp, vp, vp, C.c_int64, C.POINTER(Params), vp, ip,
                                     vp, vp, vp, vp, vp, C.c_int64]
    L.oracle_event_rates.restype = C.c_int64
    L.oracle_site_rates.argtypes = [vp, vp, vp, vp, vp, C.c_int64, C.POINTER(Params), vp, vp, vp]
    L.oracle_site_rates.restype = C.c_int64
    L.oracle_pysum.argtypes = [vp, C.c_int64]
    L.oracle_pysum.restype = C.c_double
    L.oracle_thermal_cet.argtypes = [vp, vp, C.c_int64, C.c_int64, C.c_int64] + [C.c_double] * 4 + \
                                    [C.c_int, C.c_double]
    L.oracle_thermal_cet.restype = None
    L.oracle_thermal_full.argtypes = [vp, vp, vp, vp, C.c_int64, C.c_int64, C.c_int64,
                                      C.c_double, C.c_double, C.c_double, vp, C.c_double,
                                      C.c_double, C.c_double, C.c_double]
    L.oracle_thermal_full.restype = None
    L.oracle_kmc_run.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int64, C.POINTER(Params),
                                 C.c_int64, C.c_int64, C.c_double, C.POINTER(ThermalParams),
                                 vp, ip, vp, ip, vp, ip, dp, ip, C.POINTER(C.c_int),
                                 vp, vp, vp, vp, vp, vp]
    L.oracle_kmc_run.restype = C.c_int64
    L.oracle_clusters.argtypes = [vp, vp, vp, C.c_int64, C.c_double, vp, C.c_int64, vp, vp, vp]
    L.oracle_clusters.restype = C.c_int64
    _lib = L
    return L


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def bcc_neighbors(i, j, k, L):
    out = np.zeros((14, 3), dtype=np.int64)
    n = lib().oracle_bcc_neighbors(i, j, k, L, _ptr(out))
    return out[:n]


def misorientation(t1, p1, t2, p2):
    return lib().oracle_misorientation(t1, p1, t2, p2)


def event_rates(state, theta, phi, T, defects, L, params: Params, species_draws=None):
    """SoA event list in the reference's order (kmc_event_rates.py:162-176)."""
    state, defects = _i64(state), _i64(defects)
    theta, phi, T = _f64(theta), _f64(phi), _f64(T)
    sd = None if species_draws is None else _f64(species_draws)
    used = C.c_int64(0)
    cap = max(1024, 2 * L ** 3)
    while True:
        ty = np.empty(cap, np.uint8); po = np.empty(cap, np.int64); ra = np.empty(cap, np.float64)
        ta = np.empty(cap, np.int64); at = np.empty(cap, np.int32)
        n = lib().oracle_event_rates(_ptr(state), _ptr(theta), _ptr(phi), _ptr(T), _ptr(defects), L,
                                     C.byref(params), _ptr(sd), C.byref(used),
                                     _ptr(ty), _ptr(po), _ptr(ra), _ptr(ta), _ptr(at), cap)
        if n <= cap:
            break
        cap = n
    return dict(type=ty[:n], pos=po[:n], rate=ra[:n], target=ta[:n], atom=at[:n],
                draws_used=used.value)


def events_as_tuples(ev, L):
    """The reference's list-of-tuples form (bytes, (i,j,k), float, (i,j,k), int)."""
    out = []
    LL = L * L
    for ty, po, ra, ta, at in zip(ev["type"].tolist(), ev["pos"].tolist(), ev["rate"].tolist(),
                                  ev["target"].tolist(), ev["atom"].tolist()):
        p = (po // LL, (po // L) % L, po % L)
        t = (-1, -1, -1) if ta < 0 else (ta // LL, (ta // L) % L, ta % L)
        out.append((EV_NAMES[ty], p, ra, t, at))
    return out


def site_rates(state, theta, phi, T, defects, L, params: Params, want_counts=False):
    state, defects = _i64(state), _i64(defects)
    theta, phi, T = _f64(theta), _f64(phi), _f64(T)
    sr = np.zeros(L ** 3, np.float64)
    dep = np.full(L * L, np.nan)
    cnt = np.zeros(L ** 3, np.int32) if want_counts else None
    n = lib().oracle_site_rates(_ptr(state), _ptr(theta), _ptr(phi), _ptr(T), _ptr(defects), L,
                                C.byref(params), _ptr(sr), _ptr(dep), _ptr(cnt))
    return sr.reshape(L, L, L), dep.reshape(L, L), n, (None if cnt is None else cnt.reshape(L, L, L))


def pysum(x):
    x = _f64(x)
    return lib().oracle_pysum(_ptr(x), x.size)


def thermal_cet(T, dt=1e-6, nan_to_num=False, consts=None):
    """thermal_solver.py:107-117 (optionally preceded by kmc_simulation.py:249)."""
    T = _f64(T)
    tp = make_thermal_params(dt, consts)
    out = np.empty_like(T)
    n0, n1, n2 = T.shape
    lib().oracle_thermal_cet(_ptr(T), _ptr(out), n0, n1, n2, tp.dt_alpha, tp.inv_dx2, tp.lo, tp.hi,
                             1 if nan_to_num else 0, tp.nan_value)
    return out


def laser_q_top(L, laser_pos, laser_power, beam_radius=50e-6, absorptivity=0.35, voxel=5e-6):
    """thermal_solver.py:80-94: I_surface / VOXEL_SIZE on the top plane (note j0 used twice)."""
    i0, j0 = laser_pos
    jj = np.arange(L, dtype=np.float64)
    JJ, KK = np.meshgrid(jj, jj, indexing="ij")
    r_m = np.sqrt((JJ - j0) ** 2 + (KK - j0) ** 2) * voxel
    area = np.pi * beam_radius * beam_radius
    I = (laser_power * absorptivity / area) * np.exp(-(r_m ** 2) / (beam_radius ** 2))
    return I / voxel


def thermal_full(T, state, prev_state, dt, laser_pos, laser_power, beam_radius=50e-6,
                 absorptivity=0.35, consts=None):
    """thermal_solver.py:36-105."""
    k = dict(CONSTANTS)
    if consts:
        k.update(consts)
    T = _f64(T); state = _i64(state); prev_state = _i64(prev_state)
    n0, n1, n2 = T.shape
    q = _f64(laser_q_top(n0, laser_pos, laser_power, beam_radius, absorptivity, k["VOXEL_SIZE"]))
    out = np.empty_like(T)
    lib().oracle_thermal_full(_ptr(T), _ptr(state), _ptr(prev_state), _ptr(out), n0, n1, n2,
                              dt, ALPHA, 1.0 / (k["VOXEL_SIZE"] * k["VOXEL_SIZE"]), _ptr(q),
                              RHO * CP, 200e3 / CP, float(k["T_SUB"]), k["T_MELT"] * 1.1)
    return out


def kmc_run(state, atom_type, theta, phi, T, defects, L, params, step0, n_steps, defect_fraction,
            py_draws, np_draws, sp_draws, thermal: Optional[ThermalParams] = None, log=True,
            total_time0=0.0):
    """kmc_simulation.py:246-332 for steps [step0, step0+n_steps); arrays are updated in place
    (they must be C-contiguous int64 / float64).  Returns a dict with counters and the event log."""
    for a, dt in ((state, np.int64), (atom_type, np.int64), (theta, np.float64), (phi, np.float64),
                  (T, np.float64)):
        assert a.dtype == dt and a.flags.c_contiguous
    defects = _i64(defects)
    py_draws, np_draws = _f64(py_draws), _f64(np_draws)
    sp = None if sp_draws is None else _f64(sp_draws)
    tp = thermal or make_thermal_params()
    pp, npp, spp = C.c_int64(0), C.c_int64(0), C.c_int64(0)
    tt, nc, term = C.c_double(total_time0), C.c_int64(0), C.c_int(0)
    lg = {}
    if log:
        lg = dict(type=np.zeros(n_steps, np.uint8), pos=np.zeros(n_steps, np.int64),
                  target=np.zeros(n_steps, np.int64), atom=np.zeros(n_steps, np.int32),
                  rate=np.zeros(n_steps, np.float64), total=np.zeros(n_steps, np.float64))
    done = lib().oracle_kmc_run(_ptr(state), _ptr(atom_type), _ptr(theta), _ptr(phi), _ptr(T),
                                _ptr(defects), L, C.byref(params), step0, n_steps, defect_fraction,
                                C.byref(tp), _ptr(py_draws), C.byref(pp), _ptr(np_draws), C.byref(npp),
                                _ptr(sp), C.byref(spp), C.byref(tt), C.byref(nc), C.byref(term),
                                _ptr(lg.get("type")), _ptr(lg.get("pos")), _ptr(lg.get("target")),
                                _ptr(lg.get("atom")), _ptr(lg.get("rate")), _ptr(lg.get("total")))
    res = dict(steps_done=done, py_used=pp.value, np_used=npp.value, sp_used=spp.value,
               total_time=tt.value, nucleation_count=nc.value, terminated=bool(term.value))
    for k_, v in lg.items():
        res["log_" + k_] = v[:done]
    return res


# ---------------------------------------------------------------------------
# Host set-up restatements (NumPy; out of the hot path, needed to build inputs)
# ---------------------------------------------------------------------------
def initialize_lattice(L, n_seeds=20, T_sub=2800, T_melt=3695, random_seed=42, impurity_c=0.0,
                       rng=None):
    """lattice_init.py:10-59 — same draws in the same order from the legacy MT19937 stream.
    `rng` defaults to a fresh RandomState(random_seed) (the reference seeds the global one)."""
    rs = rng if rng is not None else np.random.RandomState(random_seed)
    state = np.zeros((L, L, L), dtype=np.int64)
    theta = np.zeros((L, L, L)); phi = np.zeros((L, L, L))
    atom_type = np.zeros((L, L, L), dtype=np.int64)
    G = (T_melt - T_sub) / L                                      # :28
    T = np.ascontiguousarray(np.broadcast_to(T_sub + G * np.arange(L), (L, L, L)))   # :29-32
    seeds = rs.choice(L * L, n_seeds, replace=False)              # :35
    for s in seeds:
        x, y = int(s) // L, int(s) % L
        r = rs.random_sample()                                    # :43
        atom = 2 if r < 0.10 else (3 if r < 0.10 + impurity_c else 1)   # :44-49
        state[x, y, 0] = atom; atom_type[x, y, 0] = atom
        theta[x, y, 0] = rs.uniform(0, np.pi)                     # :53
        phi[x, y, 0] = rs.uniform(0, 2 * np.pi)                   # :54
    return state, theta, phi, T, atom_type


def track_defects(atom_type, T, rng):
    """defects.py:4-19 — defect mask on carbon sites, one legacy-stream draw per C site."""
    k = CONSTANTS
    defects = np.zeros(atom_type.shape, dtype=np.int64)
    c_sites = (atom_type == 3)
    if np.any(c_sites):
        Tv = T[c_sites]
        valid = np.where(Tv > 0, Tv, k["T_SUB"])
        prob = np.clip(k["DEFECT_PROB_BASE"] * np.exp(-0.3 / (k["K_T"] * valid)), 0.0, 1.0)
        defects[c_sites] = (rng.random_sample(int(c_sites.sum())) < prob).astype(np.int64)
    return defects


def half_grown_lattice(L, seed=1234, T_updates=0, grain=8):
    """SURVEY §8(d)(ii) synthetic 'half-grown' input (salt-and-pepper solid below a wavy front,
    8^3 piecewise-constant grains, linear gradient along axis 2, defects on 5 % of C sites)."""
    rng = np.random.default_rng(seed)
    i = np.arange(L)[:, None, None]; k = np.arange(L)[None, None, :]
    front = L / 2 + 8 * np.sin(2 * np.pi * i / L)
    u = rng.random((L, L, L))
    species = rng.choice(np.array([1, 2, 3, 4]), size=(L, L, L), p=[.85, .10, .04, .01])
    state = np.where((u < 0.5) & (k < front), species, 0).astype(np.int64)
    g = (L + grain - 1) // grain
    th_b = rng.uniform(0, np.pi, (g, g, g)); ph_b = rng.uniform(0, 2 * np.pi, (g, g, g))
    up = lambda a: np.repeat(np.repeat(np.repeat(a, grain, 0), grain, 1), grain, 2)[:L, :L, :L]
    solid = (state >= 1) & (state <= 3)
    theta = np.where(solid, up(th_b), 0.0); phi = np.where(solid, up(ph_b), 0.0)
    T = np.ascontiguousarray(np.broadcast_to(2800 + 895.0 * np.arange(L) / L, (L, L, L))).copy()
    for _ in range(T_updates):
        T = thermal_cet(T)
    defects = ((state == 3) & (rng.random((L, L, L)) < 0.05)).astype(np.int64)
    return state, np.ascontiguousarray(theta), np.ascontiguousarray(phi), T, defects


class DrawStreams:
    """The three MT19937 streams run_kmc consumes, pre-generated so both the
    oracle and the GPU path can be fed identical draws."""

    def __init__(self, seed=42, n_py=0, n_np=0, n_sp=0):
        import random as _r
        r = _r.Random(seed)
        self.py = np.array([r.random() for _ in range(n_py)], dtype=np.float64)
        self.np = np.random.RandomState(seed + 1).random_sample(n_np)
        self.sp = np.random.RandomState(seed + 2).random_sample(n_sp)


def pi_times(u):
    return math.pi * u


def run_kmc(L=30, n_steps=20000, temp=2800, defect_fraction=0.0, n_seeds=5, impurity_c=0.0,
            seed=42, species_seed=42, on_metric_step=None, log=False):
    """kmc_simulation.py:203-398 with the three RNG streams made explicit.

    `seed` plays RANDOM_SEED for Python's `random` (:220) and NumPy's legacy global stream
    (:219, lattice_init.py:20); `species_seed` seeds the stream that stands in for Numba's
    private generator (kmc_event_rates.py:65; == nb_seed(species_seed) in refharness).
    `on_metric_step(step, state, atom_type, theta, phi, T, defects)` is called where the
    reference computes a metrics row (:341).  Returns (state, atom_type, total_time, theta,
    phi, info)."""
    import random as _random
    rs = np.random.RandomState(seed)
    pyr = _random.Random(seed)
    spr = np.random.RandomState(species_seed)
    state, theta, phi, T, atom_type = initialize_lattice(L, n_seeds, temp, random_seed=seed,
                                                         impurity_c=impurity_c, rng=rs)
    defects = track_defects(atom_type, T, rs)                       # :231
    params = make_params(impurity_c)
    per_step = 3 if defect_fraction > 0.0 else 2
    total_time, nuc, step, terminated = 0.0, 0, 0, False
    logs = []
    while step < n_steps and not terminated:
        last = min(((step + 199) // 200) * 200, n_steps - 1)
        nb = last - step + 1
        st_py, st_np, st_sp = pyr.getstate(), rs.get_state(), spr.get_state()
        py = np.array([pyr.random() for _ in range(per_step * nb)])
        npd = rs.random_sample(2 * nb)
        spd = spr.random_sample(nb * L * L)
        res = kmc_run(state, atom_type, theta, phi, T, defects, L, params, step, nb,
                      defect_fraction, py, npd, spd, log=log, total_time0=total_time)
        pyr.setstate(st_py); [pyr.random() for _ in range(res["py_used"])]
        rs.set_state(st_np); rs.random_sample(res["np_used"])
        spr.set_state(st_sp); spr.random_sample(res["sp_used"])
        total_time = res["total_time"]; nuc += res["nucleation_count"]
        if log:
            logs.append(res)
        step += res["steps_done"]
        if res["terminated"]:
            terminated = True
            break
        s_last = step - 1
        if s_last % 200 == 0:                                        # :335-338
            defects = track_defects(atom_type, T, rs)
        if on_metric_step is not None and (s_last % 200 == 0 or s_last == n_steps - 1):
            on_metric_step(s_last, state, atom_type, theta, phi, T, defects)
    info = dict(steps=step, terminated=terminated, nucleation_count=nuc, T=T, defects=defects,
                logs=logs)
    return state, atom_type, total_time, theta, phi, info


# ---------------------------------------------------------------------------
# Grain clustering and observables (utils.py:28-84,104-111; metrics.py:41-105) — pure-Python
# restatement for small lattices (the reference's own DFS is pure Python too).
# ---------------------------------------------------------------------------
_NB_OFFSETS = ((1, 1, 0), (1, -1, 0), (-1, 1, 0), (-1, -1, 0), (0, 1, 1), (0, 1, -1), (0, -1, 1), (0, -1, -1),
               (2, 0, 0), (-2, 0, 0), (0, 2, 0), (0, -2, 0), (0, 0, 2), (0, 0, -2))   # kmc_event_rates.py:29-36


def _py_misorientation(t1, p1, t2, p2):
    """kmc_event_rates.py:10-23"""
    v1 = (math.sin(t1) * math.cos(p1), math.sin(t1) * math.sin(p1), math.cos(t1))
    v2 = (math.sin(t2) * math.cos(p2), math.sin(t2) * math.sin(p2), math.cos(t2))
    dot = v1[0] * v2[0] + v1[1] * v2[1] + v1[2] * v2[2]
    return math.acos(max(min(dot, 1.0), -1.0))


def get_clusters(state, theta, phi=None, theta_threshold=0.5):
    """utils.py:28-84: DFS over the 14-neighbourhood; clusters in discovery (raster) order, each
    a list of (i, j, k) in DFS order; `visited` holds the cluster number 1.. per site."""
    Lx, Ly, Lz = state.shape
    visited = np.zeros(state.shape, dtype=np.int32)
    occ = (state != 0).tolist()
    th = theta.tolist()
    ph = phi.tolist() if phi is not None else None
    vis = visited.tolist()
    clusters, label = [], 1
    for i in range(Lx):
        for j in range(Ly):
            for k in range(Lz):
                if not occ[i][j][k] or vis[i][j][k]:
                    continue
                cluster, stack = [(i, j, k)], [(i, j, k)]
                vis[i][j][k] = label
                while stack:
                    ci, cj, ck = stack.pop()
                    for di, dj, dk in _NB_OFFSETS:
                        ni, nj, nk = ci + di, cj + dj, ck + dk
                        # get_bcc_neighbors bounds every axis by Lx (utils.py:46), then :49 by the true extents
                        if not (0 <= ni < Lx and 0 <= nj < Lx and 0 <= nk < Lx):
                            continue
                        if not (nj < Ly and nk < Lz) or not occ[ni][nj][nk] or vis[ni][nj][nk]:
                            continue
                        if ph is None:
                            mis = abs(th[ci][cj][ck] - th[ni][nj][nk])
                        else:
                            mis = _py_misorientation(th[ci][cj][ck], ph[ci][cj][ck], th[ni][nj][nk], ph[ni][nj][nk])
                        if mis < theta_threshold:
                            vis[ni][nj][nk] = label
                            cluster.append((ni, nj, nk))
                            stack.append((ni, nj, nk))
                clusters.append(cluster)
                label += 1
    return clusters, np.array(vis, dtype=np.int32).reshape(state.shape)


def clusters_fast(state, theta, phi=None, theta_threshold=0.5):
    """utils.py:28-84 in C (oracle.c: oracle_clusters) for lattices the pure-Python DFS above is too
    slow for.  Returns dict(n, visited, size, box_lo, box_hi) with clusters in discovery order."""
    L = state.shape[0]
    assert state.shape == (L, L, L)
    st, th = _i64(state), _f64(theta)
    ph = None if phi is None else _f64(phi)
    visited = np.zeros(state.shape, np.int32)
    cap = int(np.count_nonzero(st))
    size = np.zeros(max(cap, 1), np.int32); lo = np.zeros((max(cap, 1), 3), np.int32); hi = np.zeros((max(cap, 1), 3), np.int32)
    n = lib().oracle_clusters(_ptr(st), _ptr(th), _ptr(ph), L, float(theta_threshold), _ptr(visited), cap,
                              _ptr(size), _ptr(lo), _ptr(hi))
    return dict(n=int(n), visited=visited, size=size[:n], box_lo=lo[:n], box_hi=hi[:n])


def calculate_aspect_ratio(cluster):
    """utils.py:104-111"""
    c = np.array(cluster)
    dims = c.max(axis=0) - c.min(axis=0) + 1
    return float(np.max(dims)) / float(max(np.min(dims), 1))


def compute_metrics(state, theta, phi, defects=None, voxel_size=None):
    """metrics.py:41-96, the keys the driver consumes (kmc_simulation.py:341-378)."""
    voxel_size = CONSTANTS["VOXEL_SIZE"] if voxel_size is None else voxel_size
    clusters, visited = get_clusters(state, theta, phi, 0.5)
    if not clusters:
        return {"AspectRatio": 0.0, "EquiaxedFraction": 0.0, "NucleationDensity": 0.0, "AvgGrainSize": 0.0,
                "GrainCount": 0, "DefectDensity": 0.0, "Grain_d50_um": 0.0, "Grain_d90_um": 0.0}
    ars = [calculate_aspect_ratio(c) for c in clusters]
    volume = state.size * (voxel_size ** 3)
    def_count = np.sum(defects) if defects is not None else 0
    diam = ((6.0 * (np.array(visited) * (voxel_size ** 3)) / np.pi) ** (1.0 / 3.0)) * 1e6   # metrics.py:43,76
    return {"AspectRatio": np.mean(ars), "EquiaxedFraction": np.mean(np.array(ars) < CONSTANTS["CET_AR_THRESHOLD"]),
            "NucleationDensity": len(clusters) / volume, "AvgGrainSize": np.mean([len(c) for c in clusters]) * voxel_size * 1e6,
            "GrainCount": len(clusters), "DefectDensity": def_count / volume,
            "Grain_d50_um": np.median(diam), "Grain_d90_um": np.percentile(diam, 90)}


def grown_lattice(L, seed=1, grain=5, fill=0.7, jitter=0.15):
    """Clustering input: `fill` of the sites occupied, piecewise-constant grain orientations of
    edge `grain` with a per-site orientation jitter (rad), so that edges fall on both sides of
    the 0.5 rad threshold; 2 % defect sites (state 4, theta = phi = 0)."""
    rng = np.random.default_rng(seed)
    g = (L + grain - 1) // grain
    up = lambda a: np.repeat(np.repeat(np.repeat(a, grain, 0), grain, 1), grain, 2)[:L, :L, :L]
    th = up(rng.uniform(0, np.pi, (g, g, g))) + jitter * rng.standard_normal((L, L, L))
    ph = up(rng.uniform(0, 2 * np.pi, (g, g, g))) + jitter * rng.standard_normal((L, L, L))
    state = np.where(rng.random((L, L, L)) < fill, rng.choice(np.array([1, 2, 3]), size=(L, L, L), p=[.8, .15, .05]), 0)
    state = np.where((state != 0) & (rng.random((L, L, L)) < 0.02), 4, state).astype(np.int64)
    solid = (state >= 1) & (state <= 3)
    return state, np.ascontiguousarray(np.where(solid, th, 0.0)), np.ascontiguousarray(np.where(solid, ph, 0.0))
