"""Large-lattice drivers and on-disk contracts around the hot path.

    save_lattice / load_lattice      the reference's five-file .npy snapshot (lattice_init.py:98-105)
    checkpoint / resume              the same set plus the defect mask and the run counters: a
                                     run can be stopped and continued (the reference cannot)
    run_cet_sublattice               the large-lattice analogue of run_kmc (kmc_simulation.py:203-398):
                                     synchronous-sublattice sweeps on the resident lattice, the
                                     200-step cadence (defect-mask refresh, clustering, metrics row)
                                     on the device, outputs/<prefix>/metrics.csv with the reference's
                                     18 columns; optionally the laser / latent-heat thermal step
                                     (thermal_solver.update_temperature, :36-105) with a melt pool
                                     moving along axis 1 at a fixed speed — SURVEY §8d config 3
    run_gr_sweep                     config 5: independent lattices over a grid of (G, R), one per
                                     GPU (replicas only, no communication), one cet_map.csv and a tree
                                     the unmodified plot_cet.py reads (write_plot_cet_tree)
"""
from __future__ import annotations

import json
import os

import numpy as np

from . import _lib
from . import defects as _defects
from . import kmc_simulation as _ks
from . import metrics as _metrics
from ._config import constants, rate_params, thermal_full_params, thermal_params
from .thermal_solver import DEFAULT_ABSORPTIVITY, DEFAULT_BEAM_RADIUS, laser_source_top

_FILES = ("state", "orientation_theta", "orientation_phi", "temperature", "atom_type")


def save_lattice(state, orientation_theta, orientation_phi, T, atom_type, prefix="init"):
    """lattice_init.py:98-105 — `<prefix>_{state,orientation_theta,orientation_phi,temperature,atom_type}.npy`."""
    for name, a in zip(_FILES, (state, orientation_theta, orientation_phi, T, atom_type)):
        np.save(f"{prefix}_{name}.npy", a)


def load_lattice(prefix="init"):
    """Inverse of save_lattice: (state, theta, phi, T, atom_type) in the reference's dtypes."""
    st, th, ph, T, at = (np.load(f"{prefix}_{name}.npy") for name in _FILES)
    return (np.ascontiguousarray(st, dtype=np.int64), np.ascontiguousarray(th, dtype=np.float64),
            np.ascontiguousarray(ph, dtype=np.float64), np.ascontiguousarray(T, dtype=np.float64),
            np.ascontiguousarray(at, dtype=np.int64))


def checkpoint(ctx, prefix, **meta):
    """Write the resident lattice as the reference's snapshot set plus `<prefix>_defects.npy` and
    `<prefix>_meta.json` (run counters)."""
    f = ctx.download(state=True, atom_type=True, theta=True, phi=True, T=True)
    save_lattice(f["state"], f["theta"], f["phi"], f["T"], f["atom_type"], prefix=prefix)
    np.save(f"{prefix}_defects.npy", (ctx.download_packed() >> 4).astype(np.int64))
    idx, tau, t = ctx.sweep_state()
    meta = dict(meta, sweep_clock=[idx, float.hex(tau), float.hex(t)], numpy_rng=_rng_to_json(np.random.get_state()))
    with open(f"{prefix}_meta.json", "w") as fh:
        json.dump(meta, fh)


def resume(ctx, prefix):
    """Load a checkpoint into `ctx`; returns the meta dict."""
    st, th, ph, T, _at = load_lattice(prefix)
    df = np.load(f"{prefix}_defects.npy") if os.path.exists(f"{prefix}_defects.npy") else None
    ctx.upload(state=st, theta=th, phi=ph, T=T, defects=df)
    meta = {}
    if os.path.exists(f"{prefix}_meta.json"):
        with open(f"{prefix}_meta.json") as fh:
            meta = json.load(fh)
    if "sweep_clock" in meta:
        idx, tau, t = meta["sweep_clock"]
        ctx.sweep_set_state(int(idx), float.fromhex(tau), float.fromhex(t))
    if "numpy_rng" in meta:
        np.random.set_state(_rng_from_json(meta["numpy_rng"]))
    return meta


def _rng_to_json(st):
    return [st[0], np.asarray(st[1]).tolist(), int(st[2]), int(st[3]), float(st[4])]


def _rng_from_json(j):
    return (j[0], np.array(j[1], dtype=np.uint32), int(j[2]), int(j[3]), float(j[4]))


def _gr(L, temp, nu_dep):
    """kmc_simulation.py:236-239 with the run's own T_sub / NU_DEP."""
    G = (constants.T_MELT - temp) / (L * constants.VOXEL_SIZE)
    R = nu_dep * 2.74e-10 / constants.VOXEL_SIZE
    R_phys = nu_dep * constants.ATOMIC_SPACING_W
    return G, R, R_phys, (G / R_phys if R_phys > 0 else np.inf)


def _slab_comm(world):
    """Rendezvous helpers of a multi-GPU run (one process per GPU under torch.distributed.run): the
    process group is only plumbing — NCCL unique id broadcast and the small host-side gathers of the
    metrics cadence; the lattice exchange itself is libcetkmc's (comm.cu)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        dist.init_process_group("nccl" if torch.cuda.is_available() else "gloo")

    def all_gather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    def bcast_bytes(b):
        box = [b]
        dist.broadcast_object_list(box, src=0)
        return box[0]
    return all_gather, bcast_bytes


def run_cet_sublattice(L=None, n_sweeps=2000, temp=None, defect_fraction=0.0, n_seeds=5, impurity_c=0.0,
                       output_prefix="cet_sublattice", metrics_every=None, events_per_sweep=None, p_max=0.1,
                       nu_dep=None, laser=None, thermal_every=_ks.THERMAL_EVERY, device=0, seed=None,
                       checkpoint_every=0, resume_from=None, lattice=None, verbose=True, rank=0, world=1,
                       mask_stream="numpy"):
    """run_kmc's large-lattice sibling.  Same set-up (initialize_lattice + introduce_defects with the
    run's seed), same cadence and CSV; the steps are synchronous-sublattice sweeps (csrc/sweep.cu), so
    `Step` counts sweeps and `Time` is the accumulated sweep interval.

    laser: None -> update_temperature_cet every `thermal_every` sweeps (the reference's driver);
           dict(power=W, speed=voxels per thermal step along axis 1, start=(i0, j0), dt=s,
           beam_radius=m, absorptivity=) -> thermal_solver.update_temperature with the moving source.
    lattice: optional (state, theta, phi, T, atom_type[, defects]) to start from instead of
           initialize_lattice.  Returns (state, atom_type, total_time, theta, phi) like run_kmc.
    world > 1: the lattice is split into z-slabs, one per rank / GPU (launch every rank with the same
           arguments plus its rank; torch.distributed.run provides the rendezvous).  The trajectory and
           the CSV do not depend on the number of slabs.  Rank 0 writes the CSV; every rank returns its
           own planes.  The laser variant and checkpoints are single-GPU.
    mask_stream: "numpy" — the 200-step defect-mask refresh consumes NumPy's global stream in C order
           like defects.py:18 (single GPU only); "philox" — the device's counter-based stream keyed by
           (seed, refresh number, global site), which a slab run always uses (an ordered host stream
           cannot be split over slabs); a single-GPU run with "philox" equals the slab run bit for bit.
    `Time` is the accumulated sweep interval (a physical clock); run_kmc's CSV carries the reference's
    clock, which its 1e-12 s floor makes a step counter (Time = 1e-12 s x steps) — the two
    columns are not comparable and level-3 parity is therefore matched by executed events (tests).
    """
    L = constants.LATTICE_SIZE if L is None else int(L)
    temp = constants.T_SUB if temp is None else temp
    nu_dep = constants.NU_DEP if nu_dep is None else nu_dep
    seed = constants.RANDOM_SEED if seed is None else int(seed)
    every = constants.METRIC_UPDATE_STEP if metrics_every is None else int(metrics_every)
    np.random.seed(seed)                                    # a resumed run restores the stream position below
    output_dir = f"outputs/{output_prefix}"
    os.makedirs(output_dir, exist_ok=True)
    consts = _gr(L, temp, nu_dep)

    if world > 1 and (laser or checkpoint_every or resume_from):
        raise ValueError("run_cet_sublattice: the laser variant and checkpoint / resume are single-GPU")
    if mask_stream not in ("numpy", "philox"):
        raise ValueError("mask_stream must be 'numpy' or 'philox'")
    philox_mask = world > 1 or mask_stream == "philox"
    all_gather = None
    i_begin, i_end = _ks.slab_bounds(L, world, rank)
    ctx = _lib.Context(L=L, device=device, i_begin=i_begin, i_end=i_end, halo=_ks.SWEEP_HALO if world > 1 else 0)
    try:
        ctx.set_rate_params(rate_params(impurity_c, 1, 2, 3, overrides={"NU_DEP": nu_dep}))
        sweep0, total_time, nucleation_count, cet_detected, rows = 0, 0.0, 0, False, []
        if resume_from:
            meta = resume(ctx, resume_from)
            sweep0, total_time = int(meta.get("sweep", 0)), float(meta.get("time", 0.0))
            nucleation_count, cet_detected = int(meta.get("nucleation_count", 0)), bool(meta.get("cet_detected", False))
            rows = list(meta.get("rows", []))
        else:
            if lattice is None:
                state, theta, phi, T, atom_type = _ks.initialize_lattice(lattice_size=L, n_seeds=n_seeds, T_sub=temp,
                                                                         impurity_c=impurity_c)
                defects_mask, _ = _ks.introduce_defects(state, atom_type, T, apply_to_state=False)
            else:
                state, theta, phi, T, atom_type = lattice[:5]
                defects_mask = lattice[5] if len(lattice) > 5 else None
            own = slice(i_begin, i_end)
            ctx.upload(state=np.ascontiguousarray(state[own], dtype=np.int64), theta=theta[own], phi=phi[own], T=T[own],
                       defects=None if defects_mask is None else defects_mask[own])
        if world > 1:
            all_gather, bcast_bytes = _slab_comm(world)
            ctx.comm_init(bcast_bytes(_lib.comm_unique_id() if rank == 0 else None), rank, world)
            ctx.halo_exchange(7)
        sp = _lib.SweepParams()
        sp.seed = seed
        sp.events_per_sweep = float(events_per_sweep if events_per_sweep is not None else 0.005 * L ** 3)
        sp.p_max, sp.defect_fraction = float(p_max), float(defect_fraction)
        sp.thermal_every = 0 if laser else int(thermal_every)
        tp = None if laser else thermal_params(_ks.THERMAL_DT, nan_to_num=True, t_floor=temp)
        if laser and int(thermal_every) <= 0:
            raise ValueError("run_cet_sublattice: the laser variant needs thermal_every >= 1")
        if laser:
            l_dt = float(laser.get("dt", _ks.THERMAL_DT))
            l_pos = [float(x) for x in laser.get("start", (0.0, 0.0))]
            tfp = thermal_full_params(l_dt)

        def advance(n):
            """n sweeps; with a laser the thermal step is driven from here every `thermal_every` sweeps."""
            nonlocal total_time, nucleation_count
            done, terminated = 0, False
            while done < n and not terminated:
                blk = n - done if not laser else min(n - done, thermal_every - ((sweep + done) % thermal_every))
                if laser and (sweep + done) % thermal_every == 0:
                    ctx.thermal_full(tfp, laser_source_top(L, tuple(l_pos), float(laser["power"]),
                                                           float(laser.get("beam_radius", DEFAULT_BEAM_RADIUS)),
                                                           float(laser.get("absorptivity", DEFAULT_ABSORPTIVITY))))
                    ctx.snapshot_state()                     # prev_state of the next latent-heat term
                    l_pos[1] += float(laser.get("speed", 0.0))
                res = ctx.sweep_run(blk, sp, tp)
                total_time += res["time"]
                nucleation_count += res["nucleation_count"] if world == 1 else int(np.sum(all_gather(res["nucleation_count"])))
                done += res["sweeps_done"]
                terminated = bool(res["terminated"]) or res["sweeps_done"] == 0
                if res["overflow"]:
                    raise RuntimeError("fired-event list overflowed: lower events_per_sweep")
            return done, terminated

        sweep = sweep0
        if laser:
            ctx.snapshot_state()
        terminated = False
        while sweep < n_sweeps and not terminated:
            stop_at = min(((sweep + every - 1) // every) * every, n_sweeps - 1)      # run_kmc's cadence
            done, terminated = advance(stop_at - sweep + 1)
            sweep += done
            last = sweep - 1
            if last % every == 0:                                                     # kmc_simulation.py:335-338
                if not philox_mask:
                    _defects.refresh_resident(ctx)
                else:
                    ctx.defects_refresh(draws=None, seed=seed, epoch=last // every, prob_base=constants.DEFECT_PROB_BASE,
                                        e_mig=_defects.E_MIGRATION, kT=constants.K_T, T_default=constants.T_SUB,
                                        carbon_id=_defects.CARBON_ID, defect_id=constants.DEFECT_ID)
                    if world > 1:
                        ctx.halo_exchange(1)                                          # the ghost copies of the mask nibble
            row, cet_detected = _ks._metrics_row(last, total_time, nucleation_count, cet_detected, consts, ctx,
                                                 verbose=verbose and rank == 0, all_gather=all_gather)
            rows.append(row)
            if checkpoint_every and (len(rows) % checkpoint_every == 0):
                checkpoint(ctx, os.path.join(output_dir, "checkpoint"), sweep=sweep, time=total_time,
                           nucleation_count=nucleation_count, cet_detected=cet_detected,
                           rows=[{k: (v if not isinstance(v, (np.generic,)) else v.item()) for k, v in r.items()} for r in rows])
        if rank == 0:
            _ks._write_csv(rows, output_dir)
        f = ctx.download(state=True, atom_type=True, theta=True, phi=True)
    finally:
        ctx.close()
    return f["state"], f["atom_type"], total_time, f["theta"], f["phi"]


def run_gr_sweep(temps, nu_deps, L=64, n_sweeps=400, impurity_c=0.0, n_seeds=20, defect_fraction=0.0,
                 output_root="gr_sweep", rank=None, world=None, devices=None, **kw):
    """one independent lattice per (T_sub, NU_DEP) grid point — T_sub sets the
    thermal gradient G = (T_MELT - T_sub) / (L dx), NU_DEP the growth velocity R (kmc_simulation.py:
    236-239).  Replicas only: case q runs on rank q % world (torchrun: RANK / WORLD_SIZE / LOCAL_RANK)
    or, in a single process, on device q % len(devices).  Every case writes its own metrics.csv; the
    rank that owns a case appends its final row to outputs/<output_root>/cet_map_rank<r>.csv, and
    merge_cet_map() joins them into cet_map.csv (columns G, R, G_over_R, AspectRatio,
    EquiaxedFraction, GrainCount, CET_Class, ...)."""
    rank = int(os.environ.get("RANK", 0)) if rank is None else rank
    world = int(os.environ.get("WORLD_SIZE", 1)) if world is None else world
    local = int(os.environ.get("LOCAL_RANK", 0))
    devices = [local] if devices is None else list(devices)
    cases = [(float(t), float(r)) for t in temps for r in nu_deps]
    out_dir = f"outputs/{output_root}"
    os.makedirs(out_dir, exist_ok=True)
    rows = []
    for q, (temp, nu_dep) in enumerate(cases):
        if q % world != rank:
            continue
        prefix = f"{output_root}/T{temp:g}_R{nu_dep:g}"
        run_cet_sublattice(L=L, n_sweeps=n_sweeps, temp=temp, nu_dep=nu_dep, impurity_c=impurity_c, n_seeds=n_seeds,
                           defect_fraction=defect_fraction, output_prefix=prefix,
                           device=devices[(q // world) % len(devices)], verbose=False, **kw)
        import csv
        with open(f"outputs/{prefix}/metrics.csv") as fh:
            last = list(csv.DictReader(fh))[-1]
        G, R, _rp, _gr_phys = _gr(L, temp, nu_dep)
        rows.append({"case": q, "T_sub": temp, "NU_DEP": nu_dep, "G": G, "R": R, "G_over_R": G / R if R > 0 else np.inf,
                     **{k: last[k] for k in ("Step", "Time", "AspectRatio", "EquiaxedFraction", "GrainCount", "AvgGrainSize",
                                             "NucleationCount", "CET_Class", "CET_Detected")}})
    _write_rows(rows, os.path.join(out_dir, f"cet_map_rank{rank}.csv"))
    write_plot_cet_tree([(r["case"], f"outputs/{output_root}/T{r['T_sub']:g}_R{r['NU_DEP']:g}/metrics.csv") for r in rows],
                        os.path.join(out_dir, "plot_cet"))
    return rows


def write_plot_cet_tree(cases, root):
    """Lay per-case metrics files out the way the reference's plot_cet.py discovers them
    (plot_cet.py:26: `outputs/impurity_c_<id>/metrics_<id>.csv` below the working directory; it
    labels a series "<id>% C", reads Step / AspectRatio / DefectDensity / EquiaxedFraction, :49-71):
    `<root>/outputs/impurity_c_<id>/metrics_<id>.csv` for every (id, path of a metrics.csv) in `cases`.
    Run the unmodified script from `<root>`: `cd <root> && python /path/to/plot_cet.py`; for a G-R
    sweep the id is the case number of cet_map.csv (which holds its T_sub, NU_DEP, G and R).  Every
    rank of a sweep adds its own cases."""
    import shutil
    for cid, path in cases:
        d = os.path.join(root, "outputs", f"impurity_c_{int(cid)}")
        os.makedirs(d, exist_ok=True)
        shutil.copyfile(path, os.path.join(d, f"metrics_{int(cid)}.csv"))
    return root


def _write_rows(rows, path):
    import csv
    if not rows:
        return
    with open(path, "w", newline="") as fh:
        w = csv.DictWriter(fh, fieldnames=list(rows[0].keys()))
        w.writeheader()
        w.writerows(rows)


def merge_cet_map(output_root="gr_sweep"):
    """Join the per-rank files of run_gr_sweep into outputs/<output_root>/cet_map.csv (sorted by case)."""
    import csv
    import glob
    rows = []
    for p in sorted(glob.glob(f"outputs/{output_root}/cet_map_rank*.csv")):
        with open(p) as fh:
            rows += list(csv.DictReader(fh))
    rows.sort(key=lambda r: int(r["case"]))
    _write_rows(rows, f"outputs/{output_root}/cet_map.csv")
    return rows
