#!/usr/bin/env python
"""bench.py — KMC site-updates/s of the sublattice sweep on the 512^3 lattice (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl cetkmc|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One step = one synchronous-sublattice sweep over the whole lattice (csrc/sweep.cu): thermal stencil
+ dense rate rebuild when due (every 20 sweeps, kmc_simulation.py:248), one fire decision per site
against its resident rate sum, event pick + conflict resolution + apply for the fired sites,
neighbour-rate refresh of the sites the events touched, totals for the next time increment, and
(N > 1) the halo exchange.
Workload: the 'half-grown' synthetic lattice of SURVEY §8(d)(ii), 512 x 512 x 512 sites per GPU;
N GPUs hold a (512 N) x 512 x 512 lattice split into z-slabs (weak scaling).

Prints ONE JSON line (rank 0).  Keys beyond the base contract: roofline (dominant kernel, live
CUDA-event timing), cpu_baseline (the oracle port on the host cores, bounded sample), e2e (the
public API call with host buffers, copies inside the timed region), clocks, gpu_launches.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L_BENCH = 512
THERMAL_EVERY = 20
EVENTS_FRACTION = 0.005         # events_per_sweep = 0.5 % of the sites (the level-3 parity tests run at <= this)
P_MAX = 0.1
# algorithmic bytes per unit of each dense kernel
BYTES_STREAM = 8                # resident rate sum read once per site
BYTES_RATES = 41                # 1 B state + 8 B T + 24 B unit vector read, 8 B rate sum written
BYTES_STAMP = 1                 # refresh scan: one stamp byte per site
BYTES_THERMAL = 16              # T read + T written
CPU_SAMPLE_L = 160              # cpu_baseline / reference arm: one 160^3 block of the same workload


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_rate_sample(threads):
    """The oracle port (oracle/oracle.c, OpenMP over planes) evaluating every event rate of a
    CPU_SAMPLE_L^3 block of the benchmark workload — the reference's get_event_rates sweep, which
    is what one 'site-update' costs on the CPU path (kmc_simulation.py:253)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    from cetkmc import _synth
    os.environ["OMP_NUM_THREADS"] = str(threads)
    O.build()
    Ls = CPU_SAMPLE_L
    packed, th, ph, T = _synth.half_grown(Ls, seed=1234)
    st, df = _synth.unpack(packed)
    p = O.make_params(0.1)
    t0 = time.perf_counter()
    reps = 0
    while True:
        O.site_rates(st, th, ph, T, df, Ls, p)
        reps += 1
        dt = time.perf_counter() - t0
        if dt > 10.0 or reps >= 400:                 # ~10 s of host work (bounded sample)
            break
    return Ls ** 3 * reps / dt, dt, reps


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU algorithm for this path (oracle port; the Python
    reference cannot travel to the GPU box) on all host threads, bounded sample per step."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    from cetkmc import _synth
    os.environ["OMP_NUM_THREADS"] = str(threads)
    O.build()
    Ls = CPU_SAMPLE_L
    packed, th, ph, T = _synth.half_grown(Ls, seed=1234)
    st, df = _synth.unpack(packed)
    p = O.make_params(0.1)
    for _ in range(max(args.warmup, 1)):
        O.site_rates(st, th, ph, T, df, Ls, p)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.site_rates(st, th, ph, T, df, Ls, p)        # one full rate sweep = one visit of every site
    dt = time.perf_counter() - t0
    value = Ls ** 3 * args.steps / dt
    sample = f"{Ls}^3 block of the half-grown workload, one full event-rate sweep per step (oracle.c, OpenMP)"
    print(json.dumps({
        "impl": "reference", "metric": "kmc_site_updates_per_s", "value": value, "unit": "site-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"half-grown {L_BENCH}^3 per GPU (reference arm: {sample})"},
        "cpu_baseline": {"value": value, "unit": "site-updates/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "site-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def pinned(shape, dtype):
    """Page-locked host array (torch allocator) — falls back to pageable memory."""
    try:
        import torch
        t = torch.empty(tuple(shape), dtype={np.float64: torch.float64, np.uint8: torch.uint8}[dtype], pin_memory=True)
        return t.numpy(), t
    except Exception:
        return np.empty(shape, dtype=dtype), None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cetkmc", choices=["cetkmc", "reference"])
    ap.add_argument("--L", type=int, default=L_BENCH, help="sites per edge in a plane and planes per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--debug-flags", type=int, default=0, help="refresh variant for A/B runs (cet_debug_flags)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    warm = max(args.warmup, 3)

    import cetkmc
    from cetkmc import _synth
    from cetkmc._config import rate_params, thermal_params
    cetkmc._lib.require_gpu()                       # no CPU fallback
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    L = args.L
    n0 = L * world
    i_begin, i_end = rank * L, (rank + 1) * L
    halo = 6 if world > 1 else 0
    packed, th, ph, T = _synth.half_grown(L, seed=1234, planes=(i_begin, i_end), n0=n0)
    sites_local, sites_total = L ** 3, L ** 3 * world

    ctx = cetkmc.Context(L=L, n0=n0, device=local_rank, i_begin=i_begin, i_end=i_end, halo=halo)
    ctx.debug_flags(args.debug_flags)
    ctx.set_rate_params(rate_params(0.1))
    ctx.upload_packed(packed)
    ctx.upload(theta=th, phi=ph, T=T)
    if world > 1:
        import torch
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(cetkmc._lib.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        ctx.comm_init(bytes(uid.cpu().numpy().tobytes()), rank, world)
        ctx.halo_exchange(7)
    sp = cetkmc._lib.SweepParams()
    sp.seed, sp.events_per_sweep, sp.p_max = 42, EVENTS_FRACTION * sites_total, P_MAX
    sp.defect_fraction, sp.thermal_every = 3e-3, THERMAL_EVERY
    tp = thermal_params(1e-6, nan_to_num=True)

    def barrier():
        ctx.sync()
        if dist is not None:
            dist.barrier()

    ctx.sweep_run(warm, sp, tp)                     # warm-up (sweep 0 only measures the total rate)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    ctx.profile_enable(True)
    ctx.timer_begin()
    res = ctx.sweep_run(args.steps, sp, tp)
    ms = ctx.timer_end_ms()
    clocks = sampler.stop()
    ctx.profile_enable(False)
    barrier()
    if dist is not None:
        import torch
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        ev = torch.tensor([res["events_applied"]], dtype=torch.float64, device="cuda")
        dist.all_reduce(ev)
        events = float(ev.item())
    else:
        events = float(res["events_applied"])
    KINDS = ("decide", "pick", "apply", "refresh", "thermal", "rates", "halo", "allreduce", "boundary", "step")
    prof = {k: ctx.profile_read(k) for k in KINDS}
    value = sites_total * args.steps / (ms * 1e-3)
    peak, peak_kind = peaks()
    # planes a dense kernel processes: the owned planes plus the ghost planes it must evaluate (N > 1)
    eval_planes = (i_end - i_begin) + (0 if world == 1 else (4 if rank in (0, world - 1) else 8))
    eval_sites = eval_planes * L * L

    def roof(kind, nbytes):
        t_ms, n = prof[kind]
        if n == 0 or t_ms <= 0:
            return None
        ach = nbytes / (t_ms / n * 1e-3) / 1e9
        return {"achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "launches": int(n),
                "ms_per_launch": t_ms / n, "bytes_per_launch": nbytes}

    def roof_rates():
        # full rebuilds (one per thermal step) plus, for N > 1, the 6-plane ghost-zone rebuilds of every sweep
        t_ms, n = prof["rates"]
        n_full = prof["thermal"][1]
        if n == 0 or t_ms <= 0:
            return None
        faces = 0 if world == 1 else (1 if rank in (0, world - 1) else 2)
        nbytes = BYTES_RATES * (n_full * eval_sites + args.steps * faces * 6 * L * L)
        ach = nbytes / (t_ms * 1e-3) / 1e9
        return {"achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "launches": int(n),
                "ms_total": t_ms, "bytes_total": nbytes}

    refreshed = res["sites_refreshed"] / max(args.steps, 1)
    rl = {
        "sweep_stream_kernel": roof("decide", BYTES_STREAM * eval_sites),
        "dirty_scan+dirty_eval (neighbour-rate refresh)": roof("refresh", BYTES_RATES * refreshed + BYTES_STAMP * eval_sites),
        "rates_tile_kernel (dense rebuild after the thermal step)": roof_rates(),
        "thermal_kernel": roof("thermal", BYTES_THERMAL * (i_end - i_begin) * L * L),
    }
    share = {k: prof[k][0] for k in ("decide", "pick", "apply", "refresh", "thermal", "rates", "halo", "allreduce")}
    share["boundary_other"] = max(prof["boundary"][0] - (prof["rates"][0] - 0.0 if world > 1 else 0.0), 0.0) if world > 1 else 0.0
    dominant = max(share, key=share.get)
    dom_name = {"decide": "sweep_stream_kernel", "refresh": "dirty_scan+dirty_eval (neighbour-rate refresh)",
                "rates": "rates_tile_kernel (dense rebuild after the thermal step)", "thermal": "thermal_kernel"}.get(dominant)
    traffic = None                   # dram__bytes_read+write per launch from the committed ncu capture
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            per = json.load(f)["per_kernel"]
        pick_k = {"decide": ["sweep_stream_kernel"], "refresh": ["dirty_scan_kernel", "dirty_eval_kernel"],
                  "rates": ["rates_tile_kernel"], "thermal": ["thermal_kernel_v2"]}.get(dominant, [])
        vals = [v for k, v in per.items() if any(n in k for n in pick_k)]
        traffic = sum(vals[:len(pick_k)]) if vals else None
    except Exception:
        pass
    main_roof = dict(rl[dom_name]) if dom_name and rl.get(dom_name) else {"achieved": None, "peak": peak, "unit": "GB/s", "frac": None}
    main_roof.update({"bound": "hbm", "kernel": dom_name or dominant, "traffic": traffic, "peak_kind": peak_kind,
                      "share_of_step": share[dominant] / max(sum(share.values()), 1e-9)})
    n_rates = prof["rates"][1]
    out = {
        "metric": "kmc_site_updates_per_s", "value": value, "unit": "site-updates/s", "n_gpus": world,
        "steps": args.steps, "warmup": warm, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"half-grown lattice {n0}x{L}x{L} (SURVEY 8d-ii: 25% solid, salt-and-pepper below a "
                               f"wavy front, linear G), {L} planes per GPU, synchronous-sublattice sweeps with resident "
                               f"rates + neighbour-rate refresh, thermal stencil and dense rate rebuild every "
                               f"{THERMAL_EVERY} sweeps, events_per_sweep={EVENTS_FRACTION}N, p_max={P_MAX}",
                   "l2": "fields swept per step (>= 1 GB rate sums per sweep) exceed the 126 MB L2; no flush needed",
                   "parallelism": f"zslab{world}"},
        "executed_events_per_s": events / (ms * 1e-3),
        "kernel_ms_per_step": {k: v / args.steps for k, v in share.items()},
        "step_span_ms": prof["step"][0] / args.steps,
        "roofline": main_roof,
        "roofline_all": rl,
        # per sweep: reset, stream, plane-reduce, pick, apply, dirty-scan, dirty-eval, finalize
        "gpu_launches": int(8 * args.steps + prof["thermal"][1] + n_rates),
        "clocks": clocks,
    }

    # ---- e2e: the public API call with host buffers (copies inside the timed region) -----------
    if not args.no_e2e:
        from cetkmc.kmc_simulation import run_kmc_sublattice_slab
        hp, _k1 = pinned(packed.shape, np.uint8); hp[...] = packed
        hth, _k2 = pinned(th.shape, np.float64); hth[...] = th
        hph, _k3 = pinned(ph.shape, np.float64); hph[...] = ph
        hT, _k4 = pinned(T.shape, np.float64); hT[...] = T
        res_buf = {"packed": pinned(packed.shape, np.uint8), "theta": pinned(th.shape, np.float64),
                   "phi": pinned(ph.shape, np.float64)}
        res_keep = {k: v[1] for k, v in res_buf.items()}            # the torch owners of the pinned pages
        res_out = {k: v[0] for k, v in res_buf.items()}
        for v in res_out.values():
            v[...] = 0                                                # touch the pages outside the timed region
        barrier()
        t0 = time.perf_counter()
        r = run_kmc_sublattice_slab(ctx, hp, hth, hph, hT, args.steps, sp, tp, out=res_out)
        barrier()
        dt = time.perf_counter() - t0
        if dist is not None:
            import torch
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        h2d = hp.nbytes + hth.nbytes + hph.nbytes + hT.nbytes
        d2h = r["packed"].nbytes + r["theta"].nbytes + r["phi"].nbytes
        out["e2e"] = {"value": sites_total * args.steps / dt, "unit": "site-updates/s",
                      "h2d_bytes_per_step": h2d / args.steps, "d2h_bytes_per_step": d2h / args.steps,
                      "what": f"upload packed state+theta+phi+T from pinned host memory, {args.steps} sweeps, "
                              "download packed state+theta+phi into pinned host memory; one call"}
        del res_keep
    ctx.close()

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, dt, reps = cpu_rate_sample(threads)
        out["cpu_baseline"] = {"value": v, "unit": "site-updates/s", "cores": threads, "kind": "port",
                               "sample": f"{reps} full event-rate sweeps of a {CPU_SAMPLE_L}^3 block of the same "
                                         f"workload ({dt:.1f} s, oracle.c with OpenMP over planes)"}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
