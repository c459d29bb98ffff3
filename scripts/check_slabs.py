"""Multi-GPU check (launch with torch.distributed.run, one rank per GPU):
the z-slab decomposed sublattice run must reproduce the single-GPU run of the same global
lattice bit for bit (state, theta, phi, T, event counters), because every decision is keyed by the
global site index and plane sums are combined in a fixed order.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/check_slabs.py [--L 48] [--n0 96] [--sweeps 12]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import cetkmc
from cetkmc import _synth
from cetkmc._config import rate_params, thermal_params
from cetkmc.kmc_simulation import slab_bounds, SWEEP_HALO


def sweep_params(seed, n_sites, eps, p_max, defect_fraction, thermal_every):
    sp = cetkmc._lib.SweepParams()
    sp.seed, sp.events_per_sweep, sp.p_max = seed, eps * n_sites, p_max
    sp.defect_fraction, sp.thermal_every = defect_fraction, thermal_every
    return sp


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--L", type=int, default=48)
    ap.add_argument("--n0", type=int, default=0)
    ap.add_argument("--sweeps", type=int, default=12)
    ap.add_argument("--eps", type=float, default=0.01)
    ap.add_argument("--flags", type=int, default=0, help="cet_debug_flags of the slab run (the single-GPU run uses the default kernels)")
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = args.L
    n0 = args.n0 or L
    i_begin, i_end = slab_bounds(n0, world, rank)
    sp = sweep_params(7, n0 * L * L, args.eps, 0.2, 0.01, 5)
    tp = thermal_params(1e-6, nan_to_num=True)

    packed, th, ph, T = _synth.half_grown(L, seed=99, grain=4, planes=(i_begin, i_end), n0=n0)
    ctx = cetkmc.Context(L=L, n0=n0, device=local, i_begin=i_begin, i_end=i_end, halo=SWEEP_HALO)
    ctx.debug_flags(args.flags)
    ctx.set_rate_params(rate_params(0.1))
    ctx.upload_packed(packed)
    ctx.upload(theta=th, phi=ph, T=T)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(cetkmc._lib.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    ctx.comm_init(uid.cpu().numpy().tobytes(), rank, world)
    ctx.halo_exchange(7)
    res = ctx.sweep_run(args.sweeps, sp, tp)
    mine = dict(packed=ctx.download_packed(), **ctx.download(theta=True, phi=True, T=True))
    ctx.close()

    counters = torch.tensor([res["events_fired"], res["events_applied"], res["nucleation_count"]],
                            dtype=torch.int64, device="cuda")
    dist.all_reduce(counters)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    ok = True
    if rank == 0:
        full = {k: np.concatenate([g[k] for g in gathered], axis=0) for k in mine}
        p1, t1, f1, T1 = _synth.half_grown(L, seed=99, grain=4, n0=n0)
        ref = cetkmc.Context(L=L, n0=n0, device=local)
        ref.set_rate_params(rate_params(0.1))
        ref.upload_packed(p1)
        ref.upload(theta=t1, phi=f1, T=T1)
        r1 = ref.sweep_run(args.sweeps, sp, tp)
        one = dict(packed=ref.download_packed(), **ref.download(theta=True, phi=True, T=True))
        ref.close()
        for k in one:
            same = np.array_equal(one[k], full[k])
            ok &= same
            print(f"{k:7s} identical to the single-GPU run: {same}"
                  + ("" if same else f"  ({int((one[k] != full[k]).sum())} sites differ)"))
        c1 = [r1["events_fired"], r1["events_applied"], r1["nucleation_count"]]
        print("counters (fired, applied, nucleations): slabs", counters.tolist(), "single", c1)
        ok &= counters.tolist() == c1
        ok &= r1["events_applied"] > 0
        for key in ("time", "last_total_rate", "last_max_rate", "last_tau"):
            same = r1[key] == res[key]
            ok &= same
            print(f"{key}: slabs {res[key]!r} single {r1[key]!r} {'==' if same else '!='}")
        print("SLAB CHECK", "PASSED" if ok else "FAILED", f"(world={world}, lattice {n0}x{L}x{L}, {args.sweeps} sweeps, "
              f"{r1['events_applied']} events)")
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
