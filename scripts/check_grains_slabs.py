"""Multi-GPU check (launch with torch.distributed.run, one rank per GPU): grains of a lattice split
into z-slabs (metrics.grains_distributed) == grains of the same lattice in one context — count,
first voxels, sizes, bounding boxes, and the metrics row derived from them.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 \
        scripts/check_grains_slabs.py [--L 48]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import cetkmc
from cetkmc import metrics as M
from cetkmc._config import rate_params
from cetkmc.kmc_simulation import slab_bounds, SWEEP_HALO


def grown_lattice(L, seed, grain, fill, jitter):
    """Solid blocks of constant orientation (+ jitter), `fill` of the sites occupied — grains that span many planes."""
    rng = np.random.default_rng(seed)
    g = (L + grain - 1) // grain
    tb = rng.uniform(0, np.pi, (g, g, g)); pb = rng.uniform(0, 2 * np.pi, (g, g, g))
    rep = lambda a: np.repeat(np.repeat(np.repeat(a, grain, 0), grain, 1), grain, 2)[:L, :L, :L]
    th = np.clip(rep(tb) + jitter * rng.standard_normal((L, L, L)), 0, np.pi)
    ph = np.clip(rep(pb) + jitter * rng.standard_normal((L, L, L)), 0, 2 * np.pi)
    st = np.where(rng.random((L, L, L)) < fill, rng.choice([1, 2, 3, 4], size=(L, L, L), p=[.85, .1, .04, .01]), 0).astype(np.int64)
    th = np.where(st > 0, th, 0.0); ph = np.where(st > 0, ph, 0.0)
    return st, th, ph


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--L", type=int, default=48)
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = args.L
    ok = True
    for case, (grain, fill, jitter) in enumerate(((12, 0.7, 0.1), (6, 0.45, 0.25), (L, 0.9, 0.02))):
        st, th, ph = grown_lattice(L, 100 + case, grain, fill, jitter)
        i_begin, i_end = slab_bounds(L, world, rank)
        ctx = cetkmc.Context(L=L, device=local, i_begin=i_begin, i_end=i_end, halo=SWEEP_HALO)
        ctx.set_rate_params(rate_params(0.1))
        ctx.upload(state=st[i_begin:i_end], theta=th[i_begin:i_end], phi=ph[i_begin:i_end])
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(cetkmc._lib.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        ctx.comm_init(uid.cpu().numpy().tobytes(), rank, world)
        ctx.halo_exchange(7)

        def all_gather(obj):
            out = [None] * world
            dist.all_gather_object(out, obj)
            return out

        g = M.grains_distributed(ctx, all_gather, 0.5)
        ctx.close()
        if rank == 0:
            one = cetkmc.Context(L=L, device=local)
            one.upload(state=st, theta=th, phi=ph)
            w = one.grains(0.5)
            one.close()
            same = g["n"] == w["n"] and all(np.array_equal(g[k], w[k]) for k in ("root", "size", "box_lo", "box_hi"))
            m1, m2 = M.metrics_from_grains(g, L ** 3), M.metrics_from_grains(w, L ** 3)
            same &= all(m1[k] == m2[k] for k in ("AspectRatio", "EquiaxedFraction", "GrainCount", "AvgGrainSize", "Grain_d50_um"))
            print(f"case {case}: {w['n']} grains (largest {int(w['size'].max())} voxels), slabs == single: {same}")
            ok &= bool(same)
    if rank == 0:
        print("GRAIN SLAB CHECK", "PASSED" if ok else "FAILED", f"(world={world}, L={L})")
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
