"""BASELINE configs[4]: the G-R parameter sweep (batched independent lattices, replicas only) over the
GPUs of one box, producing cet_map.csv and the tree the reference's plot_cet.py reads.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
        scripts/run_gr_sweep.py [--L 64] [--sweeps 400] [--out gr_sweep]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

from cetkmc import campaign


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--L", type=int, default=64)
    ap.add_argument("--sweeps", type=int, default=401)
    ap.add_argument("--out", default="gr_sweep")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    temps = [2600.0, 2800.0, 3000.0, 3200.0]                      # T_sub -> G = (T_MELT - T_sub) / (L dx)
    nu_deps = [1e12, 5e12, 2e13, 1e14]                            # NU_DEP -> R
    t0 = time.perf_counter()
    rows = campaign.run_gr_sweep(temps, nu_deps, L=args.L, n_sweeps=args.sweeps, n_seeds=20, defect_fraction=3e-3,
                                 impurity_c=0.1, output_root=args.out, metrics_every=100)
    dt = time.perf_counter() - t0
    # replicas only: the one rendezvous is "everybody is done" before rank 0 merges the per-rank files
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")
        dist.barrier()
    if rank == 0:
        merged = campaign.merge_cet_map(args.out)
        wall = time.perf_counter() - t0
        sites = len(temps) * len(nu_deps) * args.L ** 3 * args.sweeps
        print(json.dumps({"cases": len(merged), "world": world, "L": args.L, "sweeps": args.sweeps, "wall_s": wall,
                          "rank0_s": dt, "site_updates_per_s": sites / wall,
                          "classes": sorted(set(r["CET_Class"] for r in merged))}))
        for r in merged:
            print(f"case {int(r['case']):2d} T_sub {float(r['T_sub']):6.0f} NU_DEP {float(r['NU_DEP']):8.1e} G/R {float(r['G_over_R']):9.3e} "
                  f"AR {float(r['AspectRatio']):.3f} eq {float(r['EquiaxedFraction']):.3f} grains {int(float(r['GrainCount']))} {r['CET_Class']}")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
