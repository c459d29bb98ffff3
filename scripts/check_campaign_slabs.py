"""Multi-GPU check (torch.distributed.run, one rank per GPU): campaign.run_cet_sublattice over z-slabs
writes the same metrics.csv and returns the same lattice as the single-GPU run (mask_stream="philox").

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
        scripts/check_campaign_slabs.py
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

from cetkmc import campaign


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tmp = tempfile.mkdtemp() if rank == 0 else None
    box = [tmp]
    dist.broadcast_object_list(box, src=0)
    os.chdir(box[0])
    kw = dict(L=40, n_sweeps=61, temp=2800, defect_fraction=3e-3, n_seeds=20, impurity_c=0.2, metrics_every=20,
              events_per_sweep=0.004 * 40 ** 3, thermal_every=10, verbose=False, seed=7)
    st, at, t, th, ph = campaign.run_cet_sublattice(output_prefix="slabs", device=local, rank=rank, world=world, **kw)
    parts = [None] * world
    dist.all_gather_object(parts, (st, th, ph))
    ok = True
    if rank == 0:
        import pandas as pd
        st1, at1, t1, th1, ph1 = campaign.run_cet_sublattice(output_prefix="single", device=local, mask_stream="philox", **kw)
        full = [np.concatenate([p[q] for p in parts], axis=0) for q in range(3)]
        same = np.array_equal(full[0], st1) and np.array_equal(full[1], th1) and np.array_equal(full[2], ph1) and t == t1
        a, b = pd.read_csv("outputs/slabs/metrics.csv"), pd.read_csv("outputs/single/metrics.csv")
        csv_same = a.equals(b)
        print(f"lattice + time identical: {same}; metrics.csv identical: {csv_same} ({len(a)} rows, "
              f"{int(a['GrainCount'].iloc[-1])} grains, occupied {int((st1 != 0).sum())})")
        if not csv_same:
            print(a.compare(b))
        ok = bool(same and csv_same and len(a) >= 3 and (st1 != 0).sum() > 200)
        print("CAMPAIGN SLAB CHECK", "PASSED" if ok else "FAILED", f"(world={world})")
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
