"""CPU, world_size 2 over gloo: the host-side slab logic — plane partition, per-plane seeded
synthetic slabs, and the halo plane convention of csrc/comm.cu (a rank sends its first / last H
owned planes and receives into its H ghost planes) — checked against the global lattice."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

H = 6


def _worker(rank, world, port, L, n0, q):
    sys.path.insert(0, ROOT)
    import torch
    from cetkmc import _synth
    from cetkmc.kmc_simulation import SWEEP_HALO, slab_bounds
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert SWEEP_HALO == H
        i0, i1 = slab_bounds(n0, world, rank)
        packed, th, ph, T = _synth.half_grown(L, seed=5, grain=4, planes=(i0, i1), n0=n0)
        # local array with H ghost planes per side, as cet_ctx lays it out
        loc = np.full((i1 - i0 + 2 * H, L, L), -1.0)
        loc[H:H + (i1 - i0)] = th
        lower, upper = rank - 1, rank + 1
        reqs = []
        own = i1 - i0
        if lower >= 0:
            reqs.append(dist.isend(torch.from_numpy(loc[H:2 * H].copy()), lower))
            lo_buf = torch.empty((H, L, L), dtype=torch.float64)
            reqs.append(dist.irecv(lo_buf, lower))
        if upper < world:
            reqs.append(dist.isend(torch.from_numpy(loc[own:own + H].copy()), upper))      # planes [np-2H, np-H)
            hi_buf = torch.empty((H, L, L), dtype=torch.float64)
            reqs.append(dist.irecv(hi_buf, upper))
        for r in reqs:
            r.wait()
        if lower >= 0:
            loc[0:H] = lo_buf.numpy()
        if upper < world:
            loc[own + H:] = hi_buf.numpy()
        # compare with the same planes of the global lattice
        _, th_all, _, _ = _synth.half_grown(L, seed=5, grain=4, n0=n0)
        g0, g1 = max(i0 - H, 0), min(i1 + H, n0)
        got = loc[(g0 - (i0 - H)):(g1 - (i0 - H))]
        ok = np.array_equal(got, th_all[g0:g1])
        # a sum-of-counts all-reduce stands in for the event counters
        t = torch.tensor([float((packed & 15 != 0).sum())], dtype=torch.float64)
        dist.all_reduce(t)
        p_all, _, _, _ = _synth.half_grown(L, seed=5, grain=4, n0=n0)
        ok = ok and int(t.item()) == int((p_all & 15 != 0).sum())
        q.put((rank, bool(ok), (i0, i1)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("L,n0", [(12, 24), (10, 13)])
def test_slab_partition_and_halo_convention_gloo(L, n0):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 200) + n0
    procs = [ctx.Process(target=_worker, args=(r, world, port, L, n0, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    bounds = dict((r, b) for r, _, b in res)
    assert bounds[0][0] == 0 and bounds[0][1] == bounds[1][0] and bounds[1][1] == n0
