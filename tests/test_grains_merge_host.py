"""CPU: the seam merge of metrics.grains_distributed (grains of a lattice split into z-slabs) against the
whole-lattice clustering, with the slab-local device labelling replaced by the host restatement
(tests/hostref.py) on the same planes — owned + 2 ghost planes per cut face."""
import threading

import numpy as np
import pytest

import hostref


class FakeSlab:
    """What metrics.grains_distributed needs from a Context, computed on the host."""

    def __init__(self, st, th, ph, i_begin, i_end, rank):
        self.n0 = st.shape[0]
        self.i_begin, self.i_end, self.rank = i_begin, i_end, rank
        self.owned_shape = (i_end - i_begin,) + st.shape[1:]
        self.lo, self.hi = max(i_begin - 2, 0), min(i_end + 2, self.n0)
        sl = slice(self.lo, self.hi)
        lab, n = hostref.label_grains(st[sl], th[sl], ph[sl], 0.5)
        plane = st.shape[1] * st.shape[2]
        flat = lab.ravel()
        idx = np.arange(flat.size, dtype=np.int64) + self.lo * plane          # global site index
        occ = flat > 0
        first = np.full(n + 1, np.iinfo(np.int64).max, np.int64)
        np.minimum.at(first, flat[occ], idx[occ])
        self.labels = np.where(lab > 0, first[lab], -1).astype(np.int32)      # root per site, like the device labelling
        own = np.zeros_like(lab, bool)
        own[i_begin - self.lo:i_end - self.lo] = True
        sel = occ & own.ravel()
        size = np.bincount(flat[sel], minlength=n + 1)[1:]
        coords = np.stack(np.unravel_index(np.flatnonzero(sel), lab.shape), axis=1) + np.array([self.lo, 0, 0])
        blo = np.full((n + 1, 3), 0x7fffffff, np.int64); bhi = np.full((n + 1, 3), -1, np.int64)
        np.minimum.at(blo, flat[sel], coords); np.maximum.at(bhi, flat[sel], coords)
        self.loc = dict(root=first[1:], size=size.astype(np.int64), box_lo=blo[1:].astype(np.int32), box_hi=bhi[1:].astype(np.int32))

    def grains_local(self, thr):
        return self.loc

    def grain_label_planes(self, i_lo, i_hi):
        return self.labels[i_lo - self.lo:i_hi - self.lo]


@pytest.mark.parametrize("world,grain,fill", [(2, 8, 0.7), (3, 4, 0.5), (4, 24, 0.9)])
def test_seam_merge_equals_whole_lattice_clustering(oracle, world, grain, fill):
    from cetkmc import metrics as M
    from cetkmc.kmc_simulation import slab_bounds
    L = 24
    st, th, ph = oracle.grown_lattice(L, seed=world, grain=grain, fill=fill, jitter=0.1)
    lab, n = hostref.label_grains(st, th, ph, 0.5)
    sizes, _ = hostref.grain_statistics(lab, n)
    slabs = [FakeSlab(st, th, ph, *slab_bounds(L, world, r), r) for r in range(world)]
    barrier = threading.Barrier(world)
    box, results = [None] * world, [None] * world

    def run(r):
        def all_gather(obj):
            box[r] = obj
            barrier.wait()
            out = list(box)
            barrier.wait()
            return out
        results[r] = M.grains_distributed(slabs[r], all_gather, 0.5)

    ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    g = results[0]
    assert g["n"] == n
    np.testing.assert_array_equal(g["size"], sizes)
    first = np.full(n + 1, L ** 3, np.int64)
    np.minimum.at(first, lab.ravel(), np.arange(L ** 3))
    np.testing.assert_array_equal(g["root"], first[1:])
    for r in range(1, world):
        for k in ("root", "size", "box_lo", "box_hi"):
            np.testing.assert_array_equal(results[r][k], g[k])
    # bounding boxes -> the reference's aspect ratios
    _, ar = hostref.grain_statistics(lab, n)
    np.testing.assert_array_equal(M.grain_aspect_ratios(g), ar)
