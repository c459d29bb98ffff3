"""CPU: the product's per-site rate code (csrc/site_rates.cuh), compiled for the host by g++ as
a test-only shim, against the oracle / reference fixtures.  Checks event set, order and rates of
the device arithmetic before any GPU time is spent (the same header is what the kernels inline)."""
import ctypes as C
import glob
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, golden

SRC = os.path.join(ROOT, "tests", "hostsim", "site_rates_host.cpp")
OUT = os.path.join(ROOT, "tests", "hostsim", "_build")


@pytest.fixture(scope="module")
def hostsim():
    os.makedirs(OUT, exist_ok=True)
    so = os.path.join(OUT, "libhostsim.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", so, SRC],
                   check=True)
    return C.CDLL(so)


@pytest.mark.parametrize("name", sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "rates_*.npz"))))
def test_site_events_match_reference(hostsim, name):
    import cetkmc
    from cetkmc._config import rate_params
    g = golden(name)
    L = g["state"].shape[0]
    P = rate_params(float(g["impurity_c"]))
    vox = (g["state"].astype(np.uint8) | (g["defects"].astype(np.uint8) << 4)).ravel()
    cap = 16 * L ** 3
    ty = np.zeros(cap, np.uint8); pos = np.zeros(cap, np.int64); rate = np.zeros(cap); tgt = np.zeros(cap, np.int64)
    atom = np.zeros(cap, np.int32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    hostsim.hostsim_events.restype = C.c_longlong
    th, ph, T = (np.ascontiguousarray(g[k], dtype=np.float64) for k in ("theta", "phi", "T"))
    n = hostsim.hostsim_events(vp(vox), vp(th), vp(ph), vp(T), L, C.byref(P), C.c_longlong(cap), vp(ty), vp(pos),
                               vp(rate), vp(tgt), vp(atom))
    assert n == g["ev_type"].size
    np.testing.assert_array_equal(ty[:n], g["ev_type"])
    np.testing.assert_array_equal(pos[:n], g["ev_pos"])
    np.testing.assert_array_equal(tgt[:n], g["ev_target"])
    nondep = g["ev_type"] != 0                      # species of dep events come from the draw stream
    np.testing.assert_array_equal(atom[:n][nondep], g["ev_atom"][nondep])
    np.testing.assert_allclose(rate[:n], g["ev_rate"], rtol=1e-13, atol=0.0)


def test_neighbour_class_words(hostsim):
    """The cache word the dense kernel decodes with shifts and ANDs (rate_tile.cuh): nibble o of a
    site = class of neighbour o — bit 0 occupied, bit 1 Re, bit 2 C, bit 3 attachable species, 8
    alone = outside the lattice — and bits 56 / 57 flag k == 0 / k == L-1."""
    from cetkmc._config import rate_params
    L = 7
    rng = np.random.default_rng(4)
    st = rng.integers(0, 6, (L, L, L)).astype(np.uint8)            # 0 empty, 1 W, 2 Re, 3 C, 4 defect, 5 unknown species
    vox = (st | (rng.integers(0, 2, (L, L, L)).astype(np.uint8) << 4)).ravel()
    P = rate_params(0.1)
    out = np.zeros(L ** 3, np.uint64)
    hostsim.hostsim_nst_words(vox.ctypes.data_as(C.c_void_p), L, C.byref(P), out.ctypes.data_as(C.c_void_p))
    code = {0: 0, 1: 8 | 1, 2: 8 | 2 | 1, 3: 8 | 4 | 1, 4: 1, 5: 1}
    offs = [(1, 1, 0), (1, -1, 0), (-1, 1, 0), (-1, -1, 0), (0, 1, 1), (0, 1, -1), (0, -1, 1), (0, -1, -1),
            (2, 0, 0), (-2, 0, 0), (0, 2, 0), (0, -2, 0), (0, 0, 2), (0, 0, -2)]       # kmc_event_rates.py:29-36
    for i in range(L):
        for j in range(L):
            for k in range(L):
                w = (1 << 56 if k == 0 else 0) | (1 << 57 if k == L - 1 else 0)
                for o, (di, dj, dk) in enumerate(offs):
                    ni, nj, nk = i + di, j + dj, k + dk
                    inb = 0 <= ni < L and 0 <= nj < L and 0 <= nk < L
                    w |= (code[int(st[ni, nj, nk])] if inb else 8) << (4 * o)
                assert int(out[(i * L + j) * L + k]) == w, (i, j, k)


@pytest.mark.parametrize("name", sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "rates_*.npz"))))
def test_tile_evaluation_equals_per_event_code(hostsim, name):
    """The fused sweep kernel evaluates a site from class codes + one pair operand per neighbour
    (csrc/tile_state.cuh).  On the reference's lattices (empty sites carry no orientation) that must
    give the bits of the per-event code (site_rate_sum), which the fixtures pin to the reference."""
    from cetkmc._config import rate_params
    g = golden(name)
    L = g["state"].shape[0]
    P = rate_params(float(g["impurity_c"]))
    st = g["state"].astype(np.uint8)
    rng = np.random.default_rng(L)
    st = np.where(rng.random(st.shape) < 0.03, 5, st).astype(np.uint8)       # a few sites of an unknown species
    vox = (st | (g["defects"].astype(np.uint8) << 4)).ravel()
    th, ph, T = (np.ascontiguousarray(g[k], dtype=np.float64).copy() for k in ("theta", "phi", "T"))
    canonical = name != "rates_general9.npz"
    if canonical:
        th[st == 0] = 0.0
        ph[st == 0] = 0.0
    tile = np.zeros(L ** 3); gen = np.zeros(L ** 3)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    hostsim.hostsim_tile_rates.restype = C.c_longlong
    oriented = hostsim.hostsim_tile_rates(vp(vox), vp(th), vp(ph), vp(T), L, C.byref(P), vp(tile), vp(gen))
    if canonical:
        assert oriented == 0
        np.testing.assert_array_equal(tile, gen)
        assert np.count_nonzero(gen) > L ** 3 // 4
    else:
        assert oriented > 0                          # the invariant check sends such lattices to the gather kernels
