"""CPU: the C-ABI library loads and exports every symbol include/cetkmc.h declares; host-side
helpers that need no GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "cetkmc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cet_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import cetkmc
    cetkmc.build()
    lib = C.CDLL(cetkmc._lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in cetkmc.h but not exported"
    assert cetkmc._lib.lib().cet_abi_version() == 1
    # every bound function is declared in the header and vice versa
    bound = set(cetkmc._lib.SIGNATURES) | {"cet_last_error", "cet_abi_version"}
    assert bound == set(names)


def test_struct_sizes_match_header():
    import cetkmc
    L = cetkmc._lib
    assert C.sizeof(L.RateParams) == 8 * 19 + 16
    assert C.sizeof(L.ThermalParams) == 48
    assert C.sizeof(L.ThermalFullParams) == 56
    assert C.sizeof(L.KmcResult) == 72
    assert C.sizeof(L.SweepParams) == 40
    assert C.sizeof(L.SweepResult) == 88


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry fails loudly."""
    import cetkmc
    if cetkmc.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CUDA device|no CPU fallback"):
        cetkmc.Context(L=8)
    from cetkmc import thermal_solver
    with pytest.raises(RuntimeError):
        thermal_solver.update_temperature_cet(np.full((4, 4, 4), 3000.0), None)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "cet-driven-simulation-for-3d-printing-am-kmc-approach_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "liboracle" not in txt and "refharness" not in txt, f


def test_host_helpers_match_oracle(oracle):
    from cetkmc import kmc_event_rates as ker
    from cetkmc.kmc_simulation import slab_bounds
    for ijk in [(0, 0, 0), (1, 5, 3), (5, 5, 5), (4, 0, 2), (2, 2, 2)]:
        np.testing.assert_array_equal(ker.get_bcc_neighbors(*ijk, 6), oracle.bcc_neighbors(*ijk, 6))
    assert ker.get_bcc_neighbors(0, 0, 0, 6).dtype == np.int64
    rng = np.random.default_rng(0)
    for _ in range(50):
        a = rng.uniform(0, np.pi, 2); b = rng.uniform(0, 2 * np.pi, 2)
        assert abs(ker.compute_misorientation(a[0], b[0], a[1], b[1]) - oracle.misorientation(a[0], b[0], a[1], b[1])) < 1e-14
    for L, w in [(512, 8), (30, 4), (7, 3), (1024, 8)]:
        b = [slab_bounds(L, w, r) for r in range(w)]
        assert b[0][0] == 0 and b[-1][1] == L and all(b[r][1] == b[r + 1][0] for r in range(w - 1))
        assert max(e - s for s, e in b) - min(e - s for s, e in b) <= 1


def test_host_lattice_setup_matches_oracle(oracle):
    from cetkmc import _host
    a = _host.initialize_lattice(9, n_seeds=7, T_sub=2800, random_seed=3, impurity_c=0.2)
    b = oracle.initialize_lattice(9, n_seeds=7, T_sub=2800, random_seed=3, impurity_c=0.2)
    for x, y in zip(a, b):
        np.testing.assert_array_equal(x, y)
    st, th, ph, T, at = a
    at = at.copy(); at[st == 0] = 0
    at[2:5, 2:5, 2:5] = 3
    np.random.seed(5)
    m, _ = _host.introduce_defects(st, at, T)
    m2 = oracle.track_defects(at, T, np.random.RandomState(5))
    np.testing.assert_array_equal(m, m2)
