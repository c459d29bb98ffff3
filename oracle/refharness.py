"""Import harness for the UNMODIFIED reference (test infrastructure, never product).

Only usable where /root/reference exists (the build container).  It is used by
`oracle/gen_golden.py` to generate the committed fixtures under `tests/golden/`
and by `tests/test_oracle_vs_reference.py` to pin the C restatement
(`oracle/oracle.c`) to the reference itself.  Nothing on the GPU box imports it.

The reference needs matplotlib (absent here): `lattice_init.py:2-3` imports it at
module level, so empty stub modules are registered first.  Numba keeps a private
MT19937 (`kmc_event_rates.py:65`) that `np.random.seed` does not reach; `nb_seed`
seeds it from inside a jitted function so that the stream equals
`np.random.RandomState(seed).random_sample()`.
"""
import os
import sys
import types

REF_DIR = os.environ.get("CETKMC_REFERENCE_DIR", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "kmc_event_rates.py"))


_loaded = {}


def load():
    """Return a dict of the reference modules (imported once)."""
    if _loaded:
        return _loaded
    if not available():
        raise RuntimeError(f"reference not found under {REF_DIR}")
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/cetkmc_numba_cache")
    for n in ("matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.mplot3d"):
        if n not in sys.modules:
            sys.modules[n] = types.ModuleType(n)
    sys.modules["mpl_toolkits.mplot3d"].Axes3D = object
    # The reference modules import each other by bare name and Numba's on-disk cache
    # re-imports them by that name, so they stay registered in sys.modules.  The drop-in
    # modules are always imported through the package (`<pkg>.kmc_event_rates`), never by
    # bare name in the same process, so the two do not collide.
    saved_path = list(sys.path)
    for k in _REF_NAMES:
        if k in sys.modules and not getattr(sys.modules[k], "__file__", "").startswith(REF_DIR):
            raise RuntimeError(f"module {k!r} already imported from elsewhere")
    sys.path.insert(0, REF_DIR)
    try:
        import numba
        import numpy as np

        @numba.njit
        def nb_seed(s):
            np.random.seed(s)

        import constants, kmc_event_rates, thermal_solver, defects, lattice_init  # noqa
        import utils, metrics, kmc_simulation  # noqa
        for k in _REF_NAMES:
            if k in sys.modules:
                _loaded[k] = sys.modules[k]
        _loaded["nb_seed"] = nb_seed
    finally:
        sys.path[:] = saved_path
    return _loaded


_REF_NAMES = ("constants", "kmc_event_rates", "thermal_solver", "defects", "lattice_init",
              "utils", "metrics", "kmc_simulation", "visualization", "graphs")
