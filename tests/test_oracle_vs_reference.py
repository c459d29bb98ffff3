"""CPU, build container only: the C oracle against the reference imported live from
/root/reference (skipped where the reference is absent, e.g. on the GPU box)."""
import numpy as np
import pytest

import refharness

pytestmark = pytest.mark.skipif(not refharness.available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    return refharness.load()


@pytest.mark.parametrize("L,seed,ups,c", [(7, 1, 0, 0.0), (11, 2, 2, 0.15), (14, 3, 25, 0.2)])
def test_event_rates_live(oracle, ref, L, seed, ups, c):
    st, th, ph, T, df = oracle.half_grown_lattice(L, seed=seed, grain=3)
    for _ in range(ups):
        T = ref["thermal_solver"].update_temperature_cet(T, st)
    ref["nb_seed"](seed)
    want = ref["kmc_event_rates"].get_event_rates(st, th, ph, T, st.copy(), df, L, 1, 2, 3, impurity_c=c)
    draws = np.random.RandomState(seed).random_sample(L * L)
    got = oracle.events_as_tuples(oracle.event_rates(st, th, ph, T, df, L, oracle.make_params(c), draws), L)
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert a[0] == b[0] and tuple(a[1]) == tuple(b[1]) and a[2] == b[2] and tuple(a[3]) == tuple(b[3]) \
            and a[4] == b[4]


def test_thermal_live(oracle, ref):
    rng = np.random.default_rng(5)
    T = 2800 + 900 * rng.random((6, 17, 9))
    Tr = T
    for _ in range(30):
        T = oracle.thermal_cet(T)
        Tr = ref["thermal_solver"].update_temperature_cet(Tr, None)
        np.testing.assert_array_equal(T, Tr)


def test_initialize_lattice_and_defects_live(oracle, ref):
    a = ref["lattice_init"].initialize_lattice(lattice_size=9, n_seeds=7, T_sub=2800, random_seed=3, impurity_c=0.2)
    b = oracle.initialize_lattice(9, n_seeds=7, T_sub=2800, random_seed=3, impurity_c=0.2)
    for x, y in zip(a, b):
        np.testing.assert_array_equal(x, y)


def test_helpers_live(oracle, ref):
    ker = ref["kmc_event_rates"]
    for ijk in [(0, 0, 0), (1, 5, 3), (5, 5, 5), (4, 0, 2)]:
        np.testing.assert_array_equal(oracle.bcc_neighbors(*ijk, 6), ker.get_bcc_neighbors(*ijk, 6))
    rng = np.random.default_rng(1)
    for _ in range(20):
        a = rng.uniform(0, np.pi, 2); b = rng.uniform(0, 2 * np.pi, 2)
        assert oracle.misorientation(a[0], b[0], a[1], b[1]) == ker.compute_misorientation(a[0], b[0], a[1], b[1])


def test_clusters_and_metrics_live(oracle, ref):
    """utils.get_clusters / metrics.compute_metrics (utils.py:69, metrics.py:41) vs the restatement."""
    st, th, ph = oracle.grown_lattice(11, seed=5, grain=4, fill=0.6, jitter=0.25)
    clusters, visited = ref["utils"].get_clusters(st, th, ph, theta_threshold=0.5)
    o_clusters, o_visited = oracle.get_clusters(st, th, ph, 0.5)
    np.testing.assert_array_equal(np.asarray(visited), o_visited)
    assert [sorted(c) for c in clusters] == [sorted(c) for c in o_clusters]
    m, om = ref["metrics"].compute_metrics(st, th, ph), oracle.compute_metrics(st, th, ph)
    for k in ("AspectRatio", "EquiaxedFraction", "NucleationDensity", "AvgGrainSize", "GrainCount", "Grain_d50_um", "Grain_d90_um"):
        assert m[k] == om[k], k


def test_defects_live(oracle, ref):
    rng = np.random.default_rng(2)
    st = rng.choice(np.array([0, 1, 2, 3]), size=(9, 9, 9), p=[.3, .3, .1, .3]).astype(np.int64)
    T = 2600 + 1200 * rng.random(st.shape); T[0, 0, :3] = 0.0
    np.random.seed(8)
    mask, _ = ref["defects"].introduce_defects(st.copy(), st, T)
    rs = np.random.RandomState(8)
    np.testing.assert_array_equal(mask, oracle.track_defects(st, T, rs))
    assert np.random.random() == rs.random_sample()
