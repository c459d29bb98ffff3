"""GPU: the defect-mask refresh on the resident lattice (csrc/defects.cu, cetkmc/defects.py; SURVEY
§8f N3) against the oracle's restatement of defects.py:4-19 fed with the same draw stream.
Bit-exact mask (a site could differ only if a draw fell within an ulp of its probability)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _lattice(L, seed, c_frac=0.3):
    rng = np.random.default_rng(seed)
    st = rng.choice(np.array([0, 1, 2, 3, 4]), size=(L, L, L), p=[.3, .4 - c_frac / 2, .2 - c_frac / 2, c_frac, .1]).astype(np.int64)
    T = 2500 + 1500 * rng.random((L, L, L))
    T[rng.random((L, L, L)) < 0.02] = -5.0           # defects.py:13 falls back to T_SUB
    return st, T


@pytest.mark.parametrize("L,seed", [(7, 1), (24, 2), (50, 3), (64, 4)])
def test_defect_mask_matches_oracle_stream(cet, oracle, L, seed):
    st, T = _lattice(L, seed)
    want = oracle.track_defects(st, T, np.random.RandomState(100 + seed))
    ctx = cet.Context(L=L)
    try:
        ctx.upload(state=st, T=T, defects=(st == 1).astype(np.int64))     # a stale mask on W sites must be erased
        draws = np.random.RandomState(100 + seed).random_sample(int((st == 3).sum()))
        n_c, n_d = ctx.defects_refresh(draws=draws)
        got = ctx.download_packed()
    finally:
        ctx.close()
    assert n_c == int((st == 3).sum()) and n_d == int(want.sum())
    np.testing.assert_array_equal(got >> 4, want)
    np.testing.assert_array_equal(got & 15, st)


def test_defects_dropin_consumes_numpy_stream_like_the_reference(cet, oracle):
    from cetkmc import defects as D
    st, T = _lattice(20, 9)
    np.random.seed(5)
    mask, dens = D.introduce_defects(st.copy(), st, T)
    after = np.random.random()
    rs = np.random.RandomState(5)
    want = oracle.track_defects(st, T, rs)
    np.testing.assert_array_equal(mask, want)
    assert after == rs.random_sample()                       # exactly n_carbon draws were consumed
    assert dens == want.sum() / (want.size * (5e-6) ** 3)
    # no carbon: no draw (defects.py:9)
    np.random.seed(6)
    empty = D.track_defects(np.zeros((5, 5, 5), int), np.ones((5, 5, 5), int), 5, None)
    assert not empty.any() and np.random.random() == np.random.RandomState(6).random_sample()
    # T=None: flat probability DEFECT_PROB_BASE, apply_to_state marks the sites as defects
    np.random.seed(7)
    s2 = st.copy()
    m2, _ = D.introduce_defects(s2, st, None, apply_to_state=True)
    u = np.random.RandomState(7).random_sample(int((st == 3).sum()))
    np.testing.assert_array_equal(m2[st == 3], (u < 0.12).astype(int))
    assert np.all(s2[m2 == 1] == 4)


def test_defects_philox_mode(cet):
    L = 48
    st, T = _lattice(L, 12, c_frac=0.35)
    T = np.full_like(T, 3000.0)
    ctx = cet.Context(L=L)
    try:
        ctx.upload(state=st, T=T)
        a = ctx.defects_refresh(seed=3, epoch=1)
        m1 = ctx.download_packed() >> 4
        b = ctx.defects_refresh(seed=3, epoch=1)
        m1b = ctx.download_packed() >> 4
        ctx.defects_refresh(seed=3, epoch=2)
        m2 = ctx.download_packed() >> 4
    finally:
        ctx.close()
    assert a == b and np.array_equal(m1, m1b)                 # deterministic in (seed, epoch, site)
    assert not np.array_equal(m1, m2)
    assert not m1[st != 3].any()
    p = 0.12 * np.exp(-0.3 / (8.617333262e-5 * 3000.0))
    n_c = int((st == 3).sum())
    assert abs(m1.sum() - p * n_c) < 5 * np.sqrt(p * (1 - p) * n_c)


@pytest.mark.parametrize("name", ["defects_c20.npz", "defects_c35.npz"])
def test_defects_dropin_golden(cet, name):
    """cetkmc.defects.introduce_defects against the reference's own output (tests/golden/defects_*.npz)."""
    from conftest import golden
    from cetkmc import defects as D
    g = golden(name)
    st, T, seed = g["state"].astype(np.int64), g["T"], int(g["seed"])
    np.random.seed(seed)
    mask, density = D.introduce_defects(st.copy(), st, T, apply_to_state=False)
    np.testing.assert_array_equal(mask, g["mask"])
    assert density == float(g["density"]) and np.random.random() == float(g["next_draw"])
    np.random.seed(seed)
    mask_noT, _ = D.introduce_defects(st.copy(), st, None)
    np.testing.assert_array_equal(mask_noT, g["mask_noT"])
