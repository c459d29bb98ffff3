"""CPU: the division-by-multiplication of the list-driven refresh kernel (csrc/rates_refresh.cu: fast_div with the
host-side magic numbers of rates_refresh_list, m = floor(2^32 / d) + 1): the quotient estimate umulhi(n, m) is the
quotient or one above it for every site index 0 <= n < 2^31 and every divisor L or L^2 the lattice can have, so the
kernel's single correction step is enough."""
import numpy as np


def _fast_div(n, d):
    m = (1 << 32) // d + 1                      # rates_refresh_list: a.mLL / a.mL
    q = (n * m) >> 32                           # __umulhi(n, m)
    r = n - q * d
    neg = r < 0
    return np.where(neg, q - 1, q), np.where(neg, r + d, r)


def test_fast_div_matches_integer_division():
    rng = np.random.default_rng(0)
    for L in list(range(8, 130)) + [160, 192, 255, 256, 257, 384, 500, 511, 512, 513, 768, 1000, 1023, 1024]:
        for d in (L, L * L):
            n = np.concatenate([rng.integers(0, 2 ** 31, 4000, dtype=np.int64),
                                np.array([0, 1, d - 1, d, d + 1, 2 ** 31 - 2, 2 ** 31 - 1], dtype=np.int64)])
            q, r = _fast_div(n, d)
            assert np.array_equal(q, n // d) and np.array_equal(r, n % d), (L, d)
