"""Host-side check of index arithmetic the kernels rely on (restated from popnet_b200/csrc/conv.cuh)."""
import numpy as np


def test_fast_divmod_is_exact():
    """fast_divmod: q = umulhi(n, 0xFFFFFFFF // d + 1); r = n - q*d; one fix-up when r < 0 -- exact for every 32-bit n."""
    rng = np.random.default_rng(0)
    for d in list(range(2, 600)) + [1000, 4097, 65535]:
        m = 0xFFFFFFFF // d + 1
        ns = [int(v) for v in rng.integers(0, 2 ** 31 - 1, 300)] + list(range(0, 1200)) + \
             [2 ** 31 - 1, 2 ** 31 - 2, d * 12345 - 1, d * 12345, d * 12345 + 1]
        for n in ns:
            q = (n * m) >> 32
            r = n - q * d
            if r < 0:
                q -= 1
                r += d
            assert q == n // d and r == n % d, (n, d)


def test_c8p_position_count():
    """P = (2 + N*(H+1)) * (W+1): the bench batch at the three resolutions."""
    P = lambda N, H, W: (2 + N * (H + 1)) * (W + 1)
    assert P(64, 112, 112) == 817442 and P(64, 56, 56) == 208050 and P(64, 28, 28) == 53882
    assert (P(64, 28, 28) + 127) // 128 == 421          # M blocks of the 28x28 layers
