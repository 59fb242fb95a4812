"""CPU: the property the forward rasteriser's exact row intervals rest on (homan_b200/csrc/raster.cu::clip_edge).

The reference skips a sample when  A < (xp - xk) * eky  with A = (yp - yk) * ekx, every operation rounded to fp32
(oracle/csrc/nmr_raster.c). For a fixed raster row the right-hand side is a monotone function of
the pixel index because each rounding step is monotone, hence the samples that pass form a prefix (eky > 0), a suffix
(eky < 0) or all / none (eky == 0) of the row - whatever the magnitudes involved. Checked here in numpy float32 over
random and adversarial edges, power-of-two and other raster sizes; the three-edge intersection is then an interval."""
import numpy as np
import pytest


def _pass_row(A, xk, eky, S):
    xi = np.arange(S, dtype=np.int32)
    num = (2 * xi + 1 - S).astype(np.float32)
    xp = (num / np.float32(S)).astype(np.float32)
    rhs = ((xp - np.float32(xk)).astype(np.float32) * np.float32(eky)).astype(np.float32)
    with np.errstate(invalid="ignore"):
        return ~(np.float32(A) < rhs)


def _is_interval(mask, kind):
    idx = np.flatnonzero(mask)
    if idx.size == 0 or idx.size == mask.size:
        return True
    contiguous = idx[-1] - idx[0] + 1 == idx.size
    if kind > 0:
        return contiguous and idx[0] == 0            # prefix
    if kind < 0:
        return contiguous and idx[-1] == mask.size - 1  # suffix
    return False                                      # eky == 0: all or none


@pytest.mark.parametrize("S", [512, 256, 192, 96, 1000])
def test_edge_predicate_passes_a_prefix_or_suffix_of_every_row(S):
    rng = np.random.default_rng(S)
    n = 4000
    scales = 10.0 ** rng.uniform(-6, 1, size=n)
    x0 = (rng.uniform(-1.5, 1.5, n)).astype(np.float32)
    y0 = (rng.uniform(-1.5, 1.5, n)).astype(np.float32)
    ex = (rng.normal(size=n) * scales).astype(np.float32)
    ey = (rng.normal(size=n) * scales).astype(np.float32)
    ey[::17] = 0.0                                   # horizontal edges
    ex[::19] = 0.0                                   # vertical edges
    x0[::23] = ((2 * rng.integers(0, S, size=x0[::23].shape) + 1 - S) / S).astype(np.float32)  # on pixel centres
    rows = rng.integers(0, S, size=n)
    for k in range(n):
        yp = np.float32(np.float32(2 * rows[k] + 1 - S) / np.float32(S))
        A = np.float32(np.float32(yp - y0[k]) * ex[k])
        m = _pass_row(A, x0[k], ey[k], S)
        assert _is_interval(m, np.sign(ey[k])), (k, x0[k], y0[k], ex[k], ey[k], rows[k])


def test_triangle_rows_are_intervals_and_match_the_bruteforce_inside_test():
    rng = np.random.default_rng(7)
    S = 256
    for _ in range(300):
        f = rng.uniform(-1.1, 1.1, size=(3, 2)).astype(np.float32)
        if rng.random() < 0.3:   # slivers
            f[2] = (f[0] + (f[1] - f[0]) * np.float32(rng.uniform(0.2, 0.8)) + rng.normal(size=2).astype(np.float32) * 1e-4).astype(np.float32)
        yi = int(rng.integers(0, S))
        yp = np.float32(np.float32(2 * yi + 1 - S) / np.float32(S))
        inside = np.ones(S, bool)
        for a, b in ((0, 1), (1, 2), (2, 0)):
            A = np.float32(np.float32(yp - f[a, 1]) * np.float32(f[b, 0] - f[a, 0]))
            inside &= _pass_row(A, f[a, 0], np.float32(f[b, 1] - f[a, 1]), S)
        idx = np.flatnonzero(inside)
        assert idx.size == 0 or idx[-1] - idx[0] + 1 == idx.size
