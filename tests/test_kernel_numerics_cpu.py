"""CPU: the two numerical claims the raster kernels rest on, restated in numpy float32 (no GPU needed).

1. Fast depth ranking of the forward pass (homan_b200/csrc/raster.cu, "Depth of the winner search"): the kernel ranks
   samples with ws / (w0*iz0 + w1*iz1 + w2*iz2) instead of the reference's 1 / (w0/ws/z0 + w1/ws/z1 + w2/ws/z2) and
   treats two depths within AMB = 48 ulp as a tie to be re-resolved exactly. The two evaluation orders must therefore
   differ by well under AMB / 2 ulp (each key may be off by that much): measured here, plus 2 ulp for __fdividef.
2. Closed-form far part of a sweep item in the backward pass (harmonic_span): sum_{k<n} 1/(z+k) for z >= NEAR_N = 4 from
   the digamma asymptotic series, against a float64 direct sum."""
import numpy as np

f32 = np.float32


def _ulp_diff(a, b):
    ia = a.view(np.int32).astype(np.int64)
    ib = b.view(np.int32).astype(np.int64)
    return np.abs(ia - ib)


def test_fast_depth_is_within_a_few_ulp_of_the_reference_expression():
    rng = np.random.default_rng(0)
    n = 400000
    w = rng.dirichlet((0.6, 0.6, 0.6), size=n).astype(f32)            # clamped barycentrics, incl. near-zero ones
    w[: n // 10] += rng.uniform(0, 2e-3, size=(n // 10, 3)).astype(f32)  # sum not exactly 1
    z = (10.0 ** rng.uniform(-0.9, 1.9, size=(n, 3))).astype(f32)     # depths between the near and far planes
    spread = rng.uniform(0, 1, size=(n, 1)).astype(f32)
    z = (z[:, :1] * (1 + spread * (z / z[:, :1] - 1))).astype(f32)     # mostly similar corner depths, some very different
    ws = ((w[:, 0] + w[:, 1]).astype(f32) + w[:, 2]).astype(f32)
    wn = (w / ws[:, None]).astype(f32)
    s = (((wn[:, 0] / z[:, 0]).astype(f32) + (wn[:, 1] / z[:, 1]).astype(f32)).astype(f32) +
         (wn[:, 2] / z[:, 2]).astype(f32)).astype(f32)
    zp_ref = (f32(1) / s).astype(f32)
    iz = (f32(1) / z).astype(f32)
    q = (((w[:, 0] * iz[:, 0]).astype(f32) + (w[:, 1] * iz[:, 1]).astype(f32)).astype(f32) +
         (w[:, 2] * iz[:, 2]).astype(f32)).astype(f32)
    zp_fast = (ws / q).astype(f32)
    d = _ulp_diff(zp_ref, zp_fast)
    assert d.max() + 2 <= 8, d.max()     # + 2 ulp: __fdividef instead of a correctly rounded division
    assert 2 * (d.max() + 2) < 48        # two keys, each off by that much, still inside the ambiguity window


def _harmonic_span_f32(z1, n):
    """numpy float32 restatement of harmonic_span() in raster.cu."""
    z1, n = f32(z1), f32(n)
    z2 = f32(z1 + n)
    i1, i2 = f32(1) / z1, f32(1) / z2
    a1, a2 = f32(i1 * i1), f32(i2 * i2)
    r = f32(np.log1p(f32(n * i1)))
    r = f32(r + f32(0.5) * f32(i1 - i2))
    r = f32(r + f32(1 / 12) * f32(a1 - a2))
    r = f32(r - f32(1 / 120) * f32(f32(a1 * a1) - f32(a2 * a2)))
    r = f32(r + f32(1 / 252) * f32(f32(f32(a1 * a1) * a1) - f32(f32(a2 * a2) * a2)))
    return r


def test_harmonic_span_matches_the_direct_sum():
    rng = np.random.default_rng(1)
    worst = 0.0
    for _ in range(4000):
        z = float(rng.uniform(4.0, 5.0)) if rng.random() < 0.5 else float(rng.uniform(4.0, 600.0))
        n = int(rng.integers(0, 500))
        direct = float(np.sum(1.0 / (z + np.arange(n, dtype=np.float64)))) if n else 0.0
        got = float(_harmonic_span_f32(z, n))
        worst = max(worst, abs(got - direct))
        assert abs(got - direct) <= 2e-6 * max(direct, 1.0) + 1e-7, (z, n, got, direct)
    assert _harmonic_span_f32(7.3, 0) == 0.0   # an item without a far part contributes exactly nothing
    assert worst < 1e-5
