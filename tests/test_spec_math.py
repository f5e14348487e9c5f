"""Bounds the "spec" transcendental kernels (oracle/zro_math.h restates them, the device
implements the same sequence in zrt_math.cuh) against float64 truth, and the counter RNG against basic
uniformity checks.  The reference's own transcendentals (Zig std.math) are not pinned by any reference test;
these bounds show the spec kernels are as good a restatement as libm is."""
import numpy as np

from oracle import zro_py

f32 = np.float32


def _ulp_err(got, true64):
    t32 = true64.astype(f32)
    ulp = np.spacing(np.abs(t32)).astype(np.float64)
    return np.abs(got.astype(np.float64) - true64) / np.maximum(ulp, 1e-45)


def test_sincos_on_0_2pi():  # sample.zig:50-52 evaluates cos/sin(2*pi*r2)
    rng = np.random.default_rng(1)
    x = (f32(6.2831855) * rng.random(2_000_000, dtype=f32)).astype(f32)
    x = np.concatenate([x, np.array([0.0, 6.2831855, 3.1415927, 1.5707964, 4.712389], f32)])
    s, c = zro_py.math_eval(0, x)
    x64 = x.astype(np.float64)
    assert np.abs(s - np.sin(x64)).max() < 1.2e-7 and np.abs(c - np.cos(x64)).max() < 1.2e-7
    big = np.abs(np.sin(x64)) > 0.01
    assert _ulp_err(s[big], np.sin(x64[big])).max() < 2.0
    big = np.abs(np.cos(x64)) > 0.01
    assert _ulp_err(c[big], np.cos(x64[big])).max() < 2.0


def test_acos():  # sphere.zig:47 theta = acos(-n.y)
    rng = np.random.default_rng(2)
    x = (rng.random(2_000_000, dtype=f32) * 2 - 1).astype(f32)
    x = np.concatenate([x, np.array([-1.0, 1.0, 0.0, 0.5, -0.5], f32)])
    a, _ = zro_py.math_eval(1, x)
    assert _ulp_err(a, np.arccos(x.astype(np.float64))).max() < 3.0
    assert np.isnan(zro_py.math_eval(1, np.array([1.5], f32))[0][0])


def test_atan2():  # sphere.zig:48 phi = atan2(-n.z, -n.x) + pi
    rng = np.random.default_rng(3)
    x = rng.standard_normal(2_000_000).astype(f32)
    y = rng.standard_normal(2_000_000).astype(f32)
    a, _ = zro_py.math_eval(2, x, y)
    t = np.arctan2(y.astype(np.float64), x.astype(np.float64))
    assert np.abs(a - t).max() < 3.6e-7  # 1.5 ulp of pi
    ax = np.array([1, -1, 0, 0, 1, -1], f32)
    ay = np.array([0, 0, 1, -1, 1, -1], f32)
    a, _ = zro_py.math_eval(2, ax, ay)
    np.testing.assert_allclose(a, np.arctan2(ay.astype(np.float64), ax.astype(np.float64)), atol=3e-7)


def test_pow5_is_square_and_multiply():  # material.zig:127 pow(f32, 1 - cosine, 5.0)
    rng = np.random.default_rng(4)
    x = rng.random(1_000_000, dtype=f32)
    p, _ = zro_py.math_eval(3, x)
    assert np.array_equal(p, x * ((x * x) * (x * x)))
    assert _ulp_err(p, x.astype(np.float64) ** 5).max() < 3.0


def test_counter_rng_uniformity_and_independence():
    """pcg4d keyed (pixel, sample, bounce, seed) is fed small sequential integers; check the three words the
    path consumes for uniformity (chi-square over 64 bins) and absence of linear correlation between
    neighbouring pixels / samples / bounces."""
    n = 1 << 16
    words = np.zeros((n, 4), np.uint32)
    for i in range(n):
        words[i] = zro_py.rng_ctr(i % 1000, i // 1000, 1, 42)
    u = ((words >> 9).astype(np.float64)) / (1 << 23)
    for k in range(3):
        hist, _ = np.histogram(u[:, k], bins=64, range=(0, 1))
        chi2 = ((hist - n / 64) ** 2 / (n / 64)).sum()
        assert chi2 < 130, (k, chi2)  # 63 dof: P(chi2 > 130) ~ 1e-6
        assert abs(np.corrcoef(u[:-1, k], u[1:, k])[0, 1]) < 0.02
    assert abs(np.corrcoef(u[:, 0], u[:, 1])[0, 1]) < 0.02
    a = np.array([zro_py.rng_ctr(7, 3, b, 42) for b in range(31)])
    assert len({tuple(r) for r in a}) == 31
    assert abs(((a[:, 2] >> 31).mean()) - 0.5) < 0.3
