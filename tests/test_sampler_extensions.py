"""Sampler extensions behind flags (the reference's own TODO list src/README.md:5-13):
Russian roulette and Halton pixel jitter.  They change the estimator, so there is nothing in the reference to
compare with; the CPU part checks the oracle's restatement of the spec (include/zrt.h) for what the spec promises
(same expectation, shorter paths, stratified jitter), the GPU part checks device == oracle draw for draw."""
import numpy as np
import pytest

from oracle import zro_py
from tests import scenes_py
from zraytrace_b200 import _abi as A


def _pooled(img, k):
    h, w, _ = img.shape
    return img.reshape(h // k, k, w // k, k, 3).mean(axis=(1, 3))


def test_halton_points_are_the_radical_inverses():
    """Primary hits of sample s with ZRT_FLAG_SAMPLER_HALTON use jitter (h2(s+1), h3(s+1)) + pixel offsets mod 1:
    check the offsets cancel, i.e. differences between samples are the differences of the radical inverses."""
    def h(i, b):
        f, r = 1.0, 0.0
        while i:
            f /= b
            r += f * (i % b)
            i //= b
        return r
    sc, cam = scenes_py.three_balls()
    # a 1x1 image: colour of sample s is a deterministic function of the jitter; use the oracle's own plain jitter
    # path as the reference by brute force over a 64x64 image instead: stratification shows up as lower variance
    p_plain = A.make_params(48, 48, 16, 1, bvh=False)
    p_halt = A.make_params(48, 48, 16, 1, bvh=False, flags=A.ZRT_FLAG_SAMPLER_HALTON)
    ref, _, _ = zro_py.render(sc, cam, A.make_params(48, 48, 1024, 1, bvh=False), threads=4)
    a, ca, _ = zro_py.render(sc, cam, p_plain)
    b, cb, _ = zro_py.render(sc, cam, p_halt)
    assert ca.samples_processed == cb.samples_processed == 48 * 48 * 16
    err_plain = np.sqrt(((a - ref) ** 2).mean())
    err_halton = np.sqrt(((b - ref) ** 2).mean())
    # depth 1: the image is pure pixel-footprint integration, where a (2,3) Halton set beats 16 random points
    assert err_halton < 0.8 * err_plain, (err_halton, err_plain)
    assert abs(h(5, 2) - 0.625) < 1e-12 and abs(h(5, 3) - (2 / 3 + 1 / 9)) < 1e-12


def test_halton_needs_the_counter_rng():
    sc, cam = scenes_py.three_balls()
    with pytest.raises(RuntimeError):
        zro_py.render(sc, cam, A.make_params(8, 8, 2, 5, bvh=False, flags=A.ZRT_FLAG_SAMPLER_HALTON), rng=zro_py.RNG_REF)


def test_russian_roulette_is_unbiased_and_shortens_paths():
    sc, cam = scenes_py.three_balls()
    p0 = A.make_params(64, 64, 256, 30, bvh=False)
    p1 = A.make_params(64, 64, 256, 30, bvh=False, flags=A.ZRT_FLAG_RUSSIAN_ROULETTE)
    a, ca, _ = zro_py.render(sc, cam, p0, threads=4)
    b, cb, _ = zro_py.render(sc, cam, p1, threads=4)
    assert cb.samples_processed == ca.samples_processed
    assert cb.rays_processed < ca.rays_processed            # paths end early
    assert cb.recursion_depth_hits <= ca.recursion_depth_hits
    # same expectation: 8x8-pooled images agree within Monte-Carlo noise (256 spp x 64 pixels per cell)
    d = _pooled(a, 8) - _pooled(b, 8)
    assert np.sqrt((d ** 2).mean()) < 0.01, np.sqrt((d ** 2).mean())
    # paths shorter than 3 rays never meet the roulette: a depth-2 render is bit-identical with and without it
    a2, ca2, _ = zro_py.render(sc, cam, A.make_params(32, 32, 8, 2, bvh=False))
    b2, cb2, _ = zro_py.render(sc, cam, A.make_params(32, 32, 8, 2, bvh=False, flags=A.ZRT_FLAG_RUSSIAN_ROULETTE))
    assert ca2.as_dict() == cb2.as_dict()
    np.testing.assert_allclose(a2, b2, rtol=1e-6, atol=1e-7)


def test_russian_roulette_with_the_reference_stream():
    sc, cam = scenes_py.three_balls()
    a, ca, _ = zro_py.render(sc, cam, A.make_params(48, 48, 64, 30, bvh=False), rng=zro_py.RNG_REF)
    b, cb, _ = zro_py.render(sc, cam, A.make_params(48, 48, 64, 30, bvh=False, flags=A.ZRT_FLAG_RUSSIAN_ROULETTE),
                             rng=zro_py.RNG_REF)
    assert cb.rays_processed < ca.rays_processed
    assert np.sqrt(((_pooled(a, 8) - _pooled(b, 8)) ** 2).mean()) < 0.03


@pytest.mark.gpu
@pytest.mark.parametrize("name,w,spp,depth", [("three_balls", 96, 16, 30), ("teapot", 64, 8, 30), ("bunny_glass", 48, 8, 30)])
@pytest.mark.parametrize("flags", [A.ZRT_FLAG_RUSSIAN_ROULETTE, A.ZRT_FLAG_SAMPLER_HALTON,
                                   A.ZRT_FLAG_RUSSIAN_ROULETTE | A.ZRT_FLAG_SAMPLER_HALTON])
def test_device_matches_oracle_draw_for_draw(name, w, spp, depth, flags):
    from zraytrace_b200 import lib as Z
    builders = {"three_balls": scenes_py.three_balls, "teapot": scenes_py.teapot_and_ball,
                "bunny_glass": lambda: scenes_py.bunny_and_ball(dielectric=True)}
    sc, cam = builders[name]()
    p = A.make_params(w, w, spp, depth, sample_chunks=1, flags=flags)
    img_o, c_o, _ = zro_py.render(sc, cam, p)
    with Z.Scene(sc, device=0) as dev:
        img_g, c_g, _ = dev.render(cam, p)
        assert c_g.as_dict() == c_o.as_dict()
        np.testing.assert_allclose(img_g, img_o, rtol=3e-5, atol=1e-6)
        # the split-sample contract holds for the extensions too (keys are global sample indices)
        tot = {}
        acc = np.zeros_like(img_g)
        for b, e in ((0, spp // 2), (spp // 2, spp)):
            q = A.make_params(w, w, spp, depth, sample_chunks=1, flags=flags | A.ZRT_FLAG_RAW_SUM, sample_begin=b, sample_end=e)
            part, c, _ = dev.render(cam, q)
            acc += part
            for k, v in c.as_dict().items():
                tot[k] = tot.get(k, 0) + v
        assert tot == c_g.as_dict()
        np.testing.assert_allclose(acc * np.float32(1.0 / spp), img_g, rtol=1e-5, atol=1e-6)
        with pytest.raises(Z.ZrtError):
            dev.render(cam, A.make_params(w, w, spp, depth, flags=flags | A.ZRT_FLAG_KERNEL_WARP))
