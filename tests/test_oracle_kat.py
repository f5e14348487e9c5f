"""Pins the CPU oracle against every known-answer test / golden value the reference holds for the
hot path (reference file:line cited per test).  CPU only."""
import ctypes as C
import os

import numpy as np
import pytest
from PIL import Image

from oracle import zro_py
from oracle.zro_py import fv
from tests import scenes_py
from zraytrace_b200 import _abi as A

L = zro_py.lib()
f32 = np.float32


def out3():
    return (C.c_float * 3)()


def test_ray_at():  # ray.zig:32-39
    o = out3()
    L.zro_ray_at(fv(1, 1, 1), fv(1, 2, 3), C.c_float(2.0), o)
    assert tuple(f32(x) for x in o) == (f32(1.53452253e+00), f32(2.06904506e+00), f32(2.60356736e+00))


def test_triangle_miss():  # triangle.zig:84-96
    t = C.c_float()
    hit = L.zro_triangle_hit(fv(1, 0, 0), fv(0, 1, 0), fv(0, 0, 1), fv(1, 1, 1), fv(1, 1, 1), C.c_float(0.1),
                             C.c_float(10000.0), C.byref(t), None, None, None, None)
    assert hit == 0


def test_triangle_hit():  # triangle.zig:98-118
    t, loc, n, ff, uv = C.c_float(), out3(), out3(), C.c_int(), (C.c_float * 2)()
    hit = L.zro_triangle_hit(fv(10, 5, 1), fv(-10, -10, 1), fv(-10, 10, 1), fv(0, 0, -10), fv(0, 0, 1),
                             C.c_float(0.1), C.c_float(10000.0), C.byref(t), loc, n, C.byref(ff), uv)
    assert hit == 1
    assert tuple(loc) == (0.0, 0.0, 1.0)
    assert tuple(n) == (0.0, 0.0, -1.0)
    assert t.value == 11.0
    assert ff.value == 1


def test_aabb_hit():  # aabb.zig:244-254
    mn, mx = fv(-1, -1, -1), fv(1, 1, 1)
    assert L.zro_aabb_hit(mn, mx, fv(-10, 0, 0), fv(-1, 0, 0), C.c_float(0.0), C.c_float(1e5)) == 0
    assert L.zro_aabb_hit(mn, mx, fv(-10, 0, 0), fv(1, 0, 0), C.c_float(0.0), C.c_float(1e5)) == 1


def test_aabb_flat_box_never_hit():  # aabb.zig:121 `tmax <= tmin`
    assert L.zro_aabb_hit(fv(-1, -1, 0), fv(1, 1, 0), fv(0, 0, -5), fv(0.1, 0.1, 1), C.c_float(0.001),
                          C.c_float(1e30)) == 0


def test_aabb_constructors():  # aabb.zig:151-232
    mn, mx, mid = out3(), out3(), out3()
    L.zro_aabb_min_max(fv(-1, 2, 3), fv(4, -3, 7), mn, mx, mid)
    assert tuple(mn) == (-1.0, -3.0, 3.0) and tuple(mx) == (4.0, 2.0, 7.0)
    mn2, mx2, mid2 = out3(), out3(), out3()
    L.zro_aabb_min_max(fv(4, -3, 7), fv(-1, 2, 3), mn2, mx2, mid2)
    assert tuple(mn) == tuple(mn2) and tuple(mx) == tuple(mx2) and tuple(mid) == tuple(mid2)
    assert tuple(mid) == (1.5, -0.5, 5.0)
    pts = np.array([[1, 2, 3], [-4, 5, -6], [7, -8, 9]], np.float32)
    L.zro_aabb_vertexes(pts.ctypes.data_as(C.POINTER(C.c_float)), 3, mn, mx)
    assert tuple(mn) == (-4.0, -8.0, -6.0) and tuple(mx) == (7.0, 5.0, 9.0)
    L.zro_aabb_union(fv(0, 0, 0), fv(1, 1, 1), fv(-1, 0.5, 0.5), fv(0.5, 3, 0.75), mn, mx)
    assert tuple(mn) == (-1.0, 0.0, 0.0) and tuple(mx) == (1.0, 3.0, 1.0)


def test_aabb_surface_area_and_volume():  # aabb.zig:234-242 (2*sum d^2)
    assert L.zro_aabb_surface_area(fv(0, 0, 0), fv(1, -2, 3)) == 28.0
    assert L.zro_aabb_volume(fv(0, 0, 3), fv(-3.5, 2, 4)) == 7.0


def test_vec3():  # vector.zig:169-255
    assert L.zro_vec3_dot(fv(0, 0, 0), fv(1, 0, 0)) == 0.0
    assert L.zro_vec3_dot(fv(1, 0, 0), fv(1, 0, 0)) == 1.0
    assert L.zro_vec3_dot(fv(1, 0, 0), fv(0, 1, 0)) == 0.0
    o = out3()
    L.zro_vec3_unit(fv(0, 0, 0), o)
    assert all(np.isnan(x) for x in o)
    L.zro_vec3_unit(fv(1, 0, 0), o)
    assert tuple(o) == (1.0, 0.0, 0.0)
    L.zro_vec3_unit(fv(3, -4, 0), o)
    assert tuple(f32(x) for x in o) == (f32(0.6), f32(-0.8), f32(0.0))
    pts = np.array([[1, 2, 4], [-1, -2, -4], [-3, 6, -12]], np.float32)
    L.zro_vec3_center(pts.ctypes.data_as(C.POINTER(C.c_float)), 3, o)
    assert tuple(o) == (-1.0, 2.0, -4.0)
    L.zro_vec3_center(pts.ctypes.data_as(C.POINTER(C.c_float)), 1, o)
    assert tuple(o) == (1.0, 2.0, 4.0)


def test_image_texture_albedo_earthmap():  # texture.zig:90-103 (pins PNG decode, row flip, 1-u, nearest)
    px = scenes_py.read_png_bottom_up("earthmap.png")
    tex = A.Texture(A.ZRT_TEXTURE_IMAGE, 0, 0, 0, px.shape[1], px.shape[0], px.shape[2], px.ctypes.data, 0.0, 0.0)
    o = out3()
    want = {(0.0, 0.0): (9.21568632e-01, 9.37254905e-01, 9.49019610e-01),
            (0.1, 0.1): (9.25490200e-01, 9.45098042e-01, 9.56862747e-01),
            (0.5, 0.5): (0.0e+00, 7.84313771e-03, 2.07843139e-01),
            (1.0, 1.0): (1.0, 1.0, 1.0)}
    for (u, v), rgb in want.items():
        L.zro_texture_albedo(C.byref(tex), C.c_float(u), C.c_float(v), o)
        assert tuple(f32(x) for x in o) == tuple(f32(x) for x in rgb), (u, v)


def test_color_texture_albedo():  # texture.zig:83-88
    tex = A.Texture(A.ZRT_TEXTURE_COLOR, 0.1, 0.2, 0.3, 0, 0, 0, None, 0, 0)
    o = out3()
    L.zro_texture_albedo(C.byref(tex), C.c_float(0.1), C.c_float(0.1), o)
    assert tuple(f32(x) for x in o) == (f32(0.1), f32(0.2), f32(0.3))


@pytest.mark.parametrize("which,want,len_check", [
    (0, (-0.7746, 0.3873, -0.7065), "gt1"),   # sample.zig:70-80 randomVector
    (1, (0.1846, 0.8305, -0.0479), "lt1"),    # sample.zig:82-92 randomVectorInUnitSphere
    (2, (0.2167, 0.9746, -0.0562), "unit"),   # sample.zig:94-105 randomUnitVector_old
    (3, (-0.344, -0.932, 0.113), "unit"),     # sample.zig:107-118 randomUnitVector
])
def test_sampling_goldens_xoroshiro(which, want, len_check):
    """Pins the restated Zig std.rand (Xoroshiro128+ / SplitMix64 / float and boolean bit recipes)."""
    o = out3()
    L.zro_sample(which, 0, o)
    v = np.array(tuple(o))
    assert np.all(np.abs(v - np.array(want)) < 0.01)
    n = np.linalg.norm(v)
    assert {"gt1": n > 1.0, "lt1": n < 1.0, "unit": 0.99 < n < 1.01}[len_check]


def test_bvh_random_spheres_statistical():  # bvh.zig:262-291
    hits = L.zro_bvh_random_test(3127, 2000, 42, zro_py.TRAVERSAL_REF)
    assert 10 < hits < 1500
    assert hits == L.zro_bvh_random_test(3127, 2000, 42, zro_py.TRAVERSAL_TIGHT)


def test_quantize():  # png_image.zig:136-140
    assert L.zro_quantize(0.0) == 0 and L.zro_quantize(1.0) == 255 and L.zro_quantize(2.0) == 255
    assert L.zro_quantize(-1.0) == 0 and L.zro_quantize(0.5) == 127


def test_counter_identities_small_render():
    """raytrace.zig:64-99: rays = bg + absorbed + reflections; samples = bg + absorbed + depth_hits."""
    sc, cam = scenes_py.three_balls()
    p = A.make_params(60, 60, 8, 30)
    for rng in (zro_py.RNG_REF, zro_py.RNG_CTR):
        _, c, st = zro_py.render(sc, cam, p, rng=rng)
        absorbed = st.metal_absorbed
        assert c.rays_processed == c.background_hits + absorbed + c.reflections
        assert c.samples_processed == c.background_hits + absorbed + c.recursion_depth_hits
        assert c.samples_processed == 60 * 60 * 8 and c.pixels_processed == 3600


def test_reference_unit_render_configs_run():
    """raytrace.zig:214-271 + scenes.zig:280-289: the reference's own tiny integration renders."""
    sc, cam = scenes_py.small_test_scene()
    img, c, _ = zro_py.render(sc, cam, A.make_params(20, 20, 5, 5, bvh=False), rng=zro_py.RNG_REF)
    assert np.isfinite(img).all() and c.samples_processed == 2000
    sc, cam = scenes_py.man_and_ball()
    img, c, st = zro_py.render(sc, cam, A.make_params(30, 30, 5, 5, bvh=True), rng=zro_py.RNG_REF)
    assert np.isfinite(img).all() and st.bvh_nodes > 0


@pytest.mark.slow
def test_published_counters_7spheres():
    """README.md:49-61: 1000x1000, 1000 spp, depth 30 -> rays 2144645362, reflections 1144753226,
    background 999892115, samples 1e9.  Ratios per sample must agree within 0.5 % on a 4 M-sample run of
    the same image plane (1000x1000 x 4 spp), in both RNG modes."""
    sc, cam = scenes_py.three_balls()
    p = A.make_params(1000, 1000, 4, 30)
    pub = {"rays": 2144645362 / 1e9, "refl": 1144753226 / 1e9, "bg": 999892115 / 1e9}
    for rng, threads in ((zro_py.RNG_REF, 1), (zro_py.RNG_CTR, 4)):
        _, c, _ = zro_py.render(sc, cam, p, rng=rng, threads=threads)
        n = c.samples_processed
        got = {"rays": c.rays_processed / n, "refl": c.reflections / n, "bg": c.background_hits / n}
        for k in pub:
            assert abs(got[k] / pub[k] - 1) < 0.005, (rng, k, got[k], pub[k])
        # depth-limit terminations are rare (published 1.08e-4 per sample): same order of magnitude
        assert 0.5e-4 < c.recursion_depth_hits / n < 2e-4


@pytest.mark.slow
def test_showcase_image_7spheres():
    """showcase/7-spheres.png (README.md:39): the reference's own 1000x1000x1000spp render.  A 1000x1000
    oracle render at 16 spp, box-filtered 4x4 on both sides to average out Monte-Carlo noise, must agree
    with it: per-channel RMSE < 2 % of full scale (noise floor of 16 spp after 16-pixel pooling)."""
    gold = np.array(Image.open(os.path.join(os.path.dirname(__file__), "golden", "showcase_7spheres_1000.png")))
    gold = gold[::-1].astype(np.float64) / 255.0  # PNG row 0 is the top; image row 0 is the bottom
    sc, cam = scenes_py.three_balls()
    img, _, _ = zro_py.render(sc, cam, A.make_params(1000, 1000, 16, 30), rng=zro_py.RNG_CTR, threads=8)
    q = np.floor(np.clip(255.999 * img.astype(np.float64), 0, 255)) / 255.0

    def pool(a):
        return a.reshape(250, 4, 250, 4, 3).mean(axis=(1, 3))

    rmse = np.sqrt(((pool(q) - pool(gold)) ** 2).mean(axis=(0, 1)))
    assert (rmse < 0.02).all(), rmse


@pytest.mark.parametrize("name", ["teapot", "bunny_glass", "man"])
def test_tight_traversal_returns_the_hits_of_the_literal_aabb_test(name):
    """aabb.zig:109-127 does not carry the interval between axes; the oracle's interval-carrying test on
    the same pointer tree must be result-neutral: same first hits bit for bit, same full-depth paths (every counter
    and every pixel), only the visit counts differ.  The GPU tests at BASELINE sizes lean on this."""
    from tests import scenes_py
    from zraytrace_b200 import _abi as A
    sc, cam = {"teapot": scenes_py.teapot_and_ball, "man": scenes_py.man_and_ball,
               "bunny_glass": lambda: scenes_py.bunny_and_ball(dielectric=True)}[name]()
    p = A.make_params(48, 48, 2, 30)
    ids_r, t_r = zro_py.primary_hits(sc, cam, p, traversal=zro_py.TRAVERSAL_REF, threads=2)
    ids_t, t_t = zro_py.primary_hits(sc, cam, p, traversal=zro_py.TRAVERSAL_TIGHT)
    assert np.array_equal(ids_r, ids_t) and np.array_equal(t_r.view(np.uint32), t_t.view(np.uint32))
    assert (ids_r != A.ZRT_NO_HIT).mean() > 0.2
    p = A.make_params(24, 24, 2, 30)
    for rng, math in ((zro_py.RNG_CTR, zro_py.MATH_SPEC), (zro_py.RNG_REF, zro_py.MATH_LIBM)):
        img_r, c_r, st_r = zro_py.render(sc, cam, p, rng=rng, math=math, traversal=zro_py.TRAVERSAL_REF)
        img_t, c_t, st_t = zro_py.render(sc, cam, p, rng=rng, math=math, traversal=zro_py.TRAVERSAL_TIGHT)
        assert c_r.as_dict() == c_t.as_dict() and np.array_equal(img_r.view(np.uint32), img_t.view(np.uint32))
        assert st_t.box_tests * 5 < st_r.box_tests  # the literal test barely culls
