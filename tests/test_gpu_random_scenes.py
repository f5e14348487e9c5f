"""Randomised GPU parity (-m gpu): seeded random scenes — spheres of both radius signs, triangle soups with shared
edges and degenerate members, all three materials, constant and image textures, list mode and both BVH topologies —
rendered by libzrt through the C ABI and by the oracle.  Same bars as tests/test_gpu_parity.py: first-hit surface
ids and t bit-identical, the six u64 counters equal, pixels equal up to re-association of the attenuation product.
The fixed scenes of the reference pin the common cases; these pin the code paths they do not reach (triangles in
list mode next to spheres, image-textured triangles using barycentric (u, v), rays that start inside glass spheres,
coplanar leaves that the reference's flat-box rule hides, exact ties on shared edges)."""
import numpy as np
import pytest

from oracle import zro_py
from zraytrace_b200 import _abi as A
from zraytrace_b200 import lib as Z
from zraytrace_b200.scene import SceneBuilder

pytestmark = pytest.mark.gpu


def random_scene(seed, n_spheres, n_tris, with_ground=True):
    rng = np.random.default_rng(seed)
    b = SceneBuilder()
    tex_img = rng.integers(0, 256, size=(17, 23, 3 + (seed & 1)), dtype=np.uint8)
    mats = [b.lambertian(b.color_texture(*rng.uniform(0.1, 1.0, 3))), b.metal(b.color_texture(*rng.uniform(0.3, 1.0, 3))),
            b.dielectric(float(rng.uniform(1.2, 1.8))), b.lambertian(b.image_texture(tex_img)),
            b.metal(b.image_texture(tex_img, u_offset=float(rng.uniform(0, 0.9)), v_offset=float(rng.uniform(0, 0.9))))]
    if with_ground:
        b.sphere((0.0, -103.0, 5.0), 100.0, mats[0])
    for _ in range(n_spheres):
        c = rng.uniform((-4, -2, 2), (4, 3, 9)).astype(np.float32)
        r = float(rng.uniform(0.3, 1.4)) * (-1.0 if rng.random() < 0.2 else 1.0)
        b.sphere(tuple(c), r, int(rng.choice(mats)))
    # triangle soup made of small fans: neighbours share an edge, so exact ties in t do occur
    for _ in range(max(n_tris // 4, 0)):
        centre = rng.uniform((-4, -2, 2), (4, 3, 9)).astype(np.float32)
        ring = centre + rng.uniform(-1.2, 1.2, size=(5, 3)).astype(np.float32)
        m = int(rng.choice(mats))
        for k in range(4):
            b.triangle(tuple(centre), tuple(ring[k]), tuple(ring[k + 1]), m)
    if n_tris:
        p = rng.uniform(-1, 1, 3).astype(np.float32)
        b.triangle(tuple(p), tuple(p), tuple(p + 1), mats[1])                              # zero area
        q = np.array([0.5, 0.25, 4.0], np.float32)                                          # axis-aligned, flat box
        b.triangle(tuple(q), tuple(q + np.array([1, 0, 0], np.float32)), tuple(q + np.array([0, 1, 0], np.float32)), mats[3])
    cam = zro_py.camera_init(tuple(rng.uniform((-1, -0.5, -8), (1, 1.5, -5))), (0.0, 0.0, 1.0), (0.0, 1.0, 0.0), 45.0, 1.0)
    return b.build(), cam


CASES = [(1, 3, 0), (2, 6, 8), (3, 0, 40), (4, 9, 24), (5, 12, 120), (6, 2, 300), (7, 8, 0), (8, 5, 64)]


@pytest.mark.parametrize("seed,n_spheres,n_tris", CASES)
def test_random_scene_primary_hits_and_paths(seed, n_spheres, n_tris):
    sc, cam = random_scene(seed, n_spheres, n_tris)
    with Z.Scene(sc, device=0) as dev:
        for bvh, flags in ((False, 0), (True, 0), (True, A.ZRT_FLAG_BVH_REFERENCE)):
            p = A.make_params(72, 72, 6, 12, bvh=bvh, sample_chunks=1, flags=flags, seed=1000 + seed)
            for jitter in (0, 1):
                ids_o, t_o = zro_py.primary_hits(sc, cam, p, jitter=jitter)
                ids_g, t_g = dev.primary_hits(cam, p, jitter=jitter)
                assert np.array_equal(ids_g, ids_o), (seed, bvh, flags, jitter, int((ids_g != ids_o).sum()))
                assert np.array_equal(t_g.view(np.uint32), t_o.view(np.uint32))
            img_o, c_o, _ = zro_py.render(sc, cam, p)
            img_g, c_g, _ = dev.render(cam, p)
            assert c_g.as_dict() == c_o.as_dict(), (seed, bvh, flags)
            np.testing.assert_allclose(img_g, img_o, rtol=3e-5, atol=1e-6)
            if bvh and not flags:  # the opt-in kernels trace the same paths
                for kf in [A.ZRT_FLAG_KERNEL_WARP] + ([A.ZRT_FLAG_KERNEL_SORTED] if Z.has_experiments() else []):
                    p2 = A.make_params(72, 72, 6, 12, bvh=bvh, sample_chunks=1, flags=kf, seed=1000 + seed)
                    img_k, c_k, _ = dev.render(cam, p2)
                    assert c_k.as_dict() == c_o.as_dict()
                    assert np.array_equal(img_k.view(np.uint32), img_g.view(np.uint32))


def test_camera_inside_a_glass_sphere_and_no_ground():
    sc, cam = random_scene(11, 4, 16, with_ground=False)
    b = SceneBuilder()
    glass = b.dielectric(1.5)
    red = b.lambertian(b.color_texture(0.9, 0.2, 0.2))
    b.sphere((0.0, 0.0, -6.0), 3.0, glass)  # the camera below sits inside this one
    b.sphere((0.0, 0.0, 4.0), 1.0, red)
    b.sphere((2.5, 0.5, 5.0), -1.0, glass)
    inside = b.build()
    cam_in = zro_py.camera_init((0.0, 0.0, -6.5), (0.0, 0.0, 1.0), (0.0, 1.0, 0.0), 45.0, 1.0)
    for scene, camera in ((sc, cam), (inside, cam_in)):
        p = A.make_params(64, 64, 8, 30, sample_chunks=1)
        img_o, c_o, _ = zro_py.render(scene, camera, p)
        with Z.Scene(scene, device=0) as dev:
            img_g, c_g, _ = dev.render(camera, p)
            ids_o, t_o = zro_py.primary_hits(scene, camera, p)
            ids_g, t_g = dev.primary_hits(camera, p)
        assert c_g.as_dict() == c_o.as_dict()
        assert np.array_equal(ids_g, ids_o) and np.array_equal(t_g.view(np.uint32), t_o.view(np.uint32))
        np.testing.assert_allclose(img_g, img_o, rtol=3e-5, atol=1e-6)
