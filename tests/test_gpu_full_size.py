"""BASELINE.json's BVH configurations at their stated sizes (bench.py's c2, c3, c4) against the CPU oracle:

* first hits (surface id AND t, bit for bit) on the FULL image plane of every configuration, through the oracle's
  pointer tree on all host cores — c2 and c3 with the literal aabb.zig test, c4 (322 k surfaces, where the
  literal test visits ~10^5 nodes per ray) with the interval-carrying test on the same tree plus the literal test on
  a 120x67 plane of the same camera;
* full-depth paths, draw for draw (counter RNG + spec math): scene, camera, aspect ratio and x_limit of the
  configuration on a reduced plane (256x256 / 240x135, 16 spp) — all six counters equal, pixels to f32 association;
* at the full plane and depth, invariance of the device path under what must not matter: which tree is traversed
  (binned SAH or the flattened reference topology) and how a pixel's samples are cut into launches (the multi-GPU split)."""
import os

import numpy as np
import pytest

import bench
from oracle import zro_py
from zraytrace_b200 import _abi as A
from zraytrace_b200 import host
from zraytrace_b200 import lib as Z

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["c2", "c3", "c4"])
def workload(request):
    wl = bench.WORKLOADS[request.param]
    hs = host.HostScene(wl["scene"], variant=wl["variant"], aspect_ratio=wl.get("aspect", 1.0))
    dev = Z.Scene(hs, device=0)
    yield request.param, wl, hs, dev
    dev.close()
    hs.close()


def test_full_plane_first_hits_do_not_depend_on_the_tree(workload):
    name, wl, hs, dev = workload
    p = bench.params_for(wl)
    ids_s, t_s = dev.primary_hits(hs.camera, p)
    ids_r, t_r = dev.primary_hits(hs.camera, bench.params_for(wl, flags=A.ZRT_FLAG_BVH_REFERENCE))
    assert ids_s.shape == (wl["h"], wl["w"]) and (ids_s != A.ZRT_NO_HIT).mean() > 0.3
    assert np.array_equal(ids_s, ids_r) and np.array_equal(t_s.view(np.uint32), t_r.view(np.uint32))
    cores = os.cpu_count() or 1
    # the whole plane through the oracle's pointer tree (scenes.zig:102-128, 206-260 at BASELINE size)
    trav = zro_py.TRAVERSAL_TIGHT if name == "c4" else zro_py.TRAVERSAL_REF
    ids_o, t_o = zro_py.primary_hits(hs, hs.camera, p, traversal=trav, threads=cores)
    assert np.array_equal(ids_o, ids_s), f"{(ids_o != ids_s).sum()} of {ids_o.size} surface ids differ"
    assert np.array_equal(t_o.view(np.uint32), t_s.view(np.uint32)), "hit distances are not bit-identical"
    if name == "c4":  # the literal aabb.zig traversal on the same camera, 1/16 of the resolution per axis
        ps = A.make_params(wl["w"] // 16, wl["h"] // 16, 1, wl["depth"], x_limit=wl.get("x_limit", A.ZRT_XLIMIT_HEIGHT))
        ids_l, t_l = zro_py.primary_hits(hs, hs.camera, ps, traversal=zro_py.TRAVERSAL_REF, threads=cores)
        ids_g, t_g = dev.primary_hits(hs.camera, ps)
        assert (ids_g != A.ZRT_NO_HIT).mean() > 0.3
        assert np.array_equal(ids_l, ids_g) and np.array_equal(t_l.view(np.uint32), t_g.view(np.uint32))


def test_full_depth_paths_match_the_oracle_at_the_configuration_camera(workload):
    """Scene, camera, aspect ratio, x_limit and depth of the BASELINE configuration; the plane reduced so that the oracle
    finishes in seconds.  Counter RNG + spec math: every draw and rounding is the same on both sides."""
    name, wl, hs, dev = workload
    w, h = (240, 135) if name == "c4" else (256, 256)
    p = A.make_params(w, h, 16, wl["depth"], x_limit=wl.get("x_limit", A.ZRT_XLIMIT_HEIGHT), seed=42, sample_chunks=1)
    img_o, c_o, st = zro_py.render(hs, hs.camera, p, rng=zro_py.RNG_CTR, math=zro_py.MATH_SPEC,
                                   traversal=zro_py.TRAVERSAL_TIGHT, threads=os.cpu_count() or 1)
    for flags in (0, A.ZRT_FLAG_KERNEL_THREAD, A.ZRT_FLAG_KERNEL_WARP, A.ZRT_FLAG_KERNEL_POOL, A.ZRT_FLAG_BVH_REFERENCE,
                  A.ZRT_FLAG_KERNEL_POOL | A.ZRT_FLAG_BVH_REFERENCE):
        p.flags = flags
        img_g, c_g, _ = dev.render(hs.camera, p)
        assert c_g.as_dict() == c_o.as_dict(), (name, flags, c_g.as_dict(), c_o.as_dict())
        np.testing.assert_allclose(img_g, img_o, rtol=2e-5, atol=1e-6)
    assert c_o.samples_processed == w * h * 16 and st.triangle_tests > 0
    if name == "c4":
        assert st.texture_lookups > 0  # image-textured metal triangles and the earth-mapped Man are on the plane


def test_full_size_paths_do_not_depend_on_tree_or_sample_split(workload):
    """Full image plane, full depth, 16 of the configuration's samples per pixel: (a) SAH and reference topology trace
    the same paths — all six counters and every pixel bit for bit; (b) two launches over disjoint sample ranges (two
    ranks of the spp split) add up to the same counters exactly and to the same image up to f32 association."""
    name, wl, hs, dev = workload
    spp = 16
    kw = dict(bvh=True, x_limit=wl.get("x_limit", A.ZRT_XLIMIT_HEIGHT), seed=42)
    full, c_full, _ = dev.render(hs.camera, A.make_params(wl["w"], wl["h"], spp, wl["depth"], **kw))
    ref, c_ref, _ = dev.render(hs.camera, A.make_params(wl["w"], wl["h"], spp, wl["depth"], flags=A.ZRT_FLAG_BVH_REFERENCE, **kw))
    assert c_full.as_dict() == c_ref.as_dict() and np.array_equal(full, ref)
    assert c_full.samples_processed == spp * c_full.pixels_processed
    assert c_full.rays_processed == c_full.reflections + c_full.samples_processed - c_full.recursion_depth_hits
    acc, tot = np.zeros_like(full), {}
    for b, e in ((0, 7), (7, 16)):
        part, c, _ = dev.render(hs.camera, A.make_params(wl["w"], wl["h"], spp, wl["depth"], sample_begin=b, sample_end=e,
                                                          flags=A.ZRT_FLAG_RAW_SUM, **kw))
        acc += part
        for k, v in c.as_dict().items():
            tot[k] = tot.get(k, 0) + v
    assert tot == c_full.as_dict()
    np.testing.assert_allclose(acc * np.float32(1.0 / spp), full, rtol=2e-5, atol=1e-6)
