"""BASELINE.json's BVH configurations at their full image sizes (bench.py's c2, c3, c4), through properties that do
not need the CPU oracle to trace a megapixel: the oracle checks first hits on the whole c2 plane; everything else is
invariance of the device path under the things that must not matter — which tree is traversed (binned SAH or the
flattened reference topology, both carrying the reference's tie-break slots and pruning) and how the samples of a
pixel are cut into launches (the multi-GPU split)."""
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
    if name == "c2":  # 262 144 rays through the oracle's pointer tree: ~15 s on one host core
        ids_o, t_o = zro_py.primary_hits(hs, hs.camera, p)
        assert np.array_equal(ids_o, ids_s) and np.array_equal(t_o.view(np.uint32), t_s.view(np.uint32))


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
