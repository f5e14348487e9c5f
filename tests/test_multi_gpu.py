"""Multi-GPU behind the C ABI (include/zrt.h "multi-GPU", zrt_multi_*): the replacement for raytrace.render()
(raytrace.zig:136-138) on W devices - spp split, ONE NCCL reduce of the f32 accumulators, 1/spp on the root.

CPU part: the group refuses to exist without a device, NCCL loads, the id has the documented size.
GPU part (-m gpu): a group of one device is zrt_render bit for bit; on a box with W >= 2 devices the W-GPU image equals the
1-GPU image to f32 association, all six counters exactly, for W = 2 .. all visible devices, including the README
headline plane against the published counters and showcase image, and the fewer-samples-than-ranks corner."""
import os

import numpy as np
import pytest
from PIL import Image

from tests import scenes_py
from zraytrace_b200 import _abi as A
from zraytrace_b200 import lib as Z


def test_group_needs_a_device_and_nccl_loads():
    assert len(Z.comm_id()) == A.ZRT_COMM_ID_BYTES == 128
    assert Z.nccl_version() >= 20000
    if Z.device_count() > 0:
        pytest.skip("a device is visible")
    sc, cam = scenes_py.three_balls()
    for kw in (dict(devices=[0]), dict(devices=[0, 1]), dict(device=0, comm_id=Z.comm_id(), rank=0, world=2)):
        with pytest.raises(Z.ZrtError) as e:
            Z.MultiScene(sc, **kw)
        assert e.value.code == A.ZRT_ERR_NO_DEVICE


def test_group_argument_checks():
    sc, cam = scenes_py.three_balls()
    for kw in (dict(devices=[]), dict(devices=[0, 0]) if Z.device_count() else dict(devices=list(range(65))),
               dict(device=0, comm_id=Z.comm_id(), rank=2, world=2)):
        with pytest.raises(Z.ZrtError) as e:
            Z.MultiScene(sc, **kw)
        assert e.value.code in (A.ZRT_ERR_INVALID, A.ZRT_ERR_NO_DEVICE)


@pytest.mark.gpu
def test_group_of_one_is_zrt_render():
    sc, cam = scenes_py.three_balls()
    p = A.make_params(96, 64, 12, 30, x_limit=A.ZRT_XLIMIT_WIDTH)
    with Z.Scene(sc, device=0) as dev, Z.MultiScene(sc, devices=[0]) as grp:
        img1, c1, _ = dev.render(cam, p)
        img2, c2, tm = grp.render(cam, p)
        assert c1.as_dict() == c2.as_dict()
        np.testing.assert_allclose(img2, img1, rtol=1e-6, atol=1e-7)  # sum * (1/spp) on both sides, one association
        assert tm.launches >= 2 and tm.total_ms > 0
        img3, c3, _ = grp.render(cam, p, to_host=False)
        assert img3 is None and c3.as_dict() == c1.as_dict()
        for bad in (A.make_params(96, 64, 12, 30, flags=A.ZRT_FLAG_RAW_SUM), A.make_params(96, 64, 12, 30, sample_begin=1, sample_end=3)):
            with pytest.raises(Z.ZrtError):
                grp.render(cam, bad)


@pytest.mark.gpu
def test_reload_swaps_the_scene_on_every_local_device():
    """zrt_multi_reload: a new scene for the same group (communicator and accumulators kept), created on one host thread per
    local device; what bench.py's end-to-end leg does every step."""
    n = min(Z.device_count(), 4)
    sc_a, cam_a = scenes_py.three_balls()
    sc_b, cam_b = scenes_py.small_test_scene()
    p = A.make_params(80, 60, 9, 30, x_limit=A.ZRT_XLIMIT_WIDTH)
    with Z.Scene(sc_a, device=0) as da, Z.Scene(sc_b, device=0) as db:
        img_a, c_a, _ = da.render(cam_a, p)
        img_b, c_b, _ = db.render(cam_b, p)
    with Z.MultiScene(sc_a, devices=list(range(n))) as grp:
        img, c, _ = grp.render(cam_a, p)
        assert c.as_dict() == c_a.as_dict()
        for _ in range(2):
            grp.reload(sc_b)
            img, c, _ = grp.render(cam_b, p)
            assert c.as_dict() == c_b.as_dict()
            np.testing.assert_allclose(img, img_b, rtol=1e-5, atol=1e-6)
            grp.reload(sc_a)
            img, c, _ = grp.render(cam_a, p)
            assert c.as_dict() == c_a.as_dict()
            np.testing.assert_allclose(img, img_a, rtol=1e-5, atol=1e-6)


def _worlds():
    n = Z.device_count()
    return [w for w in (2, 3, 4, 8) if w <= n]


@pytest.mark.gpu
def test_n_gpu_image_equals_one_gpu_image():
    """The north-star correctness statement at N > 1, on hardware: same paths, same counters, image to f32 association."""
    if Z.device_count() < 2:
        pytest.skip("needs at least 2 devices")
    cases = [("three_balls", scenes_py.three_balls, A.make_params(160, 120, 37, 30, x_limit=A.ZRT_XLIMIT_WIDTH)),
             ("teapot", scenes_py.teapot_and_ball, A.make_params(96, 96, 16, 30)),
             ("bunny_glass", lambda: scenes_py.bunny_and_ball(dielectric=True), A.make_params(96, 96, 16, 30)),
             ("three_balls_spp1", scenes_py.three_balls, A.make_params(64, 64, 1, 30))]  # fewer samples than ranks
    for name, make, p in cases:
        sc, cam = make()
        with Z.Scene(sc, device=0) as dev:
            img1, c1, _ = dev.render(cam, p)
        for w in _worlds():
            with Z.MultiScene(sc, devices=list(range(w))) as grp:
                for _ in range(2):  # the second call reuses every buffer
                    img, c, tm = grp.render(cam, p)
                    assert c.as_dict() == c1.as_dict(), (name, w)
                    np.testing.assert_allclose(img, img1, rtol=1e-5, atol=1e-6, err_msg=f"{name} on {w} GPUs")


@pytest.mark.gpu
def test_headline_on_all_gpus_against_published_numbers(capsys):
    """C5 (7-spheres 1000x1000, 1000 spp, depth 30) on every visible device: counters exactly those of one GPU and within
    0.5 % of README.md:49-61, per-channel RMSE against showcase/7-spheres.png below 1 %."""
    n = Z.device_count()
    if n < 2:
        pytest.skip("needs at least 2 devices")
    sc, cam = scenes_py.three_balls()
    p = A.make_params(1000, 1000, 1000, 30)
    with Z.Scene(sc, device=0) as dev:
        img1, c1, _ = dev.render(cam, p)
    with Z.MultiScene(sc, devices=list(range(n))) as grp:
        img, c, tm = grp.render(cam, p)
    assert c.as_dict() == c1.as_dict()
    np.testing.assert_allclose(img, img1, rtol=1e-5, atol=1e-6)
    pub = {"rays_processed": 2144645362, "reflections": 1144753226, "background_hits": 999892115}
    for k, v in pub.items():
        assert abs(getattr(c, k) / v - 1) < 0.005
    gold = np.array(Image.open(os.path.join(os.path.dirname(__file__), "golden", "showcase_7spheres_1000.png")))
    gold = gold[::-1].astype(np.float64) / 255.0
    q = np.floor(np.clip(255.999 * img.astype(np.float64), 0, 255)) / 255.0
    rmse = np.sqrt(((q - gold) ** 2).mean(axis=(0, 1)))
    with capsys.disabled():
        print(f"\n[headline x{n} GPUs] RMSE vs showcase {rmse.round(5).tolist()}, rays {c.rays_processed}, "
              f"device {tm.total_ms:.2f} ms (trace {tm.kernel_ms:.2f} ms)")
    assert (rmse < 0.01).all()
