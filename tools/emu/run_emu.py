#!/usr/bin/env python
"""TEST INFRASTRUCTURE: run small renders through tools/emu/libzrt_emu.so (the kernels as CPU fibers, see cuda_runtime.h)
and compare every kernel variant with the oracle and with the thread kernel.  ZRT_LIB_PATH must point at the emulation
library BEFORE zraytrace_b200.lib is imported; this script sets it itself.
  python tools/emu/run_emu.py [scene ...] [--size N] [--spp N] [--depth N]"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["ZRT_LIB_PATH"] = os.path.join(ROOT, "tools", "emu", "libzrt_emu.so")
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from oracle import zro_py  # noqa: E402
from tests import scenes_py  # noqa: E402
from zraytrace_b200 import _abi as A  # noqa: E402
from zraytrace_b200 import lib as Z  # noqa: E402

SCENES = {"three_balls": scenes_py.three_balls, "teapot": scenes_py.teapot_and_ball,
          "bunny_glass": lambda: scenes_py.bunny_and_ball(dielectric=True), "man": scenes_py.man_and_ball,
          "teapot_circle": scenes_py.teapot_and_ball_circle}
KERNELS = {"thread": A.ZRT_FLAG_KERNEL_THREAD, "warp": A.ZRT_FLAG_KERNEL_WARP, "pool": A.ZRT_FLAG_KERNEL_POOL}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("scenes", nargs="*", default=["three_balls", "teapot"])
    ap.add_argument("--size", type=int, default=24)
    ap.add_argument("--spp", type=int, default=6)
    ap.add_argument("--depth", type=int, default=30)
    ap.add_argument("--chunks", type=int, nargs="*", default=[0, 1])
    ap.add_argument("--slots", nargs="*", default=["64", "96", "128"])
    a = ap.parse_args()
    bad = 0
    for name in a.scenes:
        sc, cam = SCENES[name]()
        with Z.Scene(sc, device=0) as dev:
            for chunks in a.chunks:
                p = A.make_params(a.size, a.size, a.spp, a.depth, sample_chunks=chunks)
                img_o, c_o, _ = zro_py.render(sc, cam, p, rng=zro_py.RNG_CTR, math=zro_py.MATH_SPEC)
                ref = None
                for kname, flag in KERNELS.items():
                    for slots in (a.slots if kname == "pool" else [""]):
                        if slots:
                            os.environ["ZRT_POOL_SLOTS"] = slots
                        p.flags = flag
                        t0 = time.time()
                        img, c, _ = dev.render(cam, p)
                        dt = time.time() - t0
                        ok_c = c.as_dict() == c_o.as_dict()
                        ok_i = np.allclose(img, img_o, rtol=2e-5, atol=1e-6)
                        if ref is None:
                            ref = img
                        # pinned slice counts: the same f32 sums in the same order, bit for bit.  chunks = 0: every kernel picks
                        # its own slice count (k_trace_pool3 takes 16 below 384 spp), so only the association may differ
                        ok_b = np.array_equal(ref.view(np.uint32), img.view(np.uint32)) if chunks else np.allclose(ref, img, rtol=1e-5, atol=1e-6)
                        bad += not (ok_c and ok_i and ok_b)
                        print(f"{name:14s} chunks={chunks} {kname:6s} {slots:4s} counters={'ok' if ok_c else 'DIFF'} "
                              f"oracle={'ok' if ok_i else 'DIFF'} bits={'ok' if ok_b else 'DIFF'} rays={c.rays_processed} {dt:.1f}s", flush=True)
                        if not ok_c:
                            print("   emu   ", c.as_dict(), "\n   oracle", c_o.as_dict())
    print("FAILED" if bad else "all ok")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
