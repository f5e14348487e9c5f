#!/usr/bin/env python
"""TEST INFRASTRUCTURE: section profile of a kernel variant under the emulator (tools/emu/cuda_runtime.h): how often a warp
ran each section and with how many lanes, per ray.  Used to tune the warp schedulers before spending GPU time.
  python tools/emu/prof_emu.py --scene teapot --size 64 --spp 16 --kernel pool [--slots 128] [--thresholds 12,2,8,16]"""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["ZRT_LIB_PATH"] = os.path.join(ROOT, "tools", "emu", "libzrt_emu.so")
sys.path.insert(0, ROOT)
from tests import scenes_py  # noqa: E402
from zraytrace_b200 import _abi as A  # noqa: E402
from zraytrace_b200 import lib as Z  # noqa: E402

NAMES = {0: "p.iterations", 1: "p.sched slow path", 2: "p.N node step", 3: "p.L leaf test", 13: "p.X sections", 4: "p.X hand-over", 5: "p.X re-arm",
         6: "p.S regen", 7: "p.S lambertian", 8: "p.S metal", 9: "p.S glass", 10: "p.S lamb img", 11: "p.S metal img", 12: "p.pop",
         20: "w.iterations", 21: "w.sched slow path", 22: "w.N node step", 23: "w.L leaf test", 24: "w.S sections", 25: "w.S background",
         26: "w.S hit record", 27: "w.S lambertian", 28: "w.S metal", 29: "w.S glass", 30: "w.S image texture", 31: "w.S regenerate",
         32: "w.S unit + bookkeeping", 33: "w.pop",
         40: "q.iterations", 41: "q.S regen", 42: "q.S lambertian", 43: "q.S metal", 44: "q.S glass", 45: "q.S lamb img", 46: "q.S hand-over",
         47: "q.common (unit + spheres)", 48: "q.primary spheres", 49: "q.secondary spheres",
         50: "t.iterations", 51: "t.top block", 52: "t.unit + closest hit", 53: "t.background", 54: "t.draw", 55: "t.regenerate",
         56: "t.hit record", 57: "t.lambertian", 58: "t.metal", 59: "t.glass", 60: "t.image texture"}
SCENES = {"three_balls": scenes_py.three_balls, "teapot": scenes_py.teapot_and_ball,
          "bunny_glass": lambda: scenes_py.bunny_and_ball(dielectric=True), "man": scenes_py.man_and_ball,
          "teapot_circle": scenes_py.teapot_and_ball_circle}
KERNELS = {"thread": A.ZRT_FLAG_KERNEL_THREAD, "warp": A.ZRT_FLAG_KERNEL_WARP, "pool": A.ZRT_FLAG_KERNEL_POOL}


def profile(scene, size, spp, kernel, depth=30, chunks=0, env=None):
    for k, v in (env or {}).items():
        os.environ[k] = v
    L = Z.lib()
    sc, cam = SCENES[scene]()
    p = A.make_params(size, size, spp, depth, sample_chunks=chunks, flags=KERNELS[kernel])
    with Z.Scene(sc, device=0) as dev:
        L.zrt_emu_prof_reset()
        img, c, _ = dev.render(cam, p)
        calls, lanes = (C.c_ulonglong * 64)(), (C.c_ulonglong * 64)()
        L.zrt_emu_prof_get(calls, lanes)
    return c, list(calls), list(lanes)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="teapot")
    ap.add_argument("--size", type=int, default=64)
    ap.add_argument("--spp", type=int, default=16)
    ap.add_argument("--depth", type=int, default=30)
    ap.add_argument("--chunks", type=int, default=0)
    ap.add_argument("--kernel", default="pool")
    ap.add_argument("--slots", default=None)
    ap.add_argument("--thresholds", default=None)
    a = ap.parse_args()
    env = {}
    if a.slots:
        env["ZRT_POOL_SLOTS"] = a.slots
    if a.thresholds:
        env["ZRT_POOL_THRESHOLDS" if a.kernel == "pool" else "ZRT_WS_THRESHOLDS"] = a.thresholds
    c, calls, lanes = profile(a.scene, a.size, a.spp, a.kernel, a.depth, a.chunks, env)
    rays = c.rays_processed
    print(f"{a.scene} {a.size}x{a.size} {a.spp} spp kernel={a.kernel} {env}: {rays} rays, {c.samples_processed} samples")
    print(f"{'section':28s} {'warp runs':>10s} {'per 32 rays':>12s} {'lanes/run':>10s}")
    for i in range(64):
        if calls[i]:
            print(f"{NAMES.get(i, str(i)):28s} {calls[i]:10d} {32.0 * calls[i] / rays:12.3f} {lanes[i] / calls[i]:10.2f}")


if __name__ == "__main__":
    main()
