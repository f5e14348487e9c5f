#!/usr/bin/env python
"""Threshold sweep of the warp-scheduled BVH kernel (ZRT_FLAG_KERNEL_WARP) against the default kernel, in one process,
at a quarter of each workload's samples.  ZRT_WS_THRESHOLDS="node,leaf,shade" is read by makePlan on every render.
  gpurun -- 'python tools/ws_sweep.py > gpurun_out/ws_sweep.log'"""
import itertools
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from zraytrace_b200 import _abi as A  # noqa: E402
from zraytrace_b200 import host, lib as Z  # noqa: E402


def best_ms(sc, cam, p, reps=3):
    return min(sc.render(cam, p)[2].kernel_ms for _ in range(reps))


for name in sys.argv[1:] or ["c2", "c3", "c4"]:
    wl = dict(bench.WORKLOADS[name])
    wl["spp"] //= 4
    hs = host.HostScene(wl["scene"], variant=wl["variant"], aspect_ratio=wl.get("aspect", 1.0))
    with Z.Scene(hs, device=0) as sc:
        os.environ.pop("ZRT_WS_THRESHOLDS", None)
        thread = best_ms(sc, hs.camera, bench.params_for(wl))
        print(f"{name} thread {thread:.3f}", flush=True)
        rows = []
        for n, l, s in itertools.product((4, 8, 12, 16), (2, 4, 8), (12, 16, 20, 24, 28)):
            os.environ["ZRT_WS_THRESHOLDS"] = f"{n},{l},{s}"
            ms = best_ms(sc, hs.camera, bench.params_for(wl, flags=A.ZRT_FLAG_KERNEL_WARP))
            rows.append((ms, n, l, s))
            print(f"{name} warp {n},{l},{s} {ms:.3f} {ms / thread:.3f}", flush=True)
        rows.sort()
        print(f"{name} best", rows[:5], flush=True)
