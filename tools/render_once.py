#!/usr/bin/env python
"""Render one workload a few times through the C ABI and print timing; the short command used under ncu.
  python tools/render_once.py --workload c5 [--spp N] [--reps R] [--reftree] [--chunks C]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from zraytrace_b200 import _abi as A  # noqa: E402
from zraytrace_b200 import host, lib as Z  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c5")
ap.add_argument("--spp", type=int, default=0)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--reftree", action="store_true")
ap.add_argument("--chunks", type=int, default=0)
ap.add_argument("--primary", action="store_true")
ap.add_argument("--size", type=int, default=0, help="override width = height")
ap.add_argument("--kernel", default="auto", choices=["auto", "thread", "sorted", "warp", "x2", "pool"])
a = ap.parse_args()
wl = dict(bench.WORKLOADS[a.workload])
if a.spp:
    wl["spp"] = a.spp
if a.size:
    wl["w"] = wl["h"] = a.size
hs = host.HostScene(wl["scene"], variant=wl["variant"], aspect_ratio=wl.get("aspect", 1.0))
kflag = {"auto": 0, "thread": A.ZRT_FLAG_KERNEL_THREAD, "sorted": A.ZRT_FLAG_KERNEL_SORTED, "warp": A.ZRT_FLAG_KERNEL_WARP, "x2": A.ZRT_FLAG_KERNEL_X2, "pool": A.ZRT_FLAG_KERNEL_POOL}[a.kernel]
p = bench.params_for(wl, flags=(A.ZRT_FLAG_BVH_REFERENCE if a.reftree else 0) | kflag, sample_chunks=a.chunks)
with Z.Scene(hs, device=0) as sc:
    for i in range(a.reps):
        if a.primary:
            sc.primary_hits(hs.camera, p)
        img, c, t = sc.render(hs.camera, p)
        print(json.dumps({"rep": i, "kernel": a.kernel, "workload": a.workload, "spp": wl["spp"], "kernel_ms": t.kernel_ms, "total_ms": t.total_ms,
                          "prepare_ms": t.prepare_ms, "launches": t.launches, "bvh_nodes": t.bvh_nodes,
                          "Mrays_s_kernel": c.rays_processed / t.kernel_ms / 1e3, **c.as_dict()}))
