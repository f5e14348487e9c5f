#!/usr/bin/env python
"""Host-side breakdown of one end-to-end step (scene create -> render -> destroy) through the C ABI."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from zraytrace_b200 import host, lib as Z
wl = dict(bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c5"])
if len(sys.argv) > 2: wl["spp"] = int(sys.argv[2])
chunks = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hs = host.HostScene(wl["scene"], variant=wl["variant"], aspect_ratio=wl.get("aspect", 1.0))
p = bench.params_for(wl, sample_chunks=chunks)
for i in range(5):
    t0 = time.perf_counter(); sc = Z.Scene(hs, device=0)
    t1 = time.perf_counter(); img, c, tm = sc.render(hs.camera, p)
    t2 = time.perf_counter(); img2, c2, tm2 = sc.render(hs.camera, p)
    t3 = time.perf_counter(); sc.close()
    t4 = time.perf_counter()
    print(json.dumps({"create_ms": 1e3*(t1-t0), "render1_ms": 1e3*(t2-t1), "render2_ms": 1e3*(t3-t2), "destroy_ms": 1e3*(t4-t3),
                      "kernel_ms": tm.kernel_ms, "kernel2_ms": tm2.kernel_ms, "dev_total_ms": tm.total_ms, "prepare_ms": tm.prepare_ms, "Mrays_s": c.rays_processed/tm2.kernel_ms/1e3}))
