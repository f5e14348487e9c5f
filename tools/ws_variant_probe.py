import os, sys
sys.path.insert(0, os.getcwd())
import bench
from zraytrace_b200 import _abi as A, host, lib as Z
for variant, name in ((0, "bunny metal"), (1, "bunny glass")):
    wl = dict(bench.WORKLOADS["c3"]); wl["spp"] = 128
    hs = host.HostScene(2, variant=variant)
    with Z.Scene(hs, device=0) as sc:
        t = min(sc.render(hs.camera, bench.params_for(wl))[2].kernel_ms for _ in range(3))
        w = min(sc.render(hs.camera, bench.params_for(wl, flags=A.ZRT_FLAG_KERNEL_WARP))[2].kernel_ms for _ in range(3))
        c = sc.render(hs.camera, bench.params_for(wl))[1]
    print(name, "thread", round(t, 3), "warp", round(w, 3), round(w / t, 3), "rays/sample", round(c.rays_processed / c.samples_processed, 2))
