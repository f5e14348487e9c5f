#!/usr/bin/env python
"""Small renders of every kernel variant for compute-sanitizer (memcheck): spheres, list, BVH (both trees), primary hits."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zraytrace_b200 import _abi as A, host, lib as Z
for idx, var, bvh in ((1, 0, True), (3, 0, True), (3, 0, False), (0, 0, True), (2, 1, True)):
    hs = host.HostScene(idx, variant=var)
    with Z.Scene(hs, device=0) as sc:
        for flags in (0, A.ZRT_FLAG_BVH_REFERENCE):
            w = 24 if not bvh and idx == 3 else 48
            p = A.make_params(w, w, 4, 8, bvh=bvh, flags=flags)
            sc.primary_hits(hs.camera, p)
            img, c, t = sc.render(hs.camera, p)
            print(idx, var, bvh, flags, c.rays_processed, flush=True)
print("selftest", Z.selftest(0))
