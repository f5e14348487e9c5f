#!/usr/bin/env python
"""Run-to-run determinism on the device: the image bits and counters of repeated renders must be identical although
which lane / slot traces which item depends on timing (per-item partial sums land in fixed slabs, k_resolve adds them in order).
  python tools/determinism_check.py [c5 c2 ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import bench  # noqa: E402
from zraytrace_b200 import host, lib as Z  # noqa: E402

bad = 0
for name in sys.argv[1:] or ["c5", "c2", "c3"]:
    wl = dict(bench.WORKLOADS[name])
    wl["spp"] = max(32, wl["spp"] // 4)
    hs = host.HostScene(wl["scene"], variant=wl["variant"], aspect_ratio=wl.get("aspect", 1.0))
    with Z.Scene(hs, device=0) as sc:
        ref = None
        for rep in range(4):
            img, c, tm = sc.render(hs.camera, bench.params_for(wl))
            if ref is None:
                ref = (img.copy(), c.as_dict())
            same = np.array_equal(ref[0].view(np.uint32), img.view(np.uint32)) and ref[1] == c.as_dict()
            bad += not same
            print(f"{name} rep {rep}: {'identical' if same else 'DIFFERENT'} ({tm.kernel_ms:.2f} ms)", flush=True)
print("FAILED" if bad else "deterministic")
sys.exit(1 if bad else 0)
