#!/usr/bin/env python
"""TEST INFRASTRUCTURE: sweep the K1p scheduler thresholds under the emulator and rank them by a static cost model
(warp instructions per section run, from tools/sass_lines.py).  python tools/emu/sweep_emu.py --scene teapot"""
import argparse
import itertools
import json
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

COST = {"p.iterations": 10, "p.sched slow path": 45, "p.N node step": 60, "p.L leaf test": 90, "p.X sections": 85, "p.pop": 30,
        "p.S regen": 260, "p.S lambertian": 330, "p.S metal": 230, "p.S glass": 300}


def run(args):
    scene, size, spp, slots, th = args
    out = subprocess.run([sys.executable, "tools/emu/prof_emu.py", "--scene", scene, "--size", str(size), "--spp", str(spp),
                          "--kernel", "pool", "--slots", str(slots), "--thresholds", th], capture_output=True, text=True).stdout
    cost, rows = 0.0, {}
    for line in out.splitlines()[2:]:
        name = line[:28].strip()
        f = line[28:].split()
        rows[name] = (float(f[1]), float(f[2]))
        cost += COST.get(name, 0) * float(f[1])
    return th, slots, cost / 32.0, rows


ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="teapot")
ap.add_argument("--size", type=int, default=48)
ap.add_argument("--spp", type=int, default=32)
ap.add_argument("--slots", type=int, nargs="*", default=[128])
ap.add_argument("--node", type=int, nargs="*", default=[8, 12, 16, 20, 24])
ap.add_argument("--leaf", type=int, nargs="*", default=[2, 4, 8])
ap.add_argument("--idle", type=int, nargs="*", default=[2, 4, 8, 12])
ap.add_argument("--batch", type=int, nargs="*", default=[16])
a = ap.parse_args()
jobs = [(a.scene, a.size, a.spp, s, f"{n},{l},{i},{b}") for s in a.slots for n, l, i, b in itertools.product(a.node, a.leaf, a.idle, a.batch)]
with ThreadPoolExecutor(8) as ex:
    res = list(ex.map(run, jobs))
res.sort(key=lambda r: r[2])
for th, slots, cost, rows in res[:12] + res[-3:]:
    n, l, x = rows.get("p.N node step", (0, 0)), rows.get("p.L leaf test", (0, 0)), rows.get("p.X sections", (0, 0))
    print(f"slots={slots} th={th:12s} est {cost:6.1f} warp-inst/ray | N {n[0]:5.2f}@{n[1]:4.1f} L {l[0]:5.2f}@{l[1]:4.1f} X {x[0]:5.2f} it {rows['p.iterations'][0]:5.2f}")
