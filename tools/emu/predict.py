#!/usr/bin/env python
"""TEST INFRASTRUCTURE: predicted thread instructions per ray of a kernel variant, without a GPU.

  executions per source line   from a gcov build of the emulator (make -C tools/emu COV=1: every CUDA thread a fiber)
x SASS instructions per line   from nvdisasm --print-line-info on the real sm_100a object (zraytrace_b200/csrc/build)
= thread instructions, by source line / region; divided by rays.  Packed f32x2 intrinsics (their SASS carries the line of a
CUDA header) are counted from the emulator's own __ffma2_rn / __fadd2_rn / __fmul2_rn at one instruction each.
Calibration: k_trace_pool<7,128,7> on the 7-spheres scene measures 465 thread instructions per ray under ncu
(profiles/r2_a_pool128_summary.txt); this tool's figure for the same source is printed beside whatever it is asked for.

  python tools/emu/predict.py --scene three_balls --size 64 --spp 16 --kernel pool --symbol k_trace_poolILi7ELi128 [--top 40]"""
import argparse
import os
import re
import subprocess
import sys
import tempfile
from collections import Counter, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
EMU = os.path.join(ROOT, "tools", "emu")
COVDIR = os.path.join(EMU, "build_cov")
OBJ = os.path.join(ROOT, "zraytrace_b200", "csrc", "build", "zrt_kernels.o")

RUN = r"""
import os, sys, json
sys.path.insert(0, %(root)r)
from tests import scenes_py
from zraytrace_b200 import _abi as A, lib as Z
S = {"three_balls": scenes_py.three_balls, "teapot": scenes_py.teapot_and_ball, "bunny_glass": lambda: scenes_py.bunny_and_ball(dielectric=True),
     "man": scenes_py.man_and_ball, "teapot_circle": scenes_py.teapot_and_ball_circle}
K = {"thread": A.ZRT_FLAG_KERNEL_THREAD, "warp": A.ZRT_FLAG_KERNEL_WARP, "pool": A.ZRT_FLAG_KERNEL_POOL, "auto": 0}
sc, cam = S[%(scene)r]()
p = A.make_params(%(size)d, %(size)d, %(spp)d, 30, sample_chunks=%(chunks)d, flags=K[%(kernel)r])
with Z.Scene(sc, device=0) as dev:
    img, c, _ = dev.render(cam, p)
print(json.dumps(c.as_dict()))
"""


def static_sass(symbol):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", OBJ], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    cnt, cur, active, packed = Counter(), None, False, 0
    for line in dis.splitlines():
        if line.startswith("\t.section\t.text."):
            active = symbol in line
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", line):
            cnt[cur] += 1
    return cnt


def gcov_counts():
    tmp = tempfile.mkdtemp()
    os.symlink(os.path.join(ROOT, "zraytrace_b200"), os.path.join(os.path.dirname(tmp), "zraytrace_b200")) if False else None
    # gcov resolves the source paths recorded at compile time (relative to tools/emu), so it runs there; its *.gcov files are removed again
    subprocess.run(["gcov", "-o", COVDIR, os.path.join(COVDIR, "zrt_kernels.gcda")], cwd=EMU, capture_output=True)
    out = defaultdict(dict)
    for f in os.listdir(EMU):
        if not f.endswith(".gcov"):
            continue
        name = f[:-5]
        path = os.path.join(EMU, f)
        lines = open(path, errors="replace").read().split("\n")
        os.remove(path)
        for line in lines:
            m = re.match(r"\s*([0-9#=\-*]+)\*?:\s*(\d+):", line)
            if m and m.group(1)[0].isdigit():
                n = int(m.group(1).rstrip("*"))
                ln = int(m.group(2))
                out[name][ln] = max(out[name].get(ln, 0), n)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="three_balls")
    ap.add_argument("--size", type=int, default=64)
    ap.add_argument("--spp", type=int, default=16)
    ap.add_argument("--chunks", type=int, default=0)
    ap.add_argument("--kernel", default="pool")
    ap.add_argument("--symbol", default="k_trace_poolILi7ELi128")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--env", nargs="*", default=[])
    a = ap.parse_args()
    for root, _, files in os.walk(COVDIR):
        for f in files:
            if f.endswith(".gcda"):
                os.remove(os.path.join(root, f))
    env = dict(os.environ, ZRT_LIB_PATH=os.path.join(EMU, "libzrt_emu_cov.so"))
    for kv in a.env:
        k, v = kv.split("=", 1)
        env[k] = v
    r = subprocess.run([sys.executable, "-c", RUN % dict(root=ROOT, scene=a.scene, size=a.size, spp=a.spp, chunks=a.chunks, kernel=a.kernel)],
                       env=env, capture_output=True, text=True)
    if r.returncode:
        print(r.stderr)
        return 1
    import json
    c = json.loads(r.stdout.strip().splitlines()[-1])
    rays = c["rays_processed"]
    sass = static_sass(a.symbol)
    cov = gcov_counts()
    src = {n: open(os.path.join(ROOT, "zraytrace_b200", "csrc", n)).read().split("\n") for n in ("zrt_kernels.cu", "zrt_pool_bvh.cuh", "zrt_pool_spheres.cuh", "zrt_math.cuh")}
    total, rows, unmatched = 0.0, [], 0
    for (f, ln), n in sass.items():
        if f in cov and ln in cov[f]:
            t = cov[f][ln] * n
            total += t
            rows.append((t, f, ln, n, cov[f][ln]))
        else:
            unmatched += n
    # packed intrinsics: one SASS instruction per call
    emu = cov.get("cuda_runtime.h", {})
    esrc = open(os.path.join(EMU, "cuda_runtime.h")).read().split("\n")
    packed = 0
    for i, line in enumerate(esrc, 1):
        if re.match(r"inline float2 __(fadd2|fmul2|ffma2)_rn", line):
            packed += emu.get(i, 0)
    total += packed
    print(f"{a.scene} {a.size}x{a.size} {a.spp} spp kernel={a.kernel} symbol={a.symbol}: {rays} rays")
    print(f"predicted thread instructions per ray: {total / rays:.1f}   (packed f32x2: {packed / rays:.1f}; static SASS on lines the emulator never "
          f"mapped: {unmatched} of {sum(sass.values())})")
    rows.sort(reverse=True)
    for t, f, ln, n, ex in rows[:a.top]:
        txt = src[f][ln - 1].strip()[:88] if f in src and ln <= len(src[f]) else ""
        print(f"{t / rays:7.2f}  {f}:{ln:<5d} sass={n:<4d} exec/ray={ex / rays:6.3f}  {txt}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
