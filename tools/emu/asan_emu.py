#!/usr/bin/env python
"""TEST INFRASTRUCTURE: every kernel under AddressSanitizer + UBSan (compute-sanitizer is closed on the GPU pool).
`make -C tools/emu ASAN=1` builds the kernel sources with -fsanitize=address,undefined against the emulator's
cuda_runtime.h; shared memory, the slice buffers, rings, stacks and scene arrays are then ordinary host memory, so an index
that strays is reported with a stack trace.  Run as  python tools/emu/asan_emu.py  (re-executes itself with libasan preloaded)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB = os.path.join(ROOT, "tools", "emu", "libzrt_emu_asan.so")

if os.environ.get("ZRT_ASAN_CHILD") != "1":
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tools", "emu"), "ASAN=1", "-j8"], stdout=subprocess.DEVNULL)
    asan = subprocess.check_output(["gcc", "-print-file-name=libasan.so"], text=True).strip()
    env = dict(os.environ, ZRT_ASAN_CHILD="1", LD_PRELOAD=asan, ZRT_LIB_PATH=LIB,
               ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0:abort_on_error=1", UBSAN_OPTIONS="halt_on_error=1:print_stacktrace=1")
    sys.exit(subprocess.call([sys.executable, os.path.abspath(__file__)], env=env))

sys.path.insert(0, ROOT)
from tests import scenes_py  # noqa: E402
from zraytrace_b200 import _abi as A, lib as Z  # noqa: E402

SCENES = (("three_balls", scenes_py.three_balls), ("teapot", scenes_py.teapot_and_ball),
          ("bunny_glass", lambda: scenes_py.bunny_and_ball(dielectric=True)), ("teapot_circle", scenes_py.teapot_and_ball_circle))
for name, make in SCENES:
    sc, cam = make()
    with Z.Scene(sc, device=0) as dev:
        for flag in (A.ZRT_FLAG_KERNEL_THREAD, A.ZRT_FLAG_KERNEL_WARP, A.ZRT_FLAG_KERNEL_POOL, A.ZRT_FLAG_BVH_REFERENCE, A.ZRT_FLAG_RUSSIAN_ROULETTE):
            for (w, h, spp, chunks) in ((20, 17, 5, 0), (1, 9, 3, 1), (33, 2, 40, 4)):
                for slots in ("64", "128"):
                    os.environ["ZRT_POOL_SLOTS"] = slots
                    img, c, _ = dev.render(cam, A.make_params(w, h, spp, 30, sample_chunks=chunks, flags=flag, x_limit=A.ZRT_XLIMIT_WIDTH))
        dev.primary_hits(cam, A.make_params(33, 9, 1, 30))
        dev.render_rgb8(cam, A.make_params(16, 8, 2, 30)) if hasattr(dev, "render_rgb8") else None
    print(f"{name}: clean", flush=True)
print("asan/ubsan: all kernels clean")
