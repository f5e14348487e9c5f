"""The warp-level control flow of the persistent kernels, exercised WITHOUT a GPU (test infrastructure, tools/emu).

tools/emu compiles the library's own kernel sources with g++ against a stand-in cuda_runtime.h: every CUDA thread is a
fiber, every *_sync intrinsic a rendezvous of the 32 fibers of a warp, one thread block at a time.  What this checks is
the part the oracle cannot see from outside and a GPU box is slow to iterate on: the item queue, the slot pools and rings
of k_trace_pool3 / k_trace_bpool, the warp schedulers of k_trace_ws - no deadlock, no lost or duplicated slot, no lanes of
one warp in different collectives - by demanding the oracle's counters and pixels from every kernel variant.
It is not a CPU path of the product: zraytrace_b200 loads libzrt.so (nvcc, sm_100a only); the emulation library is
loaded only by tools/emu/*.py through ZRT_LIB_PATH, in a subprocess."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tools", "emu")


@pytest.fixture(scope="module")
def emu_lib():
    r = subprocess.run(["make", "-C", EMU, "-j8"], capture_output=True, text=True)
    if r.returncode != 0 or not os.path.exists(os.path.join(EMU, "libzrt_emu.so")):
        pytest.skip("tools/emu does not build here: " + r.stderr[-300:])
    return os.path.join(EMU, "libzrt_emu.so")


@pytest.mark.parametrize("scenes,size,spp", [(["three_balls"], 40, 12), (["teapot", "bunny_glass"], 28, 8), (["teapot_circle"], 24, 6)])
def test_every_kernel_variant_matches_the_oracle_under_emulation(emu_lib, scenes, size, spp):
    r = subprocess.run([sys.executable, os.path.join(EMU, "run_emu.py"), *scenes, "--size", str(size), "--spp", str(spp),
                        "--chunks", "0", "1", "4"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and r.stdout.strip().endswith("all ok"), r.stdout[-3000:] + r.stderr[-2000:]
    assert "pool" in r.stdout and "warp" in r.stdout and "thread" in r.stdout
