"""ctypes binding of libzrt.so (include/zrt.h).  This is the host-side mirror of the reference's
`raytrace.render(allocator, random, camera, surfaces, render_params)` (raytrace.zig:136-138): same
arguments in, image + Progress counters out.  There is no CPU fallback anywhere in this package: if the
CUDA library is missing or no device is visible, calls raise."""
import ctypes as C
import os

import numpy as np

from . import _abi as A

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ZRT_LIB_PATH") or os.path.join(_HERE, "libzrt.so")  # override: A/B builds of the library
_lib = None


class ZrtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libzrt error {code}: {msg}")
        self.code = code


def lib():
    """Load libzrt.so (built in-tree by __graft_entry__.build() / make -C zraytrace_b200/csrc)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ZrtError(A.ZRT_ERR_NO_DEVICE, f"{LIB_PATH} is not built; run __graft_entry__.build() "
                                                "(there is no Python/CPU fallback)")
        L = C.CDLL(LIB_PATH)
        P = C.POINTER
        L.zrt_device_count.restype = C.c_int
        L.zrt_last_error.restype = C.c_char_p
        L.zrt_scene_create.argtypes = [P(A.SceneDesc), C.c_int, P(C.c_void_p)]
        L.zrt_scene_destroy.argtypes = [C.c_void_p]
        L.zrt_scene_destroy.restype = None
        L.zrt_render.argtypes = [C.c_void_p, P(A.Camera), P(A.Params), C.c_void_p, P(A.Counters), P(A.Timing)]
        L.zrt_render_rgb8.argtypes = [C.c_void_p, P(A.Camera), P(A.Params), C.c_void_p, P(A.Counters), P(A.Timing)]
        L.zrt_render_device.argtypes = [C.c_void_p, P(A.Camera), P(A.Params), C.c_void_p, C.c_void_p, C.c_void_p]
        L.zrt_primary_hits.argtypes = [C.c_void_p, P(A.Camera), P(A.Params), C.c_int, C.c_void_p, C.c_void_p]
        L.zrt_scene_bvh_info.argtypes = [C.c_void_p, C.c_uint32, P(A.BvhInfo)]
        L.zrt_scene_bvh_order.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.zrt_scene_launch_count.argtypes = [C.c_void_p]
        L.zrt_scene_launch_count.restype = C.c_uint64
        L.zrt_build_features.restype = C.c_uint32
        L.zrt_comm_id.argtypes = [C.c_void_p]
        L.zrt_multi_create.argtypes = [P(A.SceneDesc), P(C.c_int), C.c_int, P(C.c_void_p)]
        L.zrt_multi_create_rank.argtypes = [P(A.SceneDesc), C.c_int, C.c_void_p, C.c_int, C.c_int, P(C.c_void_p)]
        L.zrt_multi_destroy.argtypes = [C.c_void_p]
        L.zrt_multi_reload.argtypes = [C.c_void_p, P(A.SceneDesc)]
        L.zrt_multi_destroy.restype = None
        L.zrt_multi_render.argtypes = [C.c_void_p, P(A.Camera), P(A.Params), C.c_void_p, P(A.Counters), P(A.Timing)]
        L.zrt_multi_world_size.argtypes = [C.c_void_p]
        L.zrt_multi_launch_count.argtypes = [C.c_void_p]
        L.zrt_multi_launch_count.restype = C.c_uint64
        L.zrt_multi_image_device.argtypes = [C.c_void_p]
        L.zrt_multi_image_device.restype = C.c_void_p
        L.zrt_trace_statistics.argtypes = [C.c_void_p, P(A.Camera), P(A.Params), P(A.TraceStats)]
        L.zrt_selftest.argtypes = [C.c_int, P(C.c_uint64)]
        L.zrt_measure_peaks.argtypes = [C.c_int, P(C.c_double), C.c_int]
        L.zrt_pinned_alloc.argtypes = [C.c_size_t, P(C.c_void_p)]
        L.zrt_pinned_free.argtypes = [C.c_void_p]
        L.zrt_pinned_free.restype = None
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise ZrtError(rc, lib().zrt_last_error().decode())


def device_count():
    return lib().zrt_device_count()


def has_experiments():
    """True if the library was built with EXPERIMENTS=1 (k_trace_sorted / k_trace_x2 compiled in)."""
    return bool(lib().zrt_build_features() & A.ZRT_FEATURE_EXPERIMENTS)


class HostImage:
    """Page-locked host buffer (`zrt_pinned_alloc`) seen as a numpy array: pass `.array` as `out=` to Scene.render /
    render_rgb8 so the image comes back in one DMA.  The array is only valid until close()."""

    def __init__(self, shape, dtype=np.float32):
        self._p = C.c_void_p()
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        _check(lib().zrt_pinned_alloc(n, C.byref(self._p)))
        self.array = np.frombuffer((C.c_char * n).from_address(self._p.value), dtype=dtype).reshape(shape)

    def close(self):
        if self._p:
            self.array = None
            lib().zrt_pinned_free(self._p)
            self._p = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _out_buffer(out, shape, dtype):
    if out is None:
        return np.empty(shape, dtype)
    if out.dtype != dtype or out.shape != shape or not out.flags["C_CONTIGUOUS"]:
        raise ValueError(f"out must be a C-contiguous {np.dtype(dtype).name} array of shape {shape}")
    return out


class Scene:
    """Device-resident flattened scene (`zrt_scene*`).  device=-1 keeps it on the host (inspection only)."""

    def __init__(self, built_scene_or_desc, device=0):
        desc = getattr(built_scene_or_desc, "desc", built_scene_or_desc)
        self.n_surfaces = desc.n_surfaces
        self.device = device
        self._h = C.c_void_p()
        _check(lib().zrt_scene_create(C.byref(desc), device, C.byref(self._h)))

    def close(self):
        if self._h:
            lib().zrt_scene_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def render(self, camera, params, out=None):
        """-> (image float32 [H][W][3] row 0 = bottom, Counters, Timing)   (raytrace.zig:136-203)
        out: optional caller-owned C-contiguous float32 [H][W][3] host buffer (page-locked memory makes the
        device-to-host copy one DMA instead of a staged copy into fresh pages)."""
        img = _out_buffer(out, (params.height, params.width, 3), np.float32)
        cnt, tm = A.Counters(), A.Timing()
        _check(lib().zrt_render(self._h, C.byref(camera), C.byref(params), img.ctypes.data, C.byref(cnt), C.byref(tm)))
        return img, cnt, tm

    def render_rgb8(self, camera, params, out=None):
        """-> (uint8 [H][W][3] with row 0 = TOP scanline, quantised on the device like png_image.zig:136-140,
        Counters, Timing)"""
        img = _out_buffer(out, (params.height, params.width, 3), np.uint8)
        cnt, tm = A.Counters(), A.Timing()
        _check(lib().zrt_render_rgb8(self._h, C.byref(camera), C.byref(params), img.ctypes.data, C.byref(cnt), C.byref(tm)))
        return img, cnt, tm

    def render_device(self, camera, params, d_rgb_ptr, d_counters_ptr, stream_ptr=0):
        """Asynchronous render into device memory (raw pointers, e.g. torch tensor .data_ptr())."""
        _check(lib().zrt_render_device(self._h, C.byref(camera), C.byref(params), C.c_void_p(d_rgb_ptr),
                                       C.c_void_p(d_counters_ptr), C.c_void_p(stream_ptr)))

    def primary_hits(self, camera, params, jitter=0):
        ids = np.empty((params.height, params.width), np.uint32)
        t = np.empty((params.height, params.width), np.float32)
        _check(lib().zrt_primary_hits(self._h, C.byref(camera), C.byref(params), jitter, ids.ctypes.data, t.ctypes.data))
        return ids, t

    def trace_statistics(self, camera, params):
        """Event counts (node visits, primitive tests, ...) of one render, from the instrumented kernel."""
        st = A.TraceStats()
        _check(lib().zrt_trace_statistics(self._h, C.byref(camera), C.byref(params), C.byref(st)))
        return st

    def launch_count(self):
        return int(lib().zrt_scene_launch_count(self._h))

    def bvh_info(self, flags=0):
        info = A.BvhInfo()
        _check(lib().zrt_scene_bvh_info(self._h, flags, C.byref(info)))
        return info

    def bvh_order(self):
        order = np.zeros(self.n_surfaces, np.uint32)
        vis = np.zeros(self.n_surfaces, np.uint8)
        _check(lib().zrt_scene_bvh_order(self._h, order.ctypes.data, vis.ctypes.data))
        return order, vis.astype(bool)


def comm_id():
    """128 opaque bytes (an ncclUniqueId) that every process of a one-process-per-GPU group must be given."""
    buf = (C.c_uint8 * A.ZRT_COMM_ID_BYTES)()
    _check(lib().zrt_comm_id(buf))
    return bytes(buf)


def nccl_version():
    return int(lib().zrt_nccl_version())


class MultiScene:
    """`zrt_multi*`: the scene replicated on several GPUs, samples-per-pixel split across them, ONE NCCL reduce of the
    f32 accumulators to rank 0 (include/zrt.h "multi-GPU").

    MultiScene(scene, devices=[0, 1, ...])                      one process drives all devices
    MultiScene(scene, device=d, comm_id=id, rank=r, world=w)    one process per device (torchrun / MPI)"""

    def __init__(self, built_scene_or_desc, devices=None, device=None, comm_id=None, rank=0, world=1):
        desc = getattr(built_scene_or_desc, "desc", built_scene_or_desc)
        self._h = C.c_void_p()
        if device is None:
            devs = list(devices) if devices is not None else [0]
            arr = (C.c_int * len(devs))(*devs)
            _check(lib().zrt_multi_create(C.byref(desc), arr, len(devs), C.byref(self._h)))
            self.rank, self.world = 0, len(devs)
        else:
            idbuf = (C.c_uint8 * A.ZRT_COMM_ID_BYTES)(*(comm_id or bytes(A.ZRT_COMM_ID_BYTES)))
            _check(lib().zrt_multi_create_rank(C.byref(desc), device, idbuf, rank, world, C.byref(self._h)))
            self.rank, self.world = rank, world

    def close(self):
        if self._h:
            lib().zrt_multi_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def render(self, camera, params, out=None, to_host=True):
        """-> (image float32 [H][W][3] or None, Counters, Timing).  Image and counters are meaningful on the process
        that owns rank 0; to_host=False leaves the image on rank 0's device (timing.total_ms is device time)."""
        img = None
        if to_host and self.rank == 0:
            img = _out_buffer(out, (params.height, params.width, 3), np.float32)
        cnt, tm = A.Counters(), A.Timing()
        _check(lib().zrt_multi_render(self._h, C.byref(camera), C.byref(params), img.ctypes.data if img is not None else None,
                                      C.byref(cnt), C.byref(tm)))
        return img, cnt, tm

    def launch_count(self):
        return int(lib().zrt_multi_launch_count(self._h))

    def reload(self, built_scene_or_desc):
        """Replace the scene on every local device (flatten + H2D again); the communicator is kept."""
        desc = getattr(built_scene_or_desc, "desc", built_scene_or_desc)
        _check(lib().zrt_multi_reload(self._h, C.byref(desc)))


def selftest(device=0):
    """-> number of bit mismatches between the kernels' exact-division fast paths and IEEE division (must be 0)"""
    n = C.c_uint64()
    _check(lib().zrt_selftest(device, C.byref(n)))
    return int(n.value)


def measure_peaks(device=0):
    out = (C.c_double * 5)()
    _check(lib().zrt_measure_peaks(device, out, 5))
    return {"fp32_nofma_ops": out[0], "ffma_instr": out[1], "l2_read_gbs": out[2], "hbm_read_gbs": out[3],
            "sm_max_mhz": out[4]}
