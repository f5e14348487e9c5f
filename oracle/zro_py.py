"""ctypes loader for the CPU oracle (oracle/libzro.so).  TEST INFRASTRUCTURE: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg import this module."""
import ctypes as C
import os
import subprocess

import numpy as np

from zraytrace_b200 import _abi as A

HERE = os.path.dirname(os.path.abspath(__file__))
RNG_REF, RNG_CTR = 0, 1
TRAVERSAL_REF, TRAVERSAL_TIGHT = 0, 1
MATH_SPEC, MATH_LIBM = 0, 1


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "sphere_tests", "sphere_sqrt", "sphere_accepts", "triangle_tests", "triangle_accepts",
        "box_tests", "box_passes", "lambertian", "metal", "metal_absorbed", "dielectric_reflect",
        "dielectric_refract", "texture_lookups", "background", "bvh_nodes", "bvh_max_depth")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


def build(force=False):
    so = os.path.join(HERE, "libzro.so")
    srcs = [os.path.join(HERE, f) for f in ("zro.cpp", "zro.h", "zro_math.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", HERE, "-s"])
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        f3 = C.POINTER(C.c_float)
        L.zro_render.argtypes = [C.POINTER(A.SceneDesc), C.POINTER(A.Camera), C.POINTER(A.Params), C.c_int, C.c_int,
                                 C.c_int, C.c_int, C.c_void_p, C.POINTER(A.Counters), C.POINTER(Stats)]
        L.zro_primary_hits.argtypes = [C.POINTER(A.SceneDesc), C.POINTER(A.Camera), C.POINTER(A.Params), C.c_int,
                                       C.c_int, C.c_void_p, C.c_void_p]
        L.zro_primary_hits_mt.argtypes = [C.POINTER(A.SceneDesc), C.POINTER(A.Camera), C.POINTER(A.Params), C.c_int,
                                          C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.zro_bvh_order.argtypes = [C.POINTER(A.SceneDesc), C.c_void_p, C.c_void_p, C.POINTER(Stats)]
        L.zro_camera_init.argtypes = [f3, f3, f3, C.c_float, C.c_float, C.POINTER(A.Camera)]
        L.zro_vec3_dot.restype = C.c_float
        L.zro_aabb_surface_area.restype = C.c_float
        L.zro_aabb_volume.restype = C.c_float
        L.zro_bvh_random_test.restype = C.c_uint64
        L.zro_bvh_random_test.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64, C.c_int]
        L.zro_quantize.restype = C.c_uint8
        L.zro_quantize.argtypes = [C.c_float]
        L.zro_rng_ctr.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_void_p]
        L.zro_sample.argtypes = [C.c_int, C.c_uint64, f3]
        L.zro_math_eval.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        _lib = L
    return _lib


def fv(*xs):
    return (C.c_float * len(xs))(*xs)


def camera_init(look_from, look_at, vup, vfov, aspect):
    cam = A.Camera()
    lib().zro_camera_init(fv(*look_from), fv(*look_at), fv(*vup), vfov, aspect, C.byref(cam))
    return cam


def render(scene, camera, params, rng=RNG_CTR, traversal=TRAVERSAL_REF, math=MATH_SPEC, threads=1):
    """-> (image float32 [H][W][3] row 0 = bottom, Counters, Stats)"""
    img = np.zeros((params.height, params.width, 3), np.float32)
    cnt, st = A.Counters(), Stats()
    rc = lib().zro_render(C.byref(scene.desc), C.byref(camera), C.byref(params), rng, traversal, math, threads,
                          img.ctypes.data, C.byref(cnt), C.byref(st))
    if rc != 0:
        raise RuntimeError(f"zro_render failed: {rc}")
    return img, cnt, st


def primary_hits(scene, camera, params, jitter=0, traversal=TRAVERSAL_REF, threads=1):
    """First rayColor iteration per pixel through the oracle's pointer tree.  threads > 1 only spreads the rows
    over host cores (pixels are independent); TRAVERSAL_TIGHT visits ~50x fewer nodes than the literal aabb.zig
    test and returns the same hits."""
    ids = np.zeros((params.height, params.width), np.uint32)
    t = np.zeros((params.height, params.width), np.float32)
    rc = lib().zro_primary_hits_mt(C.byref(scene.desc), C.byref(camera), C.byref(params), jitter, traversal, threads,
                                   ids.ctypes.data, t.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"zro_primary_hits failed: {rc}")
    return ids, t


def bvh_order(scene):
    n = scene.n_surfaces
    order = np.zeros(n, np.uint32)
    vis = np.zeros(n, np.uint8)
    st = Stats()
    rc = lib().zro_bvh_order(C.byref(scene.desc), order.ctypes.data, vis.ctypes.data, C.byref(st))
    if rc != 0:
        raise RuntimeError(f"zro_bvh_order failed: {rc}")
    return order, vis.astype(bool), st


def math_eval(which, x, y=None):
    x = np.ascontiguousarray(x, np.float32)
    y = np.ascontiguousarray(y if y is not None else x, np.float32)
    o, o2 = np.empty_like(x), np.empty_like(x)
    lib().zro_math_eval(which, x.ctypes.data, y.ctypes.data, o.ctypes.data, o2.ctypes.data, x.size)
    return o, o2


def rng_ctr(pixel, sample, bounce, seed):
    out = np.zeros(4, np.uint32)
    lib().zro_rng_ctr(pixel, sample, bounce, seed, out.ctypes.data)
    return out
