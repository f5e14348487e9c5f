"""Python face of the C++ host mirror (include/zrt_host.h): the reference's scenes.zig / camera.zig /
obj_reader.zig / png_image.zig API, implemented in C++ inside libzrt.so and bound here with ctypes."""
import ctypes as C
import os

import numpy as np

from . import _abi as A
from .lib import ZrtError, lib

ASSETS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "assets")
VARIANT_REFERENCE, VARIANT_BUNNY_GLASS, VARIANT_GOAT_SUBSTITUTE = 0, 1, 2
# scenes.zig:267-277
SCENE_MAN, SCENE_THREE_BALLS, SCENE_BUNNY, SCENE_TEAPOT, SCENE_TEAPOT_CIRCLE, SCENE_GOAT = range(6)

_ready = False


def _L():
    global _ready
    L = lib()
    if not _ready:
        P = C.POINTER
        f3 = P(C.c_float)
        L.zrt_host_camera_init.argtypes = [f3, f3, f3, C.c_float, C.c_float, P(A.Camera)]
        L.zrt_host_read_obj.argtypes = [C.c_char_p, C.c_uint32, P(P(A.Triangle)), P(C.c_uint32)]
        L.zrt_host_png_read.argtypes = [C.c_char_p, P(P(C.c_uint8)), P(C.c_uint32), P(C.c_uint32), P(C.c_uint32)]
        L.zrt_host_png_write.argtypes = [C.c_char_p, C.c_void_p, C.c_uint32, C.c_uint32]
        L.zrt_host_png_write_rgb8.argtypes = [C.c_char_p, C.c_void_p, C.c_uint32, C.c_uint32]
        L.zrt_host_free.argtypes = [C.c_void_p]
        L.zrt_host_free.restype = None
        L.zrt_host_scene_load.argtypes = [C.c_uint32, C.c_char_p, C.c_uint32, C.c_float, P(C.c_void_p)]
        L.zrt_host_scene_desc.argtypes = [C.c_void_p]
        L.zrt_host_scene_desc.restype = P(A.SceneDesc)
        L.zrt_host_scene_camera.argtypes = [C.c_void_p]
        L.zrt_host_scene_camera.restype = P(A.Camera)
        L.zrt_host_scene_free.argtypes = [C.c_void_p]
        L.zrt_host_scene_free.restype = None
        L.zrt_host_scene_pin.argtypes = [C.c_void_p]
        L.zrt_host_render_scene.argtypes = [C.c_uint32, C.c_char_p, C.c_uint32, P(A.Params), C.c_int, C.c_void_p,
                                            P(A.Counters), P(A.Timing)]
        _ready = True
    return L


def _check(rc, what):
    if rc != 0:
        raise ZrtError(rc, f"{what}: {lib().zrt_last_error().decode()}")


def camera_init(look_from, look_at, vup, vfov, aspect_ratio):
    """Camera.init (camera.zig:17-35)"""
    cam = A.Camera()
    v = lambda t: (C.c_float * 3)(*t)
    _check(_L().zrt_host_camera_init(v(look_from), v(look_at), v(vup), vfov, aspect_ratio, C.byref(cam)), "camera_init")
    return cam


def read_obj(path, material=0):
    """ObjReader.readObjFile (obj_reader.zig:114-198) -> float32 [N][3][3]"""
    tris, n = C.POINTER(A.Triangle)(), C.c_uint32()
    _check(_L().zrt_host_read_obj(path.encode(), material, C.byref(tris), C.byref(n)), f"read_obj {path}")
    raw = np.ctypeslib.as_array(C.cast(tris, C.POINTER(C.c_float)), shape=(n.value, 10)).copy()
    _L().zrt_host_free(tris)
    return np.ascontiguousarray(raw[:, :9].reshape(-1, 3, 3))


def png_read(path):
    """png_image.readFile (png_image.zig:19-94) -> uint8 [H][W][C], row 0 = bottom scanline"""
    px, w, h, ch = C.POINTER(C.c_uint8)(), C.c_uint32(), C.c_uint32(), C.c_uint32()
    _check(_L().zrt_host_png_read(path.encode(), C.byref(px), C.byref(w), C.byref(h), C.byref(ch)), f"png_read {path}")
    out = np.ctypeslib.as_array(px, shape=(h.value, w.value, ch.value)).copy()
    _L().zrt_host_free(px)
    return out


def png_write(path, image):
    """png_image.writeFile (png_image.zig:96-148): image float32 [H][W][3], row 0 = bottom"""
    img = np.ascontiguousarray(image, np.float32)
    _check(_L().zrt_host_png_write(path.encode(), img.ctypes.data, img.shape[1], img.shape[0]), f"png_write {path}")


def png_write_rgb8(path, image8_top_down):
    """8-bit top-down image (Scene.render_rgb8) -> PNG, no arithmetic"""
    img = np.ascontiguousarray(image8_top_down, np.uint8)
    _check(_L().zrt_host_png_write_rgb8(path.encode(), img.ctypes.data, img.shape[1], img.shape[0]), f"png_write_rgb8 {path}")


class HostScene:
    """One of the reference's scenes (scenes.zig:26-260) built by the C++ host: `.desc` and `.camera` are
    what raytrace.render() receives."""

    def __init__(self, scene_index, assets_dir=ASSETS, variant=VARIANT_REFERENCE, aspect_ratio=1.0):
        self._h = C.c_void_p()
        _check(_L().zrt_host_scene_load(scene_index, assets_dir.encode(), variant, aspect_ratio, C.byref(self._h)),
               f"scene {scene_index}")
        self.desc = _L().zrt_host_scene_desc(self._h).contents
        self.camera = _L().zrt_host_scene_camera(self._h).contents
        self.n_surfaces = self.desc.n_surfaces

    def close(self):
        if self._h:
            _L().zrt_host_scene_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def pin(self):
        """Page-lock the texel arrays (zrt_host_scene_pin): scene uploads become one DMA per texture."""
        _check(_L().zrt_host_scene_pin(self._h), "scene pin")
        return self

    def upload_bytes(self):
        """host->device bytes zrt_scene_create moves for this scene (texels + primitives + materials)."""
        d = self.desc
        n = d.n_spheres * 32 + d.n_triangles * (48 + 8) + d.n_surfaces * 4 + d.n_materials * 64
        for i in range(d.n_textures):
            t = d.textures[i]
            if t.kind == A.ZRT_TEXTURE_IMAGE:
                n += t.width * t.height * t.channels
        return n


def render_scene(scene_index, params, assets_dir=ASSETS, variant=VARIANT_REFERENCE, device=0):
    """scenes.render_scene (scenes.zig:267-277) -> (image, Counters, Timing)"""
    img = np.empty((params.height, params.width, 3), np.float32)
    cnt, tm = A.Counters(), A.Timing()
    _check(_L().zrt_host_render_scene(scene_index, assets_dir.encode(), variant, C.byref(params), device,
                                      img.ctypes.data, C.byref(cnt), C.byref(tm)), f"render_scene {scene_index}")
    return img, cnt, tm
