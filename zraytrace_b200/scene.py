"""Host-side scene description builder: the Python face of `ArrayList(Surface)` + materials + textures
that the reference hands to raytrace.render() (raytrace.zig:136-138).  Produces the POD `zrt_scene_desc`
of include/zrt.h; numpy arrays own the memory."""
import ctypes as C

import numpy as np

from . import _abi as A

_SPHERE_DT = np.dtype([("center", np.float32, 3), ("radius", np.float32), ("material", np.uint32)])
_TRI_DT = np.dtype([("a", np.float32, 3), ("b", np.float32, 3), ("c", np.float32, 3), ("material", np.uint32)])
_SURF_DT = np.dtype([("kind", np.uint32), ("index", np.uint32)])
assert _SPHERE_DT.itemsize == C.sizeof(A.Sphere) and _TRI_DT.itemsize == C.sizeof(A.Triangle)


class SceneBuilder:
    """Append-only scene, mirroring how scenes.zig builds its surface list."""

    def __init__(self):
        self._spheres, self._tris, self._surfaces = [], [], []
        self._materials, self._textures, self._keep = [], [], []

    # -- textures (texture.zig:7-16) --
    def color_texture(self, r, g, b):
        self._textures.append(A.Texture(A.ZRT_TEXTURE_COLOR, r, g, b, 0, 0, 0, None, 0.0, 0.0))
        return len(self._textures) - 1

    def image_texture(self, pixels_bottom_up, u_offset=0.19, v_offset=0.1):
        """pixels_bottom_up: uint8 [H][W][3|4], row 0 = bottom scanline (png_image.zig:82-87 flip)."""
        px = np.ascontiguousarray(pixels_bottom_up, dtype=np.uint8)
        h, w, ch = px.shape
        assert ch in (3, 4)
        self._keep.append(px)
        self._textures.append(A.Texture(A.ZRT_TEXTURE_IMAGE, 0, 0, 0, w, h, ch, px.ctypes.data, u_offset, v_offset))
        return len(self._textures) - 1

    # -- materials (material.zig:16-37) --
    def lambertian(self, texture):
        self._materials.append(A.Material(A.ZRT_MATERIAL_LAMBERTIAN, texture, 0.0))
        return len(self._materials) - 1

    def metal(self, texture):
        self._materials.append(A.Material(A.ZRT_MATERIAL_METAL, texture, 0.0))
        return len(self._materials) - 1

    def dielectric(self, index_of_refraction):
        self._materials.append(A.Material(A.ZRT_MATERIAL_DIELECTRIC, 0, index_of_refraction))
        return len(self._materials) - 1

    # -- surfaces --
    def sphere(self, center, radius, material):
        self._surfaces.append((A.ZRT_SURFACE_SPHERE, len(self._spheres)))
        self._spheres.append((tuple(center), radius, material))
        return len(self._surfaces) - 1

    def triangle(self, a, b, c, material):
        self._surfaces.append((A.ZRT_SURFACE_TRIANGLE, len(self._tris)))
        self._tris.append((tuple(a), tuple(b), tuple(c), material))
        return len(self._surfaces) - 1

    def triangles(self, abc, material):
        """abc: float32 [N][3][3]"""
        abc = np.asarray(abc, dtype=np.float32)
        base = len(self._tris)
        for t in abc:
            self._tris.append((tuple(t[0]), tuple(t[1]), tuple(t[2]), material))
        self._surfaces.extend((A.ZRT_SURFACE_TRIANGLE, base + i) for i in range(len(abc)))

    def build(self):
        return BuiltScene(self)


class BuiltScene:
    """Frozen arrays + the ctypes zrt_scene_desc pointing into them."""

    def __init__(self, b):
        self.spheres = np.array(b._spheres, dtype=_SPHERE_DT) if b._spheres else np.zeros(0, _SPHERE_DT)
        self.triangles = np.array(b._tris, dtype=_TRI_DT) if b._tris else np.zeros(0, _TRI_DT)
        self.surfaces = np.array(b._surfaces, dtype=_SURF_DT) if b._surfaces else np.zeros(0, _SURF_DT)
        self.materials = (A.Material * max(1, len(b._materials)))(*b._materials)
        self.textures = (A.Texture * max(1, len(b._textures)))(*b._textures)
        self._keep = list(b._keep)
        self.n_materials, self.n_textures = len(b._materials), len(b._textures)
        self.desc = A.SceneDesc(
            len(self.surfaces), self.surfaces.ctypes.data_as(C.POINTER(A.Surface)),
            len(self.spheres), self.spheres.ctypes.data_as(C.POINTER(A.Sphere)),
            len(self.triangles), self.triangles.ctypes.data_as(C.POINTER(A.Triangle)),
            self.n_materials, self.materials, self.n_textures, self.textures)

    @property
    def n_surfaces(self):
        return len(self.surfaces)
