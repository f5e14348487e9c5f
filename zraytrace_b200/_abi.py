"""ctypes mirror of include/zrt.h (the C ABI of libzrt).  Field order and types must match the header."""
import ctypes as C

ZRT_OK, ZRT_ERR_INVALID, ZRT_ERR_NO_DEVICE, ZRT_ERR_CUDA, ZRT_ERR_OOM, ZRT_ERR_IO, ZRT_ERR_NCCL = 0, -1, -2, -3, -4, -5, -6
ZRT_COMM_ID_BYTES = 128
ZRT_SURFACE_SPHERE, ZRT_SURFACE_TRIANGLE = 0, 1
ZRT_MATERIAL_LAMBERTIAN, ZRT_MATERIAL_METAL, ZRT_MATERIAL_DIELECTRIC = 0, 1, 2
ZRT_TEXTURE_COLOR, ZRT_TEXTURE_IMAGE = 0, 1
ZRT_XLIMIT_HEIGHT, ZRT_XLIMIT_WIDTH = 0, 1
ZRT_FLAG_RAW_SUM, ZRT_FLAG_BVH_REFERENCE, ZRT_FLAG_KERNEL_THREAD, ZRT_FLAG_KERNEL_SORTED, ZRT_FLAG_KERNEL_WARP = 1, 2, 4, 8, 16
ZRT_FLAG_RUSSIAN_ROULETTE, ZRT_FLAG_SAMPLER_HALTON, ZRT_FLAG_KERNEL_X2, ZRT_FLAG_KERNEL_POOL = 32, 64, 128, 256
ZRT_NO_HIT = 0xFFFFFFFF
ZRT_FEATURE_EXPERIMENTS = 1


class Vec3(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]

    def __init__(self, x=0.0, y=0.0, z=0.0):
        super().__init__(x, y, z)

    def tuple(self):
        return (self.x, self.y, self.z)


class Sphere(C.Structure):
    _fields_ = [("center", Vec3), ("radius", C.c_float), ("material", C.c_uint32)]


class Triangle(C.Structure):
    _fields_ = [("a", Vec3), ("b", Vec3), ("c", Vec3), ("material", C.c_uint32)]


class Surface(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("index", C.c_uint32)]


class Material(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("texture", C.c_uint32), ("index_of_refraction", C.c_float)]


class Texture(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("r", C.c_float), ("g", C.c_float), ("b", C.c_float),
                ("width", C.c_uint32), ("height", C.c_uint32), ("channels", C.c_uint32),
                ("pixels", C.c_void_p), ("u_offset", C.c_float), ("v_offset", C.c_float)]


class SceneDesc(C.Structure):
    _fields_ = [("n_surfaces", C.c_uint32), ("surfaces", C.POINTER(Surface)),
                ("n_spheres", C.c_uint32), ("spheres", C.POINTER(Sphere)),
                ("n_triangles", C.c_uint32), ("triangles", C.POINTER(Triangle)),
                ("n_materials", C.c_uint32), ("materials", C.POINTER(Material)),
                ("n_textures", C.c_uint32), ("textures", C.POINTER(Texture))]


class Camera(C.Structure):
    _fields_ = [("origin", Vec3), ("lower_left_corner", Vec3), ("horizontal", Vec3), ("vertical", Vec3)]


class Params(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("samples_per_pixel", C.c_uint32),
                ("max_depth", C.c_uint32), ("bounded_volume_hierarchy", C.c_uint32), ("x_limit", C.c_uint32),
                ("seed", C.c_uint64), ("sample_begin", C.c_uint32), ("sample_end", C.c_uint32),
                ("flags", C.c_uint32), ("sample_chunks", C.c_uint32)]


class Counters(C.Structure):
    _fields_ = [("recursion_depth_hits", C.c_uint64), ("reflections", C.c_uint64),
                ("background_hits", C.c_uint64), ("pixels_processed", C.c_uint64),
                ("samples_processed", C.c_uint64), ("rays_processed", C.c_uint64)]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class Timing(C.Structure):
    _fields_ = [("prepare_ms", C.c_float), ("kernel_ms", C.c_float), ("resolve_ms", C.c_float),
                ("total_ms", C.c_float), ("launches", C.c_uint32), ("bvh_nodes", C.c_uint32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class TraceStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("rays", "samples", "node_visits", "triangle_tests", "sphere_tests",
                                          "texture_lookups")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class BvhInfo(C.Structure):
    _fields_ = [("nodes", C.c_uint32), ("leaves", C.c_uint32), ("max_depth", C.c_uint32),
                ("pruned_surfaces", C.c_uint32), ("reference_nodes", C.c_uint32),
                ("reference_max_depth", C.c_uint32)]


def make_params(width, height, spp, max_depth, bvh=True, x_limit=ZRT_XLIMIT_HEIGHT, seed=42,
                sample_begin=0, sample_end=0, flags=0, sample_chunks=0):
    return Params(width, height, spp, max_depth, 1 if bvh else 0, x_limit, seed, sample_begin, sample_end, flags,
                  sample_chunks)
