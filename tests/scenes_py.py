"""Independent Python restatement of the reference's scene builders (scenes.zig:26-277),
obj_reader.zig:21-198 and png_image.zig:19-94, used by the tests to feed the SAME scene description to
the oracle and to libzrt, and to cross-check the C++ host mirror.  Test infrastructure."""
import gzip
import os

import numpy as np
from PIL import Image

from oracle import zro_py
from zraytrace_b200.scene import SceneBuilder

ASSETS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "assets")

# image.zig:14-20
SILVER = (0.752, 0.752, 0.752)
GREEN = (0.01, 1.0, 0.01)
BLUE = (0.01, 0.01, 1.0)


def read_png_bottom_up(name):
    """png_image.zig:19-94: 8-bit RGB/RGBA, rows flipped so that row 0 is the bottom scanline."""
    im = np.array(Image.open(os.path.join(ASSETS, "images", name)))
    assert im.dtype == np.uint8 and im.shape[2] in (3, 4)
    return np.ascontiguousarray(im[::-1])


def read_obj(name):
    """obj_reader.zig:114-198: `v` and `f` lines, 1-based indices, v/vt/vn suffixes ignored, faces of
    3..6 vertices fan-triangulated as (0,1,2),(2,3,0),(3,4,0),(4,5,0) (obj_reader.zig:64-111)."""
    path = os.path.join(ASSETS, "models", name + ".obj.gz")
    verts, tris = [], []
    with gzip.open(path, "rt") as f:
        for line in f:
            line = line.rstrip("\n").rstrip("\r")
            if len(line) < 2:
                continue
            if line[0] == "v" and line[1] == " ":
                p = line.split()
                verts.append((float(p[1]), float(p[2]), float(p[3])))
            elif line[0] == "f" and line[1] == " ":
                idx = [int(tok.split("/")[0]) - 1 for tok in line.split()[1:]]
                assert 3 <= len(idx) <= 6
                tris.append((idx[0], idx[1], idx[2]))
                for k in range(3, len(idx)):
                    tris.append((idx[k - 1], idx[k], idx[0]))
    v = np.array(verts, dtype=np.float64).astype(np.float32)
    t = np.array(tris, dtype=np.int64)
    return v[t]  # [N][3][3]


def _ground(b, top):
    green = b.lambertian(b.color_texture(*GREEN))
    radius = np.float32(100.0)
    center = (np.float32(1.66445508e-01), np.float32(top) - radius, np.float32(7.37018966e+00))
    b.sphere(center, 100.0, green)


def _cam(look_from, aspect=1.0):
    return zro_py.camera_init(look_from, (0.0, 0.0, 1.0), (0.0, 1.0, 0.0), 45.0, aspect)


def three_balls():
    """scenes.zig:54-100 (scene index 1): the 7-spheres showcase."""
    b = SceneBuilder()
    mirror = b.metal(b.color_texture(*SILVER))
    nitor = b.lambertian(b.image_texture(read_png_bottom_up("nitor-logo-25.png")))
    green = b.lambertian(b.color_texture(*GREEN))
    glass = b.dielectric(1.52)
    earth = b.metal(b.image_texture(read_png_bottom_up("earthmap.png")))
    b.sphere((1.0, -102.5, 4.0), 100.0, green)
    b.sphere((0.0, 0.0, 8.0), 2.0, nitor)
    b.sphere((-3.0, -1.5, 3.0), 1.0, mirror)
    b.sphere((3.0, -1.0, 4.0), 1.5, earth)
    b.sphere((-1.0, -1.0, 2.0), 0.7, glass)
    b.sphere((0.85, -0.7, 1.5), 0.9, glass)
    b.sphere((0.85, -0.7, 1.5), -0.8, glass)
    return b.build(), _cam((0.0, 0.0, -7.0))


def man_and_ball():
    """scenes.zig:26-52 (scene 0)"""
    b = SceneBuilder()
    blue = b.metal(b.color_texture(*BLUE))
    tris = read_obj("Man")
    _ground(b, -2.33)
    b.triangles(tris, blue)
    return b.build(), _cam((0.0, 0.0, -30.0))


def bunny_and_ball(dielectric=False):
    """scenes.zig:102-128 (scene 2); BASELINE config 3 swaps the bunny material to Dielectric(1.52)."""
    b = SceneBuilder()
    mat = b.dielectric(1.52) if dielectric else b.metal(b.color_texture(*SILVER))
    tris = read_obj("bunny")
    _ground(b, -0.33)
    b.triangles(tris, mat)
    return b.build(), _cam((0.0, 0.0, -0.5))


def teapot_and_ball():
    """scenes.zig:206-232 (scene 3)"""
    b = SceneBuilder()
    blue = b.metal(b.color_texture(*BLUE))
    tris = read_obj("teapot")
    _ground(b, -2.33)
    b.triangles(tris, blue)
    return b.build(), _cam((0.0, 0.0, -10.0))


def teapot_and_ball_circle():
    """scenes.zig:130-204 (scene 4)"""
    b = SceneBuilder()
    blue = b.metal(b.color_texture(*BLUE))
    silver = b.metal(b.color_texture(*SILVER))
    purple = b.lambertian(b.image_texture(read_png_bottom_up("earthmap.png")))
    tris = read_obj("teapot")
    b.sphere((0.0, 0.0, 6.0), -2.0, silver)
    b.sphere((3.0, -1.0, 4.0), 1.0, purple)
    _ground(b, -2.33)
    b.triangles(tris, blue)
    return b.build(), _cam((-8.0, 0.0, -10.0))


def small_test_scene():
    """raytrace.zig:214-239 "Render something": 3 spheres, list mode."""
    b = SceneBuilder()
    gold = b.metal(b.color_texture(1.0, 0.843, 0.0))
    green = b.lambertian(b.color_texture(*GREEN))
    purple = b.lambertian(b.color_texture(0.5, 0.0, 0.5))
    b.sphere((0.0, 0.0, 6.0), 2.0, gold)
    b.sphere((3.0, 1.0, 4.0), 1.0, purple)
    b.sphere((1.0, 102.5, 4.0), 100.0, green)
    return b.build(), _cam((0.0, 0.0, -7.0))
