"""CPU tests of the C++ host mirror (scenes / camera / OBJ / PNG) against independent Python restatements,
and of the C-ABI surface: the library loads, exports every symbol the headers declare, and refuses to
compute without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
from PIL import Image

from oracle import zro_py
from tests import scenes_py
from zraytrace_b200 import _abi as A
from zraytrace_b200 import host
from zraytrace_b200 import lib as Z

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    L = Z.lib()
    declared = set()
    for h in ("zrt.h", "zrt_host.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        declared |= set(re.findall(r"\b(zrt_[a-z0-9_]+)\s*\(", src))
    assert len(declared) >= 18
    for name in sorted(declared):
        assert hasattr(L, name), f"libzrt.so does not export {name}"


def test_abi_struct_sizes_match_header():
    # sizes computed from include/zrt.h by a C compiler would be: see zrt.h; keep ctypes in sync
    assert C.sizeof(A.Sphere) == 20 and C.sizeof(A.Triangle) == 40 and C.sizeof(A.Surface) == 8
    assert C.sizeof(A.Material) == 12 and C.sizeof(A.Texture) == 48 and C.sizeof(A.Camera) == 48
    assert C.sizeof(A.Params) == 48 and C.sizeof(A.Counters) == 48 and C.sizeof(A.SceneDesc) == 80


def test_no_cpu_fallback_without_device():
    if Z.device_count() > 0:
        pytest.skip("a device is visible")
    sc, cam = scenes_py.three_balls()
    with pytest.raises(Z.ZrtError) as e:
        Z.Scene(sc, device=0)
    assert e.value.code == A.ZRT_ERR_NO_DEVICE
    with Z.Scene(sc, device=-1) as hs:
        with pytest.raises(Z.ZrtError) as e:
            hs.render(cam, A.make_params(8, 8, 1, 2))
        assert e.value.code == A.ZRT_ERR_NO_DEVICE


def test_host_alloc_needs_the_driver_and_checks_arguments():
    """zrt_pinned_alloc page-locks through the CUDA driver: without a device it fails like every compute call, and
    it never hands back a pointer on an error."""
    p = C.c_void_p(1)
    assert Z.lib().zrt_pinned_alloc(0, C.byref(p)) == A.ZRT_ERR_INVALID and not p.value
    assert Z.lib().zrt_pinned_alloc(64, None) == A.ZRT_ERR_INVALID
    Z.lib().zrt_pinned_free(None)
    if Z.device_count() == 0:
        p = C.c_void_p(1)
        assert Z.lib().zrt_pinned_alloc(64, C.byref(p)) == A.ZRT_ERR_NO_DEVICE and not p.value
        with pytest.raises(Z.ZrtError):
            Z.HostImage((4, 4, 3))


def test_host_scene_pin_without_a_device_leaves_the_scene_alone():
    hs = host.HostScene(host.SCENE_THREE_BALLS)
    before = [bytes(C.string_at(hs.desc.textures[i].pixels, 64)) for i in range(hs.desc.n_textures)
              if hs.desc.textures[i].kind == A.ZRT_TEXTURE_IMAGE]
    assert len(before) == 2
    if Z.device_count() == 0:
        with pytest.raises(Z.ZrtError) as e:
            hs.pin()
        assert e.value.code == A.ZRT_ERR_NO_DEVICE
    else:
        hs.pin().pin()  # idempotent
    after = [bytes(C.string_at(hs.desc.textures[i].pixels, 64)) for i in range(hs.desc.n_textures)
             if hs.desc.textures[i].kind == A.ZRT_TEXTURE_IMAGE]
    assert before == after
    hs.close()


def test_invalid_scene_rejected():
    from zraytrace_b200.scene import SceneBuilder
    b = SceneBuilder()
    b.sphere((0, 0, 0), 1.0, 3)  # material index out of range
    with pytest.raises(Z.ZrtError) as e:
        Z.Scene(b.build(), device=-1)
    assert e.value.code == A.ZRT_ERR_INVALID


def test_camera_init_matches_oracle():  # camera.zig:17-35
    for args in (((0, 0, -7), (0, 0, 1), (0, 1, 0), 45.0, 1.0), ((1, 0, 0), (0, 0, 1), (0, 1, 0), 45.0, 1.0),
                 ((-8, 0, -10), (0, 0, 1), (0, 1, 0), 45.0, 16 / 9)):
        a, b = host.camera_init(*args), zro_py.camera_init(*args)
        assert bytes(a) == bytes(b)
    assert host.camera_init((1, 0, 0), (0, 0, 1), (0, 1, 0), 45.0, 1.0).origin.tuple() == (1.0, 0.0, 0.0)  # camera.zig:58-66


@pytest.mark.parametrize("name", ["Man", "bunny", "teapot"])
def test_obj_reader_matches_python_parser(name):  # obj_reader.zig:201-226
    got = host.read_obj(os.path.join(host.ASSETS, "models", name + ".obj.gz"))
    want = scenes_py.read_obj(name)
    assert got.shape == want.shape and np.array_equal(got, want)
    assert len(got) == {"Man": 3933, "bunny": 4968, "teapot": 6320}[name]


def _obj(tmp_path, text, name="m.obj"):
    path = str(tmp_path / name)
    with open(path, "wb") as f:
        f.write(text.encode() if isinstance(text, str) else text)
    return path


def test_obj_reader_grammar_and_errors(tmp_path):
    """obj_reader.zig:21-198 line by line: CRLF, `v/vt/vn` and `v//vn` faces, fans of 3..6 vertices, numbers the fast
    parser hands to strtof, an unterminated last line (never returned by readUntilDelimiterAlloc), and every way the
    reference's loop fails (`try`): bad numbers, 2 or 7 face vertices, an index beyond the vertices read so far."""
    ok = ("# comment\r\nv 0 0 0\r\nv +1 0 0\nv 0 1e0 0 9 9\nv  0   0  0x1p0\nv 1 1 1\nv 2 2 2\nvn 0 0 1\nvt 0.5 0.5\n"
          "f 1 2 3\nf 1/1/1 2/2/2 3/3/3 4/4/4\nf 1//1 2//2 3//3 4//4 5//5\nf 1/7 2/7 3/7 4/7 5/7 6/7\n"
          "g ignored\nf 3 2 1")  # the last face has no newline: dropped
    t = host.read_obj(_obj(tmp_path, ok))
    v = np.array([(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 1), (2, 2, 2)], np.float32)
    fans = [(0, 1, 2), (0, 1, 2), (2, 3, 0), (0, 1, 2), (2, 3, 0), (3, 4, 0), (0, 1, 2), (2, 3, 0), (3, 4, 0), (4, 5, 0)]
    assert np.array_equal(t, v[np.array(fans)])
    assert len(host.read_obj(_obj(tmp_path, ""))) == 0
    bad = ["v 0 0\n", "v 0 0 zero\n", "v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2\n", "v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3 1 2 3 1\n",
           "v 0 0 0\nv 1 0 0\nf 1 2 3\nv 0 1 0\n",  # forward reference: only two vertices had been read
           "v 0 0 0\nv 1 0 0\nv 0 1 0\nf 0 1 2\n", "v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 -3\n",
           "v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1/x 2 3\n", "vn 0 0\n", "v 0 0 0 " + " " * 20001 + "\n"]
    for text in bad:
        with pytest.raises(Z.ZrtError) as e:
            host.read_obj(_obj(tmp_path, text))
        assert e.value.code == A.ZRT_ERR_INVALID, text[:40]
    with pytest.raises(Z.ZrtError) as e:
        host.read_obj(str(tmp_path / "missing.obj"))
    assert e.value.code == A.ZRT_ERR_IO


def test_obj_reader_parallel_chunks_keep_file_order(tmp_path):
    """A file large enough to be parsed by several chunks on separate threads: vertices and faces interleaved, faces
    that refer far back, and one that refers one vertex too far ahead (an error only because of where it stands)."""
    rng = np.random.default_rng(7)
    n = 60000
    verts = rng.standard_normal((n, 3)).astype(np.float32)
    lines, tris = [], []
    for i in range(n):
        lines.append("v %r %r %r" % tuple(float(x) for x in verts[i]))
        if i >= 2:
            a, b = int(rng.integers(0, i + 1)), int(rng.integers(0, i + 1))
            lines.append(f"f {i + 1} {a + 1}/1 {b + 1}//2")
            tris.append((i, a, b))
    text = "\n".join(lines) + "\n"
    assert len(text) > 4 * (256 << 10)
    got = host.read_obj(_obj(tmp_path, text))
    assert np.array_equal(got, verts[np.array(tris)])
    k = len(lines) * 3 // 4
    lines.insert(k, f"f 1 2 {sum(1 for x in lines[:k] if x[0] == 'v') + 1}")
    with pytest.raises(Z.ZrtError) as e:
        host.read_obj(_obj(tmp_path, "\n".join(lines) + "\n"))
    assert e.value.code == A.ZRT_ERR_INVALID


@pytest.mark.parametrize("name", ["earthmap.png", "nitor-logo-25.png"])
def test_png_reader_matches_pil(name):  # png_image.zig:19-94 incl. the row flip
    got = host.png_read(os.path.join(host.ASSETS, "images", name))
    want = scenes_py.read_png_bottom_up(name)
    assert got.shape == want.shape and np.array_equal(got, want)


def test_png_writer_quantisation_and_flip(tmp_path):  # png_image.zig:96-148
    rng = np.random.default_rng(3)
    img = rng.random((7, 5, 3)).astype(np.float32) * 1.2 - 0.1
    path = str(tmp_path / "o.png")
    host.png_write(path, img)
    back = np.array(Image.open(path))
    want = np.array([[[zro_py.lib().zro_quantize(float(c)) for c in px] for px in row] for row in img], np.uint8)[::-1]
    assert back.shape == (7, 5, 3) and np.array_equal(back, want)
    assert np.array_equal(host.png_read(path)[..., :3], want[::-1])


def _desc_arrays(d):
    def arr(ptr, n, ctype_size):
        return bytes(C.string_at(ptr, n * ctype_size)) if n else b""
    tex = []
    for i in range(d.n_textures):
        t = d.textures[i]
        px = bytes(C.string_at(t.pixels, t.width * t.height * t.channels)) if t.kind == A.ZRT_TEXTURE_IMAGE else b""
        tex.append((t.kind, t.r, t.g, t.b, t.width, t.height, t.channels, t.u_offset, t.v_offset, px))
    return (arr(d.surfaces, d.n_surfaces, 8), arr(d.spheres, d.n_spheres, 20), arr(d.triangles, d.n_triangles, 40),
            arr(d.materials, d.n_materials, 12), tex)


@pytest.mark.parametrize("index,variant,py", [
    (host.SCENE_MAN, 0, scenes_py.man_and_ball), (host.SCENE_THREE_BALLS, 0, scenes_py.three_balls),
    (host.SCENE_BUNNY, 0, scenes_py.bunny_and_ball),
    (host.SCENE_BUNNY, host.VARIANT_BUNNY_GLASS, lambda: scenes_py.bunny_and_ball(dielectric=True)),
    (host.SCENE_TEAPOT, 0, scenes_py.teapot_and_ball), (host.SCENE_TEAPOT_CIRCLE, 0, scenes_py.teapot_and_ball_circle)])
def test_scene_builders_match_python_restatement(index, variant, py):  # scenes.zig:26-260
    hs = host.HostScene(index, variant=variant)
    sc, cam = py()
    assert bytes(hs.camera) == bytes(cam)
    a, b = _desc_arrays(hs.desc), _desc_arrays(sc.desc)
    # surfaces, spheres and triangles are identical; material/texture tables may be ordered differently, so
    # compare what every surface resolves to
    assert a[0] == b[0] and hs.desc.n_spheres == sc.desc.n_spheres and hs.desc.n_triangles == sc.desc.n_triangles

    def resolved(d, tex, kind, i):
        m = d.materials[(d.spheres[i] if kind == 0 else d.triangles[i]).material]
        return (m.kind, m.index_of_refraction if m.kind == 2 else 0.0, tex[m.texture] if m.kind != 2 else None)
    for kind, n in ((0, sc.desc.n_spheres), (1, min(sc.desc.n_triangles, 50))):
        for i in range(n):
            assert resolved(hs.desc, a[4], kind, i) == resolved(sc.desc, b[4], kind, i)
    g = lambda d, i: bytes(d.spheres[i])[:16]
    assert all(g(hs.desc, i) == g(sc.desc, i) for i in range(sc.desc.n_spheres))
    t = lambda d, i: bytes(d.triangles[i])[:36]
    assert all(t(hs.desc, i) == t(sc.desc, i) for i in range(sc.desc.n_triangles))
    hs.close()


def test_goat_scene_reports_missing_model_and_substitute_builds():
    with pytest.raises(Z.ZrtError) as e:  # models/high_poly_goat.obj is absent from the reference checkout
        host.HostScene(host.SCENE_GOAT)
    assert e.value.code == A.ZRT_ERR_IO
    hs = host.HostScene(host.SCENE_GOAT, variant=host.VARIANT_GOAT_SUBSTITUTE, aspect_ratio=16 / 9)
    assert hs.desc.n_triangles == 3933 + 4968 * 64 and hs.desc.n_spheres == 1
    hs.close()


@pytest.mark.parametrize("name,fn", [("teapot", scenes_py.teapot_and_ball), ("man", scenes_py.man_and_ball),
                                     ("bunny", scenes_py.bunny_and_ball)])
def test_flattener_order_and_pruning_match_oracle_tree(name, fn):
    """The product's index-based rebuild of bvh.zig:62-185 must give the reference tree: same left-first DFS
    order of surfaces (tie-break keys) and same set of surfaces hidden under zero-thickness boxes (Q4)."""
    sc, _ = fn()
    o_order, o_vis, st = zro_py.bvh_order(sc)
    with Z.Scene(sc, device=-1) as hs:
        z_order, z_vis = hs.bvh_order()
        info = hs.bvh_info(A.ZRT_FLAG_BVH_REFERENCE)
        sah = hs.bvh_info()
    assert np.array_equal(o_order, z_order) and np.array_equal(o_vis, z_vis)
    assert info.reference_nodes == st.bvh_nodes and info.reference_max_depth == st.bvh_max_depth
    assert info.leaves == int(o_vis.sum()) and info.pruned_surfaces == int((~o_vis).sum())
    assert info.nodes == info.leaves - 1 and sah.nodes == info.nodes and sah.max_depth <= info.max_depth
    if name == "man":
        assert info.pruned_surfaces > 0  # Man.obj has axis-aligned flat triangles


def _tie_heavy_scene(seed, n_tris, n_spheres):
    """Surfaces whose box midpoints sit on a coarse lattice: every axis has long runs of exactly equal sort keys
    (duplicates, -0.0 next to +0.0, axis-aligned flat triangles), the case in which the order a stable sort leaves
    depends on the whole history of earlier sorts."""
    from zraytrace_b200.scene import SceneBuilder
    rng = np.random.default_rng(seed)
    b = SceneBuilder()
    m = b.lambertian(b.color_texture(0.5, 0.5, 0.5))
    cells = rng.integers(-3, 4, size=(n_tris, 3)).astype(np.float32)
    half = rng.choice(np.array([0.5, 1.0], np.float32), size=(n_tris, 3))
    flat = rng.random(n_tris) < 0.2
    tris = np.zeros((n_tris, 3, 3), np.float32)
    for i in range(n_tris):
        lo, hi = cells[i] - half[i], cells[i] + half[i]
        z_top = lo[2] if flat[i] else hi[2]  # flat: all three vertices share z, a zero-thickness box (Q4)
        if flat[i] and z_top == 0.0 and rng.random() < 0.5:
            lo[2] = z_top = np.float32(-0.0)  # midpoint -0.0: sorts as equal to +0.0
        tris[i] = ((lo[0], lo[1], lo[2]), (hi[0], lo[1], z_top), (lo[0], hi[1], z_top))
    b.triangles(tris, m)
    for _ in range(n_spheres):
        c = rng.integers(-3, 4, size=3).astype(np.float32)
        b.sphere(tuple(float(x) for x in c), float(rng.choice([0.25, 0.5, -0.5])), m)
    return b.build()


@pytest.mark.parametrize("seed,n_tris,n_spheres", [(1, 2500, 40), (2, 6000, 0), (3, 2100, 300), (4, 300, 20), (5, 11, 0),
                                                   (6, 9, 2), (7, 2040, 7)])  # > 10 surfaces: raytrace.zig:111-133
def test_presorted_tree_build_matches_oracle_on_tie_heavy_scenes(seed, n_tris, n_spheres):
    """From 2048 surfaces on the reference tree is rebuilt from five presorted index lists instead of by sorting
    inside the recursion (zrt_flatten.cpp), below that literally; either way slots and pruning are assigned where
    the leaves are created.  The oracle's pointer tree sorts like bvh.zig:71-120 and walks the tree for both."""
    sc = _tie_heavy_scene(seed, n_tris, n_spheres)
    o_order, o_vis, st = zro_py.bvh_order(sc)
    with Z.Scene(sc, device=-1) as hs:
        z_order, z_vis = hs.bvh_order()
        info = hs.bvh_info(A.ZRT_FLAG_BVH_REFERENCE)
    assert np.array_equal(o_order, z_order) and np.array_equal(o_vis, z_vis)
    assert info.reference_nodes == st.bvh_nodes and info.reference_max_depth == st.bvh_max_depth
    assert info.leaves == int(o_vis.sum()) and info.pruned_surfaces == int((~o_vis).sum())
    if n_tris >= 300:
        assert (~o_vis).sum() > 0


def test_presorted_tree_build_matches_literal_build_on_config4(monkeypatch):
    """317 952 + 3 933 triangles: the presorted build and the literal sort-in-the-recursion build of the same
    library (ZRT_BVH_BUILD=literal) give the same DFS order, pruning, node count and depth."""
    hsrc = host.HostScene(host.SCENE_GOAT, variant=host.VARIANT_GOAT_SUBSTITUTE, aspect_ratio=16 / 9)
    got = []
    for how in ("literal", "presorted"):
        monkeypatch.setenv("ZRT_BVH_BUILD", how)
        with Z.Scene(hsrc.desc, device=-1) as s:
            order, vis = s.bvh_order()
            info = s.bvh_info(A.ZRT_FLAG_BVH_REFERENCE)
            sah = s.bvh_info()
            got.append((order.copy(), vis.copy(), info.reference_nodes, info.reference_max_depth, info.nodes,
                        info.max_depth, sah.nodes, sah.max_depth))
    hsrc.close()
    assert np.array_equal(got[0][0], got[1][0]) and np.array_equal(got[0][1], got[1][1])
    assert got[0][2:] == got[1][2:]
