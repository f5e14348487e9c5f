//! raytrace_zrt.zig — drop-in for src/raytrace.zig's `render` (raytrace.zig:136-138) that runs the path tracer on
//! B200 GPUs through libzrt's C ABI (include/zrt.h).
//!
//! NOT COMPILED in this repository's image (no Zig toolchain; see zig/build.zig).  Written against the reference's
//! dialect (Zig 0.9.0-dev: `*Allocator`, `std.debug.warn`, `@intCast(T, x)`) and its types as they are in
//! src/{surface,sphere,triangle,material,texture,image,camera}.zig.  The exercised bindings of the same ABI are the
//! C++ host (zraytrace_b200/csrc/host/) and the ctypes binding (zraytrace_b200/lib.py) the GPU tests call through.
//!
//! What changes for a caller: nothing in the signature.  `random` is unused (the device path draws from a
//! counter-based generator keyed (pixel, sample, bounce, seed), BASELINE.json north_star); the BVH is built by the
//! library from the flat surface list (so `preprocessSufraces`, raytrace.zig:111-133, is not called); `ngpus` below
//! selects how many devices share the samples of every pixel.
const std = @import("std");
const Allocator = std.mem.Allocator;
const ArrayList = std.ArrayList;
const Random = std.rand.Random;
const Camera = @import("camera.zig").Camera;
const Surface = @import("surface.zig").Surface;
const Material = @import("material.zig").Material;
const Texture = @import("texture.zig").Texture;
const Image = @import("image.zig").Image;
const Vec3 = @import("vector.zig").Vec3;
pub const RenderParams = @import("raytrace.zig").RenderParams; // raytrace.zig:102-108, unchanged

const c = @cImport({
    @cInclude("zrt.h");
});

pub const ZrtError = error{ ZrtSceneCreate, ZrtRender, ZrtNoDevice };

/// Number of GPUs that share the work: samples-per-pixel are split across them and the f32 accumulators are summed with
/// one NCCL reduce (zrt_multi_render).  0 = every visible device.
pub var ngpus: u32 = 0;
/// RNG key; the reference seeds its PRNG with 42 in every scene (scenes.zig:32,60,108,174,212,239).
pub var seed: u64 = 42;
/// `false` reproduces `while (x < image.height)` (raytrace.zig:168); `true` renders the full width of a non-square image.
pub var full_width: bool = false;

fn v3(v: Vec3) c.zrt_vec3 {
    return .{ .x = v.x, .y = v.y, .z = v.z };
}

/// *const Material -> index into zrt_scene_desc.materials, textures folded in (material.zig:16-29, texture.zig:7-16)
const MaterialTable = struct {
    allocator: *Allocator,
    keys: ArrayList(*const Material),
    materials: ArrayList(c.zrt_material),
    textures: ArrayList(c.zrt_texture),
    texels: ArrayList([]u8), // 8-bit copies of the ImageTexture pixels, freed in deinit

    fn init(allocator: *Allocator) MaterialTable {
        return .{
            .allocator = allocator,
            .keys = ArrayList(*const Material).init(allocator),
            .materials = ArrayList(c.zrt_material).init(allocator),
            .textures = ArrayList(c.zrt_texture).init(allocator),
            .texels = ArrayList([]u8).init(allocator),
        };
    }

    fn deinit(self: *MaterialTable) void {
        for (self.texels.items) |t| self.allocator.free(t);
        self.texels.deinit();
        self.textures.deinit();
        self.materials.deinit();
        self.keys.deinit();
    }

    fn addTexture(self: *MaterialTable, texture: Texture) !u32 {
        var t = std.mem.zeroes(c.zrt_texture);
        switch (texture) {
            .color => |ct| {
                t.kind = c.ZRT_TEXTURE_COLOR;
                t.r = ct.color.r;
                t.g = ct.color.g;
                t.b = ct.color.b;
            },
            .image => |it| {
                // png_image.readFile stores byte / 255.0 as f32 (png_image.zig:87), rows already flipped (:86).
                // libzrt wants the bytes back and divides by 255 on lookup, bit-identically: round(c * 255) is exact.
                const n = it.image.pixels.len;
                var bytes = try self.allocator.alloc(u8, n * 3);
                for (it.image.pixels) |p, i| {
                    bytes[3 * i + 0] = @floatToInt(u8, p.r * 255.0 + 0.5);
                    bytes[3 * i + 1] = @floatToInt(u8, p.g * 255.0 + 0.5);
                    bytes[3 * i + 2] = @floatToInt(u8, p.b * 255.0 + 0.5);
                }
                try self.texels.append(bytes);
                t.kind = c.ZRT_TEXTURE_IMAGE;
                t.width = it.image.width;
                t.height = it.image.height;
                t.channels = 3;
                t.pixels = bytes.ptr;
                t.u_offset = it.u_offset; // texture.zig:14-16 default (0.19, 0.1)
                t.v_offset = it.v_offset;
            },
        }
        try self.textures.append(t);
        return @intCast(u32, self.textures.items.len - 1);
    }

    fn index(self: *MaterialTable, material: *const Material) !u32 {
        for (self.keys.items) |k, i| {
            if (k == material) return @intCast(u32, i);
        }
        var m = std.mem.zeroes(c.zrt_material);
        switch (material.*) {
            .lambertian => |l| {
                m.kind = c.ZRT_MATERIAL_LAMBERTIAN;
                m.texture = try self.addTexture(l.texture);
            },
            .metal => |mt| {
                m.kind = c.ZRT_MATERIAL_METAL;
                m.texture = try self.addTexture(mt.texture);
            },
            .dielectric => |d| {
                m.kind = c.ZRT_MATERIAL_DIELECTRIC;
                m.index_of_refraction = d.index_of_refraction;
            },
        }
        try self.keys.append(material);
        try self.materials.append(m);
        return @intCast(u32, self.materials.items.len - 1);
    }
};

/// Render a scene (same contract as raytrace.render, raytrace.zig:136-203): returns an Image the caller deinit()s,
/// pixels in image.zig:74-103 layout (row 0 = bottom scanline), prints the Progress counters (raytrace.zig:191-201).
pub fn render(allocator: *Allocator, random: *Random, camera: Camera, surfaces: ArrayList(Surface), render_params: RenderParams) !*Image {
    _ = random;
    var spheres = ArrayList(c.zrt_sphere).init(allocator);
    defer spheres.deinit();
    var tris = ArrayList(c.zrt_triangle).init(allocator);
    defer tris.deinit();
    var list = ArrayList(c.zrt_surface).init(allocator);
    defer list.deinit();
    var mats = MaterialTable.init(allocator);
    defer mats.deinit();

    // ArrayList(Surface) in order: the position in this list is the surface id and decides ties (raytrace.zig:75-81)
    for (surfaces.items) |*s| {
        switch (s.*) {
            .sphere => |sp| {
                try list.append(.{ .kind = c.ZRT_SURFACE_SPHERE, .index = @intCast(u32, spheres.items.len) });
                try spheres.append(.{ .center = v3(sp.center), .radius = sp.radius, .material = try mats.index(sp.material) });
            },
            .triangle => |t| {
                try list.append(.{ .kind = c.ZRT_SURFACE_TRIANGLE, .index = @intCast(u32, tris.items.len) });
                try tris.append(.{ .a = v3(t.a), .b = v3(t.b), .c = v3(t.c), .material = try mats.index(t.material) });
            },
            .bvh_node => unreachable, // callers pass the flat list; libzrt builds the hierarchy (bvh.zig:62-185) itself
        }
    }
    var desc = c.zrt_scene_desc{
        .n_surfaces = @intCast(u32, list.items.len),
        .surfaces = list.items.ptr,
        .n_spheres = @intCast(u32, spheres.items.len),
        .spheres = spheres.items.ptr,
        .n_triangles = @intCast(u32, tris.items.len),
        .triangles = tris.items.ptr,
        .n_materials = @intCast(u32, mats.materials.items.len),
        .materials = mats.materials.items.ptr,
        .n_textures = @intCast(u32, mats.textures.items.len),
        .textures = mats.textures.items.ptr,
    };

    const visible = c.zrt_device_count();
    if (visible <= 0) return ZrtError.ZrtNoDevice; // there is no CPU fallback: keep raytrace.zig for that
    const n = if (ngpus == 0 or ngpus > @intCast(u32, visible)) @intCast(u32, visible) else ngpus;

    var group: ?*c.zrt_multi = null;
    if (c.zrt_multi_create(&desc, null, @intCast(c_int, n), &group) != c.ZRT_OK) {
        std.debug.warn("zrt_multi_create: {s}\n", .{c.zrt_last_error()});
        return ZrtError.ZrtSceneCreate;
    }
    defer c.zrt_multi_destroy(group);

    var params = std.mem.zeroes(c.zrt_params);
    params.width = render_params.width;
    params.height = render_params.height;
    params.samples_per_pixel = render_params.samples_per_pixel;
    params.max_depth = render_params.max_depth;
    params.bounded_volume_hierarchy = @boolToInt(render_params.bounded_volume_hierarchy);
    params.x_limit = if (full_width) c.ZRT_XLIMIT_WIDTH else c.ZRT_XLIMIT_HEIGHT;
    params.seed = seed;
    var cam = c.zrt_camera{
        .origin = v3(camera.origin),
        .lower_left_corner = v3(camera.lower_left_corner),
        .horizontal = v3(camera.horizontal),
        .vertical = v3(camera.vertical),
    };

    var image = try Image.init(allocator, render_params.width, render_params.height);
    errdefer image.deinit();
    var counters: c.zrt_counters = undefined;
    var timing: c.zrt_timing = undefined;
    // Color is {r, g, b: f32} (image.zig:9-12): Image.pixels is exactly the float[W * H * 3] zrt_multi_render fills
    if (c.zrt_multi_render(group, &cam, &params, @ptrCast([*]f32, image.pixels.ptr), &counters, &timing) != c.ZRT_OK) {
        std.debug.warn("zrt_multi_render: {s}\n", .{c.zrt_last_error()});
        return ZrtError.ZrtRender;
    }
    // raytrace.zig:191-201
    std.debug.warn("Render stats\n", .{});
    std.debug.warn("  Recursion depth hits:  {}\n", .{counters.recursion_depth_hits});
    std.debug.warn("  Reflections:           {}\n", .{counters.reflections});
    std.debug.warn("  Background hits:       {}\n", .{counters.background_hits});
    std.debug.warn("  Samples:               {}\n", .{counters.samples_processed});
    std.debug.warn("  Total rays:            {}\n", .{counters.rays_processed});
    std.debug.warn("  Prepare runtime:       {d:.2} ms (flatten, BVH, upload)\n", .{timing.prepare_ms});
    std.debug.warn("  Render runtime:        {d:.2} ms on {} GPU(s), slowest trace {d:.2} ms\n", .{ timing.total_ms, n, timing.kernel_ms });
    return image;
}
