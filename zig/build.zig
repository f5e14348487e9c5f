// build.zig — how a zraytrace checkout links libzrt.  NOT COMPILED in this repository's image: there is no Zig
// toolchain here (probed: zig, go, javac, node, rustc all absent), so this file is the binding a maintainer would
// add, written against the reference's own dialect (Zig 0.9.0-dev, `std.build.Builder`) and modelled line for line
// on its build.zig:1-30, where libpng is linked the same way (build.zig:17-19).
//
// Layout assumed: this directory copied into the zraytrace checkout as `zrt/`, with
//   zrt/include/zrt.h               (include/zrt.h of this repository)
//   zrt/lib/libzrt.so               (zraytrace_b200/libzrt.so, built by `make -C zraytrace_b200/csrc`)
//   src/raytrace_zrt.zig            (zig/raytrace_zrt.zig)
// and `scenes.zig` importing `raytrace_zrt.zig` instead of `raytrace.zig` (one line: scenes.zig:10).
const Builder = @import("std").build.Builder;

pub fn build(b: *Builder) void {
    const target = b.standardTargetOptions(.{});
    const mode = b.standardReleaseOptions();

    const exe = b.addExecutable("raytrace", "src/main.zig");
    exe.setTarget(target);
    exe.setBuildMode(mode);
    exe.linkLibC();
    exe.addIncludeDir("/usr/local/include");
    exe.linkSystemLibrary("png"); // png_image.zig keeps reading and writing the PNGs (build.zig:19)
    // --- libzrt: the B200 path-tracing core behind raytrace.render() ---
    exe.addIncludeDir("zrt/include");
    exe.addLibPath("zrt/lib");
    exe.linkSystemLibrary("zrt"); // libzrt.so carries its CUDA runtime statically; NCCL is dlopen'ed only by zrt_multi_*
    exe.addRPath("zrt/lib");
    exe.install();

    const run_cmd = exe.run();
    run_cmd.step.dependOn(b.getInstallStep());
    if (b.args) |args| {
        run_cmd.addArgs(args);
    }

    const run_step = b.step("run", "Run the app");
    run_step.dependOn(&run_cmd.step);
}
