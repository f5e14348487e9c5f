// main.cpp — CLI with the reference's argv (main.zig:12-36):
//     zrt_cli width height samples depth scene_index filename [--assets DIR] [--device N] [--variant V]
// Forces bounded_volume_hierarchy = true like main.zig:30-32 and writes an 8-bit RGB PNG.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "zrt_host.h"

int main(int argc, char **argv) {
    std::fprintf(stderr, "raytrace (libzrt, sm_100a)\nUSAGE;\nzrt_cli width heigth samples depth scene_index filename "
                         "[--assets DIR] [--device N] [--variant V] [--xlimit-width]\n");
    if (argc < 7) return 2;
    zrt_params p{};
    p.width = (uint32_t)std::strtoul(argv[1], nullptr, 10);
    p.height = (uint32_t)std::strtoul(argv[2], nullptr, 10);
    p.samples_per_pixel = (uint32_t)std::strtoul(argv[3], nullptr, 10);
    p.max_depth = (uint32_t)std::strtoul(argv[4], nullptr, 10);
    const uint32_t scene_index = (uint32_t)std::strtoul(argv[5], nullptr, 10);
    const char *filename = argv[6];
    p.bounded_volume_hierarchy = 1;
    p.seed = 42; // scenes.zig:32,60,...
    std::string assets = ".";
    int device = 0;
    uint32_t variant = 0;
    for (int i = 7; i < argc; i++) {
        if (!std::strcmp(argv[i], "--assets") && i + 1 < argc) assets = argv[++i];
        else if (!std::strcmp(argv[i], "--device") && i + 1 < argc) device = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--variant") && i + 1 < argc) variant = (uint32_t)std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--xlimit-width")) p.x_limit = ZRT_XLIMIT_WIDTH;
    }
    std::vector<float> image((size_t)p.width * p.height * 3);
    zrt_counters c{};
    zrt_timing t{};
    const int rc = zrt_host_render_scene(scene_index, assets.c_str(), variant, &p, device, image.data(), &c, &t);
    if (rc != ZRT_OK) {
        std::fprintf(stderr, "render failed (%d): %s\n", rc, zrt_last_error());
        return 1;
    }
    // raytrace.zig:190-201
    std::fprintf(stderr, "Rendering ready\n  Total reflections:     %llu\n  Total background hits: %llu\n  Total pixels:          %llu\n"
                         "  Total samples:         %llu\n  Total rays:            %llu\n  Recursion limit hits:  %llu\n"
                         "    Prepare runtime:     %.2f ms\n    Render runtime:      %.2f ms (kernel %.2f ms)\n",
                 (unsigned long long)c.reflections, (unsigned long long)c.background_hits, (unsigned long long)c.pixels_processed,
                 (unsigned long long)c.samples_processed, (unsigned long long)c.rays_processed,
                 (unsigned long long)c.recursion_depth_hits, t.prepare_ms, t.total_ms, t.kernel_ms);
    if (zrt_host_png_write(filename, image.data(), p.width, p.height) != ZRT_OK) {
        std::fprintf(stderr, "Can't open file %s\n", filename);
        return 1;
    }
    return 0;
}
