// zrt_scene.h — the device-resident scene behind the opaque zrt_scene* of include/zrt.h, shared by zrt_api.cu (one
// device) and zrt_multi.cu (the same scene replicated on several devices, one NCCL reduce of the accumulators).
#pragma once
#include <cuda_runtime.h>

#include <memory>
#include <string>
#include <vector>

#include "zrt_internal.h"

namespace zrt {

// records the message zrt_last_error() returns on this thread and passes the code through
int fail(int code, const std::string &msg);
#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return fail(ZRT_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));                \
    } while (0)

// Device buffers come from the device's default stream-ordered memory pool (cudaMallocAsync) with the release
// threshold raised, so creating and destroying scenes or scratch images does not pay cudaMalloc/cudaFree
// (the 100+ ms spikes seen in the first end-to-end measurements) after the first use.
cudaStream_t g_alloc_stream(int device);

template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    // H2D on `st`: ordered before everything enqueued on `st` afterwards.  For pageable memory the call returns once
    // the source has been staged, so the caller may free `src` right away; the scene builders synchronise `st` once
    // at the end, which makes the data visible to every other stream as well.
    cudaError_t upload(const std::vector<T> &v, cudaStream_t st) { return upload(v.data(), v.size(), st); }
    cudaError_t upload(const T *src, size_t count, cudaStream_t st) {
        cudaError_t e = reserve(count);
        if (e != cudaSuccess || count == 0) return e;
        return cudaMemcpyAsync(p, src, count * sizeof(T), cudaMemcpyHostToDevice, st);
    }
    cudaError_t reserve(size_t count) {
        if (count <= n) return cudaSuccess;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaStream_t st = g_alloc_stream(dev);
        release();
        cudaError_t e = cudaMallocAsync(&p, count * sizeof(T), st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st); // usable from any stream afterwards
        if (e == cudaSuccess) n = count;
        else p = nullptr;
        return e;
    }
    void release() {
        if (p) {
            int dev = 0;
            cudaGetDevice(&dev);
            cudaFreeAsync(p, g_alloc_stream(dev));
        }
        p = nullptr;
        n = 0;
    }
};

// Host half of a scene representation (surface list, reference-order tree, or SAH tree): the flattened arrays,
// ready for upload.  Built once per scene; the replicas of a multi-GPU scene share it.
struct HostRep {
    int mode = MODE_LIST;
    uint32_t n_list = 0, root = REF_EMPTY;
    FlatBvh info;
    std::vector<DevSphere> spheres;
    std::vector<float4> A, E1, E2; // triangle planes, slot order
    std::vector<TriMeta> meta;
    std::vector<uint32_t> list;
};

// one device-resident representation of the scene
struct DevRep {
    bool ready = false;
    int mode = MODE_LIST;
    DevBuf<DevSphere> spheres;
    DevBuf<float4> triA, triE1, triE2;
    DevBuf<TriMeta> triMeta;
    DevBuf<uint32_t> list;
    DevBuf<DevNode> nodes;
    std::shared_ptr<const HostRep> host;
    uint32_t n_spheres = 0, n_list = 0, root = REF_EMPTY;
    float prepare_ms = 0.0f;
    void release() {
        spheres.release(); triA.release(); triE1.release(); triE2.release();
        triMeta.release(); list.release(); nodes.release();
        ready = false;
    }
};

} // namespace zrt

struct zrt_scene {
    int device = -1;
    zrt::HostScene host;
    bool all_spheres = false;
    zrt::DevBuf<zrt::DevMaterial> mats;
    std::vector<zrt::DevBuf<uint8_t>> d_texels;
    zrt::DevRep rep_list, rep_bvh, rep_sah;
    zrt::FlatBvh host_bvh[2]; // host-only inspection (device == -1)
    bool host_bvh_ready[2] = {false, false};
    // scratch owned by the scene
    zrt::DevBuf<float> part, image;
    zrt::DevBuf<uint8_t> image8;
    zrt::DevBuf<unsigned long long> counters;
    zrt::DevBuf<uint32_t> hit_id, work;
    zrt::DevBuf<float> hit_t;
    uint64_t launch_count = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    // zrt_render_device enqueues on the CALLER's stream while the scene's buffers are freed on the allocation stream:
    // ev_user marks the end of the last such render, and everything that frees, grows or reuses a scene buffer waits
    // for it first (quiesce / orderAfterUser).  A scene is used from one stream at a time.
    cudaEvent_t ev_user = nullptr;
    bool user_pending = false;
};

namespace zrt {

struct Plan {
    KParams P;
    int mode;
    uint32_t n_samples;
    size_t n_floats;
};

// zrt_api.cu
int requireDevice(zrt_scene *sc);
DevRep *repFor(zrt_scene *sc, const zrt_params *p); // which representation `p` selects (raytrace.zig:127)
int selectRep(zrt_scene *sc, const zrt_params *p, DevRep **out);
int makePlan(zrt_scene *sc, const zrt_camera *cam, const zrt_params *p, DevRep *r, Plan *plan);
// enqueue everything for one render on `st`; d_rgb receives the final image (or the raw sum with ZRT_FLAG_RAW_SUM)
int enqueueRender(zrt_scene *sc, Plan &plan, float *d_rgb, unsigned long long *d_counters, cudaStream_t st, cudaEvent_t e_k0,
                  cudaEvent_t e_k1, cudaEvent_t e_r1, uint32_t *launches, uint8_t *d_rgb8 = nullptr);
void quiesce(zrt_scene *sc);

} // namespace zrt
