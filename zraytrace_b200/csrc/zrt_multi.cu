// zrt_multi.cu — the multi-GPU form of raytrace.render() behind the C ABI (include/zrt.h "multi-GPU").
// BASELINE.json north_star / SURVEY §8(e): the scene is replicated, rank g of W traces the global samples
// [g*spp/W, (g+1)*spp/W) of every pixel, and the f32 accumulators are summed with ONE NCCL reduce over NVLink (plus
// one of the six u64 counters); the root applies the reference's final 1/spp (raytrace.zig:157,182).
// The RNG is keyed on the global sample index, so the union of the traced paths - and every counter - is the same
// for any W; the image differs only by the association of the f32 sum.
//
// Two ways to form the group, same render call:
//   zrt_multi_create        one process, n devices:   ncclCommInitAll, every rank is local
//   zrt_multi_create_rank   one process per device:   ncclCommInitRank with an id the caller distributes (MPI,
//                           torch.distributed, a file): what `torchrun bench.py` uses
// NCCL is loaded with dlopen("libnccl.so.2"): a host that links libzrt the way the reference links libpng
// (build.zig:17-19) needs no NCCL at build time, a single-GPU host none at run time, and inside a PyTorch process the
// already loaded libnccl.so.2 (same soname) is the one that is used.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <chrono>
#include <cstring>
#include <initializer_list>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "zrt_scene.h"

namespace zrt {
void launch_resolve(const float *part, float *out, uint32_t n, uint32_t chunks, float scale, cudaStream_t st);
}
using namespace zrt;

namespace {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Reduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};

NcclApi g_nccl;
const std::string &ncclLoadError() { return g_nccl.error; }

NcclApi *nccl() { // nullptr (and ncclLoadError()) if the library cannot be loaded
    NcclApi &api = g_nccl;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names)
            if ((api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
        if (!api.lib) {
            api.error = std::string("cannot load libnccl.so.2: ") + dlerror();
            return;
        }
        bool ok = true;
        auto sym = [&](const char *name) {
            void *p = dlsym(api.lib, name);
            if (!p) { ok = false; api.error = std::string("libnccl.so.2 lacks ") + name; }
            return p;
        };
        api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(sym("ncclCommInitAll"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.Reduce = reinterpret_cast<decltype(api.Reduce)>(sym("ncclReduce"));
        api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
        api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
        if (!ok) { dlclose(api.lib); api.lib = nullptr; }
    });
    return api.lib ? &api : nullptr;
}

#define NCCL_TRY(expr)                                                                                   \
    do {                                                                                                 \
        ncclResult_t _r = (expr);                                                                        \
        if (_r != ncclSuccess) return fail(ZRT_ERR_NCCL, std::string(#expr) + ": " + N->GetErrorString(_r)); \
    } while (0)

struct Replica {
    zrt_scene *scene = nullptr;
    int device = -1, rank = -1; // global rank in the group
    ncclComm_t comm = nullptr;
    DevBuf<float> accum;               // raw f32 sums of this rank's samples; on the root: the reduced sums
    DevBuf<unsigned long long> counts; // six u64 counters, same
    cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
};

} // namespace

struct zrt_multi {
    std::vector<Replica> local; // the ranks this process drives
    int world = 0;
    DevBuf<float> image;        // root only: accum * 1/spp
};

namespace {

// rank g of W: global samples [g*spp/W, (g+1)*spp/W)  (zraytrace_b200/distributed.py uses the same split)
void sampleRange(uint32_t spp, int rank, int world, uint32_t *b, uint32_t *e) {
    *b = (uint32_t)(((uint64_t)rank * spp) / (uint64_t)world);
    *e = (uint32_t)((((uint64_t)rank + 1) * spp) / (uint64_t)world);
}

void destroyMulti(zrt_multi *m) {
    if (!m) return;
    NcclApi *N = nccl();
    for (Replica &r : m->local) {
        if (r.device >= 0) cudaSetDevice(r.device);
        if (r.scene && r.scene->stream) cudaStreamSynchronize(r.scene->stream);
        if (r.comm && N) N->CommDestroy(r.comm);
        r.accum.release();
        r.counts.release();
        if (r.rank == 0) m->image.release();
        for (cudaEvent_t e : {r.e0, r.e1, r.e2})
            if (e) cudaEventDestroy(e);
        if (r.scene) zrt_scene_destroy(r.scene);
    }
    delete m;
}

int initReplica(Replica &r, const zrt_scene_desc *desc, int device, int rank) {
    r.device = device;
    r.rank = rank;
    int rc = zrt_scene_create(desc, device, &r.scene);
    if (rc != ZRT_OK) return rc;
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(cudaEventCreate(&r.e0));
    CUDA_TRY(cudaEventCreate(&r.e1));
    CUDA_TRY(cudaEventCreate(&r.e2));
    return ZRT_OK;
}

} // namespace

extern "C" {

int zrt_nccl_version(void) {
    NcclApi *N = nccl();
    int v = 0;
    if (!N || N->GetVersion(&v) != ncclSuccess) return 0;
    return v;
}

int zrt_comm_id(uint8_t id[ZRT_COMM_ID_BYTES]) {
    static_assert(sizeof(ncclUniqueId) == ZRT_COMM_ID_BYTES, "ZRT_COMM_ID_BYTES must be sizeof(ncclUniqueId)");
    if (!id) return fail(ZRT_ERR_INVALID, "id is NULL");
    NcclApi *N = nccl();
    if (!N) return fail(ZRT_ERR_NCCL, ncclLoadError());
    ncclUniqueId u;
    NCCL_TRY(N->GetUniqueId(&u));
    std::memcpy(id, &u, sizeof(u));
    return ZRT_OK;
}

int zrt_multi_create(const zrt_scene_desc *desc, const int *devices, int n_devices, zrt_multi **out) {
    if (!out) return fail(ZRT_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (n_devices < 1 || n_devices > 64) return fail(ZRT_ERR_INVALID, "n_devices must be 1..64");
    const int visible = zrt_device_count();
    if (visible == 0) return fail(ZRT_ERR_NO_DEVICE, "no CUDA device visible; libzrt has no CPU path");
    std::vector<int> devs(n_devices);
    for (int i = 0; i < n_devices; i++) {
        devs[i] = devices ? devices[i] : i;
        if (devs[i] < 0 || devs[i] >= visible) return fail(ZRT_ERR_INVALID, "device index out of range");
        for (int j = 0; j < i; j++)
            if (devs[j] == devs[i]) return fail(ZRT_ERR_INVALID, "a device is listed twice");
    }
    zrt_multi *m = new (std::nothrow) zrt_multi();
    if (!m) return fail(ZRT_ERR_OOM, "out of memory");
    m->world = n_devices;
    m->local.resize(n_devices);
    for (int i = 0; i < n_devices; i++) {
        const int rc = initReplica(m->local[i], desc, devs[i], i);
        if (rc != ZRT_OK) { destroyMulti(m); return rc; }
    }
    if (n_devices > 1) {
        NcclApi *N = nccl();
        if (!N) { destroyMulti(m); return fail(ZRT_ERR_NCCL, "more than one device needs NCCL: " + ncclLoadError()); }
        std::vector<ncclComm_t> comms(n_devices, nullptr);
        const ncclResult_t r = N->CommInitAll(comms.data(), n_devices, devs.data());
        if (r != ncclSuccess) {
            destroyMulti(m);
            return fail(ZRT_ERR_NCCL, std::string("ncclCommInitAll: ") + N->GetErrorString(r));
        }
        for (int i = 0; i < n_devices; i++) m->local[i].comm = comms[i];
    }
    *out = m;
    return ZRT_OK;
}

int zrt_multi_create_rank(const zrt_scene_desc *desc, int device, const uint8_t id[ZRT_COMM_ID_BYTES], int rank, int world,
                          zrt_multi **out) {
    if (!out) return fail(ZRT_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (world < 1 || rank < 0 || rank >= world) return fail(ZRT_ERR_INVALID, "bad rank / world size");
    if (world > 1 && !id) return fail(ZRT_ERR_INVALID, "id is NULL");
    const int visible = zrt_device_count();
    if (visible == 0) return fail(ZRT_ERR_NO_DEVICE, "no CUDA device visible; libzrt has no CPU path");
    if (device < 0 || device >= visible) return fail(ZRT_ERR_INVALID, "device index out of range");
    zrt_multi *m = new (std::nothrow) zrt_multi();
    if (!m) return fail(ZRT_ERR_OOM, "out of memory");
    m->world = world;
    m->local.resize(1);
    int rc = initReplica(m->local[0], desc, device, rank);
    if (rc != ZRT_OK) { destroyMulti(m); return rc; }
    if (world > 1) {
        NcclApi *N = nccl();
        if (!N) { destroyMulti(m); return fail(ZRT_ERR_NCCL, "more than one rank needs NCCL: " + ncclLoadError()); }
        ncclUniqueId u;
        std::memcpy(&u, id, sizeof(u));
        const ncclResult_t r = N->CommInitRank(&m->local[0].comm, world, u, rank); // collective over all ranks
        if (r != ncclSuccess) {
            destroyMulti(m);
            return fail(ZRT_ERR_NCCL, std::string("ncclCommInitRank: ") + N->GetErrorString(r));
        }
    }
    *out = m;
    return ZRT_OK;
}

void zrt_multi_destroy(zrt_multi *m) { destroyMulti(m); }

int zrt_multi_reload(zrt_multi *m, const zrt_scene_desc *desc) {
    if (!m) return fail(ZRT_ERR_INVALID, "group is NULL");
    // one host thread per local device: copying, validating and uploading the description are independent per replica
    // (8 devices in one process: 7.1 -> 5.7 ms end to end on the headline render)
    const size_t n = m->local.size();
    std::vector<zrt_scene *> fresh(n, nullptr);
    std::vector<int> rcs(n, ZRT_OK);
    std::vector<std::string> msgs(n);
    auto job = [&](size_t i) {
        rcs[i] = zrt_scene_create(desc, m->local[i].device, &fresh[i]); // validates before anything is torn down
        if (rcs[i] != ZRT_OK) msgs[i] = zrt_last_error();
    };
    try {
        std::vector<Worker> th;
        for (size_t i = 1; i < n; i++) th.emplace_back([&job, i] { job(i); });
        job(0);
        for (Worker &t : th) t.join();
    } catch (const std::exception &e) {
        for (zrt_scene *f : fresh)
            if (f) zrt_scene_destroy(f);
        return fail(ZRT_ERR_OOM, std::string("scene reload failed: ") + e.what());
    }
    for (size_t i = 0; i < n; i++)
        if (rcs[i] != ZRT_OK) {
            for (zrt_scene *f : fresh)
                if (f) zrt_scene_destroy(f);
            return fail(rcs[i], msgs[i]);
        }
    for (size_t i = 0; i < n; i++) {
        zrt_scene_destroy(m->local[i].scene);
        m->local[i].scene = fresh[i];
    }
    return ZRT_OK;
}

int zrt_multi_world_size(const zrt_multi *m) { return m ? m->world : 0; }

uint64_t zrt_multi_launch_count(const zrt_multi *m) {
    uint64_t n = 0;
    if (m)
        for (const Replica &r : m->local) n += zrt_scene_launch_count(r.scene);
    return n;
}

int zrt_multi_render(zrt_multi *m, const zrt_camera *camera, const zrt_params *params, float *out_rgb,
                     zrt_counters *counters, zrt_timing *timing) {
    if (!m || m->local.empty()) return fail(ZRT_ERR_INVALID, "group is NULL");
    if (!camera || !params) return fail(ZRT_ERR_INVALID, "camera/params is NULL");
    if (params->flags & ZRT_FLAG_RAW_SUM) return fail(ZRT_ERR_INVALID, "ZRT_FLAG_RAW_SUM belongs to zrt_render_device; the group normalises itself");
    if (params->sample_begin != 0 || params->sample_end != 0)
        return fail(ZRT_ERR_INVALID, "the group splits the samples itself: leave sample_begin / sample_end at 0");
    NcclApi *N = m->world > 1 ? nccl() : nullptr;
    const size_t n_floats = (size_t)params->width * params->height * 3;
    Replica *root = nullptr;
    uint32_t launches = 0;
    const auto t_prep0 = std::chrono::steady_clock::now();

    // ---- every local rank: upload on first use (the host flattening is shared), then its share of the samples ----
    DevRep *first_rep = nullptr;
    for (Replica &r : m->local) {
        int rc = requireDevice(r.scene);
        if (rc != ZRT_OK) return rc;
        DevRep *rep = repFor(r.scene, params);
        if (!rep->ready && !rep->host && first_rep && first_rep->host) rep->host = first_rep->host; // one flattening
        rc = selectRep(r.scene, params, &rep);
        if (rc != ZRT_OK) return rc;
        if (!first_rep) first_rep = rep;
    }
    const float prep_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_prep0).count();
    for (Replica &r : m->local) {
        int rc = requireDevice(r.scene);
        if (rc != ZRT_OK) return rc;
        cudaStream_t st = r.scene->stream;
        CUDA_TRY(r.accum.reserve(n_floats));
        CUDA_TRY(r.counts.reserve(6));
        zrt_params p = *params;
        sampleRange(params->samples_per_pixel, r.rank, m->world, &p.sample_begin, &p.sample_end);
        p.flags |= ZRT_FLAG_RAW_SUM;
        CUDA_TRY(cudaEventRecord(r.e0, st));
        if (p.sample_begin == p.sample_end && params->samples_per_pixel > 0) {
            // fewer samples than ranks: nothing to trace here.  (An empty range must not reach makePlan as (0, 0),
            // which means "all samples".)
            CUDA_TRY(cudaMemsetAsync(r.accum.p, 0, n_floats * sizeof(float), st));
            CUDA_TRY(cudaMemsetAsync(r.counts.p, 0, 6 * sizeof(unsigned long long), st));
        } else {
            DevRep *rep = nullptr;
            rc = selectRep(r.scene, &p, &rep);
            if (rc != ZRT_OK) return rc;
            Plan plan;
            rc = makePlan(r.scene, camera, &p, rep, &plan);
            if (rc != ZRT_OK) return rc;
            uint32_t l = 0;
            rc = enqueueRender(r.scene, plan, r.accum.p, r.counts.p, st, nullptr, nullptr, nullptr, &l);
            if (rc != ZRT_OK) return rc;
            launches += l;
        }
        CUDA_TRY(cudaEventRecord(r.e1, st));
        if (r.rank == 0) root = &r;
    }

    // ---- ONE reduce of the accumulators (and one of the counters) to rank 0, on every rank's own stream ----
    if (m->world > 1) {
        if (!N) return fail(ZRT_ERR_NCCL, "NCCL unavailable");
        NCCL_TRY(N->GroupStart());
        for (Replica &r : m->local) {
            NCCL_TRY(N->Reduce(r.accum.p, r.accum.p, n_floats, ncclFloat32, ncclSum, 0, r.comm, r.scene->stream));
            NCCL_TRY(N->Reduce(r.counts.p, r.counts.p, 6, ncclUint64, ncclSum, 0, r.comm, r.scene->stream));
        }
        NCCL_TRY(N->GroupEnd());
    }

    // ---- root: the reference's `* (1 / samples_per_pixel)` (raytrace.zig:157,182), then the image goes home ----
    unsigned long long h_counters[6] = {0, 0, 0, 0, 0, 0};
    if (root) {
        CUDA_TRY(cudaSetDevice(root->device));
        cudaStream_t st = root->scene->stream;
        CUDA_TRY(m->image.reserve(n_floats));
        const float scale = 1.0f / (float)params->samples_per_pixel;
        launch_resolve(root->accum.p, m->image.p, (uint32_t)n_floats, 1, scale, st);
        root->scene->launch_count++;
        launches++;
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaEventRecord(root->e2, st));
        if (out_rgb) CUDA_TRY(cudaMemcpyAsync(out_rgb, m->image.p, n_floats * sizeof(float), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(h_counters, root->counts.p, sizeof(h_counters), cudaMemcpyDeviceToHost, st));
    }
    for (Replica &r : m->local) {
        CUDA_TRY(cudaSetDevice(r.device));
        if (&r != root) CUDA_TRY(cudaEventRecord(r.e2, r.scene->stream));
        CUDA_TRY(cudaStreamSynchronize(r.scene->stream));
    }
    if (counters && root) {
        counters->recursion_depth_hits = h_counters[0];
        counters->reflections = h_counters[1];
        counters->background_hits = h_counters[2];
        counters->pixels_processed = h_counters[3];
        counters->samples_processed = h_counters[4];
        counters->rays_processed = h_counters[5];
    }
    if (timing) {
        // device milliseconds: the slowest local rank's trace, and first launch -> reduced and scaled image on the
        // root (which waits for every rank's contribution): the whole job
        float k_max = 0.0f, tot_max = 0.0f, red = 0.0f;
        for (Replica &r : m->local) {
            float k = 0.0f, tot = 0.0f;
            cudaEventElapsedTime(&k, r.e0, r.e1);
            cudaEventElapsedTime(&tot, r.e0, r.e2);
            k_max = k > k_max ? k : k_max;
            tot_max = tot > tot_max ? tot : tot_max;
            if (&r == root) cudaEventElapsedTime(&red, r.e1, r.e2);
        }
        timing->prepare_ms = prep_ms;
        timing->kernel_ms = k_max;
        timing->resolve_ms = red; // root: wait for the peers + reduce + scale
        timing->total_ms = tot_max;
        timing->launches = launches;
        timing->bvh_nodes = first_rep && first_rep->host ? (uint32_t)first_rep->host->info.nodes.size() : 0u;
    }
    return ZRT_OK;
}

const float *zrt_multi_image_device(const zrt_multi *m) { return m ? m->image.p : nullptr; }

} // extern "C"
