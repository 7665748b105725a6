// rc_api.cu — the C ABI of libraycore_cuda.so (include/raycore_cuda.h): the mutable-TLAS lifecycle of
// src/instanced-bvh.jl:261-1134 (handles, push!/delete!/update*/sync!, compaction) kept on the host in C++,
// with all geometry work on the GPU (rc_build.cu), and the batched query / analysis entry points.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "rc_build.h"
#include "rc_device.cuh"
#include "rc_trace.h"
#include "rc_wave.h"
#include "rc_wave_core.cuh"

namespace {

struct HandleInfo {
    uint32_t start, count;
    bool deleted;
};

thread_local std::string g_create_error;  // message of a failed rc_create / rc_check_exported on this thread (read back by the same thread with rc_last_error(NULL))

}  // namespace

// Page-locked storage for the host mirror of the instance descriptors: the TLAS build / refit uploads it every sync!, and from pageable
// memory that copy (1 MB for 10,000 instances) is staged by the driver at a few GB/s — most of a refit frame.  Falls back to malloc when
// page-locked memory cannot be had (a 64-byte header remembers which).
template <class T>
struct RcPinnedAlloc {
    using value_type = T;
    RcPinnedAlloc() = default;
    template <class U>
    RcPinnedAlloc(const RcPinnedAlloc<U> &) {}
    T *allocate(size_t n) {
        const size_t bytes = n * sizeof(T) + 64;
        void *p = nullptr;
        if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) == cudaSuccess) {
            *static_cast<uint64_t *>(p) = 1;
        } else {
            (void)cudaGetLastError();
            p = malloc(bytes);
            if (!p) throw std::bad_alloc();
            *static_cast<uint64_t *>(p) = 0;
        }
        return reinterpret_cast<T *>(static_cast<char *>(p) + 64);
    }
    void deallocate(T *q, size_t) {
        char *p = reinterpret_cast<char *>(q) - 64;
        if (*reinterpret_cast<uint64_t *>(p)) cudaFreeHost(p);
        else free(p);
    }
    template <class U>
    bool operator==(const RcPinnedAlloc<U> &) const { return true; }
    template <class U>
    bool operator!=(const RcPinnedAlloc<U> &) const { return false; }
};
typedef std::vector<rc_instance_desc, RcPinnedAlloc<rc_instance_desc>> RcInstanceVec;

#define RC_OVF_LIST_CAP 65536u  // rays with an overflowed short stack that a trace call can list for its fix-up pass (more: the pass scans)

struct rc_context {
    // queries on a synced TLAS may come from several host threads at once (the reference calls them under Threads.@threads,
    // src/kernels.jl:64,82); entry points that touch the GPU or the shared launch resources serialise on this lock (RC_ENTER)
    std::recursive_mutex mu;
    int device = 0;
    cudaStream_t stream = nullptr, s_h2d = nullptr, s_d2h = nullptr;
    bool owns_stream = true;
    std::string last_error;
    std::vector<RcDeviceBlas> blas;           // blas_index b+1 <-> blas[b]
    RcInstanceVec instances;                  // host mirror of tlas.instances (page-locked: uploaded by every sync!)
    std::map<uint32_t, HandleInfo> handles;   // ordered by id (Julia: Dict{TLASHandle,UnitRange})
    bool dirty = true, transforms_dirty = false, built = false, last_update_refitted = false;
    uint32_t build_flags = 0;  // RC_BUILD_* defaults of this context (rc_set_build_flags; initialised from the RC_BUILD_FLAGS environment variable)
    uint32_t next_handle = 1;
    RcDeviceTlas tlas;
    RcFlatBlas *d_flat = nullptr;
    uint32_t n_flat_blas = 0, n_flat_prims = 0, synced_blas_nodes = 0, synced_tlas_nodes = 0;
    // trace resources
    unsigned long long *d_work = nullptr;
    RcCounters *d_counters = nullptr;
    uint32_t *d_overflow = nullptr;
    uint32_t ovf_cap = RC_OVF_LIST_CAP;  // entries of the overflow list a trace may use (RC_OVF_LIST_CAP environment variable: smaller, 0 = none; tests)
    uint32_t *h_err = nullptr;  // pinned host word: the hard-error counter is read back with the copy queued BEFORE a call's final sync (no extra round trip)
    rc_ray *d_rays = nullptr;
    rc_hit *d_hits = nullptr;
    size_t cap_rays = 0, cap_hits = 0;
    static const int NEV = 8;
    cudaEvent_t ev_h2d[NEV], ev_k[NEV], ev_t0 = nullptr, ev_t1 = nullptr;
    cudaEvent_t ev_pc_src[2] = {nullptr, nullptr}, ev_pc_done[2] = {nullptr, nullptr};  // rc_peer_copy_async's own pair (host-staged traces re-record ev_k / ev_h2d)
    float last_ms = 0.f, last_build_ms = 0.f;
    bool copy_pending[2] = {false, false};
    uint32_t last_launches = 0;
    int max_blocks = 148;
    // wavefront stages: device table of per-BLAS normal arrays, re-uploaded when normals_version moves
    const float **d_normal_ptrs = nullptr;
    std::vector<const float *> normal_ptrs;  // what d_normal_ptrs currently holds
    // view_factors: row -> flat primitive map of the synced scene (usable when every metadata value occurs at most once)
    uint32_t *d_vf_row_pos = nullptr;
    bool vf_map_built = false, vf_map_usable = false;
    uint32_t vf_out_of_range = 0;
};

#define RC_FAIL(ctx, code, msg)        \
    do {                               \
        (ctx)->last_error = (msg);     \
        return (code);                 \
    } while (0)

#define RC_CUDA(ctx, call)                                                                       \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess) {                                                                 \
            (ctx)->last_error = std::string(#call) + ": " + cudaGetErrorString(e_);              \
            return e_ == cudaErrorMemoryAllocation ? RC_ERR_OUT_OF_MEMORY : RC_ERR_CUDA;         \
        }                                                                                        \
    } while (0)

static inline void use_device(const rc_context *ctx) { cudaSetDevice(ctx->device); }
// select the context's device and hold its lock for the rest of the calling function
#define RC_ENTER(ctx)  \
    use_device(ctx);   \
    std::lock_guard<std::recursive_mutex> rc_guard_((ctx)->mu)
// host-state-only entry points (no GPU work): the lock alone.  Every entry point that reads or writes context state takes one of the two,
// so a query thread interleaving with the (single) mutating thread never sees a half-updated handle table or a reallocating vector.
#define RC_LOCK(ctx) std::lock_guard<std::recursive_mutex> rc_guard_(const_cast<rc_context *>(ctx)->mu)
static int32_t builder_error_code(const std::string &err) { return err.find("supported range") != std::string::npos ? RC_ERR_INVALID_ARGUMENT : RC_ERR_CUDA; }

// Stream-ordered temporaries of one entry point: returned to the pool when the call leaves scope, on every path (an RC_CUDA early
// return included — the pool's release threshold is unbounded, so a leaked block would stay allocated for good).
struct ApiTemps {
    cudaStream_t st;
    std::vector<void *> ptrs;
    explicit ApiTemps(cudaStream_t s) : st(s) {}
    ~ApiTemps() {
        for (void *p : ptrs) cudaFreeAsync(p, st);
    }
    template <class T>
    cudaError_t get(T **out, size_t bytes) {
        void *p = nullptr;
        cudaError_t e = cudaMallocAsync(&p, bytes ? bytes : 1, st);
        if (e == cudaSuccess) { ptrs.push_back(p); *out = static_cast<T *>(p); }
        return e;
    }
};

static RcScene make_scene(const rc_context *ctx) {
    RcScene sc;
    sc.tlas4 = ctx->tlas.nodes4;
    sc.tlas2 = ctx->tlas.nodes2;
    sc.inst = ctx->tlas.rec;
    sc.aux = ctx->tlas.aux;
    sc.n_instances = ctx->tlas.n;
    return sc;
}

extern "C" {

int32_t rc_abi_version(void) { return RC_ABI_VERSION; }

const char *rc_last_error(const rc_context *ctx) { return ctx ? ctx->last_error.c_str() : g_create_error.c_str(); }

int32_t rc_create(int32_t device, rc_context **out) {
    if (!out) return RC_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_create_error = std::string("no CUDA device available: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                         " (libraycore_cuda has no CPU fallback)";
        return RC_ERR_CUDA;
    }
    if (device < 0) {
        if (cudaGetDevice(&device) != cudaSuccess) device = 0;
    }
    if (device >= count) { g_create_error = "device index out of range"; return RC_ERR_INVALID_ARGUMENT; }
    rc_context *ctx = new rc_context();
    ctx->device = device;
    for (int k = 0; k < 3; k++) { ctx->tlas.root_aabb[k] = INFINITY; ctx->tlas.root_aabb[3 + k] = -INFINITY; }
#define CREATE_CK(call)                                                          \
    do {                                                                         \
        cudaError_t e2 = (call);                                                 \
        if (e2 != cudaSuccess) {                                                 \
            g_create_error = std::string(#call) + ": " + cudaGetErrorString(e2); \
            delete ctx;                                                          \
            return RC_ERR_CUDA;                                                  \
        }                                                                        \
    } while (0)
    CREATE_CK(cudaSetDevice(device));
    CREATE_CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CREATE_CK(cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking));
    CREATE_CK(cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking));
    for (int i = 0; i < rc_context::NEV; i++) {
        CREATE_CK(cudaEventCreateWithFlags(&ctx->ev_h2d[i], cudaEventDisableTiming));
        CREATE_CK(cudaEventCreateWithFlags(&ctx->ev_k[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < 2; i++) {
        CREATE_CK(cudaEventCreateWithFlags(&ctx->ev_pc_src[i], cudaEventDisableTiming));
        CREATE_CK(cudaEventCreateWithFlags(&ctx->ev_pc_done[i], cudaEventDisableTiming));
    }
    CREATE_CK(cudaEventCreate(&ctx->ev_t0));
    CREATE_CK(cudaEventCreate(&ctx->ev_t1));
    CREATE_CK(cudaMalloc(&ctx->d_work, sizeof(unsigned long long) * (2 + RC_OVF_LIST_CAP)));  // work counter, overflow-list length, overflow list (rc_trace.h)
    CREATE_CK(cudaMalloc(&ctx->d_counters, sizeof(RcCounters)));
    CREATE_CK(cudaMalloc(&ctx->d_overflow, 4 * sizeof(uint32_t)));  // [0] rays flagged for k_trace_fixup, [1] hard errors, [2] flagged-ray scratch of the inline re-trace policies
    CREATE_CK(cudaMemset(ctx->d_counters, 0, sizeof(RcCounters)));
    CREATE_CK(cudaMemset(ctx->d_overflow, 0, 4 * sizeof(uint32_t)));
    CREATE_CK(cudaHostAlloc(&ctx->h_err, sizeof(uint32_t), cudaHostAllocDefault));
    *ctx->h_err = 0;
    // keep freed blocks cached in the stream-ordered pool: rebuild-per-frame workloads reuse them
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long thresh = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh);
    }
    ctx->max_blocks = rc_trace_max_blocks(device);
    if (const char *e = getenv("RC_OVF_LIST_CAP")) ctx->ovf_cap = std::min<uint32_t>((uint32_t)strtoul(e, nullptr, 0), RC_OVF_LIST_CAP);
    if (const char *e = getenv("RC_BUILD_FLAGS")) ctx->build_flags = (uint32_t)strtoul(e, nullptr, 0) & (RC_BUILD_KEEP_BVH2 | RC_BUILD_ALLOW_REFIT);
#undef CREATE_CK
    *out = ctx;
    return RC_OK;
}

int32_t rc_destroy(rc_context *ctx) {
    if (!ctx) return RC_OK;
    use_device(ctx);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->s_h2d);
    cudaStreamSynchronize(ctx->s_d2h);
    for (auto &b : ctx->blas) rc_free_blas(&b, ctx->stream);
    rc_free_tlas(&ctx->tlas, ctx->stream);
    if (ctx->d_flat) cudaFreeAsync(ctx->d_flat, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->d_work); cudaFree(ctx->d_counters); cudaFree(ctx->d_overflow);
    if (ctx->h_err) cudaFreeHost(ctx->h_err);
    if (ctx->d_rays) cudaFree(ctx->d_rays);
    if (ctx->d_hits) cudaFree(ctx->d_hits);
    if (ctx->d_normal_ptrs) cudaFree(ctx->d_normal_ptrs);
    if (ctx->d_vf_row_pos) cudaFree(ctx->d_vf_row_pos);
    for (int i = 0; i < rc_context::NEV; i++) { cudaEventDestroy(ctx->ev_h2d[i]); cudaEventDestroy(ctx->ev_k[i]); }
    for (int i = 0; i < 2; i++) { cudaEventDestroy(ctx->ev_pc_src[i]); cudaEventDestroy(ctx->ev_pc_done[i]); }
    cudaEventDestroy(ctx->ev_t0); cudaEventDestroy(ctx->ev_t1);
    if (ctx->owns_stream) cudaStreamDestroy(ctx->stream);
    cudaStreamDestroy(ctx->s_h2d); cudaStreamDestroy(ctx->s_d2h);
    delete ctx;
    return RC_OK;
}

void *rc_stream(rc_context *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int32_t rc_set_stream(rc_context *ctx, void *stream) {
    if (!ctx) return RC_ERR_INVALID_ARGUMENT;
    RC_ENTER(ctx);
    RC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->owns_stream) { cudaStreamDestroy(ctx->stream); ctx->owns_stream = false; }
    if (stream) ctx->stream = (cudaStream_t)stream;
    else {
        RC_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        ctx->owns_stream = true;
    }
    return RC_OK;
}

// ------------------------------------------------------------------------------------------------ mutation
static int32_t build_blas_from(rc_context *ctx, const float *verts, uint32_t n_faces, const uint32_t *face_meta, uint32_t flags, RcDeviceBlas *out) {
    if (!verts || n_faces == 0) RC_FAIL(ctx, RC_ERR_NO_VALID_TRIANGLES, "Geometry has no valid triangles");
    const float *d_verts = verts;
    const uint32_t *d_meta = face_meta;
    float *tmp_v = nullptr;
    uint32_t *tmp_m = nullptr;
    ApiTemps tmp(ctx->stream);
    if (!(flags & RC_VERTS_ON_DEVICE)) {
        RC_CUDA(ctx, tmp.get(&tmp_v, sizeof(float) * 9 * (size_t)n_faces));
        RC_CUDA(ctx, cudaMemcpyAsync(tmp_v, verts, sizeof(float) * 9 * (size_t)n_faces, cudaMemcpyHostToDevice, ctx->stream));
        d_verts = tmp_v;
        if (face_meta) {
            RC_CUDA(ctx, tmp.get(&tmp_m, sizeof(uint32_t) * (size_t)n_faces));
            RC_CUDA(ctx, cudaMemcpyAsync(tmp_m, face_meta, sizeof(uint32_t) * (size_t)n_faces, cudaMemcpyHostToDevice, ctx->stream));
            d_meta = tmp_m;
        }
    }
    std::string err;
    cudaEventRecord(ctx->ev_t0, ctx->stream);
    bool ok = rc_build_blas(ctx->stream, d_verts, d_meta, n_faces, (flags | ctx->build_flags) & (RC_BUILD_KEEP_BVH2 | RC_BUILD_ALLOW_REFIT), out, err);
    cudaEventRecord(ctx->ev_t1, ctx->stream);
    if (ok && cudaEventSynchronize(ctx->ev_t1) == cudaSuccess) cudaEventElapsedTime(&ctx->last_build_ms, ctx->ev_t0, ctx->ev_t1);
    if (!ok) {
        rc_free_blas(out, ctx->stream);
        if (err == "Geometry has no valid triangles") RC_FAIL(ctx, RC_ERR_NO_VALID_TRIANGLES, err);
        RC_FAIL(ctx, builder_error_code(err), err);
    }
    return RC_OK;
}

// append_instances_with_handle!, :612-623: the BLAS joins the table, m descriptors join the instance list under a fresh handle
static uint32_t append_blas_with_instances(rc_context *ctx, const RcDeviceBlas &b, const float *transforms, const float *inv_transforms,
                                           const uint32_t *instance_ids, uint32_t m) {
    ctx->blas.push_back(b);
    uint32_t blas_idx = (uint32_t)ctx->blas.size();  // :607
    uint32_t start = (uint32_t)ctx->instances.size();
    if (ctx->instances.capacity() < (size_t)start + m)  // (page-locked allocations are slow: grow by at least half)
        ctx->instances.reserve(std::max((size_t)start + m, ctx->instances.capacity() + ctx->instances.capacity() / 2));
    for (uint32_t i = 0; i < m; i++) {
        rc_instance_desc d;
        d.blas_index = blas_idx;
        d.instance_id = instance_ids ? instance_ids[i] : 0u;  // :672
        memcpy(d.transform, transforms + 12 * (size_t)i, 48);
        if (inv_transforms) memcpy(d.inv_transform, inv_transforms + 12 * (size_t)i, 48);
        else x_mat3x4_inverse(d.transform, d.inv_transform);  // :644, :673
        d.flags = 0;
        ctx->instances.push_back(d);
    }
    uint32_t h = ctx->next_handle++;
    ctx->handles[h] = HandleInfo{start, m, false};
    ctx->dirty = true;
    return h;
}

int32_t rc_push(rc_context *ctx, const float *verts, uint32_t n_faces, const uint32_t *face_meta, const float *transforms, const float *inv_transforms,
                const uint32_t *instance_ids, uint32_t m, uint32_t flags, uint32_t *handle_out) {
    if (!ctx) return RC_ERR_INVALID_ARGUMENT;
    RC_ENTER(ctx);
    if (!transforms || m == 0 || !handle_out) RC_FAIL(ctx, RC_ERR_INVALID_ARGUMENT, "rc_push: transforms, m >= 1 and handle_out are required");
    RcDeviceBlas b;
    int32_t rc = build_blas_from(ctx, verts, n_faces, face_meta, flags, &b);
    if (rc != RC_OK) return rc;
    *handle_out = append_blas_with_instances(ctx, b, transforms, inv_transforms, instance_ids, m);
    return RC_OK;
}

static int32_t find_handle(rc_context *ctx, uint32_t handle, HandleInfo **out) {
    auto it = ctx->handles.find(handle);
    if (it == ctx->handles.end()) RC_FAIL(ctx, RC_ERR_INVALID_HANDLE, "Invalid handle");
    if (it->second.deleted) RC_FAIL(ctx, RC_ERR_DELETED_HANDLE, "Handle has been deleted");
    *out = &it->second;
    return RC_OK;
}

int32_t rc_delete(rc_context *ctx, uint32_t handle, int32_t *deleted) {  // :690-699
    if (!ctx) return RC_ERR_INVALID_ARGUMENT;
    RC_LOCK(ctx);
    if (deleted) *deleted = 0;
    auto it = ctx->handles.find(handle);
    if (it == ctx->handles.end() || it->second.deleted) return RC_OK;
    it->second.deleted = true;
    ctx->dirty = true;
    if (deleted) *deleted = 1;
    return RC_OK;
}

int32_t rc_update_transforms(rc_context *ctx, uint32_t handle, const float *transforms, const float *inv_transforms, uint32_t m) {  // :755-797
    if (!ctx) return RC_ERR_INVALID_ARGUMENT;
    RC_LOCK(ctx);
    HandleInfo *hi = nullptr;
    int32_t rc = find_handle(ctx, handle, &hi);
    if (rc != RC_OK) return rc;
    if (!transforms) RC_FAIL(ctx, RC_ERR_INVALID_ARGUMENT, "rc_update_transforms: transforms is NULL");
    if (m != hi->count)
        RC_FAIL(ctx, RC_ERR_INVALID_ARGUMENT, "Transform count (" + std::to_string(m) + ") != instance count (" + std::to_string(hi->count) + ")");
    for (uint32_t i = 0; i < m; i++) {  // update_instance_transforms_offset_kernel!, kernels.jl:455-476
        rc_instance_desc &d = ctx->instances[hi->start + i];
        memcpy(d.transform, transforms + 12 * (size_t)i, 48);
        if (inv_transforms) memcpy(d.inv_transform, inv_transforms + 12 * (size_t)i, 48);
        else x_mat3x4_inverse(d.transform, d.inv_transform);
    }
    ctx->transforms_dirty = true;
    return RC_OK;
}

// Transforms produced on the device (a simulation kernel's output): staged through the host mirror, which stays the single
// source of truth for the instance list (refit uploads it, compaction reorders it), so this costs one 48 B/instance read-back.
int32_t rc_update_transforms_device(rc_context *ctx, uint32_t handle, const float *d_transforms, const float *d_inv_transforms, uint32_t m) {
    if (!ctx) return RC_ERR_INVALID_ARGUMENT;
    RC_ENTER(ctx);
    if (!d_transforms) RC_FAIL(ctx, RC_ERR_INVALID_ARGUMENT, "rc_update_transforms_device: transforms is NULL");
    {   // validate before touching the caller's buffer (a wrong count must not read 48 * m bytes of it)
        HandleInfo *hi = nullptr;
        int32_t rc = find_handle(ctx, handle, &hi);
        if (rc != RC_OK) return rc;
        if (m != hi->count)
            RC_FAIL(ctx, RC_ERR_INVALID_ARGUMENT, "Transform count (" + std::to_string(m) + ") != instance count (" + std::to_string(hi->count) + ")");
    }
    std::vector<float> xf(12 * (size_t)m), inv(d_inv_transforms ? 12 * (size_t)m : 0);
    if (m) {
        RC_CUDA(ctx, cudaMemcpyAsync(xf.data(), d_transforms, xf.size() * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        if (d_inv_transforms) RC_CUDA(ctx, cudaMemcpyAsync(inv.data(), d_inv_transforms, inv.size() * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        RC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return rc_update_transforms(ctx, handle, xf.data(), d_inv_transforms ? inv.data() : nullptr, m);
}

int32_t rc_update_geometry(rc_context *ctx, uint32_t handle, const float *verts, uint32_t n_faces, const uint32_t *face_meta, uint32_t flags) {  // :808-857
    if (!ctx) return RC_ERR_INVALID_ARGUMENT;
    RC_ENTER(ctx);
    HandleInfo *hi = nullptr;
    int32_t rc = find_handle(ctx, handle, &hi);
    if (rc != RC_OK) return rc;
    if (hi->count == 0) RC_FAIL(ctx, RC_ERR_INVALID_ARGUMENT, "Handle has no instances");
    uint32_t blas_idx = ctx->instances[hi->start].blas_index;
    ctx->last_update_refitted = false;
    RcDeviceBlas &old = ctx->blas[blas_idx - 1];
    if ((flags & RC_UPDATE_REFIT) && old.topo && verts && n_faces == old.n_faces_in) {
        // the refit kernel for mesh updates: same faces, moved vertices -> re-fit the kept radix tree (leaf boxes, bottom-up fit, collapse)
        const float *d_verts = verts;
        float *tmp_v = nullptr;
        ApiTemps tmp(ctx->stream);
        if (!(flags & RC_VERTS_ON_DEVICE)) {
            RC_CUDA(ctx, tmp.get(&tmp_v, sizeof(float) * 9 * (size_t)n_faces));
            RC_CUDA(ctx, cudaMemcpyAsync(tmp_v, verts, sizeof(float) * 9 * (size_t)n_faces, cudaMemcpyHostToDevice, ctx->stream));
            d_verts = tmp_v;
        }
        std::string err;
        bool refitted = false;
        cudaEventRecord(ctx->ev_t0, ctx->stream);
        if (!rc_refit_blas(ctx->stream, d_verts, n_faces, &old, &refitted, err)) RC_FAIL(ctx, builder_error_code(err), err);
        cudaEventRecord(ctx->ev_t1, ctx->stream);
        if (refitted) {
            if (cudaEventSynchronize(ctx->ev_t1) == cudaSuccess) cudaEventElapsedTime(&ctx->last_build_ms, ctx->ev_t0, ctx->ev_t1);
            if (old.normals) { cudaFreeAsync(old.normals, ctx->stream); old.normals = nullptr; }  // rc_update_geometry drops the normals
            ctx->last_update_refitted = true;
            ctx->dirty = true;  // the root box moved: the TLAS is rebuilt over the same BLAS table at the next sync (update! marks dirty, :855)
            return RC_OK;
        }
        // not refittable (the degenerate set changed; nothing was touched): fall through to a rebuild from the new soup
    }
    RcDeviceBlas nb;
    rc = build_blas_from(ctx, verts, n_faces, face_meta, flags | (old.nodes2 ? RC_BUILD_KEEP_BVH2 : 0u) | (old.topo ? RC_BUILD_ALLOW_REFIT : 0u), &nb);
    if (rc != RC_OK) {
        if (rc == RC_ERR_NO_VALID_TRIANGLES) ctx->last_error = "New geometry has no valid triangles";
        return rc;
    }
    rc_free_blas(&ctx->blas[blas_idx - 1], ctx->stream);
    ctx->blas[blas_idx - 1] = nb;
    ctx->dirty = true;
    return RC_OK;
}
int32_t rc_last_update_refitted(const rc_context *ctx) { return ctx && ctx->last_update_refitted ? 1 : 0; }
int32_t rc_set_build_flags(rc_context *ctx, uint32_t flags) {
    if (!ctx) return RC_ERR_INVALID_ARGUMENT;
    RC_LOCK(ctx);
    ctx->build_flags = flags & (RC_BUILD_KEEP_BVH2 | RC_BUILD_ALLOW_REFIT);
    return RC_OK;
}

// ---- serialised geometry (rc_build.cu, "Serialised BLAS") ----
int32_t rc_export_geometry(rc_context *ctx, uint32_t handle, void *blob, uint64_t capacity, uint64_t *size) {
    if (!ctx) return RC_ERR_INVALID_ARGUMENT;
    RC_ENTER(ctx);
    HandleInfo *hi = nullptr;
    int32_t rc = find_handle(ctx, handle, &hi);
    if (rc != RC_OK) return rc;
    if (hi->count == 0) RC_FAIL(ctx, RC_ERR_INVALID_ARGUMENT, "Handle has no instances");
    const RcDeviceBlas &B = ctx->blas[ctx->instances[hi->start].blas_index - 1];
    const uint64_t need = rc_blas_blob_bytes(B);
    if (size) *size = need;
    if (!blob) {
        if (!size) RC_FAIL(ctx, RC_ERR_INVALID_ARGUMENT, "rc_export_geometry: blob and size are both NULL");
        return RC_OK;  // size query
    }
    if (capacity < need) RC_FAIL(ctx, RC_ERR_INVALID_ARGUMENT, "rc_export_geometry: capacity " + std::to_string(capacity) + " < " + std::to_string(need) + " bytes");
    std::string err;
    if (!rc_blas_export(ctx->stream, B, blob, capacity, err)) RC_FAIL(ctx, RC_ERR_CUDA, err);
    return RC_OK;
}

// host-only: would rc_push_exported accept these bytes as far as header, size and payload hash go?  Needs neither a context nor a GPU;
// the message of a refusal is returned by rc_last_error(NULL).
int32_t rc_check_exported(const void *blob, uint64_t size, uint32_t *n_triangles, uint32_t *n_faces, uint32_t *has_normals) {
    std::string err;
    if (!rc_blas_blob_check(blob, size, n_triangles, n_faces, has_normals, err)) { g_create_error = err; return RC_ERR_INVALID_ARGUMENT; }
    return RC_OK;
}

int32_t rc_push_exported(rc_context *ctx, const void *blob, uint64_t size, const float *transforms, const float *inv_transforms, const uint32_t *instance_ids,
                         uint32_t m, uint32_t *handle_out) {
    if (!ctx) return RC_ERR_INVALID_ARGUMENT;
    RC_ENTER(ctx);
    if (!blob || !transforms || m == 0 || !handle_out) RC_FAIL(ctx, RC_ERR_INVALID_ARGUMENT, "rc_push_exported: blob, transforms, m >= 1 and handle_out are required");
    RcDeviceBlas b;
    std::string err;
    if (!rc_blas_import(ctx->stream, blob, size, &b, err)) {
        rc_free_blas(&b, ctx->stream);
        const bool cuda = err.find("cuda") == 0;  // CK() messages start with the failing call
        RC_FAIL(ctx, cuda ? (err.find("out of memory") != std::string::npos ? RC_ERR_OUT_OF_MEMORY : RC_ERR_CUDA) : RC_ERR_INVALID_ARGUMENT, err);
    }
    *handle_out = append_blas_with_instances(ctx, b, transforms, inv_transforms, instance_ids, m);
    return RC_OK;
}

// compact_instances!, :996-1065.  Handles are visited in ascending id order (Julia iterates its Dict in hash
// order, which no caller can rely on); unreferenced BLASes are freed and blas_index values remapped.
static void compact_instances(rc_context *ctx) {
    // in place: the list is kept in ascending handle order (push! appends under the largest id, this pass keeps the order), so a live
    // range only ever moves towards the front — no second page-locked buffer (allocating one costs milliseconds)
    RcInstanceVec &ni = ctx->instances;
    size_t w = 0;
    bool ordered = true;
    for (auto it = ctx->handles.begin(); it != ctx->handles.end(); ++it)
        if (!it->second.deleted) { ordered = ordered && it->second.start >= w; w += it->second.count; }
    if (ordered) {
        w = 0;
        for (auto it = ctx->handles.begin(); it != ctx->handles.end();) {
            if (it->second.deleted) { it = ctx->handles.erase(it); continue; }
            if (it->second.start != w && it->second.count) memmove(ni.data() + w, ni.data() + it->second.start, sizeof(rc_instance_desc) * it->second.count);
            it->second.start = (uint32_t)w;
            w += it->second.count;
            ++it;
        }
        ni.resize(w);
    } else {  // (not reachable through the API; kept so that a broken invariant costs time, not correctness)
        RcInstanceVec tmp;
        tmp.reserve(ni.size());
        for (auto it = ctx->handles.begin(); it != ctx->handles.end();) {
            if (it->second.deleted) { it = ctx->handles.erase(it); continue; }
            uint32_t ns = (uint32_t)tmp.size();
            tmp.insert(tmp.end(), ni.begin() + it->second.start, ni.begin() + it->second.start + it->second.count);
            it->second.start = ns;
            ++it;
        }
        ni.swap(tmp);
    }
    std::vector<uint32_t> remap(ctx->blas.size() + 1, 0);
    for (auto &d : ni) remap[d.blas_index] = 1;
    uint32_t used = 0;
    for (size_t b = 1; b < remap.size(); b++) used += remap[b];
    if (!ctx->blas.empty() && used < ctx->blas.size()) {
        std::vector<RcDeviceBlas> nb;
        uint32_t k = 0;
        for (size_t b = 1; b < remap.size(); b++) {
            if (remap[b]) { remap[b] = ++k; nb.push_back(ctx->blas[b - 1]); }
            else rc_free_blas(&ctx->blas[b - 1], ctx->stream);
        }
        for (auto &d : ni) d.blas_index = remap[d.blas_index];
        ctx->blas.swap(nb);
    }
}

// L2 residency hint (experiment switch, off by default: profiles/README.md r2 has the measurement).  RC_L2_PERSIST=1: the wide nodes of the
// largest geometry are marked persisting for the kernels of the context's stream (cudaAccessPolicyWindow; the ray / hit streams are
// already evict-first), up to the device's persisting-L2 carve-out.  Meant for one large mesh (C2: 64 MB of wide nodes share the 126 MB
// L2 with a 1 GiB ray / hit stream); an instanced scene's working set is a few MB and stays resident anyway.
static void configure_l2_window(rc_context *ctx) {
    static const int mode = [] { const char *e = getenv("RC_L2_PERSIST"); return e ? atoi(e) : 0; }();
    if (!mode) return;
    cudaStreamAttrValue v;
    memset(&v, 0, sizeof v);
    const RcDeviceBlas *big = nullptr;
    for (const RcDeviceBlas &B : ctx->blas)
        if (B.nodes4 && (!big || B.n > big->n)) big = &B;
    int max_persist = 0, max_window = 0;
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, ctx->device);
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, ctx->device);
    if (big && max_persist > 0 && max_window > 0) {
        const size_t bytes = sizeof(RcNode4) * ((size_t)big->n + 1), window = bytes < (size_t)max_window ? bytes : (size_t)max_window;
        const size_t carve = window < (size_t)max_persist ? window : (size_t)max_persist;
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve);
        v.accessPolicyWindow.base_ptr = (void *)big->nodes4;
        v.accessPolicyWindow.num_bytes = window;
        v.accessPolicyWindow.hitRatio = mode == 2 ? 0.6f : (float)((double)carve / (double)window);
        v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        if (getenv("RC_L2_PERSIST_VERBOSE")) fprintf(stderr, "rc: L2 window %zu B, carve-out %zu B (max %d), hit ratio %.2f\n", window, carve, max_persist, v.accessPolicyWindow.hitRatio);
    }
    cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &v);
    cudaGetLastError();
}

static int32_t rebuild(rc_context *ctx) {  // rebuild_bvh! :962-993 + build_flat_blas_arrays! :470-517 + rebuild_static_tlas! :930-959
    bool any_deleted = false;
    for (auto &kv : ctx->handles) any_deleted |= kv.second.deleted;
    if (any_deleted) compact_instances(ctx);
    std::vector<RcBlasPtrs> ptrs(ctx->blas.size());
    std::vector<float> roots(6 * ctx->blas.size());
    std::vector<RcFlatBlas> flat(ctx->blas.size());
    uint32_t off = 0, nodes = 0;
    for (size_t b = 0; b < ctx->blas.size(); b++) {
        const RcDeviceBlas &B = ctx->blas[b];
        ptrs[b] = RcBlasPtrs{B.nodes2, B.nodes4, B.tris, B.hull, B.n, 0, {B.sphere[0], B.sphere[1], B.sphere[2], B.sphere[3]}};
        memcpy(&roots[6 * b], B.root_aabb, 24);
        flat[b] = RcFlatBlas{B.tris, off, B.n};
        off += B.n;
        nodes += 2 * B.n - 1;
    }
    uint32_t n = (uint32_t)ctx->instances.size();
    std::string err;
    if (!rc_build_tlas(ctx->stream, ctx->instances.data(), n, ptrs, roots, &ctx->tlas, err)) RC_FAIL(ctx, builder_error_code(err), err);
    if (ctx->d_flat) { cudaFreeAsync(ctx->d_flat, ctx->stream); ctx->d_flat = nullptr; }
    // the reference drains the flat arrays when no instance is left (:969-978) but keeps every BLAS otherwise
    ctx->n_flat_blas = n == 0 && ctx->blas.empty() ? 0 : (uint32_t)flat.size();
    ctx->n_flat_prims = off;
    ctx->synced_blas_nodes = nodes;
    ctx->synced_tlas_nodes = n == 0 ? 0 : (2 * n - 1 > 1 ? 2 * n - 1 : 1);
    if (!flat.empty()) {
        RC_CUDA(ctx, cudaMallocAsync(&ctx->d_flat, sizeof(RcFlatBlas) * flat.size(), ctx->stream));
        RC_CUDA(ctx, cudaMemcpyAsync(ctx->d_flat, flat.data(), sizeof(RcFlatBlas) * flat.size(), cudaMemcpyHostToDevice, ctx->stream));
    }
    RC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->dirty = false;
    ctx->transforms_dirty = false;
    ctx->built = true;
    ctx->vf_map_built = false;
    configure_l2_window(ctx);
    return RC_OK;
}

int32_t rc_sync(rc_context *ctx, int32_t *action) {  // sync!, :894-921
    if (!ctx) return RC_ERR_INVALID_ARGUMENT;
    if (action) *action = RC_SYNC_NONE;
    RC_ENTER(ctx);
    if (!ctx->dirty && !ctx->transforms_dirty && ctx->built) return RC_OK;  // clean fast path: no GPU work, no sync (:898-900)
    if (ctx->dirty || !ctx->built) {
        int32_t rc = rebuild(ctx);
        if (rc != RC_OK) return rc;
        if (action) *action = RC_SYNC_REBUILD;
    } else {
        std::string err;  // refit_tlas!, :2197-2222
        if (!rc_refit_tlas(ctx->stream, ctx->instances.data(), (uint32_t)ctx->instances.size(), &ctx->tlas, err)) RC_FAIL(ctx, builder_error_code(err), err);
        ctx->transforms_dirty = false;
        if (action) *action = RC_SYNC_REFIT;
    }
    RC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // KA.synchronize, :919
    return RC_OK;
}

// ------------------------------------------------------------------------------------------------ introspection
int32_t rc_is_valid(const rc_context *ctx, uint32_t handle) {
    if (!ctx) return 0;
    RC_LOCK(ctx);
    auto it = ctx->handles.find(handle);
    return it != ctx->handles.end() && !it->second.deleted;
}
uint32_t rc_n_total_instances(const rc_context *ctx) {
    if (!ctx) return 0;
    RC_LOCK(ctx);
    return (uint32_t)ctx->instances.size();
}
uint32_t rc_n_instances(const rc_context *ctx) {  // :2391-2398
    if (!ctx) return 0;
    RC_LOCK(ctx);
    uint32_t pending = 0;
    for (auto &kv : ctx->handles)
        if (kv.second.deleted) pending += kv.second.count;
    return (uint32_t)ctx->instances.size() - pending;
}
uint32_t rc_n_instances_of(const rc_context *ctx, uint32_t handle) {
    if (!ctx) return 0;
    RC_LOCK(ctx);
    auto it = ctx->handles.find(handle);
    return (it == ctx->handles.end() || it->second.deleted) ? 0 : it->second.count;
}
uint32_t rc_n_geometries(const rc_context *ctx) {
    if (!ctx) return 0;
    RC_LOCK(ctx);
    return (uint32_t)ctx->blas.size();
}
int32_t rc_is_dirty(const rc_context *ctx, int32_t *dirty, int32_t *transforms_dirty) {
    if (!ctx) return RC_ERR_INVALID_ARGUMENT;
    RC_LOCK(ctx);
    if (dirty) *dirty = ctx->dirty;
    if (transforms_dirty) *transforms_dirty = ctx->transforms_dirty;
    return RC_OK;
}
int32_t rc_get_instances(const rc_context *cctx, uint32_t handle, rc_instance_desc *out) {
    rc_context *ctx = const_cast<rc_context *>(cctx);
    if (!ctx || !out) return RC_ERR_INVALID_ARGUMENT;
    RC_LOCK(ctx);
    HandleInfo *hi = nullptr;
    int32_t rc = find_handle(ctx, handle, &hi);
    if (rc != RC_OK) return rc;
    memcpy(out, ctx->instances.data() + hi->start, sizeof(rc_instance_desc) * hi->count);
    return RC_OK;
}
int32_t rc_world_bound(const rc_context *ctx, float out[6]) {
    if (!ctx || !out) return RC_ERR_INVALID_ARGUMENT;
    RC_LOCK(ctx);
    memcpy(out, ctx->tlas.root_aabb, 24);
    return RC_OK;
}
int32_t rc_wait(rc_context *ctx) {
    if (!ctx) return RC_ERR_INVALID_ARGUMENT;
    RC_ENTER(ctx);
    RC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    RC_CUDA(ctx, cudaStreamSynchronize(ctx->s_h2d));
    RC_CUDA(ctx, cudaStreamSynchronize(ctx->s_d2h));
    return RC_OK;
}
int32_t rc_sizes(const rc_context *ctx, uint32_t *tlas_nodes, uint32_t *blas_nodes, uint32_t *blas_prims, uint32_t *pending_deletes) {
    if (!ctx) return RC_ERR_INVALID_ARGUMENT;
    RC_LOCK(ctx);
    if (tlas_nodes) *tlas_nodes = ctx->synced_tlas_nodes;
    if (blas_nodes) *blas_nodes = ctx->synced_blas_nodes;
    if (blas_prims) *blas_prims = ctx->n_flat_prims;
    if (pending_deletes) {
        uint32_t p = 0;
        for (auto &kv : ctx->handles) p += kv.second.deleted ? 1u : 0u;
        *pending_deletes = p;
    }
    return RC_OK;
}

static int32_t read_nodes2(rc_context *ctx, const RcNode2 *d_nodes, uint32_t count, rc_bvh_node2 *out, uint32_t capacity) {
    if (capacity < count) RC_FAIL(ctx, RC_ERR_INVALID_ARGUMENT, "capacity too small");
    if (count == 0) return RC_OK;
    std::vector<RcNode2> tmp(count);
    RC_CUDA(ctx, cudaMemcpyAsync(tmp.data(), d_nodes, sizeof(RcNode2) * count, cudaMemcpyDeviceToHost, ctx->stream));
    RC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (uint32_t i = 0; i < count; i++) memcpy(&out[i], &tmp[i], sizeof(rc_bvh_node2));  // drop the pad word
    return RC_OK;
}
int32_t rc_read_tlas_nodes(rc_context *ctx, rc_bvh_node2 *out, uint32_t capacity) {
    if (!ctx || !out) return RC_ERR_INVALID_ARGUMENT;
    RC_ENTER(ctx);
    if (!ctx->built || ctx->dirty) RC_FAIL(ctx, RC_ERR_NOT_SYNCED, "call rc_sync first");
    return read_nodes2(ctx, ctx->tlas.nodes2, ctx->synced_tlas_nodes, out, capacity);
}
int32_t rc_read_blas_nodes(rc_context *ctx, uint32_t blas_index, rc_bvh_node2 *out, uint32_t capacity) {
    if (!ctx || !out) return RC_ERR_INVALID_ARGUMENT;
    RC_ENTER(ctx);
    if (blas_index < 1 || blas_index > ctx->blas.size()) RC_FAIL(ctx, RC_ERR_INVALID_ARGUMENT, "blas_index out of range");
    const RcDeviceBlas &B = ctx->blas[blas_index - 1];
    if (!B.nodes2) RC_FAIL(ctx, RC_ERR_INVALID_ARGUMENT, "this geometry was built without RC_BUILD_KEEP_BVH2: there is no BVH2 to read back");
    return read_nodes2(ctx, B.nodes2, 2 * B.n - 1, out, capacity);
}
int32_t rc_read_blas_order(rc_context *ctx, uint32_t blas_index, uint32_t *out, uint32_t capacity) {
    if (!ctx || !out) return RC_ERR_INVALID_ARGUMENT;
    RC_ENTER(ctx);
    if (blas_index < 1 || blas_index > ctx->blas.size()) RC_FAIL(ctx, RC_ERR_INVALID_ARGUMENT, "blas_index out of range");
    const RcDeviceBlas &B = ctx->blas[blas_index - 1];
    if (capacity < B.n) RC_FAIL(ctx, RC_ERR_INVALID_ARGUMENT, "capacity too small");
    std::vector<RcTri> tmp(B.n);
    RC_CUDA(ctx, cudaMemcpyAsync(tmp.data(), B.tris, sizeof(RcTri) * B.n, cudaMemcpyDeviceToHost, ctx->stream));
    RC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (uint32_t i = 0; i < B.n; i++) out[i] = tmp[i].prim_id;
    return RC_OK;
}
int32_t rc_read_blas_faces(rc_context *ctx, uint32_t blas_index, uint32_t *out, uint32_t capacity) {
    if (!ctx || !out) return RC_ERR_INVALID_ARGUMENT;
    RC_ENTER(ctx);
    if (blas_index < 1 || blas_index > ctx->blas.size()) RC_FAIL(ctx, RC_ERR_INVALID_ARGUMENT, "blas_index out of range");
    const RcDeviceBlas &B = ctx->blas[blas_index - 1];
    if (capacity < B.n) RC_FAIL(ctx, RC_ERR_INVALID_ARGUMENT, "capacity too small");
    std::vector<RcTri> tmp(B.n);
    RC_CUDA(ctx, cudaMemcpyAsync(tmp.data(), B.tris, sizeof(RcTri) * B.n, cudaMemcpyDeviceToHost, ctx->stream));
    RC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (uint32_t i = 0; i < B.n; i++) out[tmp[i].prim_id] = tmp[i].face_index;
    return RC_OK;
}
int32_t rc_get_instance_handles(const rc_context *ctx, uint32_t *out, uint32_t capacity) {
    if (!ctx || !out) return RC_ERR_INVALID_ARGUMENT;
    RC_LOCK(ctx);
    if (capacity < ctx->instances.size()) return RC_ERR_INVALID_ARGUMENT;
    for (auto &kv : ctx->handles)
        for (uint32_t i = 0; i < kv.second.count; i++) out[kv.second.start + i] = kv.first;
    return RC_OK;
}
uint32_t rc_blas_n_prims(const rc_context *ctx, uint32_t blas_index) {
    if (!ctx) return 0;
    RC_LOCK(ctx);
    if (blas_index < 1 || blas_index > ctx->blas.size()) return 0;
    return ctx->blas[blas_index - 1].n;
}

// ------------------------------------------------------------------------------------------------ queries
static int32_t ensure_capacity(rc_context *ctx, void **buf, size_t *cap, size_t bytes) {
    if (*cap >= bytes) return RC_OK;
    if (*buf) { cudaFree(*buf); *buf = nullptr; *cap = 0; }
    RC_CUDA(ctx, cudaMalloc(buf, bytes));
    *cap = bytes;
    return RC_OK;
}

// queue the read-back of the hard-error counter on the context stream; the caller's own final cudaStreamSynchronize completes it and
// overflow_result() then only looks at host memory (a per-ray caller pays one sync per trace, not two)
static inline void queue_overflow_read(rc_context *ctx) { cudaMemcpyAsync(ctx->h_err, ctx->d_overflow + 1, 4, cudaMemcpyDeviceToHost, ctx->stream); }
static int32_t overflow_result(rc_context *ctx) {
    const uint32_t ov = *ctx->h_err;
    if (ov) {
        *ctx->h_err = 0;
        cudaMemsetAsync(ctx->d_overflow + 1, 0, 4, ctx->stream);
        RC_FAIL(ctx, RC_ERR_STACK_OVERFLOW, "traversal stack overflow for " + std::to_string(ov) + " ray(s)");
    }
    return RC_OK;
}
static int32_t check_overflow(rc_context *ctx) {
    uint32_t ov = 0;
    RC_CUDA(ctx, cudaMemcpyAsync(&ov, ctx->d_overflow + 1, 4, cudaMemcpyDeviceToHost, ctx->stream));
    RC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ov) {
        cudaMemsetAsync(ctx->d_overflow + 1, 0, 4, ctx->stream);
        RC_FAIL(ctx, RC_ERR_STACK_OVERFLOW, "traversal stack overflow for " + std::to_string(ov) + " ray(s)");
    }
    return RC_OK;
}

static int32_t trace_common(rc_context *ctx, const rc_ray *rays, rc_hit *hits, uint64_t n, uint32_t flags, bool any) {
    if (!ctx) return RC_ERR_INVALID_ARGUMENT;
    RC_ENTER(ctx);
    if (n == 0) return RC_OK;
    if (!rays || !hits) RC_FAIL(ctx, RC_ERR_INVALID_ARGUMENT, "rays / hits is NULL");
    if (!ctx->built || ctx->dirty || ctx->transforms_dirty) RC_FAIL(ctx, RC_ERR_NOT_SYNCED, "TLAS has pending mutations: call rc_sync before tracing");
    const bool rays_dev = flags & RC_RAYS_ON_DEVICE, hits_dev = flags & RC_HITS_ON_DEVICE;
    RcTraceLaunch L;
    L.scene = make_scene(ctx);
    L.any = any;
    L.wide = !(flags & RC_MODE_REFERENCE_ORDER);
    if (!L.wide)
        for (const RcDeviceBlas &B : ctx->blas)
            if (!B.nodes2) RC_FAIL(ctx, RC_ERR_INVALID_ARGUMENT, "RC_MODE_REFERENCE_ORDER needs the reference-layout BVH2: build the geometry with RC_BUILD_KEEP_BVH2");
    L.count = flags & RC_COUNTERS;
    L.watertight = flags & RC_MODE_WATERTIGHT;
    L.zero_tmin = flags & RC_IGNORE_TMIN;
    if (L.zero_tmin && !L.wide) RC_FAIL(ctx, RC_ERR_INVALID_ARGUMENT, "RC_IGNORE_TMIN is not available together with RC_MODE_REFERENCE_ORDER");
    if (L.watertight && L.count) RC_FAIL(ctx, RC_ERR_INVALID_ARGUMENT, "RC_COUNTERS is not available together with RC_MODE_WATERTIGHT");
    L.work = ctx->d_work;
    L.ovf_cap = ctx->ovf_cap;
    L.counters = ctx->d_counters;
    L.overflow = ctx->d_overflow;
    L.max_blocks = ctx->max_blocks;
    std::string err;
    ctx->last_launches = 0;
    if (rays_dev && hits_dev) {
        L.rays = rays; L.hits = hits; L.n = n;
        cudaEventRecord(ctx->ev_t0, ctx->stream);
        if (!rc_launch_trace(ctx->stream, L, err)) RC_FAIL(ctx, RC_ERR_CUDA, err);
        cudaEventRecord(ctx->ev_t1, ctx->stream);
        ctx->last_launches = 1;
        if (flags & RC_NO_SYNC) return RC_OK;
        queue_overflow_read(ctx);
        RC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaEventElapsedTime(&ctx->last_ms, ctx->ev_t0, ctx->ev_t1);
        return overflow_result(ctx);
    }
    // host buffers: stage through device memory in chunks; H2D of chunk c+1, trace of chunk c and D2H of chunk c-1 overlap
    const rc_ray *d_rays = rays;
    rc_hit *d_hits = hits;
    if (!rays_dev) {
        int32_t rc = ensure_capacity(ctx, (void **)&ctx->d_rays, &ctx->cap_rays, n * sizeof(rc_ray));
        if (rc != RC_OK) return rc;
        d_rays = ctx->d_rays;
    }
    if (!hits_dev) {
        int32_t rc = ensure_capacity(ctx, (void **)&ctx->d_hits, &ctx->cap_hits, n * sizeof(rc_hit));
        if (rc != RC_OK) return rc;
        d_hits = ctx->d_hits;
    }
    // Chunk schedule: the copies of chunk c+1 / c-1 hide behind the trace of chunk c, but the first chunk's upload and the last chunk's trace +
    // download hide behind nothing, and every launch of the persistent kernel costs ~0.2 ms of ramp and tail (a ray is ~80 us of lane time).
    // So: large chunks in the middle (2 M rays; RC_HOST_CHUNK_RAYS), a geometric ramp from 128 K at the start and back down at the end.
    // Measured on C3, 10^8 rays from pinned host memory, on a box whose links give 1.53-1.56 Grays/s: fixed 1 M chunks 1.41, fixed 2 M
    // 1.46-1.47, fixed 4 M 1.45, this schedule 1.47 (profiles/README.md) — the ramps matter for calls of a few million rays, where a
    // fixed 2 M chunk would not overlap anything.
    static const uint64_t chunk_max = [] { const char *e = getenv("RC_HOST_CHUNK_RAYS"); uint64_t v = e ? strtoull(e, nullptr, 10) : 0; return v ? v : (1ull << 21); }();
    const uint64_t chunk_min = std::min<uint64_t>(chunk_max, 1ull << 17);
    cudaEventRecord(ctx->ev_t0, ctx->stream);
    int c = 0;
    uint64_t ramp = chunk_min;
    for (uint64_t off = 0, cn = 0; off < n; off += cn, c++, ramp = std::min(chunk_max, ramp * 2)) {
        const uint64_t rem = n - off;
        cn = ramp;
        if (rem <= cn) cn = rem;
        else if (rem < 2 * cn) cn = std::max(chunk_min, rem / 2);
        int e = c % rc_context::NEV;
        if (!rays_dev) {
            RC_CUDA(ctx, cudaMemcpyAsync((void *)(d_rays + off), rays + off, cn * sizeof(rc_ray), cudaMemcpyHostToDevice, ctx->s_h2d));
            RC_CUDA(ctx, cudaEventRecord(ctx->ev_h2d[e], ctx->s_h2d));
            RC_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_h2d[e], 0));
        }
        L.rays = d_rays + off; L.hits = d_hits + off; L.n = cn;
        if (!rc_launch_trace(ctx->stream, L, err)) RC_FAIL(ctx, RC_ERR_CUDA, err);
        ctx->last_launches++;
        if (!hits_dev) {
            RC_CUDA(ctx, cudaEventRecord(ctx->ev_k[e], ctx->stream));
            RC_CUDA(ctx, cudaStreamWaitEvent(ctx->s_d2h, ctx->ev_k[e], 0));
            RC_CUDA(ctx, cudaMemcpyAsync(hits + off, d_hits + off, cn * sizeof(rc_hit), cudaMemcpyDeviceToHost, ctx->s_d2h));
        }
    }
    cudaEventRecord(ctx->ev_t1, ctx->stream);
    queue_overflow_read(ctx);
    RC_CUDA(ctx, cudaStreamSynchronize(ctx->s_h2d));
    RC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    RC_CUDA(ctx, cudaStreamSynchronize(ctx->s_d2h));
    cudaEventElapsedTime(&ctx->last_ms, ctx->ev_t0, ctx->ev_t1);
    return overflow_result(ctx);
}

int32_t rc_trace_closest(rc_context *ctx, const rc_ray *rays, rc_hit *hits, uint64_t n, uint32_t flags) { return trace_common(ctx, rays, hits, n, flags, false); }
int32_t rc_trace_any(rc_context *ctx, const rc_ray *rays, rc_hit *hits, uint64_t n, uint32_t flags) { return trace_common(ctx, rays, hits, n, flags, true); }

int32_t rc_get_counters(rc_context *ctx, uint64_t out[6], int32_t reset) {
    if (!ctx || !out) return RC_ERR_INVALID_ARGUMENT;
    RC_ENTER(ctx);
    RcCounters c;
    RC_CUDA(ctx, cudaMemcpyAsync(&c, ctx->d_counters, sizeof c, cudaMemcpyDeviceToHost, ctx->stream));
    if (reset) RC_CUDA(ctx, cudaMemsetAsync(ctx->d_counters, 0, sizeof c, ctx->stream));
    RC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    out[0] = c.rays; out[1] = c.nodes; out[2] = c.box_tests; out[3] = c.tri_tests; out[4] = c.inst_entries; out[5] = c.max_stack;
    return RC_OK;
}
float rc_last_kernel_ms(const rc_context *ctx) { return ctx ? ctx->last_ms : 0.f; }
float rc_last_build_ms(const rc_context *ctx) { return ctx ? ctx->last_build_ms : 0.f; }
uint32_t rc_last_kernel_launches(const rc_context *ctx) { return ctx ? ctx->last_launches : 0; }

// ------------------------------------------------------------------------------------------------ analysis
static int32_t require_synced(rc_context *ctx) {
    if (!ctx->built || ctx->dirty || ctx->transforms_dirty) RC_FAIL(ctx, RC_ERR_NOT_SYNCED, "TLAS has pending mutations: call rc_sync first");
    return RC_OK;
}

static int32_t grid_common(rc_context *ctx, const float viewdir[3], uint32_t grid, rc_hit *hits, float *points, float *illum, uint32_t n_illum, double *centroid4) {
    RC_ENTER(ctx);
    int32_t rc = require_synced(ctx);
    if (rc != RC_OK) return rc;
    if (!viewdir || grid == 0) RC_FAIL(ctx, RC_ERR_INVALID_ARGUMENT, "viewdir / grid");
    if (grid > 65535u) RC_FAIL(ctx, RC_ERR_INVALID_ARGUMENT, "grid_size must be <= 65535 (grid^2 cells are indexed with 32 bits)");
    size_t n = (size_t)grid * grid;
    RcGridFrame f;
    rc_grid_frame(ctx->tlas.root_aabb, viewdir, grid, &f);
    rc_hit *d_hits = nullptr;
    float *d_points = nullptr, *d_illum = nullptr;
    double *d_cent = nullptr;
    ApiTemps tmp(ctx->stream);
    if (hits) RC_CUDA(ctx, tmp.get(&d_hits, n * sizeof(rc_hit)));
    if (points) RC_CUDA(ctx, tmp.get(&d_points, n * 3 * sizeof(float)));
    if (illum) {
        RC_CUDA(ctx, tmp.get(&d_illum, (size_t)n_illum * sizeof(float)));
        RC_CUDA(ctx, cudaMemsetAsync(d_illum, 0, (size_t)n_illum * sizeof(float), ctx->stream));
    }
    if (centroid4) {
        RC_CUDA(ctx, tmp.get(&d_cent, 4 * sizeof(double)));
        RC_CUDA(ctx, cudaMemsetAsync(d_cent, 0, 4 * sizeof(double), ctx->stream));
    }
    rc_launch_grid_trace(ctx->stream, make_scene(ctx), f, d_hits, d_points, d_illum, n_illum, d_cent, ctx->d_overflow + 1, ctx->max_blocks);
    ctx->last_launches = 1;
    if (hits) RC_CUDA(ctx, cudaMemcpyAsync(hits, d_hits, n * sizeof(rc_hit), cudaMemcpyDeviceToHost, ctx->stream));
    if (points) RC_CUDA(ctx, cudaMemcpyAsync(points, d_points, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (illum) RC_CUDA(ctx, cudaMemcpyAsync(illum, d_illum, (size_t)n_illum * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (centroid4) RC_CUDA(ctx, cudaMemcpyAsync(centroid4, d_cent, 4 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    RC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return check_overflow(ctx);
}

int32_t rc_hits_from_grid(rc_context *ctx, const float viewdir[3], uint32_t grid, rc_hit *hits, float *points) {
    if (!ctx || !hits) return RC_ERR_INVALID_ARGUMENT;
    return grid_common(ctx, viewdir, grid, hits, points, nullptr, 0, nullptr);
}

int32_t rc_get_illumination(rc_context *ctx, const float viewdir[3], uint32_t grid, float *out, uint32_t n_out) {
    if (!ctx || !out) return RC_ERR_INVALID_ARGUMENT;
    if (n_out == 0) return RC_OK;
    return grid_common(ctx, viewdir, grid, nullptr, nullptr, out, n_out, nullptr);
}

int32_t rc_get_centroid(rc_context *ctx, const float viewdir[3], uint32_t grid, float centroid[3], uint32_t *n_hits, float *points) {
    if (!ctx || !centroid) return RC_ERR_INVALID_ARGUMENT;
    size_t n = (size_t)grid * grid;
    std::vector<rc_hit> hits(points ? n : 0);
    std::vector<float> pts(points ? 3 * n : 0);
    double acc[4] = {0, 0, 0, 0};
    int32_t rc = grid_common(ctx, viewdir, grid, points ? hits.data() : nullptr, points ? pts.data() : nullptr, nullptr, 0, acc);
    if (rc != RC_OK) return rc;
    uint32_t cnt = (uint32_t)acc[3];
    if (n_hits) *n_hits = cnt;
    for (int k = 0; k < 3; k++) centroid[k] = cnt ? (float)(acc[k] / acc[3]) : NAN;  // mean of an empty collection is NaN
    if (points) {  // [hit.point for hit in hits if hit.hit] (:108), in cell order
        size_t w = 0;
        for (size_t k = 0; k < n; k++)
            if (hits[k].hit) { memcpy(points + 3 * w, &pts[3 * k], 12); w++; }
    }
    return RC_OK;
}

int32_t rc_view_factors(rc_context *ctx, uint32_t rays_per_triangle, uint64_t seed, uint32_t *out, uint32_t row_base, uint32_t n_rows, uint32_t flags,
                        uint64_t *skipped) {
    return rc_view_factors_strided(ctx, rays_per_triangle, seed, out, row_base, 1, n_rows, flags, skipped);
}

int32_t rc_view_factors_strided(rc_context *ctx, uint32_t rays_per_triangle, uint64_t seed, uint32_t *out, uint32_t row_base, uint32_t row_stride, uint32_t n_rows,
                                uint32_t flags, uint64_t *skipped) {
    if (!ctx || !out) return RC_ERR_INVALID_ARGUMENT;
    RC_ENTER(ctx);
    int32_t rc = require_synced(ctx);
    if (rc != RC_OK) return rc;
    uint32_t n_cols = ctx->n_flat_prims;
    if (skipped) *skipped = 0;
    if (n_cols == 0 || n_rows == 0) return RC_OK;
    if (row_stride == 0 || row_base >= n_cols || (uint64_t)row_base + (uint64_t)(n_rows - 1) * row_stride >= n_cols)
        RC_FAIL(ctx, RC_ERR_INVALID_ARGUMENT, "view_factors: row block exceeds the matrix");
    size_t bytes = (size_t)n_rows * n_cols * sizeof(uint32_t);
    uint32_t *d_out = out;
    ApiTemps tmp(ctx->stream);
    if (!(flags & RC_HITS_ON_DEVICE)) RC_CUDA(ctx, tmp.get(&d_out, bytes));
    if (!ctx->vf_map_built) {  // once per synced scene: which flat primitive carries row r's metadata
        if (ctx->d_vf_row_pos) { cudaFree(ctx->d_vf_row_pos); ctx->d_vf_row_pos = nullptr; }
        RC_CUDA(ctx, cudaMalloc(&ctx->d_vf_row_pos, sizeof(uint32_t) * ((size_t)n_cols + 2)));
        uint32_t *info = ctx->d_vf_row_pos + n_cols, h_info[2] = {0, 0};
        rc_launch_vf_row_map(ctx->stream, ctx->d_flat, ctx->n_flat_blas, ctx->n_flat_prims, n_cols, ctx->d_vf_row_pos, info);
        RC_CUDA(ctx, cudaMemcpyAsync(h_info, info, sizeof(h_info), cudaMemcpyDeviceToHost, ctx->stream));
        RC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->vf_map_built = true;
        ctx->vf_map_usable = h_info[0] == 0;  // duplicated metadata values: several triangles feed one row, fall back to the scan
        ctx->vf_out_of_range = h_info[1];
    }
    const uint32_t *row_pos = ctx->vf_map_usable ? ctx->d_vf_row_pos : nullptr;
    unsigned long long *d_skipped = nullptr;
    RC_CUDA(ctx, tmp.get(&d_skipped, 8));
    RC_CUDA(ctx, cudaMemsetAsync(d_skipped, 0, 8, ctx->stream));
    // the timed region starts before the matrix is zeroed: view_factors allocates-and-zeros its result (src/kernels.jl:74-78), and
    // clearing 9.9 GB (C4) is part of the job
    cudaEventRecord(ctx->ev_t0, ctx->stream);
    RC_CUDA(ctx, cudaMemsetAsync(d_out, 0, bytes, ctx->stream));
    rc_launch_view_factors(ctx->stream, make_scene(ctx), ctx->d_flat, ctx->n_flat_blas, ctx->n_flat_prims, rays_per_triangle, seed, row_base, n_rows, n_cols, d_out,
                           nullptr, d_skipped, ctx->d_overflow + 1, ctx->max_blocks, ctx->d_work, row_pos, row_stride);
    cudaEventRecord(ctx->ev_t1, ctx->stream);
    ctx->last_launches = 2;
    unsigned long long sk = 0;
    RC_CUDA(ctx, cudaMemcpyAsync(&sk, d_skipped, 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (!(flags & RC_HITS_ON_DEVICE)) RC_CUDA(ctx, cudaMemcpyAsync(out, d_out, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (flags & RC_NO_SYNC) return RC_OK;
    RC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->last_ms, ctx->ev_t0, ctx->ev_t1);
    if (row_pos && row_base == 0) sk = ctx->vf_out_of_range;  // the map never visits them; the scan counts them itself
    if (skipped) *skipped = sk;
    return check_overflow(ctx);
}

int32_t rc_view_factor_rays(rc_context *ctx, uint32_t rays_per_triangle, uint64_t seed, uint32_t row_base, uint32_t n_rows, rc_ray *out) {
    if (!ctx || !out) return RC_ERR_INVALID_ARGUMENT;
    RC_ENTER(ctx);
    int32_t rc = require_synced(ctx);
    if (rc != RC_OK) return rc;
    size_t n = (size_t)n_rows * rays_per_triangle;
    if (n == 0) return RC_OK;
    rc_ray *d_rays = nullptr;
    ApiTemps tmp(ctx->stream);
    RC_CUDA(ctx, tmp.get(&d_rays, n * sizeof(rc_ray)));
    RC_CUDA(ctx, cudaMemsetAsync(d_rays, 0, n * sizeof(rc_ray), ctx->stream));
    rc_launch_view_factors(ctx->stream, make_scene(ctx), ctx->d_flat, ctx->n_flat_blas, ctx->n_flat_prims, rays_per_triangle, seed, row_base, n_rows, ctx->n_flat_prims,
                           nullptr, d_rays, nullptr, ctx->d_overflow + 1, ctx->max_blocks, ctx->d_work, nullptr, 1);
    RC_CUDA(ctx, cudaMemcpyAsync(out, d_rays, n * sizeof(rc_ray), cudaMemcpyDeviceToHost, ctx->stream));
    RC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RC_OK;
}

int32_t rc_read_flat_metadata(rc_context *ctx, uint32_t *out, uint32_t capacity) {
    if (!ctx || !out) return RC_ERR_INVALID_ARGUMENT;
    RC_ENTER(ctx);
    int32_t rc = require_synced(ctx);
    if (rc != RC_OK) return rc;
    if (capacity < ctx->n_flat_prims) RC_FAIL(ctx, RC_ERR_INVALID_ARGUMENT, "capacity too small");
    if (ctx->n_flat_prims == 0) return RC_OK;
    uint32_t *d = nullptr;
    ApiTemps tmp(ctx->stream);
    RC_CUDA(ctx, tmp.get(&d, sizeof(uint32_t) * ctx->n_flat_prims));
    rc_launch_flat_metadata(ctx->stream, ctx->d_flat, ctx->n_flat_blas, ctx->n_flat_prims, d);
    RC_CUDA(ctx, cudaMemcpyAsync(out, d, sizeof(uint32_t) * ctx->n_flat_prims, cudaMemcpyDeviceToHost, ctx->stream));
    RC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RC_OK;
}

// ------------------------------------------------------------------------------------------------ collision (§8f row 1)
int32_t rc_collide_instances(rc_context *ctx, rc_contact_pair *contacts, uint64_t capacity, uint64_t *n_contacts) {
    if (!ctx || !n_contacts) return RC_ERR_INVALID_ARGUMENT;
    RC_ENTER(ctx);
    int32_t rc = require_synced(ctx);
    if (rc != RC_OK) return rc;
    *n_contacts = 0;
    uint32_t n = ctx->tlas.n;
    if (n == 0) return RC_OK;
    uint32_t *d_counts = nullptr, *d_excl = nullptr, *d_tile = nullptr, *d_total = nullptr;
    ApiTemps tmp(ctx->stream);
    RC_CUDA(ctx, tmp.get(&d_counts, sizeof(uint32_t) * n));
    RC_CUDA(ctx, tmp.get(&d_excl, sizeof(uint32_t) * n));
    RC_CUDA(ctx, tmp.get(&d_tile, sizeof(uint32_t) * ((n + 2047) / 2048)));
    RC_CUDA(ctx, tmp.get(&d_total, sizeof(uint32_t)));
    rc_collide_count(ctx->stream, ctx->tlas, d_counts, ctx->d_overflow + 1);
    rc_exclusive_scan_u32(ctx->stream, d_counts, d_excl, n, d_tile, d_total);
    uint32_t total = 0;
    RC_CUDA(ctx, cudaMemcpyAsync(&total, d_total, 4, cudaMemcpyDeviceToHost, ctx->stream));
    RC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *n_contacts = total;
    if (contacts && capacity >= total && total > 0) {
        rc_contact_pair *d_c = nullptr;
        RC_CUDA(ctx, tmp.get(&d_c, sizeof(rc_contact_pair) * (size_t)total));
        rc_collide_write(ctx->stream, ctx->tlas, d_counts, d_excl, d_c, ctx->d_overflow + 1);
        RC_CUDA(ctx, cudaMemcpyAsync(contacts, d_c, sizeof(rc_contact_pair) * (size_t)total, cudaMemcpyDeviceToHost, ctx->stream));
    }
    RC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return check_overflow(ctx);
}

int32_t rc_collide_instances_any(rc_context *ctx, uint32_t handle_a, uint32_t handle_b, int32_t *overlap) {
    if (!ctx || !overlap) return RC_ERR_INVALID_ARGUMENT;
    RC_ENTER(ctx);
    *overlap = 0;
    int32_t rc = require_synced(ctx);
    if (rc != RC_OK) return rc;
    HandleInfo *ha = nullptr, *hb = nullptr;
    if ((rc = find_handle(ctx, handle_a, &ha)) != RC_OK) return rc;
    if ((rc = find_handle(ctx, handle_b, &hb)) != RC_OK) return rc;
    uint32_t n = ctx->tlas.n;
    if (n == 0) return RC_OK;
    std::vector<float> boxes(8 * (size_t)n);  // RcBox = 8 floats; reference-identical world boxes, indexed by instance position
    RC_CUDA(ctx, cudaMemcpyAsync(boxes.data(), ctx->tlas.inst_boxes, sizeof(float) * 8 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    RC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (uint32_t ia = ha->start; ia < ha->start + ha->count; ia++)
        for (uint32_t ib = hb->start; ib < hb->start + hb->count; ib++) {
            const float *a = &boxes[8 * (size_t)ia], *b = &boxes[8 * (size_t)ib];
            if (a[4] >= b[0] && a[5] >= b[1] && a[6] >= b[2] && a[0] <= b[4] && a[1] <= b[5] && a[2] <= b[6]) { *overlap = 1; return RC_OK; }
        }
    return RC_OK;
}

// ---- wavefront stages around the trace (docs/src/wavefront-renderer.jl:185-362) --------------------------------------------
int32_t rc_set_normals(rc_context *ctx, uint32_t handle, const float *normals, uint32_t n_faces, uint32_t flags) {
    if (!ctx) return RC_ERR_INVALID_ARGUMENT;
    RC_ENTER(ctx);
    HandleInfo *hi = nullptr;
    int32_t rc = find_handle(ctx, handle, &hi);
    if (rc != RC_OK) return rc;
    if (hi->count == 0) RC_FAIL(ctx, RC_ERR_INVALID_ARGUMENT, "Handle has no instances");
    RcDeviceBlas &B = ctx->blas[ctx->instances[hi->start].blas_index - 1];
    if (!normals || n_faces != B.n_faces_in) RC_FAIL(ctx, RC_ERR_INVALID_ARGUMENT, "rc_set_normals: need 9 floats for each of the " + std::to_string(B.n_faces_in) + " submitted faces");
    const float *d_in = normals;
    float *tmp = nullptr;
    ApiTemps temps(ctx->stream);
    const size_t bytes = sizeof(float) * 9 * (size_t)n_faces;
    if (!(flags & RC_VERTS_ON_DEVICE)) {
        RC_CUDA(ctx, temps.get(&tmp, bytes));
        RC_CUDA(ctx, cudaMemcpyAsync(tmp, normals, bytes, cudaMemcpyHostToDevice, ctx->stream));
        d_in = tmp;
    }
    if (!B.normals) RC_CUDA(ctx, cudaMallocAsync(&B.normals, sizeof(float) * 9 * (size_t)B.n, ctx->stream));
    rc_launch_gather_normals(ctx->stream, B.tris, B.n, d_in, B.normals);
    RC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the caller may reuse `normals` on return
    return RC_OK;
}

// every BLAS gets a normal array (geometric normals where the caller set none) and the device pointer table follows ctx->blas
static int32_t ensure_normals(rc_context *ctx) {
    std::vector<const float *> want(ctx->blas.size());
    for (size_t b = 0; b < ctx->blas.size(); b++) {
        RcDeviceBlas &B = ctx->blas[b];
        if (!B.normals && B.n) {
            RC_CUDA(ctx, cudaMallocAsync(&B.normals, sizeof(float) * 9 * (size_t)B.n, ctx->stream));
            rc_launch_gather_normals(ctx->stream, B.tris, B.n, nullptr, B.normals);
        }
        want[b] = B.normals;
    }
    if (want != ctx->normal_ptrs || !ctx->d_normal_ptrs) {
        if (want.size() > ctx->normal_ptrs.capacity() || !ctx->d_normal_ptrs) {
            RC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            if (ctx->d_normal_ptrs) cudaFree(ctx->d_normal_ptrs);
            ctx->d_normal_ptrs = nullptr;
            ctx->normal_ptrs.reserve(want.size() * 2 + 8);
            RC_CUDA(ctx, cudaMalloc(&ctx->d_normal_ptrs, sizeof(float *) * ctx->normal_ptrs.capacity()));
        }
        ctx->normal_ptrs.assign(want.begin(), want.end());
        RC_CUDA(ctx, cudaMemcpyAsync(ctx->d_normal_ptrs, ctx->normal_ptrs.data(), sizeof(float *) * want.size(), cudaMemcpyHostToDevice, ctx->stream));
        RC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the host vector may be reassigned by the next call
    }
    return RC_OK;
}

static int32_t finish_stage(rc_context *ctx, uint32_t flags, uint32_t launches, bool traced) {
    cudaEventRecord(ctx->ev_t1, ctx->stream);
    ctx->last_launches = launches;
    RC_CUDA(ctx, cudaGetLastError());
    if (flags & RC_NO_SYNC) return RC_OK;
    RC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->last_ms, ctx->ev_t0, ctx->ev_t1);
    return traced ? check_overflow(ctx) : RC_OK;
}

static int32_t primary_common(rc_context *ctx, const RcCamera &cam, uint32_t width, uint32_t height, uint32_t n_samples, uint64_t seed, rc_ray *d_rays, uint32_t flags) {
    RC_ENTER(ctx);
    if (width == 0 || height == 0 || n_samples == 0) return RC_OK;
    if (!d_rays) RC_FAIL(ctx, RC_ERR_INVALID_ARGUMENT, "rays is NULL");
    cudaEventRecord(ctx->ev_t0, ctx->stream);
    rc_launch_primary_rays(ctx->stream, cam, width, height, n_samples, seed, d_rays);
    return finish_stage(ctx, flags, 1, false);
}

int32_t rc_generate_primary_rays(rc_context *ctx, uint32_t width, uint32_t height, uint32_t n_samples, const float camera_pos[3], float focal_length, float aspect,
                                 uint64_t seed, rc_ray *rays, uint32_t flags) {
    if (!ctx) return RC_ERR_INVALID_ARGUMENT;
    if (!camera_pos) RC_FAIL(ctx, RC_ERR_INVALID_ARGUMENT, "camera_pos is NULL");
    RcCamera cam;
    memset(&cam, 0, sizeof(cam));
    memcpy(cam.pos, camera_pos, 12);
    cam.forward[2] = focal_length;
    cam.half_width = aspect;
    cam.half_height = 1.0f;
    cam.lookat = 0;
    cam.jitter = (flags & RC_WAVE_NO_JITTER) ? 0 : 1;
    return primary_common(ctx, cam, width, height, n_samples, seed, rays, flags);
}

int32_t rc_generate_primary_rays_lookat(rc_context *ctx, uint32_t width, uint32_t height, uint32_t n_samples, const float camera_pos[3], const float right[3],
                                        const float up[3], const float forward[3], float half_width, float half_height, uint64_t seed, rc_ray *rays, uint32_t flags) {
    if (!ctx) return RC_ERR_INVALID_ARGUMENT;
    if (!camera_pos || !right || !up || !forward) RC_FAIL(ctx, RC_ERR_INVALID_ARGUMENT, "camera vectors are NULL");
    RcCamera cam;
    memcpy(cam.pos, camera_pos, 12);
    memcpy(cam.right, right, 12);
    memcpy(cam.up, up, 12);
    memcpy(cam.forward, forward, 12);
    cam.half_width = half_width;
    cam.half_height = half_height;
    cam.lookat = 1;
    cam.jitter = (flags & RC_WAVE_NO_JITTER) ? 0 : 1;
    return primary_common(ctx, cam, width, height, n_samples, seed, rays, flags);
}

static int32_t make_lights(rc_context *ctx, const float *lights, uint32_t n_lights, float bias, RcLights *out) {
    if (!lights || n_lights == 0 || n_lights > RC_MAX_LIGHTS) RC_FAIL(ctx, RC_ERR_INVALID_ARGUMENT, "need 1.." + std::to_string(RC_MAX_LIGHTS) + " lights");
    memset(out, 0, sizeof(*out));
    memcpy(out->pos, lights, 12 * (size_t)n_lights);
    out->n = n_lights;
    out->bias = bias;
    return RC_OK;
}

static int32_t shadow_source(rc_context *ctx, const rc_ray *rays, const rc_hit *hits, RcShadowSource *out) {
    if (!rays || !hits) RC_FAIL(ctx, RC_ERR_INVALID_ARGUMENT, "rays / hits is NULL");
    int32_t rc = require_synced(ctx);
    if (rc != RC_OK) return rc;
    rc = ensure_normals(ctx);
    if (rc != RC_OK) return rc;
    *out = RcShadowSource{rays, hits, ctx->tlas.d_inst, ctx->d_normal_ptrs};
    return RC_OK;
}

int32_t rc_generate_shadow_rays(rc_context *ctx, const rc_ray *rays, const rc_hit *hits, uint64_t n, const float *lights, uint32_t n_lights, float shadow_bias,
                                rc_ray *shadow_rays, uint32_t flags) {
    if (!ctx) return RC_ERR_INVALID_ARGUMENT;
    RC_ENTER(ctx);
    if (n == 0) return RC_OK;
    RcLights L;
    RcShadowSource src;
    int32_t rc = make_lights(ctx, lights, n_lights, shadow_bias, &L);
    if (rc != RC_OK) return rc;
    if (!shadow_rays) RC_FAIL(ctx, RC_ERR_INVALID_ARGUMENT, "shadow_rays is NULL");
    rc = shadow_source(ctx, rays, hits, &src);
    if (rc != RC_OK) return rc;
    cudaEventRecord(ctx->ev_t0, ctx->stream);
    rc_launch_shadow_rays(ctx->stream, src, L, n, shadow_rays);
    return finish_stage(ctx, flags, 1, false);
}

int32_t rc_test_shadow_rays(rc_context *ctx, const rc_ray *shadow_rays, uint64_t n, uint8_t *visible, uint32_t flags) {
    if (!ctx) return RC_ERR_INVALID_ARGUMENT;
    RC_ENTER(ctx);
    if (n == 0) return RC_OK;
    if (!shadow_rays || !visible) RC_FAIL(ctx, RC_ERR_INVALID_ARGUMENT, "shadow_rays / visible is NULL");
    int32_t rc = require_synced(ctx);
    if (rc != RC_OK) return rc;
    cudaEventRecord(ctx->ev_t0, ctx->stream);
    rc_launch_test_shadow_rays(ctx->stream, make_scene(ctx), shadow_rays, n, visible, ctx->d_overflow, ctx->max_blocks, ctx->d_work);
    return finish_stage(ctx, flags, 2, true);
}

int32_t rc_shadow_visibility(rc_context *ctx, const rc_ray *rays, const rc_hit *hits, uint64_t n, const float *lights, uint32_t n_lights, float shadow_bias,
                             uint8_t *visible, uint32_t flags) {
    if (!ctx) return RC_ERR_INVALID_ARGUMENT;
    RC_ENTER(ctx);
    if (n == 0) return RC_OK;
    RcLights L;
    RcShadowSource src;
    int32_t rc = make_lights(ctx, lights, n_lights, shadow_bias, &L);
    if (rc != RC_OK) return rc;
    if (!visible) RC_FAIL(ctx, RC_ERR_INVALID_ARGUMENT, "visible is NULL");
    rc = shadow_source(ctx, rays, hits, &src);
    if (rc != RC_OK) return rc;
    cudaEventRecord(ctx->ev_t0, ctx->stream);
    rc_launch_shadow_visibility(ctx->stream, make_scene(ctx), src, L, n, visible, ctx->d_overflow, ctx->max_blocks, ctx->d_work);
    return finish_stage(ctx, flags, 2, true);
}


// ------------------------------------------------------------------------------------------------ BLAS4 (src/bvh4.jl)
// A BLAS4 is one geometry traversed on its own 4-wide BVH: the library's wide BVH is exactly that, so the object is a private context
// holding the geometry under one identity instance (the single-instance kernel variant skips the top level entirely).
struct rc_blas4 {
    rc_context *ctx = nullptr;
    std::string last_error;
};
static thread_local std::string g_blas4_error;
static_assert(sizeof(rc_wide_node) == sizeof(RcNode4), "rc_wide_node is the public name of the wide-node layout");

int32_t rc_blas4_build(int32_t device, const float *verts, uint32_t n_faces, const uint32_t *face_meta, uint32_t flags, rc_blas4 **out) {  // build_blas4, :511-523
    if (!out) return RC_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    rc_context *ctx = nullptr;
    int32_t rc = rc_create(device, &ctx);
    if (rc != RC_OK) { g_blas4_error = g_create_error; return rc; }
    const float ident[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    uint32_t h = 0;
    rc = rc_push(ctx, verts, n_faces, face_meta, ident, ident, nullptr, 1, flags & (RC_VERTS_ON_DEVICE | RC_BUILD_KEEP_BVH2 | RC_BUILD_ALLOW_REFIT), &h);
    if (rc == RC_OK) rc = rc_sync(ctx, nullptr);
    if (rc != RC_OK) {
        g_blas4_error = rc == RC_ERR_NO_VALID_TRIANGLES ? "Cannot build BLAS4 from empty primitive list" : ctx->last_error;  // :513
        rc_destroy(ctx);
        return rc;
    }
    rc_blas4 *b = new rc_blas4();
    b->ctx = ctx;
    *out = b;
    return RC_OK;
}
int32_t rc_blas4_destroy(rc_blas4 *b) {
    if (!b) return RC_OK;
    rc_destroy(b->ctx);
    delete b;
    return RC_OK;
}
const char *rc_blas4_last_error(const rc_blas4 *b) { return b ? b->ctx->last_error.c_str() : g_blas4_error.c_str(); }
rc_context *rc_blas4_context(rc_blas4 *b) { return b ? b->ctx : nullptr; }
int32_t rc_blas4_info(const rc_blas4 *b, uint32_t *n_primitives, uint32_t *n_node_slots, float root_aabb[6]) {
    if (!b) return RC_ERR_INVALID_ARGUMENT;
    RC_LOCK(b->ctx);
    const RcDeviceBlas &B = b->ctx->blas[0];
    if (n_primitives) *n_primitives = B.n;
    if (n_node_slots) *n_node_slots = B.n + 1;
    if (root_aabb) memcpy(root_aabb, B.root_aabb, 24);
    return RC_OK;
}
// closest_hit4 / any_hit4 (:606-766): ray.t_min is ignored as the reference does (ray_mint = 0, :610)
int32_t rc_blas4_trace_closest(rc_blas4 *b, const rc_ray *rays, rc_hit *hits, uint64_t n, uint32_t flags) {
    if (!b) return RC_ERR_INVALID_ARGUMENT;
    return trace_common(b->ctx, rays, hits, n, flags | RC_IGNORE_TMIN, false);
}
int32_t rc_blas4_trace_any(rc_blas4 *b, const rc_ray *rays, rc_hit *hits, uint64_t n, uint32_t flags) {
    if (!b) return RC_ERR_INVALID_ARGUMENT;
    return trace_common(b->ctx, rays, hits, n, flags | RC_IGNORE_TMIN, true);
}
int32_t rc_blas4_read_nodes(rc_blas4 *b, rc_wide_node *out, uint32_t capacity) {
    if (!b || !out) return RC_ERR_INVALID_ARGUMENT;
    rc_context *ctx = b->ctx;
    RC_ENTER(ctx);
    const RcDeviceBlas &B = ctx->blas[0];
    if (capacity < B.n + 1) RC_FAIL(ctx, RC_ERR_INVALID_ARGUMENT, "capacity too small");
    RC_CUDA(ctx, cudaMemcpyAsync(out, B.nodes4, sizeof(RcNode4) * ((size_t)B.n + 1), cudaMemcpyDeviceToHost, ctx->stream));
    RC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RC_OK;
}

// ------------------------------------------------------------------------------------------------ memory helpers
int32_t rc_device_alloc(rc_context *ctx, size_t bytes, void **out) {
    if (!ctx || !out) return RC_ERR_INVALID_ARGUMENT;
    RC_ENTER(ctx);
    RC_CUDA(ctx, cudaMalloc(out, bytes));
    return RC_OK;
}
int32_t rc_device_free(rc_context *ctx, void *ptr) {
    if (!ctx) return RC_ERR_INVALID_ARGUMENT;
    RC_ENTER(ctx);
    RC_CUDA(ctx, cudaFree(ptr));
    return RC_OK;
}
int32_t rc_host_alloc(rc_context *ctx, size_t bytes, void **out) {
    if (!ctx || !out) return RC_ERR_INVALID_ARGUMENT;
    RC_ENTER(ctx);
    RC_CUDA(ctx, cudaHostAlloc(out, bytes, cudaHostAllocPortable));  // pinned for every device of the process (rc_multi_* stage from it on all of them)
    return RC_OK;
}
int32_t rc_host_free(rc_context *ctx, void *ptr) {
    if (!ctx) return RC_ERR_INVALID_ARGUMENT;
    RC_CUDA(ctx, cudaFreeHost(ptr));
    return RC_OK;
}
int32_t rc_memcpy_h2d(rc_context *ctx, void *dst, const void *src, size_t bytes) {
    if (!ctx) return RC_ERR_INVALID_ARGUMENT;
    RC_ENTER(ctx);
    RC_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    RC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RC_OK;
}
int32_t rc_memcpy_d2h(rc_context *ctx, void *dst, const void *src, size_t bytes) {
    if (!ctx) return RC_ERR_INVALID_ARGUMENT;
    RC_ENTER(ctx);
    RC_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    RC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RC_OK;
}
int32_t rc_ipc_export(rc_context *ctx, void *ptr, uint8_t handle_out[64]) {
    if (!ctx || !ptr || !handle_out) return RC_ERR_INVALID_ARGUMENT;
    RC_ENTER(ctx);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    RC_CUDA(ctx, cudaIpcGetMemHandle(&h, ptr));
    memcpy(handle_out, &h, 64);
    return RC_OK;
}
int32_t rc_ipc_open(rc_context *ctx, const uint8_t handle[64], void **out) {
    if (!ctx || !handle || !out) return RC_ERR_INVALID_ARGUMENT;
    RC_ENTER(ctx);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    RC_CUDA(ctx, cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
    return RC_OK;
}
// Copy-engine gather for N > 4 GPUs: push a finished hit buffer into (a slice of) a peer-mapped buffer on the copy stream,
// ordered after everything already enqueued on the context stream, while the next trace runs.  slot selects one of two
// completion events so a double-buffering caller can make the context stream wait before it overwrites that buffer again.
int32_t rc_peer_copy_async(rc_context *ctx, void *dst, const void *src, size_t bytes, uint32_t slot) {
    if (!ctx || !dst || !src || slot > 1) return RC_ERR_INVALID_ARGUMENT;
    RC_ENTER(ctx);
    RC_CUDA(ctx, cudaEventRecord(ctx->ev_pc_src[slot], ctx->stream));
    RC_CUDA(ctx, cudaStreamWaitEvent(ctx->s_d2h, ctx->ev_pc_src[slot], 0));
    RC_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->s_d2h));
    RC_CUDA(ctx, cudaEventRecord(ctx->ev_pc_done[slot], ctx->s_d2h));
    ctx->copy_pending[slot] = true;
    return RC_OK;
}
int32_t rc_stream_wait_copy(rc_context *ctx, uint32_t slot) {
    if (!ctx || slot > 1) return RC_ERR_INVALID_ARGUMENT;
    RC_ENTER(ctx);
    if (ctx->copy_pending[slot]) RC_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_pc_done[slot], 0));
    return RC_OK;
}

int32_t rc_ipc_close(rc_context *ctx, void *ptr) {
    if (!ctx) return RC_ERR_INVALID_ARGUMENT;
    RC_ENTER(ctx);
    RC_CUDA(ctx, cudaIpcCloseMemHandle(ptr));
    return RC_OK;
}

}  // extern "C"
