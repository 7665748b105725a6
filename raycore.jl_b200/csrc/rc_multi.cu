// rc_multi.cu — several GPUs of one node behind one handle, inside the library ("one process, all devices").
//
// An rc_multi owns one rc_context per device.  The scene is REPLICATED: every mutation is applied to every device's context (the
// builder is deterministic, so the replicas are byte-identical; host inputs are uploaded once per device, concurrently), and the
// queries are SHARDED with no data-path collective:
//   rc_multi_trace_*      rays [k n / G, (k+1) n / G) go to device k.  Host buffers: every device stages its own slice straight from /
//                         to the caller's arrays over its own PCIe link (H2D, trace and D2H of all devices overlap; no hop through a
//                         root GPU).  Device buffers (resident on devices[0]): peer access is enabled between the devices and every
//                         device's traversal kernel reads its ray slice and stores its hit records through NVLink directly — the
//                         gather is the kernel's own epilogue.
//   rc_multi_view_factors source rows [k N / G, (k+1) N / G) on device k, each row block copied to its place in the caller's matrix.
// The reference's only concurrency is Threads.@threads over rays / source triangles (src/kernels.jl:64,82); this is the same split
// across devices.  Everything goes through the single-device C ABI (rc_api.cu), one host thread per device while a call is in flight.
#include <cuda_runtime.h>

#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/raycore_cuda.h"

struct rc_multi {
    std::vector<rc_context *> ctx;
    std::vector<int> devices;
    bool peer_ok = true;  // every device can map devices[0]'s memory (needed by the device-resident sharded trace only)
    std::string last_error;
};

namespace {
thread_local std::string g_multi_create_error;

// run fn(k) for every device on its own host thread; returns the first non-zero status (and remembers that device's message)
template <class F>
int32_t for_each_device(rc_multi *m, F fn) {
    const size_t g = m->ctx.size();
    std::vector<int32_t> rc(g, RC_OK);
    if (g == 1) {
        rc[0] = fn(0);
    } else {
        std::vector<std::thread> th;
        th.reserve(g);
        for (size_t k = 0; k < g; k++) th.emplace_back([&, k] { rc[k] = fn(k); });
        for (auto &t : th) t.join();
    }
    for (size_t k = 0; k < g; k++)
        if (rc[k] != RC_OK) {
            m->last_error = "device " + std::to_string(m->devices[k]) + ": " + rc_last_error(m->ctx[k]);
            return rc[k];
        }
    return RC_OK;
}
}  // namespace

extern "C" {

int32_t rc_multi_create(const int32_t *devices, uint32_t n_devices, rc_multi **out) {
    if (!out) return RC_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        g_multi_create_error = "no CUDA device available (libraycore_cuda has no CPU fallback)";
        return RC_ERR_CUDA;
    }
    rc_multi *m = new rc_multi();
    if (!devices || n_devices == 0) {  // all visible devices
        for (int d = 0; d < count; d++) m->devices.push_back(d);
    } else {
        for (uint32_t k = 0; k < n_devices; k++) {
            if (devices[k] < 0 || devices[k] >= count) { g_multi_create_error = "device index out of range"; delete m; return RC_ERR_INVALID_ARGUMENT; }
            for (uint32_t j = 0; j < k; j++)
                if (devices[j] == devices[k]) { g_multi_create_error = "a device is listed twice"; delete m; return RC_ERR_INVALID_ARGUMENT; }
            m->devices.push_back(devices[k]);
        }
    }
    for (int d : m->devices) {
        rc_context *c = nullptr;
        int32_t rc = rc_create(d, &c);
        if (rc != RC_OK) {
            g_multi_create_error = std::string("device ") + std::to_string(d) + ": " + rc_last_error(nullptr);
            for (rc_context *x : m->ctx) rc_destroy(x);
            delete m;
            return rc;
        }
        m->ctx.push_back(c);
    }
    // peer access towards devices[0] (and back), for the device-resident sharded trace; already-enabled is fine
    for (size_t k = 1; k < m->devices.size(); k++) {
        int can = 0;
        cudaDeviceCanAccessPeer(&can, m->devices[k], m->devices[0]);
        if (!can) { m->peer_ok = false; continue; }
        cudaSetDevice(m->devices[k]);
        cudaError_t e = cudaDeviceEnablePeerAccess(m->devices[0], 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) m->peer_ok = false;
        cudaSetDevice(m->devices[0]);
        e = cudaDeviceEnablePeerAccess(m->devices[k], 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) m->peer_ok = false;
        cudaGetLastError();
    }
    *out = m;
    return RC_OK;
}

int32_t rc_multi_destroy(rc_multi *m) {
    if (!m) return RC_OK;
    for (rc_context *c : m->ctx) rc_destroy(c);
    delete m;
    return RC_OK;
}

const char *rc_multi_last_error(const rc_multi *m) { return m ? m->last_error.c_str() : g_multi_create_error.c_str(); }
uint32_t rc_multi_device_count(const rc_multi *m) { return m ? (uint32_t)m->ctx.size() : 0; }
rc_context *rc_multi_context(rc_multi *m, uint32_t k) { return (m && k < m->ctx.size()) ? m->ctx[k] : nullptr; }

// ---- replicated mutation: the same call on every device; handles are identical because every context sees the same call sequence
int32_t rc_multi_push(rc_multi *m, const float *verts, uint32_t n_faces, const uint32_t *face_meta, const float *transforms, const float *inv_transforms,
                      const uint32_t *instance_ids, uint32_t n_instances, uint32_t flags, uint32_t *handle_out) {
    if (!m || !handle_out) return RC_ERR_INVALID_ARGUMENT;
    if (flags & RC_VERTS_ON_DEVICE) { m->last_error = "rc_multi_push takes host vertices (every device uploads its own copy)"; return RC_ERR_INVALID_ARGUMENT; }
    std::vector<uint32_t> h(m->ctx.size(), 0);
    int32_t rc = for_each_device(m, [&](size_t k) { return rc_push(m->ctx[k], verts, n_faces, face_meta, transforms, inv_transforms, instance_ids, n_instances, flags, &h[k]); });
    if (rc != RC_OK) return rc;
    for (uint32_t x : h)
        if (x != h[0]) { m->last_error = "replicas diverged: handle ids differ (a device context was mutated behind the multi handle)"; return RC_ERR_INVALID_ARGUMENT; }
    *handle_out = h[0];
    return RC_OK;
}
int32_t rc_multi_delete(rc_multi *m, uint32_t handle, int32_t *deleted) {
    if (!m) return RC_ERR_INVALID_ARGUMENT;
    int32_t d0 = 0;
    int32_t rc = for_each_device(m, [&](size_t k) { int32_t d = 0; int32_t r = rc_delete(m->ctx[k], handle, &d); if (k == 0) d0 = d; return r; });
    if (deleted) *deleted = d0;
    return rc;
}
int32_t rc_multi_update_transforms(rc_multi *m, uint32_t handle, const float *transforms, const float *inv_transforms, uint32_t n) {
    if (!m) return RC_ERR_INVALID_ARGUMENT;
    return for_each_device(m, [&](size_t k) { return rc_update_transforms(m->ctx[k], handle, transforms, inv_transforms, n); });
}
int32_t rc_multi_update_geometry(rc_multi *m, uint32_t handle, const float *verts, uint32_t n_faces, const uint32_t *face_meta, uint32_t flags) {
    if (!m) return RC_ERR_INVALID_ARGUMENT;
    if (flags & RC_VERTS_ON_DEVICE) { m->last_error = "rc_multi_update_geometry takes host vertices"; return RC_ERR_INVALID_ARGUMENT; }
    return for_each_device(m, [&](size_t k) { return rc_update_geometry(m->ctx[k], handle, verts, n_faces, face_meta, flags); });
}
int32_t rc_multi_sync(rc_multi *m, int32_t *action) {
    if (!m) return RC_ERR_INVALID_ARGUMENT;
    int32_t a0 = RC_SYNC_NONE;
    int32_t rc = for_each_device(m, [&](size_t k) { int32_t a = 0; int32_t r = rc_sync(m->ctx[k], &a); if (k == 0) a0 = a; return r; });
    if (action) *action = a0;
    return rc;
}

// ---- sharded queries
static int32_t multi_trace(rc_multi *m, const rc_ray *rays, rc_hit *hits, uint64_t n, uint32_t flags, bool any) {
    if (!m) return RC_ERR_INVALID_ARGUMENT;
    if (n == 0) return RC_OK;
    if (!rays || !hits) { m->last_error = "rays / hits is NULL"; return RC_ERR_INVALID_ARGUMENT; }
    const bool dev = flags & (RC_RAYS_ON_DEVICE | RC_HITS_ON_DEVICE);
    if (dev && !m->peer_ok && m->ctx.size() > 1) {
        m->last_error = "device-resident buffers need peer access between the devices; pass host buffers instead";
        return RC_ERR_INVALID_ARGUMENT;
    }
    const uint64_t g = m->ctx.size();
    return for_each_device(m, [&](size_t k) {
        const uint64_t lo = n * k / g, hi = n * (k + 1) / g;
        if (hi == lo) return (int32_t)RC_OK;
        rc_context *c = m->ctx[k];
        // (device buffers live on devices[0]: the other devices' kernels reach them through the peer mapping)
        return any ? rc_trace_any(c, rays + lo, hits + lo, hi - lo, flags & ~RC_NO_SYNC) : rc_trace_closest(c, rays + lo, hits + lo, hi - lo, flags & ~RC_NO_SYNC);
    });
}
int32_t rc_multi_trace_closest(rc_multi *m, const rc_ray *rays, rc_hit *hits, uint64_t n, uint32_t flags) { return multi_trace(m, rays, hits, n, flags, false); }
int32_t rc_multi_trace_any(rc_multi *m, const rc_ray *rays, rc_hit *hits, uint64_t n, uint32_t flags) { return multi_trace(m, rays, hits, n, flags, true); }

int32_t rc_multi_view_factors(rc_multi *m, uint32_t rays_per_triangle, uint64_t seed, uint32_t *out, uint64_t *skipped) {
    if (!m || !out) return RC_ERR_INVALID_ARGUMENT;
    uint32_t n_prims = 0;
    int32_t rc = rc_sizes(m->ctx[0], nullptr, nullptr, &n_prims, nullptr);
    if (rc != RC_OK) return rc;
    if (skipped) *skipped = 0;
    if (n_prims == 0) return RC_OK;
    const uint64_t g = m->ctx.size();
    std::vector<uint64_t> sk(g, 0);
    rc = for_each_device(m, [&](size_t k) {
        const uint32_t lo = (uint32_t)((uint64_t)n_prims * k / g), hi = (uint32_t)((uint64_t)n_prims * (k + 1) / g);
        if (hi == lo) return (int32_t)RC_OK;
        return rc_view_factors(m->ctx[k], rays_per_triangle, seed, out + (size_t)lo * n_prims, lo, hi - lo, 0, &sk[k]);
    });
    if (skipped) *skipped = sk[0];  // out-of-range metadata is reported by the block that starts at row 0
    return rc;
}

}  // extern "C"
