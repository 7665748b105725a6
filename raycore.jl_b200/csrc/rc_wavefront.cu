// rc_wavefront.cu — the wavefront stages either side of the trace, on device-resident queues (SURVEY §8f row 2;
// reference: docs/src/wavefront-renderer.jl:185-362).  Stage 1 (primary rays) and stage 3 (shadow rays) are streaming kernels;
// stage 4 (occlusion test) and the fused stage 3+4 run on the scheduler kernel of rc_trace_fast.cuh through an IO policy, so a
// shadow ray is generated in the refill step and its result is one byte — the shadow-ray queue never exists in HBM.
#include "rc_trace_fast.cuh"
#include "rc_wave_core.cuh"
#include "rc_trace.h"
#include "rc_wave.h"

// ------------------------------------------------------------------------------------------------ normals
// normals_out[prim_id] = normals_in[face_index] (submitted-face order -> degenerate-filtered order of hit.primitive_id)
__global__ void k_gather_normals(const RcTri *__restrict__ tris, uint32_t n, const float *__restrict__ in, float *__restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * 9u) return;
    uint32_t pos = i / 9u, c = i % 9u;
    out[(size_t)tris[pos].prim_id * 9u + c] = in[(size_t)tris[pos].face_index * 9u + c];
}
__global__ void k_geometric_normals(const RcTri *__restrict__ tris, uint32_t n, float *__restrict__ out) {
    uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= n) return;
    RcTri t = tris[pos];
    f3 g = rc_geometric_normal(t);
    float *o = out + (size_t)t.prim_id * 9u;
    for (int k = 0; k < 3; k++) { o[3 * k] = g.x; o[3 * k + 1] = g.y; o[3 * k + 2] = g.z; }
}
void rc_launch_gather_normals(cudaStream_t st, const RcTri *tris, uint32_t n, const float *d_in, float *d_out) {
    if (n == 0) return;
    if (d_in) k_gather_normals<<<(n * 9u + 255) / 256, 256, 0, st>>>(tris, n, d_in, d_out);
    else k_geometric_normals<<<(n + 255) / 256, 256, 0, st>>>(tris, n, d_out);
}

// ------------------------------------------------------------------------------------------------ stage 1
__global__ void k_primary_rays(RcCamera cam, uint32_t width, uint32_t height, uint32_t n_samples, unsigned long long seed, rc_ray *__restrict__ rays) {
    const unsigned long long total = (unsigned long long)width * height * n_samples;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (unsigned long long)gridDim.x * blockDim.x) {
        rc_ray r = rc_primary_ray(cam, width, height, n_samples, seed, i);
        float4 *dst = reinterpret_cast<float4 *>(rays + i);
        dst[0] = make_float4(r.origin[0], r.origin[1], r.origin[2], r.tmin);
        dst[1] = make_float4(r.dir[0], r.dir[1], r.dir[2], r.tmax);
    }
}
void rc_launch_primary_rays(cudaStream_t st, const RcCamera &cam, uint32_t width, uint32_t height, uint32_t n_samples, unsigned long long seed, rc_ray *d_rays) {
    const unsigned long long total = (unsigned long long)width * height * n_samples;
    if (total == 0) return;
    unsigned long long want = (total + 255) / 256;
    int blocks = (int)(want < 148ull * 16 ? want : 148ull * 16);
    k_primary_rays<<<blocks, 256, 0, st>>>(cam, width, height, n_samples, seed, d_rays);
}

// ------------------------------------------------------------------------------------------------ stage 3
// shadow ray g = hit g / n_lights towards light g % n_lights ((idx-1)*NLights + light_idx in the reference's 1-based terms, :300)
__device__ __forceinline__ rc_ray shadow_ray_of(const RcShadowSource &s, const RcLights &lights, unsigned long long g) {
    const unsigned long long k = g / lights.n;
    const uint32_t l = (uint32_t)(g % lights.n);
    const float4 *hp = reinterpret_cast<const float4 *>(s.hits + k);
    const float4 h0 = __ldg(hp), h1 = __ldg(hp + 1);
    rc_hit h;
    h.hit = __float_as_uint(h0.x); h.t = h0.y; h.primitive_id = __float_as_uint(h0.z); h.instance_custom_index = __float_as_uint(h0.w);
    h.bary_u = h1.x; h.bary_v = h1.y; h.instance_id = __float_as_uint(h1.z); h.metadata = __float_as_uint(h1.w);
    if (!h.hit) return rc_dummy_shadow_ray();
    const rc_ray r = rc_load_ray(s.rays, k);
    const rc_instance_desc *inst = s.inst + h.instance_id;
    const float *nrm = s.blas_normals[inst->blas_index - 1u] + (size_t)h.primitive_id * 9u;
    float n9[9], inv[12];
#pragma unroll
    for (int c = 0; c < 9; c++) n9[c] = __ldg(nrm + c);
#pragma unroll
    for (int c = 0; c < 12; c++) inv[c] = __ldg(inst->inv_transform + c);
    return rc_shadow_ray(r, h, n9, inv, lights.pos[l], lights.bias);
}

__global__ void k_shadow_rays(RcShadowSource s, RcLights lights, unsigned long long n_hits, rc_ray *__restrict__ out) {
    const unsigned long long total = n_hits * lights.n;
    for (unsigned long long g = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (unsigned long long)gridDim.x * blockDim.x) {
        rc_ray r = shadow_ray_of(s, lights, g);
        float4 *dst = reinterpret_cast<float4 *>(out + g);
        dst[0] = make_float4(r.origin[0], r.origin[1], r.origin[2], r.tmin);
        dst[1] = make_float4(r.dir[0], r.dir[1], r.dir[2], r.tmax);
    }
}
void rc_launch_shadow_rays(cudaStream_t st, const RcShadowSource &s, const RcLights &lights, unsigned long long n_hits, rc_ray *d_out) {
    const unsigned long long total = n_hits * lights.n;
    if (total == 0) return;
    unsigned long long want = (total + 255) / 256;
    int blocks = (int)(want < 148ull * 16 ? want : 148ull * 16);
    k_shadow_rays<<<blocks, 256, 0, st>>>(s, lights, n_hits, d_out);
}

// ------------------------------------------------------------------------------------------------ stage 4 and fused 3+4
// test_shadow_rays! (:337-362): visible = t_max > 0 ? !any_hit(ray) : false.  A dummy ray (t_max = 0) is replaced by a ray no
// box accepts, so it retires at the TLAS root.
__device__ __forceinline__ rc_ray dead_ray() {
    rc_ray r;
    r.origin[0] = r.origin[1] = r.origin[2] = 0.f; r.dir[0] = 1.f; r.dir[1] = r.dir[2] = 0.f; r.tmin = 1.f; r.tmax = -1.f;
    return r;
}

// Both policies write visible[g] = (t_max > 0) when the ray is issued and clear it when the traversal reports an occluder: the
// same lane does both in program order, so no read-back of the ray (or re-generation) is needed at retirement.  A ray whose
// short stack overflowed is marked RC_VIS_RETRACE and redone by k_shadow_fixup with the deep-stack generic body.
#define RC_VIS_RETRACE 2

struct RcShadowQueueSource {
    const rc_ray *rays;
    __device__ __forceinline__ rc_ray get(unsigned long long g) const { return rc_load_ray(rays, g); }
};
struct RcShadowFusedSource {
    RcShadowSource src;
    RcLights lights;
    __device__ __forceinline__ rc_ray get(unsigned long long g) const { return shadow_ray_of(src, lights, g); }
};

template <class SRC>
struct RcIoShadow {
    static constexpr uint32_t kFetchMinMulti = RC_FETCH_MIN_MULTI, kTWMulti = RC_T_W_MULTI, kXWMulti = RC_X_W;  // scheduler constants, rc_trace_fast.cuh
    static constexpr uint32_t kFetchMinSingle = RC_FETCH_MIN_SINGLE, kTWSingle = RC_T_W_SINGLE;
    SRC source;
    uint8_t *visible;
    __device__ __forceinline__ rc_ray load(unsigned long long g) const {
        rc_ray r = source.get(g);
        const bool live = r.tmax > 0.0f;
        visible[g] = live ? 1 : 0;
        return live ? r : dead_ray();
    }
    __device__ __forceinline__ void store(unsigned long long g, rc_hit h) const {
        if (h.hit) visible[g] = h.hit == RC_OVERFLOW_MARK ? RC_VIS_RETRACE : 0;
    }
};

template <class SRC>
__global__ void __launch_bounds__(RC_TRACE_THREADS) k_shadow_fixup(RcScene sc, SRC source, unsigned long long n, uint8_t *__restrict__ visible, uint32_t *__restrict__ overflow) {
    if (*reinterpret_cast<volatile uint32_t *>(overflow + 2) == 0) return;  // no ray was flagged by the scheduler kernel
    for (unsigned long long g = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; g < n; g += (unsigned long long)gridDim.x * blockDim.x) {
        if (visible[g] != RC_VIS_RETRACE) continue;
        rc_ray r = source.get(g);
        rc_hit h;
        if (!rc_trace_wide<true, false>(sc, r, h, nullptr)) atomicAdd(overflow + 1, 1u);
        visible[g] = h.hit ? 0 : 1;
    }
}

// Empty TLAS (every handle deleted, then sync!): nothing can occlude, so a ray is visible iff it is a real ray (t_max > 0).  The
// scheduler kernel must not run here: there are no TLAS nodes to fetch (test/test_tlas_stress.jl:808-831 pins "empty => miss").
template <class SRC>
__global__ void k_shadow_empty(SRC source, unsigned long long n, uint8_t *__restrict__ visible) {
    for (unsigned long long g = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; g < n; g += (unsigned long long)gridDim.x * blockDim.x)
        visible[g] = source.get(g).tmax > 0.0f ? 1 : 0;
}

// overflow: the context's counter words ([1] hard errors, [2] rays flagged for the fix-up pass)
template <class SRC>
static void launch_shadow(cudaStream_t st, const RcScene &sc, const SRC &source, unsigned long long total, uint8_t *d_visible, uint32_t *overflow, int max_blocks,
                          unsigned long long *work) {
    if (sc.n_instances == 0) {
        const unsigned long long want0 = (total + 255) / 256;
        k_shadow_empty<SRC><<<(int)(want0 < 148ull * 16 ? want0 : 148ull * 16), 256, 0, st>>>(source, total, d_visible);
        return;
    }
    unsigned long long want = (total + RC_TRACE_THREADS - 1) / RC_TRACE_THREADS;
    const unsigned long long cap = sc.n_instances == 1u ? (unsigned long long)max_blocks * RC_MIN_BLOCKS_SINGLE / RC_MIN_BLOCKS : (unsigned long long)max_blocks;
    int blocks = (int)(want < cap ? want : cap);
    cudaMemsetAsync(work, 0, sizeof(unsigned long long), st);
    cudaMemsetAsync(overflow + 2, 0, sizeof(uint32_t), st);
    RcIoShadow<SRC> io{source, d_visible};
    if (sc.n_instances == 1u) k_trace_wide<true, false, RcIoShadow<SRC>, true><<<blocks, RC_TRACE_THREADS, 0, st>>>(sc, io, total, work, nullptr, overflow + 2);
    else k_trace_wide<true, false, RcIoShadow<SRC>, false><<<blocks, RC_TRACE_THREADS, 0, st>>>(sc, io, total, work, nullptr, overflow + 2);
    k_shadow_fixup<SRC><<<blocks, RC_TRACE_THREADS, 0, st>>>(sc, source, total, d_visible, overflow);
}

void rc_launch_test_shadow_rays(cudaStream_t st, const RcScene &sc, const rc_ray *d_rays, unsigned long long n, uint8_t *d_visible, uint32_t *overflow, int max_blocks,
                                unsigned long long *work) {
    if (n == 0) return;
    launch_shadow(st, sc, RcShadowQueueSource{d_rays}, n, d_visible, overflow, max_blocks, work);
}

void rc_launch_shadow_visibility(cudaStream_t st, const RcScene &sc, const RcShadowSource &s, const RcLights &lights, unsigned long long n_hits, uint8_t *d_visible,
                                 uint32_t *overflow, int max_blocks, unsigned long long *work) {
    const unsigned long long total = n_hits * lights.n;
    if (total == 0) return;
    launch_shadow(st, sc, RcShadowFusedSource{s, lights}, total, d_visible, overflow, max_blocks, work);
}
