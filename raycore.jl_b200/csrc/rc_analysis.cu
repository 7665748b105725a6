// rc_analysis.cu — fused ray-generation + trace + accumulate kernels for the analysis functions of
// src/kernels.jl: hits_from_grid / get_centroid / get_illumination (:58-72, :106-124) and view_factors! (:80-104).
#include <cuda_runtime.h>
#include <math.h>

#include <string>

#include "rc_trace.h"
#include "rc_trace_core.cuh"
#include "rc_trace_fast.cuh"

// ------------------------------------------------------------------------------------------------ grid
// Host-side restatement of generate_ray_grid's frame (src/kernels.jl:10-47).  Float32 arithmetic in the
// reference's order (this TU is compiled with -ffp-contract=off for host code).
static inline void h_normalize(const float a[3], float o[3]) {  // StaticArrays: inv(norm(a)) * a
    float n = sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    float inv = 1.0f / n;
    for (int k = 0; k < 3; k++) o[k] = inv * a[k];
}
static inline void h_cross(const float a[3], const float b[3], float o[3]) {
    float x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    o[0] = x; o[1] = y; o[2] = z;
}
static inline float h_dot(const float a[3], const float b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

bool rc_grid_frame(const float b[6], const float viewdir[3], uint32_t grid, RcGridFrame *f) {
    float d0[3], direction[3];
    h_normalize(viewdir, d0);     // hits_from_grid :59
    h_normalize(d0, direction);   // generate_ray_grid :11
    for (int k = 0; k < 3; k++) f->dir[k] = d0[k];
    float temp[3] = {1.0f, 0.0f, 0.0f};
    if (!(fabsf(direction[0]) < 0.9f)) { temp[0] = 0.0f; temp[1] = 1.0f; }
    float c1[3], c2[3];
    h_cross(direction, temp, c1); h_normalize(c1, f->basis1);
    h_cross(direction, f->basis1, c2); h_normalize(c2, f->basis2);
    float min1 = INFINITY, max1 = -INFINITY, min2 = INFINITY, max2 = -INFINITY, mind = INFINITY;
    for (int c = 0; c < 8; c++) {  // corners of the world box (bounds.jl:53-59); only extrema are used
        float p[3] = {(c & 1) ? b[3] : b[0], (c & 2) ? b[4] : b[1], (c & 4) ? b[5] : b[2]};
        float p1 = h_dot(p, f->basis1), p2 = h_dot(p, f->basis2), pd = h_dot(p, direction);
        min1 = fminf(min1, p1); max1 = fmaxf(max1, p1);
        min2 = fminf(min2, p2); max2 = fmaxf(max2, p2);
        mind = fminf(mind, pd);
    }
    float margin = 0.05f * fmaxf(max1 - min1, max2 - min2);
    float grid_width = max1 - min1 + 2 * margin;
    float grid_height = max2 - min2 + 2 * margin;
    float min_depth = mind - margin;
    float h1 = (min1 + max1) / 2, h2 = (min2 + max2) / 2;
    for (int k = 0; k < 3; k++) f->gc[k] = ((0.0f + min_depth * direction[k]) + h1 * f->basis1[k]) + h2 * f->basis2[k];
    f->cell_w = grid_width / (float)grid;
    f->cell_h = grid_height / (float)grid;
    f->grid = grid;
    return true;
}

// ray of cell k (Julia column-major: i = k % grid + 1, j = k / grid + 1); origin evaluated in Float64 then
// rounded, as the reference's mixed Int/Float64/Float32 expression does (:50-53)
__device__ __forceinline__ rc_ray grid_ray(const RcGridFrame &f, uint32_t k) {
    uint32_t i = k % f.grid + 1, j = k / f.grid + 1;
    double half = ((double)f.grid + 1.0) / 2.0;
    double u = __dmul_rn((double)i - half, (double)f.cell_w);
    double v = __dmul_rn((double)j - half, (double)f.cell_h);
    rc_ray r;
#pragma unroll
    for (int c = 0; c < 3; c++)
        r.origin[c] = (float)__dadd_rn(__dadd_rn((double)f.gc[c], __dmul_rn(u, (double)f.basis1[c])), __dmul_rn(v, (double)f.basis2[c]));
    r.dir[0] = f.dir[0]; r.dir[1] = f.dir[1]; r.dir[2] = f.dir[2];
    r.tmin = 0.0f;
    r.tmax = INFINITY;
    return r;
}

__global__ void k_grid_rays(RcGridFrame f, rc_ray *__restrict__ rays) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= f.grid * f.grid) return;
    rays[k] = grid_ray(f, k);
}

void rc_launch_grid_rays(cudaStream_t st, const RcGridFrame &f, rc_ray *rays) {
    uint32_t n = f.grid * f.grid;
    k_grid_rays<<<(n + 255) / 256, 256, 0, st>>>(f, rays);
}

__global__ void __launch_bounds__(RC_TRACE_THREADS) k_grid_trace(RcScene sc, RcGridFrame f, rc_hit *__restrict__ hits, float *__restrict__ points, float *__restrict__ illum,
                                                                 uint32_t n_illum, double *__restrict__ centroid_acc, uint32_t *__restrict__ overflow) {
    uint32_t n = f.grid * f.grid;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        rc_ray r = grid_ray(f, k);
        rc_hit h;
        const RcTri *tri = nullptr;
        if (!rc_trace_wide<false, false>(sc, r, h, nullptr, &tri)) atomicAdd(overflow, 1u);
        if (hits) hits[k] = h;
        float px = 0.f, py = 0.f, pz = 0.f;
        if (h.hit) {  // sum_mul(bary, prim.vertices) (math.jl:52), bary = (1-u-v, u, v) (:2015)
            float w = x_sub(x_sub(1.0f, h.bary_u), h.bary_v);
            px = x_add(x_add(x_mul(w, tri->v0[0]), x_mul(h.bary_u, tri->v1[0])), x_mul(h.bary_v, tri->v2[0]));
            py = x_add(x_add(x_mul(w, tri->v0[1]), x_mul(h.bary_u, tri->v1[1])), x_mul(h.bary_v, tri->v2[1]));
            pz = x_add(x_add(x_mul(w, tri->v0[2]), x_mul(h.bary_u, tri->v1[2])), x_mul(h.bary_v, tri->v2[2]));
            if (illum && h.metadata >= 1 && h.metadata <= n_illum) atomicAdd(&illum[h.metadata - 1], 1.0f);
            if (centroid_acc) {
                atomicAdd(&centroid_acc[0], (double)px);
                atomicAdd(&centroid_acc[1], (double)py);
                atomicAdd(&centroid_acc[2], (double)pz);
                atomicAdd(&centroid_acc[3], 1.0);
            }
        }
        if (points) { points[3 * (size_t)k] = px; points[3 * (size_t)k + 1] = py; points[3 * (size_t)k + 2] = pz; }
    }
}

void rc_launch_grid_trace(cudaStream_t st, const RcScene &sc, const RcGridFrame &f, rc_hit *hits, float *points, float *illum, uint32_t n_illum,
                          double *centroid_acc, uint32_t *overflow, int max_blocks) {
    uint32_t n = f.grid * f.grid;
    uint32_t want = (n + RC_TRACE_THREADS - 1) / RC_TRACE_THREADS;
    int blocks = (int)(want < (uint32_t)max_blocks ? want : (uint32_t)max_blocks);
    if (blocks < 1) blocks = 1;
    k_grid_trace<<<blocks, RC_TRACE_THREADS, 0, st>>>(sc, f, hits, points, illum, n_illum, centroid_acc, overflow);
}

// ------------------------------------------------------------------------------------------------ view factors
// One ray of view_factors! (src/kernels.jl:84-92) for source triangle `tri`: random_triangle_point (math.jl:158-174),
// origin offset 0.01*normal (:91), random_hemisphere_uniform (math.jl:125-141), frame from get_orthogonal_basis (:143-156).
// Randomness: counter RNG keyed by (seed, ray_index, dim) since Julia's task-local rand() is not reproducible.
__device__ __noinline__ rc_ray vf_make_ray(const RcTri *tri, unsigned long long seed, unsigned long long ray_index) {
    f3 p1 = mk3(tri->v0[0], tri->v0[1], tri->v0[2]), p2 = mk3(tri->v1[0], tri->v1[1], tri->v1[2]), p3 = mk3(tri->v2[0], tri->v2[1], tri->v2[2]);
    f3 normal = x_normalize(x_cross(x_sub3(p2, p1), x_sub3(p3, p1)));  // GB.orthogonal_vector ∝ (v2-v1)x(v3-v1)
    f3 n = x_normalize(normal);
    float ax = fabsf(normal.x), ay = fabsf(normal.y), az = fabsf(normal.z);
    int mi = 0;
    float best = ax;
    if (ay < best) { best = ay; mi = 1; }
    if (az < best) { best = az; mi = 2; }
    f3 cand = mk3(mi == 0 ? 1.f : 0.f, mi == 1 ? 1.f : 0.f, mi == 2 ? 1.f : 0.f);
    f3 vv = x_normalize(x_cross(n, cand));
    f3 uu = x_normalize(x_cross(vv, n));
    float r1 = rc_rng_uniform(seed, ray_index, 0), r2 = rc_rng_uniform(seed, ray_index, 1);
    float sq = x_sqrt(r1);
    float bu = x_sub(1.0f, sq), bv = x_mul(sq, x_sub(1.0f, r2)), bw = x_mul(sq, r2);
    f3 pt = mk3(x_add(x_add(x_mul(bu, p1.x), x_mul(bv, p2.x)), x_mul(bw, p3.x)), x_add(x_add(x_mul(bu, p1.y), x_mul(bv, p2.y)), x_mul(bw, p3.y)),
                x_add(x_add(x_mul(bu, p1.z), x_mul(bv, p2.z)), x_mul(bw, p3.z)));
    rc_ray r;
    r.origin[0] = x_add(pt.x, x_mul(normal.x, 0.01f));
    r.origin[1] = x_add(pt.y, x_mul(normal.y, 0.01f));
    r.origin[2] = x_add(pt.z, x_mul(normal.z, 0.01f));
    float xi1 = rc_rng_uniform(seed, ray_index, 2), xi2 = rc_rng_uniform(seed, ray_index, 3);
    float theta = acosf(xi1);
    float phi = x_mul(x_mul(2.0f, 3.14159265358979323846f), xi2);
    float st, ct, sp, cp;
    sincosf(theta, &st, &ct);
    sincosf(phi, &sp, &cp);
    float xl = x_mul(st, cp), yl = x_mul(st, sp), zl = ct;
    r.dir[0] = x_add(x_add(x_mul(uu.x, xl), x_mul(vv.x, yl)), x_mul(normal.x, zl));
    r.dir[1] = x_add(x_add(x_mul(uu.y, xl), x_mul(vv.y, yl)), x_mul(normal.y, zl));
    r.dir[2] = x_add(x_add(x_mul(uu.z, xl), x_mul(vv.z, yl)), x_mul(normal.z, zl));
    r.tmin = 0.0f;
    r.tmax = INFINITY;
    return r;
}

__device__ __forceinline__ const RcTri *flat_tri(const RcFlatBlas *flat, uint32_t n_blas, uint32_t pos) {
    uint32_t b = 0;
    while (b + 1 < n_blas && pos >= flat[b + 1].offset) b++;
    return flat[b].tris + (pos - flat[b].offset);
}

// Ray source / hit sink that turns the scheduler kernel into view_factors! (src/kernels.jl:80-104): "ray" g is ray g % rpt of the
// flat primitive g / rpt; it is generated on the fly in the refill step, and its result is one atomicAdd into the matrix block.
#ifndef RC_VF_FETCH_MIN
#define RC_VF_FETCH_MIN 16
#endif
#ifndef RC_VF_T_W
#define RC_VF_T_W 1u
#endif
#ifndef RC_VF_X_W
#define RC_VF_X_W 1u
#endif
struct RcIoViewFactors {
    // scheduler constants (rc_trace_fast.cuh): the refill generates the ray (RNG, point on the triangle, hemisphere direction), so it waits for more lanes
    static constexpr uint32_t kFetchMinMulti = RC_VF_FETCH_MIN, kTWMulti = RC_VF_T_W, kXWMulti = RC_VF_X_W;
    static constexpr uint32_t kFetchMinSingle = RC_VF_FETCH_MIN, kTWSingle = RC_VF_T_W;
    RcScene sc;
    const RcFlatBlas *flat;
    uint32_t n_blas, rpt, row_base, n_rows, n_cols;  // owned rows: row_base + k * row_stride, k < n_rows (output row k)
    unsigned long long seed;
    uint32_t *out;
    unsigned long long *skipped;
    uint32_t *overflow;       // hard errors (no stack could hold the ray)
    uint32_t *retrace_bits;   // one bit per ray: short stack overflowed
    const uint32_t *row_pos;  // nullable: row -> flat primitive (dense metadata)
    uint32_t row_stride;
    __device__ __forceinline__ bool owned(uint32_t row, uint32_t &local) const {
        const uint32_t d = row - row_base;
        local = d / row_stride;
        return row >= row_base && d % row_stride == 0u && local < n_rows;
    }
    // work item g -> (source triangle, its row).  With the row map (dense metadata: every row has exactly one triangle) the items are
    // the rays of the owned rows only; without it every flat primitive is visited and rows outside the block yield a dead ray.
    __device__ __forceinline__ const RcTri *locate(unsigned long long g, uint32_t &row, uint32_t &i) const {
        i = (uint32_t)(g % rpt);
        if (row_pos) {
            row = row_base + (uint32_t)(g / rpt) * row_stride;
            const uint32_t pos = __ldg(row_pos + row);
            return pos == RC_INVALID ? nullptr : flat_tri(flat, n_blas, pos);
        }
        const RcTri *tri = flat_tri(flat, n_blas, (uint32_t)(g / rpt));
        const uint32_t meta = tri->metadata;
        row = meta - 1u;
        if (meta < 1u || meta > n_cols) {
            if (i == 0 && skipped && row_base == 0) atomicAdd(skipped, 1ull);  // reference: unchecked index (:85,95-97)
            return nullptr;
        }
        uint32_t local;
        return owned(row, local) ? tri : nullptr;
    }
    __device__ __forceinline__ rc_ray load(unsigned long long g) const {
        uint32_t row, i;
        const RcTri *tri = locate(g, row, i);
        if (!tri) {
            rc_ray r;  // a ray no box can accept: retires after the TLAS root
            r.origin[0] = r.origin[1] = r.origin[2] = 0.f; r.dir[0] = 1.f; r.dir[1] = r.dir[2] = 0.f; r.tmin = 1.f; r.tmax = -1.f;
            return r;
        }
        return vf_make_ray(tri, seed, (unsigned long long)row * rpt + i);
    }
    __device__ __forceinline__ void store(unsigned long long g, rc_hit h) const {
        if (h.hit == RC_OVERFLOW_MARK) {  // short stack overflowed: k_view_factor_fixup redoes this ray with the deep-stack generic body
            atomicOr(&retrace_bits[g >> 5], 1u << (uint32_t)(g & 31u));
            return;
        }
        if (!h.hit) return;
        accumulate(g, h);
    }
    __device__ __forceinline__ void accumulate(unsigned long long g, const rc_hit &h) const {
        uint32_t row, local;
        if (row_pos) {
            local = (uint32_t)(g / rpt);
            row = row_base + local * row_stride;
        } else {
            const RcTri *tri = flat_tri(flat, n_blas, (uint32_t)(g / rpt));
            const uint32_t meta = tri->metadata;
            row = meta - 1u;
            if (meta < 1u || meta > n_cols || !owned(row, local)) return;
        }
        const uint32_t meta = row + 1u;
        if (h.hit && h.metadata != meta && h.metadata >= 1u && h.metadata <= n_cols) atomicAdd(&out[(size_t)local * n_cols + (h.metadata - 1u)], 1u);
    }
};

// Row map for view_factors: row_pos[meta-1] = flat position of the triangle carrying that metadata.  info[0] counts duplicates
// (two triangles with one metadata value: the map cannot be used), info[1] the triangles whose metadata is outside 1..n_cols.
__global__ void k_vf_row_map(const RcFlatBlas *__restrict__ flat, uint32_t n_blas, uint32_t n_prims, uint32_t n_cols, uint32_t *__restrict__ row_pos,
                             uint32_t *__restrict__ info) {
    uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= n_prims) return;
    const uint32_t meta = flat_tri(flat, n_blas, pos)->metadata;
    if (meta < 1u || meta > n_cols) { atomicAdd(&info[1], 1u); return; }
    if (atomicCAS(&row_pos[meta - 1u], RC_INVALID, pos) != RC_INVALID) atomicAdd(&info[0], 1u);
}
void rc_launch_vf_row_map(cudaStream_t st, const RcFlatBlas *d_flat, uint32_t n_blas, uint32_t n_prims, uint32_t n_cols, uint32_t *row_pos, uint32_t *info) {
    cudaMemsetAsync(row_pos, 0xFF, sizeof(uint32_t) * (size_t)n_cols, st);
    cudaMemsetAsync(info, 0, 2 * sizeof(uint32_t), st);
    if (n_prims) k_vf_row_map<<<(n_prims + 255) / 256, 256, 0, st>>>(d_flat, n_blas, n_prims, n_cols, row_pos, info);
}

// Deep-stack pass over the rays the scheduler kernel flagged (one bit per ray); exits at once when none was.
__global__ void __launch_bounds__(RC_TRACE_THREADS) k_view_factor_fixup(RcIoViewFactors io, unsigned long long total, const uint32_t *__restrict__ flagged) {
    if (*reinterpret_cast<const volatile uint32_t *>(flagged) == 0) return;
    const unsigned long long words = (total + 31) / 32;
    for (unsigned long long w = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; w < words; w += (unsigned long long)gridDim.x * blockDim.x) {
        uint32_t m = io.retrace_bits[w];
        while (m) {
            const unsigned long long g = w * 32 + (unsigned long long)(__ffs((int)m) - 1);
            m &= m - 1;
            rc_ray r = io.load(g);
            rc_hit h;
            if (!rc_trace_wide<false, false>(io.sc, r, h, nullptr)) atomicAdd(io.overflow, 1u);
            io.accumulate(g, h);
        }
    }
}

// the generated rays themselves (tests: the oracle traces exactly these)
__global__ void k_view_factor_rays(const RcFlatBlas *__restrict__ flat, uint32_t n_blas, uint32_t n_prims, uint32_t rpt, unsigned long long seed, uint32_t row_base,
                                   uint32_t n_rows, uint32_t n_cols, rc_ray *__restrict__ rays_out) {
    unsigned long long total = (unsigned long long)n_prims * rpt;
    for (unsigned long long g = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t pos = (uint32_t)(g / rpt), i = (uint32_t)(g % rpt);
        const RcTri *tri = flat_tri(flat, n_blas, pos);
        const uint32_t meta = tri->metadata, row = meta - 1u;
        if (meta < 1u || meta > n_cols || row < row_base || row >= row_base + n_rows) continue;
        rays_out[(size_t)(row - row_base) * rpt + i] = vf_make_ray(tri, seed, (unsigned long long)row * rpt + i);
    }
}

void rc_launch_view_factors(cudaStream_t st, const RcScene &sc, const RcFlatBlas *d_flat, uint32_t n_blas, uint32_t n_prims, uint32_t rpt, unsigned long long seed,
                            uint32_t row_base, uint32_t n_rows, uint32_t n_cols, uint32_t *out, rc_ray *rays_out, unsigned long long *skipped,
                            uint32_t *overflow, int max_blocks, unsigned long long *work, const uint32_t *row_pos, uint32_t row_stride) {
    unsigned long long total = (unsigned long long)(row_pos && !rays_out ? n_rows : n_prims) * rpt;
    if (total == 0) return;
    unsigned long long want = (total + RC_TRACE_THREADS - 1) / RC_TRACE_THREADS;
    const unsigned long long cap = (sc.n_instances == 1u && !rays_out) ? (unsigned long long)max_blocks * RC_MIN_BLOCKS_SINGLE / RC_MIN_BLOCKS : (unsigned long long)max_blocks;
    int blocks = (int)(want < cap ? want : cap);
    if (rays_out) {
        k_view_factor_rays<<<blocks, RC_TRACE_THREADS, 0, st>>>(d_flat, n_blas, n_prims, rpt, seed, row_base, n_rows, n_cols, rays_out);
        return;
    }
    if (sc.n_instances == 0) return;  // nothing to hit: the zeroed matrix is the answer
    uint32_t *bits = nullptr;
    const size_t bit_bytes = (size_t)((total + 31) / 32) * sizeof(uint32_t);
    cudaMallocAsync(&bits, bit_bytes, st);
    cudaMemsetAsync(bits, 0, bit_bytes, st);
    RcIoViewFactors io{sc, d_flat, n_blas, rpt, row_base, n_rows, n_cols, seed, out, skipped, overflow, bits, row_pos, row_stride ? row_stride : 1u};
    cudaMemsetAsync(work, 0, sizeof(unsigned long long), st);
    cudaMemsetAsync(overflow + 1, 0, sizeof(uint32_t), st);  // overflow = &d_overflow[1]; [2] counts the rays flagged for the fix-up pass
    if (sc.n_instances == 1u) k_trace_wide<false, false, RcIoViewFactors, true><<<blocks, RC_TRACE_THREADS, 0, st>>>(sc, io, total, work, nullptr, overflow + 1);
    else k_trace_wide<false, false, RcIoViewFactors, false><<<blocks, RC_TRACE_THREADS, 0, st>>>(sc, io, total, work, nullptr, overflow + 1);
    k_view_factor_fixup<<<blocks, RC_TRACE_THREADS, 0, st>>>(io, total, overflow + 1);
    cudaFreeAsync(bits, st);
}

__global__ void k_flat_metadata(const RcFlatBlas *__restrict__ flat, uint32_t n_blas, uint32_t n_prims, uint32_t *__restrict__ out) {
    uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= n_prims) return;
    out[pos] = flat_tri(flat, n_blas, pos)->metadata;
}

void rc_launch_flat_metadata(cudaStream_t st, const RcFlatBlas *d_flat, uint32_t n_blas, uint32_t n_prims, uint32_t *out) {
    if (n_prims == 0) return;
    k_flat_metadata<<<(n_prims + 255) / 256, 256, 0, st>>>(d_flat, n_blas, n_prims, out);
}
