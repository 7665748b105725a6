// rc_build.cu — GPU LBVH builder for sm_100a: degenerate filter + compaction, scene bounds,
// 30-bit Morton codes, hand-written stable LSD radix sort, Karras radix tree, atomic bottom-up fit,
// reference-layout BVH2 emission and collapse to the quantised BVH4 the fast traversal uses.
//
// Replaces build_blas (src/instanced-bvh.jl:1376-1443), build_tlas_topology (:1485-1594),
// refit_tlas! (:2197-2222) and kernels K0-K11 of src/instanced-bvh-kernels.jl.  Everything runs on
// the caller's stream; the only device->host traffic is 4 B (valid-triangle count) + 24 B (root box)
// per BLAS and 24 B per TLAS build/refit.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <string>
#include <vector>

#include "rc_build.h"
#include "rc_build_core.cuh"

#define CK(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            err = std::string(#call) + ": " + cudaGetErrorString(e_);                             \
            return false;                                                                         \
        }                                                                                         \
    } while (0)

static inline uint32_t cdiv(uint64_t a, uint32_t b) { return (uint32_t)((a + b - 1) / b); }

// =================================================================================================
// Scan (exclusive, u32) — block tiles of 2048 + single-block scan of the tile sums
// =================================================================================================
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
    int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xFFFFFFFFu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// exclusive scan of one value per thread across the block; returns exclusive prefix, writes block total
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *smem /* >= 33 */, uint32_t &total) {
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    uint32_t inc = warp_incl_scan(v);
    if (lane == 31) smem[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t s = lane < nw ? smem[lane] : 0;
        uint32_t si = warp_incl_scan(s);
        smem[lane] = si - s;
        if (lane == 31) smem[32] = si;
    }
    __syncthreads();
    uint32_t r = smem[wid] + inc - v;
    total = smem[32];
    __syncthreads();
    return r;
}

__global__ void k_tile_sums(const uint32_t *__restrict__ in, uint32_t n, uint32_t *__restrict__ tile_sums) {
    __shared__ uint32_t sm[33];
    uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++)
        if (base + i < n) s += in[base + i];
    uint32_t total;
    block_excl_scan(s, sm, total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// single block: exclusive scan of `len` values in place; total -> *total_out
__global__ void k_scan_single(uint32_t *__restrict__ data, uint32_t len, uint32_t *__restrict__ total_out) {
    __shared__ uint32_t sm[33];
    uint32_t per = (len + blockDim.x - 1) / blockDim.x;
    uint32_t b = threadIdx.x * per, e = min(b + per, len);
    uint32_t s = 0;
    for (uint32_t i = b; i < e; i++) s += data[i];
    uint32_t total;
    uint32_t off = block_excl_scan(s, sm, total);
    for (uint32_t i = b; i < e; i++) {
        uint32_t v = data[i];
        data[i] = off;
        off += v;
    }
    if (threadIdx.x == 0 && total_out) *total_out = total;
}

__global__ void k_scan_apply(const uint32_t *__restrict__ in, uint32_t n, const uint32_t *__restrict__ tile_offs, uint32_t *__restrict__ out) {
    __shared__ uint32_t sm[33];
    uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS], s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        v[i] = (base + i < n) ? in[base + i] : 0;
        s += v[i];
    }
    uint32_t total;
    uint32_t off = block_excl_scan(s, sm, total) + tile_offs[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        if (base + i < n) out[base + i] = off;
        off += v[i];
    }
}

// out[i] = sum of in[0..i); *d_total = sum of all.  tile_tmp: >= cdiv(n, SCAN_TILE) words.
static void exclusive_scan_u32(cudaStream_t st, const uint32_t *in, uint32_t *out, uint32_t n, uint32_t *tile_tmp, uint32_t *d_total) {
    uint32_t tiles = cdiv(n, SCAN_TILE);
    k_tile_sums<<<tiles, SCAN_THREADS, 0, st>>>(in, n, tile_tmp);
    k_scan_single<<<1, 1024, 0, st>>>(tile_tmp, tiles, d_total);
    k_scan_apply<<<tiles, SCAN_THREADS, 0, st>>>(in, n, tile_tmp, out);
}

void rc_exclusive_scan_u32(cudaStream_t st, const uint32_t *in, uint32_t *out, uint32_t n, uint32_t *tile_tmp, uint32_t *d_total) {
    exclusive_scan_u32(st, in, out, n, tile_tmp, d_total);
}

// =================================================================================================
// Stable LSD radix sort of (u32 key, u32 value) pairs, 8 bits per pass.
//   per pass:  k_radix_hist   per-tile digit histogram            -> hist[digit * tiles + tile]
//              k_scan_single  exclusive scan of the 256*tiles table (digit-major => global offsets)
//              k_radix_scatter stable in-tile ranking (warp match_any) + scatter
// =================================================================================================
constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;
constexpr int RS_WARPS = RS_THREADS / 32;

__global__ void __launch_bounds__(RS_THREADS) k_radix_hist(const uint32_t *__restrict__ keys, uint32_t n, int shift, uint32_t tiles, uint32_t *__restrict__ hist) {
    __shared__ uint32_t sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    uint32_t base = blockIdx.x * RS_TILE;
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
        uint32_t idx = base + i * RS_THREADS + threadIdx.x;
        if (idx < n) atomicAdd(&sh[(keys[idx] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[threadIdx.x * tiles + blockIdx.x] = sh[threadIdx.x];
}

__global__ void __launch_bounds__(RS_THREADS) k_radix_scatter(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in, uint32_t *__restrict__ keys_out,
                                                             uint32_t *__restrict__ vals_out, uint32_t n, int shift, uint32_t tiles, const uint32_t *__restrict__ offs,
                                                             const uint32_t *__restrict__ totals) {
    __shared__ uint32_t wh[RS_WARPS][256];
    __shared__ uint32_t sm[33];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&wh[0][0])[i] = 0;
    __syncthreads();
    // warp w owns the contiguous chunk [w*256, (w+1)*256) of the tile; item i of lane l = chunk + i*32 + l,
    // so (i, l) lexicographic order == memory order and ranks are stable.
    const uint32_t base = blockIdx.x * RS_TILE + wid * (32 * RS_ITEMS);
    uint32_t key[RS_ITEMS], val[RS_ITEMS], rank[RS_ITEMS], dig[RS_ITEMS];
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
        uint32_t idx = base + i * 32 + lane;
        bool ok = idx < n;
        key[i] = ok ? keys_in[idx] : 0xFFFFFFFFu;
        val[i] = ok ? vals_in[idx] : 0u;
        dig[i] = ok ? ((key[i] >> shift) & 255u) : 256u;  // 256 = padding lane group
        uint32_t peers = __match_any_sync(0xFFFFFFFFu, dig[i]);
        uint32_t leader = __ffs(peers) - 1;
        uint32_t prev = 0;
        if (ok && lane == (int)leader) {
            prev = wh[wid][dig[i]];
            wh[wid][dig[i]] = prev + __popc(peers);
        }
        prev = __shfl_sync(0xFFFFFFFFu, prev, leader);
        rank[i] = prev + __popc(peers & lt_mask);
        __syncwarp();
    }
    uint32_t total_;
    const uint32_t digit_base = block_excl_scan(totals[threadIdx.x], sm, total_);  // keys with a smaller digit (includes a __syncthreads)
    {   // thread d: turn per-warp counts of digit d into exclusive prefixes starting at the global offset
        uint32_t d = threadIdx.x;
        uint32_t off = digit_base + offs[d * tiles + blockIdx.x];
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) {
            uint32_t c = wh[w][d];
            wh[w][d] = off;
            off += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
        if (dig[i] < 256u) {
            uint32_t pos = wh[wid][dig[i]] + rank[i];
            keys_out[pos] = key[i];
            vals_out[pos] = val[i];
        }
    }
}

// Row-wise exclusive scan of the digit-major table hist[256][tiles] in place, one warp per digit row (256 warps in 32 blocks):
// each lane sums a contiguous segment (independent loads), one warp scan orders the segments, the lane rewrites its segment.
// The row totals go to totals[256]; k_radix_scatter turns them into the per-digit bases itself (a 256-entry block scan), so no
// single-block pass over the whole table is needed (the one-block version was 30 us per pass at 1 M keys, a third of the build).
__global__ void __launch_bounds__(256) k_scan_rows(uint32_t *__restrict__ hist, uint32_t tiles, uint32_t *__restrict__ totals) {
    const uint32_t lane = threadIdx.x & 31, d = blockIdx.x * 8 + (threadIdx.x >> 5);
    uint32_t *row = hist + (size_t)d * tiles;
    const uint32_t seg = (tiles + 31) / 32, lo = min(lane * seg, tiles), hi = min(lo + seg, tiles);
    uint32_t sum = 0;
    for (uint32_t i = lo; i < hi; i++) sum += row[i];
    const uint32_t inc = warp_incl_scan(sum);
    uint32_t run = inc - sum;
    for (uint32_t i = lo; i < hi; i++) {
        const uint32_t v = row[i];
        row[i] = run;
        run += v;
    }
    if (lane == 31) totals[d] = inc;
}

// sorts in place: on return keys/vals hold the sorted pairs (4 passes ping-pong through tmp buffers)
static void radix_sort_pairs(cudaStream_t st, uint32_t *keys, uint32_t *vals, uint32_t *keys_tmp, uint32_t *vals_tmp, uint32_t n, uint32_t *hist /* 256*tiles + 256 */) {
    uint32_t tiles = cdiv(n, RS_TILE);
    uint32_t *ki = keys, *vi = vals, *ko = keys_tmp, *vo = vals_tmp;
    for (int pass = 0; pass < 4; pass++) {
        int shift = pass * 8;
        k_radix_hist<<<tiles, RS_THREADS, 0, st>>>(ki, n, shift, tiles, hist);
        uint32_t *totals = hist + (size_t)256 * tiles;
        k_scan_rows<<<32, 256, 0, st>>>(hist, tiles, totals);
        k_radix_scatter<<<tiles, RS_THREADS, 0, st>>>(ki, vi, ko, vo, n, shift, tiles, hist, totals);
        std::swap(ki, ko);
        std::swap(vi, vo);
    }
}

// =================================================================================================
// BLAS front end: filter, compact, bounds, Morton
// =================================================================================================
__device__ __forceinline__ f3 ld3(const float *p) { return mk3(p[0], p[1], p[2]); }

__global__ void k_face_flags(const float *__restrict__ verts, uint32_t n_faces, uint32_t *__restrict__ flags) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_faces) return;
    const float *v = verts + (size_t)i * 9;
    flags[i] = x_is_degenerate(ld3(v), ld3(v + 3), ld3(v + 6)) ? 0u : 1u;  // is_degenerate_face, :573-577
}

__device__ __forceinline__ void bounds_atomic(uint32_t *bounds, f3 lo, f3 hi) {
    // warp reduce (REDUX) then block reduce in ordered-uint space: 6 atomics per block (per-warp atomics on one 32-B sector
    // serialised in L2: 134 us for 1 M triangles in profiles/r1_launches_v6)
    __shared__ uint32_t part[6][32];
    uint32_t v[6] = {rc_float_to_ordered(lo.x), rc_float_to_ordered(lo.y), rc_float_to_ordered(lo.z),
                     rc_float_to_ordered(hi.x), rc_float_to_ordered(hi.y), rc_float_to_ordered(hi.z)};
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int c = 0; c < 6; c++) {
        uint32_t r = c < 3 ? __reduce_min_sync(0xFFFFFFFFu, v[c]) : __reduce_max_sync(0xFFFFFFFFu, v[c]);
        if (lane == 0) part[c][wid] = r;
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int c = 0; c < 6; c++) {
            uint32_t x = lane < nw ? part[c][lane] : (c < 3 ? 0xFFFFFFFFu : 0u);
            uint32_t r = c < 3 ? __reduce_min_sync(0xFFFFFFFFu, x) : __reduce_max_sync(0xFFFFFFFFu, x);
            if (lane == 0) {
                if (c < 3) atomicMin(&bounds[c], r);
                else atomicMax(&bounds[c], r);
            }
        }
    }
}

__global__ void k_init_bounds(uint32_t *bounds) {
    if (threadIdx.x < 3) bounds[threadIdx.x] = rc_float_to_ordered(INFINITY);  // Bounds3(): (+Inf, -Inf)
    else if (threadIdx.x < 6) bounds[threadIdx.x] = rc_float_to_ordered(-INFINITY);
}

// valid face i -> compacted slot pos[i]: unsorted RcTri (prim_id = slot, metadata), triangle box, scene bounds
__global__ void k_compact_faces(const float *__restrict__ verts, const uint32_t *__restrict__ face_meta, const uint32_t *__restrict__ flags,
                                const uint32_t *__restrict__ pos, uint32_t n_faces, RcTri *__restrict__ tris_in, RcBox *__restrict__ tri_boxes,
                                uint32_t *__restrict__ bounds) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    f3 lo = mk3(INFINITY, INFINITY, INFINITY), hi = mk3(-INFINITY, -INFINITY, -INFINITY);
    if (i < n_faces && flags[i]) {
        const float *v = verts + (size_t)i * 9;
        f3 a = ld3(v), b = ld3(v + 3), c = ld3(v + 6);
        uint32_t k = pos[i];
        RcTri t;
        t.v0[0] = a.x; t.v0[1] = a.y; t.v0[2] = a.z; t.prim_id = k;
        t.v1[0] = b.x; t.v1[1] = b.y; t.v1[2] = b.z; t.metadata = face_meta ? face_meta[i] : i + 1u;  // :595
        t.v2[0] = c.x; t.v2[1] = c.y; t.v2[2] = c.z; t.face_index = i;
        tris_in[k] = t;
        lo = jl_min3(jl_min3(a, b), c);  // world_bound(tri), triangle_mesh.jl:37
        hi = jl_max3(jl_max3(a, b), c);
        RcBox bx;
        bx.lo[0] = lo.x; bx.lo[1] = lo.y; bx.lo[2] = lo.z; bx.pad0 = 0;
        bx.hi[0] = hi.x; bx.hi[1] = hi.y; bx.hi[2] = hi.z; bx.pad1 = 0;
        tri_boxes[k] = bx;
    }
    bounds_atomic(bounds, lo, hi);
}

// calculate_morton_code_for_prim, kernels.jl:88-98 (extent unguarded, :1388)
__global__ void k_morton_prims(const RcBox *__restrict__ tri_boxes, uint32_t n, const uint32_t *__restrict__ bounds, uint32_t *__restrict__ codes, uint32_t *__restrict__ idx) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    f3 smin = mk3(rc_ordered_to_float(bounds[0]), rc_ordered_to_float(bounds[1]), rc_ordered_to_float(bounds[2]));
    f3 smax = mk3(rc_ordered_to_float(bounds[3]), rc_ordered_to_float(bounds[4]), rc_ordered_to_float(bounds[5]));
    f3 ext = x_sub3(smax, smin);
    RcBox b = tri_boxes[i];
    f3 c = mk3(x_mul(0.5f, x_add(b.lo[0], b.hi[0])), x_mul(0.5f, x_add(b.lo[1], b.hi[1])), x_mul(0.5f, x_add(b.lo[2], b.hi[2])));
    f3 nrm = mk3(x_div(x_sub(c.x, smin.x), ext.x), x_div(x_sub(c.y, smin.y), ext.y), x_div(x_sub(c.z, smin.z), ext.z));
    codes[i] = rc_morton30(nrm);
    idx[i] = i;
}

// sorted triangle j = tris_in[perm[j]]; also the bounding-sphere radius^2 about the centre of the scene bounds (bits of a
// non-negative float order like the float: one atomicMax per warp)
__global__ void k_gather_tris(const RcTri *__restrict__ tris_in, const uint32_t *__restrict__ perm, uint32_t n, RcTri *__restrict__ tris,
                              const uint32_t *__restrict__ bounds, uint32_t *__restrict__ r2_bits) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    float r2 = 0.0f;
    if (j < n) {
        const float4 *s = reinterpret_cast<const float4 *>(tris_in + perm[j]);
        float4 *d = reinterpret_cast<float4 *>(tris + j);
        const float4 a = s[0], b = s[1], c = s[2];
        d[0] = a; d[1] = b; d[2] = c;
        const f3 ctr = mk3(0.5f * (rc_ordered_to_float(bounds[0]) + rc_ordered_to_float(bounds[3])), 0.5f * (rc_ordered_to_float(bounds[1]) + rc_ordered_to_float(bounds[4])),
                           0.5f * (rc_ordered_to_float(bounds[2]) + rc_ordered_to_float(bounds[5])));
        r2 = rc_far2(ctr, mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), mk3(c.x, c.y, c.z));
    }
    const uint32_t m = __reduce_max_sync(0xFFFFFFFFu, r2 == r2 ? __float_as_uint(r2) : 0x7F800000u);  // NaN vertices: infinite radius (no cull)
    if ((threadIdx.x & 31u) == 0) atomicMax(r2_bits, m);
}

// =================================================================================================
// Topology, fit, BVH2 emission, collapse (shared by BLAS and TLAS)
// =================================================================================================
__global__ void k_topology(const uint32_t *__restrict__ codes, uint32_t n, RcTopo *__restrict__ topo, uint32_t *__restrict__ parent) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;  // internal node i+1
    if (i + 1 >= n) return;
    RcTopo t = rc_topology_for_node((int)(i + 1), codes, (int)n);
    topo[i] = t;
    parent[t.child0 - 1] = i + 1;  // set_parents_for_node, kernels.jl:159-180
    parent[t.child1 - 1] = i + 1;
    if (i == 0) parent[0] = RC_INVALID;
}

__device__ __forceinline__ RcBox ld_box_cg(const RcBox *p) {
    const float4 *q = reinterpret_cast<const float4 *>(p);
    float4 a = __ldcg(q), b = __ldcg(q + 1);
    RcBox r;
    r.lo[0] = a.x; r.lo[1] = a.y; r.lo[2] = a.z; r.pad0 = 0;
    r.hi[0] = b.x; r.hi[1] = b.y; r.hi[2] = b.z; r.pad1 = 0;
    return r;
}
__device__ __forceinline__ void st_box(RcBox *p, f3 lo, f3 hi) {
    float4 *q = reinterpret_cast<float4 *>(p);
    q[0] = make_float4(lo.x, lo.y, lo.z, 0.f);
    q[1] = make_float4(hi.x, hi.y, hi.z, 0.f);
}
__device__ __forceinline__ void st_node2(RcNode2 *p, f3 a0n, f3 a0x, f3 a1n, f3 a1x, uint32_t c0, uint32_t c1, uint32_t par) {
    float4 *q = reinterpret_cast<float4 *>(p);
    q[0] = make_float4(a0n.x, a0n.y, a0n.z, a0x.x);
    q[1] = make_float4(a0x.y, a0x.z, a1n.x, a1n.y);
    q[2] = make_float4(a1n.z, a1x.x, a1x.y, a1x.z);
    q[3] = make_float4(__uint_as_float(c0), __uint_as_float(c1), __uint_as_float(par), 0.f);
}

// One thread per leaf: write the leaf's own box + reference-layout leaf node, then climb; the second
// arriver at an internal node computes it (refit_aabbs_kernel!, kernels.jl:239-286 / :381-428).
//   BLAS (tris != null): leaf box = bounds of the sorted triangle, leaf node = (v0,v1,v2,0 | INVALID, p, parent)
//   TLAS (tris == null): leaf box = inst_boxes[leaf_map[p-1]],     leaf node = (lo,hi,0,0 | INVALID, inst, parent)
__global__ void k_fit(const RcTri *__restrict__ tris, const RcBox *__restrict__ inst_boxes, const uint32_t *__restrict__ leaf_map, uint32_t n,
                      const RcTopo *__restrict__ topo, const uint32_t *__restrict__ parent, uint32_t *__restrict__ flags, RcBox *__restrict__ boxes,
                      RcNode2 *__restrict__ nodes2) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;  // sorted primitive p+1
    if (p >= n) return;
    uint32_t leaf = n - 1 + (p + 1);
    uint32_t par = parent[leaf - 1];
    f3 lo, hi;
    if (tris) {
        const float4 *t = reinterpret_cast<const float4 *>(tris + p);
        float4 a = t[0], b = t[1], c = t[2];
        f3 v0 = mk3(a.x, a.y, a.z), v1 = mk3(b.x, b.y, b.z), v2 = mk3(c.x, c.y, c.z);
        lo = jl_min3(jl_min3(v0, v1), v2);  // get_node_aabb leaf branch, :1148-1158
        hi = jl_max3(jl_max3(v0, v1), v2);
        if (nodes2) st_node2(nodes2 + (leaf - 1), v0, v1, v2, mk3(0, 0, 0), RC_INVALID, p + 1, par);
    } else {
        uint32_t inst = leaf_map[p];
        RcBox b = inst_boxes[inst];
        lo = mk3(b.lo[0], b.lo[1], b.lo[2]);
        hi = mk3(b.hi[0], b.hi[1], b.hi[2]);
        if (nodes2) st_node2(nodes2 + (leaf - 1), lo, hi, mk3(0, 0, 0), mk3(0, 0, 0), RC_INVALID, inst, par);
    }
    st_box(boxes + (leaf - 1), lo, hi);
    uint32_t node = par;
    while (node != RC_INVALID) {
        __threadfence();  // publish the box written above before signalling
        uint32_t old = atomicAdd(&flags[node - 1], 1u);
        if (old == 0) return;  // first arriver: the sibling subtree is not ready
        RcTopo tp = topo[node - 1];
        RcBox b0 = ld_box_cg(boxes + (tp.child0 - 1)), b1 = ld_box_cg(boxes + (tp.child1 - 1));
        f3 l0 = mk3(b0.lo[0], b0.lo[1], b0.lo[2]), h0 = mk3(b0.hi[0], b0.hi[1], b0.hi[2]);
        f3 l1 = mk3(b1.lo[0], b1.lo[1], b1.lo[2]), h1 = mk3(b1.hi[0], b1.hi[1], b1.hi[2]);
        uint32_t up = parent[node - 1];
        if (nodes2) st_node2(nodes2 + (node - 1), l0, h0, l1, h1, tp.child0, tp.child1, up);
        st_box(boxes + (node - 1), jl_min3(l0, l1), jl_max3(h0, h1));  // get_node_aabb interior branch, :1142-1147
        node = up;
    }
}

__global__ void k_collapse(const RcBox *__restrict__ boxes, const RcTopo *__restrict__ topo, uint32_t n, uint32_t leaf_max, const uint32_t *__restrict__ leaf_map,
                           RcNode4 *__restrict__ nodes4) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;  // node i+1
    uint32_t n_int = n > 1 ? n - 1 : 1;                   // n == 1: synthetic root over the single leaf
    if (i >= n_int) return;
    RcNode4 nd = rc_collapse_node(i + 1, boxes, topo, n, leaf_max, leaf_map);
    const float4 *s = reinterpret_cast<const float4 *>(&nd);
    float4 *d = reinterpret_cast<float4 *>(nodes4 + (i + 1));
    d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; d[3] = s[3];
}

// RC_HULL_BOXES subtree boxes that together cover the BLAS: open the largest-area internal node until the budget is used.
// One warp, lane 0 does the (tiny) serial selection.
__global__ void k_blas_hull(const RcBox *__restrict__ boxes, const RcTopo *__restrict__ topo, uint32_t n, RcBox *__restrict__ hull) {
    if (threadIdx.x != 0) return;
    uint32_t ids[RC_HULL_BOXES];
    int cnt = 1;
    ids[0] = 1;
    while (cnt < RC_HULL_BOXES) {
        int best = -1;
        float best_area = -1.0f;
        for (int k = 0; k < cnt; k++)
            if (ids[k] < n) {
                float a = rc_half_area(boxes[ids[k] - 1]);
                if (a > best_area) { best_area = a; best = k; }
            }
        if (best < 0) break;
        RcTopo tp = topo[ids[best] - 1];
        ids[best] = tp.child0;
        ids[cnt++] = tp.child1;
    }
    for (int k = 0; k < RC_HULL_BOXES; k++) {
        RcBox b;
        if (k < cnt) b = boxes[ids[k] - 1];
        else { b.lo[0] = b.lo[1] = b.lo[2] = INFINITY; b.hi[0] = b.hi[1] = b.hi[2] = -INFINITY; b.pad0 = b.pad1 = 0; }
        hull[k] = b;
    }
}

// out6 = root box; sphere4 (nullable) = (centre of the scene bounds, radius^2 inflated against the rounding of its own evaluation)
__global__ void k_read_root(const RcBox *__restrict__ boxes, float *__restrict__ out6, const uint32_t *__restrict__ bounds = nullptr,
                            const uint32_t *__restrict__ r2_bits = nullptr, float *__restrict__ sphere4 = nullptr) {
    if (threadIdx.x < 3) out6[threadIdx.x] = boxes[0].lo[threadIdx.x];
    else if (threadIdx.x < 6) out6[threadIdx.x] = boxes[0].hi[threadIdx.x - 3];
    else if (sphere4 && threadIdx.x < 9) {
        const int k = threadIdx.x - 6;
        sphere4[k] = 0.5f * (rc_ordered_to_float(bounds[k]) + rc_ordered_to_float(bounds[3 + k]));
    } else if (sphere4 && threadIdx.x == 9) {
        sphere4[3] = __uint_as_float(*r2_bits) * 1.000002f;
    }
}

// codes (sorted) -> topology, fit, BVH2, BVH4
static void build_tree(cudaStream_t st, const uint32_t *codes_sorted, uint32_t n, const RcTri *tris, const RcBox *inst_boxes, const uint32_t *leaf_map,
                       uint32_t leaf_max, RcTopo *topo, uint32_t *parent, uint32_t *flags, RcBox *boxes, RcNode2 *nodes2, RcNode4 *nodes4,
                       const RcBox *inst_boxes_tight = nullptr, RcBox *boxes_tight = nullptr) {
    const int T = 256;
    if (n > 1) {
        k_topology<<<cdiv(n - 1, T), T, 0, st>>>(codes_sorted, n, topo, parent);
        cudaMemsetAsync(flags, 0, sizeof(uint32_t) * (n - 1), st);
    } else {
        cudaMemsetAsync(parent, 0xFF, sizeof(uint32_t), st);  // single leaf: parent = INVALID
    }
    k_fit<<<cdiv(n, T), T, 0, st>>>(tris, inst_boxes, leaf_map, n, topo, parent, flags, boxes, nodes2);
    if (inst_boxes_tight) {  // wide TLAS from the tighter instance bounds (same topology, second fit without BVH2 output)
        if (n > 1) cudaMemsetAsync(flags, 0, sizeof(uint32_t) * (n - 1), st);
        k_fit<<<cdiv(n, T), T, 0, st>>>(nullptr, inst_boxes_tight, leaf_map, n, topo, parent, flags, boxes_tight, nullptr);
        boxes = boxes_tight;
    }
    k_collapse<<<cdiv(n > 1 ? n - 1 : 1, T), T, 0, st>>>(boxes, topo, n, leaf_max, leaf_map, nodes4);
}

// refit only (topology kept): recompute boxes, BVH2 and BVH4
static void refit_tree(cudaStream_t st, uint32_t n, const RcBox *inst_boxes, const uint32_t *leaf_map, uint32_t leaf_max, const RcTopo *topo,
                       const uint32_t *parent, uint32_t *flags, RcBox *boxes, RcNode2 *nodes2, RcNode4 *nodes4, const RcBox *inst_boxes_tight,
                       RcBox *boxes_tight) {
    const int T = 256;
    if (n > 1) cudaMemsetAsync(flags, 0, sizeof(uint32_t) * (n - 1), st);
    k_fit<<<cdiv(n, T), T, 0, st>>>(nullptr, inst_boxes, leaf_map, n, topo, parent, flags, boxes, nodes2);
    if (n > 1) cudaMemsetAsync(flags, 0, sizeof(uint32_t) * (n - 1), st);
    k_fit<<<cdiv(n, T), T, 0, st>>>(nullptr, inst_boxes_tight, leaf_map, n, topo, parent, flags, boxes_tight, nullptr);
    k_collapse<<<cdiv(n > 1 ? n - 1 : 1, T), T, 0, st>>>(boxes_tight, topo, n, leaf_max, leaf_map, nodes4);
}

// =================================================================================================
// Public (library-internal) entry points
// =================================================================================================
void rc_free_blas(RcDeviceBlas *b, cudaStream_t st) {
    if (!b) return;
    if (b->nodes2) cudaFreeAsync(b->nodes2, st);
    if (b->nodes4) cudaFreeAsync(b->nodes4, st);
    if (b->tris) cudaFreeAsync(b->tris, st);
    if (b->hull) cudaFreeAsync(b->hull, st);
    if (b->normals) cudaFreeAsync(b->normals, st);
    b->normals = nullptr;
    b->nodes2 = nullptr; b->nodes4 = nullptr; b->tris = nullptr; b->hull = nullptr; b->n = 0;
}

// Stream-ordered temporaries that are returned to the pool when the builder leaves scope (also on every error path).
struct RcTemps {
    cudaStream_t st;
    std::vector<void *> ptrs;
    explicit RcTemps(cudaStream_t s) : st(s) {}
    ~RcTemps() {
        for (void *p : ptrs) cudaFreeAsync(p, st);
    }
    template <class T>
    bool get(T **out, size_t count, std::string &err) {
        void *p = nullptr;
        cudaError_t e = cudaMallocAsync(&p, sizeof(T) * (count ? count : 1), st);
        if (e != cudaSuccess) { err = std::string("cudaMallocAsync: ") + cudaGetErrorString(e); return false; }
        ptrs.push_back(p);
        *out = static_cast<T *>(p);
        return true;
    }
};
#define TMP(ptr, count) \
    if (!tmp.get(&(ptr), (count), err)) return false

// The wide nodes quantise against 2^(e-127) with e <= RC_QUANT_EXP_MAX: extents beyond 255 * 2^103 (where the reference's own
// Moeller-Trumbore already overflows) are refused instead of being traversed with boxes that do not cover them.
static bool extent_supported(const float aabb[6], std::string &err) {
    const float limit = 255.0f * 1.0141204801825835e31f;  // 255 * 2^103
    for (int k = 0; k < 3; k++) {
        const float e = aabb[3 + k] - aabb[k];
        if (e > limit) { err = "geometry extent exceeds the supported range (255 * 2^103)"; return false; }
    }
    return true;
}

bool rc_build_blas(cudaStream_t st, const float *d_verts, const uint32_t *d_face_meta, uint32_t n_faces, RcDeviceBlas *out, std::string &err) {
    *out = RcDeviceBlas();
    if (n_faces == 0) { err = "Geometry has no valid triangles"; return false; }
    const int T = 256;
    RcTemps tmp(st);
    uint32_t *d_flags = nullptr, *d_pos = nullptr, *d_tile = nullptr, *d_small = nullptr;
    TMP(d_flags, n_faces);
    TMP(d_pos, n_faces);
    TMP(d_tile, cdiv(n_faces, SCAN_TILE));
    TMP(d_small, 24);  // [0] = valid count, [1] = sphere radius^2 bits, [4..9] = scene bounds (ordered uints), [10..15] = root box, [16..19] = sphere
    k_face_flags<<<cdiv(n_faces, T), T, 0, st>>>(d_verts, n_faces, d_flags);
    exclusive_scan_u32(st, d_flags, d_pos, n_faces, d_tile, d_small);
    uint32_t n = 0;
    CK(cudaMemcpyAsync(&n, d_small, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (n == 0) { err = "Geometry has no valid triangles"; return false; }  // src/instanced-bvh.jl:601
    if (n > RC_LEAF_START_MASK - 16u) { err = "BLAS too large (max 2^28 triangles)"; return false; }
    uint32_t *d_bounds = d_small + 4;
    RcTri *d_tris_in = nullptr;
    RcBox *d_tri_boxes = nullptr, *d_boxes = nullptr;
    uint32_t *d_codes = nullptr, *d_idx = nullptr, *d_codes2 = nullptr, *d_idx2 = nullptr, *d_hist = nullptr, *d_parent = nullptr, *d_fl = nullptr;
    RcTopo *d_topo = nullptr;
    TMP(d_tris_in, n);
    TMP(d_tri_boxes, n);
    TMP(d_codes, n);
    TMP(d_idx, n);
    TMP(d_codes2, n);
    TMP(d_idx2, n);
    TMP(d_hist, 256 * (size_t)cdiv(n, RS_TILE) + 256);
    TMP(d_topo, n - 1);
    TMP(d_parent, 2 * (size_t)n - 1);
    TMP(d_fl, n - 1);
    TMP(d_boxes, 2 * (size_t)n - 1);
    // the results outlive this call; on failure the caller releases them with rc_free_blas
    CK(cudaMallocAsync(&out->nodes2, sizeof(RcNode2) * (2 * (size_t)n - 1), st));
    CK(cudaMallocAsync(&out->nodes4, sizeof(RcNode4) * ((size_t)n + 1), st));
    CK(cudaMallocAsync(&out->tris, sizeof(RcTri) * (size_t)n, st));
    CK(cudaMallocAsync(&out->hull, sizeof(RcBox) * RC_HULL_BOXES, st));
    out->n = n;
    out->n_faces_in = n_faces;

    k_init_bounds<<<1, 32, 0, st>>>(d_bounds);
    k_compact_faces<<<cdiv(n_faces, T), T, 0, st>>>(d_verts, d_face_meta, d_flags, d_pos, n_faces, d_tris_in, d_tri_boxes, d_bounds);
    k_morton_prims<<<cdiv(n, T), T, 0, st>>>(d_tri_boxes, n, d_bounds, d_codes, d_idx);
    radix_sort_pairs(st, d_codes, d_idx, d_codes2, d_idx2, n, d_hist);
    cudaMemsetAsync(d_small + 1, 0, 4, st);
    k_gather_tris<<<cdiv(n, T), T, 0, st>>>(d_tris_in, d_idx, n, out->tris, d_bounds, d_small + 1);
    build_tree(st, d_codes, n, out->tris, nullptr, nullptr, RC_BLAS_LEAF_MAX, d_topo, d_parent, d_fl, d_boxes, out->nodes2, out->nodes4);
    k_blas_hull<<<1, 32, 0, st>>>(d_boxes, d_topo, n, out->hull);
    k_read_root<<<1, 32, 0, st>>>(d_boxes, reinterpret_cast<float *>(d_small + 10), d_bounds, d_small + 1, reinterpret_cast<float *>(d_small + 16));
    float h_out[10];
    CK(cudaMemcpyAsync(h_out, d_small + 10, 40, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    memcpy(out->root_aabb, h_out, 24);
    memcpy(out->sphere, h_out + 6, 16);
    return extent_supported(out->root_aabb, err);
}


// ---------------------------------------------------------------------------------------------- TLAS
// instance world boxes + scene bounds (compute_instance_aabbs_kernel!, kernels.jl:65-78; host reduction :1499-1512)
__global__ void k_instance_boxes(const rc_instance_desc *__restrict__ inst, const float *__restrict__ blas_roots /* 6 per BLAS */, uint32_t n,
                                 RcBox *__restrict__ inst_boxes, uint32_t *__restrict__ bounds, const RcBlasPtrs *__restrict__ blas,
                                 RcBox *__restrict__ tight_boxes) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    f3 lo = mk3(INFINITY, INFINITY, INFINITY), hi = mk3(-INFINITY, -INFINITY, -INFINITY);
    if (i < n) {
        const rc_instance_desc *d = inst + i;
        rc_instance_world_aabb(d->transform, blas_roots + 6 * (d->blas_index - 1), lo, hi);
        st_box(inst_boxes + i, lo, hi);
        // tighter conservative bound for the wide TLAS: union of the transformed BLAS hull boxes (a rotated root box is up to
        // sqrt(3) wider per axis than the geometry it holds)
        const RcBox *hull = blas[d->blas_index - 1].hull;
        f3 tl = mk3(INFINITY, INFINITY, INFINITY), th = mk3(-INFINITY, -INFINITY, -INFINITY);
        for (int k = 0; k < RC_HULL_BOXES; k++) {
            RcBox hb = hull[k];
            if (!(hb.lo[0] <= hb.hi[0])) continue;
            float loc[6] = {hb.lo[0], hb.lo[1], hb.lo[2], hb.hi[0], hb.hi[1], hb.hi[2]};
            f3 a, b;
            rc_instance_world_aabb(d->transform, loc, a, b);
            tl = mk3(fminf(tl.x, a.x), fminf(tl.y, a.y), fminf(tl.z, a.z));
            th = mk3(fmaxf(th.x, b.x), fmaxf(th.y, b.y), fmaxf(th.z, b.z));
        }
        // never larger than the reference box
        tl = mk3(fmaxf(tl.x, lo.x), fmaxf(tl.y, lo.y), fmaxf(tl.z, lo.z));
        th = mk3(fminf(th.x, hi.x), fminf(th.y, hi.y), fminf(th.z, hi.z));
        st_box(tight_boxes + i, tl, th);
    }
    if (bounds) bounds_atomic(bounds, lo, hi);
}

// calculate_tlas_morton_code, kernels.jl:295-313; extent clamp :1517-1521
__global__ void k_morton_instances(const rc_instance_desc *__restrict__ inst, const float *__restrict__ blas_roots, uint32_t n, const uint32_t *__restrict__ bounds,
                                   uint32_t *__restrict__ codes, uint32_t *__restrict__ idx) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    f3 smin = mk3(rc_ordered_to_float(bounds[0]), rc_ordered_to_float(bounds[1]), rc_ordered_to_float(bounds[2]));
    f3 smax = mk3(rc_ordered_to_float(bounds[3]), rc_ordered_to_float(bounds[4]), rc_ordered_to_float(bounds[5]));
    f3 ext = mk3(jl_max(x_sub(smax.x, smin.x), 1e-6f), jl_max(x_sub(smax.y, smin.y), 1e-6f), jl_max(x_sub(smax.z, smin.z), 1e-6f));
    const rc_instance_desc *d = inst + i;
    const float *la = blas_roots + 6 * (d->blas_index - 1);
    f3 lc = mk3(x_mul(0.5f, x_add(la[0], la[3])), x_mul(0.5f, x_add(la[1], la[4])), x_mul(0.5f, x_add(la[2], la[5])));
    f3 wc = x_transform_point(d->transform, lc);
    f3 nrm = mk3(x_div(x_sub(wc.x, smin.x), ext.x), x_div(x_sub(wc.y, smin.y), ext.y), x_div(x_sub(wc.z, smin.z), ext.z));
    codes[i] = rc_morton30(nrm);
    idx[i] = i;
}

__global__ void k_instance_records(const rc_instance_desc *__restrict__ inst, const RcBlasPtrs *__restrict__ blas, uint32_t n, RcInstanceRec *__restrict__ rec,
                                   RcInstanceAux *__restrict__ aux) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const rc_instance_desc *d = inst + i;
    RcBlasPtrs b = blas[d->blas_index - 1];
    RcInstanceRec r;
    for (int k = 0; k < 12; k++) r.inv[k] = d->inv_transform[k];
    r.nodes4 = b.nodes4;
    r.tris = b.tris;
    for (int k = 0; k < 4; k++) { r.sphere[k] = b.sphere[k]; r.pad[k] = 0.f; }
    rec[i] = r;
    RcInstanceAux a;
    a.nodes2 = b.nodes2;
    a.n_prims = b.n;
    a.custom_index = d->instance_id;
    aux[i] = a;
}

void rc_free_tlas(RcDeviceTlas *t, cudaStream_t st) {
    if (!t) return;
    for (void *p : {(void *)t->nodes2, (void *)t->nodes4, (void *)t->rec, (void *)t->aux, (void *)t->d_inst, (void *)t->d_blas_roots, (void *)t->d_blas_ptrs,
                    (void *)t->inst_boxes, (void *)t->inst_boxes_tight, (void *)t->boxes_tight, (void *)t->leaf_map, (void *)t->topo, (void *)t->parent, (void *)t->flags, (void *)t->boxes, (void *)t->d_small})
        if (p) cudaFreeAsync(p, st);
    *t = RcDeviceTlas();
}

// Upload descriptors + rebuild records (used by both build and refit)
static bool upload_instances(cudaStream_t st, RcDeviceTlas *t, const rc_instance_desc *h_inst, uint32_t n, std::string &err) {
    CK(cudaMemcpyAsync(t->d_inst, h_inst, sizeof(rc_instance_desc) * n, cudaMemcpyHostToDevice, st));
    k_instance_records<<<cdiv(n, 256), 256, 0, st>>>(t->d_inst, t->d_blas_ptrs, n, t->rec, t->aux);
    return true;
}

bool rc_build_tlas(cudaStream_t st, const rc_instance_desc *h_inst, uint32_t n, const std::vector<RcBlasPtrs> &blas, const std::vector<float> &blas_roots,
                   RcDeviceTlas *t, std::string &err) {
    rc_free_tlas(t, st);
    for (int k = 0; k < 3; k++) { t->root_aabb[k] = INFINITY; t->root_aabb[3 + k] = -INFINITY; }  // Bounds3()
    t->n = n;
    if (n == 0) return true;  // empty TLAS: zero nodes (:969-978)
    const int T = 256;
    uint32_t nb = (uint32_t)blas.size();
    uint32_t rs_tiles = cdiv(n, RS_TILE);
    uint32_t *d_codes = nullptr, *d_codes2 = nullptr, *d_idx2 = nullptr, *d_hist = nullptr;
    RcTemps tmp(st);
    CK(cudaMallocAsync(&t->d_inst, sizeof(rc_instance_desc) * n, st));
    CK(cudaMallocAsync(&t->d_blas_roots, sizeof(float) * 6 * nb, st));
    CK(cudaMallocAsync(&t->d_blas_ptrs, sizeof(RcBlasPtrs) * nb, st));
    CK(cudaMallocAsync(&t->rec, sizeof(RcInstanceRec) * n, st));
    CK(cudaMallocAsync(&t->aux, sizeof(RcInstanceAux) * n, st));
    CK(cudaMallocAsync(&t->inst_boxes, sizeof(RcBox) * n, st));
    CK(cudaMallocAsync(&t->inst_boxes_tight, sizeof(RcBox) * n, st));
    CK(cudaMallocAsync(&t->boxes_tight, sizeof(RcBox) * (2 * n - 1), st));
    CK(cudaMallocAsync(&t->leaf_map, sizeof(uint32_t) * n, st));
    CK(cudaMallocAsync(&t->topo, sizeof(RcTopo) * std::max(1u, n - 1), st));
    CK(cudaMallocAsync(&t->parent, sizeof(uint32_t) * (2 * n - 1), st));
    CK(cudaMallocAsync(&t->flags, sizeof(uint32_t) * std::max(1u, n - 1), st));
    CK(cudaMallocAsync(&t->boxes, sizeof(RcBox) * (2 * n - 1), st));
    CK(cudaMallocAsync(&t->nodes2, sizeof(RcNode2) * (2 * n - 1), st));
    CK(cudaMallocAsync(&t->nodes4, sizeof(RcNode4) * (n + 1), st));
    CK(cudaMallocAsync(&t->d_small, sizeof(uint32_t) * 16, st));
    TMP(d_codes, n);
    TMP(d_codes2, n);
    TMP(d_idx2, n);
    TMP(d_hist, 256 * (size_t)rs_tiles + 256);
    CK(cudaMemcpyAsync(t->d_blas_roots, blas_roots.data(), sizeof(float) * 6 * nb, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(t->d_blas_ptrs, blas.data(), sizeof(RcBlasPtrs) * nb, cudaMemcpyHostToDevice, st));
    if (!upload_instances(st, t, h_inst, n, err)) return false;
    uint32_t *d_bounds = t->d_small + 4;
    k_init_bounds<<<1, 32, 0, st>>>(d_bounds);
    k_instance_boxes<<<cdiv(n, T), T, 0, st>>>(t->d_inst, t->d_blas_roots, n, t->inst_boxes, d_bounds, t->d_blas_ptrs, t->inst_boxes_tight);
    k_morton_instances<<<cdiv(n, T), T, 0, st>>>(t->d_inst, t->d_blas_roots, n, d_bounds, d_codes, t->leaf_map);
    radix_sort_pairs(st, d_codes, t->leaf_map, d_codes2, d_idx2, n, d_hist);  // leaf_map = sorted position -> instance index
    build_tree(st, d_codes, n, nullptr, t->inst_boxes, t->leaf_map, 1, t->topo, t->parent, t->flags, t->boxes, t->nodes2, t->nodes4, t->inst_boxes_tight,
               t->boxes_tight);
    k_read_root<<<1, 32, 0, st>>>(t->boxes, reinterpret_cast<float *>(t->d_small + 10));
    CK(cudaMemcpyAsync(t->root_aabb, t->d_small + 10, 24, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    return extent_supported(t->root_aabb, err);
}

bool rc_refit_tlas(cudaStream_t st, const rc_instance_desc *h_inst, uint32_t n, RcDeviceTlas *t, std::string &err) {
    if (n == 0) return true;
    if (n != t->n) { err = "refit: instance count changed"; return false; }
    if (!upload_instances(st, t, h_inst, n, err)) return false;
    // update_tlas_leaf_aabbs_kernel! (kernels.jl:487-519) + refit_tlas_aabbs_kernel! (:381-428), then re-quantise the wide nodes
    k_instance_boxes<<<cdiv(n, 256), 256, 0, st>>>(t->d_inst, t->d_blas_roots, n, t->inst_boxes, nullptr, t->d_blas_ptrs, t->inst_boxes_tight);
    refit_tree(st, n, t->inst_boxes, t->leaf_map, 1, t->topo, t->parent, t->flags, t->boxes, t->nodes2, t->nodes4, t->inst_boxes_tight, t->boxes_tight);
    k_read_root<<<1, 32, 0, st>>>(t->boxes, reinterpret_cast<float *>(t->d_small + 10));
    CK(cudaMemcpyAsync(t->root_aabb, t->d_small + 10, 24, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    return extent_supported(t->root_aabb, err);
}

// =================================================================================================
// Serialised BLAS: the built structure moved host <-> device as one blob, the role of
// to_gpu(ArrayType, blas::BLAS) (src/kernel-abstractions.jl:31-36: a BLAS built elsewhere is uploaded, not rebuilt).
//   [RcBlobHeader 128 B][nodes2 64·(2n-1)][nodes4 64·(n+1)][tris 48·n][hull 32·RC_HULL_BOXES][normals 36·n, optional]
// every section starts on a 64-byte boundary.  An import restores byte-identical device arrays, so traces of an imported
// geometry are bit-identical to traces of the original.
// =================================================================================================
static_assert(sizeof(RcBox) == 32, "blob layout");
static_assert(sizeof(RcNode2) == 64 && sizeof(RcNode4) == 64 && sizeof(RcTri) == 48, "blob layout");

struct RcBlobHeader {
    char magic[8];  // "RCBLAS\0\1"
    uint32_t abi_version, leaf_max, hull_boxes, n, n_faces_in, has_normals;
    float root_aabb[6];
    uint64_t total_bytes, payload_hash;
    uint64_t off_nodes2, off_nodes4, off_tris, off_hull, off_normals;  // from the blob start
    float sphere[4];  // bounding sphere (centre, radius^2) of the instance-entry cull
};
static_assert(sizeof(RcBlobHeader) == 128, "blob header is 128 bytes");
static const char RC_BLOB_MAGIC[8] = {'R', 'C', 'B', 'L', 'A', 'S', 0, 1};

static inline uint64_t up64(uint64_t x) { return (x + 63u) & ~(uint64_t)63u; }

static void blob_layout(uint32_t n, bool normals, RcBlobHeader *h) {
    uint64_t o = sizeof(RcBlobHeader);
    h->off_nodes2 = o; o = up64(o + sizeof(RcNode2) * (2 * (uint64_t)n - 1));
    h->off_nodes4 = o; o = up64(o + sizeof(RcNode4) * ((uint64_t)n + 1));
    h->off_tris = o;   o = up64(o + sizeof(RcTri) * (uint64_t)n);
    h->off_hull = o;   o = up64(o + sizeof(RcBox) * RC_HULL_BOXES);
    h->off_normals = normals ? o : 0;
    if (normals) o = up64(o + sizeof(float) * 9 * (uint64_t)n);
    h->total_bytes = o;
}

// word-wise multiply-xorshift hash of the payload (everything after the header); detects truncation and bit rot, not an adversary
static uint64_t blob_hash(const uint8_t *p, uint64_t bytes) {
    uint64_t h = 0x9E3779B97F4A7C15ull ^ bytes;
    const uint64_t nw = bytes / 8;
    for (uint64_t i = 0; i < nw; i++) {
        uint64_t w;
        memcpy(&w, p + 8 * i, 8);
        h = (h ^ w) * 0xD6E8FEB86659FD93ull;
        h ^= h >> 32;
    }
    for (uint64_t i = 8 * nw; i < bytes; i++) h = (h ^ p[i]) * 0x100000001B3ull;
    return h;
}

uint64_t rc_blas_blob_bytes(const RcDeviceBlas &b) {
    RcBlobHeader h;
    blob_layout(b.n, b.normals != nullptr, &h);
    return h.total_bytes;
}

bool rc_blas_export(cudaStream_t st, const RcDeviceBlas &b, void *blob, uint64_t capacity, std::string &err) {
    if (b.n == 0 || !b.nodes2 || !b.nodes4 || !b.tris || !b.hull) { err = "export: geometry is not built"; return false; }
    RcBlobHeader h;
    memset(&h, 0, sizeof h);
    memcpy(h.magic, RC_BLOB_MAGIC, 8);
    h.abi_version = RC_ABI_VERSION;
    h.leaf_max = RC_BLAS_LEAF_MAX;
    h.hull_boxes = RC_HULL_BOXES;
    h.n = b.n;
    h.n_faces_in = b.n_faces_in;
    h.has_normals = b.normals ? 1u : 0u;
    memcpy(h.root_aabb, b.root_aabb, 24);
    memcpy(h.sphere, b.sphere, 16);
    blob_layout(b.n, b.normals != nullptr, &h);
    if (capacity < h.total_bytes) { err = "export: capacity too small"; return false; }
    uint8_t *p = static_cast<uint8_t *>(blob);
    memset(p + sizeof h, 0, h.total_bytes - sizeof h);  // alignment gaps are part of the hashed payload
    const uint64_t n = b.n;
    CK(cudaMemcpyAsync(p + h.off_nodes2, b.nodes2, sizeof(RcNode2) * (2 * n - 1), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(p + h.off_nodes4, b.nodes4, sizeof(RcNode4) * (n + 1), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(p + h.off_tris, b.tris, sizeof(RcTri) * n, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(p + h.off_hull, b.hull, sizeof(RcBox) * RC_HULL_BOXES, cudaMemcpyDeviceToHost, st));
    if (b.normals) CK(cudaMemcpyAsync(p + h.off_normals, b.normals, sizeof(float) * 9 * n, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    // wide-node slots no kernel writes (slot 0; slot n when there are internal nodes) are zeroed so equal geometry gives equal blobs
    memset(p + h.off_nodes4, 0, sizeof(RcNode4));
    if (n > 1) memset(p + h.off_nodes4 + sizeof(RcNode4) * n, 0, sizeof(RcNode4));
    h.payload_hash = blob_hash(p + sizeof h, h.total_bytes - sizeof h);
    memcpy(p, &h, sizeof h);
    return true;
}

// structural check of an uploaded blob (rc_validate_blas_elem, rc_build_core.cuh): every reference stays inside the arrays and no
// cycle is reachable from a root, so a damaged blob can neither send a traversal out of bounds nor make it spin
__global__ void k_validate_blas(const RcNode2 *__restrict__ nodes2, const RcNode4 *__restrict__ nodes4, const RcTri *__restrict__ tris, uint32_t n,
                                uint32_t n_faces_in, uint32_t *__restrict__ bad) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t errs = rc_validate_blas_elem(i, nodes2, nodes4, tris, n, RC_BLAS_LEAF_MAX, n_faces_in);
    if (errs) atomicAdd(bad, errs);
}

// host-side checks of a blob (no GPU involved): magic, layout version, section table, size, payload hash, supported extent
static bool blob_check(const void *blob, uint64_t size, RcBlobHeader &h, std::string &err) {
    if (!blob || size < sizeof(RcBlobHeader)) { err = "import: blob too small"; return false; }
    memcpy(&h, blob, sizeof h);
    if (memcmp(h.magic, RC_BLOB_MAGIC, 8) != 0) { err = "import: not a raycore BLAS blob"; return false; }
    if (h.abi_version != RC_ABI_VERSION || h.leaf_max != RC_BLAS_LEAF_MAX || h.hull_boxes != RC_HULL_BOXES) {
        err = "import: blob was written by an incompatible library build";
        return false;
    }
    if (h.n == 0 || h.n > RC_LEAF_START_MASK - 16u || h.n_faces_in < h.n) { err = "import: bad triangle count"; return false; }
    RcBlobHeader want = h;
    blob_layout(h.n, h.has_normals != 0, &want);
    if (want.total_bytes != h.total_bytes || want.off_nodes2 != h.off_nodes2 || want.off_nodes4 != h.off_nodes4 || want.off_tris != h.off_tris ||
        want.off_hull != h.off_hull || want.off_normals != h.off_normals) {
        err = "import: section table does not match the triangle count";
        return false;
    }
    if (size < h.total_bytes) { err = "import: blob is truncated"; return false; }
    const uint8_t *p = static_cast<const uint8_t *>(blob);
    if (blob_hash(p + sizeof h, h.total_bytes - sizeof h) != h.payload_hash) { err = "import: payload hash mismatch (corrupted blob)"; return false; }
    return extent_supported(h.root_aabb, err);
}

bool rc_blas_blob_check(const void *blob, uint64_t size, uint32_t *n_triangles, uint32_t *n_faces_in, uint32_t *has_normals, std::string &err) {
    RcBlobHeader h;
    if (!blob_check(blob, size, h, err)) return false;
    if (n_triangles) *n_triangles = h.n;
    if (n_faces_in) *n_faces_in = h.n_faces_in;
    if (has_normals) *has_normals = h.has_normals;
    return true;
}

bool rc_blas_import(cudaStream_t st, const void *blob, uint64_t size, RcDeviceBlas *out, std::string &err) {
    *out = RcDeviceBlas();
    RcBlobHeader h;
    if (!blob_check(blob, size, h, err)) return false;
    const uint8_t *p = static_cast<const uint8_t *>(blob);
    const uint64_t n = h.n;
    uint32_t *d_bad = nullptr;
    RcTemps tmp(st);
    TMP(d_bad, 1);
    CK(cudaMallocAsync(&out->nodes2, sizeof(RcNode2) * (2 * n - 1), st));
    CK(cudaMallocAsync(&out->nodes4, sizeof(RcNode4) * (n + 1), st));
    CK(cudaMallocAsync(&out->tris, sizeof(RcTri) * n, st));
    CK(cudaMallocAsync(&out->hull, sizeof(RcBox) * RC_HULL_BOXES, st));
    if (h.has_normals) CK(cudaMallocAsync(&out->normals, sizeof(float) * 9 * n, st));
    out->n = h.n;
    out->n_faces_in = h.n_faces_in;
    memcpy(out->root_aabb, h.root_aabb, 24);
    memcpy(out->sphere, h.sphere, 16);
    CK(cudaMemcpyAsync(out->nodes2, p + h.off_nodes2, sizeof(RcNode2) * (2 * n - 1), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(out->nodes4, p + h.off_nodes4, sizeof(RcNode4) * (n + 1), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(out->tris, p + h.off_tris, sizeof(RcTri) * n, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(out->hull, p + h.off_hull, sizeof(RcBox) * RC_HULL_BOXES, cudaMemcpyHostToDevice, st));
    if (h.has_normals) CK(cudaMemcpyAsync(out->normals, p + h.off_normals, sizeof(float) * 9 * n, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(d_bad, 0, 4, st));
    k_validate_blas<<<cdiv((uint32_t)(2 * n), 256), 256, 0, st>>>(out->nodes2, out->nodes4, out->tris, h.n, h.n_faces_in, d_bad);
    uint32_t bad = 0;
    CK(cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));  // the caller may release the blob on return
    CK(cudaGetLastError());
    if (bad) { err = "import: blob fails the structural check (" + std::to_string(bad) + " bad references)"; return false; }
    return true;
}
