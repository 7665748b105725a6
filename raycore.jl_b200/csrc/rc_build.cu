// rc_build.cu — GPU LBVH builder for sm_100a: degenerate filter + stable compaction, scene bounds, 30-bit Morton codes, hand-written stable
// sorts (large inputs: LSD radix, 3 passes of 10 bits, inside one cooperative front-end kernel; small inputs: counting sort of 256-key runs +
// merge by binary searches), Karras radix tree, block-local bottom-up fit, collapse to the quantised BVH4 the fast traversal uses, optional
// reference-layout BVH2 emission.
//
// Replaces build_blas (src/instanced-bvh.jl:1376-1443), build_tlas_topology (:1485-1594),
// refit_tlas! (:2197-2222) and kernels K0-K11 of src/instanced-bvh-kernels.jl.  Everything runs on
// the caller's stream.  A BLAS build is 5 dependent launches — ONE cooperative kernel for <= 32,768 faces (k_build_small), likewise a TLAS
// of <= 32,768 instances (k_tlas_small) — with NO host round trip in the middle: the number of valid triangles stays on the device
// (every kernel reads it from the build's control block), and the only device->host traffic is one 104-byte read-back at the end
// (valid count, root box, bounding sphere, invariant flag).
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
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

// Programmatic dependent launch between the builder's dependent kernels: every block of a kernel signals at once that the next kernel
// of the stream may be scheduled (RC_PDL_TRIGGER), and a kernel waits for the completion (and memory flush) of its predecessor before it
// touches anything (RC_PDL_WAIT, the first statement) — so the launch latency and the block ramp of kernel k+1 hide behind the tail of
// kernel k instead of following it (a few microseconds per boundary; four boundaries per build).  Without the launch attribute both are
// no-ops.  RC_NO_PDL=1 in the environment turns the attribute off (A/B measurements).
#define RC_PDL_TRIGGER() asm volatile("griddepcontrol.launch_dependents;")
#define RC_PDL_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")
// Used from 128 K elements up: below that the builder waits for the host's launches rather than for the GPU, and the attribute makes a launch
// a little more expensive (50 k faces: 0.124 -> 0.127 ms with it, 250 k: 0.161 -> 0.156, 1 M: 0.342 -> 0.337, 4 M: 1.009 -> 0.998).
constexpr uint32_t PDL_MIN_ELEMENTS = 1u << 17;
template <class... KArgs, class... Args>
static void launch_dependent(bool pdl, void (*kernel)(KArgs...), uint32_t grid, uint32_t block, size_t smem, cudaStream_t st, Args... args) {
    static const bool off = getenv("RC_NO_PDL") != nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (off || !pdl) ? 0 : 1;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// Control block of one build (u32 words, zeroed by one memset): the builder's kernels communicate through it.
//   [CTL_N] valid primitives   [CTL_R2] bits of the bounding-sphere radius^2   [CTL_BOUNDS..+6) scene bounds, encoded so that 0 is the
//   identity of atomicMax: word k < 3 holds ~ordered(min_k), word 3+k holds ordered(max_k)   [CTL_TILE] scratch of the refit check
//   [CTL_NSPAN] length of the fit's spanning-node list   [CTL_BAR] arrival counter of the grid barriers (k_tlas_small: + 1 = blocks that have left)   [CTL_ERR] internal-invariant flag
//   [CTL_OUT..+10) floats read back by the host: root box (6), sphere (4)
enum { CTL_N = 0, CTL_R2 = 1, CTL_BOUNDS = 2, CTL_TILE = 8, CTL_NSPAN = 10, CTL_BAR = 11, CTL_ERR = 15, CTL_OUT = 16, CTL_WORDS = 32 };

__device__ __forceinline__ f3 ctl_bounds_min(const uint32_t *ctl) {
    return mk3(rc_ordered_to_float(~ctl[CTL_BOUNDS]), rc_ordered_to_float(~ctl[CTL_BOUNDS + 1]), rc_ordered_to_float(~ctl[CTL_BOUNDS + 2]));
}
__device__ __forceinline__ f3 ctl_bounds_max(const uint32_t *ctl) {
    return mk3(rc_ordered_to_float(ctl[CTL_BOUNDS + 3]), rc_ordered_to_float(ctl[CTL_BOUNDS + 4]), rc_ordered_to_float(ctl[CTL_BOUNDS + 5]));
}
// the element count of a build: on the device (BLAS: written by k_front) or a host constant (TLAS)
__device__ __forceinline__ uint32_t count_of(const uint32_t *n_ptr, uint32_t n_host) { return n_ptr ? *n_ptr : n_host; }

// =================================================================================================
// Scan (exclusive, u32) — block tiles of 2048 + single-block scan of the tile sums (used by the collision broad phase)
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
void rc_exclusive_scan_u32(cudaStream_t st, const uint32_t *in, uint32_t *out, uint32_t n, uint32_t *tile_tmp, uint32_t *d_total) {
    uint32_t tiles = cdiv(n, SCAN_TILE);
    k_tile_sums<<<tiles, SCAN_THREADS, 0, st>>>(in, n, tile_tmp);
    k_scan_single<<<1, 1024, 0, st>>>(tile_tmp, tiles, d_total);
    k_scan_apply<<<tiles, SCAN_THREADS, 0, st>>>(in, n, tile_tmp, out);
}

// =================================================================================================
// Stable LSD radix sort of (u32 key, u32 value) pairs, 10 bits per pass: a 30-bit Morton code is three passes.
//   per pass:  k_radix_hist    per-tile digit histogram (pass 0: written by the kernel that produces the keys) -> hist[digit * tiles + tile]
//              k_scan_rows     row-wise exclusive scan of the digit-major table + row totals
//              k_radix_scatter stable in-tile ranking (warp match_any) + scatter; the digit bases come from the row totals
// Stability is what makes the topology equal the reference's (AK.sortperm is a stable merge sort, src/instanced-bvh.jl:1399).
// =================================================================================================
constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_BITS = 10;
constexpr int RS_DIGITS = 1 << RS_BITS;
constexpr int RS_PASSES = 3;
constexpr int RS_DPT = RS_DIGITS / RS_THREADS;  // digits per thread in the table steps

static size_t radix_hist_words(uint32_t n_bound) { return (size_t)RS_DIGITS * cdiv(n_bound, RS_TILE) + RS_DIGITS; }

// =================================================================================================
// BLAS front end: filter + stable compaction + bounds in one pass, then Morton codes (+ the first radix histogram)
// =================================================================================================
__device__ __forceinline__ f3 ld3(const float *p) { return mk3(p[0], p[1], p[2]); }

// block-wide min / max of a box into the control block (REDUX per warp, six atomics per block).  Every thread of the block calls it.
__device__ __forceinline__ void bounds_atomic(uint32_t *ctl, f3 lo, f3 hi) {
    __shared__ uint32_t part[6][32];
    // 0 is the identity of both encodings: ~ordered(+Inf) and ordered(-Inf) are > 0 only for real boxes' complements... (empty lanes pass +Inf / -Inf)
    uint32_t v[6] = {~rc_float_to_ordered(lo.x), ~rc_float_to_ordered(lo.y), ~rc_float_to_ordered(lo.z),
                     rc_float_to_ordered(hi.x), rc_float_to_ordered(hi.y), rc_float_to_ordered(hi.z)};
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int c = 0; c < 6; c++) {
        uint32_t r = __reduce_max_sync(0xFFFFFFFFu, v[c]);
        if (lane == 0) part[c][wid] = r;
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int c = 0; c < 6; c++) {
            uint32_t x = lane < nw ? part[c][lane] : 0u;
            uint32_t r = __reduce_max_sync(0xFFFFFFFFu, x);
            if (lane == 0) atomicMax(&ctl[CTL_BOUNDS + c], r);
        }
    }
}

// =================================================================================================
// The front end as ONE persistent cooperative kernel: filter + stable compaction + scene bounds, Morton codes, and the three radix passes,
// separated by grid barriers instead of kernel boundaries (a launch boundary costs 3-5 us of drain + ramp; the 1 M-triangle front end was
// 10 dependent launches).  Every block is resident (cooperative launch, grid <= occupancy x SMs) and loops over its tiles:
//   F   per 2048-face tile: exact degenerate test, valid count -> tile_counts; scene bounds (one set of atomics per block)     | barrier
//   M   prefix of the tile over tile_counts (no look-back chain), faces re-read (L2), compacted RcTri records, Morton codes    | barrier
//   per pass: H per-tile digit histograms | barrier | R row scans of the digit-major table + digit totals | barrier | S stable scatter | barrier
// Data that other blocks wrote earlier in the same launch is read with ld.cg (L1 is not coherent across a grid barrier).
// The TLAS build runs the same kernel without F / M (its keys come from k_morton_instances).
// =================================================================================================
constexpr int FT_TILE = 2048;  // faces per filter tile (8 per thread): 1 M faces are 490 tiles, one per resident block
__device__ __forceinline__ uint32_t ld_acquire(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
#ifdef RC_FRONT_PROF  // diagnosis build (tools/build_variant.sh): block 0 prints the time of every phase boundary
__device__ __forceinline__ unsigned long long prof_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define RC_PROF_MARK(name) if (blockIdx.x == 0 && threadIdx.x == 0 && prof_k < 48) { prof_t[prof_k] = prof_now(); prof_name[prof_k++] = name; }
#define RC_PROF_DUMP() if (blockIdx.x == 0 && threadIdx.x == 0) { for (int q = 0; q < prof_k; q++) printf("k_front %-10s %8llu ns\n", prof_name[q], prof_t[q] - (q ? prof_t[q - 1] : prof_t0)); }
// finer marks inside the phases of k_build_small (device-global so that the shared bodies can set them)
__device__ unsigned long long g_sb_t[96];
__device__ const char *g_sb_n[96];
__device__ int g_sb_k;
#define SB_MARK(name) if (blockIdx.x == 0 && threadIdx.x == 0 && g_sb_k < 96) { g_sb_t[g_sb_k] = prof_now(); g_sb_n[g_sb_k++] = name; }
#define SB_DUMP(what) if (blockIdx.x == 0 && threadIdx.x == 0) { for (int q = 1; q < g_sb_k; q++) printf("  %s %-8s %6llu ns\n", what, g_sb_n[q], g_sb_t[q] - g_sb_t[q - 1]); g_sb_k = 0; }
#else
#define RC_PROF_MARK(name)
#define RC_PROF_DUMP()
#define SB_MARK(name)
#define SB_DUMP(what)
#endif
__device__ __forceinline__ void grid_barrier(uint32_t *bar, uint32_t &target) {
    if (gridDim.x == 1) {  // a one-block grid (tiny meshes, the one-instance TLAS of every single-mesh scene): the block barrier is the grid barrier
        __syncthreads();
        return;
    }
    target += gridDim.x;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        while (ld_acquire(bar) < target) __nanosleep(40);  // back off: hundreds of pollers on one L2 line delay the arrivals they wait for
        __threadfence();
    }
    __syncthreads();
}
struct FrontArgs {
    const float *verts;         // null: sort only (n = n_host, keys / vals already in keys0 / vals0)
    const uint32_t *face_meta;  // nullable
    uint32_t n_faces;
    RcTri *tris_in;
    uint32_t *tile_counts;      // cdiv(n_faces, FT_TILE)
    uint32_t *ctl;              // bounds, CTL_N
    uint32_t *keys0, *vals0, *keys1, *vals1;  // three passes: the sorted pairs end up in (keys1, vals1)
    uint32_t n_host;
    uint32_t *hist;             // RS_DIGITS x tiles_stride, then RS_DIGITS totals
    uint32_t tiles_stride;
    uint32_t *bar;              // zeroed
};
__device__ __forceinline__ f3 ldcg3(const uint32_t *ctl, int k, bool inv) {
    const uint32_t a = __ldcg(ctl + k), b = __ldcg(ctl + k + 1), c = __ldcg(ctl + k + 2);
    return inv ? mk3(rc_ordered_to_float(~a), rc_ordered_to_float(~b), rc_ordered_to_float(~c)) : mk3(rc_ordered_to_float(a), rc_ordered_to_float(b), rc_ordered_to_float(c));
}
__global__ void __launch_bounds__(RS_THREADS, 4) k_front(const FrontArgs A) {
    __shared__ uint32_t wh[RS_WARPS][RS_DIGITS];
    __shared__ uint32_t sm[40];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5, lt_mask = (1u << lane) - 1u;
    uint32_t target = 0, n = A.n_host;
#ifdef RC_FRONT_PROF
    unsigned long long prof_t0 = prof_now(), prof_t[48];
    const char *prof_name[48];
    int prof_k = 0;
#endif
    if (A.verts) {
        const uint32_t f_tiles = (A.n_faces + FT_TILE - 1) / FT_TILE;
        // ---- F: valid faces per tile, scene bounds
        f3 lo = mk3(INFINITY, INFINITY, INFINITY), hi = mk3(-INFINITY, -INFINITY, -INFINITY);
        for (uint32_t t = blockIdx.x; t < f_tiles; t += gridDim.x) {
            uint32_t cnt = 0;
#pragma unroll
            for (int it = 0; it < FT_TILE / RS_THREADS; it++) {
                const uint32_t i = t * FT_TILE + it * RS_THREADS + tid;
                if (i < A.n_faces) {
                    const float *v = A.verts + (size_t)i * 9;
                    const f3 a = ld3(v), b = ld3(v + 3), c = ld3(v + 6);
                    if (!x_is_degenerate(a, b, c)) {
                        cnt++;
                        lo = jl_min3(lo, jl_min3(jl_min3(a, b), c));  // world_bound(tri), triangle_mesh.jl:37
                        hi = jl_max3(hi, jl_max3(jl_max3(a, b), c));
                    }
                }
            }
            cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
            if (lane == 0) sm[wid] = cnt;
            __syncthreads();
            if (tid == 0) {
                uint32_t s = 0;
#pragma unroll
                for (int w = 0; w < RS_WARPS; w++) s += sm[w];
                A.tile_counts[t] = s;
            }
            __syncthreads();
        }
        bounds_atomic(A.ctl, lo, hi);
        RC_PROF_MARK("F")
        grid_barrier(A.bar, target);
        RC_PROF_MARK("F-barrier")
        // ---- M: compacted records + Morton codes (calculate_morton_code_for_prim, kernels.jl:88-98; extent unguarded, :1388)
        const f3 smin = ldcg3(A.ctl, CTL_BOUNDS, true), smax = ldcg3(A.ctl, CTL_BOUNDS + 3, false);
        const f3 ext = x_sub3(smax, smin);
        uint32_t total = 0;
        {   // every block needs the valid count
            uint32_t s = 0;
            for (uint32_t k = tid; k < f_tiles; k += RS_THREADS) s += __ldcg(A.tile_counts + k);
            uint32_t tot_;
            block_excl_scan(s, sm, tot_);
            total = tot_;
        }
        n = total;
        if (blockIdx.x == 0 && tid == 0) A.ctl[CTL_N] = n;
        for (uint32_t t = blockIdx.x; t < f_tiles; t += gridDim.x) {
            uint32_t s = 0;
            for (uint32_t k = tid; k < t; k += RS_THREADS) s += __ldcg(A.tile_counts + k);
            uint32_t prefix;
            block_excl_scan(s, sm, prefix);  // (its total = the faces kept before this tile)
            // warp w owns the contiguous faces [w * 256, (w + 1) * 256) of the tile; item `it` of lane l = chunk + it * 32 + l, so (it, l)
            // lexicographic order == face order and the ranks are stable (filter! keeps the order, src/instanced-bvh.jl:591-600)
            constexpr int ITEMS = FT_TILE / RS_THREADS;
            uint32_t bal[ITEMS], wsum = 0;
#pragma unroll
            for (int it = 0; it < ITEMS; it++) {
                const uint32_t i = t * FT_TILE + wid * (32 * ITEMS) + it * 32 + lane;
                bool valid = false;
                if (i < A.n_faces) {
                    const float *v = A.verts + (size_t)i * 9;
                    valid = !x_is_degenerate(ld3(v), ld3(v + 3), ld3(v + 6));
                }
                bal[it] = __ballot_sync(0xFFFFFFFFu, valid);
                wsum += __popc(bal[it]);
            }
            uint32_t tile_total;
            const uint32_t woff = block_excl_scan(lane == 0 ? wsum : 0u, sm, tile_total);  // exclusive over the threads: lane 0 of warp w holds the warps before it
            uint32_t base = prefix + __shfl_sync(0xFFFFFFFFu, woff, 0);
#pragma unroll
            for (int it = 0; it < ITEMS; it++) {
                if ((bal[it] >> lane) & 1u) {
                    const uint32_t i = t * FT_TILE + wid * (32 * ITEMS) + it * 32 + lane;
                    const uint32_t k = base + __popc(bal[it] & lt_mask);
                    const float *v = A.verts + (size_t)i * 9;  // (second read of the face: L1)
                    const f3 a = ld3(v), b = ld3(v + 3), c = ld3(v + 6);
                    float4 *d = reinterpret_cast<float4 *>(A.tris_in + k);
                    d[0] = make_float4(a.x, a.y, a.z, __uint_as_float(k));
                    d[1] = make_float4(b.x, b.y, b.z, __uint_as_float(A.face_meta ? A.face_meta[i] : i + 1u));  // :595
                    d[2] = make_float4(c.x, c.y, c.z, __uint_as_float(i));
                    const f3 blo = jl_min3(jl_min3(a, b), c), bhi = jl_max3(jl_max3(a, b), c);
                    const f3 ctr = mk3(x_mul(0.5f, x_add(blo.x, bhi.x)), x_mul(0.5f, x_add(blo.y, bhi.y)), x_mul(0.5f, x_add(blo.z, bhi.z)));
                    const f3 nrm = mk3(x_div(x_sub(ctr.x, smin.x), ext.x), x_div(x_sub(ctr.y, smin.y), ext.y), x_div(x_sub(ctr.z, smin.z), ext.z));
                    A.keys0[k] = rc_morton30(nrm);
                    A.vals0[k] = k;
                }
                base += __popc(bal[it]);
            }
        }
        RC_PROF_MARK("M")
        grid_barrier(A.bar, target);
        RC_PROF_MARK("M-barrier")
    }
    if (n == 0) return;  // (uniform over the grid)
    const uint32_t tiles = (n + RS_TILE - 1) / RS_TILE, stride = A.tiles_stride;
    uint32_t *totals = A.hist + (size_t)RS_DIGITS * stride;
    uint32_t *ki = A.keys0, *vi = A.vals0, *ko = A.keys1, *vo = A.vals1;
    for (int pass = 0; pass < RS_PASSES; pass++) {
        const int shift = pass * RS_BITS;
        // ---- H: per-tile digit histograms -> hist[digit * stride + tile]
        uint32_t *sh = &wh[0][0];
        for (uint32_t t = blockIdx.x; t < tiles; t += gridDim.x) {
#pragma unroll
            for (int k = 0; k < RS_DPT; k++) sh[tid + k * RS_THREADS] = 0;
            __syncthreads();
#pragma unroll
            for (int i = 0; i < RS_ITEMS; i++) {
                const uint32_t idx = t * RS_TILE + i * RS_THREADS + tid;
                if (idx < n) atomicAdd(&sh[(__ldcg(ki + idx) >> shift) & (RS_DIGITS - 1u)], 1u);
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < RS_DPT; k++) {
                const uint32_t d = tid + k * RS_THREADS;
                A.hist[(size_t)d * stride + t] = sh[d];
            }
            __syncthreads();
        }
        RC_PROF_MARK("H")
        grid_barrier(A.bar, target);
        RC_PROF_MARK("H-barrier")
        // ---- R: row-wise exclusive scan of the digit-major table, one warp per digit row; row totals -> totals[]
        // (rows are dealt to the blocks round-robin — warp 0 of every block first —, a row is read with coalesced, independent loads: 16 in
        // flight per lane, so a row of <= 512 tiles costs one L2 round trip)
        for (uint32_t d = blockIdx.x + wid * gridDim.x; d < (uint32_t)RS_DIGITS; d += gridDim.x * RS_WARPS) {
            uint32_t *row = A.hist + (size_t)d * stride;
            uint32_t run = 0;
            for (uint32_t base = 0; base < tiles; base += 32u * 16u) {
                uint32_t v[16];
#pragma unroll
                for (int u = 0; u < 16; u++) {
                    const uint32_t i = base + u * 32u + lane;
                    v[u] = i < tiles ? __ldcg(row + i) : 0u;
                }
#pragma unroll
                for (int u = 0; u < 16; u++) {
                    const uint32_t i = base + u * 32u + lane;
                    const uint32_t inc = warp_incl_scan(v[u]);
                    if (i < tiles) row[i] = run + inc - v[u];
                    run += __shfl_sync(0xFFFFFFFFu, inc, 31);
                }
            }
            if (lane == 0) totals[d] = run;
        }
        RC_PROF_MARK("R")
        grid_barrier(A.bar, target);
        RC_PROF_MARK("R-barrier")
        // ---- S: stable in-tile ranking (warp match_any) + scatter; the digit bases come from the row totals
        for (uint32_t t = blockIdx.x; t < tiles; t += gridDim.x) {
            for (int i = tid; i < RS_WARPS * RS_DIGITS; i += RS_THREADS) (&wh[0][0])[i] = 0;
            __syncthreads();
            // warp w owns the contiguous chunk [w*256, (w+1)*256) of the tile; item i of lane l = chunk + i*32 + l,
            // so (i, l) lexicographic order == memory order and ranks are stable.
            const uint32_t base = t * RS_TILE + wid * (32 * RS_ITEMS);
            uint32_t key[RS_ITEMS], val[RS_ITEMS], rank[RS_ITEMS], dig[RS_ITEMS];
#pragma unroll
            for (int i = 0; i < RS_ITEMS; i++) {
                const uint32_t idx = base + i * 32 + lane;
                const bool ok = idx < n;
                key[i] = ok ? __ldcg(ki + idx) : 0xFFFFFFFFu;
                val[i] = ok ? __ldcg(vi + idx) : 0u;
            }
#pragma unroll
            for (int i = 0; i < RS_ITEMS; i++) {
                const bool ok = base + i * 32 + lane < n;
                dig[i] = ok ? ((key[i] >> shift) & (RS_DIGITS - 1u)) : (uint32_t)RS_DIGITS;  // RS_DIGITS = padding lane group
                const uint32_t peers = __match_any_sync(0xFFFFFFFFu, dig[i]);
                const uint32_t leader = __ffs(peers) - 1;
                uint32_t prev = 0;
                if (ok && lane == leader) {
                    prev = wh[wid][dig[i]];
                    wh[wid][dig[i]] = prev + __popc(peers);
                }
                prev = __shfl_sync(0xFFFFFFFFu, prev, leader);
                rank[i] = prev + __popc(peers & lt_mask);
                __syncwarp();
            }
            // keys with a smaller digit: exclusive scan of the 1024 row totals, RS_DPT consecutive digits per thread
            uint32_t tot[RS_DPT], local = 0;
#pragma unroll
            for (int k = 0; k < RS_DPT; k++) { tot[k] = __ldcg(totals + tid * RS_DPT + k); local += tot[k]; }
            uint32_t total_;
            uint32_t digit_base = block_excl_scan(local, sm, total_);  // (includes a __syncthreads: the per-warp counts are complete)
#pragma unroll
            for (int k = 0; k < RS_DPT; k++) {  // digit d: turn the per-warp counts into exclusive prefixes starting at the global offset
                const uint32_t d = tid * RS_DPT + k;
                uint32_t off = digit_base + __ldcg(A.hist + (size_t)d * stride + t);
                digit_base += tot[k];
#pragma unroll
                for (int w = 0; w < RS_WARPS; w++) {
                    const uint32_t c = wh[w][d];
                    wh[w][d] = off;
                    off += c;
                }
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < RS_ITEMS; i++) {
                if (dig[i] < (uint32_t)RS_DIGITS) {
                    const uint32_t pos = wh[wid][dig[i]] + rank[i];
                    ko[pos] = key[i];
                    vo[pos] = val[i];
                }
            }
            __syncthreads();
        }
        RC_PROF_MARK("S")
        if (pass + 1 < RS_PASSES) grid_barrier(A.bar, target);
        RC_PROF_MARK("S-barrier")
        if (pass + 1 == RS_PASSES) { RC_PROF_DUMP() }
        uint32_t *tk = ki; ki = ko; ko = tk;
        uint32_t *tv = vi; vi = vo; vo = tv;
    }
}

// grid of the front-end kernel: one block per tile, capped at what is co-resident on this device
static bool launch_front(cudaStream_t st, FrontArgs &A, uint32_t tiles_wanted, std::string &err) {
    static std::atomic<int> max_blocks[64];  // per device; several host threads build concurrently under rc_multi_* (zero-initialised)
    int dev = 0;
    CK(cudaGetDevice(&dev));
    int cap = (dev >= 0 && dev < 64) ? max_blocks[dev].load(std::memory_order_relaxed) : 0;
    if (cap == 0) {
        int per_sm = 0, sms = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_front, RS_THREADS, 0));
        CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        cap = per_sm * sms;
        if (cap < 1) { err = "k_front does not fit on this device"; return false; }
        if (dev >= 0 && dev < 64) max_blocks[dev].store(cap, std::memory_order_relaxed);
    }
    // at least 128 blocks: the row scans of a pass are 1024 independent rows, one warp each
    const uint32_t grid = std::min(std::max(tiles_wanted, 128u), (uint32_t)cap);
    void *args[] = {(void *)&A};
    CK(cudaLaunchCooperativeKernel((const void *)k_front, dim3(grid), dim3(RS_THREADS), args, 0, st));
    return true;
}

// =================================================================================================
// Topology, fit, BVH2 emission, collapse (shared by BLAS and TLAS)
// =================================================================================================
__global__ void k_topology(const uint32_t *__restrict__ codes, const uint32_t *__restrict__ n_ptr, uint32_t n_host, RcTopo *__restrict__ topo, uint32_t *__restrict__ parent) {
    RC_PDL_TRIGGER();
    RC_PDL_WAIT();
    const uint32_t n = count_of(n_ptr, n_host);
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;  // internal node i+1
    if (n == 1 && i == 0) parent[0] = RC_INVALID;  // single leaf: no internal node, the leaf's parent is INVALID
    if (i + 1 >= n) return;
    RcTopo t = rc_topology_for_node((int)(i + 1), codes, (int)n);
    topo[i] = t;
    parent[t.child0 - 1] = i + 1;  // set_parents_for_node, kernels.jl:159-180
    parent[t.child1 - 1] = i + 1;
    if (i == 0) parent[0] = RC_INVALID;
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
__device__ __forceinline__ void st_node4(RcNode4 *p, const RcNode4 &nd) {
    const float4 *s = reinterpret_cast<const float4 *>(&nd);
    float4 *d = reinterpret_cast<float4 *>(p);
    d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; d[3] = s[3];
}
__device__ __forceinline__ void st_node4_zero(RcNode4 *p) {
    float4 *d = reinterpret_cast<float4 *>(p);
    d[0] = d[1] = d[2] = d[3] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// Bottom-up fit (refit_aabbs_kernel!, kernels.jl:239-286 / :381-428) without a dependency chain through global memory.
//
// A block of k_fit_local owns FIT_T consecutive sorted leaves and the internal nodes of the same numbers.  Every internal node whose
// span lies inside that range (all but a few per block) is fitted in shared memory — arrival counters, boxes, topology and parent
// links of the range are staged there — and collapsed into its wide node from the same shared-memory copies, so the boxes and the
// topology are not read back from global memory at all.  The climb of a thread ends at the first ancestor that reaches beyond the
// block; the subtree it finished is one of the block's maximal in-block subtrees ("segments").  The segments hang off the two root
// paths that end at the block's boundaries, so there are at most 2 x 64 of them (the common-prefix length, 0..63, grows along a path
// of the radix tree).  The block stores, for every segment, the union of the leaf boxes from its first leaf to the end of the block
// (sfx) and from the start of the block to its last leaf (pfx).
//
// A node that spans several blocks ("spanning node") starts at a segment start and ends at a segment end, hence
//     box(node) = sfx[first leaf] u (whole blocks in between) u pfx[last leaf]
// and k_fit_span evaluates that with one warp per spanning node and no ordering between nodes: min / max are exact and the
// reference's min / max (NaN-propagating, -0 < +0) are associative and commutative on the GPU (canonical NaN), so the union over
// the leaves gives the bits the pairwise climb gives.  (The round-1 kernel climbed the spanning levels with a release-ordered
// atomic + ld.cg per level: ~25 dependent L2 round trips on the critical path of every block, 184 us for 1 M triangles.)
//   BLAS build (tris_in != null): sorted triangle p = tris_in[perm[p]] is gathered here and written to tris; leaf box = its bounds;
//                                 leaf node = (v0,v1,v2,0 | INVALID, p+1, parent); the bounding-sphere radius is reduced on the way
//   BLAS refit (tris_in == null, tris != null): the triangles are already in place
//   TLAS (tris == null): leaf box = inst_boxes[leaf_map[p]], leaf node = (lo,hi,0,0 | INVALID, inst, parent)
#ifndef RC_FIT_T
#define RC_FIT_T 512
#endif
constexpr int FIT_T = RC_FIT_T;  // 48 registers x 512 threads: two blocks per SM, so one block's barriers and gather latency overlap the other's work
constexpr int FIT_SEG_MAX = 128;
struct FitSeg {  // 64 B; the table of a block is sorted by position, entry 0 starts at the block's first leaf (its sfx = the whole block)
    uint32_t s, e, count, overflow;
    float sfx[6], pfx[6];
};
static_assert(sizeof(FitSeg) == 64, "segment table entry");
struct FitWork {  // per-fit scratch: segment tables [blocks][FIT_SEG_MAX], spanning-node list [blocks * FIT_SEG_MAX] of (node, span_lo, span_hi, -), its length
    FitSeg *seg;
    unsigned char *pos_idx;  // per sorted position: rank (in its block's table) of the segment that starts or ends there, 0xFF = none
    uint4 *span_list;
    uint32_t *span_count;
    uint32_t *err_flag;  // set when a block found more than FIT_SEG_MAX segments (impossible for a radix tree; checked by the host, never silent)
};
static size_t fit_work_bytes(uint32_t n_bound, uint32_t fit_t = FIT_T) {
    const size_t blocks = cdiv(n_bound, fit_t);
    return blocks * FIT_SEG_MAX * (sizeof(FitSeg) + sizeof(uint4)) + blocks * fit_t + 64;
}
static FitWork fit_work_at(void *base, uint32_t n_bound, uint32_t fit_t = FIT_T) {
    const size_t blocks = cdiv(n_bound, fit_t);
    FitWork w;
    w.seg = reinterpret_cast<FitSeg *>(base);
    w.span_list = reinterpret_cast<uint4 *>(w.seg + blocks * FIT_SEG_MAX);
    w.pos_idx = reinterpret_cast<unsigned char *>(w.span_list + blocks * FIT_SEG_MAX);
    w.span_count = reinterpret_cast<uint32_t *>(w.pos_idx + blocks * fit_t);
    return w;
}
template <int T>  // T = leaves per fit block = threads per block (FIT_T for the multi-launch path, SB_T inside the small-build kernel)
struct FitSmemT {
    RcTopo topo[T];
    uint32_t par_int[T], par_leaf[T], flag[T];
    float box_int[T][6], box_leaf[T][6];
    uint32_t seg_s[FIT_SEG_MAX], seg_e[FIT_SEG_MAX];
    float seg_box[FIT_SEG_MAX][6];
    uint32_t nseg;
    uint32_t pos_idx[T / 4];  // bytes, see FitWork::pos_idx
};
using FitSmem = FitSmemT<FIT_T>;
__device__ __forceinline__ RcBox box_from6(const float *b) {
    RcBox r;
    r.lo[0] = b[0]; r.lo[1] = b[1]; r.lo[2] = b[2]; r.pad0 = 0.f;
    r.hi[0] = b[3]; r.hi[1] = b[4]; r.hi[2] = b[5]; r.pad1 = 0.f;
    return r;
}
// build_list: append this block's spanning nodes to work.span_list (skipped when a list of the same topology is already there)
// nodes4 != null: collapse the in-block nodes into their wide nodes here (leaf_max, leaf_map as in rc_collapse_node)
template <int T>
__device__ __forceinline__ void fit_local_body(unsigned char *fit_raw, const uint32_t block, const RcTri *__restrict__ tris_in, const uint32_t *__restrict__ perm, RcTri *__restrict__ tris,
                                                     const RcBox *__restrict__ inst_boxes, const uint32_t *__restrict__ leaf_map, const uint32_t *__restrict__ n_ptr,
                                                     uint32_t n_host, const RcTopo *__restrict__ topo, const uint32_t *__restrict__ parent, RcBox *__restrict__ boxes,
                                                     RcNode2 *__restrict__ nodes2, uint32_t *__restrict__ ctl, FitWork work, bool build_list, RcNode4 *__restrict__ nodes4,
                                                     uint32_t leaf_max) {
    FitSmemT<T> &S = *reinterpret_cast<FitSmemT<T> *>(fit_raw);
    const uint32_t n = count_of(n_ptr, n_host);
    const uint32_t tid = threadIdx.x;
    const uint32_t blk_lo = block * T + 1u;  // first sorted primitive (1-based) = first internal node number of the range
    if (blk_lo > n) return;
    const uint32_t blk_hi = min(blk_lo + T - 1u, n);
    const uint32_t p1 = blk_lo + tid;  // this thread's primitive (1-based) and the internal node number it stages
    const bool has_leaf = p1 <= blk_hi, has_node = p1 <= blk_hi && p1 < n;
    if (has_leaf) {
        if (has_node) {
            S.topo[tid] = topo[p1 - 1];
            S.par_int[tid] = parent[p1 - 1];
        }
        S.par_leaf[tid] = parent[n - 1 + p1 - 1];
        S.flag[tid] = 0;
    }
    SB_MARK("enter")
    if (tid == 0) S.nseg = 0;
    if (tid < T / 4) S.pos_idx[tid] = 0xFFFFFFFFu;
    __syncthreads();
    SB_MARK("stage")
    float r2 = 0.0f;
    uint32_t node = RC_INVALID, cs = p1, ce = p1;  // parent of / span of the subtree this thread has finished
    f3 lo = mk3(0, 0, 0), hi = lo;
    if (has_leaf) {
        const uint32_t p = p1 - 1, leaf = n - 1 + p1;
        node = S.par_leaf[tid];
        if (tris) {
            float4 a, b, c;
            if (tris_in) {
                const float4 *s = reinterpret_cast<const float4 *>(tris_in + perm[p]);
                a = s[0]; b = s[1]; c = s[2];
                float4 *d = reinterpret_cast<float4 *>(tris + p);
                d[0] = a; d[1] = b; d[2] = c;
            } else {
                const float4 *s = reinterpret_cast<const float4 *>(tris + p);
                a = s[0]; b = s[1]; c = s[2];
            }
            const f3 v0 = mk3(a.x, a.y, a.z), v1 = mk3(b.x, b.y, b.z), v2 = mk3(c.x, c.y, c.z);
            lo = jl_min3(jl_min3(v0, v1), v2);  // get_node_aabb leaf branch, :1148-1158
            hi = jl_max3(jl_max3(v0, v1), v2);
            if (nodes2) st_node2(nodes2 + (leaf - 1), v0, v1, v2, mk3(0, 0, 0), RC_INVALID, p1, node);
            const f3 smin = ctl_bounds_min(ctl), smax = ctl_bounds_max(ctl);
            r2 = rc_far2(mk3(0.5f * (smin.x + smax.x), 0.5f * (smin.y + smax.y), 0.5f * (smin.z + smax.z)), v0, v1, v2);
            if (!(r2 == r2)) r2 = INFINITY;  // NaN vertices: infinite radius (no cull)
        } else {
            const uint32_t inst = leaf_map[p];
            const RcBox b = inst_boxes[inst];
            lo = mk3(b.lo[0], b.lo[1], b.lo[2]);
            hi = mk3(b.hi[0], b.hi[1], b.hi[2]);
            if (nodes2) st_node2(nodes2 + (leaf - 1), lo, hi, mk3(0, 0, 0), mk3(0, 0, 0), RC_INVALID, inst, node);
        }
        st_box(boxes + (leaf - 1), lo, hi);
        S.box_leaf[tid][0] = lo.x; S.box_leaf[tid][1] = lo.y; S.box_leaf[tid][2] = lo.z;
        S.box_leaf[tid][3] = hi.x; S.box_leaf[tid][4] = hi.y; S.box_leaf[tid][5] = hi.z;
    }
    __syncthreads();
    SB_MARK("leaves")
    // The climb, one level per round with a block barrier between rounds: a thread that holds a finished subtree counts its arrival at the
    // parent; the second arriver — both children's boxes were stored in earlier rounds — fits the parent and carries on.  (Barrier-
    // ordered: no fences, clean under racecheck; a 1024-leaf range of a Morton-ordered tree is 12-20 levels deep.  Letting a thread climb
    // several levels per round when the sibling arrived in an earlier round was measured slower: the rounds are issue-bound, not
    // barrier-bound.)
    bool active = has_leaf;
    for (;;) {
        if (active) {
            bool local = false;
            uint32_t s = 0;
            RcTopo tp = {0, 0, 0, 0};
            if (node != RC_INVALID && node >= blk_lo && node <= blk_hi) {
                s = node - blk_lo;
                tp = S.topo[s];
                local = tp.span_lo >= blk_lo && tp.span_hi <= blk_hi;
            }
            if (!local) {  // the parent reaches beyond the block (or there is none): [cs, ce] is a segment
                const uint32_t k = atomicAdd(&S.nseg, 1u);
                if (k < (uint32_t)FIT_SEG_MAX) {
                    S.seg_s[k] = cs; S.seg_e[k] = ce;
                    S.seg_box[k][0] = lo.x; S.seg_box[k][1] = lo.y; S.seg_box[k][2] = lo.z;
                    S.seg_box[k][3] = hi.x; S.seg_box[k][4] = hi.y; S.seg_box[k][5] = hi.z;
                }
                active = false;
            } else if (atomicAdd(&S.flag[s], 1u) == 0u) {
                active = false;  // first arriver: the sibling subtree is not ready
            } else {
                // second arriver: its own box is in registers, the sibling's was stored before the last barrier
                const bool me_left = cs == tp.span_lo;  // the left child covers [span_lo, split]
                const uint32_t sib = me_left ? tp.child1 : tp.child0, up = S.par_int[s];
                const float *bs = sib >= n ? S.box_leaf[sib - (n - 1) - blk_lo] : S.box_int[sib - blk_lo];
                const f3 ls = mk3(bs[0], bs[1], bs[2]), hs = mk3(bs[3], bs[4], bs[5]);
                const f3 l0 = me_left ? lo : ls, h0 = me_left ? hi : hs, l1 = me_left ? ls : lo, h1 = me_left ? hs : hi;
                lo = jl_min3(l0, l1);  // get_node_aabb interior branch, :1142-1147
                hi = jl_max3(h0, h1);
                if (nodes2) st_node2(nodes2 + (node - 1), l0, h0, l1, h1, tp.child0, tp.child1, up);
                st_box(boxes + (node - 1), lo, hi);
                float *bi = S.box_int[s];
                bi[0] = lo.x; bi[1] = lo.y; bi[2] = lo.z; bi[3] = hi.x; bi[4] = hi.y; bi[5] = hi.z;
                cs = tp.span_lo; ce = tp.span_hi;
                node = up;
            }
        }
        if (!__syncthreads_or(active)) break;
    }
    SB_MARK("climb")
    if (tris) {  // bits of a non-negative float order like the float: one atomicMax per warp (lanes without a leaf carry 0)
        const uint32_t m = __reduce_max_sync(0xFFFFFFFFu, __float_as_uint(r2));
        if ((tid & 31u) == 0u) atomicMax(&ctl[CTL_R2], m);
    }
    __syncthreads();
    // ---- segment table: position-sorted entries with the suffix / prefix unions of the block's segments
    const uint32_t m_all = S.nseg, m = min(m_all, (uint32_t)FIT_SEG_MAX);
    if (tid == 0 && m_all > (uint32_t)FIT_SEG_MAX) atomicOr(work.err_flag, 1u);
    if (tid < m) {
        const uint32_t sj = S.seg_s[tid], ej = S.seg_e[tid];
        uint32_t rank = 0;
        f3 slo = mk3(INFINITY, INFINITY, INFINITY), shi = mk3(-INFINITY, -INFINITY, -INFINITY), plo = slo, phi = shi;
        for (uint32_t k = 0; k < m; k++) {
            const uint32_t sk = S.seg_s[k];
            const f3 kl = mk3(S.seg_box[k][0], S.seg_box[k][1], S.seg_box[k][2]), kh = mk3(S.seg_box[k][3], S.seg_box[k][4], S.seg_box[k][5]);
            rank += sk < sj ? 1u : 0u;
            if (sk >= sj) { slo = jl_min3(slo, kl); shi = jl_max3(shi, kh); }
            if (sk <= sj) { plo = jl_min3(plo, kl); phi = jl_max3(phi, kh); }
        }
        FitSeg e;
        e.s = sj; e.e = ej; e.count = m; e.overflow = m_all > (uint32_t)FIT_SEG_MAX ? 1u : 0u;
        e.sfx[0] = slo.x; e.sfx[1] = slo.y; e.sfx[2] = slo.z; e.sfx[3] = shi.x; e.sfx[4] = shi.y; e.sfx[5] = shi.z;
        e.pfx[0] = plo.x; e.pfx[1] = plo.y; e.pfx[2] = plo.z; e.pfx[3] = phi.x; e.pfx[4] = phi.y; e.pfx[5] = phi.z;
        const float4 *src = reinterpret_cast<const float4 *>(&e);
        float4 *dst = reinterpret_cast<float4 *>(work.seg + (size_t)block * FIT_SEG_MAX + rank);
        dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
        // (segments are disjoint: a position is the start or the end of at most one of them)
        reinterpret_cast<unsigned char *>(S.pos_idx)[sj - blk_lo] = (unsigned char)rank;
        reinterpret_cast<unsigned char *>(S.pos_idx)[ej - blk_lo] = (unsigned char)rank;
    }
    __syncthreads();
    if (tid < T / 4) reinterpret_cast<uint32_t *>(work.pos_idx)[(size_t)block * (T / 4) + tid] = S.pos_idx[tid];
    SB_MARK("segtab")
    // ---- this block's internal nodes: spanning ones go to the list, the others are collapsed from shared memory
    bool spanning = false;
    RcTopo tp = {0, 0, 0, 0};
    if (has_node) {
        tp = S.topo[tid];
        spanning = tp.span_lo < blk_lo || tp.span_hi > blk_hi;
    }
    if (build_list) {
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, spanning);
        uint32_t base = 0;
        if ((tid & 31u) == 0u && bal) base = atomicAdd(work.span_count, (uint32_t)__popc(bal));
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (spanning) work.span_list[base + __popc(bal & ((1u << (tid & 31u)) - 1u))] = make_uint4(p1, tp.span_lo, tp.span_hi, 0u);
    }
    if (nodes4) {
        // a third of the nodes head no wide node: the others are compacted into a dense list first, so the (long) collapse runs with full warps
        const bool own = has_node && !spanning;
        const bool skip = own && p1 > 1u && tp.span_hi - tp.span_lo + 1u <= leaf_max;  // never the head of a wide node: the slot stays empty (zeroed: exported blobs are deterministic)
        if (skip) st_node4_zero(nodes4 + p1);
        const uint32_t lane = tid & 31u, wid = tid >> 5;
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, own && !skip);
        uint32_t *list = S.flag, *woff = S.par_leaf;  // the arrival counters and the leaves' parent links are dead since the climb's last barrier
        if (lane == 0) woff[wid] = __popc(bal);
        __syncthreads();
        if (wid == 0) {
            const uint32_t c = lane < T / 32 ? woff[lane] : 0u, inc = warp_incl_scan(c);
            woff[lane] = inc - c;
            if (lane == 31) woff[32] = inc;
        }
        __syncthreads();
        if (own && !skip) list[woff[wid] + __popc(bal & ((1u << lane) - 1u))] = p1;
        __syncthreads();
        SB_MARK("compact")
        if (tid < woff[32]) {
            const uint32_t v = list[tid];
            const RcNode4 nd = rc_collapse_node_t(
                v, [&](uint32_t c) -> RcBox { return box_from6(c >= n ? S.box_leaf[c - (n - 1) - blk_lo] : S.box_int[c - blk_lo]); },
                [&](uint32_t c) -> RcTopo { return S.topo[c - blk_lo]; }, n, leaf_max, leaf_map);
            st_node4(nodes4 + v, nd);
        }
        SB_MARK("collapse")
    }
}

__global__ void __launch_bounds__(FIT_T, FIT_T <= 512 ? 3 : 1) k_fit_local(const RcTri *__restrict__ tris_in, const uint32_t *__restrict__ perm, RcTri *__restrict__ tris,
                                                     const RcBox *__restrict__ inst_boxes, const uint32_t *__restrict__ leaf_map, const uint32_t *__restrict__ n_ptr,
                                                     uint32_t n_host, const RcTopo *__restrict__ topo, const uint32_t *__restrict__ parent, RcBox *__restrict__ boxes,
                                                     RcNode2 *__restrict__ nodes2, uint32_t *__restrict__ ctl, FitWork work, bool build_list, RcNode4 *__restrict__ nodes4,
                                                     uint32_t leaf_max) {
    extern __shared__ __align__(16) unsigned char fit_raw[];
    RC_PDL_TRIGGER();
    RC_PDL_WAIT();
    fit_local_body<FIT_T>(fit_raw, blockIdx.x, tris_in, perm, tris, inst_boxes, leaf_map, n_ptr, n_host, topo, parent, boxes, nodes2, ctl, work, build_list, nodes4, leaf_max);
}

// union of the leaf boxes of sorted positions [a, b] (1-based), which lie in different blocks, a at a segment start and b at a segment
// end of their blocks; cooperative over a group of FIT_G consecutive lanes (mask = the group's lanes), every lane returns the result
constexpr uint32_t FIT_G = 32;  // a warp per node: with 8-lane groups the few nodes that span hundreds of blocks take 4x longer and set the kernel time (52 vs 18 us)
template <int T>
__device__ __forceinline__ void fit_range_box(const FitSeg *__restrict__ seg, const unsigned char *__restrict__ pos_idx, uint32_t a, uint32_t b, uint32_t mask, f3 &lo,
                                              f3 &hi) {
    const uint32_t gl = threadIdx.x & (FIT_G - 1u);
    const uint32_t bA = (a - 1u) / T, bB = (b - 1u) / T;
    const FitSeg *tA = seg + (size_t)bA * FIT_SEG_MAX, *tB = seg + (size_t)bB * FIT_SEG_MAX;
    const uint32_t ia = pos_idx[a - 1u], ib = pos_idx[b - 1u];  // (the same address in every lane of the group: one transaction)
    lo = mk3(INFINITY, INFINITY, INFINITY);
    hi = mk3(-INFINITY, -INFINITY, -INFINITY);
    if (bB > bA + 1u) {  // whole blocks in between: lane-strided, then a butterfly over the group
        for (uint32_t k = bA + 1u + gl; k < bB; k += FIT_G) {
            const float *x = seg[(size_t)k * FIT_SEG_MAX].sfx;
            lo = jl_min3(lo, mk3(x[0], x[1], x[2]));
            hi = jl_max3(hi, mk3(x[3], x[4], x[5]));
        }
#pragma unroll
        for (int d = FIT_G / 2; d > 0; d >>= 1) {
            const f3 ol = mk3(__shfl_xor_sync(mask, lo.x, d), __shfl_xor_sync(mask, lo.y, d), __shfl_xor_sync(mask, lo.z, d));
            const f3 oh = mk3(__shfl_xor_sync(mask, hi.x, d), __shfl_xor_sync(mask, hi.y, d), __shfl_xor_sync(mask, hi.z, d));
            lo = jl_min3(lo, ol);
            hi = jl_max3(hi, oh);
        }
    }
    if (ia < (uint32_t)FIT_SEG_MAX && ib < (uint32_t)FIT_SEG_MAX && tA[ia].s == a && tB[ib].e == b) {
        const float *x = tA[ia].sfx, *y = tB[ib].pfx;
        lo = jl_min3(jl_min3(lo, mk3(x[0], x[1], x[2])), mk3(y[0], y[1], y[2]));
        hi = jl_max3(jl_max3(hi, mk3(x[3], x[4], x[5])), mk3(y[3], y[4], y[5]));
    } else {  // cannot happen for a radix tree (see k_fit_local); poison the box so that a broken invariant is loud, not subtle
        lo = hi = mk3(__int_as_float(0x7FFFFFFF), __int_as_float(0x7FFFFFFF), __int_as_float(0x7FFFFFFF));
    }
}
template <int T>
__device__ __forceinline__ bool fit_is_spanning(const RcTopo &t) { return (t.span_lo - 1u) / T != (t.span_hi - 1u) / T; }

// boxes (and BVH2 records) of the spanning nodes: a group of FIT_G lanes per node (the kernel is a chain of 3 dependent loads per node, so the
// parallelism is in nodes, not in lanes), no ordering between nodes
template <int T>
__device__ __forceinline__ void fit_span_body(const uint32_t *__restrict__ n_ptr, uint32_t n_host, const RcTopo *__restrict__ topo, const uint32_t *__restrict__ parent,
                                              RcBox *__restrict__ boxes, RcNode2 *__restrict__ nodes2, FitWork work) {
    const uint32_t n = count_of(n_ptr, n_host);
    const uint32_t count = *work.span_count, lane = threadIdx.x & 31u, gl = lane & (FIT_G - 1u);
    const uint32_t mask = (FIT_G == 32u ? 0xFFFFFFFFu : ((1u << (FIT_G & 31u)) - 1u) << (lane & ~(FIT_G - 1u)));
    const uint32_t groups = (gridDim.x * blockDim.x) / FIT_G;
    for (uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) / FIT_G; w < count; w += groups) {
        const uint4 ent = work.span_list[w];
        const uint32_t v = ent.x;
        f3 lo, hi;
        fit_range_box<T>(work.seg, work.pos_idx, ent.y, ent.z, mask, lo, hi);
        if (gl == 0) st_box(boxes + (v - 1), lo, hi);
        if (nodes2) {  // the BVH2 record holds the children's boxes: a spanning child's box comes from the same formula
            const RcTopo tp = topo[v - 1];
            f3 cl[2], ch[2];
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const uint32_t c = k == 0 ? tp.child0 : tp.child1;
                bool span_c = false;
                RcTopo tc = {0, 0, 0, 0};
                if (c < n) { tc = topo[c - 1]; span_c = fit_is_spanning<T>(tc); }
                if (span_c) {
                    fit_range_box<T>(work.seg, work.pos_idx, tc.span_lo, tc.span_hi, mask, cl[k], ch[k]);
                } else {  // written by k_fit_local (an earlier launch)
                    const RcBox b = boxes[c - 1];
                    cl[k] = mk3(b.lo[0], b.lo[1], b.lo[2]);
                    ch[k] = mk3(b.hi[0], b.hi[1], b.hi[2]);
                }
            }
            if (gl == 0) st_node2(nodes2 + (v - 1), cl[0], ch[0], cl[1], ch[1], tp.child0, tp.child1, parent[v - 1]);
        }
    }
}

__global__ void __launch_bounds__(256) k_fit_span(const uint32_t *__restrict__ n_ptr, uint32_t n_host, const RcTopo *__restrict__ topo, const uint32_t *__restrict__ parent,
                                                  RcBox *__restrict__ boxes, RcNode2 *__restrict__ nodes2, FitWork work) {
    RC_PDL_TRIGGER();
    RC_PDL_WAIT();
    fit_span_body<FIT_T>(n_ptr, n_host, topo, parent, boxes, nodes2, work);
}

// Wide nodes of the spanning BVH2 nodes (global arrays; the in-block ones were collapsed by k_fit_local), one thread per node.  A BVH2
// node covering <= leaf_max primitives can never be the root of a wide node (its parent turns it into a leaf reference): zeroed.
// The last block finishes the build:
//   hull (hull != null): the RC_HULL_BOXES subtrees four levels below the root — together they cover the whole BLAS — give the instance
//     bounds of the wide TLAS (union of the transformed hull boxes);
//   n == 1: the synthetic root over the single leaf;  the two wide-node slots nothing else writes (0, and n) are zeroed;
//   out10 (nullable): root box (6 floats); with ctl also the bounding sphere of a BLAS (centre of the scene bounds, radius^2 inflated
//     against the rounding of its own evaluation), next to the valid count, so the host reads everything back in one transfer.
__device__ __forceinline__ void collapse_span_body(const RcBox *__restrict__ boxes, const RcTopo *__restrict__ topo, const uint32_t *__restrict__ n_ptr, uint32_t n_host,
                                                   uint32_t leaf_max, const uint32_t *__restrict__ leaf_map, RcNode4 *__restrict__ nodes4, RcBox *__restrict__ hull,
                                                   FitWork work, float *__restrict__ out10, const uint32_t *__restrict__ ctl, const RcBox *__restrict__ root_boxes) {
    const uint32_t n = count_of(n_ptr, n_host);
    if (n == 0) return;
    if (blockIdx.x == gridDim.x - 1) {
        const uint32_t k = threadIdx.x;
        if (hull && k < RC_HULL_BOXES) {
            uint32_t node = 1;
#pragma unroll
            for (int l = 3; l >= 0; l--) {
                if (node >= n) break;  // a leaf (or the single-leaf tree): stays
                const RcTopo tp = topo[node - 1];
                node = (k >> l) & 1u ? tp.child1 : tp.child0;
            }
            hull[k] = n == 1 ? boxes[0] : boxes[node - 1];
        }
        if (nodes4) {
            if (k == 32 && n == 1) st_node4(nodes4 + 1, rc_collapse_node(1, boxes, topo, n, leaf_max, leaf_map));
            if (k == 33) st_node4_zero(nodes4);
            if (k == 34 && n > 1) st_node4_zero(nodes4 + n);
        }
        if (out10) {
            const RcBox *rb = root_boxes ? root_boxes : boxes;
            if (k >= 64 && k < 67) out10[k - 64] = rb[0].lo[k - 64];
            else if (k >= 67 && k < 70) out10[k - 64] = rb[0].hi[k - 67];
            else if (ctl && k >= 70 && k < 73) {
                const int a = k - 70;
                const f3 smin = ctl_bounds_min(ctl), smax = ctl_bounds_max(ctl);
                const float lo = a == 0 ? smin.x : (a == 1 ? smin.y : smin.z), hi = a == 0 ? smax.x : (a == 1 ? smax.y : smax.z);
                out10[6 + a] = 0.5f * (lo + hi);
            } else if (ctl && k == 73) {
                out10[9] = __uint_as_float(ctl[CTL_R2]) * 1.000002f;
            }
        }
        return;
    }
    if (!nodes4) return;
    // (entry w goes to block w mod nb: a short list is spread over the SMs instead of filling the first blocks' warps)
    const uint32_t count = *work.span_count, nb = gridDim.x - 1, threads = nb * blockDim.x;
    for (uint32_t w = threadIdx.x * nb + blockIdx.x; w < count; w += threads) {
        const uint4 ent = work.span_list[w];
        const uint32_t v = ent.x;
        if (v > 1u && ent.z - ent.y + 1u <= leaf_max) st_node4_zero(nodes4 + v);
        else st_node4(nodes4 + v, rc_collapse_node_cached(v, boxes, topo, n, leaf_max, leaf_map));
    }
}

__global__ void __launch_bounds__(256) k_collapse_span(const RcBox *__restrict__ boxes, const RcTopo *__restrict__ topo, const uint32_t *__restrict__ n_ptr, uint32_t n_host,
                                                       uint32_t leaf_max, const uint32_t *__restrict__ leaf_map, RcNode4 *__restrict__ nodes4, RcBox *__restrict__ hull,
                                                       FitWork work, float *__restrict__ out10, const uint32_t *__restrict__ ctl, const RcBox *__restrict__ root_boxes) {
    RC_PDL_TRIGGER();
    RC_PDL_WAIT();
    collapse_span_body(boxes, topo, n_ptr, n_host, leaf_max, leaf_map, nodes4, hull, work, out10, ctl, root_boxes);
}

// =================================================================================================
// Small meshes (<= SB_MAX_FACES faces): the WHOLE BLAS build as one cooperative kernel.  The five-launch path above spends a small build
// waiting: 8,192 faces are 4 radix tiles (4 working blocks per phase), every phase is a chain of L2 round trips, and the kernel
// boundaries cost as much as the kernels (131 us, of which the kernels themselves are 110).  Here a block (SB_T threads) owns SB_FT
// faces / sorted leaves / internal nodes in every phase, the grid is cdiv(faces, SB_FT) <= 128 blocks (all resident: one per SM), and
//   F   one face per thread: exact degenerate test, scene bounds, valid count of the block                                   | grid barrier
//   M   compacted RcTri records + Morton codes (the reference's arithmetic, as in k_front)                                    | grid barrier
//   S1  the block sorts its own run of SB_FT keys by counting (rank = keys of the run that are smaller, or equal and earlier: stable)    | grid barrier
//   S2  every block reads all runs into shared memory and places ITS keys: final position = rank in the run + for every other run the
//       number of its keys that are smaller (later runs) or not larger (earlier runs: they hold the earlier faces) — one binary
//       search per (key, run), 4 threads per key; sorted keys and the permutation go to global memory                                    | grid barrier
//       (A full shared-memory radix sort of all keys in every block — 4 passes of 8 bits, atomicOr peer masks for the ranks — took 22 us
//       of a 65 us build, a MATCH-based one 45 us; this merge of sorted runs is the same stable order for a fraction of the work.)
//   T   every block reads the sorted keys into shared memory; Karras topology of the block's own nodes from there; parent links scattered  | grid barrier
//   the fit of the five-launch path, its bodies reused as they are: fit_local_body<SB_FT> | barrier | fit_span_body<SB_FT> | barrier |
//   collapse_span_body (the last block finishes: hull, sphere, read-back words).
// Same arithmetic, same output arrays (tris, wide nodes, hull, BVH2 / kept topology when requested) byte for byte.
// =================================================================================================
#ifndef RC_SMALL_BUILD
#define RC_SMALL_BUILD 1
#endif
constexpr int SB_T = 1024;   // threads per block: the sort wants many warps
constexpr int SB_FT = 256;   // faces / sorted leaves / internal nodes a block owns in the other phases: 8,192 faces spread over 32 SMs (the climb and the
                             // collapse are issue-bound per SM: with 1,024 leaves per block they took 9 + 4.5 us of a 65 us build)
#ifndef RC_SB_ITEMS_MAX
#define RC_SB_ITEMS_MAX 32
#endif
constexpr int SB_ITEMS_MAX = RC_SB_ITEMS_MAX;  // x SB_T = the largest small build: 32,768 faces = 128 blocks (all co-resident: one block per SM), 128 KB of keys in shared memory
constexpr uint32_t SB_MAX_FACES = RC_SMALL_BUILD ? SB_T * SB_ITEMS_MAX : 0;
constexpr size_t SB_SORT_BYTES = (size_t)SB_MAX_FACES * 4 + SB_FT * 4;  // all keys + the rank accumulators of the block's run
constexpr size_t SB_SMEM_BYTES = SB_SORT_BYTES > sizeof(FitSmemT<SB_FT>) ? SB_SORT_BYTES : sizeof(FitSmemT<SB_FT>);
struct SmallArgs {
    const float *verts;
    const uint32_t *face_meta;  // nullable
    uint32_t n_faces;
    RcTri *tris_in;
    uint32_t *tile_counts;  // gridDim words
    uint32_t *ctl;
    uint32_t *keys;   // n_faces: Morton codes of the compacted faces; after the merge: the sorted codes
    uint32_t *perm;   // n_faces: sorted position -> compacted index
    uint32_t *run_keys, *run_idx;  // n_faces each: the sorted runs (SB_FT keys each) between the two sort phases
    RcTopo *topo;
    uint32_t *parent;
    RcTri *tris;
    RcBox *boxes;
    RcNode2 *nodes2;  // nullable
    RcNode4 *nodes4;
    RcBox *hull;
    uint32_t leaf_max;
    FitWork work;     // laid out for SB_FT leaves per block
};
// The stable sort of the small-build kernels (S1 / S2 above) + the sorted keys into every block's shared memory.
//   keys[0..n): codes in input order; on return: the sorted codes (also in skey[0..n)), perm[p] = input position of sorted position p.
// Every thread of every block calls it (two grid barriers inside).
__device__ __forceinline__ void sb_sort_runs(uint32_t *__restrict__ keys, uint32_t *__restrict__ run_keys, uint32_t *__restrict__ run_idx, uint32_t *__restrict__ perm,
                                             const uint32_t n, uint32_t *skey, uint32_t *bar, uint32_t &target) {
    const uint32_t tid = threadIdx.x;
    // ---- S1: the block's run (compacted faces [run0, run0 + len)) sorted by counting, 4 threads per key
    uint32_t *srank = skey + SB_MAX_FACES;
    const uint32_t runs = (n + SB_FT - 1) / SB_FT, run0 = blockIdx.x * SB_FT;
    const uint32_t len = blockIdx.x < runs ? min((uint32_t)SB_FT, n - run0) : 0u;
    const uint32_t ki = tid & (SB_FT - 1u), kq = tid / SB_FT;  // key of the run, quarter of the work on it
    if (tid < (uint32_t)SB_FT) {
        skey[tid] = tid < len ? __ldcg(keys + run0 + tid) : 0xFFFFFFFFu;  // (padding never counts: codes are < 2^30)
        srank[tid] = 0u;
    }
    __syncthreads();
    {
        const uint32_t mine = skey[ki];
        if (ki < len) {
            uint32_t c = 0;
#pragma unroll 8
            for (uint32_t j = kq * (SB_FT / 4); j < (kq + 1u) * (SB_FT / 4); j++) {
                const uint32_t other = skey[j];  // (the same address in every lane: a broadcast)
                c += (other < mine || (other == mine && j < ki)) ? 1u : 0u;
            }
            atomicAdd(&srank[ki], c);
        }
        __syncthreads();
        if (tid < len) {
            const uint32_t pos = run0 + srank[tid];
            run_keys[pos] = mine;
            run_idx[pos] = run0 + tid;
        }
    }
    SB_MARK("s1")
    grid_barrier(bar, target);
    SB_MARK("s1-bar")
    // ---- S2: all runs into shared memory; the block merges its own run into the final order
    for (uint32_t k = tid; k < n; k += SB_T) skey[k] = __ldcg(run_keys + k);
    if (tid < (uint32_t)SB_FT) srank[tid] = 0u;
    __syncthreads();
    SB_MARK("s2-load")
    {
        const uint32_t mine = ki < len ? skey[run0 + ki] : 0u;
        if (ki < len) {
            uint32_t c = 0;
            for (uint32_t r = kq; r < runs; r += SB_T / SB_FT) {
                if (r == blockIdx.x) continue;
                const uint32_t *rk = skey + r * SB_FT;
                const uint32_t rl = min((uint32_t)SB_FT, n - r * SB_FT);
                const uint32_t bound = mine + (r < blockIdx.x ? 1u : 0u);  // keys of earlier runs also precede when equal
                uint32_t lo = 0;  // number of keys of the run below the bound
#pragma unroll
                for (uint32_t step = SB_FT; step > 0; step >>= 1)
                    if (lo + step <= rl && rk[lo + step - 1u] < bound) lo += step;
                c += lo;
            }
            atomicAdd(&srank[ki], c);
        }
        __syncthreads();
        if (tid < len) {
            const uint32_t pos = tid + srank[tid];
            keys[pos] = mine;  // (the unsorted codes were last read before the barrier above)
            perm[pos] = __ldcg(run_idx + run0 + tid);
        }
    }
    SB_MARK("s2")
    grid_barrier(bar, target);
    SB_MARK("s2-bar")
    for (uint32_t k = tid; k < n; k += SB_T) skey[k] = __ldcg(keys + k);
    __syncthreads();
}
__global__ void __launch_bounds__(SB_T, 1) k_build_small(const SmallArgs A) {
    extern __shared__ __align__(16) unsigned char sb_raw[];
    __shared__ uint32_t sm[40];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, lt_mask = (1u << lane) - 1u;
    uint32_t target = 0;
    uint32_t *bar = A.ctl + CTL_BAR;
#ifdef RC_FRONT_PROF
    unsigned long long prof_t0 = prof_now(), prof_t[48];
    const char *prof_name[48];
    int prof_k = 0;
#endif
    // ---- F: one face per thread
    const uint32_t face = blockIdx.x * SB_FT + tid;
    bool valid = false;
    f3 a = mk3(0, 0, 0), b = a, c = a;
    f3 lo = mk3(INFINITY, INFINITY, INFINITY), hi = mk3(-INFINITY, -INFINITY, -INFINITY);
    if (tid < (uint32_t)SB_FT && face < A.n_faces) {
        const float *v = A.verts + (size_t)face * 9;
        a = ld3(v); b = ld3(v + 3); c = ld3(v + 6);
        valid = !x_is_degenerate(a, b, c);
        if (valid) {
            lo = jl_min3(jl_min3(a, b), c);  // world_bound(tri), triangle_mesh.jl:37
            hi = jl_max3(jl_max3(a, b), c);
        }
    }
    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, valid);
    uint32_t tile_total;
    const uint32_t woff = block_excl_scan(lane == 0 ? (uint32_t)__popc(bal) : 0u, sm, tile_total);  // lane 0 of warp w: valid faces of the warps before it
    const uint32_t wbase = __shfl_sync(0xFFFFFFFFu, woff, 0);
    if (tid == 0) A.tile_counts[blockIdx.x] = tile_total;
    bounds_atomic(A.ctl, lo, hi);
    RC_PROF_MARK("F")
    grid_barrier(bar, target);
    RC_PROF_MARK("F-barrier")
    // ---- M: compacted records (filter! keeps the order, src/instanced-bvh.jl:591-600) + Morton codes (kernels.jl:88-98; extent unguarded, :1388)
    uint32_t prefix = 0, n = 0;
    {   // one load per thread + a block scan (a serial loop over the <= 48 counts was 5 us: one L2 round trip each)
        const uint32_t cnt = tid < gridDim.x ? __ldcg(A.tile_counts + tid) : 0u;
        const uint32_t ex = block_excl_scan(cnt, sm, n);
        if (tid == blockIdx.x) sm[36] = ex;
        __syncthreads();
        prefix = sm[36];
    }
    if (blockIdx.x == 0 && tid == 0) A.ctl[CTL_N] = n;
    if (valid) {
        const f3 smin = ldcg3(A.ctl, CTL_BOUNDS, true), smax = ldcg3(A.ctl, CTL_BOUNDS + 3, false);
        const f3 ext = x_sub3(smax, smin);
        const uint32_t k = prefix + wbase + __popc(bal & lt_mask);
        float4 *d = reinterpret_cast<float4 *>(A.tris_in + k);
        d[0] = make_float4(a.x, a.y, a.z, __uint_as_float(k));
        d[1] = make_float4(b.x, b.y, b.z, __uint_as_float(A.face_meta ? A.face_meta[face] : face + 1u));  // :595
        d[2] = make_float4(c.x, c.y, c.z, __uint_as_float(face));
        const f3 ctr = mk3(x_mul(0.5f, x_add(lo.x, hi.x)), x_mul(0.5f, x_add(lo.y, hi.y)), x_mul(0.5f, x_add(lo.z, hi.z)));
        const f3 nrm = mk3(x_div(x_sub(ctr.x, smin.x), ext.x), x_div(x_sub(ctr.y, smin.y), ext.y), x_div(x_sub(ctr.z, smin.z), ext.z));
        A.keys[k] = rc_morton30(nrm);
    }
    RC_PROF_MARK("M")
    grid_barrier(bar, target);
    RC_PROF_MARK("M-barrier")
    if (n == 0) return;  // (uniform over the grid)
    uint32_t *skey = reinterpret_cast<uint32_t *>(sb_raw);
    sb_sort_runs(A.keys, A.run_keys, A.run_idx, A.perm, n, skey, bar, target);
    RC_PROF_MARK("S")
    SB_DUMP("sort")
    // ---- T: topology of this block's internal nodes (k_topology's arithmetic on the shared-memory keys), the block's share of the permutation
    {
        const uint32_t p1 = tid < (uint32_t)SB_FT ? blockIdx.x * SB_FT + tid + 1u : 0xFFFFFFF0u;  // internal node number / 1-based sorted position (none for the other threads)
        // the sorted triangle record is gathered here, behind the topology search (the fit then finds its triangles in place, like a refit)
        float4 g0 = make_float4(0, 0, 0, 0), g1 = g0, g2 = g0;
        if (p1 <= n) {
            const uint32_t k = __ldcg(A.perm + (p1 - 1u));
            const float4 *src = reinterpret_cast<const float4 *>(A.tris_in + k);
            g0 = __ldcg(src); g1 = __ldcg(src + 1); g2 = __ldcg(src + 2);
        }
        if (n == 1u && p1 == 1u) A.parent[0] = RC_INVALID;
        if (p1 < n) {
            const RcTopo t = rc_topology_for_node((int)p1, skey, (int)n);
            A.topo[p1 - 1u] = t;
            A.parent[t.child0 - 1u] = p1;  // set_parents_for_node, kernels.jl:159-180
            A.parent[t.child1 - 1u] = p1;
            if (p1 == 1u) A.parent[0] = RC_INVALID;
        }
        if (p1 <= n) {
            float4 *dst = reinterpret_cast<float4 *>(A.tris + (p1 - 1u));
            dst[0] = g0; dst[1] = g1; dst[2] = g2;
        }
    }
    RC_PROF_MARK("T")
    grid_barrier(bar, target);  // (its block barrier also retires the sort's shared memory: the fit lays its own arrays over it)
    RC_PROF_MARK("T-barrier")
    const uint32_t *n_ptr = A.ctl + CTL_N;
    fit_local_body<SB_FT>(sb_raw, blockIdx.x, nullptr, A.perm, A.tris, nullptr, nullptr, n_ptr, A.n_faces, A.topo, A.parent, A.boxes, A.nodes2, A.ctl, A.work, true, A.nodes4,
                         A.leaf_max);
    RC_PROF_MARK("fit-local")
    SB_DUMP("fit")
    grid_barrier(bar, target);
    RC_PROF_MARK("L-barrier")
    if (gridDim.x > 1) {
        fit_span_body<SB_FT>(n_ptr, A.n_faces, A.topo, A.parent, A.boxes, A.nodes2, A.work);
        RC_PROF_MARK("fit-span")
        grid_barrier(bar, target);
        RC_PROF_MARK("P-barrier")
    }
    collapse_span_body(A.boxes, A.topo, n_ptr, A.n_faces, A.leaf_max, nullptr, A.nodes4, A.hull, A.work, reinterpret_cast<float *>(A.ctl + CTL_OUT), A.ctl, nullptr);
    RC_PROF_MARK("collapse")
    RC_PROF_DUMP()
}
static bool launch_small(cudaStream_t st, SmallArgs &A, std::string &err) {
    static std::atomic<bool> configured[64];  // per device (zero-initialised; setting the attribute twice is harmless)
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !configured[dev].load(std::memory_order_relaxed)) {
        CK(cudaFuncSetAttribute(k_build_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SB_SMEM_BYTES));
        if (dev >= 0 && dev < 64) configured[dev].store(true, std::memory_order_relaxed);
    }
    void *args[] = {(void *)&A};
    CK(cudaLaunchCooperativeKernel((const void *)k_build_small, dim3(cdiv(A.n_faces, SB_FT)), dim3(SB_T), args, SB_SMEM_BYTES, st));
    return true;
}

// The same for a TLAS of <= SB_MAX_FACES instances — build (instance records + boxes, Morton codes, sort, topology, the two fits over one
// topology) or refit (records + boxes + the two fits) as one cooperative kernel instead of 12 / 8 launches: a one-instance TLAS (every
// sync! after a mesh update) was 0.13 ms of launch boundaries, the refit of 10,000 instances 0.19 ms.
struct TlasSmallArgs {
    const rc_instance_desc *inst;
    const float *blas_roots;
    const RcBlasPtrs *blas;
    uint32_t n;
    bool refit;  // topology, leaf_map and the spanning-node list are kept
    RcInstanceRec *rec;
    RcInstanceAux *aux;
    RcBox *inst_boxes, *inst_boxes_tight;
    uint32_t *ctl;
    uint32_t *keys, *run_keys, *run_idx;  // build only
    uint32_t *leaf_map;
    RcTopo *topo;
    uint32_t *parent;
    RcBox *boxes, *boxes_tight;
    RcNode2 *nodes2;
    RcNode4 *nodes4;
    FitWork work_ref, work_tight;  // two sets of segment tables (the fits run back to back), one spanning-node list
};
__device__ __forceinline__ void instance_record(const rc_instance_desc *d, const RcBlasPtrs &b, RcInstanceRec *rec, RcInstanceAux *aux) {
    RcInstanceRec r;
    for (int k = 0; k < 12; k++) r.inv[k] = d->inv_transform[k];
    r.nodes4 = b.nodes4;
    r.tris = b.tris;
    for (int k = 0; k < 4; k++) r.sphere[k] = b.sphere[k];
    rc_world_sphere(d->transform, b.sphere, r.wsphere);
    *rec = r;
    RcInstanceAux a;
    a.nodes2 = b.nodes2;
    a.n_prims = b.n;
    a.custom_index = d->instance_id;
    *aux = a;
}
// reference box (8 corners of the BLAS root box, kernels.jl:38-62) and the tighter hull-derived box of one instance
__device__ __forceinline__ void instance_boxes(const rc_instance_desc *d, const float *blas_roots, const RcBlasPtrs *blas, f3 &lo, f3 &hi, f3 &tl, f3 &th) {
    rc_instance_world_aabb(d->transform, blas_roots + 6 * (d->blas_index - 1), lo, hi);
    const RcBox *hull = blas[d->blas_index - 1].hull;
    tl = mk3(INFINITY, INFINITY, INFINITY);
    th = mk3(-INFINITY, -INFINITY, -INFINITY);
    for (int k = 0; k < RC_HULL_BOXES; k++) {
        RcBox hb = hull[k];
        if (!(hb.lo[0] <= hb.hi[0])) continue;
        float loc[6] = {hb.lo[0], hb.lo[1], hb.lo[2], hb.hi[0], hb.hi[1], hb.hi[2]};
        f3 a, b;
        rc_instance_world_aabb(d->transform, loc, a, b);
        tl = mk3(fminf(tl.x, a.x), fminf(tl.y, a.y), fminf(tl.z, a.z));
        th = mk3(fmaxf(th.x, b.x), fmaxf(th.y, b.y), fmaxf(th.z, b.z));
    }
    tl = mk3(fmaxf(tl.x, lo.x), fmaxf(tl.y, lo.y), fmaxf(tl.z, lo.z));  // never larger than the reference box
    th = mk3(fminf(th.x, hi.x), fminf(th.y, hi.y), fminf(th.z, hi.z));
}
__global__ void __launch_bounds__(SB_T, 1) k_tlas_small(const TlasSmallArgs A) {
    extern __shared__ __align__(16) unsigned char sb_raw[];
    const uint32_t tid = threadIdx.x, n = A.n;
    uint32_t target = 0;
    uint32_t *bar = A.ctl + CTL_BAR;
    // ---- I: records, reference boxes (+ scene bounds for a build), tight boxes; one instance per thread (SB_FT per block)
    const uint32_t i = blockIdx.x * SB_FT + tid;
    const bool mine = tid < (uint32_t)SB_FT && i < n;
    f3 lo = mk3(INFINITY, INFINITY, INFINITY), hi = mk3(-INFINITY, -INFINITY, -INFINITY);
    if (mine) {
        const rc_instance_desc *d = A.inst + i;
        instance_record(d, A.blas[d->blas_index - 1], A.rec + i, A.aux + i);
        f3 tl, th;
        instance_boxes(d, A.blas_roots, A.blas, lo, hi, tl, th);
        st_box(A.inst_boxes + i, lo, hi);
        st_box(A.inst_boxes_tight + i, tl, th);
    }
    if (!A.refit) {
        bounds_atomic(A.ctl, lo, hi);
        grid_barrier(bar, target);
        // ---- M: calculate_tlas_morton_code, kernels.jl:295-313; extent clamp :1517-1521
        if (mine) {
            const f3 smin = ldcg3(A.ctl, CTL_BOUNDS, true), smax = ldcg3(A.ctl, CTL_BOUNDS + 3, false);
            const f3 ext = mk3(jl_max(x_sub(smax.x, smin.x), 1e-6f), jl_max(x_sub(smax.y, smin.y), 1e-6f), jl_max(x_sub(smax.z, smin.z), 1e-6f));
            const rc_instance_desc *d = A.inst + i;
            const float *la = A.blas_roots + 6 * (d->blas_index - 1);
            const f3 lc = mk3(x_mul(0.5f, x_add(la[0], la[3])), x_mul(0.5f, x_add(la[1], la[4])), x_mul(0.5f, x_add(la[2], la[5])));
            const f3 wc = x_transform_point(d->transform, lc);
            const f3 nrm = mk3(x_div(x_sub(wc.x, smin.x), ext.x), x_div(x_sub(wc.y, smin.y), ext.y), x_div(x_sub(wc.z, smin.z), ext.z));
            A.keys[i] = rc_morton30(nrm);
        }
        grid_barrier(bar, target);
        uint32_t *skey = reinterpret_cast<uint32_t *>(sb_raw);
        sb_sort_runs(A.keys, A.run_keys, A.run_idx, A.leaf_map, n, skey, bar, target);  // leaf_map = sorted position -> instance index
        // ---- T
        const uint32_t p1 = tid < (uint32_t)SB_FT ? blockIdx.x * SB_FT + tid + 1u : 0xFFFFFFF0u;
        if (n == 1u && p1 == 1u) A.parent[0] = RC_INVALID;
        if (p1 < n) {
            const RcTopo t = rc_topology_for_node((int)p1, skey, (int)n);
            A.topo[p1 - 1u] = t;
            A.parent[t.child0 - 1u] = p1;
            A.parent[t.child1 - 1u] = p1;
            if (p1 == 1u) A.parent[0] = RC_INVALID;
        }
    }
    grid_barrier(bar, target);
    // ---- the two fits over the one topology: reference boxes -> BVH2 + root box, tight boxes -> wide nodes
    fit_local_body<SB_FT>(sb_raw, blockIdx.x, nullptr, nullptr, nullptr, A.inst_boxes, A.leaf_map, nullptr, n, A.topo, A.parent, A.boxes, A.nodes2, nullptr, A.work_ref, !A.refit,
                          nullptr, 1u);
    __syncthreads();
    fit_local_body<SB_FT>(sb_raw, blockIdx.x, nullptr, nullptr, nullptr, A.inst_boxes_tight, A.leaf_map, nullptr, n, A.topo, A.parent, A.boxes_tight, nullptr, nullptr, A.work_tight,
                          false, A.nodes4, 1u);
    grid_barrier(bar, target);
    if (gridDim.x > 1) {
        fit_span_body<SB_FT>(nullptr, n, A.topo, A.parent, A.boxes, A.nodes2, A.work_ref);
        fit_span_body<SB_FT>(nullptr, n, A.topo, A.parent, A.boxes_tight, nullptr, A.work_tight);
        grid_barrier(bar, target);
    }
    collapse_span_body(A.boxes_tight, A.topo, nullptr, n, 1u, A.leaf_map, A.nodes4, nullptr, A.work_tight, reinterpret_cast<float *>(A.ctl + CTL_OUT), nullptr, A.boxes);
    // the last block out leaves the barrier's arrival counter at zero for the next refit (every block has passed the last barrier by then)
    if (tid == 0 && atomicAdd(bar + 1, 1u) == gridDim.x - 1u) { bar[0] = 0u; bar[1] = 0u; }
}
static bool launch_tlas_small(cudaStream_t st, TlasSmallArgs &A, std::string &err) {
    static std::atomic<bool> configured[64];
    int dev = 0;
    CK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !configured[dev].load(std::memory_order_relaxed)) {
        CK(cudaFuncSetAttribute(k_tlas_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SB_SMEM_BYTES));
        if (dev >= 0 && dev < 64) configured[dev].store(true, std::memory_order_relaxed);
    }
    void *args[] = {(void *)&A};
    CK(cudaLaunchCooperativeKernel((const void *)k_tlas_small, dim3(cdiv(A.n, SB_FT)), dim3(SB_T), args, SB_SMEM_BYTES, st));
    return true;
}

// Largest input the small-build kernels take on the current device: every block of the cooperative grid must be resident at once (one
// block of SB_T threads and SB_SMEM_BYTES per SM), so the limit is min(SB_MAX_FACES, SMs x SB_FT); 0 when the kernels do not fit at all.
static uint32_t small_build_limit() {
    static std::atomic<uint32_t> limit[64];  // per device, 0 = not asked yet, 1 = unusable
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
    uint32_t v = limit[dev].load(std::memory_order_relaxed);
    if (v == 0) {
        int sms = 0, per_sm_b = 0, per_sm_t = 0;
        bool ok = cudaFuncSetAttribute(k_build_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SB_SMEM_BYTES) == cudaSuccess &&
                  cudaFuncSetAttribute(k_tlas_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SB_SMEM_BYTES) == cudaSuccess &&
                  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess &&
                  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_b, k_build_small, SB_T, SB_SMEM_BYTES) == cudaSuccess &&
                  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_t, k_tlas_small, SB_T, SB_SMEM_BYTES) == cudaSuccess;
        if (!ok) (void)cudaGetLastError();
        const uint64_t blocks = ok ? (uint64_t)sms * (uint64_t)std::min(per_sm_b, per_sm_t) : 0;
        v = (uint32_t)std::min<uint64_t>(SB_MAX_FACES, blocks * SB_FT);
        if (v == 0) v = 1;
        limit[dev].store(v, std::memory_order_relaxed);
    }
    return v == 1 ? 0 : v;
}

// One fit of a built topology: leaf boxes -> boxes of every node (+ BVH2 records) (+ wide nodes, hull, finishing read-back words).
struct FitJob {
    uint32_t n_bound = 0;            // sizes the grids; the live count is *n_ptr (device) or n_bound itself
    const uint32_t *n_ptr = nullptr;
    const RcTri *tris_in = nullptr;  // leaf source, see k_fit_local
    const uint32_t *perm = nullptr;
    RcTri *tris = nullptr;
    const RcBox *inst_boxes = nullptr;
    const uint32_t *leaf_map = nullptr;
    const RcTopo *topo = nullptr;
    const uint32_t *parent = nullptr;
    RcBox *boxes = nullptr;
    RcNode2 *nodes2 = nullptr;       // nullable
    RcNode4 *nodes4 = nullptr;       // nullable: no collapse
    uint32_t leaf_max = 1;
    RcBox *hull = nullptr;           // nullable
    uint32_t *ctl = nullptr;         // BLAS control block (bounds -> sphere), nullable
    float *out10 = nullptr;          // nullable: root box (+ sphere) for the host
    const RcBox *root_boxes = nullptr;  // out10's root box comes from this fit's boxes unless given (TLAS: the reference boxes)
    bool build_list = true;          // false: work.span_list / span_count already hold this topology's spanning nodes
};
static void run_fit(cudaStream_t st, const FitJob &j, const FitWork &work) {
    // the opt-in to > 48 KB of dynamic shared memory is a per-device function attribute (several devices build concurrently under rc_multi_*)
    static std::atomic<bool> configured[64];  // (zero-initialised; setting the attribute twice is harmless)
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev].load(std::memory_order_relaxed)) {
        cudaFuncSetAttribute(k_fit_local, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FitSmem));
        if (dev >= 0 && dev < 64) configured[dev].store(true, std::memory_order_relaxed);
    }
    const uint32_t blocks = cdiv(j.n_bound, FIT_T);
    const bool pdl = j.n_bound >= PDL_MIN_ELEMENTS;
    launch_dependent(pdl, k_fit_local, blocks, FIT_T, sizeof(FitSmem), st, j.tris_in, j.perm, j.tris, j.inst_boxes, j.leaf_map, j.n_ptr, j.n_bound, j.topo, j.parent, j.boxes, j.nodes2,
                     j.ctl, work, j.build_list, j.nodes4, j.leaf_max);
    // at most FIT_SEG_MAX spanning nodes per block; a group of FIT_G lanes each, grid-stride
    const uint32_t span_bound = blocks > 1 ? blocks * FIT_SEG_MAX : 0u;
    if (span_bound) launch_dependent(pdl, k_fit_span, std::min(cdiv(span_bound, 256 / FIT_G), 148u * 8u), 256, 0, st, j.n_ptr, j.n_bound, j.topo, j.parent, j.boxes, j.nodes2, work);
    if (j.nodes4 || j.hull || j.out10)
        launch_dependent(pdl, k_collapse_span, (span_bound && j.nodes4 ? std::min(cdiv(span_bound, 256), 148u * 4u) : 0u) + 1u, 256, 0, st, j.boxes, j.topo, j.n_ptr, j.n_bound, j.leaf_max,
                         j.leaf_map, j.nodes4, j.hull, work, j.out10, j.ctl, j.root_boxes);
}

// =================================================================================================
// Public (library-internal) entry points
// =================================================================================================
void rc_free_blas(RcDeviceBlas *b, cudaStream_t st) {
    if (!b) return;
    for (void *p : {(void *)b->nodes2, (void *)b->nodes4, (void *)b->tris, (void *)b->hull, (void *)b->normals, (void *)b->topo, (void *)b->parent})
        if (p) cudaFreeAsync(p, st);
    *b = RcDeviceBlas();
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

// read back {valid count, root box, sphere} and finish the host-side record
// page-locked landing zone of the builders' small read-backs, one per host thread (a pageable destination turns the copy into a staged,
// blocking transfer); falls back to the caller's stack buffer when the allocation fails
static uint32_t *pinned_words(uint32_t *fallback) {
    thread_local uint32_t *p = nullptr;
    thread_local bool tried = false;
    if (!tried) {
        tried = true;
        void *q = nullptr;
        if (cudaHostAlloc(&q, 256, cudaHostAllocPortable) == cudaSuccess) p = static_cast<uint32_t *>(q);
        else (void)cudaGetLastError();
    }
    return p ? p : fallback;
}
static bool finish_blas(cudaStream_t st, uint32_t *d_ctl, RcDeviceBlas *out, std::string &err) {
    uint32_t h_stack[CTL_OUT + 10];
    uint32_t *h = pinned_words(h_stack);
    static_assert(sizeof h_stack <= 256, "pinned landing zone");
    CK(cudaMemcpyAsync(h, d_ctl, sizeof h_stack, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    out->n = h[CTL_N];
    if (h[CTL_ERR]) { err = "internal error: fit segment table overflow"; return false; }
    if (out->n == 0) { err = "Geometry has no valid triangles"; return false; }  // src/instanced-bvh.jl:601
    memcpy(out->root_aabb, h + CTL_OUT, 24);
    memcpy(out->sphere, h + CTL_OUT + 6, 16);
    return extent_supported(out->root_aabb, err);
}

bool rc_build_blas(cudaStream_t st, const float *d_verts, const uint32_t *d_face_meta, uint32_t n_faces, uint32_t build_flags, RcDeviceBlas *out, std::string &err) {
    *out = RcDeviceBlas();
    if (n_faces == 0) { err = "Geometry has no valid triangles"; return false; }
    if (n_faces > RC_LEAF_START_MASK - 16u) { err = "BLAS too large (max 2^28 triangles)"; return false; }
    const uint32_t nf = n_faces;  // upper bound of the valid count: sizes every array and grid; the live count stays on the device
    const bool keep_bvh2 = build_flags & RC_BUILD_KEEP_BVH2, keep_topo = build_flags & RC_BUILD_ALLOW_REFIT;
    RcTemps tmp(st);
    const uint32_t f_tiles = cdiv(nf, FT_TILE), s_tiles = cdiv(nf, RS_TILE);
    uint32_t *d_ctl = nullptr;
    RcTri *d_tris_in = nullptr;
    RcBox *d_boxes = nullptr;
    uint32_t *d_codes = nullptr, *d_idx = nullptr, *d_codes2 = nullptr, *d_idx2 = nullptr, *d_hist = nullptr, *d_parent = nullptr;
    RcTopo *d_topo = nullptr;
    const bool small = nf <= small_build_limit();  // one cooperative kernel instead of the five launches (k_build_small)
    // every temporary of the build comes out of ONE stream-ordered allocation: a small build is as long as a dozen cudaMallocAsync /
    // cudaFreeAsync calls on the host, and the GPU waits for the launch behind them
    size_t arena_bytes = 0;
    auto carve = [&](size_t bytes) { const size_t at = arena_bytes; arena_bytes += (bytes + 255) & ~(size_t)255; return at; };
    const size_t o_ctl = carve(sizeof(uint32_t) * (CTL_WORDS + std::max(f_tiles, cdiv(nf, SB_FT))));  // control block + the filter's per-tile valid counts
    const size_t o_tris_in = carve(sizeof(RcTri) * (size_t)nf);
    const size_t o_codes = carve(sizeof(uint32_t) * (size_t)nf), o_idx = carve(sizeof(uint32_t) * (size_t)nf);
    const size_t o_codes2 = carve(sizeof(uint32_t) * (size_t)nf), o_idx2 = carve(sizeof(uint32_t) * (size_t)nf);
    const size_t o_hist = small ? 0 : carve(sizeof(uint32_t) * radix_hist_words(nf));
    const size_t o_work = carve(small ? fit_work_bytes(nf, SB_FT) : fit_work_bytes(nf));
    const size_t o_boxes = carve(sizeof(RcBox) * 2 * (size_t)nf);
    const size_t o_topo = keep_topo ? 0 : carve(sizeof(RcTopo) * (size_t)nf), o_parent = keep_topo ? 0 : carve(sizeof(uint32_t) * 2 * (size_t)nf);
    unsigned char *arena = nullptr;
    TMP(arena, arena_bytes);
    d_ctl = reinterpret_cast<uint32_t *>(arena + o_ctl);
    d_tris_in = reinterpret_cast<RcTri *>(arena + o_tris_in);
    d_codes = reinterpret_cast<uint32_t *>(arena + o_codes);
    d_idx = reinterpret_cast<uint32_t *>(arena + o_idx);
    d_codes2 = reinterpret_cast<uint32_t *>(arena + o_codes2);
    d_idx2 = reinterpret_cast<uint32_t *>(arena + o_idx2);
    if (!small) d_hist = reinterpret_cast<uint32_t *>(arena + o_hist);
    unsigned char *d_work = arena + o_work;
    d_boxes = reinterpret_cast<RcBox *>(arena + o_boxes);
    // the results outlive this call; on failure the caller releases them with rc_free_blas
    if (keep_topo) {
        CK(cudaMallocAsync(&out->topo, sizeof(RcTopo) * (size_t)nf, st));
        CK(cudaMallocAsync(&out->parent, sizeof(uint32_t) * 2 * (size_t)nf, st));
        d_topo = out->topo;
        d_parent = out->parent;
    } else {
        d_topo = reinterpret_cast<RcTopo *>(arena + o_topo);
        d_parent = reinterpret_cast<uint32_t *>(arena + o_parent);
    }
    if (keep_bvh2) CK(cudaMallocAsync(&out->nodes2, sizeof(RcNode2) * 2 * (size_t)nf, st));
    CK(cudaMallocAsync(&out->nodes4, sizeof(RcNode4) * ((size_t)nf + 1), st));
    CK(cudaMallocAsync(&out->tris, sizeof(RcTri) * (size_t)nf, st));
    CK(cudaMallocAsync(&out->hull, sizeof(RcBox) * RC_HULL_BOXES, st));
    out->n_faces_in = n_faces;

    CK(cudaMemsetAsync(d_ctl, 0, sizeof(uint32_t) * CTL_WORDS, st));
    if (small) {
        SmallArgs sa;
        sa.verts = d_verts; sa.face_meta = d_face_meta; sa.n_faces = n_faces; sa.tris_in = d_tris_in;
        sa.tile_counts = d_ctl + CTL_WORDS; sa.ctl = d_ctl; sa.keys = d_codes; sa.perm = d_idx; sa.run_keys = d_codes2; sa.run_idx = d_idx2;
        sa.topo = d_topo; sa.parent = d_parent; sa.tris = out->tris; sa.boxes = d_boxes; sa.nodes2 = out->nodes2;
        sa.nodes4 = out->nodes4; sa.hull = out->hull; sa.leaf_max = RC_BLAS_LEAF_MAX;
        sa.work = fit_work_at(d_work, nf, SB_FT);
        sa.work.span_count = d_ctl + CTL_NSPAN;  // zeroed with the control block
        sa.work.err_flag = d_ctl + CTL_ERR;
        if (!launch_small(st, sa, err)) return false;
        return finish_blas(st, d_ctl, out, err);
    }
    FrontArgs fa;
    fa.verts = d_verts; fa.face_meta = d_face_meta; fa.n_faces = n_faces; fa.tris_in = d_tris_in;
    fa.tile_counts = d_ctl + CTL_WORDS; fa.ctl = d_ctl;
    fa.keys0 = d_codes; fa.vals0 = d_idx; fa.keys1 = d_codes2; fa.vals1 = d_idx2;
    fa.n_host = 0; fa.hist = d_hist; fa.tiles_stride = s_tiles; fa.bar = d_ctl + CTL_BAR;
    if (!launch_front(st, fa, f_tiles, err)) return false;
    uint32_t *codes_sorted = d_codes2, *perm = d_idx2;
    launch_dependent(nf >= PDL_MIN_ELEMENTS, k_topology, cdiv(std::max(1u, nf - 1), 256), 256, 0, st, codes_sorted, d_ctl + CTL_N, nf, d_topo, d_parent);
    FitWork work = fit_work_at(d_work, nf);
    work.span_count = d_ctl + CTL_NSPAN;  // zeroed with the control block
    work.err_flag = d_ctl + CTL_ERR;
    FitJob job;
    job.n_bound = nf; job.n_ptr = d_ctl + CTL_N;
    job.tris_in = d_tris_in; job.perm = perm; job.tris = out->tris;
    job.topo = d_topo; job.parent = d_parent; job.boxes = d_boxes; job.nodes2 = out->nodes2;
    job.nodes4 = out->nodes4; job.leaf_max = RC_BLAS_LEAF_MAX; job.hull = out->hull;
    job.ctl = d_ctl; job.out10 = reinterpret_cast<float *>(d_ctl + CTL_OUT);
    run_fit(st, job, work);
    return finish_blas(st, d_ctl, out, err);
}

// Vertex update with unchanged topology (update!, src/instanced-bvh.jl:808-857, as a refit): the kept radix tree is re-fitted to the new
// vertex positions of the same faces.  Allowed when the geometry was built with RC_BUILD_ALLOW_REFIT, the face count is unchanged and
// the set of degenerate faces is the same (*refitted = false otherwise: the caller rebuilds).
// pass A (blockIdx.y == 0): a kept face that is degenerate in the new soup is counted; pass B (blockIdx.y == 1): the valid faces of the
// new soup are counted (must equal n).  Together: the degenerate set is unchanged.  Nothing is written to the geometry.
__global__ void k_refit_check(const float *__restrict__ verts, uint32_t n_faces, const RcTri *__restrict__ tris, uint32_t n, uint32_t *__restrict__ ctl) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (blockIdx.y == 0) {
        if (i < n) {
            const float *v = verts + (size_t)tris[i].face_index * 9;
            if (x_is_degenerate(ld3(v), ld3(v + 3), ld3(v + 6))) atomicAdd(&ctl[CTL_TILE], 1u);
        }
    } else if (i < n_faces) {
        const float *v = verts + (size_t)i * 9;
        if (!x_is_degenerate(ld3(v), ld3(v + 3), ld3(v + 6))) atomicAdd(&ctl[CTL_TILE + 1], 1u);
    }
}
// every sorted triangle takes its face's new vertices (ids kept); scene bounds for the bounding sphere
__global__ void k_refit_apply(const float *__restrict__ verts, RcTri *__restrict__ tris, uint32_t n, uint32_t *__restrict__ ctl) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    f3 lo = mk3(INFINITY, INFINITY, INFINITY), hi = mk3(-INFINITY, -INFINITY, -INFINITY);
    if (i < n) {
        float4 *t = reinterpret_cast<float4 *>(tris + i);
        const float4 t0 = t[0], t1 = t[1], t2 = t[2];
        const float *v = verts + (size_t)__float_as_uint(t2.w) * 9;
        const f3 a = ld3(v), b = ld3(v + 3), c = ld3(v + 6);
        t[0] = make_float4(a.x, a.y, a.z, t0.w);
        t[1] = make_float4(b.x, b.y, b.z, t1.w);
        t[2] = make_float4(c.x, c.y, c.z, t2.w);
        lo = jl_min3(jl_min3(a, b), c);
        hi = jl_max3(jl_max3(a, b), c);
    }
    bounds_atomic(ctl, lo, hi);
}

// The vertex-update refit of a small geometry (<= small_build_limit() faces) as one cooperative kernel, with no host decision in the middle:
// the check (is the degenerate set unchanged?) is a phase, its verdict is uniform over the grid after a barrier, and a refused refit leaves
// the geometry untouched and says so in ctl[CTL_REFUSED] (the multi-launch path reads the two counts back and decides on the host).
enum { CTL_REFUSED = 12 };
struct RefitSmallArgs {
    const float *verts;
    uint32_t n_faces, n;
    RcTri *tris;
    const RcTopo *topo;
    const uint32_t *parent;
    RcBox *boxes;
    RcNode2 *nodes2;
    RcNode4 *nodes4;
    RcBox *hull;
    uint32_t *ctl;
    FitWork work;  // laid out for SB_FT leaves per block
};
__global__ void __launch_bounds__(SB_T, 1) k_refit_small(const RefitSmallArgs A) {
    extern __shared__ __align__(16) unsigned char sb_raw[];
    __shared__ uint32_t red[2];
    const uint32_t tid = threadIdx.x, n = A.n;
    uint32_t target = 0;
    uint32_t *bar = A.ctl + CTL_BAR;
    const uint32_t i = tid < (uint32_t)SB_FT ? blockIdx.x * SB_FT + tid : 0xFFFFFFFFu;
    // ---- check: a kept face that is degenerate in the new soup / the valid faces of the new soup (must equal n)
    if (tid < 2) red[tid] = 0;
    __syncthreads();
    if (i < n) {
        const float *v = A.verts + (size_t)A.tris[i].face_index * 9;
        if (x_is_degenerate(ld3(v), ld3(v + 3), ld3(v + 6))) atomicAdd(&red[0], 1u);
    }
    if (i < A.n_faces) {
        const float *v = A.verts + (size_t)i * 9;
        if (!x_is_degenerate(ld3(v), ld3(v + 3), ld3(v + 6))) atomicAdd(&red[1], 1u);
    }
    __syncthreads();
    if (tid == 0) {
        if (red[0]) atomicAdd(&A.ctl[CTL_TILE], red[0]);
        if (red[1]) atomicAdd(&A.ctl[CTL_TILE + 1], red[1]);
        if (blockIdx.x == 0) A.ctl[CTL_N] = n;
    }
    grid_barrier(bar, target);
    if (__ldcg(A.ctl + CTL_TILE) != 0u || __ldcg(A.ctl + CTL_TILE + 1) != n) {  // (uniform over the grid)
        if (blockIdx.x == 0 && tid == 0) A.ctl[CTL_REFUSED] = 1u;
        return;
    }
    // ---- apply: every sorted triangle takes its face's new vertices (ids kept); scene bounds for the bounding sphere
    f3 lo = mk3(INFINITY, INFINITY, INFINITY), hi = mk3(-INFINITY, -INFINITY, -INFINITY);
    if (i < n) {
        float4 *t = reinterpret_cast<float4 *>(A.tris + i);
        const float4 t0 = t[0], t1 = t[1], t2 = t[2];
        const float *v = A.verts + (size_t)__float_as_uint(t2.w) * 9;
        const f3 a = ld3(v), b = ld3(v + 3), c = ld3(v + 6);
        t[0] = make_float4(a.x, a.y, a.z, t0.w);
        t[1] = make_float4(b.x, b.y, b.z, t1.w);
        t[2] = make_float4(c.x, c.y, c.z, t2.w);
        lo = jl_min3(jl_min3(a, b), c);
        hi = jl_max3(jl_max3(a, b), c);
    }
    bounds_atomic(A.ctl, lo, hi);
    grid_barrier(bar, target);
    const uint32_t *n_ptr = A.ctl + CTL_N;
    fit_local_body<SB_FT>(sb_raw, blockIdx.x, nullptr, nullptr, A.tris, nullptr, nullptr, n_ptr, n, A.topo, A.parent, A.boxes, A.nodes2, A.ctl, A.work, true, A.nodes4, RC_BLAS_LEAF_MAX);
    grid_barrier(bar, target);
    if (gridDim.x > 1) {
        fit_span_body<SB_FT>(n_ptr, n, A.topo, A.parent, A.boxes, A.nodes2, A.work);
        grid_barrier(bar, target);
    }
    collapse_span_body(A.boxes, A.topo, n_ptr, n, RC_BLAS_LEAF_MAX, nullptr, A.nodes4, A.hull, A.work, reinterpret_cast<float *>(A.ctl + CTL_OUT), A.ctl, nullptr);
}

bool rc_refit_blas(cudaStream_t st, const float *d_verts, uint32_t n_faces, RcDeviceBlas *b, bool *refitted, std::string &err) {
    *refitted = false;
    if (!b->topo || !b->parent || n_faces != b->n_faces_in || b->n == 0) return true;
    const uint32_t n = b->n;
    RcTemps tmp(st);
    uint32_t *d_ctl = nullptr;
    unsigned char *d_work = nullptr;
    RcBox *d_boxes = nullptr;
    const bool small = n_faces <= small_build_limit();
    TMP(d_ctl, CTL_WORDS);
    TMP(d_work, small ? fit_work_bytes(n, SB_FT) : fit_work_bytes(n));
    TMP(d_boxes, 2 * (size_t)n);
    CK(cudaMemsetAsync(d_ctl, 0, sizeof(uint32_t) * CTL_WORDS, st));
    if (small) {
        static std::atomic<bool> configured[64];
        int dev = 0;
        CK(cudaGetDevice(&dev));
        if (dev < 0 || dev >= 64 || !configured[dev].load(std::memory_order_relaxed)) {
            CK(cudaFuncSetAttribute(k_refit_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SB_SMEM_BYTES));
            if (dev >= 0 && dev < 64) configured[dev].store(true, std::memory_order_relaxed);
        }
        RefitSmallArgs ra;
        ra.verts = d_verts; ra.n_faces = n_faces; ra.n = n; ra.tris = b->tris; ra.topo = b->topo; ra.parent = b->parent;
        ra.boxes = d_boxes; ra.nodes2 = b->nodes2; ra.nodes4 = b->nodes4; ra.hull = b->hull; ra.ctl = d_ctl;
        ra.work = fit_work_at(d_work, n, SB_FT);
        ra.work.span_count = d_ctl + CTL_NSPAN;
        ra.work.err_flag = d_ctl + CTL_ERR;
        void *args[] = {(void *)&ra};
        CK(cudaLaunchCooperativeKernel((const void *)k_refit_small, dim3(cdiv(n_faces, SB_FT)), dim3(SB_T), args, SB_SMEM_BYTES, st));
        uint32_t h_stack[CTL_OUT + 10];
        uint32_t *h = pinned_words(h_stack);
        CK(cudaMemcpyAsync(h, d_ctl, sizeof h_stack, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        CK(cudaGetLastError());
        if (h[CTL_REFUSED]) return true;  // the degenerate set changed: primitive numbering would differ (the geometry was not touched)
        if (h[CTL_ERR]) { err = "internal error: fit segment table overflow"; return false; }
        float root[6];
        memcpy(root, h + CTL_OUT, 24);
        if (!extent_supported(root, err)) return false;
        memcpy(b->root_aabb, root, 24);
        memcpy(b->sphere, h + CTL_OUT + 6, 16);
        *refitted = true;
        return true;
    }
    // check first, touch the geometry only when the refit is certain (a refused update leaves the old geometry intact, as update! does, :837)
    dim3 grid(cdiv(std::max(n, n_faces), 256), 2);
    k_refit_check<<<grid, 256, 0, st>>>(d_verts, n_faces, b->tris, n, d_ctl);
    uint32_t h[CTL_TILE + 2];
    CK(cudaMemcpyAsync(h, d_ctl, sizeof h, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (h[CTL_TILE] != 0 || h[CTL_TILE + 1] != n) return true;  // the degenerate set changed: primitive numbering would differ
    CK(cudaMemsetAsync(d_ctl + CTL_TILE, 0, 8, st));
    k_refit_apply<<<cdiv(n, 256), 256, 0, st>>>(d_verts, b->tris, n, d_ctl);
    uint32_t n_word = n;
    CK(cudaMemcpyAsync(d_ctl + CTL_N, &n_word, 4, cudaMemcpyHostToDevice, st));
    FitWork work = fit_work_at(d_work, n);
    work.span_count = d_ctl + CTL_NSPAN;  // zero since the memset of the control block
    work.err_flag = d_ctl + CTL_ERR;
    FitJob job;
    job.n_bound = n; job.tris = b->tris;
    job.topo = b->topo; job.parent = b->parent; job.boxes = d_boxes; job.nodes2 = b->nodes2;
    job.nodes4 = b->nodes4; job.leaf_max = RC_BLAS_LEAF_MAX; job.hull = b->hull;
    job.ctl = d_ctl; job.out10 = reinterpret_cast<float *>(d_ctl + CTL_OUT);
    run_fit(st, job, work);
    RcDeviceBlas probe;
    if (!finish_blas(st, d_ctl, &probe, err)) return false;
    memcpy(b->root_aabb, probe.root_aabb, 24);
    memcpy(b->sphere, probe.sphere, 16);
    *refitted = true;
    return true;
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
    f3 smin = ctl_bounds_min(bounds), smax = ctl_bounds_max(bounds);  // bounds = the build's control block
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
    for (int k = 0; k < 4; k++) r.sphere[k] = b.sphere[k];
    rc_world_sphere(d->transform, b.sphere, r.wsphere);
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
                    (void *)t->inst_boxes, (void *)t->inst_boxes_tight, (void *)t->boxes_tight, (void *)t->leaf_map, (void *)t->topo, (void *)t->parent, (void *)t->fit_work, (void *)t->boxes, (void *)t->d_small})
        if (p) cudaFreeAsync(p, st);
    *t = RcDeviceTlas();
}

// Upload descriptors + rebuild records (used by both build and refit)
static bool upload_instances(cudaStream_t st, RcDeviceTlas *t, const rc_instance_desc *h_inst, uint32_t n, std::string &err) {
    CK(cudaMemcpyAsync(t->d_inst, h_inst, sizeof(rc_instance_desc) * n, cudaMemcpyHostToDevice, st));
    k_instance_records<<<cdiv(n, 256), 256, 0, st>>>(t->d_inst, t->d_blas_ptrs, n, t->rec, t->aux);
    return true;
}

// read back {invariant flag, root box} (adjacent words of the control block) and finish the host-side record
static bool finish_tlas(cudaStream_t st, RcDeviceTlas *t, std::string &err) {
    uint32_t h_stack[7];
    uint32_t *h = pinned_words(h_stack);
    static_assert(CTL_ERR + 1 == CTL_OUT, "one transfer");
    CK(cudaMemcpyAsync(h, t->d_small + CTL_ERR, sizeof h_stack, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    if (h[0]) { err = "internal error: fit segment table overflow"; return false; }
    memcpy(t->root_aabb, h + 1, 24);
    return extent_supported(t->root_aabb, err);
}

// The two fits of a TLAS over one topology: the reference-identical instance boxes give the BVH2 (read-backs, reference-order mode) and the
// root box; the tighter hull-derived boxes give the wide nodes.  The spanning-node list depends on the topology only: built by the first
// fit after a (re)build (its counter, t->d_small[CTL_NSPAN], is zero then) and reused by every later fit.
static void fit_tlas(cudaStream_t st, RcDeviceTlas *t, bool fresh_topology) {
    FitWork work = fit_work_at(t->fit_work, t->n);
    work.span_count = t->d_small + CTL_NSPAN;
    work.err_flag = t->d_small + CTL_ERR;
    FitJob ref;
    ref.n_bound = t->n; ref.inst_boxes = t->inst_boxes; ref.leaf_map = t->leaf_map;
    ref.topo = t->topo; ref.parent = t->parent; ref.boxes = t->boxes; ref.nodes2 = t->nodes2;
    ref.build_list = fresh_topology;
    run_fit(st, ref, work);
    FitJob tight = ref;
    tight.inst_boxes = t->inst_boxes_tight; tight.boxes = t->boxes_tight; tight.nodes2 = nullptr;
    tight.nodes4 = t->nodes4; tight.leaf_max = 1; tight.build_list = false;
    tight.out10 = reinterpret_cast<float *>(t->d_small + CTL_OUT); tight.root_boxes = t->boxes;
    run_fit(st, tight, work);
}

static TlasSmallArgs tlas_small_args(RcDeviceTlas *t, bool refit) {
    TlasSmallArgs a;
    a.inst = t->d_inst; a.blas_roots = t->d_blas_roots; a.blas = t->d_blas_ptrs; a.n = t->n; a.refit = refit;
    a.rec = t->rec; a.aux = t->aux; a.inst_boxes = t->inst_boxes; a.inst_boxes_tight = t->inst_boxes_tight;
    a.ctl = t->d_small; a.keys = a.run_keys = a.run_idx = nullptr;
    a.leaf_map = t->leaf_map; a.topo = t->topo; a.parent = t->parent;
    a.boxes = t->boxes; a.boxes_tight = t->boxes_tight; a.nodes2 = t->nodes2; a.nodes4 = t->nodes4;
    a.work_ref = fit_work_at(t->fit_work, t->n, SB_FT);
    a.work_tight = fit_work_at(t->fit_work + fit_work_bytes(t->n, SB_FT), t->n, SB_FT);
    a.work_tight.span_list = a.work_ref.span_list;  // the spanning nodes depend on the topology only
    a.work_ref.span_count = a.work_tight.span_count = t->d_small + CTL_NSPAN;
    a.work_ref.err_flag = a.work_tight.err_flag = t->d_small + CTL_ERR;
    return a;
}

bool rc_build_tlas(cudaStream_t st, const rc_instance_desc *h_inst, uint32_t n, const std::vector<RcBlasPtrs> &blas, const std::vector<float> &blas_roots,
                   RcDeviceTlas *t, std::string &err) {
    uint32_t nb = (uint32_t)blas.size();
    // a rebuild with the same instance count (every sync! after a mesh update) keeps the device arrays: 16 cudaFreeAsync + 16 cudaMallocAsync
    // were a third of a small rebuild
    const bool reuse = n > 0 && t->n == n && t->nodes4 && t->d_small && t->n_blas_cap >= nb;  // (d_small is the last array allocated: a complete set)
    if (!reuse) rc_free_tlas(t, st);
    for (int k = 0; k < 3; k++) { t->root_aabb[k] = INFINITY; t->root_aabb[3 + k] = -INFINITY; }  // Bounds3()
    t->n = n;
    if (n == 0) return true;  // empty TLAS: zero nodes (:969-978)
    const int T = 256;
    uint32_t *d_codes = nullptr, *d_idx = nullptr, *d_codes2 = nullptr, *d_idx2 = nullptr, *d_hist = nullptr;
    RcTemps tmp(st);
    const bool small = n <= small_build_limit();  // one cooperative kernel (k_tlas_small); its two fits keep a set of segment tables each
    if (!reuse) {
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
        CK(cudaMallocAsync(&t->fit_work, small ? 2 * fit_work_bytes(n, SB_FT) : fit_work_bytes(n), st));
        CK(cudaMallocAsync(&t->boxes, sizeof(RcBox) * (2 * n - 1), st));
        CK(cudaMallocAsync(&t->nodes2, sizeof(RcNode2) * (2 * n - 1), st));
        CK(cudaMallocAsync(&t->nodes4, sizeof(RcNode4) * (n + 1), st));
        CK(cudaMallocAsync(&t->d_small, sizeof(uint32_t) * CTL_WORDS, st));
        t->n_blas_cap = nb;
    }
    CK(cudaMemcpyAsync(t->d_blas_roots, blas_roots.data(), sizeof(float) * 6 * nb, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(t->d_blas_ptrs, blas.data(), sizeof(RcBlasPtrs) * nb, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(t->d_small, 0, sizeof(uint32_t) * CTL_WORDS, st));
    if (small) {
        TMP(d_codes, 3 * (size_t)n);  // codes, run codes, run indices
        CK(cudaMemcpyAsync(t->d_inst, h_inst, sizeof(rc_instance_desc) * n, cudaMemcpyHostToDevice, st));
        TlasSmallArgs sa = tlas_small_args(t, false);
        sa.keys = d_codes; sa.run_keys = d_codes + n; sa.run_idx = d_codes + 2 * (size_t)n;
        if (!launch_tlas_small(st, sa, err)) return false;
        return finish_tlas(st, t, err);
    }
    TMP(d_codes, n);
    TMP(d_idx, n);
    TMP(d_codes2, n);
    TMP(d_idx2, n);
    TMP(d_hist, radix_hist_words(n));
    if (!upload_instances(st, t, h_inst, n, err)) return false;
    k_instance_boxes<<<cdiv(n, T), T, 0, st>>>(t->d_inst, t->d_blas_roots, n, t->inst_boxes, t->d_small, t->d_blas_ptrs, t->inst_boxes_tight);
    k_morton_instances<<<cdiv(n, T), T, 0, st>>>(t->d_inst, t->d_blas_roots, n, t->d_small, d_codes, d_idx);
    FrontArgs fa;
    fa.verts = nullptr; fa.face_meta = nullptr; fa.n_faces = 0; fa.tris_in = nullptr; fa.tile_counts = nullptr; fa.ctl = t->d_small;
    fa.keys0 = d_codes; fa.vals0 = d_idx; fa.keys1 = d_codes2; fa.vals1 = d_idx2;
    fa.n_host = n; fa.hist = d_hist; fa.tiles_stride = cdiv(n, RS_TILE); fa.bar = t->d_small + CTL_BAR;
    if (!launch_front(st, fa, cdiv(n, RS_TILE), err)) return false;
    uint32_t *codes_sorted = d_codes2, *order = d_idx2;
    CK(cudaMemcpyAsync(t->leaf_map, order, sizeof(uint32_t) * n, cudaMemcpyDeviceToDevice, st));  // leaf_map = sorted position -> instance index
    launch_dependent(n >= PDL_MIN_ELEMENTS, k_topology, cdiv(std::max(1u, n - 1), T), T, 0, st, codes_sorted, nullptr, n, t->topo, t->parent);
    fit_tlas(st, t, true);
    return finish_tlas(st, t, err);
}

bool rc_refit_tlas(cudaStream_t st, const rc_instance_desc *h_inst, uint32_t n, RcDeviceTlas *t, std::string &err) {
    if (n == 0) return true;
    if (n != t->n) { err = "refit: instance count changed"; return false; }
    if (n <= small_build_limit()) {  // records, boxes and both fits in one cooperative kernel
        CK(cudaMemcpyAsync(t->d_inst, h_inst, sizeof(rc_instance_desc) * n, cudaMemcpyHostToDevice, st));
        TlasSmallArgs sa = tlas_small_args(t, true);
        if (!launch_tlas_small(st, sa, err)) return false;
        return finish_tlas(st, t, err);
    }
    if (!upload_instances(st, t, h_inst, n, err)) return false;
    // update_tlas_leaf_aabbs_kernel! (kernels.jl:487-519) + refit_tlas_aabbs_kernel! (:381-428), then re-quantise the wide nodes
    k_instance_boxes<<<cdiv(n, 256), 256, 0, st>>>(t->d_inst, t->d_blas_roots, n, t->inst_boxes, nullptr, t->d_blas_ptrs, t->inst_boxes_tight);
    fit_tlas(st, t, false);
    return finish_tlas(st, t, err);
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
    char magic[8];  // "RCBLAS\0\2"
    uint32_t abi_version, leaf_max, hull_boxes, n, n_faces_in, has_normals;
    float root_aabb[6];
    uint64_t total_bytes, payload_hash;
    uint64_t off_nodes2, off_nodes4, off_tris, off_hull, off_normals;  // from the blob start
    float sphere[4];  // bounding sphere (centre, radius^2) of the instance-entry cull
};
static_assert(sizeof(RcBlobHeader) == 128, "blob header is 128 bytes");
static const char RC_BLOB_MAGIC[8] = {'R', 'C', 'B', 'L', 'A', 'S', 0, 2};

static inline uint64_t up64(uint64_t x) { return (x + 63u) & ~(uint64_t)63u; }

static void blob_layout(uint32_t n, bool normals, bool nodes2, RcBlobHeader *h) {
    uint64_t o = sizeof(RcBlobHeader);
    h->off_nodes2 = nodes2 ? o : 0;  // the reference-layout BVH2 travels only when the geometry was built with RC_BUILD_KEEP_BVH2
    if (nodes2) o = up64(o + sizeof(RcNode2) * (2 * (uint64_t)n - 1));
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
    blob_layout(b.n, b.normals != nullptr, b.nodes2 != nullptr, &h);
    return h.total_bytes;
}

bool rc_blas_export(cudaStream_t st, const RcDeviceBlas &b, void *blob, uint64_t capacity, std::string &err) {
    if (b.n == 0 || !b.nodes4 || !b.tris || !b.hull) { err = "export: geometry is not built"; return false; }
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
    blob_layout(b.n, b.normals != nullptr, b.nodes2 != nullptr, &h);
    if (capacity < h.total_bytes) { err = "export: capacity too small"; return false; }
    uint8_t *p = static_cast<uint8_t *>(blob);
    memset(p + sizeof h, 0, h.total_bytes - sizeof h);  // alignment gaps are part of the hashed payload
    const uint64_t n = b.n;
    if (b.nodes2) CK(cudaMemcpyAsync(p + h.off_nodes2, b.nodes2, sizeof(RcNode2) * (2 * n - 1), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(p + h.off_nodes4, b.nodes4, sizeof(RcNode4) * (n + 1), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(p + h.off_tris, b.tris, sizeof(RcTri) * n, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(p + h.off_hull, b.hull, sizeof(RcBox) * RC_HULL_BOXES, cudaMemcpyDeviceToHost, st));
    if (b.normals) CK(cudaMemcpyAsync(p + h.off_normals, b.normals, sizeof(float) * 9 * n, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    // (every wide-node slot is written by the builder — unused ones are zeroed — so equal geometry gives equal blobs)
    h.payload_hash = blob_hash(p + sizeof h, h.total_bytes - sizeof h);
    memcpy(p, &h, sizeof h);
    return true;
}

// structural check of an uploaded blob (rc_validate_blas_elem, rc_build_core.cuh): every reference stays inside the arrays and no
// cycle is reachable from a root, so a damaged blob can neither send a traversal out of bounds nor make it spin
__global__ void k_validate_static(const RcNode2 *__restrict__ nodes2, const RcTri *__restrict__ tris, uint32_t n, uint32_t n_faces_in, uint32_t *__restrict__ bad) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t errs = rc_validate_static_elem(i, nodes2, tris, n, n_faces_in);
    if (errs) atomicAdd(bad, errs);
}
// one breadth-first level of the wide-node check; state[0] = violations, state[1] = nodes marked for the next level, mark[k] = level of wide node k
__global__ void k_validate_wide(uint32_t level, const RcNode4 *__restrict__ nodes4, uint32_t n, uint32_t *__restrict__ mark, uint32_t *__restrict__ state) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t errs = rc_validate_wide_level(i, level, nodes4, n, RC_BLAS_LEAF_MAX, mark, state + 1);
    if (errs) atomicAdd(state, errs);
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
    blob_layout(h.n, h.has_normals != 0, h.off_nodes2 != 0, &want);
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
    TMP(d_bad, n + 3);  // [0] violations, [1] nodes marked by the current level, [2 + k] breadth-first level of wide node k
    if (h.off_nodes2) CK(cudaMallocAsync(&out->nodes2, sizeof(RcNode2) * (2 * n - 1), st));
    CK(cudaMallocAsync(&out->nodes4, sizeof(RcNode4) * (n + 1), st));
    CK(cudaMallocAsync(&out->tris, sizeof(RcTri) * n, st));
    CK(cudaMallocAsync(&out->hull, sizeof(RcBox) * RC_HULL_BOXES, st));
    if (h.has_normals) CK(cudaMallocAsync(&out->normals, sizeof(float) * 9 * n, st));
    out->n = h.n;
    out->n_faces_in = h.n_faces_in;
    memcpy(out->root_aabb, h.root_aabb, 24);
    memcpy(out->sphere, h.sphere, 16);
    if (h.off_nodes2) CK(cudaMemcpyAsync(out->nodes2, p + h.off_nodes2, sizeof(RcNode2) * (2 * n - 1), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(out->nodes4, p + h.off_nodes4, sizeof(RcNode4) * (n + 1), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(out->tris, p + h.off_tris, sizeof(RcTri) * n, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(out->hull, p + h.off_hull, sizeof(RcBox) * RC_HULL_BOXES, cudaMemcpyHostToDevice, st));
    if (h.has_normals) CK(cudaMemcpyAsync(out->normals, p + h.off_normals, sizeof(float) * 9 * n, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(d_bad, 0, sizeof(uint32_t) * (n + 3), st));
    k_validate_static<<<cdiv((uint32_t)(2 * n), 256), 256, 0, st>>>(out->nodes2, out->tris, h.n, h.n_faces_in, d_bad);
    uint32_t one = 1, state[2] = {0, 1}, bad = 0;
    CK(cudaMemcpyAsync(d_bad + 2 + 1, &one, 4, cudaMemcpyHostToDevice, st));  // mark[root] = level 1
    for (uint32_t level = 1; state[1] != 0 && state[0] == 0; level++) {  // one launch per tree level (a few dozen): an import is not a hot path
        CK(cudaMemsetAsync(d_bad + 1, 0, 4, st));
        k_validate_wide<<<cdiv((uint32_t)n + 1, 256), 256, 0, st>>>(level, out->nodes4, h.n, d_bad + 2, d_bad);
        CK(cudaMemcpyAsync(state, d_bad, 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));  // (also: the caller may release the blob on return)
        if (level > h.n + 1) { state[0]++; break; }  // cannot happen: every node is marked at most once
    }
    bad = state[0];
    CK(cudaGetLastError());
    if (bad) { err = "import: blob fails the structural check (" + std::to_string(bad) + " bad references)"; return false; }
    return true;
}
