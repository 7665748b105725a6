// rc_trace_fast.cuh — the default traversal kernel: persistent lanes over the quantised BVH4 with a
// warp-level step scheduler.
//
// Every lane owns one ray at a time and is, at any moment, ready for one or two of four step kinds:
//   N  node step      its next reference is a wide node (4 quantised child boxes)
//   T  triangle step  it has a leaf parked (one triangle per step)
//   X  level step     it must enter an instance (TLAS leaf); leaving one (sentinel popped) is folded into the settle
//   F  refill         its ray is finished (or it has none yet)
// Each lane keeps its readiness as a packed vote word (one byte per kind); every iteration the warp sums the votes
// with ONE REDUX.SUM and executes one kind for the lanes that are ready for it (node steps by default; a short step as soon as
// it has a policy-defined share of the node-step lanes, see the scheduler constants below), so a freshly fetched ray that needs
// ten box steps to reach its first leaf never stalls 31 lanes that wait to test triangles, and vice versa (profiles/r1_v2: a plain
// while-while loop ran the box code at 9.4/32 lanes; the shipped policies run node steps at 23-25/32 lanes in the warp simulator).
// The vote word is recomputed only by the lanes that just executed a step ("settle"), which also returns a lane to the TLAS when
// it popped the level sentinel and parks a freshly reached leaf so the lane can keep descending.
// Refills are warp-cooperative (one atomic on the global work counter per refill) and deferred until RC_FETCH_MIN_*
// lanes are idle, so the refill / retire code also runs with several lanes.
//
// Arithmetic: two child planes are decoded per PRMT into a half2 of subnormals (0x00qq = q * 2^-24 exactly), widened with
// HADD2.F32 on the FMA pipe (no I2F: the XU pipe saturated in profiles/r1_v1; the ALU pipe is the limiter since v4),
// and fed to one FMA against per-node (2^24 * scale * inv_d, (origin - o) * inv_d); an explicit rounding bound keeps the
// slab test conservative.  The triangle test is the exact, FMA-free Moeller-Trumbore of
// rc_device.cuh, so t/u/v are bit-identical to the reference evaluation whenever the same triangle wins.
// The traversal stack lives in shared memory ([depth][thread], conflict-free); pushes are branch-free (a rejected
// child is written to a dummy row); rays that would need more than RC_SSTACK entries are flagged and re-traced by
// k_trace_fixup with the deep-stack generic body.
#pragma once
#ifdef RC_WARPSIM
// tests/hostsim compiles this very kernel for the CPU — one fibre per lane, the warp intrinsics as lock-step exchanges — so the
// scheduler, the stack handling and the level changes are parity-tested without a GPU.  Test infrastructure only: the shim comes
// from tests/hostsim/warpsim.h, never from the library build.
#include "warpsim.h"
#else
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math_constants.h>

#include "rc_trace.h"
#endif
#include "rc_trace_core.cuh"

// RC_SIM_STEP(kind, active): step statistics of the CPU warp simulator (iterations and active lanes per step kind); nothing on the GPU
#ifndef RC_SIM_STEP
#define RC_SIM_STEP(kind, active)
#define RC_SIM_ITER()
#define RC_SIM_IDLE(vote_word, cur_ref, leaf_ref)
#endif

__device__ __forceinline__ rc_ray rc_load_ray(const rc_ray *rays, unsigned long long i) {
    const float4 *p = reinterpret_cast<const float4 *>(rays + i);
    float4 a = __ldcs(p), b = __ldcs(p + 1);  // streamed once: evict-first keeps the BVH resident in L2
    rc_ray r;
    r.origin[0] = a.x; r.origin[1] = a.y; r.origin[2] = a.z; r.tmin = a.w;
    r.dir[0] = b.x; r.dir[1] = b.y; r.dir[2] = b.z; r.tmax = b.w;
    return r;
}

__device__ __forceinline__ void rc_store_hit(rc_hit *hits, unsigned long long i, const rc_hit &h) {
    float4 *p = reinterpret_cast<float4 *>(hits + i);
    // streaming stores: hit records are write-once, keep them out of the way of the BVH working set in L2
    __stcs(p, make_float4(__uint_as_float(h.hit), h.t, __uint_as_float(h.primitive_id), __uint_as_float(h.instance_custom_index)));
    __stcs(p + 1, make_float4(h.bary_u, h.bary_v, __uint_as_float(h.instance_id), __uint_as_float(h.metadata)));
}

// Scheduler constants per kernel variant (swept with tools/exp_variant.py on the B200 and screened with tests/sched_model.py in the
// CPU warp simulator, profiles/README.md).  A refill runs when at least RC_FETCH_MIN lanes of the warp are idle (or nothing else can
// run); a T step runs when RC_T_W * nT >= nN, an X step when RC_X_W * nX > nN.  Letting the short steps (triangle test, instance entry,
// refill) run before they have a majority keeps their lanes from idling through long runs of node steps; the weight and the refill
// threshold only pay off together (each alone is within +-1 % on C2).  One mesh under one instance (SINGLE): weight 2, refill at 8
// (C2 interior rays 2.54 -> 2.36 ms per 2^23; weight 3 / refill 6: 2.39).  With a real TLAS the rays are longer and spread over four
// step kinds: weight 3, refill at 6 (C3: 13.39 -> 10.28 ms).
#ifndef RC_FETCH_MIN_SINGLE
#define RC_FETCH_MIN_SINGLE 8
#endif
#ifndef RC_FETCH_MIN_MULTI
#define RC_FETCH_MIN_MULTI 6
#endif
#ifndef RC_T_W_SINGLE
#define RC_T_W_SINGLE 2u
#endif
#ifndef RC_T_W_MULTI
#define RC_T_W_MULTI 3u
#endif
#ifndef RC_X_W
#define RC_X_W 3u
#endif
// RC_PARK_INSTANCE: a lane that reaches an instance leaf parks it and keeps walking the TLAS (as it parks a BLAS leaf); the level step
// enters the parked instance — or drops it when the bounding-sphere test culls the entry, in which case the TLAS walk was never interrupted.
#ifndef RC_PARK_INSTANCE
#define RC_PARK_INSTANCE 0
#endif
#ifndef RC_WINV_SMEM
#define RC_WINV_SMEM 1  // C3: 9.58 -> 9.49 ms per 2^24 rays (profiles/README.md r2)
#endif
#ifndef RC_MIN_BLOCKS
#define RC_MIN_BLOCKS 8     // 64 registers -> 32 resident warps per SM (swept 6..10 in r1: 8 is best, 9+ spills)
#endif
#ifndef RC_MIN_BLOCKS_SINGLE
#define RC_MIN_BLOCKS_SINGLE 9  // the single-instance variant needs 56 registers (no world-ray copy): 36 resident warps (10 spills, -5 %)
#endif
#ifndef RC_SSTACK
#define RC_SSTACK 32       // stack entries per lane (shared memory, [depth][thread]); deeper rays go through the fix-up pass
#endif
#define RC_OVERFLOW_MARK 0xFFFFFFFFu  // rc_hit.hit of a ray whose short stack overflowed (re-traced by k_trace_fixup)
#define RC_DEADLANE 0xFFFFFFFEu

#define RC_VOTE_N 0x00000001u
#define RC_VOTE_T 0x00000100u
#define RC_VOTE_X 0x00010000u
#define RC_VOTE_F 0x01000000u

// safe_invdir's clamp (src/instanced-bvh.jl:1742-1748) with MUFU.RCP (<= 1 ulp); only the conservative box test uses it.
// rcp.approx.f32 without .ftz expands to ~8 instructions (operand rescaling for subnormal inputs / results); the clamp rules out
// subnormal inputs, so the single-instruction .ftz form gives the same bits unless a component exceeds 2^126 (subnormal result),
// which takes the full form behind one warp-rarely-taken branch.  (The three reciprocals of a level change were 14 % of the
// instanced kernel's warp instructions at 4/32 lanes, profiles/r1_trace_c3_v20.)
#ifndef RC_WARPSIM
// cold path of rc_fast_inv3, out of line: kept as a call so that the hot path carries neither its ~100 instructions nor the register
// moves of a two-sided join (they were 4.8 % of the instanced kernel's warp instructions at 5/32 lanes, profiles/r2_trace_c3_final)
static __device__ __noinline__ f3 rc_slow_inv3(float x, float y, float z) {
    f3 r;
    asm("rcp.approx.f32 %0, %1;" : "=f"(r.x) : "f"(x));
    asm("rcp.approx.f32 %0, %1;" : "=f"(r.y) : "f"(y));
    asm("rcp.approx.f32 %0, %1;" : "=f"(r.z) : "f"(z));
    return r;
}
#endif
__device__ __forceinline__ f3 rc_fast_inv3(f3 d) {
    const float ooeps = 1.0e-5f;
    const float x = fabsf(d.x) > ooeps ? d.x : copysignf(ooeps, d.x);
    const float y = fabsf(d.y) > ooeps ? d.y : copysignf(ooeps, d.y);
    const float z = fabsf(d.z) > ooeps ? d.z : copysignf(ooeps, d.z);
    f3 r;
#ifdef RC_WARPSIM
    r.x = 1.0f / x; r.y = 1.0f / y; r.z = 1.0f / z;  // host stand-in for MUFU.RCP (<= 1 ulp apart; only the conservative box test sees it)
#else
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(y));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.z) : "f"(z));
    if (__builtin_expect(fmaxf(fmaxf(fabsf(x), fabsf(y)), fabsf(z)) > 8.5070591730234616e37f, 0)) r = rc_slow_inv3(x, y, z);  // 2^126: 1/x is subnormal
#endif
    return r;
}
// bound on the relative error of the slab evaluation: reciprocal (1 ulp), two roundings of (origin - o) * inv, one rounding of
// the scale product, one FMA
#define RC_BOX_EPS_FAST 4.8e-7f  // 2^-21

// 32-byte read-only load (LDG.E.256.CONSTANT on sm_100a): a 64-B wide node is two of these instead of four LDG.128,
// halving the L1 wavefronts per node step (LSU wavefronts were 70 % of peak in profiles/r1_v6)
__device__ __forceinline__ void rc_ldg256(const void *p, float4 &a, float4 &b) {
#ifdef RC_WARPSIM
    a = static_cast<const float4 *>(p)[0];
    b = static_cast<const float4 *>(p)[1];
#else
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                 : "l"(p));
#endif
}

// bytes (2j, 2j+1) of w -> two floats q * 2^-24 (exact): each byte becomes the mantissa of a subnormal fp16 (0x00qq), which
// HADD2.F32 widens exactly; the 2^24 factor is folded into the per-node slab scale.  One PRMT with immediate selector and RZ per
// pair — no magic-constant register (the 0x6400 bias form cost an extra register move per PRMT, profiles/r1_v7).
__device__ __forceinline__ float2 rc_q2f_pair(uint32_t w, int j) {
    const uint32_t h = __byte_perm(w, 0u, j == 0 ? 0x4140u : 0x4342u);
    return __half22float2(*reinterpret_cast<const __half2 *>(&h));
}

// Two FMAs per issue slot: fma.rn.f32x2 (FFMA2 on sm_100a) with the scale and offset broadcast from scalar registers
// (SASS: FFMA2 R8, R8.F32x2.HI_LO, R0.F32, R13.F32).  The node step is issue / ALU bound (profiles/r1_trace_c3_v20: issue slots 80 %,
// FMA pipe 32 %), so halving the 24 slab FMAs' issue slots is free throughput.  Each component is an IEEE RN fma, same bits as fmaf.
// Measured (profiles/README.md r2): -1.7 % on the instanced scene, +2.4 % on the single-instance variant (its 56-register budget has
// no room for the aligned register pairs), so only the multi-instance variant uses the packed form.
#ifndef RC_FFMA2
#define RC_FFMA2 1
#endif
template <bool PACKED>
__device__ __forceinline__ float2 rc_fma2(float2 q, float a, float b) {
#if defined(RC_WARPSIM) || !RC_FFMA2
    return make_float2(fmaf(q.x, a, b), fmaf(q.y, a, b));
#else
    if (!PACKED) return make_float2(fmaf(q.x, a, b), fmaf(q.y, a, b));
    unsigned long long qq, aa, bb, r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(qq) : "f"(q.x), "f"(q.y));
    asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(qq), "l"(aa), "l"(bb));
    float2 o;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(o.x), "=f"(o.y) : "l"(r));
    return o;
#endif
}

// Instance-entry cull (level step): does the local-space ray (o, d) miss the BLAS's bounding sphere (c, r2)?  Conservative: the
// closest approach l = oc - (oc.d / d.d) d is evaluated in the cancellation-free form (Haines et al., "Precision improvements for
// ray / sphere intersection"), its rounding error is bounded by ~4 ulp(|oc|), and the comparison carries 1.1e-6 (r2 + |oc|^2) of slack
// (2 r delta + delta^2 <= 1e-6 (r2 + oc2) for delta = 1e-6 |oc|).  NaN / Inf anywhere makes every comparison false: no cull.
#ifndef RC_SPHERE_CULL
#define RC_SPHERE_CULL 1
#endif
__device__ __forceinline__ bool rc_misses_sphere(f3 o, f3 d, float4 sph, float t_max) {
    const float ocx = o.x - sph.x, ocy = o.y - sph.y, ocz = o.z - sph.z;
    const float a = fmaf(d.x, d.x, fmaf(d.y, d.y, d.z * d.z));
    const float b = fmaf(ocx, d.x, fmaf(ocy, d.y, ocz * d.z));
    const float oc2 = fmaf(ocx, ocx, fmaf(ocy, ocy, ocz * ocz));
    const float s = __fdividef(b, a);  // ray parameter of the closest approach is -s
    const float lx = fmaf(-s, d.x, ocx), ly = fmaf(-s, d.y, ocy), lz = fmaf(-s, d.z, ocz);
    const float l2 = fmaf(lx, lx, fmaf(ly, ly, lz * lz));
    const float r2s = fmaf(1.1e-6f, sph.w + oc2, sph.w);
    // the line misses the sphere, or the origin is outside and the closest approach lies behind it (b > 0: moving away)
    return l2 > r2s || (oc2 > r2s && b > 0.0f);
}
// the same test without the "behind the origin" clause (cheaper; the box test has already dealt with most of those)
__device__ __forceinline__ bool rc_line_misses_sphere(f3 o, f3 d, float4 sph) {
    const float ocx = o.x - sph.x, ocy = o.y - sph.y, ocz = o.z - sph.z;
    const float a = fmaf(d.x, d.x, fmaf(d.y, d.y, d.z * d.z));
    const float b = fmaf(ocx, d.x, fmaf(ocy, d.y, ocz * d.z));
    const float oc2 = fmaf(ocx, ocx, fmaf(ocy, ocy, ocz * ocz));
    const float s = __fdividef(b, a);
    const float lx = fmaf(-s, d.x, ocx), ly = fmaf(-s, d.y, ocy), lz = fmaf(-s, d.z, ocz);
    return fmaf(lx, lx, fmaf(ly, ly, lz * lz)) > fmaf(1.1e-6f, sph.w + oc2, sph.w);
}

// How a lane addresses its column of the shared-memory stack.  RcStk<false>: a generic pointer (the warp simulator, and the any-hit
// variants).  RcStk<true>: ONE 32-bit register holding the address in the shared window, st.shared / ld.shared with immediate row
// offsets — with the pointer form the compiler carried two copies of the top (one as an address, one for the depth compare) and bumped
// both on every push, and the ALU pipe is this kernel's limiter: closest_hit on C3 9.25 -> 8.90 ms (profiles/README.md r2).  The any-hit
// variants keep the pointer form (the address form costs them 8 bytes of spills: 6.26 -> 6.49 ms).
#ifndef RC_STACK_ADDR32
#define RC_STACK_ADDR32 1
#endif
template <bool A32>
struct RcStk {
    typedef uint32_t *ptr;
    static constexpr int ROW = RC_TRACE_THREADS;
    static __device__ __forceinline__ ptr base(uint32_t *sstack, uint32_t tid) { return sstack + tid; }
    static __device__ __forceinline__ uint32_t ld(ptr b, int off) { return b[off]; }
    static __device__ __forceinline__ void st(ptr b, int off, uint32_t v) { b[off] = v; }
};
#ifndef RC_WARPSIM
template <>
struct RcStk<true> {
    typedef uint32_t ptr;
    static constexpr int ROW = RC_TRACE_THREADS * 4;
    static __device__ __forceinline__ ptr base(uint32_t *sstack, uint32_t tid) { return (uint32_t)__cvta_generic_to_shared(sstack + tid); }
    static __device__ __forceinline__ uint32_t ld(ptr b, int off) {
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(b + (uint32_t)off));
        return v;
    }
    static __device__ __forceinline__ void st(ptr b, int off, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(b + (uint32_t)off), "r"(v)); }
};
#endif

// Ray source / hit sink of the batched entry points: RTRay array in, RTHitResult array out.
struct RcIoArrays {
    // scheduler constants of the multi-instance variant for this ray source (see above): a cheap refill (one 32-B load) can run early
    static constexpr uint32_t kFetchMinMulti = RC_FETCH_MIN_MULTI, kTWMulti = RC_T_W_MULTI, kXWMulti = RC_X_W;
    static constexpr uint32_t kFetchMinSingle = RC_FETCH_MIN_SINGLE, kTWSingle = RC_T_W_SINGLE;
    const rc_ray *rays;
    rc_hit *hits;
    bool zero_tmin;  // closest_hit4 / any_hit4 ignore ray.t_min (src/bvh4.jl:610, :700)
    // optional list of the rays whose short stack overflowed (ovf_list[0] = their number, entries from [1]): k_trace_fixup then re-traces
    // the listed rays instead of scanning every hit record for the mark (0.83 ms per 100 M rays on C3, where a handful of rays overflow)
    unsigned long long *ovf_list;
    uint32_t ovf_cap;
    __device__ __forceinline__ rc_ray load(unsigned long long i) const {
        rc_ray r = rc_load_ray(rays, i);
        if (zero_tmin) r.tmin = 0.0f;
        return r;
    }
    __device__ __forceinline__ void store(unsigned long long i, const rc_hit &h) const {
        rc_store_hit(hits, i, h);
#ifndef RC_WARPSIM
        if (h.hit == RC_OVERFLOW_MARK && ovf_list) {
            const unsigned long long slot = atomicAdd(ovf_list, 1ull);
            if (slot < ovf_cap) ovf_list[1 + slot] = i;
        }
#endif
    }
};

// RC_VOTE_LUT: the vote word from a 32-entry constant table indexed by the reference's top nibble (+16 with a leaf parked) instead of
// three range tests and four selects
#ifndef RC_VOTE_LUT
#define RC_VOTE_LUT 1  // C3: 8.89 -> 8.83 ms per 2^24 rays
#endif
#if RC_VOTE_LUT
// vote word by (leaf parked ? 16 : 0) + top nibble of the lane's next reference: 0-7 wide node, 8-B a second BLAS leaf (waits for the T step),
// C-D instance leaf (index < 2^28), E level sentinel (waits for the parked leaf; without one the settle has already left the level),
// F RC_INVALID (finished).  Second half: the single-instance variants (no level step).
#ifdef RC_WARPSIM
static const uint32_t rc_vote_lut[64] = {
#else
static __constant__ uint32_t rc_vote_lut[64] = {
#endif
#define N_ RC_VOTE_N
#define T_ RC_VOTE_T
#define X_ RC_VOTE_X
#define F_ RC_VOTE_F
    N_, N_, N_, N_, N_, N_, N_, N_, 0, 0, 0, 0, X_, X_, 0, F_,
    N_ | T_, N_ | T_, N_ | T_, N_ | T_, N_ | T_, N_ | T_, N_ | T_, N_ | T_, T_, T_, T_, T_, T_, T_, T_, T_,
    N_, N_, N_, N_, N_, N_, N_, N_, 0, 0, 0, 0, 0, 0, 0, F_,
    N_ | T_, N_ | T_, N_ | T_, N_ | T_, N_ | T_, N_ | T_, N_ | T_, N_ | T_, T_, T_, T_, T_, T_, T_, T_, T_};
#undef N_
#undef T_
#undef X_
#undef F_
#endif

// IO: where ray i comes from and where its result goes (RcIoArrays for rc_trace_*; rc_analysis.cu plugs in an on-the-fly
// view-factor ray generator + matrix accumulator, so the analysis kernels run on the same scheduler).
// WT: the watertight triangle test (RC_MODE_WATERTIGHT) instead of Moeller-Trumbore in the T step; everything else is the same kernel.
template <bool ANY, bool COUNT, class IO, bool SINGLE = false, bool WT = false>
__global__ void __launch_bounds__(RC_TRACE_THREADS, SINGLE ? RC_MIN_BLOCKS_SINGLE : RC_MIN_BLOCKS) k_trace_wide(RcScene sc, IO io, unsigned long long n,
                                                                 unsigned long long *__restrict__ work, RcCounters *__restrict__ counters,
                                                                 uint32_t *__restrict__ overflow) {
    // rows: 0 guard (always RC_INVALID), 1..RC_SSTACK live entries, +4 scratch (<= 3 pushes past the limit before the overflow check, + 1 rejected store)
    // (+3 rows with RC_WINV_SMEM: the world ray's reciprocal direction, so the return to the top level reloads it instead of recomputing it)
    __shared__ uint32_t sstack[(RC_SSTACK + 5 + (RC_WINV_SMEM && !SINGLE ? 3 : 0)) * RC_TRACE_THREADS];
    const uint32_t FULL = 0xFFFFFFFFu;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, lt_mask = (1u << lane) - 1u;
    RcLocalCounters lc = {0, 0, 0, 0, 0};
    unsigned long long traced = 0, idx = 0;
    f3 wo = mk3(0, 0, 0), wd = mk3(0, 0, 0), o = wo, d = wd, inv = wo;
    float t_min = 0.f, t_max = 0.f, hit_u = 0.f, hit_v = 0.f;
    int cur_inst = -1, best_inst = -1;
#ifdef RC_WARPSIM
    typedef RcStk<false> STK;
#else
    typedef RcStk<(RC_STACK_ADDR32 != 0) && !ANY> STK;
#endif
    const typename STK::ptr sbase = STK::base(sstack, tid);  // row 0 of this lane's column (the guard row)
    typename STK::ptr spa = sbase;                           // top of the stack
#define RC_ROW STK::ROW
#define RC_LD_OFF(base, off) STK::ld((base), (off))
#define RC_ST_OFF(base, off, v) STK::st((base), (off), (v))
    uint32_t best_prim = 0, best_meta = 0;
    const RcTri *tris = nullptr;
    const RcNode4 *nodes = sc.tlas4;
    uint32_t cur = RC_INVALID, leaf = 0, leaf_k = 0, vote = RC_VOTE_F;
#if RC_PARK_INSTANCE
    uint32_t pinst = 0;  // parked instance leaf reference (0 = none)
#endif
    bool have = false, ovf = false;
    // A TLAS with a single instance needs no top-level traversal: the refill step enters that instance directly and, with no
    // sentinel under the BLAS entries, the ray finishes when the stack bottom is popped (saves a node step and two level changes).
    // SINGLE is a compile-time variant (chosen by the launcher when n_instances == 1): the world-space ray copy, the sentinel and the
    // whole level-change step drop out of the kernel.
    constexpr bool single = SINGLE;
    constexpr uint32_t FETCH_MIN = SINGLE ? IO::kFetchMinSingle : IO::kFetchMinMulti, T_W = SINGLE ? IO::kTWSingle : IO::kTWMulti, X_W = IO::kXWMulti;

    // Branch-free conditional push: the value is always stored one row above the top and the top pointer only advances when the
    // push is accepted, so a rejected value is simply overwritten by the next push (rows above the top are don't-care).  Row 0 is
    // a guard row that always holds RC_INVALID, so the speculative read of an empty stack is harmless and neither push nor pop
    // needs a clamp.
#ifndef RC_PUSH_PRED
#define RC_PUSH_PRED 1
#endif
#if RC_PUSH_PRED  // experiment: predicated store + predicated pointer bump (one ALU-pipe instruction less per push than SEL + IADD)
#define RC_PUSH_IF(cond, v)              \
    if (cond) {                          \
        RC_ST_OFF(spa, RC_ROW, (v));     \
        spa += RC_ROW;                   \
    }
#else
#define RC_PUSH_IF(cond, v)              \
    {                                    \
        RC_ST_OFF(spa, RC_ROW, (v));     \
        spa += (cond) ? RC_ROW : 0;      \
    }
#endif
#define RC_TOP() RC_LD_OFF(spa, 0)
#define RC_DEPTH() ((uint32_t)(spa - sbase) / RC_ROW)
    // A lane whose next reference is the level sentinel (and has no leaf parked) returns to the TLAS right here in the settle
    // instead of voting for a level-change step: the leave is ~10 instructions, a scheduler round for it costs more (+3.2 % on the
    // instanced scene C3; folding the much longer instance *entry* in the same way gives the gain back — profiles/README.md).
#define RC_SETTLE_LEAVE()                                                                 \
    if (!SINGLE && cur == RC_SENTINEL && leaf == 0) {                                     \
        cur_inst = -1; /* src/instanced-bvh.jl:1996-2006 */                               \
        nodes = sc.tlas4;                                                                 \
        cur = RC_TOP();                                                                   \
        spa -= RC_ROW;                                                                    \
        o = wo; d = wd;                                                                   \
        if (RC_WINV_SMEM) {                                                               \
            inv.x = __uint_as_float(RC_LD_OFF(sbase, (RC_SSTACK + 5) * RC_ROW));                     \
            inv.y = __uint_as_float(RC_LD_OFF(sbase, (RC_SSTACK + 6) * RC_ROW));                     \
            inv.z = __uint_as_float(RC_LD_OFF(sbase, (RC_SSTACK + 7) * RC_ROW));                     \
        } else {                                                                          \
            inv = rc_fast_inv3(d);                                                        \
        }                                                                                 \
    }
    // RC_WORLD_CULL (experiment, off): a lane that has just reached an instance leaf tests the ray against the instance's WORLD-space
    // bounding sphere right here in the settle (one 16-byte load, ~20 instructions, no transform).  On the instanced scene C3 three of four
    // instance leaves are culled that way, level steps drop from 16.5 to 6.7 per 32 rays and the CPU model predicts -6.9 % — but the test
    // then runs, a few lanes wide, in the settle of almost every top-level node step, and the B200 measures +7 % (9.57 -> 10.24 ms per 2^24
    // rays, results CRC-identical; profiles/README.md r2).  The level step's own local-space sphere test stays.
#ifndef RC_WORLD_CULL
#define RC_WORLD_CULL 0
#endif
#ifndef RC_WCULL_LINE_ONLY
#define RC_WCULL_LINE_ONLY 0
#endif
#define RC_SETTLE_WCULL()                                                                                          \
    if (!SINGLE && RC_WORLD_CULL && cur_inst < 0 && (cur + 0x40000000u) < 0x2FFFFFFFu) {                           \
        const float4 ws_ = __ldg(reinterpret_cast<const float4 *>(reinterpret_cast<const char *>(sc.inst + (cur & RC_LEAF_START_MASK)) + 80)); \
        if (RC_WCULL_LINE_ONLY ? rc_line_misses_sphere(o, d, ws_) : rc_misses_sphere(o, d, ws_, t_max)) {         \
            cur = RC_TOP();                                                                                        \
            spa -= RC_ROW;                                                                                         \
        }                                                                                                          \
    }
// the short-stack overflow path never runs on real scenes: keep it a branch (an empty volatile asm stops the if-conversion that turned it
// into four selects on every settle; the ALU pipe is the limiter)
#ifndef RC_OVF_BRANCH
#define RC_OVF_BRANCH 1
#endif
#if RC_OVF_BRANCH && !defined(RC_WARPSIM)
#define RC_COLD_PATH() asm volatile("");
#else
#define RC_COLD_PATH()
#endif
#ifndef RC_PARK_PRED
#define RC_PARK_PRED 0
#endif
#ifndef RC_POP_PRED
#define RC_POP_PRED 1  // C3: 8.754 -> 8.717 ms per 2^24 rays
#endif
#if RC_PARK_PRED  // experiment: the park as a predicated block instead of four selects
#define RC_SETTLE_PARK()                                                                                           \
        if (park_) { leaf = cur; leaf_k = 0u; cur = RC_TOP(); spa -= RC_ROW; }
#else
#define RC_SETTLE_PARK()                                                                                           \
        const uint32_t top_ = RC_TOP();                                                                            \
        leaf = park_ ? cur : leaf;                                                                                 \
        leaf_k = park_ ? 0u : leaf_k;                                                                              \
        cur = park_ ? top_ : cur;                                                                                  \
        spa -= park_ ? RC_ROW : 0;
#endif
    // after a step: park a freshly reached BLAS leaf (so the lane can keep descending) and recompute the lane's vote
#define RC_SETTLE()                                                                                                \
    {                                                                                                              \
        RC_SETTLE_LEAVE()                                                                                          \
        if (spa > sbase + RC_SSTACK * RC_ROW) { RC_COLD_PATH() ovf = true; cur = RC_INVALID; leaf = 0; spa = sbase; RC_CLEAR_PINST() } \
        RC_SETTLE_WCULL()                                                                                          \
        const bool park_ = ((cur ^ RC_LEAF_BIT) < 0x40000000u) && leaf == 0; /* BLAS leaf reference */              \
        RC_SETTLE_PARK()                                                                                           \
        RC_SETTLE_VOTE()                                                                                           \
    }
    /* instance leaf = [0xC0000000, RC_SENTINEL); a sentinel still here waits for the parked leaf and is left by the T step's settle */
#if RC_PARK_INSTANCE
#define RC_CLEAR_PINST() pinst = 0;
#define RC_SETTLE_VOTE()                                                                                           \
        if (!SINGLE && cur_inst < 0 && pinst == 0 && (cur + 0x40000000u) < 0x2FFFFFFFu) {                          \
            pinst = cur; cur = RC_TOP(); spa -= RC_ROW;                                                            \
        }                                                                                                          \
        vote = ((int)cur >= 0) ? RC_VOTE_N : 0u;                                                                   \
        vote |= leaf ? RC_VOTE_T : ((cur == RC_INVALID && (SINGLE || pinst == 0)) ? RC_VOTE_F : 0u);               \
        if (!SINGLE) vote |= (pinst != 0) ? RC_VOTE_X : 0u;
#else
#define RC_CLEAR_PINST()
#if RC_VOTE_LUT
#define RC_SETTLE_VOTE() vote = rc_vote_lut[(SINGLE ? 32u : 0u) + (cur >> 28) + (leaf ? 16u : 0u)];
#else
#define RC_SETTLE_VOTE()                                                                                           \
        vote = ((int)cur >= 0) ? RC_VOTE_N : 0u;                                                                   \
        vote |= leaf ? RC_VOTE_T : ((cur == RC_INVALID) ? RC_VOTE_F : 0u);                                         \
        if (!SINGLE) vote |= ((cur + 0x40000000u) < 0x2FFFFFFFu) ? RC_VOTE_X : 0u;
#endif
#endif

    // enter instance `index`: its world->local transform applied with the reference's exact arithmetic (:1961-1977).  `entered_` tells
    // whether the lane really went in: with CULL, a ray whose local copy misses the BLAS's bounding sphere stays at the top level
    // (o / d / inv / nodes untouched).
#define RC_ENTER_INSTANCE(index, CULL, entered_)                                          \
    {                                                                                     \
        const int inst_ = (index);                                                        \
        const char *ip_ = reinterpret_cast<const char *>(sc.inst + inst_);                \
        float m_[12];                                                                     \
        float4 q0_, q1_, q2_, q3_;                                                        \
        rc_ldg256(ip_, q0_, q1_);                                                         \
        rc_ldg256(ip_ + 32, q2_, q3_);                                                    \
        m_[0] = q0_.x; m_[1] = q0_.y; m_[2] = q0_.z; m_[3] = q0_.w;                       \
        m_[4] = q1_.x; m_[5] = q1_.y; m_[6] = q1_.z; m_[7] = q1_.w;                       \
        m_[8] = q2_.x; m_[9] = q2_.y; m_[10] = q2_.z; m_[11] = q2_.w;                     \
        ulonglong2 pp_;                                                                   \
        pp_.x = ((unsigned long long)__float_as_uint(q3_.y) << 32) | __float_as_uint(q3_.x); \
        pp_.y = ((unsigned long long)__float_as_uint(q3_.w) << 32) | __float_as_uint(q3_.z); \
        const f3 lo_ = x_transform_point(m_, wo);                                         \
        const f3 ld_ = x_transform_direction(m_, wd);                                     \
        entered_ = true;                                                                  \
        if (CULL && RC_SPHERE_CULL) {                                                     \
            const float4 sph_ = __ldg(reinterpret_cast<const float4 *>(ip_ + 64));        \
            entered_ = !rc_misses_sphere(lo_, ld_, sph_, t_max);                          \
        }                                                                                 \
        if (entered_) {                                                                   \
            cur_inst = inst_;                                                             \
            nodes = reinterpret_cast<const RcNode4 *>(pp_.x);                             \
            tris = reinterpret_cast<const RcTri *>(pp_.y);                                \
            o = lo_; d = ld_;                                                             \
            inv = rc_fast_inv3(d);                                                        \
        }                                                                                 \
    }

    for (;;) {
        const uint32_t votes = __reduce_add_sync(FULL, vote);
        RC_SIM_ITER()
        if (votes == 0) break;  // every lane is dead
#ifdef RC_WARPSIM
        const uint32_t nN = votes & 0xFFu, nT = (votes >> 8) & 0xFFu, nX = (votes >> 16) & 0xFFu, nF = votes >> 24;
#else  // one PRMT per count (a shift + a mask each otherwise: the scheduler header runs 32 lanes wide in every iteration)
        const uint32_t nN = __byte_perm(votes, 0u, 0x4440u), nT = __byte_perm(votes, 0u, 0x4441u), nX = __byte_perm(votes, 0u, 0x4442u), nF = votes >> 24;
#endif

        if (nF > 0 && (nF >= FETCH_MIN || (votes & 0x00FFFFFFu) == 0)) {
            // ---- F: retire + refill (warp-cooperative) -----------------------------------------------------------------
            const bool wantF = vote & RC_VOTE_F;
            RC_SIM_STEP(3, wantF)
            if (wantF && have) {
                rc_hit h;
                if (best_inst >= 0) {
                    h.hit = 1; h.t = t_max; h.primitive_id = best_prim; h.instance_custom_index = sc.aux[best_inst].custom_index;
                    h.bary_u = hit_u; h.bary_v = hit_v; h.instance_id = (uint32_t)best_inst; h.metadata = best_meta;
                } else {
                    rc_write_miss(h);
                }
                if (ovf) { rc_write_miss(h); h.hit = RC_OVERFLOW_MARK; atomicAdd(overflow, 1u); }
                io.store(idx, h);
                traced++;
                have = false;
            }
            const uint32_t mNeed = __ballot_sync(FULL, wantF);
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(work, (unsigned long long)__popc(mNeed));
            base = __shfl_sync(FULL, base, 0);
            if (wantF) {
                idx = base + (unsigned long long)__popc(mNeed & lt_mask);
                if (idx >= n) {
                    cur = RC_DEADLANE;
                    vote = 0;
                } else {
                    rc_ray r = io.load(idx);
                    RcRayIn w = rc_prepare_ray(r, ANY);
                    wo = w.o; wd = w.d;
                    t_min = w.t_min; t_max = w.t_max;
                    best_inst = -1; ovf = false;
                    RC_ST_OFF(sbase, 0, RC_INVALID);       // guard row
                    RC_ST_OFF(sbase, RC_ROW, RC_INVALID);  // stack bottom: popping it ends the ray
                    spa = sbase + RC_ROW;
                    if (single) {  // straight into the only instance: no top-level node step, no sentinel, no return step
                        bool in_;
                        RC_ENTER_INSTANCE(0, false, in_)
                        (void)in_;
                        if (COUNT) lc.inst_entries++;
                    } else {
                        o = wo; d = wd;
                        inv = rc_fast_inv3(d);
                        if (RC_WINV_SMEM) {
                            RC_ST_OFF(sbase, (RC_SSTACK + 5) * RC_ROW, __float_as_uint(inv.x));
                            RC_ST_OFF(sbase, (RC_SSTACK + 6) * RC_ROW, __float_as_uint(inv.y));
                            RC_ST_OFF(sbase, (RC_SSTACK + 7) * RC_ROW, __float_as_uint(inv.z));
                        }
                        cur_inst = -1;
                        nodes = sc.tlas4;
                    }
                    cur = 1;
                    leaf = 0;
#if RC_PARK_INSTANCE
                    pinst = 0;
#endif
                    vote = RC_VOTE_N;
                    have = true;
                }
            }
        } else if (nT * T_W >= nN && nT >= nX) {
            // ---- T: one triangle of the parked leaf per lane -------------------------------------------------------------
            RC_SIM_STEP(1, vote & RC_VOTE_T)
            if (vote & RC_VOTE_T) {
                const uint32_t start = leaf & RC_LEAF_START_MASK, count = ((leaf >> RC_LEAF_COUNT_SHIFT) & 7u) + 1u;
                const float4 *tp = reinterpret_cast<const float4 *>(tris + start + leaf_k);
                const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
                float t, u, v;
                if (COUNT) lc.tri_tests++;
                if ((WT ? x_intersect_triangle_watertight(o, d, mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), mk3(c.x, c.y, c.z), t_min, t_max, t, u, v)
                        : x_intersect_triangle(o, d, mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), mk3(c.x, c.y, c.z), t_min, t_max, t, u, v)) && t == t) {
                    t_max = t;
                    best_inst = cur_inst;
                    best_prim = __float_as_uint(a.w);
                    best_meta = __float_as_uint(b.w);
                    hit_u = u; hit_v = v;
                    if (ANY) { cur = RC_INVALID; spa = sbase; leaf_k = count; }
                }
                if (++leaf_k >= count) {
                    leaf = 0;
                    RC_SETTLE()  // the vote can only change when the parked leaf is exhausted
                }
            }
        } else if (!SINGLE && nX * X_W > nN) {
            // ---- X: enter an instance (TLAS leaf) --------------------------------------------------------------------------
            RC_SIM_STEP(2, vote & RC_VOTE_X)
            if (vote & RC_VOTE_X) {  // (the return to the TLAS happens in RC_SETTLE_LEAVE)
                bool in_;
#if RC_PARK_INSTANCE
                RC_ENTER_INSTANCE((int)(pinst & RC_LEAF_START_MASK), true, in_)
                pinst = 0;
                if (COUNT) { lc.inst_entries += in_ ? 1u : 0u; }
                // entered: the interrupted TLAS reference and the sentinel go under the BLAS root; culled: the TLAS walk simply goes on
                RC_PUSH_IF(in_, cur)
                RC_PUSH_IF(in_, RC_SENTINEL)
                if (COUNT && RC_DEPTH() > lc.max_stack) lc.max_stack = RC_DEPTH();
                cur = in_ ? 1u : cur;
#else
                RC_ENTER_INSTANCE((int)(cur & RC_LEAF_START_MASK), true, in_)
                if (COUNT) { lc.inst_entries += in_ ? 1u : 0u; }
                // entered: the sentinel goes under the BLAS root; culled: the next TLAS reference is popped instead
                const uint32_t top_x = RC_TOP();
                RC_PUSH_IF(in_, RC_SENTINEL)
                if (COUNT && RC_DEPTH() > lc.max_stack) lc.max_stack = RC_DEPTH();
                cur = in_ ? 1u : top_x;
                spa -= in_ ? 0 : RC_ROW;
#endif
                RC_SETTLE()
            }
        } else {
            // ---- N: test the 4 quantised child boxes, descend into the nearest, push the other hit children -----------------
            RC_SIM_STEP(0, vote & RC_VOTE_N)
            RC_SIM_IDLE(vote, cur, leaf)
            if (vote & RC_VOTE_N) {
                const char *np = reinterpret_cast<const char *>(nodes + cur);
                float4 n0, n1, n2, n3;
                rc_ldg256(np, n0, n1);
                rc_ldg256(np + 32, n2, n3);
                if (COUNT) { lc.nodes++; lc.box_tests += 4; }
                // a = (2^24 * scale) * inv_d (decoded planes carry 2^-24; the node stores the scaled value), b = (origin - o) * inv_d
                const float ax = n0.w * inv.x, ay = n3.z * inv.y, az = n3.w * inv.z;
                const float bx = (n0.x - o.x) * inv.x, by = (n0.y - o.y) * inv.y, bz = (n0.z - o.z) * inv.z;
                const float kq = 255.0f / 16777216.0f;
                const float slack = RC_BOX_EPS_FAST * fmaxf(fmaxf(fmaf(kq, fabsf(ax), fabsf(bx)), fmaf(kq, fabsf(ay), fabsf(by))), fmaf(kq, fabsf(az), fabsf(bz)));
                const uint32_t qlox = __float_as_uint(n1.x), qloy = __float_as_uint(n1.y), qloz = __float_as_uint(n1.z), qhix = __float_as_uint(n1.w);
                const uint32_t qhiy = __float_as_uint(n2.x), qhiz = __float_as_uint(n2.y);
                const uint32_t r0 = __float_as_uint(n2.z), r1 = __float_as_uint(n2.w), r2 = __float_as_uint(n3.x), r3 = __float_as_uint(n3.y);
                const uint32_t nx = inv.x >= 0.0f ? qlox : qhix, fx = inv.x >= 0.0f ? qhix : qlox;
                const uint32_t ny = inv.y >= 0.0f ? qloy : qhiy, fy = inv.y >= 0.0f ? qhiy : qloy;
                const uint32_t nz = inv.z >= 0.0f ? qloz : qhiz, fz = inv.z >= 0.0f ? qhiz : qloz;
                const float t_hi = t_max + slack;
                float tn[4];
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const float2 tnx = rc_fma2<!SINGLE>(rc_q2f_pair(nx, j), ax, bx), tny = rc_fma2<!SINGLE>(rc_q2f_pair(ny, j), ay, by), tnz = rc_fma2<!SINGLE>(rc_q2f_pair(nz, j), az, bz);
                    const float2 tfx = rc_fma2<!SINGLE>(rc_q2f_pair(fx, j), ax, bx), tfy = rc_fma2<!SINGLE>(rc_q2f_pair(fy, j), ay, by), tfz = rc_fma2<!SINGLE>(rc_q2f_pair(fz, j), az, bz);
                    const float lo0 = fmaxf(fmaxf(tnx.x, tny.x), fmaxf(tnz.x, t_min));
                    const float hi0 = fminf(fminf(tfx.x, tfy.x), tfz.x);
                    const float lo1 = fmaxf(fmaxf(tnx.y, tny.y), fmaxf(tnz.y, t_min));
                    const float hi1 = fminf(fminf(tfx.y, tfy.y), tfz.y);
                    tn[2 * j] = (lo0 <= fminf(hi0 + slack, t_hi)) ? lo0 : CUDART_INF_F;
                    tn[2 * j + 1] = (lo1 <= fminf(hi1 + slack, t_hi)) ? lo1 : CUDART_INF_F;
                }
                // unused slots carry an inverted box and child 0's reference: no validity test needed (rc_types.h)
                const float t0 = tn[0], t1 = tn[1], t2 = tn[2], t3 = tn[3];
                // the nearest hit child is entered next (exact argmin); the other hit children are pushed in slot order
                const float tm = fminf(fminf(t0, t1), fminf(t2, t3));
                const bool any_hit = tm < CUDART_INF_F;
                const bool e0 = t0 == tm, e1 = !e0 && t1 == tm, e2 = !e0 && !e1 && t2 == tm, e3 = !e0 && !e1 && !e2;
                RC_PUSH_IF(t3 < CUDART_INF_F && !e3, r3)
                RC_PUSH_IF(t2 < CUDART_INF_F && !e2, r2)
                RC_PUSH_IF(t1 < CUDART_INF_F && !e1, r1)
                RC_PUSH_IF(t0 < CUDART_INF_F && !e0, r0)
                if (COUNT && RC_DEPTH() > lc.max_stack) lc.max_stack = RC_DEPTH();
                const uint32_t rn = e0 ? r0 : (e1 ? r1 : (e2 ? r2 : r3));
#if RC_POP_PRED  // experiment: load the stack top only when no child was hit (the unconditional form reads it in every node step)
                cur = rn;
                if (!any_hit) { cur = RC_TOP(); spa -= RC_ROW; }
#else
                const uint32_t top = RC_TOP();
                cur = any_hit ? rn : top;
                spa -= any_hit ? 0 : RC_ROW;
#endif
                RC_SETTLE()
            }
        }
    }
#undef RC_PUSH_IF
#undef RC_TOP
#undef RC_DEPTH
#undef RC_ROW
#undef RC_LD_OFF
#undef RC_ST_OFF
#undef RC_SETTLE
#undef RC_SETTLE_VOTE
#undef RC_SETTLE_PARK
#undef RC_CLEAR_PINST
#undef RC_SETTLE_LEAVE
#undef RC_SETTLE_WCULL
#undef RC_ENTER_INSTANCE
    if (COUNT) {
        atomicAdd(&counters->rays, traced);
        atomicAdd(&counters->nodes, (unsigned long long)lc.nodes);
        atomicAdd(&counters->box_tests, (unsigned long long)lc.box_tests);
        atomicAdd(&counters->tri_tests, (unsigned long long)lc.tri_tests);
        atomicAdd(&counters->inst_entries, (unsigned long long)lc.inst_entries);
        atomicMax(&counters->max_stack, (unsigned long long)lc.max_stack);
    }
}
