// rc_trace_fast.cuh — the default traversal kernel: persistent lanes over the quantised BVH4 with a
// warp-level step scheduler.
//
// Every lane owns one ray at a time and is, at any moment, ready for one or two of four step kinds:
//   N  node step      its next reference is a wide node (4 quantised child boxes)
//   T  triangle step  it has a leaf parked (one triangle per step)
//   X  level step     it must enter an instance (TLAS leaf) or leave one (sentinel popped)
//   F  refill         its ray is finished (or it has none yet)
// Each iteration the warp counts the ready lanes per kind with ONE packed REDUX.SUM vote and executes the kind with the
// most ready lanes, so a freshly fetched ray that needs ten box steps to reach its first leaf never stalls 31 lanes that
// wait to test triangles, and vice versa (profiles/r1_v2: a plain while-while loop ran the box code at 9.4/32 lanes,
// this scheduler at 18.6/32 in r1_v3).  Refills are warp-cooperative (one atomic on the global work counter per refill)
// and deferred until RC_FETCH_MIN lanes are idle, so the refill / retire code also runs with many lanes.
//
// Arithmetic: child planes are decoded with PRMT + FADD (no I2F: the XU pipe saturated in profiles/r1_v1); the slab
// test is 24 FMAs against per-node (scale * inv_d, (origin - o) * inv_d) plus an explicit rounding bound (conservative);
// the triangle test is the exact, FMA-free Moeller-Trumbore of rc_device.cuh, so t/u/v are bit-identical to the
// reference evaluation whenever the same triangle wins.  The traversal stack lives in shared memory ([depth][thread],
// conflict-free) with a local-memory overflow area, so pushes / pops never touch the L1 tag stage.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#include "rc_trace.h"
#include "rc_trace_core.cuh"

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

#define RC_FETCH_MIN 12    // refill when at least this many lanes of the warp are idle (or nothing else can run)
#define RC_SSTACK 32       // stack entries per lane (shared memory, [depth][thread])
#define RC_OVERFLOW_MARK 0xFFFFFFFFu  // rc_hit.hit of a ray whose short stack overflowed (re-traced by k_trace_fixup)
#define RC_DEADLANE 0xFFFFFFFDu

#define RC_CE(ta, ra, tb, rb)                    \
    {                                            \
        bool sw_ = (tb) < (ta);                  \
        float tl_ = sw_ ? (tb) : (ta);           \
        float th_ = sw_ ? (ta) : (tb);           \
        uint32_t rl_ = sw_ ? (rb) : (ra);        \
        uint32_t rh_ = sw_ ? (ra) : (rb);        \
        ta = tl_; tb = th_; ra = rl_; rb = rh_;  \
    }

// safe_invdir's clamp (src/instanced-bvh.jl:1742-1748) with MUFU.RCP (<= 1 ulp); only the conservative box test uses it,
// and RC_BOX_EPS_FAST covers the extra ulp.
__device__ __forceinline__ float rc_fast_inv(float d) {
    const float ooeps = 1.0e-5f;
    float x = fabsf(d) > ooeps ? d : copysignf(ooeps, d), r;
    asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
#define RC_BOX_EPS_FAST 4.8e-7f  // 2^-21

template <bool ANY, bool COUNT>
__global__ void __launch_bounds__(RC_TRACE_THREADS) k_trace_wide(RcScene sc, const rc_ray *__restrict__ rays, rc_hit *__restrict__ hits, unsigned long long n,
                                                                 unsigned long long *__restrict__ work, RcCounters *__restrict__ counters,
                                                                 uint32_t *__restrict__ overflow) {
    __shared__ uint32_t sstack[RC_SSTACK * RC_TRACE_THREADS];
    const uint32_t FULL = 0xFFFFFFFFu;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, lt_mask = (1u << lane) - 1u;
    RcLocalCounters lc = {0, 0, 0, 0, 0};
    unsigned long long traced = 0, idx = 0;
    f3 wo = mk3(0, 0, 0), wd = mk3(0, 0, 0), o = wo, d = wd, inv = wo;
    float t_min = 0.f, t_max = 0.f, hit_u = 0.f, hit_v = 0.f;
    int cur_inst = -1, best_inst = -1, sp = 0;
    uint32_t best_prim = 0, best_meta = 0;
    const RcTri *tris = nullptr;
    const RcNode4 *nodes = sc.tlas4;
    uint32_t cur = RC_INVALID, leaf = 0, leaf_k = 0;
    bool have = false, ovf = false;

    // The stack holds RC_SSTACK entries per lane.  A push beyond that is dropped and flags the ray; flagged rays are
    // re-traced by k_trace_fixup with the deep-stack generic body, so results never depend on the short stack.
#define RC_PUSH(v)                                                     \
    {                                                                  \
        if (sp < RC_SSTACK) sstack[sp * RC_TRACE_THREADS + tid] = (v); \
        ovf |= sp >= RC_SSTACK;                                        \
        sp++;                                                          \
    }
#define RC_TOP() (sstack[min(max(sp - 1, 0), RC_SSTACK - 1) * RC_TRACE_THREADS + tid])

    for (;;) {
        // park a leaf (cheap, every iteration): a BLAS leaf reference with no leaf parked yet
        {
            const bool park = cur_inst >= 0 && (cur & RC_LEAF_BIT) && cur < RC_DEADLANE && leaf == 0;
            const uint32_t top = RC_TOP();
            leaf = park ? cur : leaf;
            leaf_k = park ? 0u : leaf_k;
            cur = park ? top : cur;
            sp -= park ? 1 : 0;
        }
        const bool wantN = !(cur & RC_LEAF_BIT);
        const bool wantT = leaf != 0;
        const bool wantX = (cur_inst < 0 && (cur & RC_LEAF_BIT) && cur < RC_DEADLANE) || (cur == RC_SENTINEL && leaf == 0);
        const bool wantF = cur == RC_INVALID && leaf == 0;
        const uint32_t votes = __reduce_add_sync(FULL, (uint32_t)wantN | ((uint32_t)wantT << 8) | ((uint32_t)wantX << 16) | ((uint32_t)wantF << 24));
        if (votes == 0) break;  // every lane is dead
        const uint32_t nN = votes & 0xFFu, nT = (votes >> 8) & 0xFFu, nX = (votes >> 16) & 0xFFu, nF = votes >> 24;

        if (nF > 0 && (nF >= RC_FETCH_MIN || (votes & 0x00FFFFFFu) == 0)) {
            // ---- F: retire + refill (warp-cooperative) -----------------------------------------------------------------
            if (wantF && have) {
                rc_hit h;
                if (best_inst >= 0) {
                    h.hit = 1; h.t = t_max; h.primitive_id = best_prim; h.instance_custom_index = sc.aux[best_inst].custom_index;
                    h.bary_u = hit_u; h.bary_v = hit_v; h.instance_id = (uint32_t)best_inst; h.metadata = best_meta;
                } else {
                    rc_write_miss(h);
                }
                if (ovf) { rc_write_miss(h); h.hit = RC_OVERFLOW_MARK; atomicAdd(overflow, 1u); }
                rc_store_hit(hits, idx, h);
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
                } else {
                    rc_ray r = rc_load_ray(rays, idx);
                    RcRayIn w = rc_prepare_ray(r, ANY);
                    wo = w.o; wd = w.d; o = wo; d = wd;
                    t_min = w.t_min; t_max = w.t_max;
                    inv = mk3(rc_fast_inv(d.x), rc_fast_inv(d.y), rc_fast_inv(d.z));
                    cur_inst = -1; best_inst = -1; ovf = false;
                    nodes = sc.tlas4;
                    sstack[tid] = RC_INVALID;
                    sp = 1;
                    cur = 1;
                    have = true;
                }
            }
        } else if (nT >= nN && nT >= nX) {
            // ---- T: one triangle of the parked leaf per lane -------------------------------------------------------------
            if (wantT) {
                const uint32_t start = leaf & RC_LEAF_START_MASK, count = ((leaf >> RC_LEAF_COUNT_SHIFT) & 7u) + 1u;
                const float4 *tp = reinterpret_cast<const float4 *>(tris + start + leaf_k);
                const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
                float t, u, v;
                if (COUNT) lc.tri_tests++;
                if (x_intersect_triangle(o, d, mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), mk3(c.x, c.y, c.z), t_min, t_max, t, u, v) && t == t) {
                    t_max = t;
                    best_inst = cur_inst;
                    best_prim = __float_as_uint(a.w);
                    best_meta = __float_as_uint(b.w);
                    hit_u = u; hit_v = v;
                    if (ANY) { cur = RC_INVALID; sp = 0; leaf_k = count; }
                }
                if (++leaf_k >= count) leaf = 0;
            }
        } else if (nX > nN) {
            // ---- X: enter an instance (TLAS leaf) or return to the TLAS (sentinel) -------------------------------------------
            if (wantX) {
                if (cur == RC_SENTINEL) {
                    cur_inst = -1;  // src/instanced-bvh.jl:1996-2006
                    nodes = sc.tlas4;
                    cur = RC_TOP();
                    sp--;
                    if (cur != RC_INVALID) {  // more TLAS work: restore the world ray (skipped when the ray is finished)
                        o = wo; d = wd;
                        inv = mk3(rc_fast_inv(d.x), rc_fast_inv(d.y), rc_fast_inv(d.z));
                    }
                } else {
                    cur_inst = (int)(cur & RC_LEAF_START_MASK);  // :1961-1977; ray transformed with the reference's exact arithmetic
                    const char *ip = reinterpret_cast<const char *>(sc.inst + cur_inst);
                    float m[12];
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        float4 r = __ldg(reinterpret_cast<const float4 *>(ip) + k);
                        m[4 * k] = r.x; m[4 * k + 1] = r.y; m[4 * k + 2] = r.z; m[4 * k + 3] = r.w;
                    }
                    const ulonglong2 pp = __ldg(reinterpret_cast<const ulonglong2 *>(ip + 48));
                    nodes = reinterpret_cast<const RcNode4 *>(pp.x);
                    tris = reinterpret_cast<const RcTri *>(pp.y);
                    o = x_transform_point(m, wo);
                    d = x_transform_direction(m, wd);
                    inv = mk3(rc_fast_inv(d.x), rc_fast_inv(d.y), rc_fast_inv(d.z));
                    RC_PUSH(RC_SENTINEL)
                    if (COUNT) { lc.inst_entries++; if ((uint32_t)sp > lc.max_stack) lc.max_stack = (uint32_t)sp; }
                    cur = 1;
                    if (ovf) { cur = RC_INVALID; leaf = 0; sp = 0; }
                }
            }
        } else {
            // ---- N: test the 4 quantised child boxes, descend into the nearest, push the rest far -> near -----------------
            if (wantN) {
                const float4 *np = reinterpret_cast<const float4 *>(nodes + cur);
                const float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3);
                if (COUNT) { lc.nodes++; lc.box_tests += 4; }
                const uint32_t e = __float_as_uint(n0.w);
                const float ax = __uint_as_float((e & 0xFFu) << 23) * inv.x, ay = __uint_as_float(((e >> 8) & 0xFFu) << 23) * inv.y,
                            az = __uint_as_float(((e >> 16) & 0xFFu) << 23) * inv.z;
                const float bx = (n0.x - o.x) * inv.x, by = (n0.y - o.y) * inv.y, bz = (n0.z - o.z) * inv.z;
                const float slack = RC_BOX_EPS_FAST * fmaxf(fmaxf(fmaf(255.0f, fabsf(ax), fabsf(bx)), fmaf(255.0f, fabsf(ay), fabsf(by))), fmaf(255.0f, fabsf(az), fabsf(bz)));
                const uint32_t qlox = __float_as_uint(n1.x), qloy = __float_as_uint(n1.y), qloz = __float_as_uint(n1.z), qhix = __float_as_uint(n1.w);
                const uint32_t qhiy = __float_as_uint(n2.x), qhiz = __float_as_uint(n2.y);
                uint32_t r0 = __float_as_uint(n2.z), r1 = __float_as_uint(n2.w), r2 = __float_as_uint(n3.x), r3 = __float_as_uint(n3.y);
                const uint32_t nx = inv.x >= 0.0f ? qlox : qhix, fx = inv.x >= 0.0f ? qhix : qlox;
                const uint32_t ny = inv.y >= 0.0f ? qloy : qhiy, fy = inv.y >= 0.0f ? qhiy : qloy;
                const uint32_t nz = inv.z >= 0.0f ? qloz : qhiz, fz = inv.z >= 0.0f ? qhiz : qloz;
                const float t_hi = t_max + slack;
                float tn[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    float lo = fmaxf(fmaxf(fmaf(rc_q2f(nx, k), ax, bx), fmaf(rc_q2f(ny, k), ay, by)), fmaxf(fmaf(rc_q2f(nz, k), az, bz), t_min));
                    float hi = fminf(fminf(fmaf(rc_q2f(fx, k), ax, bx), fmaf(rc_q2f(fy, k), ay, by)), fmaf(rc_q2f(fz, k), az, bz));
                    tn[k] = (lo <= fminf(hi + slack, t_hi)) ? lo : CUDART_INF_F;
                }
                // empty slots carry an inverted box (qlo = 255, qhi = 0) and could only pass through the slack: mask them
                float t0 = r0 == RC_INVALID ? CUDART_INF_F : tn[0], t1 = r1 == RC_INVALID ? CUDART_INF_F : tn[1];
                float t2 = r2 == RC_INVALID ? CUDART_INF_F : tn[2], t3 = r3 == RC_INVALID ? CUDART_INF_F : tn[3];
                // the nearest hit child is entered next (exact argmin); the other hit children are pushed in slot order
                const float tm = fminf(fminf(t0, t1), fminf(t2, t3));
                const bool any_hit = tm < CUDART_INF_F;
                const bool e0 = t0 == tm, e1 = !e0 && t1 == tm, e2 = !e0 && !e1 && t2 == tm, e3 = !e0 && !e1 && !e2;
                if (t3 < CUDART_INF_F && !e3) RC_PUSH(r3)
                if (t2 < CUDART_INF_F && !e2) RC_PUSH(r2)
                if (t1 < CUDART_INF_F && !e1) RC_PUSH(r1)
                if (t0 < CUDART_INF_F && !e0) RC_PUSH(r0)
                if (COUNT && (uint32_t)sp > lc.max_stack) lc.max_stack = (uint32_t)sp;
                const uint32_t top = RC_TOP();
                const uint32_t rn = e0 ? r0 : (e1 ? r1 : (e2 ? r2 : r3));
                cur = any_hit ? rn : top;
                sp -= any_hit ? 0 : 1;
                if (ovf) { cur = RC_INVALID; leaf = 0; sp = 0; }
            }
        }
    }
#undef RC_PUSH
#undef RC_TOP
    if (COUNT) {
        atomicAdd(&counters->rays, traced);
        atomicAdd(&counters->nodes, (unsigned long long)lc.nodes);
        atomicAdd(&counters->box_tests, (unsigned long long)lc.box_tests);
        atomicAdd(&counters->tri_tests, (unsigned long long)lc.tri_tests);
        atomicAdd(&counters->inst_entries, (unsigned long long)lc.inst_entries);
        atomicMax(&counters->max_stack, (unsigned long long)lc.max_stack);
    }
}
