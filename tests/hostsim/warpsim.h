// warpsim.h — what rc_trace_fast.cuh needs from CUDA, restated for g++ so that the shipped traversal kernel (k_trace_wide) can be
// compiled for the CPU and run one fibre per lane.  TEST INFRASTRUCTURE ONLY (tests/hostsim): never part of libraycore_cuda.so.
//
// Execution model: the 32 lanes of a warp are ucontext fibres on one OS thread.  A lane runs until it reaches a warp intrinsic
// (__reduce_add_sync / __ballot_sync / __shfl_sync), deposits its operand in a per-warp slot array and yields; the scheduler
// resumes a lane only after every lane of its warp has arrived, so each intrinsic sees all 32 operands — lock step exactly where
// the hardware guarantees it, arbitrary interleaving elsewhere.  Slots are double-buffered because a resumed lane may reach the
// next intrinsic before its neighbours have read the previous one.  "Shared memory" is each fibre's own copy of the array: the
// kernel only ever touches its own [depth][thread] column.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <ucontext.h>

#include <vector>

#ifndef RC_TRACE_THREADS
#define RC_TRACE_THREADS 128
#endif
#define __global__
#define __device__
#define __forceinline__ inline
#define __shared__
#define __launch_bounds__(...)
#define CUDART_INF_F (__builtin_inff())

struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) ulonglong2 { unsigned long long x, y; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
static inline float __fdividef(float a, float b) { return a / b; }  // host stand-in; only the conservative sphere cull uses it
static inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
struct __half2 { uint16_t x, y; };
static inline float ws_half2float(uint16_t h) {  // IEEE binary16 -> binary32, exact (subnormals included)
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16, e = (h >> 10) & 0x1Fu, m = h & 0x3FFu;
    float f;
    if (e == 0) f = ldexpf((float)m, -24);
    else if (e == 31) f = m ? NAN : INFINITY;
    else f = ldexpf((float)(m | 0x400u), (int)e - 25);
    uint32_t u;
    memcpy(&u, &f, 4);
    u |= sign;
    memcpy(&f, &u, 4);
    return f;
}
static inline float2 __half22float2(__half2 h) { float2 r; r.x = ws_half2float(h.x); r.y = ws_half2float(h.y); return r; }
static inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s) {  // PRMT, default mode (selector nibbles 0..7, no sign replication used here)
    const uint64_t v = ((uint64_t)y << 32) | x;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= (uint32_t)((v >> (8 * ((s >> (4 * i)) & 7u))) & 0xFFu) << (8 * i);
    return r;
}
static inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
template <class T> static inline T __ldg(const T *p) { return *p; }
template <class T> static inline T __ldcs(const T *p) { return *p; }
template <class T> static inline void __stcs(T *p, const T &v) { *p = v; }
template <class T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }  // one OS thread: plain read-modify-write
template <class T> static inline T atomicMax(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }

namespace warpsim {
struct Lane {
    Lane() { memset(&ctx, 0, sizeof ctx); }
    ucontext_t ctx;
    std::vector<char> stack;
    bool done = false;
    unsigned gen = 0;
};
struct Warp {
    Lane lane[32];
    uint64_t slot[2][32];
    unsigned first_tid = 0;
    uint64_t iter = 0, last_iter[4] = {~0ull, ~0ull, ~0ull, ~0ull};  // step statistics: scheduler iteration number, last iteration each step kind ran in
};
struct Tid { unsigned x; };
extern Warp *g_warp;     // the warp and lane the scheduler is running right now
extern int g_lane;
extern Tid g_tid;
extern ucontext_t g_sched;
extern uint64_t g_exchanges;
extern uint64_t g_idle[6];  // lanes sitting out a node step, by what they wait for: T with a second leaf reached, T at the level sentinel, T with an empty stack, X, F, dead
extern uint64_t g_step_iters[4], g_step_lanes[4];  // per step kind (N, T, X, F): warp iterations, active lanes summed
// every lane contributes v; returns the 32 operands of this exchange
static inline const uint64_t *exchange(uint64_t v) {
    Warp *w = g_warp;
    const int l = g_lane;
    const unsigned g = w->lane[l].gen++ & 1u;
    w->slot[g][l] = v;
    g_exchanges++;
    swapcontext(&w->lane[l].ctx, &g_sched);  // resumed once all lanes of the warp have deposited their operand
    return w->slot[g];
}
}  // namespace warpsim
#define threadIdx (warpsim::g_tid)
// step statistics: a step kind counts one iteration when at least one lane of the warp executes its body
#define RC_SIM_ITER() { if (warpsim::g_lane == 0) warpsim::g_warp->iter++; }
#define RC_SIM_STEP(kind, active)                                                                     \
    {                                                                                                 \
        if (active) {                                                                                 \
            warpsim::Warp *w_ = warpsim::g_warp;                                                      \
            if (w_->last_iter[kind] != w_->iter) { w_->last_iter[kind] = w_->iter; warpsim::g_step_iters[kind]++; } \
            warpsim::g_step_lanes[kind]++;                                                            \
        }                                                                                             \
    }

#define RC_SIM_IDLE(vote_word, cur_ref, leaf_ref)                                                          \
    {                                                                                                       \
        if (!((vote_word) & RC_VOTE_N)) {                                                                   \
            int k_ = 5;                                                                                     \
            if ((vote_word) & RC_VOTE_T) k_ = ((cur_ref) == RC_SENTINEL) ? 1 : ((cur_ref) == RC_INVALID ? 2 : 0); \
            else if ((vote_word) & RC_VOTE_X) k_ = 3;                                                       \
            else if ((vote_word) & RC_VOTE_F) k_ = 4;                                                       \
            warpsim::g_idle[k_]++;                                                                          \
        }                                                                                                   \
    }

static inline uint32_t __reduce_add_sync(uint32_t, uint32_t v) {
    const uint64_t *s = warpsim::exchange(v);
    uint32_t r = 0;
    for (int i = 0; i < 32; i++) r += (uint32_t)s[i];
    return r;
}
static inline uint32_t __ballot_sync(uint32_t, bool p) {
    const uint64_t *s = warpsim::exchange(p ? 1u : 0u);
    uint32_t r = 0;
    for (int i = 0; i < 32; i++) r |= (uint32_t)(s[i] & 1u) << i;
    return r;
}
static inline unsigned long long __shfl_sync(uint32_t, unsigned long long v, int src) { return warpsim::exchange(v)[src & 31]; }
