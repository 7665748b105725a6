// rc_trace.cu — batched closest_hit / any_hit launchers (replaces the per-ray device functions
// closest_hit / any_hit of src/instanced-bvh.jl:1902-2140 and the batched HW entry
// Lava.trace_closest_hits!, docs/src/hw_acceleration.md:143-146).
//
//   k_trace_wide (rc_trace_fast.cuh)  default: quantised BVH4, persistent lanes, warp-level step scheduler
//   k_trace                           reference-order BVH2 walk (bit-identical to the reference incl. ties):
//                                     thin persistent wrapper around rc_trace_reference_order (rc_trace_core.cuh)
#include <cuda_runtime.h>

#include <string>

#include "rc_trace.h"
#include "rc_trace_core.cuh"
#include "rc_trace_fast.cuh"

template <bool ANY, bool COUNT, bool WT>
__global__ void __launch_bounds__(RC_TRACE_THREADS) k_trace(RcScene sc, const rc_ray *__restrict__ rays, rc_hit *__restrict__ hits, unsigned long long n,
                                                            unsigned long long *__restrict__ work, RcCounters *__restrict__ counters, uint32_t *__restrict__ overflow) {
    RcLocalCounters lc = {0, 0, 0, 0, 0};
    unsigned long long traced = 0;
    while (true) {
        unsigned long long i = atomicAdd(work, 1ull);
        if (i >= n) break;
        rc_ray r = rc_load_ray(rays, i);
        rc_hit h;
        if (!rc_trace_reference_order<ANY, COUNT, WT>(sc, r, h, &lc)) atomicAdd(overflow + 1, 1u);
        rc_store_hit(hits, i, h);
        traced++;
    }
    if (COUNT) {
        atomicAdd(&counters->rays, traced);
        atomicAdd(&counters->nodes, (unsigned long long)lc.nodes);
        atomicAdd(&counters->box_tests, (unsigned long long)lc.box_tests);
        atomicAdd(&counters->tri_tests, (unsigned long long)lc.tri_tests);
        atomicAdd(&counters->inst_entries, (unsigned long long)lc.inst_entries);
        atomicMax(&counters->max_stack, (unsigned long long)lc.max_stack);
    }
}

// Re-trace the rays whose short (shared-memory) stack overflowed in k_trace_wide with the deep-stack generic body.
// overflow[0] = rays flagged by the fast kernel (exit immediately when 0), overflow[1] = rays no stack could hold (error).
// ovf_list (nullable): ovf_list[0] = number of flagged rays the fast kernel listed, entries from [1], capacity ovf_cap; when the list holds
// every flagged ray only those are visited, otherwise every hit record is scanned for the mark.
template <bool ANY, bool WT>
__global__ void __launch_bounds__(RC_TRACE_THREADS) k_trace_fixup(RcScene sc, const rc_ray *__restrict__ rays, rc_hit *__restrict__ hits, unsigned long long n,
                                                                  uint32_t *__restrict__ overflow, const unsigned long long *__restrict__ ovf_list, uint32_t ovf_cap, bool zero_tmin) {
    const uint32_t flagged = *reinterpret_cast<volatile uint32_t *>(overflow);
    if (flagged == 0) return;
    const bool listed = ovf_list && flagged <= ovf_cap && ovf_list[0] == flagged;
    const unsigned long long count = listed ? flagged : n;
    for (unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long i = listed ? ovf_list[1 + k] : k;
        if (hits[i].hit != RC_OVERFLOW_MARK) continue;
        rc_ray r = rc_load_ray(rays, i);
        if (zero_tmin) r.tmin = 0.0f;
        rc_hit h;
        if (!rc_trace_wide<ANY, false, WT>(sc, r, h, nullptr)) atomicAdd(overflow + 1, 1u);
        rc_store_hit(hits, i, h);
    }
}

__global__ void k_fill_miss(rc_hit *__restrict__ hits, unsigned long long n) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 *p = reinterpret_cast<float4 *>(hits + i);
    p[0] = make_float4(0.f, 0.f, 0.f, 0.f);
    p[1] = make_float4(0.f, 0.f, 0.f, 0.f);
}

bool rc_launch_trace(cudaStream_t st, const RcTraceLaunch &L, std::string &err) {
    if (L.n == 0) return true;
    if (L.scene.n_instances == 0) {  // empty TLAS: every ray misses (test/test_tlas_stress.jl:808-831)
        k_fill_miss<<<(unsigned)((L.n + 255) / 256), 256, 0, st>>>(L.hits, L.n);
    } else {
        cudaMemsetAsync(L.work, 0, sizeof(unsigned long long) * (L.ovf_cap ? 2 : 1), st);
        cudaMemsetAsync(L.overflow, 0, sizeof(uint32_t), st);
        unsigned long long want = (L.n + RC_TRACE_THREADS - 1) / RC_TRACE_THREADS;
        // the single-instance variant is compiled for RC_MIN_BLOCKS_SINGLE resident CTAs per SM (fewer registers: no world-ray copy)
        const unsigned long long cap = (L.wide && L.scene.n_instances == 1u) ? (unsigned long long)L.max_blocks * RC_MIN_BLOCKS_SINGLE / RC_MIN_BLOCKS : (unsigned long long)L.max_blocks;
        int blocks = (int)(want < cap ? want : cap);
        if (blocks < 1) blocks = 1;
#define RC_ARGS L.scene, L.rays, L.hits, L.n, L.work, L.counters, L.overflow
#define RC_WARGS L.scene, RcIoArrays{L.rays, L.hits, L.zero_tmin, L.ovf_cap ? L.work + 1 : nullptr, L.ovf_cap}, L.n, L.work, L.counters, L.overflow
        const bool single = L.scene.n_instances == 1u;
        if (L.wide) {
            // (ANY, COUNT, SINGLE, WT): the instrumented build exists for Moeller-Trumbore only
#define RC_LAUNCH_WIDE(A, C, S, W) k_trace_wide<A, C, RcIoArrays, S, W><<<blocks, RC_TRACE_THREADS, 0, st>>>(RC_WARGS)
#define RC_PICK_SINGLE(A, C, W)                                      \
    {                                                                \
        if (single) RC_LAUNCH_WIDE(A, C, true, W);                   \
        else RC_LAUNCH_WIDE(A, C, false, W);                         \
    }
            if (L.watertight) { if (L.any) RC_PICK_SINGLE(true, false, true) else RC_PICK_SINGLE(false, false, true) }
            else if (L.any) { if (L.count) RC_PICK_SINGLE(true, true, false) else RC_PICK_SINGLE(true, false, false) }
            else { if (L.count) RC_PICK_SINGLE(false, true, false) else RC_PICK_SINGLE(false, false, false) }
#undef RC_PICK_SINGLE
#undef RC_LAUNCH_WIDE
        } else if (L.watertight) {
            if (L.any) k_trace<true, false, true><<<blocks, RC_TRACE_THREADS, 0, st>>>(RC_ARGS);
            else k_trace<false, false, true><<<blocks, RC_TRACE_THREADS, 0, st>>>(RC_ARGS);
        } else {
            if (L.any) { if (L.count) k_trace<true, true, false><<<blocks, RC_TRACE_THREADS, 0, st>>>(RC_ARGS); else k_trace<true, false, false><<<blocks, RC_TRACE_THREADS, 0, st>>>(RC_ARGS); }
            else { if (L.count) k_trace<false, true, false><<<blocks, RC_TRACE_THREADS, 0, st>>>(RC_ARGS); else k_trace<false, false, false><<<blocks, RC_TRACE_THREADS, 0, st>>>(RC_ARGS); }
        }
#undef RC_ARGS
#undef RC_WARGS
        if (L.wide) {
            if (L.watertight) {
                if (L.any) k_trace_fixup<true, true><<<blocks, RC_TRACE_THREADS, 0, st>>>(L.scene, L.rays, L.hits, L.n, L.overflow, L.ovf_cap ? L.work + 1 : nullptr, L.ovf_cap, L.zero_tmin);
                else k_trace_fixup<false, true><<<blocks, RC_TRACE_THREADS, 0, st>>>(L.scene, L.rays, L.hits, L.n, L.overflow, L.ovf_cap ? L.work + 1 : nullptr, L.ovf_cap, L.zero_tmin);
            } else {
                if (L.any) k_trace_fixup<true, false><<<blocks, RC_TRACE_THREADS, 0, st>>>(L.scene, L.rays, L.hits, L.n, L.overflow, L.ovf_cap ? L.work + 1 : nullptr, L.ovf_cap, L.zero_tmin);
                else k_trace_fixup<false, false><<<blocks, RC_TRACE_THREADS, 0, st>>>(L.scene, L.rays, L.hits, L.n, L.overflow, L.ovf_cap ? L.work + 1 : nullptr, L.ovf_cap, L.zero_tmin);
            }
        }
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { err = std::string("trace launch: ") + cudaGetErrorString(e); return false; }
    return true;
}

int rc_trace_max_blocks(int device) {
    int sms = 148, per_sm = 1;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_trace_wide<false, false, RcIoArrays>, RC_TRACE_THREADS, 0);
    if (per_sm < 1) per_sm = 1;
    return sms * per_sm;
}
