// rc_trace.cu — batched closest_hit / any_hit kernels (replaces the per-ray device functions
// closest_hit / any_hit of src/instanced-bvh.jl:1902-2140 and the batched HW entry
// Lava.trace_closest_hits!, docs/src/hw_acceleration.md:143-146).
//
// Persistent threads: the grid is sized to fill the 148 SMs once; every lane pulls its next ray from a
// global work counter as soon as its current ray retires, so a long ray never idles the other 31 lanes
// of its warp behind a loop exit.
#include <cuda_runtime.h>

#include <string>

#include "rc_trace.h"
#include "rc_trace_core.cuh"

__device__ __forceinline__ rc_ray load_ray(const rc_ray *rays, unsigned long long i) {
    const float4 *p = reinterpret_cast<const float4 *>(rays + i);
    float4 a = __ldg(p), b = __ldg(p + 1);
    rc_ray r;
    r.origin[0] = a.x; r.origin[1] = a.y; r.origin[2] = a.z; r.tmin = a.w;
    r.dir[0] = b.x; r.dir[1] = b.y; r.dir[2] = b.z; r.tmax = b.w;
    return r;
}

__device__ __forceinline__ void store_hit(rc_hit *hits, unsigned long long i, const rc_hit &h) {
    float4 *p = reinterpret_cast<float4 *>(hits + i);
    // streaming stores: hit records are write-once, keep them out of the way of the BVH working set in L2
    __stcs(p, make_float4(__uint_as_float(h.hit), h.t, __uint_as_float(h.primitive_id), __uint_as_float(h.instance_custom_index)));
    __stcs(p + 1, make_float4(h.bary_u, h.bary_v, __uint_as_float(h.instance_id), __uint_as_float(h.metadata)));
}

template <bool WIDE, bool ANY, bool COUNT>
__global__ void __launch_bounds__(RC_TRACE_THREADS) k_trace(RcScene sc, const rc_ray *__restrict__ rays, rc_hit *__restrict__ hits, unsigned long long n,
                                                            unsigned long long *__restrict__ work, RcCounters *__restrict__ counters, uint32_t *__restrict__ overflow) {
    RcLocalCounters lc = {0, 0, 0, 0, 0};
    unsigned long long traced = 0;
    while (true) {
        unsigned long long i = atomicAdd(work, 1ull);
        if (i >= n) break;
        rc_ray r = load_ray(rays, i);
        rc_hit h;
        bool ok = WIDE ? rc_trace_wide<ANY, COUNT>(sc, r, h, &lc) : rc_trace_reference_order<ANY, COUNT>(sc, r, h, &lc);
        if (!ok) atomicAdd(overflow, 1u);
        store_hit(hits, i, h);
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

template <bool WIDE, bool ANY, bool COUNT>
static void launch(cudaStream_t st, int blocks, const RcScene &sc, const rc_ray *rays, rc_hit *hits, unsigned long long n, unsigned long long *work,
                   RcCounters *counters, uint32_t *overflow) {
    k_trace<WIDE, ANY, COUNT><<<blocks, RC_TRACE_THREADS, 0, st>>>(sc, rays, hits, n, work, counters, overflow);
}

bool rc_launch_trace(cudaStream_t st, const RcTraceLaunch &L, std::string &err) {
    if (L.n == 0) return true;
    cudaMemsetAsync(L.work, 0, sizeof(unsigned long long), st);
    unsigned long long want = (L.n + RC_TRACE_THREADS - 1) / RC_TRACE_THREADS;
    int blocks = (int)(want < (unsigned long long)L.max_blocks ? want : (unsigned long long)L.max_blocks);
    if (blocks < 1) blocks = 1;
#define RC_GO(W, A, C) launch<W, A, C>(st, blocks, L.scene, L.rays, L.hits, L.n, L.work, L.counters, L.overflow)
    if (L.wide) {
        if (L.any) { if (L.count) RC_GO(true, true, true); else RC_GO(true, true, false); }
        else { if (L.count) RC_GO(true, false, true); else RC_GO(true, false, false); }
    } else {
        if (L.any) { if (L.count) RC_GO(false, true, true); else RC_GO(false, true, false); }
        else { if (L.count) RC_GO(false, false, true); else RC_GO(false, false, false); }
    }
#undef RC_GO
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { err = std::string("trace launch: ") + cudaGetErrorString(e); return false; }
    return true;
}

int rc_trace_max_blocks(int device) {
    int sms = 148, per_sm = 1;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_trace<true, false, false>, RC_TRACE_THREADS, 0);
    if (per_sm < 1) per_sm = 1;
    return sms * per_sm;
}
