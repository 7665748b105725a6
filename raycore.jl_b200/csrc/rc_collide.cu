// rc_collide.cu — broad-phase instance collision on the TLAS: collide_instances (src/collision.jl:189-233,
// kernel :81-156) and collide_instances_any (:241-261).  Runs on the reference-identical BVH2 TLAS so the contact list is
// byte-identical to the reference algorithm's, including its order (per leaf, contacts are written back-to-front, :138).
#include <cuda_runtime.h>

#include <string>

#include "rc_build.h"
#include "rc_build_core.cuh"

#define RC_COLLIDE_STACK 64  // reference: MVector{16} without overflow check (:97)

// box k (0 / 1) of a reference-layout node: (aabb0_min, aabb0_max) or (aabb1_min, aabb1_max)
__device__ __forceinline__ void node_box(const RcNode2 &nd, int k, f3 &lo, f3 &hi) {
    if (k == 0) { lo = mk3(nd.aabb0_min[0], nd.aabb0_min[1], nd.aabb0_min[2]); hi = mk3(nd.aabb0_max[0], nd.aabb0_max[1], nd.aabb0_max[2]); }
    else { lo = mk3(nd.aabb1_min[0], nd.aabb1_min[1], nd.aabb1_min[2]); hi = mk3(nd.aabb1_max[0], nd.aabb1_max[1], nd.aabb1_max[2]); }
}
__device__ __forceinline__ bool overlaps(f3 amin, f3 amax, f3 bmin, f3 bmax) {  // aabb_overlaps, :49-51
    return amax.x >= bmin.x && amax.y >= bmin.y && amax.z >= bmin.z && amin.x <= bmax.x && amin.y <= bmax.y && amin.z <= bmax.z;
}

// collide_instances_kernel!: one thread per TLAS leaf.  excl == nullptr: counting pass (counts[p] = contacts of leaf p);
// else writing pass with excl = exclusive prefix sums of the counts.
__global__ void k_collide(const RcNode2 *__restrict__ nodes, uint32_t n, uint32_t *__restrict__ counts, const uint32_t *__restrict__ excl,
                          rc_contact_pair *__restrict__ contacts, uint32_t *__restrict__ overflow) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const RcNode2 leaf = nodes[n - 1 + p];
    f3 a_min, a_max;
    node_box(leaf, 0, a_min, a_max);
    const uint32_t instance_a = leaf.child1;
    const uint32_t total = excl ? counts[p] : 0u;
    uint32_t stack[RC_COLLIDE_STACK];
    int sp = 0;
    uint32_t node_index = 1, count = 0;
    for (;;) {
        const RcNode2 nd = nodes[node_index - 1];
        if (nd.child0 != RC_INVALID) {
            f3 l0, h0, l1, h1;
            node_box(nd, 0, l0, h0);
            node_box(nd, 1, l1, h1);
            const bool o0 = overlaps(a_min, a_max, l0, h0), o1 = overlaps(a_min, a_max, l1, h1);
            if (o0 && o1) {
                if (sp >= RC_COLLIDE_STACK) { atomicAdd(overflow, 1u); break; }
                stack[sp++] = nd.child1;
                node_index = nd.child0;
                continue;
            } else if (o0) { node_index = nd.child0; continue; }
            else if (o1) { node_index = nd.child1; continue; }
        } else {
            const uint32_t instance_b = nd.child1;
            f3 bl, bh;
            node_box(nd, 0, bl, bh);
            if (instance_b > instance_a && overlaps(a_min, a_max, bl, bh)) {
                count++;
                if (excl) {  // write_idx = inclusive[i] - count + 1 (1-based) == exclusive + total - count (0-based)
                    rc_contact_pair c;
                    c.instance_a = instance_a + 1u;
                    c.instance_b = instance_b + 1u;
                    contacts[excl[p] + total - count] = c;
                }
            }
        }
        if (sp > 0) node_index = stack[--sp];
        else break;
    }
    if (!excl) counts[p] = count;
}

bool rc_collide_count(cudaStream_t st, const RcDeviceTlas &t, uint32_t *d_counts, uint32_t *d_overflow) {
    if (t.n == 0) return true;
    k_collide<<<(t.n + 127) / 128, 128, 0, st>>>(t.nodes2, t.n, d_counts, nullptr, nullptr, d_overflow);
    return cudaGetLastError() == cudaSuccess;
}

bool rc_collide_write(cudaStream_t st, const RcDeviceTlas &t, uint32_t *d_counts, const uint32_t *d_excl, rc_contact_pair *d_contacts, uint32_t *d_overflow) {
    if (t.n == 0) return true;
    k_collide<<<(t.n + 127) / 128, 128, 0, st>>>(t.nodes2, t.n, d_counts, d_excl, d_contacts, d_overflow);
    return cudaGetLastError() == cudaSuccess;
}
