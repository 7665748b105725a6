// rc_types.h — device-resident data layouts of libraycore_cuda (host + device).
//
// Everything the traversal kernels fetch is 16-byte aligned and sized in 16-byte quanta so it
// moves as LDG.128: wide nodes 64 B (2 sectors), triangles 48 B, instance records 64 B.
#pragma once
#include <stdint.h>

#include "../../include/raycore_cuda.h"

#define RC_INVALID 0xFFFFFFFFu       // INVALID_NODE, src/instanced-bvh.jl:65
#define RC_SENTINEL 0xEFFFFFFFu      // TOP_LEVEL_SENTINEL (src/instanced-bvh.jl:1733 uses 0xFFFFFFFE; the value is internal to the stack)
#define RC_LEAF_BIT 0x80000000u      // wide-node child reference: leaf flag
#define RC_TLAS_LEAF_TAG 0xC0000000u // TLAS leaves carry bit 30 too, so a reference alone tells "triangles" [0x8..,0xC..) from
                                     // "change level" [0xC.., 0xF..) (instance leaf or RC_SENTINEL) without consulting the traversal level
#define RC_LEAF_COUNT_SHIFT 28       // bits 30..28: triangle count - 1
#define RC_LEAF_START_MASK 0x0FFFFFFFu
#ifndef RC_BLAS_LEAF_MAX
#define RC_BLAS_LEAF_MAX 2           // triangles per wide-BVH leaf (<= 8); 2 measured best on C2 (profiles/r1_leafmax.md)
#endif

// BVH2 node in the reference's field order (BVHNode2, src/instanced-bvh.jl:50-63) padded 60 -> 64 B.
// Child / parent / primitive indices keep the reference's 1-based values.
struct __attribute__((aligned(16))) RcNode2 {
    float aabb0_min[3], aabb0_max[3], aabb1_min[3], aabb1_max[3];
    uint32_t child0, child1, parent, pad;
};

// Wide (4-ary) node with child boxes quantised to 8 bits per plane against the node's own frame.
//   origin o, per-axis scale 2^(e-127);  child k plane = o + q * scale,  lo rounded down, hi rounded up.
//   sx/sy/sz hold the scale pre-multiplied by 2^24 (the traversal kernel's plane decode yields q * 2^-24), e <= RC_QUANT_EXP_MAX.
//   qlo*/qhi*: byte k = child k.   child[k]: RC_LEAF_BIT|count-1|start = leaf (BLAS: start = first Morton-sorted triangle,
//   TLAS: RC_TLAS_LEAF_TAG | instance index); else wide-node index.  An unused slot repeats child 0's reference under an
//   inverted box (qlo = 255, qhi = 0): it fails the slab test, and if rounding slack ever lets it through, the traversal
//   merely revisits child 0 — so the kernels need no per-slot validity test.
struct __attribute__((aligned(16))) RcNode4 {
    float ox, oy, oz;
    float sx;
    uint32_t qlox, qloy, qloz, qhix;
    uint32_t qhiy, qhiz, child0, child1;
    uint32_t child2, child3;
    float sy, sz;
};
#define RC_QUANT_EXP_MAX 230u  // 2^(e-127+24) must stay finite; extents beyond 255 * 2^103 overflow the reference's own Moeller-Trumbore

// One triangle = 3 x float4 in Morton-sorted order:  (v0, primitive_id), (v1, metadata), (v2, face_index)
//   primitive_id = position in the degenerate-filtered input list, face_index = position in the submitted soup
struct __attribute__((aligned(16))) RcTri {
    float v0[3]; uint32_t prim_id;
    float v1[3]; uint32_t metadata;
    float v2[3]; uint32_t face_index;
};

// Per-instance traversal record (96 B, 32-B aligned: the transform + pointers move as two LDG.256, the sphere as one LDG.128 —
// three L1 wavefronts per entering lane instead of five): world->local transform + the BLAS arrays it enters + the BLAS's bounding sphere in its own
// (local) space.  A ray whose transformed copy misses that sphere cannot hit any triangle of the instance, so the level step culls
// the entry before the BLAS walk starts (an instance's world AABB is ~2x the cross-section of a round mesh: profiles/README.md r2).
struct __attribute__((aligned(32))) RcInstanceRec {
    float inv[12];           // Mat3x4f rows, src/instanced-bvh.jl:94
    const RcNode4 *nodes4;   // BLAS wide nodes (root = index 1)
    const RcTri *tris;
    float sphere[4];         // centre xyz, radius^2 (conservative: every vertex lies inside); radius^2 = +Inf disables the test
    float wsphere[4];        // the same sphere in world space (centre, radius^2 scaled by the transform's largest stretch, rc_world_sphere): only read by the
                             // RC_WORLD_CULL experiment of rc_trace_fast.cuh (cull in the settle, before any transform is fetched; measured slower, off)
};

// Cold per-instance data (hit write-back and the reference-order path)
struct RcInstanceAux {
    const RcNode2 *nodes2;   // BLAS BVH2 (root = index 1, stored at [0])
    uint32_t n_prims;
    uint32_t custom_index;   // InstanceDescriptor.instance_id
};

struct RcCounters {  // instrumented build only
    unsigned long long rays, nodes, box_tests, tri_tests, inst_entries, max_stack;
};

// Everything a trace kernel needs (passed by value)
struct RcScene {
    const RcNode4 *tlas4;   // root = index 1
    const RcNode2 *tlas2;   // root = index 1 (stored at [0])
    const RcInstanceRec *inst;
    const RcInstanceAux *aux;
    uint32_t n_instances;
};
