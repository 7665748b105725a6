// rc_build.h — library-internal interface of the GPU builder (rc_build.cu)
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include <string>
#include <vector>

#include "rc_types.h"

struct RcBox;
struct RcTopo;

struct RcDeviceBlas {
    uint32_t n = 0;           // valid (non-degenerate) triangles
    uint32_t n_faces_in = 0;  // faces submitted
    RcNode2 *nodes2 = nullptr;  // 2n-1, reference layout, node k at [k-1]; only with RC_BUILD_KEEP_BVH2 (reference-order mode, BVH2 read-backs)
    RcNode4 *nodes4 = nullptr;  // n+1 slots, indexed by BVH2 internal node number, root = [1]
    RcTri *tris = nullptr;      // n, Morton-sorted
    RcBox *hull = nullptr;      // RC_HULL_BOXES boxes of BVH2 subtrees covering the whole BLAS (tight instance bounds for the wide TLAS)
    RcTopo *topo = nullptr;     // kept radix tree (RC_BUILD_ALLOW_REFIT): per internal node children + span ...
    uint32_t *parent = nullptr; // ... and parent links, so a vertex update can re-fit instead of rebuilding
    float *normals = nullptr;   // optional, 9 floats per primitive indexed by primitive_id (rc_set_normals; shading-side data of the wavefront stages)
    float root_aabb[6] = {0, 0, 0, 0, 0, 0};
    float sphere[4] = {0, 0, 0, INFINITY};  // bounding sphere in local space: centre = centre of the root box, radius^2 over all vertices (instance-entry cull)
};

#define RC_HULL_BOXES 16

struct RcBlasPtrs {  // device-visible BLAS table entry
    const RcNode2 *nodes2;
    const RcNode4 *nodes4;
    const RcTri *tris;
    const RcBox *hull;  // RC_HULL_BOXES entries (unused ones are empty boxes)
    uint32_t n, pad;
    float sphere[4];
};

struct RcDeviceTlas {
    uint32_t n = 0;
    uint32_t n_blas_cap = 0;  // entries d_blas_roots / d_blas_ptrs were allocated for
    RcNode2 *nodes2 = nullptr;
    RcNode4 *nodes4 = nullptr;
    RcInstanceRec *rec = nullptr;
    RcInstanceAux *aux = nullptr;
    rc_instance_desc *d_inst = nullptr;
    float *d_blas_roots = nullptr;
    RcBlasPtrs *d_blas_ptrs = nullptr;
    RcBox *inst_boxes = nullptr;        // reference-identical world boxes (8 corners of the BLAS root box) -> BVH2
    RcBox *inst_boxes_tight = nullptr;  // union over the BLAS hull boxes -> wide TLAS only (conservative, tighter)
    RcBox *boxes_tight = nullptr;       // per-node boxes of the tight fit
    uint32_t *leaf_map = nullptr;  // sorted position -> instance index
    RcTopo *topo = nullptr;
    uint32_t *parent = nullptr;
    unsigned char *fit_work = nullptr;  // segment tables + spanning-node list of the fit (rc_build.cu), kept for refits
    RcBox *boxes = nullptr;
    uint32_t *d_small = nullptr;
    float root_aabb[6] = {0, 0, 0, 0, 0, 0};
};

// build_flags: RC_BUILD_KEEP_BVH2 | RC_BUILD_ALLOW_REFIT (include/raycore_cuda.h)
bool rc_build_blas(cudaStream_t st, const float *d_verts, const uint32_t *d_face_meta, uint32_t n_faces, uint32_t build_flags, RcDeviceBlas *out, std::string &err);
// re-fit a BLAS built with RC_BUILD_ALLOW_REFIT to new vertex positions of the same faces; *refitted = false when that is not possible
// (no kept topology, face count or degenerate set changed) and the caller has to rebuild
bool rc_refit_blas(cudaStream_t st, const float *d_verts, uint32_t n_faces, RcDeviceBlas *b, bool *refitted, std::string &err);
void rc_free_blas(RcDeviceBlas *b, cudaStream_t st);
// serialised BLAS (host blob <-> device arrays, byte-identical restore; layout in rc_build.cu)
uint64_t rc_blas_blob_bytes(const RcDeviceBlas &b);
bool rc_blas_export(cudaStream_t st, const RcDeviceBlas &b, void *blob, uint64_t capacity, std::string &err);
bool rc_blas_import(cudaStream_t st, const void *blob, uint64_t size, RcDeviceBlas *out, std::string &err);
// the host-side part of the import checks alone (header, section table, size, payload hash): no GPU needed
bool rc_blas_blob_check(const void *blob, uint64_t size, uint32_t *n_triangles, uint32_t *n_faces_in, uint32_t *has_normals, std::string &err);
bool rc_build_tlas(cudaStream_t st, const rc_instance_desc *h_inst, uint32_t n, const std::vector<RcBlasPtrs> &blas, const std::vector<float> &blas_roots,
                   RcDeviceTlas *t, std::string &err);
bool rc_refit_tlas(cudaStream_t st, const rc_instance_desc *h_inst, uint32_t n, RcDeviceTlas *t, std::string &err);
void rc_free_tlas(RcDeviceTlas *t, cudaStream_t st);

// generic device exclusive scan (rc_build.cu): out[i] = sum(in[0..i)), *d_total = sum of all; tile_tmp >= ceil(n / 2048) words
void rc_exclusive_scan_u32(cudaStream_t st, const uint32_t *in, uint32_t *out, uint32_t n, uint32_t *tile_tmp, uint32_t *d_total);
// collision broad phase (rc_collide.cu)
bool rc_collide_count(cudaStream_t st, const RcDeviceTlas &t, uint32_t *d_counts, uint32_t *d_overflow);
bool rc_collide_write(cudaStream_t st, const RcDeviceTlas &t, uint32_t *d_counts, const uint32_t *d_excl, rc_contact_pair *d_contacts, uint32_t *d_overflow);
