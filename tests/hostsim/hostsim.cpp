// hostsim.cpp — CPU instantiation of the library's per-element device code, for unit tests on machines
// without a GPU.  TEST INFRASTRUCTURE ONLY: it is compiled by tests/ (g++ -ffp-contract=off), never linked
// into libraycore_cuda.so and never reachable from the C ABI.  It includes the very same RC_HD bodies the CUDA
// kernels call (rc_device.cuh, rc_build_core.cuh, rc_trace_core.cuh) and drives them with sequential loops
// that stand in for the kernel launches of rc_build.cu / rc_trace.cu.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>

#include "../../raycore.jl_b200/csrc/rc_build_core.cuh"
#include "../../raycore.jl_b200/csrc/rc_trace_core.cuh"
#include "../../raycore.jl_b200/csrc/rc_wave_core.cuh"
// the shipped traversal kernel itself, compiled for the CPU (one fibre per lane; see warpsim.h)
#define RC_WARPSIM 1
#include "../../raycore.jl_b200/csrc/rc_trace_fast.cuh"

namespace {

struct HsTree {
    uint32_t n = 0;
    std::vector<RcNode2> nodes2;
    std::vector<RcNode4> nodes4;
    std::vector<RcBox> boxes;
    std::vector<RcTopo> topo;
    std::vector<uint32_t> parent, leaf_map, codes;
    float root[6];
};

struct HsBlas {
    HsTree tree;
    std::vector<RcTri> tris;
    uint32_t n_faces_in = 0;
    float sphere[4] = {0, 0, 0, INFINITY};
};

struct HsScene {
    std::vector<HsBlas *> blas;
    std::vector<rc_instance_desc> inst;
    HsTree tlas;
    std::vector<RcInstanceRec> rec;
    std::vector<RcInstanceAux> aux;
    RcScene scene;
};

void set_node2(RcNode2 &nd, f3 a0n, f3 a0x, f3 a1n, f3 a1x, uint32_t c0, uint32_t c1, uint32_t par) {
    nd.aabb0_min[0] = a0n.x; nd.aabb0_min[1] = a0n.y; nd.aabb0_min[2] = a0n.z;
    nd.aabb0_max[0] = a0x.x; nd.aabb0_max[1] = a0x.y; nd.aabb0_max[2] = a0x.z;
    nd.aabb1_min[0] = a1n.x; nd.aabb1_min[1] = a1n.y; nd.aabb1_min[2] = a1n.z;
    nd.aabb1_max[0] = a1x.x; nd.aabb1_max[1] = a1x.y; nd.aabb1_max[2] = a1x.z;
    nd.child0 = c0; nd.child1 = c1; nd.parent = par; nd.pad = 0;
}
void set_box(RcBox &b, f3 lo, f3 hi) {
    b.lo[0] = lo.x; b.lo[1] = lo.y; b.lo[2] = lo.z; b.pad0 = 0;
    b.hi[0] = hi.x; b.hi[1] = hi.y; b.hi[2] = hi.z; b.pad1 = 0;
}

unsigned long long g_collapse_checked = 0, g_collapse_mismatch = 0;
// stands in for k_topology + k_fit + k_collapse
void build_tree(HsTree &t, const std::vector<uint32_t> &codes_sorted, const RcTri *tris, const std::vector<RcBox> *inst_boxes, uint32_t leaf_max) {
    uint32_t n = (uint32_t)codes_sorted.size();
    t.n = n;
    t.codes = codes_sorted;
    t.topo.assign(std::max(1u, n - 1), RcTopo{0, 0, 0, 0});
    t.parent.assign(2 * n - 1, RC_INVALID);
    t.boxes.resize(2 * n - 1);
    t.nodes2.resize(2 * n - 1);
    t.nodes4.assign(n + 1, RcNode4{});
    for (uint32_t i = 0; i + 1 < n; i++) {
        RcTopo tp = rc_topology_for_node((int)(i + 1), codes_sorted.data(), (int)n);
        t.topo[i] = tp;
        t.parent[tp.child0 - 1] = i + 1;
        t.parent[tp.child1 - 1] = i + 1;
    }
    std::vector<uint32_t> flags(std::max(1u, n - 1), 0);
    for (uint32_t p = 0; p < n; p++) {
        uint32_t leaf = n - 1 + (p + 1), par = t.parent[leaf - 1];
        f3 lo, hi;
        if (tris) {
            const RcTri &tr = tris[p];
            f3 v0 = mk3(tr.v0[0], tr.v0[1], tr.v0[2]), v1 = mk3(tr.v1[0], tr.v1[1], tr.v1[2]), v2 = mk3(tr.v2[0], tr.v2[1], tr.v2[2]);
            lo = jl_min3(jl_min3(v0, v1), v2);
            hi = jl_max3(jl_max3(v0, v1), v2);
            set_node2(t.nodes2[leaf - 1], v0, v1, v2, mk3(0, 0, 0), RC_INVALID, p + 1, par);
        } else {
            uint32_t inst = t.leaf_map[p];
            const RcBox &b = (*inst_boxes)[inst];
            lo = mk3(b.lo[0], b.lo[1], b.lo[2]);
            hi = mk3(b.hi[0], b.hi[1], b.hi[2]);
            set_node2(t.nodes2[leaf - 1], lo, hi, mk3(0, 0, 0), mk3(0, 0, 0), RC_INVALID, inst, par);
        }
        set_box(t.boxes[leaf - 1], lo, hi);
        uint32_t node = par;
        while (node != RC_INVALID) {
            if (flags[node - 1]++ == 0) break;
            RcTopo tp = t.topo[node - 1];
            const RcBox &b0 = t.boxes[tp.child0 - 1], &b1 = t.boxes[tp.child1 - 1];
            f3 l0 = mk3(b0.lo[0], b0.lo[1], b0.lo[2]), h0 = mk3(b0.hi[0], b0.hi[1], b0.hi[2]);
            f3 l1 = mk3(b1.lo[0], b1.lo[1], b1.lo[2]), h1 = mk3(b1.hi[0], b1.hi[1], b1.hi[2]);
            uint32_t up = t.parent[node - 1];
            set_node2(t.nodes2[node - 1], l0, h0, l1, h1, tp.child0, tp.child1, up);
            set_box(t.boxes[node - 1], jl_min3(l0, l1), jl_max3(h0, h1));
            node = up;
        }
    }
    uint32_t n_int = n > 1 ? n - 1 : 1;
    for (uint32_t i = 0; i < n_int; i++) {
        const uint32_t *lm = t.leaf_map.empty() ? nullptr : t.leaf_map.data();
        t.nodes4[i + 1] = rc_collapse_node(i + 1, t.boxes.data(), t.topo.data(), n, leaf_max, lm);
        // the fetch-once form k_collapse_span uses for the spanning nodes must give the same bytes
        const RcNode4 c = rc_collapse_node_cached(i + 1, t.boxes.data(), t.topo.data(), n, leaf_max, lm);
        g_collapse_checked++;
        if (memcmp(&c, &t.nodes4[i + 1], sizeof c) != 0) g_collapse_mismatch++;
    }
    for (int k = 0; k < 3; k++) { t.root[k] = t.boxes[0].lo[k]; t.root[3 + k] = t.boxes[0].hi[k]; }
}

void stable_sort_pairs(std::vector<uint32_t> &codes, std::vector<uint32_t> &idx) {
    std::vector<uint32_t> perm(codes.size());
    std::iota(perm.begin(), perm.end(), 0u);
    std::stable_sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) { return codes[a] < codes[b]; });
    std::vector<uint32_t> c2(codes.size()), i2(codes.size());
    for (size_t k = 0; k < perm.size(); k++) { c2[k] = codes[perm[k]]; i2[k] = idx[perm[k]]; }
    codes.swap(c2);
    idx.swap(i2);
}

}  // namespace

extern "C" {

// rc_collapse_node_cached against rc_collapse_node over every wide node built so far in this process: {checked, mismatches}
void hs_collapse_cached_stats(unsigned long long *out2) { out2[0] = g_collapse_checked; out2[1] = g_collapse_mismatch; }

void *hs_blas_build(const float *verts, uint32_t n_faces, const uint32_t *face_meta) {
    std::vector<RcTri> tris_in;
    std::vector<RcBox> tri_boxes;
    uint32_t bounds[6];
    for (int c = 0; c < 3; c++) { bounds[c] = rc_float_to_ordered(INFINITY); bounds[3 + c] = rc_float_to_ordered(-INFINITY); }
    for (uint32_t i = 0; i < n_faces; i++) {  // k_face_flags + scan + k_compact_faces
        const float *v = verts + (size_t)i * 9;
        f3 a = mk3(v[0], v[1], v[2]), b = mk3(v[3], v[4], v[5]), c = mk3(v[6], v[7], v[8]);
        if (x_is_degenerate(a, b, c)) continue;
        RcTri t;
        t.v0[0] = a.x; t.v0[1] = a.y; t.v0[2] = a.z; t.prim_id = (uint32_t)tris_in.size();
        t.v1[0] = b.x; t.v1[1] = b.y; t.v1[2] = b.z; t.metadata = face_meta ? face_meta[i] : i + 1u;
        t.v2[0] = c.x; t.v2[1] = c.y; t.v2[2] = c.z; t.face_index = i;
        tris_in.push_back(t);
        f3 lo = jl_min3(jl_min3(a, b), c), hi = jl_max3(jl_max3(a, b), c);
        RcBox bx;
        set_box(bx, lo, hi);
        tri_boxes.push_back(bx);
        float l[3] = {lo.x, lo.y, lo.z}, h[3] = {hi.x, hi.y, hi.z};
        for (int k = 0; k < 3; k++) {
            bounds[k] = std::min(bounds[k], rc_float_to_ordered(l[k]));
            bounds[3 + k] = std::max(bounds[3 + k], rc_float_to_ordered(h[k]));
        }
    }
    uint32_t n = (uint32_t)tris_in.size();
    if (n == 0) return nullptr;
    f3 smin = mk3(rc_ordered_to_float(bounds[0]), rc_ordered_to_float(bounds[1]), rc_ordered_to_float(bounds[2]));
    f3 smax = mk3(rc_ordered_to_float(bounds[3]), rc_ordered_to_float(bounds[4]), rc_ordered_to_float(bounds[5]));
    f3 ext = x_sub3(smax, smin);
    std::vector<uint32_t> codes(n), idx(n);
    for (uint32_t i = 0; i < n; i++) {  // k_morton_prims
        const RcBox &b = tri_boxes[i];
        f3 c = mk3(x_mul(0.5f, x_add(b.lo[0], b.hi[0])), x_mul(0.5f, x_add(b.lo[1], b.hi[1])), x_mul(0.5f, x_add(b.lo[2], b.hi[2])));
        f3 nrm = mk3(x_div(x_sub(c.x, smin.x), ext.x), x_div(x_sub(c.y, smin.y), ext.y), x_div(x_sub(c.z, smin.z), ext.z));
        codes[i] = rc_morton30(nrm);
        idx[i] = i;
    }
    stable_sort_pairs(codes, idx);
    HsBlas *B = new HsBlas();
    B->n_faces_in = n_faces;
    B->tris.resize(n);
    for (uint32_t j = 0; j < n; j++) B->tris[j] = tris_in[idx[j]];  // k_gather_tris
    build_tree(B->tree, codes, B->tris.data(), nullptr, RC_BLAS_LEAF_MAX);
    {   // bounding sphere of the instance-entry cull (k_gather_tris + k_read_root): centre of the scene bounds, farthest vertex
        f3 ctr = mk3(0.5f * (smin.x + smax.x), 0.5f * (smin.y + smax.y), 0.5f * (smin.z + smax.z));
        float r2 = 0.0f;
        for (uint32_t j = 0; j < n; j++) {
            const RcTri &t = B->tris[j];
            float f = rc_far2(ctr, mk3(t.v0[0], t.v0[1], t.v0[2]), mk3(t.v1[0], t.v1[1], t.v1[2]), mk3(t.v2[0], t.v2[1], t.v2[2]));
            r2 = f == f ? std::max(r2, f) : INFINITY;
        }
        B->sphere[0] = ctr.x; B->sphere[1] = ctr.y; B->sphere[2] = ctr.z; B->sphere[3] = r2 * 1.000002f;
    }
    return B;
}

uint32_t hs_blas_n(void *b) { return ((HsBlas *)b)->tree.n; }
void hs_blas_nodes2(void *b, rc_bvh_node2 *out) {
    HsBlas *B = (HsBlas *)b;
    for (size_t i = 0; i < B->tree.nodes2.size(); i++) memcpy(&out[i], &B->tree.nodes2[i], sizeof(rc_bvh_node2));
}
void hs_blas_root(void *b, float *out6) { memcpy(out6, ((HsBlas *)b)->tree.root, 24); }
// the wide nodes (n + 1 slots of 64 bytes, slot 0 unused): what k_fit_local / k_collapse_span must reproduce byte for byte
void hs_blas_nodes4(void *b, void *out) {
    HsBlas *B = (HsBlas *)b;
    memcpy(out, B->tree.nodes4.data(), B->tree.nodes4.size() * sizeof(RcNode4));
}
void hs_blas_order(void *b, uint32_t *out) {
    HsBlas *B = (HsBlas *)b;
    for (size_t i = 0; i < B->tris.size(); i++) out[i] = B->tris[i].prim_id;
}
void hs_blas_free(void *b) { delete (HsBlas *)b; }

void hs_mat3x4_inverse(const float *m, float *out) { x_mat3x4_inverse(m, out); }

void *hs_scene_build(void **blas, uint32_t n_blas, const rc_instance_desc *inst, uint32_t n) {
    HsScene *S = new HsScene();
    for (uint32_t i = 0; i < n_blas; i++) S->blas.push_back((HsBlas *)blas[i]);
    S->inst.assign(inst, inst + n);
    S->rec.resize(n);
    S->aux.resize(n);
    std::vector<RcBox> inst_boxes(n);
    uint32_t bounds[6];
    for (int c = 0; c < 3; c++) { bounds[c] = rc_float_to_ordered(INFINITY); bounds[3 + c] = rc_float_to_ordered(-INFINITY); }
    for (uint32_t i = 0; i < n; i++) {  // k_instance_records + k_instance_boxes
        HsBlas *B = S->blas[inst[i].blas_index - 1];
        memcpy(S->rec[i].inv, inst[i].inv_transform, 48);
        S->rec[i].nodes4 = B->tree.nodes4.data();
        S->rec[i].tris = B->tris.data();
        memcpy(S->rec[i].sphere, B->sphere, 16);
        rc_world_sphere(inst[i].transform, B->sphere, S->rec[i].wsphere);
        S->aux[i].nodes2 = B->tree.nodes2.data();
        S->aux[i].n_prims = B->tree.n;
        S->aux[i].custom_index = inst[i].instance_id;
        f3 lo, hi;
        rc_instance_world_aabb(inst[i].transform, B->tree.root, lo, hi);
        set_box(inst_boxes[i], lo, hi);
        float l[3] = {lo.x, lo.y, lo.z}, h[3] = {hi.x, hi.y, hi.z};
        for (int k = 0; k < 3; k++) {
            bounds[k] = std::min(bounds[k], rc_float_to_ordered(l[k]));
            bounds[3 + k] = std::max(bounds[3 + k], rc_float_to_ordered(h[k]));
        }
    }
    if (n > 0) {
        f3 smin = mk3(rc_ordered_to_float(bounds[0]), rc_ordered_to_float(bounds[1]), rc_ordered_to_float(bounds[2]));
        f3 smax = mk3(rc_ordered_to_float(bounds[3]), rc_ordered_to_float(bounds[4]), rc_ordered_to_float(bounds[5]));
        f3 ext = mk3(jl_max(x_sub(smax.x, smin.x), 1e-6f), jl_max(x_sub(smax.y, smin.y), 1e-6f), jl_max(x_sub(smax.z, smin.z), 1e-6f));
        std::vector<uint32_t> codes(n), idx(n);
        for (uint32_t i = 0; i < n; i++) {  // k_morton_instances
            const float *la = S->blas[inst[i].blas_index - 1]->tree.root;
            f3 lc = mk3(x_mul(0.5f, x_add(la[0], la[3])), x_mul(0.5f, x_add(la[1], la[4])), x_mul(0.5f, x_add(la[2], la[5])));
            f3 wc = x_transform_point(inst[i].transform, lc);
            f3 nrm = mk3(x_div(x_sub(wc.x, smin.x), ext.x), x_div(x_sub(wc.y, smin.y), ext.y), x_div(x_sub(wc.z, smin.z), ext.z));
            codes[i] = rc_morton30(nrm);
            idx[i] = i;
        }
        stable_sort_pairs(codes, idx);
        S->tlas.leaf_map = idx;
        build_tree(S->tlas, codes, nullptr, &inst_boxes, 1);
    }
    S->scene.tlas4 = S->tlas.nodes4.data();
    S->scene.tlas2 = S->tlas.nodes2.data();
    S->scene.inst = S->rec.data();
    S->scene.aux = S->aux.data();
    S->scene.n_instances = n;
    return S;
}

uint32_t hs_scene_tlas_nodes2(void *s, rc_bvh_node2 *out) {
    HsScene *S = (HsScene *)s;
    if (out)
        for (size_t i = 0; i < S->tlas.nodes2.size(); i++) memcpy(&out[i], &S->tlas.nodes2[i], sizeof(rc_bvh_node2));
    return (uint32_t)S->tlas.nodes2.size();
}
void hs_scene_root(void *s, float *out6) { memcpy(out6, ((HsScene *)s)->tlas.root, 24); }
void hs_scene_free(void *s) { delete (HsScene *)s; }

// returns number of rays whose traversal stack overflowed
// the same per-ray bodies with the watertight triangle test (RC_MODE_WATERTIGHT)
uint32_t hs_trace_watertight(void *s, const rc_ray *rays, rc_hit *hits, uint64_t n, int any, int wide) {
    HsScene *S = (HsScene *)s;
    uint32_t overflow = 0;
    for (uint64_t i = 0; i < n; i++) {
        bool ok;
        if (wide) ok = any ? rc_trace_wide<true, false, true>(S->scene, rays[i], hits[i], nullptr) : rc_trace_wide<false, false, true>(S->scene, rays[i], hits[i], nullptr);
        else ok = any ? rc_trace_reference_order<true, false, true>(S->scene, rays[i], hits[i], nullptr) : rc_trace_reference_order<false, false, true>(S->scene, rays[i], hits[i], nullptr);
        if (!ok) overflow++;
    }
    return overflow;
}

uint32_t hs_trace(void *s, const rc_ray *rays, rc_hit *hits, uint64_t n, int any, int wide, uint64_t *counters /* nullable, 5 */) {
    HsScene *S = (HsScene *)s;
    uint32_t overflow = 0;
    RcLocalCounters lc = {0, 0, 0, 0, 0};
    uint64_t acc[5] = {0, 0, 0, 0, 0};
    for (uint64_t i = 0; i < n; i++) {
        bool ok;
        lc = RcLocalCounters{0, 0, 0, 0, 0};
        if (wide) ok = any ? rc_trace_wide<true, true>(S->scene, rays[i], hits[i], &lc) : rc_trace_wide<false, true>(S->scene, rays[i], hits[i], &lc);
        else ok = any ? rc_trace_reference_order<true, true>(S->scene, rays[i], hits[i], &lc) : rc_trace_reference_order<false, true>(S->scene, rays[i], hits[i], &lc);
        if (!ok) overflow++;
        acc[0] += lc.nodes; acc[1] += lc.box_tests; acc[2] += lc.tri_tests; acc[3] += lc.inst_entries;
        if (lc.max_stack > acc[4]) acc[4] = lc.max_stack;
    }
    if (counters) memcpy(counters, acc, sizeof acc);
    return overflow;
}

// wide-node sanity: every child box decoded from node k contains the exact BVH2 box of that child; returns violations
uint32_t hs_check_wide(void *b) {
    HsBlas *B = (HsBlas *)b;
    HsTree &t = B->tree;
    uint32_t bad = 0, n = t.n;
    std::vector<uint32_t> todo{1};
    while (!todo.empty()) {
        uint32_t k = todo.back();
        todo.pop_back();
        const RcNode4 &nd = t.nodes4[k];
        const float k24 = 5.9604644775390625e-8f;
        float sc[3] = {nd.sx * k24, nd.sy * k24, nd.sz * k24};
        float org[3] = {nd.ox, nd.oy, nd.oz};
        uint32_t ql[3] = {nd.qlox, nd.qloy, nd.qloz}, qh[3] = {nd.qhix, nd.qhiy, nd.qhiz};
        uint32_t ch[4] = {nd.child0, nd.child1, nd.child2, nd.child3};
        for (int c = 0; c < 4; c++) {
            if (c > 0 && ch[c] == ch[0] && ((ql[0] >> (8 * c)) & 0xFF) == 255 && ((qh[0] >> (8 * c)) & 0xFF) == 0) continue;  // unused slot
            if (ch[c] == RC_INVALID) { bad++; continue; }
            float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
            if (ch[c] & RC_LEAF_BIT) {
                uint32_t start = ch[c] & RC_LEAF_START_MASK, cnt = ((ch[c] >> RC_LEAF_COUNT_SHIFT) & 7) + 1;
                if (start + cnt > n) { bad++; continue; }
                for (uint32_t q = 0; q < cnt; q++) {
                    const RcTri &tr = B->tris[start + q];
                    for (int a = 0; a < 3; a++) {
                        lo[a] = std::min(lo[a], std::min(tr.v0[a], std::min(tr.v1[a], tr.v2[a])));
                        hi[a] = std::max(hi[a], std::max(tr.v0[a], std::max(tr.v1[a], tr.v2[a])));
                    }
                }
            } else {
                for (int a = 0; a < 3; a++) { lo[a] = t.boxes[ch[c] - 1].lo[a]; hi[a] = t.boxes[ch[c] - 1].hi[a]; }
                todo.push_back(ch[c]);
            }
            for (int a = 0; a < 3; a++) {
                float dl = fmaf((float)((ql[a] >> (8 * c)) & 0xFF), sc[a], org[a]), dh = fmaf((float)((qh[a] >> (8 * c)) & 0xFF), sc[a], org[a]);
                if (!(dl <= lo[a]) || !(dh >= hi[a])) bad++;
            }
        }
    }
    return bad;
}

// structural check used on imported blobs (rc_validate_blas_elem; the loop stands in for k_validate_blas).  corrupt: 0 none,
// 1 wide-node child index past the last node, 2 leaf range past the triangle array, 3 BVH2 child out of range, 4 BVH2 leaf
// primitive out of range, 5 triangle prim_id out of range, 6 TLAS-tagged reference inside a BLAS, 7 wide node naming itself as a child
// (a cycle: the traversal would never end), 8 BVH2 node naming itself as a child, 9 triangle face_index past the submitted faces,
// 10 a reference to the root, 11 one wide node referenced twice, 12 a referenced wide slot that is empty.  `where` picks the element.
uint32_t hs_validate_blas(void *b, int corrupt, uint32_t where) {
    HsBlas *B = (HsBlas *)b;
    HsTree &t = B->tree;
    const uint32_t n = t.n, last = n > 1 ? n - 1 : 1;
    std::vector<RcNode2> nodes2 = t.nodes2;
    std::vector<RcNode4> nodes4 = t.nodes4;
    std::vector<RcTri> tris = B->tris;
    // wide-node faults are injected into a node the traversal can reach (slots nothing points at may hold anything: they are never read)
    std::vector<uint32_t> reach{1};
    for (size_t q = 0; q < reach.size() && n > 1; q++) {
        const uint32_t c[4] = {nodes4[reach[q]].child0, nodes4[reach[q]].child1, nodes4[reach[q]].child2, nodes4[reach[q]].child3};
        for (int k = 0; k < 4; k++)
            if (!(c[k] & RC_LEAF_BIT) && (k == 0 || c[k] != c[0])) reach.push_back(c[k]);
    }
    const uint32_t w = reach[where % reach.size()];
    switch (corrupt) {
        case 1: nodes4[w].child1 = last + 1; break;
        case 2: nodes4[w].child0 = RC_LEAF_BIT | ((RC_BLAS_LEAF_MAX - 1u) << RC_LEAF_COUNT_SHIFT) | (n - RC_BLAS_LEAF_MAX + 1u); break;
        case 3: if (n > 1) nodes2[where % (n - 1)].child1 = 2 * n; else nodes2[0].child0 = 1; break;
        case 4: nodes2[n - 1 + where % n].child1 = n + 1; break;
        case 5: tris[where % n].prim_id = n; break;
        case 6: nodes4[w].child2 = RC_TLAS_LEAF_TAG | 0u; break;
        case 7: nodes4[w].child1 = w; break;
        case 8: if (n > 1) nodes2[where % (n - 1)].child0 = 1 + where % (n - 1); else nodes2[0].child0 = 1; break;
        case 9: tris[where % n].face_index = B->n_faces_in; break;
        case 10: nodes4[w].child3 = 1; break;  // nothing may point at the root
        case 11: {  // two nodes naming the same child (in-degree 2: the shape a cycle reachable from the root needs)
            if (reach.size() > 2) nodes4[reach[1]].child1 = reach.back() == reach[1] ? reach[2] : reach.back(); else nodes4[1].child1 = 1;
            break;
        }
        case 12: {  // a referenced slot that is empty
            uint32_t c = nodes4[1].child0;
            if (!(c & RC_LEAF_BIT)) memset(&nodes4[c], 0, sizeof(RcNode4)); else nodes4[1].child0 = last + 7;
            break;
        }
        default: break;
    }
    uint32_t bad = 0;
    for (uint32_t i = 0; i < 2 * n; i++) bad += rc_validate_static_elem(i, nodes2.data(), tris.data(), n, B->n_faces_in);
    std::vector<uint32_t> mark(n + 1, 0u);
    mark[1] = 1;
    uint32_t marked = 1;
    for (uint32_t level = 1; marked != 0 && bad == 0 && level <= n + 1; level++) {  // the loop of rc_blas_import
        marked = 0;
        for (uint32_t i = 1; i <= last; i++) bad += rc_validate_wide_level(i, level, nodes4.data(), n, RC_BLAS_LEAF_MAX, mark.data(), &marked);
    }
    return bad;
}

// ---- wavefront stage bodies (rc_wave_core.cuh); loops stand in for k_primary_rays / k_geometric_normals / k_shadow_rays
void hs_primary_rays(const float *camera_pos, const float *right, const float *up, const float *forward, float half_width, float half_height, int lookat, int jitter,
                     uint32_t width, uint32_t height, uint32_t n_samples, uint64_t seed, rc_ray *out) {
    RcCamera cam;
    memset(&cam, 0, sizeof cam);
    memcpy(cam.pos, camera_pos, 12);
    if (right) memcpy(cam.right, right, 12);
    if (up) memcpy(cam.up, up, 12);
    memcpy(cam.forward, forward, 12);
    cam.half_width = half_width; cam.half_height = half_height;
    cam.lookat = (uint32_t)lookat; cam.jitter = (uint32_t)jitter;
    uint64_t total = (uint64_t)width * height * n_samples;
    for (uint64_t i = 0; i < total; i++) out[i] = rc_primary_ray(cam, width, height, n_samples, seed, i);
}

// blas_normals: table with one entry per BLAS (NULL entry => geometric normals), 9 floats per primitive indexed by primitive_id
void hs_shadow_rays(void *s, const rc_ray *rays, const rc_hit *hits, uint64_t n, const float *const *blas_normals, const float *lights, uint32_t n_lights,
                    float bias, rc_ray *out) {
    HsScene *S = (HsScene *)s;
    std::vector<std::vector<float>> geo(S->blas.size());
    for (uint64_t k = 0; k < n; k++) {
        for (uint32_t l = 0; l < n_lights; l++) {
            if (!hits[k].hit) { out[k * n_lights + l] = rc_dummy_shadow_ray(); continue; }
            const rc_instance_desc &inst = S->inst[hits[k].instance_id];
            uint32_t b = inst.blas_index - 1;
            const float *nrm = blas_normals ? blas_normals[b] : nullptr;
            if (!nrm) {
                if (geo[b].empty()) {
                    HsBlas *B = S->blas[b];
                    geo[b].resize(9 * (size_t)B->tree.n);
                    for (uint32_t p = 0; p < B->tree.n; p++) {
                        f3 g = rc_geometric_normal(B->tris[p]);
                        float *o = geo[b].data() + 9 * (size_t)B->tris[p].prim_id;
                        for (int q = 0; q < 3; q++) { o[3 * q] = g.x; o[3 * q + 1] = g.y; o[3 * q + 2] = g.z; }
                    }
                }
                nrm = geo[b].data();
            }
            out[k * n_lights + l] = rc_shadow_ray(rays[k], hits[k], nrm + 9 * (size_t)hits[k].primitive_id, inst.inv_transform, lights + 3 * l, bias);
        }
    }
}

}  // extern "C"

// ---- k_trace_wide (rc_trace_fast.cuh) on the CPU: n_warps persistent warps of 32 fibres share one work counter, as the CTAs of the
// real launch do.  Rays whose short stack overflowed are re-traced by the deep-stack generic body, as k_trace_fixup does.
namespace warpsim {
Warp *g_warp = nullptr;
int g_lane = 0;
Tid g_tid = {0};
ucontext_t g_sched;
uint64_t g_exchanges = 0;
uint64_t g_step_iters[4] = {0, 0, 0, 0}, g_step_lanes[4] = {0, 0, 0, 0};
uint64_t g_idle[6] = {0, 0, 0, 0, 0, 0};
struct Launch {
    RcScene sc;
    RcIoArrays io;
    unsigned long long n;
    unsigned long long *work;
    RcCounters *counters;
    uint32_t *overflow;
    int any, count;
};
static Launch g_launch;
static void lane_main() {
    const Launch &L = g_launch;
    const bool single = L.sc.n_instances == 1u;  // the launcher's choice of variant (rc_trace.cu)
#define WS_RUN(A, C)                                                                                      \
    {                                                                                                     \
        if (single) k_trace_wide<A, C, RcIoArrays, true>(L.sc, L.io, L.n, L.work, L.counters, L.overflow);  \
        else k_trace_wide<A, C, RcIoArrays, false>(L.sc, L.io, L.n, L.work, L.counters, L.overflow);        \
    }
    if (L.any) { if (L.count) WS_RUN(true, true) else WS_RUN(true, false) }
    else { if (L.count) WS_RUN(false, true) else WS_RUN(false, false) }
#undef WS_RUN
    g_warp->lane[g_lane].done = true;
    swapcontext(&g_warp->lane[g_lane].ctx, &g_sched);
}
// returns 0, or 1 when the lanes of a warp did not leave the kernel together (a divergence bug around a warp intrinsic)
static int run(std::vector<Warp> &warps) {
    const size_t stack_bytes = 512 * 1024;
    for (size_t w = 0; w < warps.size(); w++) {
        warps[w].first_tid = 32u * (unsigned)(w % (RC_TRACE_THREADS / 32));
        for (int l = 0; l < 32; l++) {
            Lane &ln = warps[w].lane[l];
            ln.stack.resize(stack_bytes);
#pragma GCC diagnostic push
#pragma GCC diagnostic ignored "-Wmaybe-uninitialized"  // getcontext *writes* the context; gcc 13 flags the argument as read
            getcontext(&ln.ctx);
#pragma GCC diagnostic pop
            ln.ctx.uc_stack.ss_sp = ln.stack.data();
            ln.ctx.uc_stack.ss_size = stack_bytes;
            ln.ctx.uc_link = &g_sched;
            makecontext(&ln.ctx, lane_main, 0);
        }
    }
    for (;;) {
        size_t live = 0;
        for (auto &w : warps) {
            int done = 0;
            for (int l = 0; l < 32; l++) done += w.lane[l].done ? 1 : 0;
            if (done == 32) continue;
            if (done != 0) return 1;
            live++;
            for (int l = 0; l < 32; l++) {
                g_warp = &w;
                g_lane = l;
                g_tid.x = w.first_tid + (unsigned)l;
                swapcontext(&g_sched, &w.lane[l].ctx);
            }
        }
        if (!live) return 0;
    }
}
}  // namespace warpsim

extern "C" {
// out_info (nullable, 18): rays flagged by the short stack, rays no stack could hold, warp exchanges executed, lock-step violation,
// then per step kind (N, T, X, F) the warp iterations [4..7] and the active lanes summed over them [8..11], then the lanes that sat
// out node steps by reason [12..17] (see g_idle)
uint32_t hs_trace_warpsim(void *s, const rc_ray *rays, rc_hit *hits, uint64_t n, int any, uint32_t n_warps, uint64_t *counters /* nullable, 6 */,
                          uint64_t *out_info) {
    HsScene *S = (HsScene *)s;
    if (out_info) memset(out_info, 0, 18 * sizeof(uint64_t));
    if (S->scene.n_instances == 0) {  // rc_launch_trace: an empty TLAS never reaches the kernel
        for (uint64_t i = 0; i < n; i++) rc_write_miss(hits[i]);
        return 0;
    }
    unsigned long long work = 0;
    RcCounters cnt;
    memset(&cnt, 0, sizeof cnt);
    uint32_t overflow[4] = {0, 0, 0, 0};
    warpsim::g_launch = warpsim::Launch{S->scene, RcIoArrays{rays, hits}, n, &work, &cnt, overflow, any, counters != nullptr};
    warpsim::g_exchanges = 0;
    for (int k = 0; k < 4; k++) warpsim::g_step_iters[k] = warpsim::g_step_lanes[k] = 0;
    for (int k = 0; k < 6; k++) warpsim::g_idle[k] = 0;
    std::vector<warpsim::Warp> warps(n_warps ? n_warps : 1);
    const int bad = warpsim::run(warps);
    uint32_t hard = 0;
    for (uint64_t i = 0; i < n && !bad; i++) {  // k_trace_fixup
        if (hits[i].hit != RC_OVERFLOW_MARK) continue;
        bool ok = any ? rc_trace_wide<true, false>(S->scene, rays[i], hits[i], nullptr) : rc_trace_wide<false, false>(S->scene, rays[i], hits[i], nullptr);
        if (!ok) hard++;
    }
    if (counters) { counters[0] = cnt.rays; counters[1] = cnt.nodes; counters[2] = cnt.box_tests; counters[3] = cnt.tri_tests; counters[4] = cnt.inst_entries; counters[5] = cnt.max_stack; }
    if (out_info) {
        out_info[0] = overflow[0]; out_info[1] = hard; out_info[2] = warpsim::g_exchanges; out_info[3] = (uint64_t)bad;
        for (int k = 0; k < 4; k++) { out_info[4 + k] = warpsim::g_step_iters[k]; out_info[8 + k] = warpsim::g_step_lanes[k]; }
        for (int k = 0; k < 6; k++) out_info[12 + k] = warpsim::g_idle[k];
    }
    return bad ? 0xFFFFFFFFu : hard;
}
}  // extern "C"
