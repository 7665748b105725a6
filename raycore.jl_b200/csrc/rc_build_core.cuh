// rc_build_core.cuh — per-element bodies of the LBVH builder (Karras radix tree, bottom-up fit,
// BVH2 -> quantised BVH4 collapse).  RC_HD so tests/hostsim can run them on the CPU.
//
// Index conventions follow the reference (src/instanced-bvh.jl:1293-1295): nodes are numbered
// 1..2n-1, internal 1..n-1, leaves n..2n-1 (leaf of sorted primitive p (1-based) = n-1+p), root = 1.
// Arrays indexed by node number are stored 0-based ([node-1]) except the wide-node array, which is
// indexed directly by the BVH2 internal-node number (slot 0 unused, root = slot 1).
#pragma once
#include "rc_device.cuh"

struct RcBox {  // own AABB of a BVH2 node, 32 B
    float lo[3], pad0;
    float hi[3], pad1;
};

struct RcTopo {  // per internal node
    uint32_t child0, child1;  // 1-based node numbers
    uint32_t span_lo, span_hi;  // 1-based sorted primitive range covered
};

// Karras 2012 as restated by the reference: find_span_for_node :1232-1262, find_split_in_span :1265-1290,
// build_topology_for_node (kernels.jl:119-143)
RC_HD RcTopo rc_topology_for_node(int idx, const uint32_t *codes, int n) {
    int d_left = rc_delta(idx, idx - 1, codes, n);
    int d_right = rc_delta(idx, idx + 1, codes, n);
    int d = d_right > d_left ? 1 : -1;
    int delta_min = rc_delta(idx, idx - d, codes, n);
    int l_max = 2;
    while (rc_delta(idx, idx + l_max * d, codes, n) > delta_min) l_max *= 2;
    int l = 0, t = l_max;
    while (t > 1) {
        t = t / 2;
        if (rc_delta(idx, idx + (l + t) * d, codes, n) > delta_min) l = l + t;
    }
    int j = idx + l * d;
    int span_left = d > 0 ? idx : j, span_right = d > 0 ? j : idx;
    int numidentical = rc_delta(span_left, span_right, codes, n);
    int left = span_left, right = span_right;
    while (right > left + 1) {
        int newsplit = (right + left) / 2;
        if (rc_delta(left, newsplit, codes, n) > numidentical) left = newsplit;
        else right = newsplit;
    }
    int split = left;
    RcTopo r;
    r.child0 = (uint32_t)((split == span_left) ? (n - 1 + split) : split);
    int c1 = split + 1;
    r.child1 = (uint32_t)((c1 == span_right) ? (n - 1 + c1) : c1);
    r.span_lo = (uint32_t)span_left;
    r.span_hi = (uint32_t)span_right;
    return r;
}

RC_HD float rc_half_area(const RcBox &b) {
    float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
    return dx * dy + dy * dz + dz * dx;
}

// squared distance from c to the farthest of three vertices (bounding-sphere radius of the instance-entry cull)
RC_HD float rc_far2(f3 c, f3 a, f3 b, f3 v) {
    const float da = (a.x - c.x) * (a.x - c.x) + (a.y - c.y) * (a.y - c.y) + (a.z - c.z) * (a.z - c.z);
    const float db = (b.x - c.x) * (b.x - c.x) + (b.y - c.y) * (b.y - c.y) + (b.z - c.z) * (b.z - c.z);
    const float dv = (v.x - c.x) * (v.x - c.x) + (v.y - c.y) * (v.y - c.y) + (v.z - c.z) * (v.z - c.z);
    return fmaxf(da, fmaxf(db, dv));
}

// Smallest biased exponent e with 255 * 2^(e-127) >= extent
RC_HD uint32_t rc_quant_exponent(float extent) {
    if (!(extent > 0.0f)) return 1u;
    float s = extent / 255.0f;
    uint32_t bits = f2u(s);
    uint32_t e = (bits >> 23) & 0xFFu;
    if (bits & 0x7FFFFFu) e += 1;
    if (e < 1u) e = 1u;
    if (e > RC_QUANT_EXP_MAX) e = RC_QUANT_EXP_MAX;
    // guard the rounding of extent/255: make sure 255 * scale really covers the extent
    while (e < RC_QUANT_EXP_MAX && u2f(e << 23) * 255.0f < extent) e += 1;
    return e;
}

// Conservative 8-bit plane codes: decoded lo plane <= v (floor), decoded hi plane >= v (ceil).  scale = 2^(e-127) with
// 1 <= e <= RC_QUANT_EXP_MAX, so x / scale == x * 2^(127-e) bit for bit (an exact scaling, one rounding either way, and the
// reciprocal is a normal float): a multiply instead of an IEEE division (24 of them per wide node were 13 % of k_fit_local).
RC_HD float rc_quant_inv_scale(float scale) { return u2f((254u << 23) - f2u(scale)); }
RC_HD uint32_t rc_quant_lo(float v, float origin, float scale) {
    float q = floorf((v - origin) * rc_quant_inv_scale(scale));
    q = q < 0.0f ? 0.0f : (q > 255.0f ? 255.0f : q);
    while (q > 0.0f && fmaf(q, scale, origin) > v) q -= 1.0f;
    return (uint32_t)q;
}
RC_HD uint32_t rc_quant_hi(float v, float origin, float scale) {
    float q = ceilf((v - origin) * rc_quant_inv_scale(scale));
    q = q < 0.0f ? 0.0f : (q > 255.0f ? 255.0f : q);
    while (q < 255.0f && fmaf(q, scale, origin) < v) q += 1.0f;
    return (uint32_t)q;
}

// Collapse the BVH2 subtree rooted at internal node `idx` into one wide node.
//   boxes/topo: BVH2 arrays; n: primitive count; leaf_max: triangles per wide leaf;
//   leaf_map (nullable): sorted position -> payload index (TLAS: instance index), only with leaf_max == 1.
// A BVH2 node covering <= leaf_max primitives becomes a leaf reference, otherwise the child with the
// largest surface area is opened until four slots are used (greedy; precedent: src/bvh4.jl:234-277).
// box_of(c) / topo_of(c): own box of BVH2 node c / topology record of internal node c (1-based numbers) — global arrays in k_collapse_span,
// the block's shared-memory copies in k_fit_local.
template <class BoxOf, class TopoOf>
RC_HD RcNode4 rc_collapse_node_t(uint32_t idx, BoxOf box_of, TopoOf topo_of, uint32_t n, uint32_t leaf_max, const uint32_t *leaf_map) {
    uint32_t slots[4];
    int ns = 0;
    auto count_of = [&](uint32_t c) -> uint32_t {
        if (c >= n) return 1u;
        const RcTopo t = topo_of(c);
        return t.span_hi - t.span_lo + 1u;
    };
    const RcBox own = box_of(idx);
    uint32_t own_count = count_of(idx);
    if (idx >= n || own_count <= leaf_max) {
        slots[ns++] = idx;  // degenerate root: whole BLAS is one leaf
    } else {
        {
            const RcTopo t = topo_of(idx);
            slots[ns++] = t.child0;
            slots[ns++] = t.child1;
        }
        while (ns < 4) {
            int best = -1;
            float best_area = -1.0f;
            for (int k = 0; k < ns; k++) {
                uint32_t c = slots[k];
                if (c < n && count_of(c) > leaf_max) {
                    float a = rc_half_area(box_of(c));
                    if (a > best_area) { best_area = a; best = k; }
                }
            }
            if (best < 0) break;
            const RcTopo t = topo_of(slots[best]);
            slots[best] = t.child0;
            slots[ns++] = t.child1;
        }
    }
    RcNode4 nd;
    nd.ox = own.lo[0]; nd.oy = own.lo[1]; nd.oz = own.lo[2];
    uint32_t ex = rc_quant_exponent(own.hi[0] - own.lo[0]);
    uint32_t ey = rc_quant_exponent(own.hi[1] - own.lo[1]);
    uint32_t ez = rc_quant_exponent(own.hi[2] - own.lo[2]);
    nd.sx = u2f((ex + 24u) << 23); nd.sy = u2f((ey + 24u) << 23); nd.sz = u2f((ez + 24u) << 23);
    float sx = u2f(ex << 23), sy = u2f(ey << 23), sz = u2f(ez << 23);
    uint32_t qlo[3] = {0, 0, 0}, qhi[3] = {0, 0, 0}, ch[4];
    for (int k = 0; k < 4; k++) {
        if (k >= ns) {  // unused slot: inverted box + child 0's reference (filled in below)
            for (int a = 0; a < 3; a++) { qlo[a] |= 255u << (8 * k); }
            continue;
        }
        uint32_t c = slots[k];
        const RcBox b = box_of(c);
        qlo[0] |= rc_quant_lo(b.lo[0], nd.ox, sx) << (8 * k);
        qlo[1] |= rc_quant_lo(b.lo[1], nd.oy, sy) << (8 * k);
        qlo[2] |= rc_quant_lo(b.lo[2], nd.oz, sz) << (8 * k);
        qhi[0] |= rc_quant_hi(b.hi[0], nd.ox, sx) << (8 * k);
        qhi[1] |= rc_quant_hi(b.hi[1], nd.oy, sy) << (8 * k);
        qhi[2] |= rc_quant_hi(b.hi[2], nd.oz, sz) << (8 * k);
        uint32_t cnt = count_of(c);
        if (c >= n) {
            uint32_t pos = c - n;  // 0-based sorted position
            ch[k] = leaf_map ? (RC_TLAS_LEAF_TAG | leaf_map[pos]) : (RC_LEAF_BIT | pos);
        } else if (cnt <= leaf_max) {
            uint32_t pos = topo_of(c).span_lo - 1u;
            ch[k] = RC_LEAF_BIT | ((cnt - 1u) << RC_LEAF_COUNT_SHIFT) | (leaf_map ? leaf_map[pos] : pos);
        } else {
            ch[k] = c;
        }
    }
    nd.qlox = qlo[0]; nd.qloy = qlo[1]; nd.qloz = qlo[2];
    nd.qhix = qhi[0]; nd.qhiy = qhi[1]; nd.qhiz = qhi[2];
    for (int k = ns; k < 4; k++) ch[k] = ch[0];
    nd.child0 = ch[0]; nd.child1 = ch[1]; nd.child2 = ch[2]; nd.child3 = ch[3];
    return nd;
}

RC_HD RcNode4 rc_collapse_node(uint32_t idx, const RcBox *boxes, const RcTopo *topo, uint32_t n, uint32_t leaf_max, const uint32_t *leaf_map) {
    return rc_collapse_node_t(
        idx, [&](uint32_t c) -> RcBox { return boxes[c - 1]; }, [&](uint32_t c) -> RcTopo { return topo[c - 1]; }, n, leaf_max, leaf_map);
}

// rc_collapse_node with every BVH2 record fetched once and the fetches of a step independent of each other: a slot carries the count, box,
// children and first position of its node, so opening a slot is ONE round of loads (its two children) instead of a chain of ~10
// dependent ones.  For the global-memory callers (k_collapse_span: a spanning node was ~40 dependent L2 round trips); same selection
// order and arithmetic, bit-identical result.  No dynamically indexed arrays (they would live in local memory).
struct RcCollapseSlot {
    uint32_t node, count, c0, c1, first;  // first: 0-based sorted position of the first primitive
    RcBox box;
};
RC_HD RcCollapseSlot rc_collapse_fetch(uint32_t c, const RcBox *boxes, const RcTopo *topo, uint32_t n) {
    RcCollapseSlot s;
    s.node = c;
    s.box = boxes[c - 1];
    if (c >= n) {
        s.count = 1u; s.c0 = 0u; s.c1 = 0u; s.first = c - n;
    } else {
        const RcTopo t = topo[c - 1];
        s.count = t.span_hi - t.span_lo + 1u; s.c0 = t.child0; s.c1 = t.child1; s.first = t.span_lo - 1u;
    }
    return s;
}
RC_HD RcNode4 rc_collapse_node_cached(uint32_t idx, const RcBox *boxes, const RcTopo *topo, uint32_t n, uint32_t leaf_max, const uint32_t *leaf_map) {
    const RcCollapseSlot own = rc_collapse_fetch(idx, boxes, topo, n);
    RcCollapseSlot sl[4];
    int ns = 0;
    if (idx >= n || own.count <= leaf_max) {
        sl[0] = own;  // degenerate root: whole BLAS is one leaf
        ns = 1;
    } else {
        sl[0] = rc_collapse_fetch(own.c0, boxes, topo, n);
        sl[1] = rc_collapse_fetch(own.c1, boxes, topo, n);
        ns = 2;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int stage = 2; stage < 4; stage++) {  // open the slot with the largest area until four are used
            if (ns != stage) break;
            int best = -1;
            float best_area = -1.0f;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int k = 0; k < 3; k++) {
                if (k < stage && sl[k].node < n && sl[k].count > leaf_max) {
                    const float a = rc_half_area(sl[k].box);
                    if (a > best_area) { best_area = a; best = k; }
                }
            }
            if (best < 0) break;
            uint32_t c0 = 0, c1 = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int k = 0; k < 3; k++)
                if (k == best) { c0 = sl[k].c0; c1 = sl[k].c1; }
            const RcCollapseSlot a = rc_collapse_fetch(c0, boxes, topo, n), b = rc_collapse_fetch(c1, boxes, topo, n);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int k = 0; k < 3; k++)
                if (k == best) sl[k] = a;
            sl[stage] = b;
            ns = stage + 1;
        }
    }
    RcNode4 nd;
    nd.ox = own.box.lo[0]; nd.oy = own.box.lo[1]; nd.oz = own.box.lo[2];
    const uint32_t ex = rc_quant_exponent(own.box.hi[0] - own.box.lo[0]);
    const uint32_t ey = rc_quant_exponent(own.box.hi[1] - own.box.lo[1]);
    const uint32_t ez = rc_quant_exponent(own.box.hi[2] - own.box.lo[2]);
    nd.sx = u2f((ex + 24u) << 23); nd.sy = u2f((ey + 24u) << 23); nd.sz = u2f((ez + 24u) << 23);
    const float sx = u2f(ex << 23), sy = u2f(ey << 23), sz = u2f(ez << 23);
    uint32_t qlo[3] = {0, 0, 0}, qhi[3] = {0, 0, 0}, ch[4] = {0, 0, 0, 0};
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 4; k++) {
        if (k >= ns) {  // unused slot: inverted box + child 0's reference (filled in below)
            for (int a = 0; a < 3; a++) qlo[a] |= 255u << (8 * k);
            continue;
        }
        const RcCollapseSlot &c = sl[k];
        qlo[0] |= rc_quant_lo(c.box.lo[0], nd.ox, sx) << (8 * k);
        qlo[1] |= rc_quant_lo(c.box.lo[1], nd.oy, sy) << (8 * k);
        qlo[2] |= rc_quant_lo(c.box.lo[2], nd.oz, sz) << (8 * k);
        qhi[0] |= rc_quant_hi(c.box.hi[0], nd.ox, sx) << (8 * k);
        qhi[1] |= rc_quant_hi(c.box.hi[1], nd.oy, sy) << (8 * k);
        qhi[2] |= rc_quant_hi(c.box.hi[2], nd.oz, sz) << (8 * k);
        if (c.node >= n) ch[k] = leaf_map ? (RC_TLAS_LEAF_TAG | leaf_map[c.first]) : (RC_LEAF_BIT | c.first);
        else if (c.count <= leaf_max) ch[k] = RC_LEAF_BIT | ((c.count - 1u) << RC_LEAF_COUNT_SHIFT) | (leaf_map ? leaf_map[c.first] : c.first);
        else ch[k] = c.node;
    }
    nd.qlox = qlo[0]; nd.qloy = qlo[1]; nd.qloz = qlo[2];
    nd.qhix = qhi[0]; nd.qhiy = qhi[1]; nd.qhiz = qhi[2];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 1; k < 4; k++)
        if (k >= ns) ch[k] = ch[0];
    nd.child0 = ch[0]; nd.child1 = ch[1]; nd.child2 = ch[2]; nd.child3 = ch[3];
    return nd;
}

// Structural check of a BLAS (used on imported blobs).  A blob that passes can neither send a traversal out of bounds nor into an
// endless loop.  Two parts:
//   rc_validate_static_elem(i), i = 0 .. 2n-1: triangle i has prim_id < n and face_index < n_faces_in (rc_set_normals gathers by it);
//       when the blob carries the reference-layout BVH2, node i+1 keeps to the numbering of src/instanced-bvh.jl:1293-1295 (internal
//       1..n-1, leaves n..2n-1 with child0 == INVALID_NODE and a 1-based primitive in child1), the two children of an internal node
//       are distinct and name it as their parent, and the root has none — so the 2n-2 child references reach 2n-2 distinct nodes and
//       everything reachable from node 1 is a tree.
//   rc_validate_wide_level(i, level), i = 1 .. max(1, n-1), level = 1, 2, ... until a level marks nothing: the wide nodes are checked
//       in breadth-first order from the root (mark[1] = 1 to start).  A node marked `level` must be a real node (the builder leaves
//       slots that head no wide node zeroed or stale; they are legal as long as nothing reachable points at them), its leaf ranges must
//       lie inside the triangle array, its node references inside 1..max(1, n-1), and every node it references must be unmarked so
//       far — a second reference to a node is what a cycle (or a DAG) reachable from the root needs, and is refused.
// Not checked (harmless for memory safety and termination): that the boxes bound their subtrees and that prim_id values are
// distinct — import blobs you wrote.
RC_HD uint32_t rc_validate_static_elem(uint32_t i, const RcNode2 *nodes2 /* nullable */, const RcTri *tris, uint32_t n, uint32_t n_faces_in) {
    uint32_t errs = 0;
    const uint32_t n_nodes2 = 2u * n - 1u;
    if (nodes2 && i < n_nodes2) {
        const RcNode2 &nd = nodes2[i];
        if (i == 0u && nd.parent != RC_INVALID) errs++;
        if (i + 1u < n) {
            if (nd.child0 < 1u || nd.child0 > n_nodes2 || nd.child1 < 1u || nd.child1 > n_nodes2 || nd.child0 == nd.child1) errs++;
            else if (nodes2[nd.child0 - 1u].parent != i + 1u || nodes2[nd.child1 - 1u].parent != i + 1u) errs++;
        } else if (nd.child0 != RC_INVALID || nd.child1 < 1u || nd.child1 > n) {
            errs++;
        }
    }
    if (i < n && (tris[i].prim_id >= n || tris[i].face_index >= n_faces_in)) errs++;
    return errs;
}
// returns violations; *marked counts the nodes this call marked for the next level
RC_HD uint32_t rc_validate_wide_level(uint32_t i, uint32_t level, const RcNode4 *nodes4, uint32_t n, uint32_t leaf_max, uint32_t *mark /* n + 1 words */,
                                      uint32_t *marked) {
    const uint32_t last = n > 1u ? n - 1u : 1u;
    if (i < 1u || i > last || mark[i] != level) return 0u;
    uint32_t errs = 0;
    const RcNode4 &w = nodes4[i];
    const uint32_t c[4] = {w.child0, w.child1, w.child2, w.child3};
    for (int k = 0; k < 4; k++) {
        if (k > 0 && c[k] == c[0]) continue;  // an unused slot repeats child 0's reference (rc_types.h): one edge, not two
        if (c[k] & RC_LEAF_BIT) {
            const uint32_t count = ((c[k] >> RC_LEAF_COUNT_SHIFT) & 7u) + 1u, start = c[k] & RC_LEAF_START_MASK;
            if ((c[k] & RC_TLAS_LEAF_TAG) == RC_TLAS_LEAF_TAG || count > leaf_max || start >= n || count > n - start) errs++;
        } else if (n == 1u || c[k] < 1u || c[k] > last) {
            errs++;  // (an empty slot fails here: reference 0)
        } else {
#if RC_ON_DEVICE
            const uint32_t old = atomicCAS(&mark[c[k]], 0u, level + 1u);
#else
            const uint32_t old = mark[c[k]];
            if (old == 0u) mark[c[k]] = level + 1u;
#endif
            if (old != 0u) errs++;
            else {
#if RC_ON_DEVICE
                atomicAdd(marked, 1u);
#else
                ++*marked;
#endif
            }
        }
    }
    return errs;
}

// world AABB of an instance: bounds of the 8 transformed corners of the BLAS root box
// (compute_instance_world_aabb, kernels.jl:38-62; corner order bounds.jl:53-59)
RC_HD void rc_instance_world_aabb(const float *xf, const float *local, f3 &mn, f3 &mx) {
    for (int c = 0; c < 8; c++) {
        f3 p = mk3((c & 1) ? local[3] : local[0], (c & 2) ? local[4] : local[1], (c & 4) ? local[5] : local[2]);
        f3 w = x_transform_point(xf, p);
        if (c == 0) { mn = w; mx = w; }
        else { mn = jl_min3(mn, w); mx = jl_max3(mx, w); }
    }
}
