// rc_trace_core.cuh — per-ray traversal bodies (RC_HD so tests/hostsim can run them on the CPU).
//
//   rc_trace_reference_order : the reference's own two-level BVH2 walk (closest_hit / any_hit,
//       src/instanced-bvh.jl:1902-2140) on the reference-identical BVH2, exact arithmetic,
//       same near/far rule, same tie behaviour -> bit-identical results.
//   rc_trace_wide            : two-level walk over the quantised BVH4 (default fast path).  Boxes are
//       conservative supersets and the slab test carries an explicit error bound, so no triangle the
//       exact Moeller-Trumbore test would accept is culled; the triangle test itself is the exact one,
//       hence t / u / v are bit-identical whenever the same triangle wins.
#pragma once
#include "rc_device.cuh"

#define RC_STACK_REF 128
#define RC_STACK_WIDE 96

struct rc_f4 {
    float x, y, z;
    uint32_t w;
};

RC_HD rc_f4 rc_load16(const void *p) {
    rc_f4 r;
#if RC_ON_DEVICE
    float4 v = __ldg(reinterpret_cast<const float4 *>(p));
    r.x = v.x; r.y = v.y; r.z = v.z; r.w = __float_as_uint(v.w);
#else
    const float *f = reinterpret_cast<const float *>(p);
    r.x = f[0]; r.y = f[1]; r.z = f[2];
    memcpy(&r.w, f + 3, 4);
#endif
    return r;
}

struct RcRayIn {  // sanitised world-space ray
    f3 o, d;
    float t_min, t_max;
};

RC_HD RcRayIn rc_prepare_ray(const rc_ray &r, bool any) {
    RcRayIn q;
    q.o = mk3(r.origin[0], r.origin[1], r.origin[2]);
    q.d = mk3(x_fix_zero(r.dir[0]), x_fix_zero(r.dir[1]), x_fix_zero(r.dir[2]));  // check_direction, src/ray.jl:39-49
    q.t_min = any ? 0.0f : r.tmin;  // any_hit ignores ray.t_min, src/instanced-bvh.jl:2039
    q.t_max = r.tmax;
    return q;
}

RC_HD void rc_write_miss(rc_hit &h) {
    h.hit = 0; h.t = 0.0f; h.primitive_id = 0; h.instance_custom_index = 0;
    h.bary_u = 0.0f; h.bary_v = 0.0f; h.instance_id = 0; h.metadata = 0;
}

struct RcLocalCounters {
    uint32_t nodes, box_tests, tri_tests, inst_entries, max_stack;
};

// ------------------------------------------------------------------------------------------------
// Reference-order traversal.  Returns false on stack overflow.
template <bool ANY, bool COUNT, bool WT = false>
RC_HD bool rc_trace_reference_order(const RcScene &sc, const rc_ray &ray, rc_hit &out, RcLocalCounters *cnt, const RcTri **tri_out = nullptr) {
    rc_write_miss(out);
    if (tri_out) *tri_out = nullptr;
    if (sc.n_instances == 0) return true;  // reference: UB on an empty TLAS; its tests require a miss (test_tlas_stress.jl:828)
    RcRayIn w = rc_prepare_ray(ray, ANY);
    f3 ray_o = w.o, ray_d = w.d;
    float ray_mint = w.t_min, ray_maxt = w.t_max;
    f3 inv = mk3(x_safe_inv(ray_d.x), x_safe_inv(ray_d.y), x_safe_inv(ray_d.z));
    uint32_t stack[RC_STACK_REF];
    int sp = 1;
    stack[0] = RC_INVALID;
    int current_instance = -1, closest_instance = -1;
    uint32_t closest_prim = RC_INVALID;
    float hit_u = 0.0f, hit_v = 0.0f;
    uint32_t node_index = 1;
    const RcNode2 *nodes = sc.tlas2;
    while (node_index != RC_INVALID) {
        const char *np = reinterpret_cast<const char *>(nodes + (node_index - 1));
        rc_f4 q0 = rc_load16(np), q1 = rc_load16(np + 16), q2 = rc_load16(np + 32), q3 = rc_load16(np + 48);
        if (COUNT) cnt->nodes++;
        uint32_t child0 = f2u(q3.x), child1 = f2u(q3.y);
        if (child0 != RC_INVALID) {
            f3 a0n = mk3(q0.x, q0.y, q0.z), a0x = mk3(u2f(q0.w), q1.x, q1.y);
            f3 a1n = mk3(q1.z, u2f(q1.w), q2.x), a1x = mk3(q2.y, q2.z, u2f(q2.w));
            float t0n, t0x, t1n, t1x;  // intersect_internal_node :1807-1832
            x_intersect_bbox(ray_o, inv, a0n, a0x, ray_mint, ray_maxt, t0n, t0x);
            x_intersect_bbox(ray_o, inv, a1n, a1x, ray_mint, ray_maxt, t1n, t1x);
            if (COUNT) cnt->box_tests += 2;
            uint32_t tr0 = (t0n <= t0x) ? child0 : RC_INVALID;
            uint32_t tr1 = (t1n <= t1x) ? child1 : RC_INVALID;
            uint32_t near_c, far_c;
            if (t0n < t1n && tr0 != RC_INVALID) { near_c = tr0; far_c = tr1; }
            else { near_c = tr1; far_c = tr0; }
            if (far_c != RC_INVALID) {
                if (sp >= RC_STACK_REF) return false;
                stack[sp++] = far_c;
                if (COUNT && (uint32_t)sp > cnt->max_stack) cnt->max_stack = (uint32_t)sp;
            }
            if (near_c != RC_INVALID) { node_index = near_c; continue; }
        } else if (current_instance < 0) {
            current_instance = (int)child1;  // TLAS leaf: 0-based instance index (:1963)
            if (sp >= RC_STACK_REF) return false;
            stack[sp++] = RC_SENTINEL;
            if (COUNT && (uint32_t)sp > cnt->max_stack) cnt->max_stack = (uint32_t)sp;
            node_index = 1;
            const RcInstanceRec *ir = sc.inst + current_instance;
            nodes = sc.aux[current_instance].nodes2;
            float m[12];
            for (int k = 0; k < 3; k++) {
                rc_f4 r = rc_load16(reinterpret_cast<const char *>(ir) + 16 * k);
                m[4 * k] = r.x; m[4 * k + 1] = r.y; m[4 * k + 2] = r.z; m[4 * k + 3] = u2f(r.w);
            }
            ray_o = x_transform_point(m, w.o);
            ray_d = x_transform_direction(m, w.d);
            inv = mk3(x_safe_inv(ray_d.x), x_safe_inv(ray_d.y), x_safe_inv(ray_d.z));
            if (COUNT) cnt->inst_entries++;
            continue;
        } else {
            // BLAS leaf: v0,v1,v2 in the box slots (BVH2IL, kernels.jl:198-215)
            f3 v0 = mk3(q0.x, q0.y, q0.z), v1 = mk3(u2f(q0.w), q1.x, q1.y), v2 = mk3(q1.z, u2f(q1.w), q2.x);
            float t, u, v;
            if (COUNT) cnt->tri_tests++;
            if (WT ? x_intersect_triangle_watertight(ray_o, ray_d, v0, v1, v2, ray_mint, ray_maxt, t, u, v)
                   : x_intersect_triangle(ray_o, ray_d, v0, v1, v2, ray_mint, ray_maxt, t, u, v)) {
                ray_maxt = t;
                closest_instance = current_instance;
                closest_prim = child1;  // 1-based sorted primitive index
                hit_u = u; hit_v = v;
                if (ANY) break;
            }
        }
        node_index = stack[--sp];
        if (node_index == RC_SENTINEL) {
            node_index = stack[--sp];
            current_instance = -1;
            nodes = sc.tlas2;
            ray_o = w.o; ray_d = w.d;
            inv = mk3(x_safe_inv(ray_d.x), x_safe_inv(ray_d.y), x_safe_inv(ray_d.z));
        }
    }
    if (closest_instance >= 0) {
        const RcTri *tri = sc.inst[closest_instance].tris + (closest_prim - 1);
        if (tri_out) *tri_out = tri;
        out.hit = 1; out.t = ray_maxt; out.bary_u = hit_u; out.bary_v = hit_v;
        out.primitive_id = tri->prim_id; out.metadata = tri->metadata;
        out.instance_id = (uint32_t)closest_instance;
        out.instance_custom_index = sc.aux[closest_instance].custom_index;
    }
    return true;
}

// ------------------------------------------------------------------------------------------------
// Wide (BVH4) traversal.
#define RC_BOX_EPS 2.4e-7f  // 2^-22: bound on the relative rounding error of the quantised slab evaluation

// byte k of w as float.  Device: PRMT builds 0x4B0000qq (= 2^23 + q exactly) and one FADD removes the bias; this
// keeps the decode on the ALU/FMA pipes (I2F.U8 runs on the quarter-rate XU pipe, which saturated in profiles/r1_v1).
RC_HD float rc_q2f(uint32_t w, int k) {
#if RC_ON_DEVICE
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540u + (uint32_t)k)) - 8388608.0f;
#else
    return (float)((w >> (8 * k)) & 0xFFu);
#endif
}

struct RcWideHit {
    float t[4];
    uint32_t ref[4];
    int n;
};

// Test the 4 quantised child boxes of `nd` against the ray (o, inv) over [t_lo, t_hi]; returns the hit
// children sorted near -> far.
RC_HD void rc_wide_node_test(const rc_f4 &n0, const rc_f4 &n1, const rc_f4 &n2, const rc_f4 &n3, f3 o, f3 inv, float t_lo, float t_hi, RcWideHit &h) {
    const float k24 = 5.9604644775390625e-8f;  // 2^-24: the node stores its scales pre-multiplied by 2^24 (exact rescale)
    float ax = (u2f(n0.w) * k24) * inv.x, ay = (n3.z * k24) * inv.y, az = (u2f(n3.w) * k24) * inv.z;
    float bx = (n0.x - o.x) * inv.x, by = (n0.y - o.y) * inv.y, bz = (n0.z - o.z) * inv.z;
    // error bound of fmaf(q, a, b) over q in [0,255] (both products rounded once, b rounded twice)
    float slack = RC_BOX_EPS * fmaxf(fmaxf(fmaf(255.0f, fabsf(ax), fabsf(bx)), fmaf(255.0f, fabsf(ay), fabsf(by))), fmaf(255.0f, fabsf(az), fabsf(bz)));
    uint32_t qlox = f2u(n1.x), qloy = f2u(n1.y), qloz = f2u(n1.z), qhix = n1.w;
    uint32_t qhiy = f2u(n2.x), qhiz = f2u(n2.y);
    uint32_t ch[4] = {f2u(n2.z), n2.w, f2u(n3.x), f2u(n3.y)};  // unused slots repeat child 0 under an inverted box
    // choose near/far planes per axis by the sign of the direction
    uint32_t nx = inv.x >= 0.0f ? qlox : qhix, fx = inv.x >= 0.0f ? qhix : qlox;
    uint32_t ny = inv.y >= 0.0f ? qloy : qhiy, fy = inv.y >= 0.0f ? qhiy : qloy;
    uint32_t nz = inv.z >= 0.0f ? qloz : qhiz, fz = inv.z >= 0.0f ? qhiz : qloz;
    h.n = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int k = 0; k < 4; k++) {
        float tnx = fmaf(rc_q2f(nx, k), ax, bx), tfx = fmaf(rc_q2f(fx, k), ax, bx);
        float tny = fmaf(rc_q2f(ny, k), ay, by), tfy = fmaf(rc_q2f(fy, k), ay, by);
        float tnz = fmaf(rc_q2f(nz, k), az, bz), tfz = fmaf(rc_q2f(fz, k), az, bz);
        float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, t_lo));
        float tf = fminf(fminf(tfx, tfy), fminf(tfz, t_hi));
        bool hit = tn <= tf + slack;
        if (hit) {
            // insertion sort by tn (<= 4 elements)
            int j = h.n++;
            while (j > 0 && h.t[j - 1] > tn) {
                h.t[j] = h.t[j - 1];
                h.ref[j] = h.ref[j - 1];
                j--;
            }
            h.t[j] = tn;
            h.ref[j] = ch[k];
        }
    }
}

template <bool ANY, bool COUNT, bool WT = false>
RC_HD bool rc_trace_wide(const RcScene &sc, const rc_ray &ray, rc_hit &out, RcLocalCounters *cnt, const RcTri **tri_out = nullptr) {
    rc_write_miss(out);
    if (tri_out) *tri_out = nullptr;
    if (sc.n_instances == 0) return true;
    RcRayIn w = rc_prepare_ray(ray, ANY);
    f3 ray_o = w.o, ray_d = w.d;
    float ray_mint = w.t_min, ray_maxt = w.t_max;
    f3 inv = mk3(x_safe_inv(ray_d.x), x_safe_inv(ray_d.y), x_safe_inv(ray_d.z));
    uint32_t stack[RC_STACK_WIDE];
    int sp = 1;
    stack[0] = RC_INVALID;
    int current_instance = -1, closest_instance = -1;
    const RcTri *closest_tri = nullptr;
    float hit_u = 0.0f, hit_v = 0.0f;
    const RcNode4 *nodes = sc.tlas4;
    const RcTri *tris = nullptr;
    uint32_t cur = 1;  // wide-node index of the TLAS root
    bool done = false;
    while (!done) {
        if (!(cur & RC_LEAF_BIT)) {
            const char *np = reinterpret_cast<const char *>(nodes + cur);
            rc_f4 n0 = rc_load16(np), n1 = rc_load16(np + 16), n2 = rc_load16(np + 32), n3 = rc_load16(np + 48);
            if (COUNT) { cnt->nodes++; cnt->box_tests += 4; }
            RcWideHit h;
            rc_wide_node_test(n0, n1, n2, n3, ray_o, inv, ray_mint, ray_maxt, h);
            if (h.n > 0) {
                if (sp + h.n - 1 > RC_STACK_WIDE) return false;
                for (int k = h.n - 1; k >= 1; k--) stack[sp++] = h.ref[k];
                if (COUNT && (uint32_t)sp > cnt->max_stack) cnt->max_stack = (uint32_t)sp;
                cur = h.ref[0];
                continue;
            }
        } else if (current_instance < 0) {
            // TLAS leaf: enter the instance (src/instanced-bvh.jl:1961-1977)
            current_instance = (int)(cur & RC_LEAF_START_MASK);
            if (sp >= RC_STACK_WIDE) return false;
            stack[sp++] = RC_SENTINEL;
            if (COUNT && (uint32_t)sp > cnt->max_stack) cnt->max_stack = (uint32_t)sp;
            const char *ip = reinterpret_cast<const char *>(sc.inst + current_instance);
            float m[12];
            for (int k = 0; k < 3; k++) {
                rc_f4 r = rc_load16(ip + 16 * k);
                m[4 * k] = r.x; m[4 * k + 1] = r.y; m[4 * k + 2] = r.z; m[4 * k + 3] = u2f(r.w);
            }
            nodes = sc.inst[current_instance].nodes4;
            tris = sc.inst[current_instance].tris;
            ray_o = x_transform_point(m, w.o);
            ray_d = x_transform_direction(m, w.d);
            inv = mk3(x_safe_inv(ray_d.x), x_safe_inv(ray_d.y), x_safe_inv(ray_d.z));
            if (COUNT) cnt->inst_entries++;
            cur = 1;
            continue;
        } else {
            uint32_t start = cur & RC_LEAF_START_MASK, count = ((cur >> RC_LEAF_COUNT_SHIFT) & 7u) + 1u;
            for (uint32_t k = 0; k < count; k++) {
                const char *tp = reinterpret_cast<const char *>(tris + start + k);
                rc_f4 a = rc_load16(tp), b = rc_load16(tp + 16), c = rc_load16(tp + 32);
                float t, u, v;
                if (COUNT) cnt->tri_tests++;
                if (WT ? x_intersect_triangle_watertight(ray_o, ray_d, mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), mk3(c.x, c.y, c.z), ray_mint, ray_maxt, t, u, v)
                       : x_intersect_triangle(ray_o, ray_d, mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), mk3(c.x, c.y, c.z), ray_mint, ray_maxt, t, u, v)) {
                    if (t == t) {  // a NaN t (ray in the triangle's plane) is rejected here; documented deviation
                        ray_maxt = t;
                        closest_instance = current_instance;
                        closest_tri = tris + start + k;
                        hit_u = u; hit_v = v;
                        if (ANY) { done = true; break; }
                    }
                }
            }
            if (done) break;
        }
        cur = stack[--sp];
        if (cur == RC_SENTINEL) {
            cur = stack[--sp];
            current_instance = -1;
            nodes = sc.tlas4;
            ray_o = w.o; ray_d = w.d;
            inv = mk3(x_safe_inv(ray_d.x), x_safe_inv(ray_d.y), x_safe_inv(ray_d.z));
        }
        if (cur == RC_INVALID) done = true;
    }
    if (closest_instance >= 0) {
        if (tri_out) *tri_out = closest_tri;
        out.hit = 1; out.t = ray_maxt; out.bary_u = hit_u; out.bary_v = hit_v;
        out.primitive_id = closest_tri->prim_id; out.metadata = closest_tri->metadata;
        out.instance_id = (uint32_t)closest_instance;
        out.instance_custom_index = sc.aux[closest_instance].custom_index;
    }
    return true;
}
