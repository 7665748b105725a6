/*
 * oracle.c — CPU restatement of Raycore.jl's ray-query hot path (see oracle.h for scope
 * and parity status).  TEST INFRASTRUCTURE ONLY: never linked into the product.
 *
 * Arithmetic rules honoured:
 *   - IEEE float32 throughout, no FMA contraction (compile with -ffp-contract=off),
 *     expressions fold left exactly as the Julia source is written;
 *   - Julia min/max semantics (NaN-propagating, -0 < +0);
 *   - `x ≈ 0f0` on Float32 means x == 0;
 *   - unsafe_trunc(UInt32, NaN) == 0 (x86-64 cvttss2si behaviour).
 * Citations are file:line in /root/reference.
 */
#include "oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_STACK 160 /* reference: MVector{32} without overflow check (instanced-bvh.jl:1912) */

/* ------------------------------------------------------------------ Julia float semantics */
static inline float jl_min(float x, float y) {
    float diff = x - y;
    if (isnan(x) || isnan(y)) return diff;
    return signbit(diff) ? x : y;
}
static inline float jl_max(float x, float y) {
    float diff = x - y;
    if (isnan(x) || isnan(y)) return diff;
    return signbit(diff) ? y : x;
}
static inline void v_min(const float a[3], const float b[3], float o[3]) {
    for (int k = 0; k < 3; k++) o[k] = jl_min(a[k], b[k]);
}
static inline void v_max(const float a[3], const float b[3], float o[3]) {
    for (int k = 0; k < 3; k++) o[k] = jl_max(a[k], b[k]);
}
/* StaticArrays cross / dot: component formulas, left fold, no muladd */
static inline void v_cross(const float a[3], const float b[3], float o[3]) {
    float x = a[1] * b[2] - a[2] * b[1];
    float y = a[2] * b[0] - a[0] * b[2];
    float z = a[0] * b[1] - a[1] * b[0];
    o[0] = x; o[1] = y; o[2] = z;
}
static inline float v_dot(const float a[3], const float b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void v_sub(const float a[3], const float b[3], float o[3]) {
    for (int k = 0; k < 3; k++) o[k] = a[k] - b[k];
}
/* StaticArrays normalize(a) = inv(norm(a)) * a, norm = sqrt(a1^2 + a2^2 + a3^2) */
static inline void v_normalize(const float a[3], float o[3]) {
    float n = sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    float inv = 1.0f / n;
    for (int k = 0; k < 3; k++) o[k] = inv * a[k];
}

/* ------------------------------------------------------------------ scalar helpers */
uint32_t orc_expand_bits(uint32_t x) { /* instanced-bvh.jl:1177-1183 */
    x = (x * 0x00010001u) & 0xFF0000FFu;
    x = (x * 0x00000101u) & 0x0F00F00Fu;
    x = (x * 0x00000011u) & 0xC30C30C3u;
    x = (x * 0x00000005u) & 0x49249249u;
    return x;
}

static inline float jl_clamp(float x, float lo, float hi) { /* Base.clamp: NaN stays NaN */
    return x > hi ? hi : (x < lo ? lo : x);
}
static inline uint32_t jl_unsafe_trunc_u32(float x) {
    if (isnan(x)) return 0u; /* x86-64: cvttss2si r64 -> 0x8000000000000000 -> low word 0 */
    return (uint32_t)(int64_t)x;
}

uint32_t orc_morton_code_30bit(const float p[3]) { /* :1189-1200 */
    const float unit_side = 1024.0f;
    float x = jl_clamp(p[0] * unit_side, 0.0f, unit_side - 1.0f);
    float y = jl_clamp(p[1] * unit_side, 0.0f, unit_side - 1.0f);
    float z = jl_clamp(p[2] * unit_side, 0.0f, unit_side - 1.0f);
    return (orc_expand_bits(jl_unsafe_trunc_u32(x)) << 2) | (orc_expand_bits(jl_unsafe_trunc_u32(y)) << 1) |
           orc_expand_bits(jl_unsafe_trunc_u32(z));
}

int32_t orc_clz32(uint32_t x) { /* :1203-1206 */
    if (x == 0) return 32;
    return (int32_t)__builtin_clz(x);
}

int32_t orc_delta(int32_t i1, int32_t i2, const uint32_t *codes, int32_t n) { /* :1212-1229 */
    int32_t left = i1 < i2 ? i1 : i2;
    int32_t right = i1 < i2 ? i2 : i1;
    if (left < 1 || right > n) return -1;
    uint32_t lc = codes[left - 1], rc = codes[right - 1];
    if (lc != rc) return orc_clz32(lc ^ rc);
    return 32 + orc_clz32((uint32_t)left ^ (uint32_t)right);
}

int orc_is_degenerate(const float v[9]) { /* triangle_mesh.jl:14-17 */
    float a[3], b[3], c[3];
    v_sub(v + 6, v + 0, a); /* vs[3] - vs[1] */
    v_sub(v + 3, v + 0, b); /* vs[2] - vs[1] */
    v_cross(a, b, c);
    return v_dot(c, c) == 0.0f;
}

void orc_mat4_to_mat3x4(const float m[16], float out[12]) { /* :1663-1669; m column-major: m[i,j] = m[(j-1)*4 + (i-1)] */
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 4; j++) out[4 * i + j] = m[j * 4 + i];
}

/* Mat3x4f element m[r,c] (1-based, SMatrix{4,3} column-major) = mem[(c-1)*4 + (r-1)] */
#define M34(m, r, c) ((m)[((c)-1) * 4 + ((r)-1)])

void orc_mat3x4_inverse(const float m[12], float out[12]) { /* :1675-1687 */
    /* R = m[1:3,1:3] (column-major 3x3: A[k] = R[(k-1)%3+1, (k-1)/3+1]).  StaticArrays inv of a
     * 3x3 (src/inv.jl, _inv(::Size{(3,3)})): x0,x1,x2 = columns; y0 = x1 x x2; d = x0.y0;
     * x0 /= d; y0 /= d; y1 = x2 x x0; y2 = x0 x x1; B = [y0 y1 y2]^T (rows). */
    float x0[3] = {M34(m, 1, 1), M34(m, 2, 1), M34(m, 3, 1)};
    float x1[3] = {M34(m, 1, 2), M34(m, 2, 2), M34(m, 3, 2)};
    float x2[3] = {M34(m, 1, 3), M34(m, 2, 3), M34(m, 3, 3)};
    float y0[3], y1[3], y2[3];
    v_cross(x1, x2, y0);
    float d = v_dot(x0, y0);
    for (int k = 0; k < 3; k++) { x0[k] = x0[k] / d; y0[k] = y0[k] / d; }
    v_cross(x2, x0, y1);
    v_cross(x0, x1, y2);
    /* B[i,j]: row i = y_{i-1} */
    float B[3][3] = {{y0[0], y0[1], y0[2]}, {y1[0], y1[1], y1[2]}, {y2[0], y2[1], y2[2]}};
    float tx = M34(m, 4, 1), ty = M34(m, 4, 2), tz = M34(m, 4, 3);
    float tix = -(B[0][0] * tx + B[1][0] * ty + B[2][0] * tz);
    float tiy = -(B[0][1] * tx + B[1][1] * ty + B[2][1] * tz);
    float tiz = -(B[0][2] * tx + B[1][2] * ty + B[2][2] * tz);
    float o[12] = {B[0][0], B[1][0], B[2][0], tix, B[0][1], B[1][1], B[2][1], tiy, B[0][2], B[1][2], B[2][2], tiz};
    memcpy(out, o, sizeof o);
}

void orc_transform_point(const float m[12], const float p[3], float out[3]) { /* :1692-1698 */
    float r[3];
    for (int i = 0; i < 3; i++) r[i] = m[4 * i + 0] * p[0] + m[4 * i + 1] * p[1] + m[4 * i + 2] * p[2] + m[4 * i + 3];
    out[0] = r[0]; out[1] = r[1]; out[2] = r[2];
}
void orc_transform_direction(const float m[12], const float v[3], float out[3]) { /* :1711-1717 */
    float r[3];
    for (int i = 0; i < 3; i++) r[i] = m[4 * i + 0] * v[0] + m[4 * i + 1] * v[1] + m[4 * i + 2] * v[2];
    out[0] = r[0]; out[1] = r[1]; out[2] = r[2];
}

void orc_safe_invdir(const float d[3], float out[3]) { /* :1742-1748 */
    const float ooeps = 1.0e-5f;
    for (int k = 0; k < 3; k++) out[k] = 1.0f / (fabsf(d[k]) > ooeps ? d[k] : copysignf(ooeps, d[k]));
}

int orc_intersect_triangle(const float o[3], const float dir[3], const float v0[3], const float v1[3],
                           const float v2[3], float t_min, float closest_t, float *t_out, float *u_out,
                           float *v_out) { /* :1756-1797 */
    float e1[3], e2[3], s1[3], s2[3], d[3];
    v_sub(v1, v0, e1);
    v_sub(v2, v0, e2);
    v_cross(dir, e2, s1);
    float determinant = v_dot(s1, e1);
    float invd = 1.0f / determinant;
    v_sub(o, v0, d);
    float u = v_dot(d, s1) * invd;
    if (u < 0.0f || u > 1.0f) return 0;
    v_cross(d, e1, s2);
    float v = v_dot(dir, s2) * invd;
    if (v < 0.0f || (u + v) > 1.0f) return 0;
    float t = v_dot(e2, s2) * invd;
    if (t < t_min || t > closest_t) return 0;
    *t_out = t; *u_out = u; *v_out = v;
    return 1;
}

/* intersect_triangle, src/triangle_mesh.jl:168-201 with _to_ray_coordinate_space (:84-117), _edge_function (:24-30), _argmax (:78-88):
 * the watertight (pbrt) test — permute so the dominant direction axis is z, shear the vertices into ray space, signed edge functions.
 * The reference tests t against ray.t_max only; a traversal passes the closest t so far there and also honours t_min, as
 * closest_hit does for Moeller-Trumbore (:1792).  u, v = barycentric weights of v1, v2 (edges[2], edges[3] * inv_det).
 * is_degenerate(vs) (:171) is not re-tested: the builder filtered those faces with the same rule. */
int orc_intersect_triangle_watertight(const float o[3], const float dir[3], const float v0[3], const float v1[3], const float v2[3],
                                      float t_min, float t_max, float *t_out, float *u_out, float *v_out) {
    int kz = 0; /* _argmax(map(abs, ray.d)): first maximum */
    float mx = fabsf(dir[0]);
    for (int i = 0; i < 3; i++) if (fabsf(dir[i]) > mx) { mx = fabsf(dir[i]); kz = i; }
    int kx = kz + 1; if (kx == 3) kx = 0;
    int ky = kx + 1; if (ky == 3) ky = 0;
    float dx = dir[kx], dy = dir[ky], dz = dir[kz];
    float denom = 1.0f / dz;
    float shx = -dx * denom, shy = -dy * denom, shz = denom;
    float rkz = o[kz];
    const float *vs[3] = {v0, v1, v2};
    float tx[3], ty[3], tz[3];
    for (int i = 0; i < 3; i++) {
        const float *v = vs[i];
        float vox = v[kx] - o[kx], voy = v[ky] - o[ky], voz = v[kz] - o[kz];
        tx[i] = vox + shx * (v[kz] - rkz);
        ty[i] = voy + shy * (v[kz] - rkz);
        tz[i] = voz + 0.0f;
    }
    float e0 = tx[1] * ty[2] - ty[1] * tx[2];
    float e1 = tx[2] * ty[0] - ty[2] * tx[0];
    float e2 = tx[0] * ty[1] - ty[0] * tx[1];
    if (e0 == 0.0f && e1 == 0.0f && e2 == 0.0f) return 0;                                   /* iszero(edges) */
    if ((e0 < 0.0f || e1 < 0.0f || e2 < 0.0f) && (e0 > 0.0f || e1 > 0.0f || e2 > 0.0f)) return 0;
    float det = (e0 + e1) + e2;
    if (det == 0.0f) return 0;                                                               /* det ≈ 0f0 */
    float t_scaled = ((e0 * tz[0]) * shz + (e1 * tz[1]) * shz) + (e2 * tz[2]) * shz;
    if (det < 0.0f && (t_scaled >= 0.0f || t_scaled < t_max * det)) return 0;
    if (det > 0.0f && (t_scaled <= 0.0f || t_scaled > t_max * det)) return 0;
    float inv_det = 1.0f / det;
    float t = t_scaled * inv_det;
    if (t < t_min) return 0;
    *t_out = t; *u_out = e1 * inv_det; *v_out = e2 * inv_det;
    return 1;
}

void orc_intersect_bbox(const float o[3], const float inv_d[3], const float pmin[3], const float pmax[3],
                        float t_min, float t_max, float *out_min, float *out_max) { /* :1841-1859 */
    float f[3], n[3], tmx[3], tmn[3];
    for (int k = 0; k < 3; k++) {
        float oxinv = -o[k] * inv_d[k];
        f[k] = pmax[k] * inv_d[k] + oxinv;
        n[k] = pmin[k] * inv_d[k] + oxinv;
        tmx[k] = jl_max(f[k], n[k]);
        tmn[k] = jl_min(f[k], n[k]);
    }
    float mn = jl_min(jl_min(tmx[0], tmx[1]), tmx[2]);
    float mx = jl_max(jl_max(tmn[0], tmn[1]), tmn[2]);
    *out_max = jl_min(mn, t_max);
    *out_min = jl_max(mx, t_min);
}

/* ------------------------------------------------------------------ filter */
uint32_t orc_filter_triangles(const float *verts, uint32_t n_faces, const uint32_t *face_meta, orc_tri *out) {
    uint32_t k = 0;
    for (uint32_t i = 0; i < n_faces; i++) {
        const float *v = verts + (size_t)i * 9;
        if (orc_is_degenerate(v)) continue; /* instanced-bvh.jl:599 */
        memcpy(out[k].v, v, 9 * sizeof(float));
        out[k].metadata = face_meta ? face_meta[i] : i + 1; /* :595 */
        out[k].input_index = k;
        k++;
    }
    return k;
}

/* ------------------------------------------------------------------ LBVH pieces shared by BLAS and TLAS */
static void stable_sortperm_u32(const uint32_t *keys, uint32_t n, uint32_t *perm) {
    /* stable LSD radix sort; Base.sortperm / AK.sortperm are stable (instanced-bvh.jl:1399,1534) */
    uint32_t *a = (uint32_t *)malloc(sizeof(uint32_t) * n), *b = (uint32_t *)malloc(sizeof(uint32_t) * n);
    for (uint32_t i = 0; i < n; i++) a[i] = i;
    for (int pass = 0; pass < 4; pass++) {
        size_t cnt[257] = {0};
        int sh = pass * 8;
        for (uint32_t i = 0; i < n; i++) cnt[((keys[a[i]] >> sh) & 255u) + 1]++;
        for (int d = 0; d < 256; d++) cnt[d + 1] += cnt[d];
        for (uint32_t i = 0; i < n; i++) b[cnt[(keys[a[i]] >> sh) & 255u]++] = a[i];
        uint32_t *t = a; a = b; b = t;
    }
    memcpy(perm, a, sizeof(uint32_t) * n);
    free(a); free(b);
}

static void find_span_for_node(int32_t idx, const uint32_t *codes, int32_t n, int32_t *lo, int32_t *hi) { /* :1232-1262 */
    int32_t d_left = orc_delta(idx, idx - 1, codes, n);
    int32_t d_right = orc_delta(idx, idx + 1, codes, n);
    int32_t d = d_right > d_left ? 1 : -1;
    int32_t delta_min = orc_delta(idx, idx - d, codes, n);
    int32_t l_max = 2;
    while (orc_delta(idx, idx + l_max * d, codes, n) > delta_min) l_max *= 2;
    int32_t l = 0, t = l_max;
    while (t > 1) {
        t = t / 2;
        if (orc_delta(idx, idx + (l + t) * d, codes, n) > delta_min) l = l + t;
    }
    int32_t j = idx + l * d;
    if (d > 0) { *lo = idx; *hi = j; } else { *lo = j; *hi = idx; }
}

static int32_t find_split_in_span(int32_t span_left, int32_t span_right, const uint32_t *codes, int32_t n) { /* :1265-1290 */
    int32_t numidentical = orc_delta(span_left, span_right, codes, n);
    int32_t left = span_left, right = span_right;
    while (right > left + 1) {
        int32_t newsplit = (right + left) / 2;
        if (orc_delta(left, newsplit, codes, n) > numidentical) left = newsplit;
        else right = newsplit;
    }
    return left;
}

static void fill_empty(orc_node2 *nodes, uint32_t count) { /* fill_bvhnode2_kernel!, kernels.jl:19-22 */
    for (uint32_t i = 0; i < count; i++) {
        memset(&nodes[i], 0, sizeof(orc_node2));
        nodes[i].child0 = nodes[i].child1 = nodes[i].parent = ORC_INVALID_NODE;
    }
}

/* emit_topology_kernel! + set_parent_pointers_kernel!, kernels.jl:119-191 */
static void emit_topology(orc_node2 *nodes, const uint32_t *codes, int32_t n) {
    for (int32_t idx = 1; idx < n; idx++) {
        int32_t lo, hi;
        find_span_for_node(idx, codes, n, &lo, &hi);
        int32_t split = find_split_in_span(lo, hi, codes, n);
        int32_t child0 = (split == lo) ? (n - 1 + split) : split;
        int32_t c1 = split + 1;
        int32_t child1 = (c1 == hi) ? (n - 1 + c1) : c1;
        nodes[idx - 1].child0 = (uint32_t)child0;
        nodes[idx - 1].child1 = (uint32_t)child1;
        nodes[idx - 1].parent = ORC_INVALID_NODE;
    }
    for (int32_t idx = 1; idx < n; idx++) {
        nodes[nodes[idx - 1].child0 - 1].parent = (uint32_t)idx;
        nodes[nodes[idx - 1].child1 - 1].parent = (uint32_t)idx;
    }
}

static void node_aabb_blas(const orc_node2 *nd, int interior, float mn[3], float mx[3]) { /* get_node_aabb :1141-1160 */
    if (interior) {
        v_min(nd->aabb0_min, nd->aabb1_min, mn);
        v_max(nd->aabb0_max, nd->aabb1_max, mx);
    } else {
        float t[3];
        v_min(nd->aabb0_min, nd->aabb0_max, t); v_min(t, nd->aabb1_min, mn);
        v_max(nd->aabb0_min, nd->aabb0_max, t); v_max(t, nd->aabb1_min, mx);
    }
}
static void node_aabb_tlas(const orc_node2 *nd, int interior, float mn[3], float mx[3]) { /* get_tlas_node_aabb :1163-1174 */
    if (interior) {
        v_min(nd->aabb0_min, nd->aabb1_min, mn);
        v_max(nd->aabb0_max, nd->aabb1_max, mx);
    } else {
        memcpy(mn, nd->aabb0_min, 12); memcpy(mx, nd->aabb0_max, 12);
    }
}

/* refit_aabbs_kernel! / refit_tlas_aabbs_kernel!, kernels.jl:239-286, 381-428, run sequentially:
 * the second arriver at a node computes it, exactly as the atomic protocol does. */
static void refit_bottom_up(orc_node2 *nodes, int32_t n, int tlas) {
    if (n < 2) return;
    uint32_t *flags = (uint32_t *)calloc((size_t)n - 1, sizeof(uint32_t));
    for (int32_t prim = 1; prim <= n; prim++) {
        uint32_t parent = nodes[(n - 1 + prim) - 1].parent;
        while (parent != ORC_INVALID_NODE) {
            uint32_t nv = ++flags[parent - 1];
            if (nv != 2) break;
            orc_node2 *nd = &nodes[parent - 1];
            uint32_t c0 = nd->child0, c1 = nd->child1;
            float a0n[3], a0x[3], a1n[3], a1x[3];
            if (tlas) {
                node_aabb_tlas(&nodes[c0 - 1], c0 < (uint32_t)n, a0n, a0x);
                node_aabb_tlas(&nodes[c1 - 1], c1 < (uint32_t)n, a1n, a1x);
            } else {
                node_aabb_blas(&nodes[c0 - 1], c0 < (uint32_t)n, a0n, a0x);
                node_aabb_blas(&nodes[c1 - 1], c1 < (uint32_t)n, a1n, a1x);
            }
            memcpy(nd->aabb0_min, a0n, 12); memcpy(nd->aabb0_max, a0x, 12);
            memcpy(nd->aabb1_min, a1n, 12); memcpy(nd->aabb1_max, a1x, 12);
            parent = nd->parent;
        }
    }
    free(flags);
}

/* ------------------------------------------------------------------ build_blas :1376-1443 */
orc_blas *orc_build_blas(const orc_tri *tris, uint32_t n) {
    if (n == 0) return NULL; /* "Cannot build BLAS from empty primitive list" */
    orc_blas *b = (orc_blas *)calloc(1, sizeof(orc_blas));
    b->n = n;
    /* scene AABB = mapreduce(world_bound, ∪, prims, init=Bounds3()) :1386 */
    float smin[3] = {INFINITY, INFINITY, INFINITY}, smax[3] = {-INFINITY, -INFINITY, -INFINITY};
    float(*tmin)[3] = (float(*)[3])malloc(sizeof(float) * 3 * n), (*tmax)[3] = (float(*)[3])malloc(sizeof(float) * 3 * n);
    for (uint32_t i = 0; i < n; i++) {
        const float *v = tris[i].v;
        float t[3];
        v_min(v, v + 3, t); v_min(t, v + 6, tmin[i]); /* world_bound(tri), triangle_mesh.jl:37 */
        v_max(v, v + 3, t); v_max(t, v + 6, tmax[i]);
        v_min(smin, tmin[i], smin);
        v_max(smax, tmax[i], smax);
    }
    float extent[3];
    v_sub(smax, smin, extent); /* :1388, unguarded */
    uint32_t *codes = (uint32_t *)malloc(sizeof(uint32_t) * n);
    for (uint32_t i = 0; i < n; i++) { /* calculate_morton_code_for_prim, kernels.jl:88-98 */
        float c[3], nrm[3];
        for (int k = 0; k < 3; k++) {
            c[k] = 0.5f * (tmin[i][k] + tmax[i][k]);
            nrm[k] = (c[k] - smin[k]) / extent[k];
        }
        codes[i] = orc_morton_code_30bit(nrm);
    }
    free(tmin); free(tmax);
    uint32_t *perm = (uint32_t *)malloc(sizeof(uint32_t) * n);
    stable_sortperm_u32(codes, n, perm); /* :1399-1402 */
    b->morton = (uint32_t *)malloc(sizeof(uint32_t) * n);
    b->prims = (orc_tri *)malloc(sizeof(orc_tri) * n);
    for (uint32_t i = 0; i < n; i++) { b->morton[i] = codes[perm[i]]; b->prims[i] = tris[perm[i]]; }
    free(codes); free(perm);

    uint32_t nn = 2 * n - 1;
    b->nodes = (orc_node2 *)malloc(sizeof(orc_node2) * nn);
    fill_empty(b->nodes, nn);
    if (n > 1) emit_topology(b->nodes, b->morton, (int32_t)n);
    for (uint32_t i = 1; i <= n; i++) { /* create_leaf_for_prim, kernels.jl:198-215 */
        orc_node2 *lf = &b->nodes[(n - 1 + i) - 1];
        uint32_t parent = lf->parent;
        memcpy(lf->aabb0_min, b->prims[i - 1].v + 0, 12);
        memcpy(lf->aabb0_max, b->prims[i - 1].v + 3, 12);
        memcpy(lf->aabb1_min, b->prims[i - 1].v + 6, 12);
        lf->aabb1_max[0] = lf->aabb1_max[1] = lf->aabb1_max[2] = 0.0f;
        lf->child0 = ORC_INVALID_NODE;
        lf->child1 = i;
        lf->parent = parent;
    }
    refit_bottom_up(b->nodes, (int32_t)n, 0);
    node_aabb_blas(&b->nodes[0], b->nodes[0].child0 != ORC_INVALID_NODE, b->root_aabb, b->root_aabb + 3); /* :1438-1440 */
    return b;
}

void orc_free_blas(orc_blas *b) {
    if (!b) return;
    free(b->nodes); free(b->prims); free(b->morton); free(b);
}

/* ------------------------------------------------------------------ TLAS */
static void corner(const float bb[6], int c, float out[3]) { /* bounds.jl:53-59, c 1-based */
    c -= 1;
    out[0] = (c & 1) == 0 ? bb[0] : bb[3];
    out[1] = (c & 2) == 0 ? bb[1] : bb[4];
    out[2] = (c & 4) == 0 ? bb[2] : bb[5];
}

static void instance_world_aabb(const orc_instance *inst, const float local[6], float mn[3], float mx[3]) {
    /* compute_instance_world_aabb kernels.jl:38-62 == create_tlas_leaf_for_instance :345-350 (same values) */
    float c[3], w[3];
    corner(local, 1, c);
    orc_transform_point(inst->transform, c, w);
    memcpy(mn, w, 12); memcpy(mx, w, 12);
    for (int k = 2; k <= 8; k++) {
        corner(local, k, c);
        orc_transform_point(inst->transform, c, w);
        v_min(mn, w, mn);
        v_max(mx, w, mx);
    }
}

/* build_tlas_topology :1485-1594.  root_aabbs[b] = local root box of BLAS b (0-based b). */
static void build_tlas_topology(const float (*root_aabbs)[6], const orc_instance *inst, uint32_t n, orc_node2 **nodes_out,
                                uint32_t *n_nodes_out, float root_aabb[6]) {
    float(*mins)[3] = (float(*)[3])malloc(sizeof(float) * 3 * n), (*maxs)[3] = (float(*)[3])malloc(sizeof(float) * 3 * n);
    for (uint32_t i = 0; i < n; i++) instance_world_aabb(&inst[i], root_aabbs[inst[i].blas_index - 1], mins[i], maxs[i]);
    float smin[3], smax[3];
    memcpy(smin, mins[0], 12); memcpy(smax, maxs[0], 12);
    for (uint32_t i = 1; i < n; i++) { v_min(smin, mins[i], smin); v_max(smax, maxs[i], smax); } /* :1502-1511 */
    float extent[3];
    for (int k = 0; k < 3; k++) extent[k] = jl_max(smax[k] - smin[k], 1e-6f); /* :1517-1521 */
    uint32_t *codes = (uint32_t *)malloc(sizeof(uint32_t) * n);
    for (uint32_t i = 0; i < n; i++) { /* calculate_tlas_morton_code kernels.jl:295-313 */
        const float *la = root_aabbs[inst[i].blas_index - 1];
        float lc[3], wc[3], nrm[3];
        for (int k = 0; k < 3; k++) lc[k] = 0.5f * (la[k] + la[3 + k]);
        orc_transform_point(inst[i].transform, lc, wc);
        for (int k = 0; k < 3; k++) nrm[k] = (wc[k] - smin[k]) / extent[k];
        codes[i] = orc_morton_code_30bit(nrm);
    }
    uint32_t *perm = (uint32_t *)malloc(sizeof(uint32_t) * n);
    stable_sortperm_u32(codes, n, perm);
    uint32_t *sorted = (uint32_t *)malloc(sizeof(uint32_t) * n);
    for (uint32_t i = 0; i < n; i++) sorted[i] = codes[perm[i]];
    uint32_t nn = n > 0 ? 2 * n - 1 : 1; /* max(1, 2n-1) :1544 */
    orc_node2 *nodes = (orc_node2 *)malloc(sizeof(orc_node2) * nn);
    fill_empty(nodes, nn);
    if (n == 1) { /* :1553-1570 */
        memcpy(nodes[0].aabb0_min, smin, 12); memcpy(nodes[0].aabb0_max, smax, 12);
        nodes[0].child0 = ORC_INVALID_NODE;
        nodes[0].child1 = perm[0]; /* original_idx - 1 */
        nodes[0].parent = ORC_INVALID_NODE;
        memcpy(root_aabb, smin, 12); memcpy(root_aabb + 3, smax, 12);
    } else {
        emit_topology(nodes, sorted, (int32_t)n);
        for (uint32_t i = 1; i <= n; i++) { /* create_tlas_leaf_for_instance kernels.jl:332-357 */
            orc_node2 *lf = &nodes[(n - 1 + i) - 1];
            uint32_t parent = lf->parent, orig = perm[i - 1];
            memcpy(lf->aabb0_min, mins[orig], 12); memcpy(lf->aabb0_max, maxs[orig], 12);
            memset(lf->aabb1_min, 0, 12); memset(lf->aabb1_max, 0, 12);
            lf->child0 = ORC_INVALID_NODE;
            lf->child1 = orig;
            lf->parent = parent;
        }
        refit_bottom_up(nodes, (int32_t)n, 1);
        node_aabb_tlas(&nodes[0], 1, root_aabb, root_aabb + 3); /* :1590-1591 */
    }
    free(mins); free(maxs); free(codes); free(perm); free(sorted);
    *nodes_out = nodes;
    *n_nodes_out = nn;
}

orc_tlas *orc_build_tlas(orc_blas *const *blas, uint32_t n_blas, const orc_instance *inst, uint32_t n) { /* :1605-1651 */
    orc_tlas *t = (orc_tlas *)calloc(1, sizeof(orc_tlas));
    t->n_instances = n;
    t->n_blas = n_blas;
    for (int k = 0; k < 3; k++) { t->root_aabb[k] = INFINITY; t->root_aabb[3 + k] = -INFINITY; } /* Bounds3() */
    if (n == 0) return t; /* :1612-1620 */
    t->instances = (orc_instance *)malloc(sizeof(orc_instance) * n);
    memcpy(t->instances, inst, sizeof(orc_instance) * n);
    float(*roots)[6] = (float(*)[6])malloc(sizeof(float) * 6 * (n_blas ? n_blas : 1));
    for (uint32_t i = 0; i < n_blas; i++) memcpy(roots[i], blas[i]->root_aabb, 24);
    build_tlas_topology((const float(*)[6])roots, t->instances, n, &t->nodes, &t->n_nodes, t->root_aabb);
    free(roots);
    t->descs = (orc_blas_desc *)malloc(sizeof(orc_blas_desc) * (n_blas ? n_blas : 1));
    uint32_t tn = 0, tp = 0;
    for (uint32_t i = 0; i < n_blas; i++) {
        t->descs[i].nodes_offset = tn;
        t->descs[i].primitives_offset = tp;
        memcpy(t->descs[i].root_aabb, blas[i]->root_aabb, 24);
        tn += 2 * blas[i]->n - 1;
        tp += blas[i]->n;
    }
    t->n_blas_nodes = tn;
    t->n_blas_prims = tp;
    t->all_blas_nodes = (orc_node2 *)malloc(sizeof(orc_node2) * (tn ? tn : 1));
    t->all_blas_prims = (orc_tri *)malloc(sizeof(orc_tri) * (tp ? tp : 1));
    for (uint32_t i = 0; i < n_blas; i++) {
        memcpy(t->all_blas_nodes + t->descs[i].nodes_offset, blas[i]->nodes, sizeof(orc_node2) * (2 * blas[i]->n - 1));
        memcpy(t->all_blas_prims + t->descs[i].primitives_offset, blas[i]->prims, sizeof(orc_tri) * blas[i]->n);
    }
    return t;
}

void orc_free_tlas(orc_tlas *t) {
    if (!t) return;
    free(t->nodes); free(t->instances); free(t->all_blas_nodes); free(t->all_blas_prims); free(t->descs); free(t);
}

void orc_refit_tlas(orc_tlas *t) { /* refit_tlas! :2197-2222 */
    uint32_t n = t->n_instances;
    if (n == 0) return;
    for (uint32_t i = 1; i <= n; i++) { /* update_tlas_leaf_aabbs_kernel!, kernels.jl:487-519 */
        orc_node2 *lf = &t->nodes[(n - 1 + i) - 1];
        const orc_instance *inst = &t->instances[lf->child1];
        float mn[3], mx[3];
        instance_world_aabb(inst, t->descs[inst->blas_index - 1].root_aabb, mn, mx);
        memcpy(lf->aabb0_min, mn, 12); memcpy(lf->aabb0_max, mx, 12);
        memset(lf->aabb1_min, 0, 12); memset(lf->aabb1_max, 0, 12);
    }
    if (n > 1) refit_bottom_up(t->nodes, (int32_t)n, 1);
    node_aabb_tlas(&t->nodes[0], t->nodes[0].child0 != ORC_INVALID_NODE, t->root_aabb, t->root_aabb + 3); /* :2218-2219 */
}

/* ------------------------------------------------------------------ traversal :1902-2140 */
static inline void check_direction(const float d[3], float out[3]) { /* ray.jl:39-49: i ≈ 0f0 ? 0f0 : i */
    for (int k = 0; k < 3; k++) out[k] = (d[k] == 0.0f) ? 0.0f : d[k];
}

static inline void intersect_internal_node(const orc_node2 *nd, const float inv_d[3], const float o[3], float t_min,
                                           float t_max, uint32_t *near_c, uint32_t *far_c) { /* :1807-1832 */
    float t0n, t0x, t1n, t1x;
    orc_intersect_bbox(o, inv_d, nd->aabb0_min, nd->aabb0_max, t_min, t_max, &t0n, &t0x);
    orc_intersect_bbox(o, inv_d, nd->aabb1_min, nd->aabb1_max, t_min, t_max, &t1n, &t1x);
    uint32_t tr0 = (t0n <= t0x) ? nd->child0 : ORC_INVALID_NODE;
    uint32_t tr1 = (t1n <= t1x) ? nd->child1 : ORC_INVALID_NODE;
    if (t0n < t1n && tr0 != ORC_INVALID_NODE) { *near_c = tr0; *far_c = tr1; }
    else { *near_c = tr1; *far_c = tr0; }
}

/* flat_prim (nullable): 0-based position of the hit triangle in all_blas_prims */
/* mode: bit 0 = any_hit, bit 1 = the watertight triangle test instead of Moeller-Trumbore (the library's RC_MODE_WATERTIGHT) */
static void traverse(const orc_tlas *tl, const orc_ray *ray, int mode, orc_hit *out, orc_counters *cnt, uint32_t *flat_prim) {
    const int any = mode & 1, wt = (mode >> 1) & 1;
    memset(out, 0, sizeof *out);
    if (tl->n_instances == 0) return; /* reference indexes an empty array here (UB); tests require a miss (test_tlas_stress.jl:828) */
    float world_d[3];
    check_direction(ray->d, world_d);
    const float *world_o = ray->o;
    float ray_o[3] = {world_o[0], world_o[1], world_o[2]};
    float ray_d[3] = {world_d[0], world_d[1], world_d[2]};
    float ray_mint = any ? 0.0f : ray->t_min; /* :1907 vs :2039 */
    float ray_maxt = ray->t_max;
    float inv_d[3];
    orc_safe_invdir(ray_d, inv_d);

    uint32_t stack[ORC_STACK];
    int32_t sp = 1;
    stack[sp - 1] = ORC_INVALID_NODE;
    uint32_t max_sp = 1;
    int32_t current_instance = -1, closest_instance = -1;
    uint32_t closest_prim = ORC_INVALID_NODE;
    float hit_u = 0.0f, hit_v = 0.0f;
    uint32_t node_index = 1, blas_offset = 0;

    while (node_index != ORC_INVALID_NODE) {
        const orc_node2 *nd = current_instance < 0 ? &tl->nodes[node_index - 1] : &tl->all_blas_nodes[blas_offset + node_index - 1];
        if (cnt) cnt->nodes++;
        int is_leaf = nd->child0 == ORC_INVALID_NODE;
        if (!is_leaf) {
            uint32_t near_c, far_c;
            intersect_internal_node(nd, inv_d, ray_o, ray_mint, ray_maxt, &near_c, &far_c);
            if (cnt) cnt->box_tests += 2;
            if (far_c != ORC_INVALID_NODE) {
                sp++;
                if (sp > ORC_STACK) { fprintf(stderr, "oracle: traversal stack overflow\n"); abort(); }
                stack[sp - 1] = far_c;
                if ((uint32_t)sp > max_sp) max_sp = (uint32_t)sp;
            }
            if (near_c != ORC_INVALID_NODE) { node_index = near_c; continue; }
        } else if (current_instance < 0) {
            current_instance = (int32_t)nd->child1;
            sp++;
            if (sp > ORC_STACK) { fprintf(stderr, "oracle: traversal stack overflow\n"); abort(); }
            stack[sp - 1] = ORC_TOP_LEVEL_SENTINEL;
            if ((uint32_t)sp > max_sp) max_sp = (uint32_t)sp;
            node_index = 1;
            const orc_instance *inst = &tl->instances[current_instance];
            blas_offset = tl->descs[inst->blas_index - 1].nodes_offset;
            orc_transform_point(inst->inv_transform, world_o, ray_o);
            orc_transform_direction(inst->inv_transform, world_d, ray_d);
            orc_safe_invdir(ray_d, inv_d);
            if (cnt) cnt->inst_entries++;
            continue;
        } else {
            float t, u, v;
            if (cnt) cnt->tri_tests++;
            if (wt ? orc_intersect_triangle_watertight(ray_o, ray_d, nd->aabb0_min, nd->aabb0_max, nd->aabb1_min, ray_mint, ray_maxt, &t, &u, &v)
                   : orc_intersect_triangle(ray_o, ray_d, nd->aabb0_min, nd->aabb0_max, nd->aabb1_min, ray_mint, ray_maxt, &t, &u, &v)) {
                if (any) { /* :2106-2115 */
                    const orc_instance *inst = &tl->instances[current_instance];
                    const orc_tri *tri = &tl->all_blas_prims[tl->descs[inst->blas_index - 1].primitives_offset + nd->child1 - 1];
                    out->hit = 1; out->t = t; out->bary_u = u; out->bary_v = v;
                    out->primitive_id = tri->input_index; out->meta = tri->metadata;
                    out->instance_id = (uint32_t)current_instance;
                    out->instance_custom_index = inst->instance_id;
                    if (flat_prim) *flat_prim = (uint32_t)(tri - tl->all_blas_prims);
                    if (cnt && max_sp > cnt->max_stack) cnt->max_stack = max_sp;
                    return;
                }
                ray_maxt = t;
                closest_instance = current_instance;
                closest_prim = nd->child1;
                hit_u = u; hit_v = v;
            }
        }
        node_index = stack[sp - 1];
        sp--;
        if (node_index == ORC_TOP_LEVEL_SENTINEL) {
            node_index = stack[sp - 1];
            sp--;
            current_instance = -1;
            memcpy(ray_o, world_o, 12); memcpy(ray_d, world_d, 12);
            orc_safe_invdir(ray_d, inv_d);
        }
    }
    if (cnt && max_sp > cnt->max_stack) cnt->max_stack = max_sp;
    if (!any && closest_instance >= 0) { /* :2010-2017 */
        const orc_instance *inst = &tl->instances[closest_instance];
        const orc_tri *tri = &tl->all_blas_prims[tl->descs[inst->blas_index - 1].primitives_offset + closest_prim - 1];
        out->hit = 1; out->t = ray_maxt; out->bary_u = hit_u; out->bary_v = hit_v;
        out->primitive_id = tri->input_index; out->meta = tri->metadata;
        out->instance_id = (uint32_t)closest_instance;
        out->instance_custom_index = inst->instance_id;
        if (flat_prim) *flat_prim = (uint32_t)(tri - tl->all_blas_prims);
    }
}

void orc_closest_hit(const orc_tlas *t, const orc_ray *ray, orc_hit *out, orc_counters *c) { traverse(t, ray, 0, out, c, NULL); }
void orc_any_hit(const orc_tlas *t, const orc_ray *ray, orc_hit *out, orc_counters *c) { traverse(t, ray, 1, out, c, NULL); }

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static void trace_batch(const orc_tlas *t, const orc_ray *rays, orc_hit *hits, uint64_t n, int threads, int any, orc_counters *sum) {
    if (threads <= 0) threads = orc_max_threads();
    uint64_t nodes = 0, boxes = 0, tris = 0, insts = 0;
    uint32_t mstack = 0;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 4096) reduction(+ : nodes, boxes, tris, insts) reduction(max : mstack)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        if (sum) {
            orc_counters c = {0, 0, 0, 0, 0};
            traverse(t, &rays[i], any, &hits[i], &c, NULL);
            nodes += c.nodes; boxes += c.box_tests; tris += c.tri_tests; insts += c.inst_entries;
            if (c.max_stack > mstack) mstack = c.max_stack;
        } else {
            traverse(t, &rays[i], any, &hits[i], NULL, NULL);
        }
    }
    if (sum) { sum->nodes = nodes; sum->box_tests = boxes; sum->tri_tests = tris; sum->inst_entries = insts; sum->max_stack = mstack; }
}
void orc_trace_closest(const orc_tlas *t, const orc_ray *rays, orc_hit *hits, uint64_t n, int threads, orc_counters *sum) {
    trace_batch(t, rays, hits, n, threads, 0, sum);
}
void orc_trace_any(const orc_tlas *t, const orc_ray *rays, orc_hit *hits, uint64_t n, int threads, orc_counters *sum) {
    trace_batch(t, rays, hits, n, threads, 1, sum);
}
void orc_trace_mode(const orc_tlas *t, const orc_ray *rays, orc_hit *hits, uint64_t n, int threads, int mode, orc_counters *sum) {
    trace_batch(t, rays, hits, n, threads, mode, sum);
}

/* ------------------------------------------------------------------ analysis, src/kernels.jl */
void orc_generate_ray_grid(const float b[6], const float dir_in[3], uint32_t grid, float *origins, float dir_out[3]) { /* :10-56 */
    float d0[3], direction[3];
    v_normalize(dir_in, d0);       /* hits_from_grid :59 */
    v_normalize(d0, direction);    /* generate_ray_grid :11 */
    /* the rays use the once-normalised direction (:59,66); the grid frame uses the twice-normalised one */
    memcpy(dir_out, d0, 12);
    float corners[8][3];
    for (int c = 1; c <= 8; c++) corner(b, c, corners[c - 1]); /* GB.decompose(Point3f, Rect3f): the 8 box corners; only extrema are used */
    float temp[3] = {1.0f, 0.0f, 0.0f};
    if (!(fabsf(direction[0]) < 0.9f)) { temp[0] = 0.0f; temp[1] = 1.0f; }
    float c1[3], c2[3], basis1[3], basis2[3];
    v_cross(direction, temp, c1); v_normalize(c1, basis1);
    v_cross(direction, basis1, c2); v_normalize(c2, basis2);
    float min1 = INFINITY, max1 = -INFINITY, min2 = INFINITY, max2 = -INFINITY, mind = INFINITY;
    for (int c = 0; c < 8; c++) {
        float p1 = v_dot(corners[c], basis1), p2 = v_dot(corners[c], basis2), pd = v_dot(corners[c], direction);
        min1 = jl_min(min1, p1); max1 = jl_max(max1, p1);
        min2 = jl_min(min2, p2); max2 = jl_max(max2, p2);
        mind = jl_min(mind, pd);
    }
    float margin = 0.05f * jl_max(max1 - min1, max2 - min2);
    float grid_width = max1 - min1 + 2 * margin;
    float grid_height = max2 - min2 + 2 * margin;
    float min_depth = mind - margin;
    float gc[3];
    float h1 = (min1 + max1) / 2, h2 = (min2 + max2) / 2;
    for (int k = 0; k < 3; k++) gc[k] = ((0.0f + min_depth * direction[k]) + h1 * basis1[k]) + h2 * basis2[k];
    float cell_w = grid_width / (float)grid; /* Float32 / Int -> Float32 */
    float cell_h = grid_height / (float)grid;
    double half = ((double)grid + 1.0) / 2.0;
    for (uint32_t i = 1; i <= grid; i++)
        for (uint32_t j = 1; j <= grid; j++) {
            double u = ((double)i - half) * (double)cell_w; /* Float64 via (grid_size + 1) / 2 */
            double v = ((double)j - half) * (double)cell_h;
            float *o = origins + 3 * ((size_t)(j - 1) * grid + (i - 1));
            for (int k = 0; k < 3; k++) o[k] = (float)(((double)gc[k] + u * (double)basis1[k]) + v * (double)basis2[k]);
        }
}

void orc_hits_from_grid(const orc_tlas *t, const float dir[3], uint32_t grid, orc_hit *hits, float *points, int threads) { /* :58-72 */
    size_t n = (size_t)grid * grid;
    float *origins = (float *)malloc(sizeof(float) * 3 * n);
    float d[3];
    float bb[6];
    memcpy(bb, t->root_aabb, 24);
    orc_generate_ray_grid(bb, dir, grid, origins, d);
    if (threads <= 0) threads = orc_max_threads();
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1024)
    for (int64_t k = 0; k < (int64_t)n; k++) {
        orc_ray r = {{origins[3 * k], origins[3 * k + 1], origins[3 * k + 2]}, 0.0f, {d[0], d[1], d[2]}, INFINITY};
        uint32_t fp = 0;
        traverse(t, &r, 0, &hits[k], NULL, &fp);
        if (points) {
            float *p = points + 3 * k;
            if (hits[k].hit) { /* sum_mul(bary, prim.vertices), math.jl:52; bary = (1-u-v, u, v) :2015 */
                const orc_tri *tri = &t->all_blas_prims[fp];
                float w = 1.0f - hits[k].bary_u - hits[k].bary_v;
                for (int c = 0; c < 3; c++) p[c] = (w * tri->v[c] + hits[k].bary_u * tri->v[3 + c]) + hits[k].bary_v * tri->v[6 + c];
            } else {
                p[0] = p[1] = p[2] = 0.0f; /* zero bary * empty_triangle vertices */
            }
        }
    }
    free(origins);
}

void orc_get_illumination(const orc_tlas *t, const float dir[3], uint32_t grid, float *out, int threads) { /* :112-124 */
    size_t n = (size_t)grid * grid;
    orc_hit *hits = (orc_hit *)malloc(sizeof(orc_hit) * n);
    orc_hits_from_grid(t, dir, grid, hits, NULL, threads);
    for (uint32_t i = 0; i < t->n_blas_prims; i++) out[i] = 0.0f;
    for (size_t k = 0; k < n; k++)
        if (hits[k].hit && hits[k].meta >= 1 && hits[k].meta <= t->n_blas_prims) out[hits[k].meta - 1] += 1.0f;
    free(hits);
}

uint32_t orc_get_centroid(const orc_tlas *t, const float dir[3], uint32_t grid, float centroid[3], int threads) { /* :106-110 */
    size_t n = (size_t)grid * grid;
    orc_hit *hits = (orc_hit *)malloc(sizeof(orc_hit) * n);
    float *pts = (float *)malloc(sizeof(float) * 3 * n);
    orc_hits_from_grid(t, dir, grid, hits, pts, threads);
    /* Statistics.mean(::Vector{Point3f}) = sum / n ; sum is pairwise in Julia, so compare with tolerance */
    double s[3] = {0, 0, 0};
    uint32_t c = 0;
    for (size_t k = 0; k < n; k++)
        if (hits[k].hit) { s[0] += pts[3 * k]; s[1] += pts[3 * k + 1]; s[2] += pts[3 * k + 2]; c++; }
    for (int k = 0; k < 3; k++) centroid[k] = c ? (float)(s[k] / c) : NAN;
    free(hits); free(pts);
    return c;
}

/* counter-based RNG shared (by specification, not by code) with the CUDA library: DESIGN.md "RNG" */
float orc_rng_uniform(uint64_t seed, uint64_t index, uint32_t dim) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (index * 4ull + (uint64_t)dim + 1ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (float)(z >> 40) * (1.0f / 16777216.0f);
}

static void vf_make_ray(const orc_tri *tri, uint64_t seed, uint64_t ray_index, orc_ray *r) {
    /* kernels.jl:84-92; math.jl:125-174 */
    const float *p1 = tri->v, *p2 = tri->v + 3, *p3 = tri->v + 6;
    float a[3], b[3], nrm[3], normal[3];
    v_sub(p2, p1, a); v_sub(p3, p1, b);
    v_cross(a, b, nrm);            /* GB.orthogonal_vector(Vec3f, Triangle) ∝ (v2-v1)x(v3-v1) (third-party, unpinned) */
    v_normalize(nrm, normal);
    /* get_orthogonal_basis math.jl:143-156 */
    float n[3];
    v_normalize(normal, n);
    int mi = 0;
    float best = fabsf(normal[0]);
    for (int k = 1; k < 3; k++) if (fabsf(normal[k]) < best) { best = fabsf(normal[k]); mi = k; } /* argmin: first minimum */
    float cand[3] = {0, 0, 0};
    cand[mi] = 1.0f;
    float t1[3], vv[3], t2[3], uu[3];
    v_cross(n, cand, t1); v_normalize(t1, vv);
    v_cross(vv, n, t2); v_normalize(t2, uu);
    /* random_triangle_point math.jl:158-174 */
    float r1 = orc_rng_uniform(seed, ray_index, 0), r2 = orc_rng_uniform(seed, ray_index, 1);
    float sq = sqrtf(r1);
    float bu = 1 - sq, bv = sq * (1 - r2), bw = sq * r2;
    float pt[3];
    for (int k = 0; k < 3; k++) pt[k] = (bu * p1[k] + bv * p2[k]) + bw * p3[k];
    for (int k = 0; k < 3; k++) r->o[k] = pt[k] + normal[k] * 0.01f; /* :91 */
    /* random_hemisphere_uniform math.jl:125-141 */
    float xi1 = orc_rng_uniform(seed, ray_index, 2), xi2 = orc_rng_uniform(seed, ray_index, 3);
    float theta = acosf(xi1);
    float phi = 2.0f * 3.14159265358979323846f * xi2;
    float xl = sinf(theta) * cosf(phi), yl = sinf(theta) * sinf(phi), zl = cosf(theta);
    for (int k = 0; k < 3; k++) r->d[k] = (uu[k] * xl + vv[k] * yl) + normal[k] * zl;
    r->t_min = 0.0f;
    r->t_max = INFINITY;
}

void orc_view_factors(const orc_tlas *t, uint32_t rpt, uint64_t seed, uint32_t row_base, uint32_t n_rows,
                      uint32_t *result, orc_ray *rays_out, int threads) { /* :80-104 */
    if (threads <= 0) threads = orc_max_threads();
    uint32_t n_cols = t->n_blas_prims;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 16)
    for (int64_t s = 0; s < (int64_t)t->n_blas_prims; s++) {
        const orc_tri *tri = &t->all_blas_prims[s];
        uint32_t tri_idx = tri->metadata;
        if (tri_idx < 1 || tri_idx > n_cols) continue; /* reference: unchecked index (:85) */
        uint32_t row = tri_idx - 1;
        if (row < row_base || row >= row_base + n_rows) continue;
        for (uint32_t i = 0; i < rpt; i++) {
            orc_ray r;
            orc_hit h;
            vf_make_ray(tri, seed, (uint64_t)row * rpt + i, &r);
            if (rays_out) rays_out[(size_t)(row - row_base) * rpt + i] = r;
            traverse(t, &r, 0, &h, NULL, NULL);
            if (h.hit && h.meta != tri_idx && h.meta >= 1 && h.meta <= n_cols) result[(size_t)(row - row_base) * n_cols + (h.meta - 1)] += 1u;
        }
    }
}

void orc_view_factors_from_rays(const orc_tlas *t, const orc_ray *rays, uint32_t rpt, uint32_t row_base, uint32_t n_rows,
                                uint32_t *result, int threads) {
    if (threads <= 0) threads = orc_max_threads();
    uint32_t n_cols = t->n_blas_prims;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 16)
    for (int64_t row = row_base; row < (int64_t)row_base + n_rows; row++) {
        uint32_t tri_idx = (uint32_t)row + 1;
        for (uint32_t i = 0; i < rpt; i++) {
            orc_hit h;
            const orc_ray *r = &rays[(size_t)(row - row_base) * rpt + i];
            if (r->d[0] == 0.0f && r->d[1] == 0.0f && r->d[2] == 0.0f) continue; /* row without a source */
            traverse(t, r, 0, &h, NULL, NULL);
            if (h.hit && h.meta != tri_idx && h.meta >= 1 && h.meta <= n_cols) result[(size_t)(row - row_base) * n_cols + (h.meta - 1)] += 1u;
        }
    }
}

/* ------------------------------------------------------------------ collision, src/collision.jl */
static inline int aabb_overlaps(const float amin[3], const float amax[3], const float bmin[3], const float bmax[3]) { /* :49-51 */
    return amax[0] >= bmin[0] && amax[1] >= bmin[1] && amax[2] >= bmin[2] && amin[0] <= bmax[0] && amin[1] <= bmax[1] && amin[2] <= bmax[2];
}

/* one pass of collide_instances_kernel! (:81-156) for leaf position i (1-based); contacts == NULL: count only */
static uint32_t collide_leaf(const orc_tlas *t, uint32_t i, const uint32_t *incl_counts, orc_contact *contacts) {
    uint32_t n = t->n_instances;
    const orc_node2 *nodes = t->nodes;
    const orc_node2 *leaf = &nodes[(n - 1 + i) - 1];
    const float *a_min = leaf->aabb0_min, *a_max = leaf->aabb0_max;
    uint32_t instance_a = leaf->child1, count = 0;
    uint32_t stack[256];
    int sp = 0;
    uint32_t node_index = 1;
    for (;;) {
        const orc_node2 *nd = &nodes[node_index - 1];
        if (nd->child0 != ORC_INVALID_NODE) {
            int o0 = aabb_overlaps(a_min, a_max, nd->aabb0_min, nd->aabb0_max);
            int o1 = aabb_overlaps(a_min, a_max, nd->aabb1_min, nd->aabb1_max);
            if (o0 && o1) {
                if (sp >= 256) { fprintf(stderr, "oracle: collision stack overflow\n"); abort(); }
                stack[sp++] = nd->child1;
                node_index = nd->child0;
                continue;
            } else if (o0) { node_index = nd->child0; continue; }
            else if (o1) { node_index = nd->child1; continue; }
        } else {
            uint32_t instance_b = nd->child1;
            if (instance_b > instance_a && aabb_overlaps(a_min, a_max, nd->aabb0_min, nd->aabb0_max)) {
                count++;
                if (contacts) {
                    uint32_t write_idx = incl_counts[i - 1] - count + 1; /* :138 */
                    contacts[write_idx - 1].instance_a = instance_a + 1;
                    contacts[write_idx - 1].instance_b = instance_b + 1;
                }
            }
        }
        if (sp > 0) node_index = stack[--sp];
        else break;
    }
    return count;
}

/* collide_instances (:189-233).  counts (n entries, nullable) receives the inclusive prefix sums the reference keeps as its
 * cache; contacts (nullable) must hold the returned total.  Returns the number of contacts. */
uint64_t orc_collide_instances(const orc_tlas *t, uint32_t *counts, orc_contact *contacts) {
    uint32_t n = t->n_instances;
    if (n == 0) return 0;
    uint32_t *c = counts ? counts : (uint32_t *)malloc(sizeof(uint32_t) * n);
    for (uint32_t i = 1; i <= n; i++) c[i - 1] = collide_leaf(t, i, NULL, NULL);
    for (uint32_t i = 1; i < n; i++) c[i] += c[i - 1]; /* AK.accumulate!(+) :215 */
    uint64_t total = c[n - 1];
    if (contacts && total)
        for (uint32_t i = 1; i <= n; i++) collide_leaf(t, i, c, contacts);
    if (!counts) free(c);
    return total;
}

/* collide_instances_any (:241-261).  literal != 0 reproduces the reference verbatim: it indexes the Morton-sorted leaf array
 * with the *instance* index (nodes[n-1+ia]), which is only right when the sort is the identity; literal == 0 looks each
 * instance's own leaf up (the evident intent). */
int orc_collide_instances_any(const orc_tlas *t, uint32_t a_start, uint32_t a_count, uint32_t b_start, uint32_t b_count, int literal) {
    uint32_t n = t->n_instances;
    uint32_t *leaf_of = (uint32_t *)malloc(sizeof(uint32_t) * (n ? n : 1));
    for (uint32_t p = 1; p <= n; p++) {
        const orc_node2 *lf = &t->nodes[(n - 1 + p) - 1];
        leaf_of[literal ? p - 1 : lf->child1] = p;
    }
    int hit = 0;
    for (uint32_t ia = a_start; ia < a_start + a_count && !hit; ia++)
        for (uint32_t ib = b_start; ib < b_start + b_count && !hit; ib++) {
            const orc_node2 *la = &t->nodes[(n - 1 + leaf_of[ia]) - 1], *lb = &t->nodes[(n - 1 + leaf_of[ib]) - 1];
            hit = aabb_overlaps(la->aabb0_min, la->aabb0_max, lb->aabb0_min, lb->aabb0_max);
        }
    free(leaf_of);
    return hit;
}

/* ------------------------------------------------------------------ wavefront stages (docs/src/wavefront-renderer.jl) */
static void primary_uv(uint32_t width, uint32_t height, uint32_t n_samples, uint64_t seed, int jitter, uint64_t ray_idx, float *u, float *v) {
    uint64_t pixel = ray_idx / n_samples;
    float x = (float)(uint32_t)(pixel % width + 1), y = (float)(uint32_t)(pixel / width + 1);
    float j1 = jitter ? orc_rng_uniform(seed, ray_idx, 0) : 0.5f, j2 = jitter ? orc_rng_uniform(seed, ray_idx, 1) : 0.5f;
    *u = 2.0f * (x - 0.5f + j1) / (float)width - 1.0f;   /* :202 / :238 */
    *v = 1.0f - 2.0f * (y - 0.5f + j2) / (float)height;  /* :203 / :239 */
}

void orc_generate_primary_rays(uint32_t width, uint32_t height, uint32_t n_samples, const float camera_pos[3], float focal_length,
                               float aspect, uint64_t seed, int jitter, orc_ray *rays) { /* :185-213 */
    uint64_t total = (uint64_t)width * height * n_samples;
    for (uint64_t i = 0; i < total; i++) {
        float u, v;
        primary_uv(width, height, n_samples, seed, jitter, i, &u, &v);
        float d[3] = {u * aspect, v, focal_length}, dn[3];
        v_normalize(d, dn);
        orc_ray r = {{camera_pos[0], camera_pos[1], camera_pos[2]}, 0.0f, {dn[0], dn[1], dn[2]}, INFINITY};
        rays[i] = r;
    }
}

void orc_generate_primary_rays_lookat(uint32_t width, uint32_t height, uint32_t n_samples, const float camera_pos[3], const float right[3],
                                      const float up[3], const float forward[3], float half_width, float half_height, uint64_t seed,
                                      int jitter, orc_ray *rays) { /* :219-253 */
    uint64_t total = (uint64_t)width * height * n_samples;
    for (uint64_t i = 0; i < total; i++) {
        float u, v;
        primary_uv(width, height, n_samples, seed, jitter, i, &u, &v);
        float a = u * half_width, b = v * half_height, d[3], dn[3];
        for (int k = 0; k < 3; k++) d[k] = (forward[k] + right[k] * a) + up[k] * b; /* :242-246 */
        v_normalize(d, dn);
        orc_ray r = {{camera_pos[0], camera_pos[1], camera_pos[2]}, 0.0f, {dn[0], dn[1], dn[2]}, INFINITY};
        rays[i] = r;
    }
}

void orc_generate_shadow_rays(const orc_tlas *t, const orc_ray *rays, const orc_hit *hits, uint64_t n, const float *const *blas_normals,
                              const float *lights, uint32_t n_lights, float shadow_bias, orc_ray *shadow_rays) { /* :277-330 */
    /* primitive_id -> sorted position, per BLAS (only needed for the geometric-normal fallback) */
    uint32_t **inv = (uint32_t **)calloc(t->n_blas ? t->n_blas : 1, sizeof(uint32_t *));
    for (uint64_t k = 0; k < n; k++) {
        const orc_hit *h = &hits[k];
        orc_ray *out = shadow_rays + k * n_lights;
        if (!h->hit) {
            orc_ray dummy = {{0, 0, 0}, 0.0f, {0, 0, 1}, 0.0f}; /* :319 */
            for (uint32_t l = 0; l < n_lights; l++) out[l] = dummy;
            continue;
        }
        const orc_instance *inst = &t->instances[h->instance_id];
        uint32_t b = inst->blas_index - 1;
        float n9[9];
        if (blas_normals && blas_normals[b]) {
            memcpy(n9, blas_normals[b] + (size_t)h->primitive_id * 9, 36);
        } else {
            uint32_t cnt = (b + 1 < t->n_blas ? t->descs[b + 1].primitives_offset : t->n_blas_prims) - t->descs[b].primitives_offset;
            const orc_tri *prims = t->all_blas_prims + t->descs[b].primitives_offset;
            if (!inv[b]) {
                inv[b] = (uint32_t *)malloc(sizeof(uint32_t) * (cnt ? cnt : 1));
                for (uint32_t p = 0; p < cnt; p++) inv[b][prims[p].input_index] = p;
            }
            const orc_tri *tri = &prims[inv[b][h->primitive_id]];
            float e1[3], e2[3], c[3], g[3];
            v_sub(tri->v + 3, tri->v, e1); v_sub(tri->v + 6, tri->v, e2);
            v_cross(e1, e2, c);
            v_normalize(c, g);
            for (int q = 0; q < 3; q++) memcpy(n9 + 3 * q, g, 12);
        }
        const orc_ray *ray = &rays[k];
        float w0 = 1.0f - h->bary_u - h->bary_v, w1 = h->bary_u, w2 = h->bary_v; /* bary = (1-u-v, u, v), :2015-2016 */
        float nl[3], nw[3], nn[3], so[3];
        for (int c = 0; c < 3; c++) nl[c] = (n9[c] * w0 + n9[3 + c] * w1) + n9[6 + c] * w2; /* :292 */
        const float *m = inst->inv_transform;
        for (int c = 0; c < 3; c++) nw[c] = (m[c] * nl[0] + m[4 + c] * nl[1]) + m[8 + c] * nl[2];
        v_normalize(nw, nn);
        for (int c = 0; c < 3; c++) so[c] = (ray->o[c] + ray->d[c] * h->t) + nn[c] * shadow_bias; /* :289, :306 */
        for (uint32_t l = 0; l < n_lights; l++) {
            float lv[3], sd[3];
            v_sub(lights + 3 * l, so, lv);
            v_normalize(lv, sd);
            orc_ray r = {{so[0], so[1], so[2]}, 0.0f, {sd[0], sd[1], sd[2]}, sqrtf(lv[0] * lv[0] + lv[1] * lv[1] + lv[2] * lv[2])};
            out[l] = r;
        }
    }
    for (uint32_t b = 0; b < t->n_blas; b++) free(inv[b]);
    free(inv);
}

void orc_test_shadow_rays(const orc_tlas *t, const orc_ray *shadow_rays, uint64_t n, uint8_t *visible, int threads) { /* :337-362 */
    if (threads <= 0) threads = orc_max_threads();
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1024)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        if (shadow_rays[i].t_max > 0.0f) {
            orc_hit h;
            traverse(t, &shadow_rays[i], 1, &h, NULL, NULL);
            visible[i] = h.hit ? 0 : 1;
        } else {
            visible[i] = 0; /* dummy ray of a sky hit */
        }
    }
}
