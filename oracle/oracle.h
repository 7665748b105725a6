/*
 * oracle.h — CPU restatement of Raycore.jl's ray-query hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked, imported or executed by
 * the product (raycore.jl_b200/, libraycore_cuda.so).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may use it, and only as the
 * checker / the timed CPU baseline.
 *
 * Parity status: the reference is pure Julia and Julia is not installed in this image, so
 * the reference itself cannot be run.  The oracle is pinned against every known-answer
 * value in the reference's own test-suite for this path (tests/test_oracle_kat.py lists them
 * with file:line); ulp-level arithmetic of third-party packages (StaticArrays `inv`, `dot`,
 * `cross`, GeometryBasics `normalize`) is restated from their published algorithms and is
 * "parity unpinned" below the tolerance those tests state.
 *
 * All indices follow the reference: nodes / prims / BLAS indices 1-based in memory,
 * descriptor offsets 0-based, TLAS leaf child1 = 0-based instance index.
 */
#ifndef RAYCORE_ORACLE_H
#define RAYCORE_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_INVALID_NODE 0xFFFFFFFFu        /* src/instanced-bvh.jl:65   */
#define ORC_TOP_LEVEL_SENTINEL 0xFFFFFFFEu  /* src/instanced-bvh.jl:1733 */

/* BVHNode2, 60 bytes, src/instanced-bvh.jl:50-63 */
typedef struct {
    float aabb0_min[3], aabb0_max[3];
    float aabb1_min[3], aabb1_max[3];
    uint32_t child0, child1, parent;
} orc_node2;

/* InstanceDescriptor, 108 bytes, src/instanced-bvh.jl:90-96.
 * transform / inv_transform are Mat3x4f = 12 floats, memory order = Vulkan row-major 3x4:
 * floats 4i..4i+3 = row i of [R|t]  (src/instanced-bvh.jl:28-31). */
typedef struct {
    uint32_t blas_index;   /* 1-based */
    uint32_t instance_id;
    float transform[12];
    float inv_transform[12];
    uint32_t flags;
} orc_instance;

/* BLASDescriptor, 32 bytes, src/instanced-bvh.jl:132-136 */
typedef struct {
    uint32_t nodes_offset, primitives_offset; /* 0-based */
    float root_aabb[6];                       /* p_min, p_max */
} orc_blas_desc;

/* What the oracle keeps of Triangle{UInt32} (src/triangle_mesh.jl:1-7): the vertices, the
 * metadata and the position in the caller's (already degenerate-filtered) input list. */
typedef struct {
    float v[9];
    uint32_t metadata;
    uint32_t input_index; /* 0-based position in the filtered input order */
} orc_tri;

/* Ray, src/ray.jl:1-7 (time omitted) == RTRay byte layout, src/rt_transport.jl:10-19 */
typedef struct {
    float o[3];
    float t_min;
    float d[3];
    float t_max;
} orc_ray;

/* RTHitResult, 32 bytes, src/rt_transport.jl:33-42.  `meta` occupies the reference's pad
 * word and carries Triangle.metadata of the hit primitive. */
typedef struct {
    uint32_t hit;
    float t;
    uint32_t primitive_id;          /* input_index of the hit triangle within its BLAS */
    uint32_t instance_custom_index; /* InstanceDescriptor.instance_id */
    float bary_u, bary_v;
    uint32_t instance_id;           /* 0-based position in instances[] */
    uint32_t meta;
} orc_hit;

typedef struct {
    uint64_t nodes;      /* node fetches */
    uint64_t box_tests;  /* child-box slab tests (2 per interior node) */
    uint64_t tri_tests;  /* triangle tests */
    uint64_t inst_entries;
    uint32_t max_stack;
} orc_counters;

typedef struct orc_blas {
    uint32_t n;          /* primitives */
    orc_node2 *nodes;    /* 2n-1 */
    orc_tri *prims;      /* Morton-sorted */
    uint32_t *morton;    /* sorted codes (kept for tests) */
    float root_aabb[6];
} orc_blas;

/* StaticTLAS, src/instanced-bvh.jl:155-168 */
typedef struct orc_tlas {
    uint32_t n_nodes, n_instances, n_blas;
    uint32_t n_blas_nodes, n_blas_prims;
    orc_node2 *nodes;
    orc_instance *instances;
    orc_node2 *all_blas_nodes;
    orc_tri *all_blas_prims;
    orc_blas_desc *descs;
    float root_aabb[6];
} orc_tlas;

/* ---- scalar helpers (KAT targets) ---- */
uint32_t orc_expand_bits(uint32_t x);                       /* instanced-bvh.jl:1177-1183 */
uint32_t orc_morton_code_30bit(const float p[3]);           /* :1189-1200 */
int32_t orc_clz32(uint32_t x);                              /* :1203-1206 */
int32_t orc_delta(int32_t i1, int32_t i2, const uint32_t *codes, int32_t n); /* :1212-1229 */
int orc_is_degenerate(const float v[9]);                    /* triangle_mesh.jl:14-17 */
void orc_mat4_to_mat3x4(const float m4_colmajor[16], float out[12]);          /* :1663-1669 */
void orc_mat3x4_inverse(const float m[12], float out[12]);  /* :1675-1687 */
void orc_transform_point(const float m[12], const float p[3], float out[3]);  /* :1692-1698 */
void orc_transform_direction(const float m[12], const float v[3], float out[3]); /* :1711-1717 */
void orc_safe_invdir(const float d[3], float out[3]);       /* :1742-1748 */
/* fast_intersect_triangle :1756-1797; returns 1 on accept */
int orc_intersect_triangle(const float o[3], const float d[3], const float v0[3], const float v1[3],
                           const float v2[3], float t_min, float closest_t, float *t, float *u, float *v);
/* fast_intersect_bbox :1841-1859 */
void orc_intersect_bbox(const float o[3], const float inv_d[3], const float pmin[3], const float pmax[3],
                        float t_min, float t_max, float *out_min, float *out_max);

/* ---- builders ---- */
/* filter (is_degenerate_face, instanced-bvh.jl:573-577, 593-600) + metadata assignment.
 * verts: n_faces*9; face_meta: NULL => metadata = 1-based face index before filtering (:595).
 * Returns number kept; out must hold n_faces entries. */
uint32_t orc_filter_triangles(const float *verts, uint32_t n_faces, const uint32_t *face_meta, orc_tri *out);
orc_blas *orc_build_blas(const orc_tri *tris, uint32_t n);  /* :1376-1443 */
void orc_free_blas(orc_blas *b);
orc_tlas *orc_build_tlas(orc_blas *const *blas, uint32_t n_blas, const orc_instance *inst, uint32_t n_inst); /* :1605-1651 */
void orc_free_tlas(orc_tlas *t);
/* refit_tlas! :2197-2222 (update_tlas_leaf_aabbs + refit_tlas_aabbs); instances already updated */
void orc_refit_tlas(orc_tlas *t);

/* ---- queries ---- */
void orc_closest_hit(const orc_tlas *t, const orc_ray *ray, orc_hit *out, orc_counters *c); /* :1902-2024 */
void orc_any_hit(const orc_tlas *t, const orc_ray *ray, orc_hit *out, orc_counters *c);     /* :2034-2140 */
/* OpenMP over rays, as the reference's Threads.@threads over rays (src/kernels.jl:64,82) */
void orc_trace_closest(const orc_tlas *t, const orc_ray *rays, orc_hit *hits, uint64_t n, int threads, orc_counters *sum);
void orc_trace_any(const orc_tlas *t, const orc_ray *rays, orc_hit *hits, uint64_t n, int threads, orc_counters *sum);
/* mode: bit 0 = any_hit, bit 1 = watertight triangle test (src/triangle_mesh.jl:168-201) instead of Moeller-Trumbore */
void orc_trace_mode(const orc_tlas *t, const orc_ray *rays, orc_hit *hits, uint64_t n, int threads, int mode, orc_counters *sum);
int orc_intersect_triangle_watertight(const float o[3], const float d[3], const float v0[3], const float v1[3], const float v2[3],
                                      float t_min, float t_max, float *t_out, float *u_out, float *v_out); /* triangle_mesh.jl:168-201 */
int orc_max_threads(void);

/* ---- analysis (src/kernels.jl) ---- */
/* generate_ray_grid :10-56. dir is normalised inside (hits_from_grid :59 + :11).
 * origins: grid*grid*3 floats, element (i,j) 1-based at ((j-1)*grid + (i-1)) (Julia column-major).
 * dir_out: the normalised direction used for every ray. */
void orc_generate_ray_grid(const float bounds[6], const float dir[3], uint32_t grid, float *origins, float dir_out[3]);
/* hits_from_grid :58-72: hits[k] for cell k in the order above; points = sum_mul(bary, vertices) */
void orc_hits_from_grid(const orc_tlas *t, const float dir[3], uint32_t grid, orc_hit *hits, float *points, int threads);
/* get_illumination :112-124: out[idx-1] = number of grid hits whose metadata == idx, idx in 1..n_prims */
void orc_get_illumination(const orc_tlas *t, const float dir[3], uint32_t grid, float *out, int threads);
/* get_centroid :106-110: returns count of hit points, writes mean into centroid */
uint32_t orc_get_centroid(const orc_tlas *t, const float dir[3], uint32_t grid, float centroid[3], int threads);

/* view_factors! :80-104 with the counter-based RNG of DESIGN.md (the reference's task-local rand() stream is
 * not reproducible).  Sources = primitives of all_blas_prims whose metadata-1 lies in [row_base, row_base+n_rows);
 * result is the row-major row block: result[(meta_src-1-row_base)*n_prims + (meta_hit-1)] (= transpose of Julia's
 * column-major result[src,hit]).  Ray index for the RNG = (meta_src-1)*rays_per_triangle + i.
 * rays_out (nullable): the generated rays at (meta_src-1-row_base)*rays_per_triangle + i. */
void orc_view_factors(const orc_tlas *t, uint32_t rays_per_triangle, uint64_t seed, uint32_t row_base, uint32_t n_rows,
                      uint32_t *result, orc_ray *rays_out, int threads);
/* accumulate caller-supplied rays (e.g. the CUDA library's own generated rays) laid out as above */
void orc_view_factors_from_rays(const orc_tlas *t, const orc_ray *rays, uint32_t rays_per_triangle, uint32_t row_base,
                                uint32_t n_rows, uint32_t *result, int threads);
/* ---- collision (src/collision.jl) ---- */
typedef struct { uint32_t instance_a, instance_b; } orc_contact; /* ContactPair :25-28, 1-based instance indices, a < b */
uint64_t orc_collide_instances(const orc_tlas *t, uint32_t *counts, orc_contact *contacts);  /* :189-233 */
/* ranges are 0-based [start, start+count) positions in instances[] */
int orc_collide_instances_any(const orc_tlas *t, uint32_t a_start, uint32_t a_count, uint32_t b_start, uint32_t b_count, int literal); /* :241-261 */

/* ---- wavefront stages (docs/src/wavefront-renderer.jl) ---- */
/* generate_primary_rays! :185-213 — ray ((y-1)*width + (x-1))*n_samples + (s-1); jitter != 0: rand(Vec2f) restated with the
 * counter RNG (seed, ray index, dims 0/1), jitter == 0: pixel centres (0.5, 0.5) */
void orc_generate_primary_rays(uint32_t width, uint32_t height, uint32_t n_samples, const float camera_pos[3], float focal_length,
                               float aspect, uint64_t seed, int jitter, orc_ray *rays);
/* generate_primary_rays_lookat! :219-253 */
void orc_generate_primary_rays_lookat(uint32_t width, uint32_t height, uint32_t n_samples, const float camera_pos[3], const float right[3],
                                      const float up[3], const float forward[3], float half_width, float half_height, uint64_t seed,
                                      int jitter, orc_ray *rays);
/* generate_shadow_rays! :277-330.  blas_normals[b] (nullable table / nullable entries): 9 floats per primitive of BLAS b+1 indexed
 * by hit.primitive_id (= orc_tri.input_index); NULL => the triangle's geometric normal at all three vertices.  The interpolated
 * normal is carried through the instance's inverse-transpose (identity for the reference renderer's identity instances). */
void orc_generate_shadow_rays(const orc_tlas *t, const orc_ray *rays, const orc_hit *hits, uint64_t n, const float *const *blas_normals,
                              const float *lights, uint32_t n_lights, float shadow_bias, orc_ray *shadow_rays);
/* test_shadow_rays! :337-362 */
void orc_test_shadow_rays(const orc_tlas *t, const orc_ray *shadow_rays, uint64_t n, uint8_t *visible, int threads);

/* the RNG itself, exposed so tests can pin GPU == oracle on the uniform stream */
float orc_rng_uniform(uint64_t seed, uint64_t index, uint32_t dim);

#ifdef __cplusplus
}
#endif
#endif
