/*
 * raycore_cuda.h — C ABI of libraycore_cuda.so, the B200 (sm_100a) implementation of
 * Raycore.jl's ray-query hot path.  This is the drop-in boundary: a Julia shim (or any FFI)
 * binds exactly these entry points; INTEGRATION.md shows the ccall stubs.
 *
 * Every entry point cites the reference interface it replaces (file:line in
 * JuliaGeometry/Raycore.jl v0.2.0).  Plain pointers and sizes only; all functions return an
 * int32 status (RC_OK == 0) and record a message retrievable with rc_last_error().
 *
 * There is no CPU fallback: every call needs a CUDA device and fails with RC_ERR_CUDA without one.
 *
 * Conventions
 *   - Mat3x4f: 12 floats, Vulkan row-major 3x4 [R|t]  (src/instanced-bvh.jl:28-31).
 *   - Triangle soups: n_faces x 9 floats (v0,v1,v2), i.e. the decomposed mesh the reference
 *     builds Triangles from (src/instanced-bvh.jl:555-566).  Degenerate faces are dropped by the
 *     library with the reference's exact rule (src/triangle_mesh.jl:14-17).
 *   - Indices returned in RTHitResult are 0-based (src/rt_transport.jl:26-31); the Julia shim
 *     adds 1 where closest_hit's tuple is 1-based (src/instanced-bvh.jl:2011).
 *   - Threading: every entry point that touches the GPU holds a per-context lock for the duration of the
 *     call, so queries on a synced context may be issued from several host threads (the reference calls
 *     closest_hit under Threads.@threads, src/kernels.jl:64,82) — they execute one after the other on the
 *     context's stream.  Mutation keeps the reference's contract: one mutating thread, no query in flight
 *     (the lock keeps the library's own state consistent, it does not make push!/sync! + trace a transaction).
 *     rc_last_error() returns the context's last message and is not per thread.
 */
#ifndef RAYCORE_CUDA_H
#define RAYCORE_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RC_ABI_VERSION 1

/* status codes */
enum {
    RC_OK = 0,
    RC_ERR_INVALID_ARGUMENT = 1, /* Julia: ArgumentError / ErrorException for arity mismatch (:664-666, :759, :788) */
    RC_ERR_INVALID_HANDLE = 2,   /* Julia: error("Invalid handle") (:715, :756, :785, :809) */
    RC_ERR_DELETED_HANDLE = 3,   /* Julia: error("Handle has been deleted") (:716, :757, :786, :810) */
    RC_ERR_NO_VALID_TRIANGLES = 4, /* Julia: error("Geometry has no valid triangles") (:601, :837) */
    RC_ERR_CUDA = 5,
    RC_ERR_NOT_SYNCED = 6,
    RC_ERR_OUT_OF_MEMORY = 7,
    RC_ERR_STACK_OVERFLOW = 8    /* traversal stack exhausted for >=1 ray (the reference has UB here, :1912) */
};

/* RTRay, 32 bytes — src/rt_transport.jl:10-19 */
typedef struct rc_ray {
    float origin[3];
    float tmin;
    float dir[3];
    float tmax;
} rc_ray;

/* RTHitResult, 32 bytes — src/rt_transport.jl:33-42.  The reference's trailing pad word carries
 * Triangle.metadata of the hit primitive (face_meta / 1-based face index, src/instanced-bvh.jl:595). */
typedef struct rc_hit {
    uint32_t hit;                   /* 0 / 1 */
    float t;                        /* miss: 0 (src/instanced-bvh.jl:2022) */
    uint32_t primitive_id;          /* 0-based index into the BLAS's degenerate-filtered input triangle list */
    uint32_t instance_custom_index; /* InstanceDescriptor.instance_id (src/instanced-bvh.jl:92) */
    float bary_u, bary_v;           /* bary = (1-u-v, u, v) (src/instanced-bvh.jl:2015-2016) */
    uint32_t instance_id;           /* 0-based position in instances[] */
    uint32_t metadata;
} rc_hit;

/* InstanceDescriptor, 108 bytes — src/instanced-bvh.jl:90-96 */
typedef struct rc_instance_desc {
    uint32_t blas_index; /* 1-based */
    uint32_t instance_id;
    float transform[12];
    float inv_transform[12];
    uint32_t flags;
} rc_instance_desc;

/* BVHNode2, 60 bytes — src/instanced-bvh.jl:50-63 (read-back format of rc_read_*_nodes) */
typedef struct rc_bvh_node2 {
    float aabb0_min[3], aabb0_max[3], aabb1_min[3], aabb1_max[3];
    uint32_t child0, child1, parent;
} rc_bvh_node2;

typedef struct rc_context rc_context; /* one mutable TLAS (src/instanced-bvh.jl:261-310) */

/* trace / pointer flags */
#define RC_RAYS_ON_DEVICE 0x1u  /* rays pointer is device memory */
#define RC_HITS_ON_DEVICE 0x2u  /* hits pointer is device memory */
#define RC_MODE_REFERENCE_ORDER 0x4u /* traverse the reference-identical BVH2 in the reference's own order
                                        (bit-identical results incl. ties); default = wide BVH4 fast path */
#define RC_MODE_WATERTIGHT 0x400u /* triangle test = the reference's watertight intersect_triangle (src/triangle_mesh.jl:168-201: dominant-axis
                                     permutation, shear, signed edge functions) instead of fast_intersect_triangle (Moeller-Trumbore, the
                                     reference's traversal default and this library's): no ray slips between two triangles that share an
                                     edge.  Hit / miss and ids can differ from the default mode at edges; t, u, v differ in the last bits.
                                     Combines with RC_MODE_REFERENCE_ORDER; not with RC_COUNTERS */
#define RC_IGNORE_TMIN 0x800u /* closest-hit trace: treat ray.tmin as 0 (what closest_hit4 does, src/bvh4.jl:610); default wide path only */
#define RC_COUNTERS 0x8u        /* accumulate per-ray work counters (rc_get_counters) — instrumented build of the same kernel */
#define RC_VERTS_ON_DEVICE 0x10u /* rc_push / rc_update_geometry: verts (and face_meta) are device pointers */
#define RC_NO_SYNC 0x20u        /* trace: do not cudaStreamSynchronize before returning (device buffers only) */
/* build flags (rc_push / rc_update_geometry flags, or per context with rc_set_build_flags) */
#define RC_BUILD_KEEP_BVH2 0x80u    /* also emit the reference-layout BVH2 (BVHNode2, src/instanced-bvh.jl:50-63) of the geometry: needed by
                                       RC_MODE_REFERENCE_ORDER and rc_read_blas_nodes; the default build writes only what the fast path reads */
#define RC_BUILD_ALLOW_REFIT 0x100u /* keep the radix-tree topology so a later vertex update can re-fit instead of rebuilding (24 B / triangle) */
#define RC_UPDATE_REFIT 0x200u      /* rc_update_geometry: re-fit the kept topology to the new vertex positions when possible (same face
                                       count, same degenerate faces, built with RC_BUILD_ALLOW_REFIT); otherwise rebuild */

/* sync actions reported by rc_sync */
enum { RC_SYNC_NONE = 0, RC_SYNC_REFIT = 1, RC_SYNC_REBUILD = 2 };

/* ---- lifecycle ----------------------------------------------------------------------- */
/* TLAS(backend) — src/instanced-bvh.jl:334-358.  device < 0: current device. */
int32_t rc_create(int32_t device, rc_context **out);
/* free!(tlas) — src/instanced-bvh.jl:383-399 */
int32_t rc_destroy(rc_context *ctx);
const char *rc_last_error(const rc_context *ctx); /* ctx may be NULL: the calling thread's last creation / blob-check error */
int32_t rc_abi_version(void);
/* the context's CUDA stream (cudaStream_t) — all work of this context is ordered on it */
void *rc_stream(rc_context *ctx);
/* adopt a caller-owned cudaStream_t (e.g. the host framework's current stream) for all subsequent work of this
 * context; NULL restores a private stream.  Lets callers bracket calls with their own CUDA events. */
int32_t rc_set_stream(rc_context *ctx, void *stream);

/* ---- mutation (handle API) ----------------------------------------------------------- */
/* push!(tlas, mesh, transform; instance_id) / push!(tlas, mesh, transforms; instance_ids)
 * — src/instanced-bvh.jl:639-676 (+ build_and_append_blas! :581-608, build_blas :1376-1443).
 * Builds the BLAS immediately on the GPU, appends m instance descriptors, marks the TLAS dirty.
 *   verts: n_faces*9 floats; face_meta: NULL => metadata = 1-based face index before filtering.
 *   transforms: m*12 floats; inv_transforms: NULL => mat3x4_inverse (:1675-1687) is evaluated by the
 *   library, else used verbatim (lets the Julia shim pass its own results for bit identity);
 *   instance_ids: NULL => 0 ("inherit", :656-657).  m must be >= 1.
 * Geometry (or, at rc_sync, a scene) whose extent exceeds 255 * 2^103 is refused with RC_ERR_INVALID_ARGUMENT: the wide
 * nodes' quantisation frame cannot cover it (and the reference's own single-precision triangle test overflows long before). */
int32_t rc_push(rc_context *ctx, const float *verts, uint32_t n_faces, const uint32_t *face_meta, const float *transforms,
                const float *inv_transforms, const uint32_t *instance_ids, uint32_t m, uint32_t flags, uint32_t *handle_out);
/* delete!(tlas, handle)::Bool — src/instanced-bvh.jl:690-699.  *deleted = 0 for unknown / already deleted. */
int32_t rc_delete(rc_context *ctx, uint32_t handle, int32_t *deleted);
/* update_transform!/update_transforms! — src/instanced-bvh.jl:755-797 (+ kernels.jl:434-476).
 * m must equal the handle's instance count. */
int32_t rc_update_transforms(rc_context *ctx, uint32_t handle, const float *transforms, const float *inv_transforms, uint32_t m);
/* the same with DEVICE-resident transform arrays (the `instance_buffer` use case of src/Raycore.jl:118-130: transforms written by
 * a caller's kernel); ordered after the work already enqueued on the context stream. */
int32_t rc_update_transforms_device(rc_context *ctx, uint32_t handle, const float *d_transforms, const float *d_inv_transforms, uint32_t m);
/* update!(tlas, handle, new_geometry) — src/instanced-bvh.jl:808-857: rebuild the handle's BLAS in place (the reference always
 * rebuilds).  With RC_UPDATE_REFIT and a geometry built with RC_BUILD_ALLOW_REFIT the library keeps the radix tree and only re-fits
 * boxes and wide nodes to the moved vertices (the refit kernel for mesh updates; test/test_mesh_update.jl:96-116 workload) when the face
 * count and the set of degenerate faces are unchanged — results then equal a fresh build's wherever the hit is unique (the tree is the
 * old frame's, so its quality degrades with large deformations).  *face_meta is ignored by a refit.  rc_last_update_refitted tells
 * which path the last call took. */
int32_t rc_update_geometry(rc_context *ctx, uint32_t handle, const float *verts, uint32_t n_faces, const uint32_t *face_meta, uint32_t flags);
int32_t rc_last_update_refitted(const rc_context *ctx);
/* build flags (RC_BUILD_*) OR-ed into every later rc_push / rc_update_geometry of this context */
int32_t rc_set_build_flags(rc_context *ctx, uint32_t flags);
/* sync!(tlas) — src/instanced-bvh.jl:894-921: no-op when clean, refit when only transforms changed,
 * else compact + rebuild; returns with the stream idle.  *action: RC_SYNC_*. */
int32_t rc_sync(rc_context *ctx, int32_t *action);

/* ---- serialised geometry -------------------------------------------
 * to_gpu(ArrayType, blas::BLAS) — src/kernel-abstractions.jl:31-36: a BLAS that was built earlier is moved to the device
 * instead of being rebuilt.  rc_export_geometry writes a handle's built geometry (reference-layout BVH2, wide nodes, sorted
 * triangles, hull boxes, normals if present) into a host blob; rc_push_exported restores byte-identical device arrays from it
 * (no builder kernels run) and appends instances exactly like rc_push, so traces of the restored geometry are bit-identical.
 * Two-call protocol: blob == NULL only reports *size.  Blobs carry a layout version and a payload hash; a blob from an
 * incompatible build, a truncated or damaged one, or one whose references leave its arrays is refused with
 * RC_ERR_INVALID_ARGUMENT (the hash is an integrity check, not authentication: import blobs you wrote). */
int32_t rc_export_geometry(rc_context *ctx, uint32_t handle, void *blob, uint64_t capacity, uint64_t *size);
int32_t rc_push_exported(rc_context *ctx, const void *blob, uint64_t size, const float *transforms, const float *inv_transforms,
                         const uint32_t *instance_ids, uint32_t m, uint32_t *handle_out);
/* The host-side part of those checks alone (magic, layout version, section table, size, payload hash, supported extent): needs neither a
 * context nor a GPU, e.g. to vet a file before a device is claimed.  Outputs are nullable; a refusal's message is rc_last_error(NULL). */
int32_t rc_check_exported(const void *blob, uint64_t size, uint32_t *n_triangles, uint32_t *n_faces, uint32_t *has_normals);

/* ---- introspection ------------------------------------------------------------------- */
int32_t rc_is_valid(const rc_context *ctx, uint32_t handle);                 /* :524-526 */
uint32_t rc_n_instances(const rc_context *ctx);                              /* live, :2391-2398 */
uint32_t rc_n_instances_of(const rc_context *ctx, uint32_t handle);          /* :533-537 */
uint32_t rc_n_total_instances(const rc_context *ctx);                        /* :544 */
uint32_t rc_n_geometries(const rc_context *ctx);                             /* :2405 */
int32_t rc_is_dirty(const rc_context *ctx, int32_t *dirty, int32_t *transforms_dirty);
/* get_instances(tlas, handle) — :732-738; out must hold rc_n_instances_of() entries */
int32_t rc_get_instances(const rc_context *ctx, uint32_t handle, rc_instance_desc *out);
/* world_bound(tlas) — :2147-2149; (p_min, p_max); Bounds3() = (+Inf, -Inf) when empty */
int32_t rc_world_bound(const rc_context *ctx, float out[6]);
/* wait_for_gpu!(accel) — :2418-2421 */
int32_t rc_wait(rc_context *ctx);
/* sizes of the synced structure in the reference's terms: TLAS nodes (max(1,2n-1), 0 if empty),
 * Σ live BLAS nodes (2n_b-1) and prims — the flat-array invariants of test/test_mesh_update.jl:261-294 */
int32_t rc_sizes(const rc_context *ctx, uint32_t *tlas_nodes, uint32_t *blas_nodes, uint32_t *blas_prims, uint32_t *pending_deletes);
/* read back the reference-layout BVH2 (for parity tests of the builder).  blas_index 1-based (post-sync numbering). */
int32_t rc_read_tlas_nodes(rc_context *ctx, rc_bvh_node2 *out, uint32_t capacity);
int32_t rc_read_blas_nodes(rc_context *ctx, uint32_t blas_index, rc_bvh_node2 *out, uint32_t capacity);
/* sorted primitive order of a BLAS: out[k] = primitive_id of the k-th Morton-sorted triangle */
int32_t rc_read_blas_order(rc_context *ctx, uint32_t blas_index, uint32_t *out, uint32_t capacity);
uint32_t rc_blas_n_prims(const rc_context *ctx, uint32_t blas_index);
/* out[primitive_id] = index of that triangle in the submitted soup (before the degenerate filter), so a shim can
 * materialise the reference's Triangle (vertices, normals, uv, metadata) from its own copy of the mesh */
int32_t rc_read_blas_faces(rc_context *ctx, uint32_t blas_index, uint32_t *out, uint32_t capacity);
/* out[i] = handle id owning instance position i (0-based), n_total_instances entries */
int32_t rc_get_instance_handles(const rc_context *ctx, uint32_t *out, uint32_t capacity);

/* ---- queries ------------------------------------------------------------------------- */
/* Batched closest_hit(::StaticTLAS, ray) — src/instanced-bvh.jl:1902-2024; batch shape of
 * Lava.trace_closest_hits!(hits, rays, accel, n) (docs/src/hw_acceleration.md:143-146) and
 * trace_rays (ext/RaycoreMakieExt.jl:81-87).  Host pointers are staged through pinned chunks with
 * copies overlapped with tracing; device pointers are used in place.  The context must be synced. */
int32_t rc_trace_closest(rc_context *ctx, const rc_ray *rays, rc_hit *hits, uint64_t n, uint32_t flags);
/* Batched any_hit — src/instanced-bvh.jl:2034-2140 (t_min forced to 0, :2039). */
int32_t rc_trace_any(rc_context *ctx, const rc_ray *rays, rc_hit *hits, uint64_t n, uint32_t flags);
/* work counters accumulated by RC_COUNTERS traces since the last reset:
 * out = {rays, node fetches, child-box tests, triangle tests, instance entries, max stack depth} */
int32_t rc_get_counters(rc_context *ctx, uint64_t out[6], int32_t reset);
/* milliseconds of the last trace's kernel(s), measured with CUDA events on the context stream */
float rc_last_kernel_ms(const rc_context *ctx);
uint32_t rc_last_kernel_launches(const rc_context *ctx);
/* milliseconds of the last BLAS build (rc_push / rc_update_geometry), CUDA events on the context stream, upload excluded */
float rc_last_build_ms(const rc_context *ctx);

/* ---- analysis (src/kernels.jl) --------------------------------------------------------- */
/* hits_from_grid(tlas, viewdir; grid_size) — :58-72 with generate_ray_grid :10-56.
 * hits: grid*grid records, cell (i,j) 1-based at (j-1)*grid + (i-1) (Julia column-major);
 * points (nullable): grid*grid*3 floats = sum_mul(bary, vertices).  Host pointers. */
int32_t rc_hits_from_grid(rc_context *ctx, const float viewdir[3], uint32_t grid, rc_hit *hits, float *points);
/* get_illumination(tlas, viewdir; grid_size=1000) — :112-124; out: n_prims floats (hit counts per metadata 1..n_prims) */
int32_t rc_get_illumination(rc_context *ctx, const float viewdir[3], uint32_t grid, float *out, uint32_t n_out);
/* get_centroid(tlas, viewdir; grid_size=32) — :106-110; points (nullable, capacity grid*grid*3): hit points, compacted */
int32_t rc_get_centroid(rc_context *ctx, const float viewdir[3], uint32_t grid, float centroid[3], uint32_t *n_hits, float *points);
/* view_factors(tlas; rays_per_triangle) — :74-104.  Sources are the primitives of the flat array
 * (_primitives(tlas), :8: every BLAS in order, local-space vertices, as the reference uses them) whose
 * metadata - 1 lies in [row_base, row_base + n_rows); each shoots rays_per_triangle rays and
 *   out[(meta_src - 1 - row_base) * n_prims + (meta_hit - 1)] += 1     (self hits skipped, :95-97)
 * i.e. out is the row block [row_base, row_base+n_rows) of the row-major [src][hit] matrix = the transpose of
 * Julia's column-major result[src, hit].  A multi-GPU caller gives each rank its own row block.
 * Requires metadata dense in 1..n_prims as the reference does (unchecked there, :85); primitives with
 * out-of-range metadata are skipped and counted in *skipped (reported for row_base == 0 only).
 * Randomness: counter-based RNG keyed by (seed, (meta_src-1)*rays_per_triangle + i, dim) — see DESIGN.md;
 * the reference's task-local rand() stream is not reproducible, parity is statistical.
 * out is zeroed first.  RC_HITS_ON_DEVICE => out is a device pointer. */
int32_t rc_view_factors(rc_context *ctx, uint32_t rays_per_triangle, uint64_t seed, uint32_t *out, uint32_t row_base, uint32_t n_rows,
                        uint32_t flags, uint64_t *skipped);
/* the same for the interleaved row set {row_first + k * row_stride, k < n_rows} (out row k): with row_first = rank and
 * row_stride = world size every GPU gets an equally expensive share of the rows. */
int32_t rc_view_factors_strided(rc_context *ctx, uint32_t rays_per_triangle, uint64_t seed, uint32_t *out, uint32_t row_first, uint32_t row_stride, uint32_t n_rows,
                                uint32_t flags, uint64_t *skipped);
/* the rays rc_view_factors generates for that row block: host buffer of n_rows*rays_per_triangle records,
 * ray (meta_src-1-row_base)*rays_per_triangle + i (rows without a source stay zero) */
int32_t rc_view_factor_rays(rc_context *ctx, uint32_t rays_per_triangle, uint64_t seed, uint32_t row_base, uint32_t n_rows, rc_ray *out);
/* metadata of the flat primitive array (position -> metadata), n_prims entries */
int32_t rc_read_flat_metadata(rc_context *ctx, uint32_t *out, uint32_t capacity);

/* ---- collision broad phase (src/collision.jl) ---------------------------- */
/* ContactPair, 8 bytes — src/collision.jl:25-28 (1-based instance indices, instance_a < instance_b) */
typedef struct rc_contact_pair {
    uint32_t instance_a, instance_b;
} rc_contact_pair;
/* collide_instances(tlas) — src/collision.jl:189-233: every pair of instances whose world AABBs overlap, in the reference's
 * output order.  Two-call protocol: contacts == NULL (or capacity too small) only reports *n_contacts.  Host pointers. */
int32_t rc_collide_instances(rc_context *ctx, rc_contact_pair *contacts, uint64_t capacity, uint64_t *n_contacts);
/* collide_instances_any(tlas, handle_a, handle_b) — src/collision.jl:241-261: do any two instances of the two handles overlap
 * (world AABBs)?  Each instance's own TLAS leaf box is used (the reference indexes the Morton-sorted leaves with the instance
 * index, which is only correct when the sort is the identity — see DESIGN.md). */
int32_t rc_collide_instances_any(rc_context *ctx, uint32_t handle_a, uint32_t handle_b, int32_t *overlap);

/* ---- wavefront stages either side of the trace (docs/src/wavefront-renderer.jl) ----
 * The reference's renderer keeps SoA work queues on the device and runs one kernel per stage; these entry points are those stages
 * on DEVICE-RESIDENT queues (every rays / hits / visible pointer below is device memory, e.g. from rc_device_alloc), so closest_hit
 * and any_hit are fed without a host round trip.  The queue fields the reference carries beside the ray (pixel_x, pixel_y,
 * sample_idx, hit_idx, light_idx) are pure functions of the queue index and are not stored.  RC_NO_SYNC applies. */
#define RC_WAVE_NO_JITTER 0x40u /* primary rays through pixel centres instead of a jittered sample */
#define RC_MAX_LIGHTS 16
/* Triangle.normals (src/triangle_mesh.jl:3,20) of a handle's geometry: 9 floats (n0, n1, n2) per SUBMITTED face, in rc_push
 * order (degenerate faces are dropped by the library).  Optional: a geometry without normals shades with its geometric normal
 * (the reference always has vertex normals, src/instanced-bvh.jl:2291-2298).  rc_update_geometry drops them.  RC_VERTS_ON_DEVICE applies. */
int32_t rc_set_normals(rc_context *ctx, uint32_t handle, const float *normals, uint32_t n_faces, uint32_t flags);
/* generate_primary_rays! — docs/src/wavefront-renderer.jl:185-213: pinhole camera at camera_pos looking down +z;
 * ray ((y-1)*width + (x-1))*n_samples + (s-1) for 1-based pixel (x, y), sample s.  The reference's rand(Vec2f) jitter is the
 * counter RNG of DESIGN.md (seed, ray index, dims 0/1).  rays: width*height*n_samples records. */
int32_t rc_generate_primary_rays(rc_context *ctx, uint32_t width, uint32_t height, uint32_t n_samples, const float camera_pos[3], float focal_length, float aspect,
                                 uint64_t seed, rc_ray *rays, uint32_t flags);
/* generate_primary_rays_lookat! — :219-253 */
int32_t rc_generate_primary_rays_lookat(rc_context *ctx, uint32_t width, uint32_t height, uint32_t n_samples, const float camera_pos[3], const float right[3],
                                        const float up[3], const float forward[3], float half_width, float half_height, uint64_t seed, rc_ray *rays, uint32_t flags);
/* generate_shadow_rays! — :277-330: shadow ray k*n_lights + l from primary hit k towards light l (origin = hit point +
 * shadow_bias * interpolated normal, t_max = distance to the light; the reference hard-codes shadow_bias = 0.01); a missed
 * primary ray yields the dummy ray (t_max = 0).  lights: n_lights*3 floats, HOST memory, n_lights <= RC_MAX_LIGHTS.
 * hits must come from rc_trace_closest on the same, still synced, TLAS.  shadow_rays: n*n_lights records. */
int32_t rc_generate_shadow_rays(rc_context *ctx, const rc_ray *rays, const rc_hit *hits, uint64_t n, const float *lights, uint32_t n_lights, float shadow_bias,
                                rc_ray *shadow_rays, uint32_t flags);
/* test_shadow_rays! — :337-362: visible[i] = shadow_rays[i].t_max > 0 ? !any_hit(shadow_rays[i]) : 0 (one byte per ray) */
int32_t rc_test_shadow_rays(rc_context *ctx, const rc_ray *shadow_rays, uint64_t n, uint8_t *visible, uint32_t flags);
/* stages 3 + 4 in one kernel: same result as rc_generate_shadow_rays followed by rc_test_shadow_rays, but each shadow ray is
 * generated inside the traversal kernel when a lane picks it up, so the shadow-ray queue is never written to memory. */
int32_t rc_shadow_visibility(rc_context *ctx, const rc_ray *rays, const rc_hit *hits, uint64_t n, const float *lights, uint32_t n_lights, float shadow_bias,
                             uint8_t *visible, uint32_t flags);

/* ---- device memory helpers for callers that keep rays/hits resident -------------------- */
int32_t rc_device_alloc(rc_context *ctx, size_t bytes, void **out);
int32_t rc_device_free(rc_context *ctx, void *ptr);
int32_t rc_host_alloc(rc_context *ctx, size_t bytes, void **out); /* pinned */
int32_t rc_host_free(rc_context *ctx, void *ptr);
int32_t rc_memcpy_h2d(rc_context *ctx, void *dst, const void *src, size_t bytes);
int32_t rc_memcpy_d2h(rc_context *ctx, void *dst, const void *src, size_t bytes);
/* cross-process peer access for the fused multi-GPU gather: export a device allocation as a 64-byte IPC
 * handle / open one exported by another rank (one process per GPU). */
int32_t rc_ipc_export(rc_context *ctx, void *ptr, uint8_t handle_out[64]);
int32_t rc_ipc_open(rc_context *ctx, const uint8_t handle[64], void **out);
int32_t rc_ipc_close(rc_context *ctx, void *ptr);
/* copy-engine gather: enqueue a device-to-device copy (dst may be a peer pointer from rc_ipc_open) on the context's copy stream,
 * ordered after the work already enqueued on the context stream; it overlaps with later traces.  slot (0/1) names the completion
 * event; rc_stream_wait_copy(slot) makes the context stream wait for it (call before tracing into that source buffer again).
 * rc_wait() drains the copy stream. */
int32_t rc_peer_copy_async(rc_context *ctx, void *dst, const void *src, size_t bytes, uint32_t slot);
int32_t rc_stream_wait_copy(rc_context *ctx, uint32_t slot);

/* ---- BLAS4 / build_blas4 / closest_hit4 / any_hit4 (src/bvh4.jl:154-163, 511-523, 606-766; exported at src/Raycore.jl:105) -------------
 * One geometry traversed on its own 4-wide BVH, without a TLAS.  The library's wide BVH is exactly that structure, so a BLAS4 is a
 * geometry under one identity instance (the single-instance kernel variant has no top level).  The reference's 120-byte BVHNode4
 * (full float boxes, src/bvh4.jl:40-72) is not reproduced: rc_blas4_read_nodes returns this library's 64-byte quantised nodes. */
typedef struct rc_blas4 rc_blas4;
/* wide node, 64 bytes: child k's box = origin + q * scale per axis, lo planes rounded down / hi planes rounded up (conservative), byte k of
 * every q word = child k; s* = scale * 2^24; child reference: wide-node index, or 0x80000000 | (count-1) << 28 | first sorted triangle;
 * an unused slot repeats child 0 under an inverted box (qlo = 255, qhi = 0).  Slot 0 and slots that head no wide node are zero. */
typedef struct rc_wide_node {
    float origin[3], sx;
    uint32_t qlo[3], qhi_x;
    uint32_t qhi_y, qhi_z, child01[2];
    uint32_t child23[2];
    float sy, sz;
} rc_wide_node;
/* build_blas4(primitives) — :511-523.  Same vertex / metadata / flag conventions as rc_push; "Cannot build BLAS4 from empty primitive
 * list" (:513) when no valid triangle is left.  Errors of a failed build: rc_blas4_last_error(NULL). */
int32_t rc_blas4_build(int32_t device, const float *verts, uint32_t n_faces, const uint32_t *face_meta, uint32_t flags, rc_blas4 **out);
int32_t rc_blas4_destroy(rc_blas4 *b);
const char *rc_blas4_last_error(const rc_blas4 *b);
/* n_primitives = length(blas.primitives); n_node_slots = wide-node slots incl. slot 0 (root = slot 1); root_aabb = blas.root_aabb */
int32_t rc_blas4_info(const rc_blas4 *b, uint32_t *n_primitives, uint32_t *n_node_slots, float root_aabb[6]);
/* closest_hit4(blas, ray) / any_hit4(blas, ray) — :606-766, batched like rc_trace_*; ray.tmin is ignored as the reference does (:610);
 * hit.instance_id = instance_custom_index = 0.  Flags: RC_RAYS_ON_DEVICE, RC_HITS_ON_DEVICE, RC_NO_SYNC, RC_MODE_WATERTIGHT. */
int32_t rc_blas4_trace_closest(rc_blas4 *b, const rc_ray *rays, rc_hit *hits, uint64_t n, uint32_t flags);
int32_t rc_blas4_trace_any(rc_blas4 *b, const rc_ray *rays, rc_hit *hits, uint64_t n, uint32_t flags);
int32_t rc_blas4_read_nodes(rc_blas4 *b, rc_wide_node *out, uint32_t capacity);
/* the context behind the BLAS4 (read-backs such as rc_read_blas_order / rc_read_blas_faces, streams, timers) */
rc_context *rc_blas4_context(rc_blas4 *b);

/* ---- several GPUs of one node behind one handle ---------------------------------------------------------
 * One process, all devices, inside the library: an rc_multi owns one context per device; the scene is replicated (every mutation is
 * applied to every device — the builder is deterministic, the replicas are byte-identical) and queries are sharded with no data-path
 * collective.  The reference's only parallelism on this path is Threads.@threads over rays / source triangles (src/kernels.jl:64,82);
 * this is the same split across GPUs.  devices == NULL / n_devices == 0: every visible device.  Errors: rc_multi_last_error. */
typedef struct rc_multi rc_multi;
int32_t rc_multi_create(const int32_t *devices, uint32_t n_devices, rc_multi **out);
int32_t rc_multi_destroy(rc_multi *m);
const char *rc_multi_last_error(const rc_multi *m); /* m may be NULL: last creation error */
uint32_t rc_multi_device_count(const rc_multi *m);
/* the k-th device's context, for everything that is not sharded (introspection, read-backs, the wavefront stages): treat it as read-only —
 * mutating one replica alone makes the replicas diverge */
rc_context *rc_multi_context(rc_multi *m, uint32_t k);
/* push! / delete! / update_transform(s)! / update! / sync! on every replica (same arguments and results as the single-device calls; host
 * pointers only — each device uploads its own copy, concurrently) */
int32_t rc_multi_push(rc_multi *m, const float *verts, uint32_t n_faces, const uint32_t *face_meta, const float *transforms, const float *inv_transforms,
                      const uint32_t *instance_ids, uint32_t n_instances, uint32_t flags, uint32_t *handle_out);
int32_t rc_multi_delete(rc_multi *m, uint32_t handle, int32_t *deleted);
int32_t rc_multi_update_transforms(rc_multi *m, uint32_t handle, const float *transforms, const float *inv_transforms, uint32_t n);
int32_t rc_multi_update_geometry(rc_multi *m, uint32_t handle, const float *verts, uint32_t n_faces, const uint32_t *face_meta, uint32_t flags);
int32_t rc_multi_sync(rc_multi *m, int32_t *action);
/* closest_hit / any_hit over n rays, rays [k n / G, (k+1) n / G) on device k; results are byte-identical to the single-device call.
 * Host buffers (flags without RC_*_ON_DEVICE): every device stages its own slice straight from / to the caller's arrays over its own
 * host link — H2D, trace and D2H of all devices overlap, nothing hops through a root GPU (pin the arrays, e.g. rc_host_alloc, for full
 * PCIe rate).  Device buffers (both RC_RAYS_ON_DEVICE and RC_HITS_ON_DEVICE, memory of the FIRST device): every device's traversal kernel
 * reads its ray slice and stores its hit records through the NVLink peer mapping directly — the gather is the kernel's epilogue. */
int32_t rc_multi_trace_closest(rc_multi *m, const rc_ray *rays, rc_hit *hits, uint64_t n, uint32_t flags);
int32_t rc_multi_trace_any(rc_multi *m, const rc_ray *rays, rc_hit *hits, uint64_t n, uint32_t flags);
/* view_factors(tlas; rays_per_triangle) into the caller's host matrix (n_prims x n_prims, row-major [src][hit]): source rows
 * [k N / G, (k+1) N / G) are computed on device k and copied to their place; no exchange between the devices. */
int32_t rc_multi_view_factors(rc_multi *m, uint32_t rays_per_triangle, uint64_t seed, uint32_t *out, uint64_t *skipped);

#ifdef __cplusplus
}
#endif
#endif /* RAYCORE_CUDA_H */
