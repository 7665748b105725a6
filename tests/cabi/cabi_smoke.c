/* Pure-C driver of the drop-in boundary (include/raycore_cuda.h): no Python, no torch — what a cgo / ccall / JNI binding sees.
 * Mirrors test/test_instanced_bvh.jl:283-301 (a quad with metadata 42 hit at t = 1, a miss beside it) plus a two-instance push,
 * a transform update (refit), an any-hit query, the watertight mode, a vertex-update refit and the BLAS4 entry points.
 * Exit code 0 = all checks passed. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "raycore_cuda.h"

#define CHECK(cond)                                                      \
    do {                                                                 \
        if (!(cond)) {                                                   \
            fprintf(stderr, "FAILED %s:%d: %s (%s)\n", __FILE__, __LINE__, #cond, rc_last_error(ctx)); \
            return 1;                                                    \
        }                                                                \
    } while (0)

int main(void) {
    rc_context *ctx = NULL;
    if (rc_create(-1, &ctx) != RC_OK) {
        fprintf(stderr, "rc_create: %s\n", rc_last_error(NULL));
        return 2;
    }
    /* unit quad in the plane z = 0, two triangles */
    const float quad[18] = {-1, -1, 0, 1, -1, 0, 1, 1, 0, -1, -1, 0, 1, 1, 0, -1, 1, 0};
    const uint32_t meta[2] = {42, 42};
    const float xf[24] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, /* identity */
                          1, 0, 0, 5, 0, 1, 0, 0, 0, 0, 1, 0 /* translated to x = 5 */};
    const uint32_t ids[2] = {7, 9};
    uint32_t handle = 0;
    CHECK(rc_push(ctx, quad, 2, meta, xf, NULL, ids, 2, 0, &handle) == RC_OK);
    int32_t action = -1;
    CHECK(rc_sync(ctx, &action) == RC_OK && action == RC_SYNC_REBUILD);
    CHECK(rc_n_instances(ctx) == 2 && rc_n_geometries(ctx) == 1);

    rc_ray rays[3] = {{{0.25f, 0.25f, 1.0f}, 0.0f, {0, 0, -1}, INFINITY},  /* hits instance 0 at t = 1 */
                      {{5.25f, 0.25f, 2.0f}, 0.0f, {0, 0, -1}, INFINITY},  /* hits instance 1 at t = 2 */
                      {{2.5f, 0.0f, 1.0f}, 0.0f, {0, 0, -1}, INFINITY}};   /* between them: miss */
    rc_hit hits[3];
    memset(hits, 0xFF, sizeof hits);
    CHECK(rc_trace_closest(ctx, rays, hits, 3, 0) == RC_OK);
    CHECK(hits[0].hit == 1 && fabsf(hits[0].t - 1.0f) < 1e-6f && hits[0].metadata == 42 && hits[0].instance_id == 0 && hits[0].instance_custom_index == 7);
    CHECK(hits[1].hit == 1 && fabsf(hits[1].t - 2.0f) < 1e-6f && hits[1].instance_id == 1 && hits[1].instance_custom_index == 9);
    CHECK(hits[2].hit == 0 && hits[2].t == 0.0f);
    CHECK(rc_trace_any(ctx, rays, hits, 3, 0) == RC_OK && hits[0].hit == 1 && hits[1].hit == 1 && hits[2].hit == 0);

    /* move the second instance under the third ray: transforms only => refit, not rebuild */
    float xf2[24];
    memcpy(xf2, xf, sizeof xf);
    xf2[12 + 3] = 2.5f;
    CHECK(rc_update_transforms(ctx, handle, xf2, NULL, 2) == RC_OK);
    CHECK(rc_trace_closest(ctx, rays, hits, 3, 0) == RC_ERR_NOT_SYNCED); /* pending mutation: the library refuses to trace */
    CHECK(rc_sync(ctx, &action) == RC_OK && action == RC_SYNC_REFIT);
    CHECK(rc_trace_closest(ctx, rays, hits, 3, 0) == RC_OK);
    CHECK(hits[2].hit == 1 && hits[2].instance_id == 1 && fabsf(hits[2].t - 1.0f) < 1e-6f && hits[1].hit == 0);

    /* serialised geometry: export the built quad, restore it in a second context, same answers (two-call size protocol) */
    {
        uint64_t size = 0;
        CHECK(rc_export_geometry(ctx, handle, NULL, 0, &size) == RC_OK && size > 128);
        void *blob = malloc(size);
        CHECK(blob != NULL && rc_export_geometry(ctx, handle, blob, size, &size) == RC_OK);
        rc_context *ctx2 = NULL;
        CHECK(rc_create(-1, &ctx2) == RC_OK);
        uint32_t h2 = 0;
        rc_hit hits2[3];
        int same = rc_push_exported(ctx2, blob, size, xf2, NULL, ids, 2, &h2) == RC_OK && rc_sync(ctx2, &action) == RC_OK &&
                   rc_trace_closest(ctx2, rays, hits2, 3, 0) == RC_OK && memcmp(hits, hits2, sizeof hits) == 0;
        ((unsigned char *)blob)[200] ^= 1; /* a damaged blob is refused */
        int refused = rc_push_exported(ctx2, blob, size, xf2, NULL, ids, 2, &h2) == RC_ERR_INVALID_ARGUMENT;
        free(blob);
        CHECK(rc_destroy(ctx2) == RC_OK);
        CHECK(same && refused);
    }

    /* watertight mode (the reference's intersect_triangle, src/triangle_mesh.jl:168-201): a ray through the quad's shared diagonal hits */
    {
        rc_ray diag = {{0.5f, 0.5f, 1.0f}, 0.0f, {0, 0, -1}, INFINITY};
        rc_hit hd;
        CHECK(rc_trace_closest(ctx, &diag, &hd, 1, RC_MODE_WATERTIGHT) == RC_OK && hd.hit == 1 && fabsf(hd.t - 1.0f) < 1e-6f);
        CHECK(rc_trace_any(ctx, &diag, &hd, 1, RC_MODE_WATERTIGHT) == RC_OK && hd.hit == 1);
    }

    /* vertex update as a refit (update!, src/instanced-bvh.jl:808-857): the geometry was built with RC_BUILD_ALLOW_REFIT, the quad moves to
     * z = -1 with the same faces => the kept radix tree is re-fitted, no rebuild of the BLAS */
    {
        rc_context *c3 = NULL;
        CHECK(rc_create(-1, &c3) == RC_OK && rc_set_build_flags(c3, RC_BUILD_ALLOW_REFIT) == RC_OK);
        uint32_t h3 = 0;
        CHECK(rc_push(c3, quad, 2, meta, xf, NULL, NULL, 1, 0, &h3) == RC_OK && rc_sync(c3, &action) == RC_OK);
        float moved[18];
        memcpy(moved, quad, sizeof quad);
        for (int k = 2; k < 18; k += 3) moved[k] = -1.0f;
        int ok = rc_update_geometry(c3, h3, moved, 2, meta, RC_UPDATE_REFIT) == RC_OK && rc_last_update_refitted(c3) == 1 &&
                 rc_sync(c3, &action) == RC_OK && rc_trace_closest(c3, rays, hits, 1, 0) == RC_OK && hits[0].hit == 1 &&
                 fabsf(hits[0].t - 2.0f) < 1e-6f;
        CHECK(rc_destroy(c3) == RC_OK);
        CHECK(ok);
        CHECK(rc_trace_closest(ctx, rays, hits, 3, 0) == RC_OK); /* (restore hits[] for the checks below) */
    }

    /* BLAS4 / build_blas4 / closest_hit4 / any_hit4 (src/bvh4.jl:511-766): one geometry on its own wide BVH, ray.tmin ignored (:610) */
    {
        rc_blas4 *b4 = NULL;
        CHECK(rc_blas4_build(-1, quad, 2, meta, 0, &b4) == RC_OK);
        uint32_t n_prims = 0, n_slots = 0;
        float box[6];
        CHECK(rc_blas4_info(b4, &n_prims, &n_slots, box) == RC_OK && n_prims == 2 && n_slots == 3 && box[0] == -1.0f && box[3] == 1.0f);
        rc_ray r4 = {{0.25f, 0.25f, 1.0f}, 5.0f /* tmin beyond the hit: ignored */, {0, 0, -1}, INFINITY};
        rc_hit h4;
        CHECK(rc_blas4_trace_closest(b4, &r4, &h4, 1, 0) == RC_OK && h4.hit == 1 && fabsf(h4.t - 1.0f) < 1e-6f && h4.metadata == 42);
        CHECK(rc_blas4_trace_any(b4, &r4, &h4, 1, 0) == RC_OK && h4.hit == 1);
        rc_wide_node nodes[3];
        CHECK(rc_blas4_read_nodes(b4, nodes, 3) == RC_OK && (nodes[1].child01[0] & 0x80000000u) != 0u); /* the root's first child is a leaf */
        rc_blas4 *none = NULL;
        const float flat[9] = {0, 0, 0, 1, 1, 1, 2, 2, 2}; /* degenerate: "Cannot build BLAS4 from empty primitive list" (:513) */
        CHECK(rc_blas4_build(-1, flat, 1, NULL, 0, &none) == RC_ERR_NO_VALID_TRIANGLES && none == NULL);
        CHECK(rc_blas4_destroy(b4) == RC_OK);
    }

    /* error behaviour of the handle API (src/instanced-bvh.jl:715-718) */
    CHECK(rc_update_transforms(ctx, handle + 100, xf2, NULL, 2) == RC_ERR_INVALID_HANDLE);
    CHECK(rc_update_transforms(ctx, handle, xf2, NULL, 1) == RC_ERR_INVALID_ARGUMENT);
    int32_t deleted = 0;
    CHECK(rc_delete(ctx, handle, &deleted) == RC_OK && deleted == 1);
    CHECK(rc_sync(ctx, &action) == RC_OK && rc_n_instances(ctx) == 0);
    CHECK(rc_trace_closest(ctx, rays, hits, 3, 0) == RC_OK && hits[0].hit == 0); /* empty TLAS => miss */
    CHECK(rc_destroy(ctx) == RC_OK);
    printf("cabi smoke ok\n");
    return 0;
}
