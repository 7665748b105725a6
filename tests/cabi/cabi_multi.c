/* Pure-C driver of the multi-GPU entry points (include/raycore_cuda.h, rc_multi_*): one process, every visible GPU, no Python, no torch,
 * no torch.distributed — what a Julia caller holding a CuTLAS gets from a ccall.  A scene of instanced tetrahedra is replicated on all
 * devices, 2^20 rays are traced (a) on one device through rc_trace_closest and (b) sharded over all devices through
 * rc_multi_trace_closest from the same host arrays; the hit records must be byte-identical.  Also: refit through the multi handle,
 * any_hit, the sharded view-factor matrix against the single-device one, and the device-resident sharded trace (peer stores).
 * Exit code 0 = all checks passed; prints the device count and the two wall times. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "raycore_cuda.h"

#define CHECK(cond)                                                                                                          \
    do {                                                                                                                     \
        if (!(cond)) {                                                                                                       \
            fprintf(stderr, "FAILED %s:%d: %s (%s | %s)\n", __FILE__, __LINE__, #cond, rc_multi_last_error(m), rc_last_error(single)); \
            return 1;                                                                                                        \
        }                                                                                                                    \
    } while (0)

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static float frand(void) { /* xorshift64*, 24-bit uniform */
    rng_state ^= rng_state >> 12; rng_state ^= rng_state << 25; rng_state ^= rng_state >> 27;
    return (float)((rng_state * 0x2545F4914F6CDD1Dull) >> 40) * (1.0f / 16777216.0f);
}
static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

int main(void) {
    rc_multi *m = NULL;
    rc_context *single = NULL;
    if (rc_multi_create(NULL, 0, &m) != RC_OK) { fprintf(stderr, "rc_multi_create: %s\n", rc_multi_last_error(NULL)); return 2; }
    if (rc_create(0, &single) != RC_OK) { fprintf(stderr, "rc_create: %s\n", rc_last_error(NULL)); return 2; }
    const uint32_t g = rc_multi_device_count(m);
    /* a tetrahedron (4 faces), metadata 1..4, 512 instances scattered in [-20, 20]^3 with scales in [0.5, 1.5] */
    const float a[3] = {1, 1, 1}, b[3] = {-1, -1, 1}, c[3] = {-1, 1, -1}, d[3] = {1, -1, -1};
    float tet[36];
    const float *f[4][3] = {{a, b, c}, {a, c, d}, {a, d, b}, {b, d, c}};
    for (int i = 0; i < 4; i++) for (int j = 0; j < 3; j++) memcpy(tet + 9 * i + 3 * j, f[i][j], 12);
    enum { NI = 512 };
    float *xf = malloc(sizeof(float) * 12 * NI);
    uint32_t ids[NI];
    for (int i = 0; i < NI; i++) {
        const float s = 0.5f + frand();
        const float t[12] = {s, 0, 0, 40 * frand() - 20, 0, s, 0, 40 * frand() - 20, 0, 0, s, 40 * frand() - 20};
        memcpy(xf + 12 * i, t, sizeof t);
        ids[i] = 1000u + (uint32_t)i;
    }
    uint32_t hm = 0, hs = 0;
    int32_t action = -1;
    CHECK(rc_multi_push(m, tet, 4, NULL, xf, NULL, ids, NI, 0, &hm) == RC_OK);
    CHECK(rc_multi_sync(m, &action) == RC_OK && action == RC_SYNC_REBUILD);
    CHECK(rc_push(single, tet, 4, NULL, xf, NULL, ids, NI, 0, &hs) == RC_OK && rc_sync(single, NULL) == RC_OK && hs == hm);
    for (uint32_t k = 0; k < g; k++) CHECK(rc_n_instances(rc_multi_context(m, k)) == NI);

    /* rays: origins in [-24, 24]^3, uniform directions; pinned host arrays so every device runs at full PCIe rate */
    const uint64_t n = 1u << 20;
    rc_ray *rays = NULL;
    rc_hit *h1 = NULL, *h2 = NULL;
    CHECK(rc_host_alloc(single, n * sizeof(rc_ray), (void **)&rays) == RC_OK);
    CHECK(rc_host_alloc(single, n * sizeof(rc_hit), (void **)&h1) == RC_OK && rc_host_alloc(single, n * sizeof(rc_hit), (void **)&h2) == RC_OK);
    for (uint64_t i = 0; i < n; i++) {
        const float z = 1 - 2 * frand(), phi = 6.2831853f * frand(), r = sqrtf(fmaxf(0.f, 1 - z * z));
        const rc_ray q = {{48 * frand() - 24, 48 * frand() - 24, 48 * frand() - 24}, 0.0f, {r * cosf(phi), r * sinf(phi), z}, INFINITY};
        rays[i] = q;
    }
    memset(h1, 0xEE, n * sizeof(rc_hit));
    memset(h2, 0xDD, n * sizeof(rc_hit));
    CHECK(rc_trace_closest(single, rays, h1, n, 0) == RC_OK && rc_multi_trace_closest(m, rays, h2, n, 0) == RC_OK); /* warm-up */
    double t0 = now();
    CHECK(rc_trace_closest(single, rays, h1, n, 0) == RC_OK);
    double t1 = now();
    CHECK(rc_multi_trace_closest(m, rays, h2, n, 0) == RC_OK);
    double t2 = now();
    CHECK(memcmp(h1, h2, n * sizeof(rc_hit)) == 0);
    uint64_t nh = 0;
    for (uint64_t i = 0; i < n; i++) nh += h1[i].hit;
    CHECK(nh > n / 50 && nh < n);
    CHECK(rc_trace_any(single, rays, h1, n, 0) == RC_OK && rc_multi_trace_any(m, rays, h2, n, 0) == RC_OK);
    for (uint64_t i = 0; i < n; i++) CHECK(h1[i].hit == h2[i].hit);
    /* an odd count that does not divide by the device count, and a count smaller than it */
    CHECK(rc_trace_closest(single, rays, h1, 1000003, 0) == RC_OK && rc_multi_trace_closest(m, rays, h2, 1000003, 0) == RC_OK && memcmp(h1, h2, 1000003 * sizeof(rc_hit)) == 0);
    CHECK(rc_trace_closest(single, rays, h1, 3, 0) == RC_OK && rc_multi_trace_closest(m, rays, h2, 3, 0) == RC_OK && memcmp(h1, h2, 3 * sizeof(rc_hit)) == 0);

    /* move every instance: transforms only => refit on every replica */
    for (int i = 0; i < NI; i++) xf[12 * i + 3] += 0.75f;
    CHECK(rc_multi_update_transforms(m, hm, xf, NULL, NI) == RC_OK && rc_multi_sync(m, &action) == RC_OK && action == RC_SYNC_REFIT);
    CHECK(rc_update_transforms(single, hs, xf, NULL, NI) == RC_OK && rc_sync(single, NULL) == RC_OK);
    CHECK(rc_trace_closest(single, rays, h1, n, 0) == RC_OK && rc_multi_trace_closest(m, rays, h2, n, 0) == RC_OK && memcmp(h1, h2, n * sizeof(rc_hit)) == 0);

    /* device-resident buffers on the first device: the other devices read rays and store hits through the peer mapping */
    {
        rc_context *c0 = rc_multi_context(m, 0);
        void *d_rays = NULL, *d_hits = NULL;
        CHECK(rc_device_alloc(c0, n * sizeof(rc_ray), &d_rays) == RC_OK && rc_device_alloc(c0, n * sizeof(rc_hit), &d_hits) == RC_OK);
        CHECK(rc_memcpy_h2d(c0, d_rays, rays, n * sizeof(rc_ray)) == RC_OK);
        int32_t rc = rc_multi_trace_closest(m, (const rc_ray *)d_rays, (rc_hit *)d_hits, n, RC_RAYS_ON_DEVICE | RC_HITS_ON_DEVICE);
        if (rc == RC_OK) {
            CHECK(rc_memcpy_d2h(c0, h2, d_hits, n * sizeof(rc_hit)) == RC_OK && memcmp(h1, h2, n * sizeof(rc_hit)) == 0);
        } else {
            CHECK(g > 1 && rc == RC_ERR_INVALID_ARGUMENT); /* no peer access on this box: refused with a message, host buffers still work */
            printf("device-resident sharded trace unavailable: %s\n", rc_multi_last_error(m));
        }
        CHECK(rc_device_free(c0, d_rays) == RC_OK && rc_device_free(c0, d_hits) == RC_OK);
    }

    /* view factors: 4 x 512 ... the flat primitive array holds the BLAS once (4 primitives, metadata 1..4): a 4 x 4 matrix, rows sharded */
    {
        uint32_t vf1[16], vf2[16];
        uint64_t sk1 = 7, sk2 = 7;
        CHECK(rc_view_factors(single, 5000, 3, vf1, 0, 4, 0, &sk1) == RC_OK);
        CHECK(rc_multi_view_factors(m, 5000, 3, vf2, &sk2) == RC_OK);
        CHECK(memcmp(vf1, vf2, sizeof vf1) == 0 && sk1 == sk2);
    }

    int32_t deleted = 0;
    CHECK(rc_multi_delete(m, hm, &deleted) == RC_OK && deleted == 1 && rc_multi_sync(m, &action) == RC_OK);
    CHECK(rc_multi_trace_closest(m, rays, h2, 1000, 0) == RC_OK);
    for (int i = 0; i < 1000; i++) CHECK(h2[i].hit == 0); /* empty TLAS => miss (test/test_tlas_stress.jl:808-831) */
    CHECK(rc_host_free(single, rays) == RC_OK && rc_host_free(single, h1) == RC_OK && rc_host_free(single, h2) == RC_OK);
    free(xf);
    CHECK(rc_multi_destroy(m) == RC_OK && rc_destroy(single) == RC_OK);
    printf("cabi multi ok: %u device(s), 2^20 rays from host buffers: %.2f ms on one device, %.2f ms sharded\n", g, 1e3 * (t1 - t0), 1e3 * (t2 - t1));
    return 0;
}
