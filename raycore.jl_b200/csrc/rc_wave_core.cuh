// rc_wave_core.cuh — per-element bodies of the wavefront stages either side of the trace:
// primary-ray generation and shadow-ray generation.  RC_HD so tests/hostsim can run them on the CPU.
// Reference: docs/src/wavefront-renderer.jl:185-362 (stages 1, 3, 4 of its wavefront renderer).
#pragma once
#include "rc_device.cuh"

struct RcCamera {  // passed by value to the kernels
    float pos[3];
    float right[3], up[3], forward[3];  // look-at basis (lookat != 0)
    float half_width, half_height;      // look-at: tan(fov/2) extents;  pinhole: half_width = aspect, forward[2] = focal_length
    uint32_t lookat;
    uint32_t jitter;  // 0: pixel centres (jitter = 0.5), 1: counter RNG dims 0/1 of the ray index (the reference calls rand(Vec2f))
};

#ifndef RC_MAX_LIGHTS
#define RC_MAX_LIGHTS 16
#endif
struct RcLights {  // point-light positions by value
    float pos[RC_MAX_LIGHTS][3];
    uint32_t n;
    float bias;  // shadow_bias (the reference hard-codes 0.01f0, :305)
};

// Ray `ray_idx` (0-based) of generate_primary_rays! (:185-213) / generate_primary_rays_lookat! (:219-253):
// ray_idx = ((y-1)*width + (x-1)) * n_samples + (s-1) with 1-based pixel (x, y) and sample s.
RC_HD rc_ray rc_primary_ray(const RcCamera &cam, uint32_t width, uint32_t height, uint32_t n_samples, unsigned long long seed, unsigned long long ray_idx) {
    const unsigned long long pixel = ray_idx / n_samples;
    const float x = (float)(uint32_t)(pixel % width + 1ull), y = (float)(uint32_t)(pixel / width + 1ull);
    const float j1 = cam.jitter ? rc_rng_uniform(seed, ray_idx, 0) : 0.5f;
    const float j2 = cam.jitter ? rc_rng_uniform(seed, ray_idx, 1) : 0.5f;
    // 2 * (x - 0.5 + j1) / W - 1   and   1 - 2 * (y - 0.5 + j2) / H
    const float u = x_sub(x_div(x_mul(2.0f, x_add(x_sub(x, 0.5f), j1)), (float)width), 1.0f);
    const float v = x_sub(1.0f, x_div(x_mul(2.0f, x_add(x_sub(y, 0.5f), j2)), (float)height));
    f3 d;
    if (cam.lookat) {
        const float a = x_mul(u, cam.half_width), b = x_mul(v, cam.half_height);
        d = mk3(x_add(x_add(cam.forward[0], x_mul(cam.right[0], a)), x_mul(cam.up[0], b)),
                x_add(x_add(cam.forward[1], x_mul(cam.right[1], a)), x_mul(cam.up[1], b)),
                x_add(x_add(cam.forward[2], x_mul(cam.right[2], a)), x_mul(cam.up[2], b)));
    } else {
        d = mk3(x_mul(u, cam.half_width), v, cam.forward[2]);
    }
    d = x_normalize(d);
    rc_ray r;
    r.origin[0] = cam.pos[0]; r.origin[1] = cam.pos[1]; r.origin[2] = cam.pos[2];
    r.dir[0] = d.x; r.dir[1] = d.y; r.dir[2] = d.z;
    r.tmin = 0.0f;
    r.tmax = INFINITY;
    return r;
}

// dummy ray of a sky hit (:319): t_max = 0 marks "no shadow test"
RC_HD rc_ray rc_dummy_shadow_ray() {
    rc_ray r;
    r.origin[0] = r.origin[1] = r.origin[2] = 0.0f;
    r.dir[0] = r.dir[1] = 0.0f; r.dir[2] = 1.0f;
    r.tmin = 0.0f;
    r.tmax = 0.0f;
    return r;
}

// Shadow ray of generate_shadow_rays! (:277-330) for one primary hit and one light.
//   normals9: the hit triangle's three vertex normals (BLAS-local), interpolated with bary = (1-u-v, u, v), carried to world
//   space with the inverse-transpose of the instance transform (identity instances — all the reference renderer uses — reproduce
//   the reference's expression bit for bit), normalised once.
RC_HD rc_ray rc_shadow_ray(const rc_ray &ray, const rc_hit &hit, const float *normals9, const float *inv_transform, const float light[3], float bias) {
    if (!hit.hit) return rc_dummy_shadow_ray();
    const float w0 = x_sub(x_sub(1.0f, hit.bary_u), hit.bary_v), w1 = hit.bary_u, w2 = hit.bary_v;
    f3 nl = mk3(x_add(x_add(x_mul(normals9[0], w0), x_mul(normals9[3], w1)), x_mul(normals9[6], w2)),
                x_add(x_add(x_mul(normals9[1], w0), x_mul(normals9[4], w1)), x_mul(normals9[7], w2)),
                x_add(x_add(x_mul(normals9[2], w0), x_mul(normals9[5], w1)), x_mul(normals9[8], w2)));
    const float *m = inv_transform;  // rows of [R^-1 | t]; n_world = (R^-1)^T n_local
    f3 nw = mk3(x_add(x_add(x_mul(m[0], nl.x), x_mul(m[4], nl.y)), x_mul(m[8], nl.z)),
                x_add(x_add(x_mul(m[1], nl.x), x_mul(m[5], nl.y)), x_mul(m[9], nl.z)),
                x_add(x_add(x_mul(m[2], nl.x), x_mul(m[6], nl.y)), x_mul(m[10], nl.z)));
    const f3 n = x_normalize(nw);
    const f3 p = mk3(x_add(ray.origin[0], x_mul(ray.dir[0], hit.t)), x_add(ray.origin[1], x_mul(ray.dir[1], hit.t)),
                     x_add(ray.origin[2], x_mul(ray.dir[2], hit.t)));
    const f3 so = mk3(x_add(p.x, x_mul(n.x, bias)), x_add(p.y, x_mul(n.y, bias)), x_add(p.z, x_mul(n.z, bias)));
    const f3 lv = mk3(x_sub(light[0], so.x), x_sub(light[1], so.y), x_sub(light[2], so.z));
    const f3 sd = x_normalize(lv);
    rc_ray r;
    r.origin[0] = so.x; r.origin[1] = so.y; r.origin[2] = so.z;
    r.dir[0] = sd.x; r.dir[1] = sd.y; r.dir[2] = sd.z;
    r.tmin = 0.0f;
    r.tmax = x_sqrt(x_add(x_add(x_mul(lv.x, lv.x), x_mul(lv.y, lv.y)), x_mul(lv.z, lv.z)));  // norm(light_vec)
    return r;
}

// geometric normal of a sorted triangle (fallback when the caller supplied no vertex normals)
RC_HD f3 rc_geometric_normal(const RcTri &t) {
    f3 p1 = mk3(t.v0[0], t.v0[1], t.v0[2]), p2 = mk3(t.v1[0], t.v1[1], t.v1[2]), p3 = mk3(t.v2[0], t.v2[1], t.v2[2]);
    return x_normalize(x_cross(x_sub3(p2, p1), x_sub3(p3, p1)));
}
