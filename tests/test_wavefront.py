"""Wavefront stages either side of the trace (docs/src/wavefront-renderer.jl:185-362), CPU side:
the oracle's restatement against an independent numpy float32 evaluation of the reference's expressions and against
known answers of a constructed scene, and the library's per-element device bodies (rc_wave_core.cuh, run through
tests/hostsim) bit for bit against the oracle."""
import numpy as np
import pytest

import hostsim_py as hs
from engines import HostsimEngine, OracleEngine
from oracle import oracle as orc
from raycore_b200 import workloads as W

F = np.float32


def np_normalize(v):
    """StaticArrays normalize: inv(norm) * v with a left-fold sum of squares, all in Float32."""
    v = v.astype(F)
    n = np.sqrt((v[..., 0] * v[..., 0] + v[..., 1] * v[..., 1]) + v[..., 2] * v[..., 2], dtype=F)
    return (F(1) / n)[..., None] * v


def np_primary_uv(width, height, ns, seed, jitter):
    idx = np.arange(width * height * ns, dtype=np.uint64)
    pixel = idx // np.uint64(ns)
    x = (pixel % np.uint64(width) + np.uint64(1)).astype(F)
    y = (pixel // np.uint64(width) + np.uint64(1)).astype(F)
    j1 = W.rng_uniform(seed, idx, 0) if jitter else np.full(len(idx), 0.5, F)
    j2 = W.rng_uniform(seed, idx, 1) if jitter else np.full(len(idx), 0.5, F)
    u = F(2) * (x - F(0.5) + j1) / F(width) - F(1)  # wavefront-renderer.jl:202
    v = F(1) - F(2) * (y - F(0.5) + j2) / F(height)  # :203
    return u.astype(F), v.astype(F)


@pytest.mark.parametrize("jitter", [False, True])
def test_primary_rays_pinhole_restatement(jitter):
    w, h, ns, seed = 37, 23, 3, 11
    pos, focal, aspect = (0.5, -1.0, 2.0), F(1.7), F(37 / 23)
    u, v = np_primary_uv(w, h, ns, seed, jitter)
    d = np_normalize(np.stack([u * aspect, v, np.full(len(u), focal, F)], axis=1))
    got = orc.generate_primary_rays(w, h, ns, pos, focal, aspect, seed, jitter)
    assert np.array_equal(got["d"].view(np.uint32), d.view(np.uint32))
    assert np.array_equal(got["o"], np.tile(np.asarray(pos, F), (len(u), 1)))
    assert (got["t_min"] == 0).all() and np.isinf(got["t_max"]).all()
    sim = hs.primary_rays(w, h, ns, pos, focal_length=focal, aspect=aspect, seed=seed, jitter=jitter)
    assert sim.tobytes() == got.tobytes()


@pytest.mark.parametrize("jitter", [False, True])
def test_primary_rays_lookat_restatement(jitter):
    w, h, ns, seed = 16, 9, 2, 5
    pos = np.array([3, 2, -4], F)
    fwd = np_normalize(np.array([[-3, -2, 4]], F))[0]
    right = np_normalize(np.cross(fwd, np.array([0, 1, 0], F)).astype(F)[None])[0]
    up = np.cross(right, fwd).astype(F)
    hw, hh = F(0.6), F(0.3375)
    u, v = np_primary_uv(w, h, ns, seed, jitter)
    a, b = u * hw, v * hh
    d = np_normalize((fwd[None] + right[None] * a[:, None]) + up[None] * b[:, None])  # :242-246
    got = orc.generate_primary_rays_lookat(w, h, ns, pos, right, up, fwd, hw, hh, seed, jitter)
    assert np.array_equal(got["d"].view(np.uint32), d.astype(F).view(np.uint32))
    sim = hs.primary_rays(w, h, ns, pos, lookat=(right, up, fwd, hw, hh), seed=seed, jitter=jitter)
    assert sim.tobytes() == got.tobytes()


def test_primary_ray_layout_and_centre():
    # 1-based pixel (x, y), sample s -> ((y-1)*W + (x-1))*NS + (s-1); without jitter pixel x maps to ndc 2x/W - 1
    w, h, ns = 4, 2, 2
    r = orc.generate_primary_rays(w, h, ns, (0, 0, 0), 1.0, 1.0, 0, False).reshape(h, w, ns)
    assert np.array_equal(r[:, :, 0], r[:, :, 1])  # no jitter: samples coincide
    x = 2  # u = 2*2/4 - 1 = 0
    assert r["d"][0, x - 1, 0][0] == 0
    assert r["d"][1, 3, 0][0] > 0 and r["d"][1, 3, 0][1] < 0  # bottom-right pixel: +x, -y (v = 1 - 2y/H = -1)
    assert np.allclose(np.linalg.norm(r["d"].reshape(-1, 3), axis=1), 1, atol=1e-6)


def flipped(mesh):
    """reverse the winding so that the geometric normal points to -z"""
    m = mesh.reshape(-1, 3, 3)[:, ::-1, :]
    return np.ascontiguousarray(m.reshape(-1, 9))


def shadow_scene():
    floor = flipped(W.quad_mesh(z=5.0, half=4.0))
    blocker = flipped(W.quad_mesh(z=3.0, half=0.5))
    I = W.identity3x4()
    return [(floor, None, I[None], np.array([7], np.uint32)), (blocker, None, I[None], np.array([9], np.uint32))]


def test_shadow_known_answers():
    eng = OracleEngine(shadow_scene())
    light = np.array([[0, 0, -1]], F)
    o = np.zeros(3, F)
    targets = np.array([[0, 0, 3], [0.3, 0.2, 5.0], [3.0, 3.0, 5.0], [30, 0, 5]], F)  # blocker, floor in the umbra (behind the blocker as seen
    # from the camera this one hits the blocker too), lit floor, sky
    d = (targets - o) / np.linalg.norm(targets - o, axis=1, keepdims=True)
    rays = orc.make_rays(np.tile(o, (4, 1)), d)
    hits = eng.tlas.closest_hit(rays)
    assert list(hits["hit"]) == [1, 1, 1, 0]
    assert list(hits["instance_custom_index"][:3]) == [9, 9, 7]
    sh = eng.tlas.generate_shadow_rays(rays, hits, light)
    # sky hit -> the dummy ray (:319)
    assert sh["t_max"][3] == 0 and tuple(sh["d"][3]) == (0, 0, 1) and tuple(sh["o"][3]) == (0, 0, 0)
    # origin = hit point + 0.01 * normal (normal = -z), t_max = distance to the light
    assert np.allclose(sh["o"][0], [0, 0, 3 - 0.01], atol=1e-6)
    assert np.isclose(sh["t_max"][0], 4 - 0.01, atol=1e-5)
    assert np.allclose(sh["d"][0], [0, 0, -1], atol=1e-6)
    vis = eng.tlas.test_shadow_rays(sh)
    assert list(vis) == [1, 1, 1, 0]
    # a floor point in the umbra, reached by a ray that passes beside the blocker
    o2 = np.array([[6.0, 0, 0]], F)
    tgt = np.array([[0.2, 0.1, 5.0]], F)
    d2 = (tgt - o2) / np.linalg.norm(tgt - o2)
    r2 = orc.make_rays(o2, d2)
    h2 = eng.tlas.closest_hit(r2)
    assert h2["hit"][0] == 1 and h2["instance_custom_index"][0] == 7
    s2 = eng.tlas.generate_shadow_rays(r2, h2, light)
    assert list(eng.tlas.test_shadow_rays(s2)) == [0]
    # the light beyond t_max does not count as an occluder: a light between the blocker and the floor sees the floor
    s3 = eng.tlas.generate_shadow_rays(r2, h2, np.array([[0.2, 0.1, 4.0]], F))
    assert list(eng.tlas.test_shadow_rays(s3)) == [1]


def test_shadow_rays_two_lights_layout():
    eng = OracleEngine(shadow_scene())
    rays = orc.generate_primary_rays(8, 8, 1, (0, 0, 0), 1.0, 1.0, 0, False)
    hits = eng.tlas.closest_hit(rays)
    lights = np.array([[0, 0, -1], [2, 2, 0]], F)
    both = eng.tlas.generate_shadow_rays(rays, hits, lights)
    for l in range(2):  # shadow ray (k, l) sits at k*NLights + l (:300)
        one = eng.tlas.generate_shadow_rays(rays, hits, lights[l : l + 1])
        assert both[l::2].tobytes() == one.tobytes()


def random_normals(n_faces, seed):
    rng = np.random.default_rng(seed)
    n = rng.normal(size=(n_faces, 3, 3)).astype(F)
    n /= np.linalg.norm(n, axis=2, keepdims=True)
    return np.ascontiguousarray(n.reshape(n_faces, 9).astype(F))


@pytest.mark.parametrize("with_normals", [False, True])
def test_shadow_rays_device_body_matches_oracle(with_normals):
    # instanced, non-identity transforms: the library's rc_shadow_ray (host simulation) against the oracle, bit for bit
    sphere = W.uv_sphere(12)
    box = W.box_mesh()
    xf_s = W.random_trs(3, seed=3, extent=3.0)
    xf_b = W.random_trs(2, seed=4, extent=3.0)
    pushes = [(sphere, None, xf_s, np.array([1, 2, 3], np.uint32)), (box, None, xf_b, None)]
    oe, he = OracleEngine(pushes), HostsimEngine(pushes)
    rays = W.box_rays(4000, seed=9, half=3.0)
    hits = oe.tlas.closest_hit(rays)
    assert 150 < hits["hit"].sum() < 4000
    lights = np.array([[0, 8, 0], [5, -3, 2], [-4, 0, -6]], F)
    normals = None
    if with_normals:
        # per BLAS, indexed by primitive_id (degenerate-filtered input order)
        normals = [random_normals(len(oe.tris[1]), 1), random_normals(len(oe.tris[2]), 2)]
    a = oe.tlas.generate_shadow_rays(rays, hits, lights, 0.01, normals)
    b = he.scene.shadow_rays(rays, hits, lights, 0.01, normals)
    assert a.tobytes() == b.tobytes()
    miss = np.repeat(hits["hit"] == 0, 3)
    assert (a["t_max"][miss] == 0).all() and (a["t_max"][~miss] > 0).all()
    # visibility of these rays: wide host simulation against the oracle (hit/miss may only differ in the graze class)
    vo = oe.tlas.test_shadow_rays(a)
    hv = he.scene.trace(a, any_hit=True)
    vh = ((a["t_max"] > 0) & (hv["hit"] == 0)).astype(np.uint8)
    assert (vo != vh).sum() <= 1
    assert 0 < vo.sum() < len(vo)
