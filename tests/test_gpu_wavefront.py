"""Wavefront stages on device-resident queues through the C ABI (docs/src/wavefront-renderer.jl:185-362)
against the oracle: generated rays bit-exact, visibility exact outside the graze class."""
import numpy as np
import pytest

from engines import GpuEngine, OracleEngine
from oracle import oracle as orc
import raycore_b200 as rc
from raycore_b200 import HIT_DTYPE, RAY_DTYPE
from raycore_b200 import workloads as W
from test_wavefront import flipped, random_normals, shadow_scene

pytestmark = pytest.mark.gpu
F = np.float32


def test_primary_rays_match_oracle():
    ge = GpuEngine(shadow_scene())
    tl = ge.tlas
    for jitter in (False, True):
        q = tl.generate_primary_rays(61, 33, (0.5, -1.0, 2.0), 1.7, 61 / 33, n_samples=3, seed=11, jitter=jitter)
        want = orc.generate_primary_rays(61, 33, 3, (0.5, -1.0, 2.0), 1.7, 61 / 33, 11, jitter)
        assert q.download().tobytes() == want.tobytes()
        right, up, fwd = np.array([0.8, 0, 0.6], F), np.array([0, 1, 0], F), np.array([-0.6, 0, 0.8], F)
        q = tl.generate_primary_rays_lookat(40, 25, (3, 2, -4), right, up, fwd, 0.6, 0.375, n_samples=2, seed=5, jitter=jitter)
        want = orc.generate_primary_rays_lookat(40, 25, 2, (3, 2, -4), right, up, fwd, 0.6, 0.375, 5, jitter)
        assert q.download().tobytes() == want.tobytes()
    # empty image: nothing to do, no error
    assert tl.generate_primary_rays(0, 5, (0, 0, 0), 1.0, 1.0).count == 0


def _pipeline(tl, oe, rays_q, lights, normals_by_blas=None):
    hits_q = tl.intersect_rays(rays_q)
    rays, hits = rays_q.download(), hits_q.download()
    sh_q = tl.generate_shadow_rays(rays_q, hits_q, lights)
    sh = sh_q.download()
    want = oe.tlas.generate_shadow_rays(rays, hits.view(orc.HIT_DTYPE), lights, 0.01, normals_by_blas)
    assert sh.tobytes() == want.tobytes()  # same arithmetic on the same (GPU) hits
    vis_q = tl.test_shadow_rays(sh_q)
    vis = vis_q.download()
    # stage 4 is any_hit on the same rays
    any_q = tl.intersect_rays(sh_q, any_hit=True)
    live = sh["t_max"] > 0
    assert np.array_equal(vis, (live & (any_q.download()["hit"] == 0)).astype(np.uint8))
    # fused stages 3+4: identical bytes, no shadow-ray queue
    fused = tl.shadow_visibility(rays_q, hits_q, lights).download()
    assert np.array_equal(fused, vis)
    # against the oracle: only a graze (an occluder the reference's slab test culls) may differ, and only towards "occluded"
    vo = oe.tlas.test_shadow_rays(sh)
    diff = np.nonzero(vo != vis)[0]
    assert len(diff) <= max(1, 2e-5 * len(vis)), (len(diff), len(vis))
    assert (vis[diff] == 0).all()
    return hits, sh, vis


def test_shadow_pipeline_known_scene():
    pushes = shadow_scene()
    ge, oe = GpuEngine(pushes), OracleEngine(pushes)
    tl = ge.tlas
    rays_q = tl.generate_primary_rays(64, 64, (0, 0, 0), 1.0, 1.0, jitter=False)
    lights = np.array([[0, 0, -1], [6, 0, 4.5]], F)
    hits, sh, vis = _pipeline(tl, oe, rays_q, lights)
    vis = vis.reshape(64, 64, 2)
    hit = hits["hit"].reshape(64, 64)
    assert 0 < hit.sum() < hit.size  # some sky
    assert (vis[hit == 0] == 0).all()  # sky hits: dummy rays, "not visible" (:357)
    on_blocker = hits["instance_custom_index"].reshape(64, 64) == 9
    assert on_blocker.any() and (vis[on_blocker & (hit == 1)][:, 0] == 1).all()  # nothing between the blocker and light 0
    floor = (hits["instance_custom_index"].reshape(64, 64) == 7) & (hit == 1)
    assert (vis[floor][:, 0] == 1).sum() > 0  # lit floor


@pytest.mark.parametrize("with_normals", [False, True])
def test_shadow_pipeline_instanced(with_normals):
    sphere = W.uv_sphere(24)
    box = W.box_mesh()
    # a degenerate face in the middle of the submitted soup: normals are given per SUBMITTED face, hit.primitive_id counts kept faces
    box = np.concatenate([box[:5], np.zeros((1, 9), F), box[5:]])
    xf_s = W.random_trs(40, seed=3, extent=6.0)
    xf_b = W.random_trs(25, seed=4, extent=6.0)
    pushes = [(sphere, None, xf_s, None), (box, None, xf_b, None)]
    ge, oe = GpuEngine(pushes), OracleEngine(pushes)
    tl = ge.tlas
    normals_by_blas = None
    if with_normals:
        ns, nb = random_normals(len(sphere), 1), random_normals(len(box), 2)
        tl.set_normals(ge.handles[0], ns)
        tl.set_normals(ge.handles[1], nb)
        keep_s = np.array([not orc.is_degenerate(v) for v in sphere])
        keep_b = np.array([not orc.is_degenerate(v) for v in box])
        assert not keep_b.all()
        assert np.array_equal(np.nonzero(keep_b)[0], tl.read_blas_faces(2))
        normals_by_blas = [ns[keep_s], nb[keep_b]]
    rays_q = tl.queue(RAY_DTYPE, 200_000).upload(W.box_rays(200_000, seed=9, half=7.0))
    lights = np.array([[0, 9, 0], [5, -3, 2], [-4, 0, -6]], F)
    hits, sh, vis = _pipeline(tl, oe, rays_q, lights, normals_by_blas)
    assert 0.05 < hits["hit"].mean() < 0.95
    assert 0 < vis.sum() < (sh["t_max"] > 0).sum()
    # errors mirror the reference's argument checks
    import raycore_b200 as rc

    with pytest.raises(rc.RaycoreError):
        tl.set_normals(ge.handles[1], random_normals(3, 0))  # wrong face count
    with pytest.raises(rc.RaycoreError):
        tl.shadow_visibility(rays_q, tl.intersect_rays(rays_q), np.zeros((17, 3), F))  # > RC_MAX_LIGHTS
    # geometry update drops the handle's normals (geometric normals again)
    tl.update(ge.handles[1], W.box_mesh())
    tl.sync()
    oe2 = OracleEngine([(sphere, None, xf_s, None), (W.box_mesh(), None, xf_b, None)])
    _pipeline(tl, oe2, rays_q, lights, None if not with_normals else [normals_by_blas[0], None])


def test_shadow_large_mesh_sample():
    # the bench mesh class (bumpy sphere), one light inside: every primary hit casts one shadow ray
    mesh = W.bumpy_sphere(200)
    pushes = [(mesh, None, W.identity3x4()[None], None)]
    ge, oe = GpuEngine(pushes), OracleEngine(pushes)
    tl = ge.tlas
    n = 300_000
    rays_q = tl.queue(RAY_DTYPE, n).upload(W.interior_rays(n, seed=2))
    hits, sh, vis = _pipeline(tl, oe, rays_q, np.array([[0.1, 0.2, 0.0]], F))
    assert hits["hit"].all() and (sh["t_max"] > 0).all()


def test_shadow_deep_trees_fixup():
    """Occlusion rays through exponentially nested geometry overflow the 32-entry short stack: the scheduler kernel marks them
    and k_shadow_fixup redoes them (queue and fused sources) — results must still match the oracle."""
    from test_gpu_parity import _deep_scene

    blas = _deep_scene(20)
    g = [2.0 ** -i for i in range(14)]
    xf = np.stack([W.trs3x4((4 * a, 4 * b, 4 * g[(i + j) % 14]), (1, 0, 0, 0), 1.0) for i, a in enumerate(g) for j, b in enumerate(g)])
    pushes = [(blas, None, xf, None)]
    ge, oe = GpuEngine(pushes), OracleEngine(pushes)
    tl = ge.tlas
    rs = np.random.RandomState(0)
    n = 4096
    d = (np.array([1, 1, 1], F) + rs.uniform(-0.9, 0.9, (n, 3))).astype(F)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = W.make_rays(np.full((n, 3), 1e-9, F), d)
    tl.adapt().trace_closest(rays, counters=True)
    assert tl.counters()["max_stack"] > 32
    rays["t_max"] = rs.uniform(0.0, 6.0, n).astype(F)
    rays["t_max"][::7] = 0  # dummy rays
    q = tl.queue(RAY_DTYPE, n).upload(rays)
    vis = tl.test_shadow_rays(q).download()
    vo = oe.tlas.test_shadow_rays(rays)
    diff = np.nonzero(vis != vo)[0]
    assert len(diff) <= 2 and (vis[diff] == 0).all()  # graze class only
    assert (vis[::7] == 0).all() and 0 < vis.sum() < n
    # fused source on the same scene: primary rays from far outside towards the nest, lights inside it
    org = (np.array([6, 6, 6], F) + rs.uniform(-1, 1, (n, 3))).astype(F)
    prim = W.make_rays(org, (-org / np.linalg.norm(org, axis=1, keepdims=True)).astype(F))
    pq = tl.queue(RAY_DTYPE, n).upload(prim)
    _pipeline(tl, oe, pq, np.array([[1.3e-3, 0.9e-3, 1.1e-3], [4.3, 4.6, 4.45]], F))  # off the 2^-k lattice the geometry sits on


def test_wavefront_on_emptied_tlas():
    """Every handle deleted, then sync!: closest_hit misses (test/test_tlas_stress.jl:808-831) and the shadow stages must not walk a
    TLAS that has no nodes — a live shadow ray is visible, a dummy one is not."""
    tl = rc.TLAS()
    h = tl.push(W.bumpy_sphere(12), None)
    tl.sync()
    tl.delete(h)
    tl.sync()
    assert tl.n_instances() == 0
    rays_q = tl.generate_primary_rays(16, 16, (0, 0, -3), 1.0, 1.0, jitter=False)
    hits_q = tl.intersect_rays(rays_q)
    assert (hits_q.download()["hit"] == 0).all()
    lights = np.array([[0, 0, -1], [6, 0, 4.5]], F)
    # staged: every primary ray missed, so every shadow ray is the dummy ray
    sh_q = tl.generate_shadow_rays(rays_q, hits_q, lights)
    assert (sh_q.download()["t_max"] == 0).all()
    assert (tl.test_shadow_rays(sh_q).download() == 0).all()
    assert (tl.shadow_visibility(rays_q, hits_q, lights).download() == 0).all()
    # caller-made live shadow rays: nothing can occlude them
    live = W.box_rays(256, 3, half=2.0)
    live["t_max"] = 5.0
    live["t_max"][::4] = 0.0
    q = tl.queue(RAY_DTYPE, len(live)).upload(live)
    vis = tl.test_shadow_rays(q).download()
    assert np.array_equal(vis, (live["t_max"] > 0).astype(np.uint8))
    tl.free()
