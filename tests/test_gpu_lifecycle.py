"""Lifecycle / handle-API tests mirroring test/test_instanced_bvh.jl:417-661, test/test_mesh_update.jl,
test/test_abstract_accel_contract.jl and test/test_tlas_stress.jl through the host mirror of the reference API."""
import numpy as np
import pytest

from raycore_b200 import workloads as W
import raycore_b200 as rc
from raycore_b200 import Ray, TLAS, TLASHandle, RaycoreError

pytestmark = pytest.mark.gpu

DOWN = (0.0, 0.0, -1.0)


def sphere(n, c=(0, 0, 0)):
    return W.uv_sphere(n, c, 1.0)


def hit_at(tlas, x, y=0.02, z=5.0):
    return tlas.closest_hit(Ray((x, y, z), DOWN))


def test_abstract_accel_contract():
    # test/test_abstract_accel_contract.jl:7-34
    tlas, handles = rc.tlas_from_meshes([sphere(8)])
    assert tlas.n_instances() == 1 and tlas.n_geometries() == 1
    assert isinstance(tlas.world_bound(), rc.Bounds3)
    assert tlas.wait_for_gpu() is tlas
    assert isinstance(tlas.adapt(), rc.StaticTLAS)


def test_return_types_and_miss_sentinel():
    # test/test_instanced_bvh.jl:595-624, test/test_intersection.jl:121-142
    st = rc.build_static_tlas([W.quad_mesh()], lambda mi, fi: 40 + fi)
    hit, tri, t, bary, inst = st.closest_hit(Ray((0.5, 0.25, 1.0), DOWN))
    assert hit is True and isinstance(tri, rc.Triangle) and t.dtype == np.float32 and bary.shape == (3,) and inst.dtype == np.uint32
    assert inst == 1 and tri.metadata in (41, 42) and abs(t - 1) < 1e-6 and abs(bary.sum() - 1) < 1e-6
    p = (bary[:, None] * tri.vertices).sum(0)
    assert np.allclose(p, (0.5, 0.25, 0.0), atol=1e-6)
    hit, tri, t, bary, inst = st.closest_hit(Ray((5, 5, 1.0), DOWN))
    assert hit is False and t == 0 and inst == 0 and not tri.vertices.any() and not bary.any()
    hit, *_ = st.any_hit(Ray((0.5, 0.25, 1.0), DOWN))
    assert hit is True


def test_handle_api_and_multi_transform_push():
    # test/test_instanced_bvh.jl:417-468
    tlas = TLAS()
    h1 = tlas.push(W.quad_mesh())
    h2 = tlas.push(W.quad_mesh(), [W.translation3x4((5, 0, 0)), W.translation3x4((10, 0, 0))], instance_ids=[7, 8])
    assert isinstance(h1, TLASHandle) and h1 != h2 and tlas.is_valid(h1) and tlas.is_valid(h2)
    assert tlas.n_geometries() == 2 and tlas.n_instances() == 3 and tlas.n_instances(h2) == 2
    assert tlas.dirty
    tlas.sync()
    assert not tlas.dirty and tlas.last_sync_action == rc.RC_SYNC_REBUILD
    assert len(tlas.read_tlas_nodes()) == 5
    inst = tlas.get_instances(h2)
    assert list(inst["instance_id"]) == [7, 8] and inst["transform"][1][3] == 10
    assert tlas.get_instance(h2, 2)["instance_id"] == 8
    hits = tlas.trace_closest(W.make_rays([(0, 0, 1), (5, 0, 1), (10, 0, 1), (20, 0, 1)], DOWN))
    assert list(hits["hit"]) == [1, 1, 1, 0] and list(hits["instance_custom_index"][:3]) == [0, 7, 8]
    with pytest.raises(ValueError):
        tlas.push(W.quad_mesh(), [W.identity3x4()] * 2, instance_ids=[1])  # ArgumentError, :664-666


def test_update_transforms_refit_keeps_static_identity():
    # test/test_instanced_bvh.jl:491-538, test/test_mesh_update.jl:184-227, test/test_tlas_stress.jl:623-650
    tlas = TLAS()
    h = tlas.push(sphere(8))
    hm = tlas.push(sphere(8), [W.translation3x4((4, 0, 0)), W.translation3x4((8, 0, 0))])
    st = tlas.adapt()
    assert hit_at(tlas, 0.01)[0] and not hit_at(tlas, 20.01)[0]
    tlas.update_transform(h, W.translation3x4((20, 0, 0)))
    assert tlas.transforms_dirty and not tlas.dirty
    tlas.sync()
    assert tlas.last_sync_action == rc.RC_SYNC_REFIT and tlas.static_tlas is st and not tlas.transforms_dirty
    assert not hit_at(tlas, 0.01)[0] and hit_at(tlas, 20.01)[0]
    assert tlas.world_bound().p_max[0] >= 20.9
    tlas.update_transforms(hm, [W.translation3x4((4, 3, 0)), W.translation3x4((8, 3, 0))])
    tlas.sync()
    assert tlas.static_tlas is st and hit_at(tlas, 4.01, 3.02)[0] and not hit_at(tlas, 4.01, 0.02)[0]
    # clean sync is a no-op
    tlas.sync()
    assert tlas.last_sync_action == rc.RC_SYNC_NONE
    with pytest.raises(RaycoreError):
        tlas.update_transform(hm, W.identity3x4())  # "use update_transforms! for multiple" (:759)
    with pytest.raises(RaycoreError):
        tlas.update_transforms(hm, [W.identity3x4()])  # count mismatch (:788)


def test_push_delete_sync_and_errors():
    # test/test_instanced_bvh.jl:540-589, test/test_tlas_stress.jl:585-617, :769-802
    tlas = TLAS()
    hs = [tlas.push(sphere(6, (3.0 * k, 0, 0))) for k in range(5)]
    tlas.sync()
    st0 = tlas.static_tlas
    assert tlas.delete(hs[0]) is True and tlas.delete(hs[0]) is False  # idempotent false
    assert tlas.delete(TLASHandle(9999)) is False
    assert tlas.n_instances() == 4 and tlas.n_total_instances() == 5 and not tlas.is_valid(hs[0])
    tlas.delete(hs[1])
    h_new = tlas.push(sphere(6, (3.0, 0, 0)))  # mixed delete + push without sync in between
    tlas.sync()
    assert tlas.static_tlas is not st0  # rebuild replaces static_tlas (test_tlas_stress.jl:676)
    with pytest.raises(RaycoreError):
        st0.trace_closest(W.make_rays([(0, 0, 5)], DOWN))  # stale adapted form
    assert tlas.n_instances() == 4 and tlas.n_total_instances() == 4 and tlas.n_geometries() == 4
    hits = tlas.trace_closest(W.make_rays([(3.0 * k + 0.01, 0.02, 5.0) for k in range(5)], DOWN))
    assert list(hits["hit"]) == [0, 1, 1, 1, 1]
    for fn in (lambda: tlas.get_instance(hs[0]), lambda: tlas.update_transform(hs[0], W.identity3x4()), lambda: tlas.update(hs[1], sphere(6)),
               lambda: tlas.get_instances(TLASHandle(4242))):
        with pytest.raises(RaycoreError):
            fn()
    s = tlas.sizes()
    assert s["pending_deletes"] == 0 and s["tlas_nodes"] == 7
    with pytest.raises(RaycoreError):
        tlas.push(np.zeros((3, 9), np.float32))  # "Geometry has no valid triangles" (:601)
    assert tlas.is_valid(h_new)


def test_empty_tlas_traces_miss():
    # test/test_tlas_stress.jl:808-831
    tlas = TLAS()
    tlas.sync()
    assert tlas.trace_closest(W.make_rays([(0, 0, 1)], DOWN))["hit"][0] == 0
    h = tlas.push(sphere(6))
    tlas.sync()
    assert hit_at(tlas, 0.01)[0]
    tlas.delete(h)
    tlas.sync()
    assert tlas.n_instances() == 0 and tlas.sizes()["tlas_nodes"] == 0
    assert tlas.trace_any(W.make_rays([(0.01, 0.02, 5)], DOWN))["hit"][0] == 0
    assert np.isinf(tlas.world_bound().p_min).all()


def test_mesh_update_schedule():
    # test/test_mesh_update.jl:89-116: delete! + push! + sync! per frame, tessellation schedule, t ≈ 4 - z
    tlas = TLAS()
    h = tlas.push(sphere(32))
    tlas.sync()
    for i, n in enumerate([32, 8, 48, 12, 64, 16, 8, 32, 96, 16]):
        z = 0.1 * i
        tlas.delete(h)
        h = tlas.push(sphere(n, (0, 0, z)))
        tlas.sync()
        hit, tri, t, bary, inst = hit_at(tlas, 0.01)
        assert hit and abs(t - (4 - z)) < 0.1 and inst == 1
        s = tlas.sizes()
        nb = rc._lib.load().rc_blas_n_prims(tlas._ctx, 1)
        assert s["tlas_nodes"] == 1 and s["blas_prims"] == nb and s["blas_nodes"] == 2 * nb - 1  # exact flat-array invariants (:261-294)
    # update!(tlas, handle, new_geometry) path (:808-857)
    tlas.update(h, sphere(20, (0, 0, 2.0)))
    assert tlas.dirty
    tlas.sync()
    assert abs(hit_at(tlas, 0.01)[2] - 2.0) < 0.1


def test_stress_many_instances_refit_frames():
    # test/test_tlas_stress.jl:233-327 (scaled): 2000 instances, batch update + refit identity, moved instances miss at old positions
    rs = np.random.RandomState(0xC0FFEE & 0xFFFF)
    n = 2000
    pos = rs.uniform(-50, 50, (n, 3)).astype(np.float32)
    tlas = TLAS()
    h = tlas.push(W.box_mesh(), [W.translation3x4(p) for p in pos])
    st = tlas.adapt()
    rays = W.make_rays(pos + np.array([0.01, 0.02, 3.0], np.float32), DOWN)
    assert tlas.trace_closest(rays)["hit"].all()
    for frame in range(5):
        pos2 = pos + np.float32(200.0 * (frame + 1))
        tlas.update_transforms(h, [W.translation3x4(p) for p in pos2])
        tlas.sync()
        assert tlas.last_sync_action == rc.RC_SYNC_REFIT and tlas.static_tlas is st
        assert not tlas.trace_closest(rays)["hit"].any()
        assert tlas.trace_closest(W.make_rays(pos2 + np.array([0.01, 0.02, 3.0], np.float32), DOWN))["hit"].all()


def test_churn_invariants():
    # test/test_tlas_stress.jl:101-181 (scaled): random push/delete/update ops with exact invariants after each sync
    rs = np.random.RandomState(7)
    tlas = TLAS()
    live = {}
    for op in range(120):
        r = rs.rand()
        if r < 0.5 or not live:
            m = int(rs.randint(1, 4))
            h = tlas.push(W.box_mesh(), [W.translation3x4(rs.uniform(-20, 20, 3)) for _ in range(m)])
            live[h] = m
        elif r < 0.8:
            h = list(live)[rs.randint(len(live))]
            assert tlas.delete(h)
            del live[h]
        else:
            h = list(live)[rs.randint(len(live))]
            tlas.update_transforms(h, [W.translation3x4(rs.uniform(-20, 20, 3)) for _ in range(live[h])])
        if op % 7 == 0:
            tlas.sync()
            s = tlas.sizes()
            ni = sum(live.values())
            assert tlas.n_instances() == ni == tlas.n_total_instances() and tlas.n_geometries() == len(live)
            assert s["pending_deletes"] == 0 and s["tlas_nodes"] == (0 if ni == 0 else max(1, 2 * ni - 1))
            assert s["blas_prims"] == 12 * len(live) and s["blas_nodes"] == 23 * len(live)
            assert not tlas.dirty and not tlas.transforms_dirty
            for h, m in live.items():
                assert tlas.n_instances(h) == m


def test_extent_beyond_quantisation_range_is_refused():
    """Wide nodes quantise against 2^(e-127), e <= 230: geometry whose extent exceeds 255 * 2^103 is refused at push / sync
    (the reference's own triangle test overflows there) instead of being traversed with boxes that do not cover it."""
    import raycore_b200 as rc
    from raycore_b200 import workloads as W

    tl = rc.TLAS()
    big = (W.box_mesh() * np.float32(1e34)).astype(np.float32)
    with pytest.raises(rc.RaycoreError):
        tl.push(big, None)
    # still usable afterwards; a large but supported scene works
    ok = (W.box_mesh() * np.float32(1e8)).astype(np.float32)
    tl.push(ok, None)
    tl.sync()
    h = tl.trace_closest(W.make_rays([[1e6, 2e6, -1e9]], [[0, 0, 1]]))
    assert h["hit"][0] == 1 and np.isclose(h["t"][0], 1e9 - 0.5e8, rtol=1e-5)
    # instances flung apart beyond the range: refused at sync
    far = W.translation3x4((3e34, 0, 0))
    tl.push(W.box_mesh(), far)
    with pytest.raises(rc.RaycoreError):
        tl.sync()
    tl.free()


def test_update_transforms_from_device_array():
    """update_transforms! fed from a device-resident Mat3x4f array (the instance_buffer use case): same refit, same hits as the
    host-array path; arity errors as in :788."""
    mesh = W.uv_sphere(16)
    xf0 = W.random_trs(50, seed=1, extent=10.0)
    xf1 = W.random_trs(50, seed=2, extent=10.0)
    a, b = TLAS(), TLAS()
    ha, hb = a.push(mesh, list(xf0)), b.push(mesh, list(xf0))
    a.sync(); b.sync()
    rays = W.box_rays(20000, seed=3, half=12.0)
    sa = a.static_tlas
    a.update_transforms(ha, list(xf1))
    q = b.queue(np.float32, 50 * 12).upload(xf1.reshape(-1))
    b.update_transforms_device(hb, q)
    assert b.transforms_dirty
    a.sync(); b.sync()
    assert a.static_tlas is sa  # refit keeps the adapted object
    assert a.trace_closest(rays).tobytes() == b.trace_closest(rays).tobytes()
    assert np.array_equal(a.get_instances(ha)["inv_transform"], b.get_instances(hb)["inv_transform"])
    with pytest.raises(RaycoreError):
        b.update_transforms_device(hb, b.queue(np.float32, 49 * 12))
    a.free(); b.free()


def test_queries_from_several_host_threads():
    """The reference's callers issue closest_hit under Threads.@threads (src/kernels.jl:64,82).  Calls on one context serialise on the
    context lock: concurrent batched traces from host buffers (which share the context's staging buffers and work counter) and per-ray
    calls give exactly the single-threaded results."""
    import threading

    tlas = TLAS()
    tlas.push(W.bumpy_sphere(40), list(W.random_trs(20, 3, extent=5.0)))
    tlas.sync()
    batches = [np.concatenate([W.box_rays(30000 + 1000 * k, 10 + k, half=7.0), W.interior_rays(5000, 20 + k, radius=6.0)]) for k in range(4)]
    want_c = [tlas.trace_closest(b) for b in batches]
    want_a = [tlas.trace_any(b) for b in batches]
    errors = []

    def worker(k):
        try:
            for it in range(6):
                got = tlas.trace_any(batches[k]) if it % 2 else tlas.trace_closest(batches[k])
                ref = want_a[k] if it % 2 else want_c[k]
                if got.tobytes() != ref.tobytes():
                    errors.append((k, it, "batch differs"))
            r = batches[k][0]
            h = tlas.closest_hit(Ray(tuple(r["o"]), tuple(r["d"])))
            if bool(h[0]) != bool(want_c[k][0]["hit"]):
                errors.append((k, "per-ray call differs"))
        except Exception as e:  # noqa: BLE001
            errors.append((k, repr(e)))

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    tlas.free()
