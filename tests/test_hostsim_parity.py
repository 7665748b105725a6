"""CPU-side checks of the library's device code (tests/hostsim = the same RC_HD bodies compiled for the host):
builder output and both traversals against the oracle.  Runs without a GPU; the `-m gpu` tests repeat the same
comparisons through the C ABI on the real kernels."""
import numpy as np
import pytest

from oracle import oracle as orc
from raycore_b200 import workloads as W
import engines
import hostsim_py as hs
import kat
import parity


def _scene_instanced(n_inst=40, tess=10, seed=11):
    xf = W.random_trs(n_inst, seed, extent=6.0)
    return [(W.bumpy_sphere(tess), None, xf, np.arange(1, n_inst + 1, dtype=np.uint32)), (W.box_mesh(), None, W.random_trs(7, seed + 1, extent=6.0), None)]


def _rays_for(pushes, n, seed):
    r1 = W.box_rays(n // 2, seed, half=8.0)
    r2 = W.interior_rays(n - n // 2, seed + 5, radius=7.0)
    return np.concatenate([r1, r2])


def test_builder_bit_exact_vs_oracle():
    for verts in (kat.TRI, W.quad_mesh(), W.box_mesh(), W.uv_sphere(9), W.bumpy_sphere(23)):
        ob = orc.OracleBLAS.from_verts(verts)
        hb = hs.HsBlas(verts)
        assert hb.n == ob.n
        assert np.array_equal(hb.order(), ob.prims["input_index"])
        assert hb.nodes2().tobytes() == ob.nodes.tobytes(), "BVH2 differs from the reference restatement"
        assert np.array_equal(hb.root(), ob.root_aabb)
        assert hb.check_wide() == 0, "a quantised child box does not contain its exact box"


def test_collapse_fetch_once_form_is_bit_identical():
    """k_collapse_span collapses the spanning nodes with rc_collapse_node_cached (each BVH2 record fetched once); the host build runs it
    beside rc_collapse_node for every wide node (BLAS: 2-triangle leaves; TLAS: tagged instance leaves through leaf_map)."""
    for verts in (kat.TRI, W.quad_mesh(), W.box_mesh(), W.uv_sphere(9), W.bumpy_sphere(31)):
        hs.HsBlas(verts)
    engines.HostsimEngine(_scene_instanced())
    checked, bad = hs.collapse_cached_stats()
    assert checked > 2000 and bad == 0


def test_tlas_bit_exact_vs_oracle():
    pushes = _scene_instanced()
    o = engines.OracleEngine(pushes)
    h = engines.HostsimEngine(pushes)
    assert np.array_equal(o.instances["inv_transform"], h.instances["inv_transform"]), "mat3x4_inverse differs"
    assert h.scene.tlas_nodes2().tobytes() == o.tlas.nodes.tobytes()
    assert np.array_equal(h.scene.root(), o.tlas.root_aabb)


def test_reference_kats_on_device_code():
    kat.check_all(lambda p: engines.HostsimEngine(p, wide=True))
    kat.check_all(lambda p: engines.HostsimEngine(p, wide=False))


@pytest.mark.parametrize("any_hit", [False, True])
def test_reference_order_traversal_bit_exact(any_hit):
    pushes = _scene_instanced()
    o = engines.OracleEngine(pushes)
    h = engines.HostsimEngine(pushes, wide=False)
    rays = _rays_for(pushes, 6000, 3)
    a, b = h.trace(rays, any_hit=any_hit), o.trace(rays, any_hit=any_hit)
    assert a.tobytes() == b.tobytes()
    assert 0.2 < b["hit"].mean() < 1.0


def test_wide_traversal_parity_closest():
    pushes = _scene_instanced()
    o = engines.OracleEngine(pushes)
    h = engines.HostsimEngine(pushes, wide=True)
    rays = _rays_for(pushes, 20000, 9)
    a, b = h.trace(rays), o.trace(rays)
    ver = parity.make_graze_verifier(orc, rays, a, o.instances, o.tris)
    cls = parity.classify(a, b, ver)
    s = parity.assert_parity(cls, len(rays), label="hostsim wide closest")
    assert s["exact"] >= 0.995 * len(rays), s


def test_wide_traversal_parity_any():
    pushes = _scene_instanced()
    o = engines.OracleEngine(pushes)
    h = engines.HostsimEngine(pushes, wide=True)
    rays = _rays_for(pushes, 20000, 10)
    a, b = h.trace(rays, any_hit=True), o.trace(rays, any_hit=True)
    # any_hit returns the first accepted triangle in traversal order, so only the hit flag is order-independent;
    # the reported triangle must be a genuine exact-MT hit
    mism = np.nonzero(a["hit"] != b["hit"])[0]
    assert len(mism) <= 1, (len(mism), mism[:5])
    ver = parity.make_graze_verifier(orc, rays, a, o.instances, o.tris)
    idx = np.nonzero(a["hit"] == 1)[0][:3000]
    assert ver(idx).all()


def test_single_blas_million_scale_structure():
    # larger BLAS: wide boxes conservative, BVH2 identical, traversal parity on interior rays (all hit, multi-candidate)
    verts = W.bumpy_sphere(96)
    ob, hb = orc.OracleBLAS.from_verts(verts), hs.HsBlas(verts)
    assert hb.nodes2().tobytes() == ob.nodes.tobytes()
    assert hb.check_wide() == 0
    pushes = [(verts, None, [kat.I34], None)]
    o, h = engines.OracleEngine(pushes), engines.HostsimEngine(pushes)
    rays = W.interior_rays(20000, 21)
    a, b = h.trace(rays), o.trace(rays)
    cls = parity.classify(a, b, parity.make_graze_verifier(orc, rays, a, o.instances, o.tris))
    s = parity.assert_parity(cls, len(rays), label="bumpy sphere interior")
    assert b["hit"].all() and s["exact"] >= 0.995 * len(rays), s


def test_blob_structural_check_accepts_built_trees_and_catches_faults():
    """The validator run on imported blobs (rc_validate_blas_elem, k_validate_blas): clean for everything the builder produces
    (single triangle, 2-triangle root leaf, duplicates, large meshes), non-zero for each class of out-of-range reference."""
    for verts in (kat.TRI, W.quad_mesh(), W.box_mesh(), W.uv_sphere(9), W.bumpy_sphere(23), np.repeat(kat.TRI.reshape(1, 9), 9, axis=0)):
        hb = hs.HsBlas(verts)
        assert hb.validate() == 0
        for where in (0, 1, 7, 1000):
            for corrupt in (2, 3, 4, 5, 6, 8, 9):
                assert hb.validate(corrupt, where) >= 1, (hb.n, corrupt, where)
            if hb.n > 1:
                # 1: child index past the last node; 7 / 10: references that stay in range but would make a traversal cycle
                for corrupt in (1, 7, 10, 11, 12):
                    assert hb.validate(corrupt, where) >= 1, (hb.n, corrupt, where)
