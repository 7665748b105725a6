"""get_centroid / get_illumination / view_factors (src/kernels.jl) on the GPU against the oracle."""
import numpy as np
import pytest

from oracle import oracle as orc
from raycore_b200 import workloads as W
import raycore_b200 as rc
import engines
import kat
import parity

pytestmark = pytest.mark.gpu


def _dense_meta_scene():
    meshes = [W.bumpy_sphere(14, (0, 0, 0), 1.0), W.bumpy_sphere(12, (2.6, 0, 0), 0.8), W.quad_mesh(-1.5, 3.0)]
    pushes, base = [], 0
    for m in meshes:
        keep = np.array([not orc.is_degenerate(v) for v in m])
        meta = np.zeros(len(m), np.uint32)
        meta[keep] = base + 1 + np.arange(keep.sum())
        base += int(keep.sum())
        pushes.append((m, meta, [kat.I34], None))
    return pushes, base


def test_hits_from_grid_centroid_illumination():
    pushes, n_prims = _dense_meta_scene()
    o, g = engines.OracleEngine(pushes), engines.GpuEngine(pushes)
    for viewdir in ((0, 0, 1), (1.0, 0.5, -0.25), (0.95, 0.1, 0.0)):
        oh, op = o.tlas.hits_from_grid(viewdir, 64)
        gh, gp = g.tlas.hits_from_grid(viewdir, 64)
        same = oh.tobytes() == gh.tobytes()
        if not same:  # near-tie classes only
            diff = np.nonzero((oh["hit"] != gh["hit"]) | (oh["meta"] != gh["meta"]))[0]
            assert len(diff) <= 2, diff
        ok = (oh["meta"] == gh["meta"]) & (oh["hit"] == gh["hit"])
        assert np.array_equal(op[ok], gp[ok]), "hit points differ (sum_mul(bary, vertices))"
        ill_o, ill_g = o.tlas.get_illumination(viewdir, 64), g.tlas.get_illumination(viewdir, 64)
        assert ill_g.shape == (n_prims,) and np.abs(ill_o - ill_g).sum() <= 2
        n_o, c_o = o.tlas.get_centroid(viewdir, 48)
        pts, c_g = g.tlas.get_centroid(viewdir, 48)
        assert abs(len(pts) - n_o) <= 1 and np.allclose(c_o, c_g, rtol=1e-5, atol=1e-5)


def test_view_factors_same_rays_exact_and_statistics():
    pushes, n_prims = _dense_meta_scene()
    o, g = engines.OracleEngine(pushes), engines.GpuEngine(pushes)
    assert np.array_equal(np.sort(g.tlas.flat_metadata()), np.arange(1, n_prims + 1))
    rpt = 64
    vf_g = g.tlas.view_factors(rpt, seed=5)
    assert vf_g.shape == (n_prims, n_prims) and g.tlas.last_vf_skipped == 0
    # structural properties (kernels.jl:93-99): no self hits, at most rpt rays leave a triangle
    assert (np.diag(vf_g) == 0).all() and (vf_g.sum(1) <= rpt).all()
    # exactness on identical rays: trace the GPU's own generated rays with the oracle
    rays = g.tlas.view_factor_rays(rpt, seed=5)
    vf_o = o.tlas.view_factors_from_rays(rays, rpt)
    assert np.abs(vf_o.astype(np.int64) - vf_g.astype(np.int64)).sum() <= 4, "view-factor counts differ on identical rays"
    # statistical agreement with the oracle's own generation (same RNG stream, different libm): totals within 4 sigma
    vf_o2 = o.tlas.view_factors(rpt, seed=5)
    tot_g, tot_o = vf_g.sum(), vf_o2.sum()
    p = tot_o / (n_prims * rpt)
    assert abs(int(tot_g) - int(tot_o)) <= 4 * np.sqrt(n_prims * rpt * p * (1 - p)) + 8
    # row blocks (multi-GPU sharding unit) reproduce the full matrix
    a = g.tlas.view_factors(rpt, seed=5, row_base=0, n_rows=n_prims // 2)
    b = g.tlas.view_factors(rpt, seed=5, row_base=n_prims // 2)
    assert np.array_equal(np.vstack([a, b]), vf_g)
    # interleaved shares (row r of rank k in a world of 3 = row k + 3 r)
    for k in range(3):
        assert np.array_equal(g.tlas.view_factors(rpt, seed=5, row_base=k, row_stride=3), vf_g[k::3])
    with pytest.raises(rc.RaycoreError):
        g.tlas.view_factors(rpt, seed=5, row_base=1, row_stride=3, n_rows=(n_prims + 2) // 3 + 1)
    # uniform stream itself is identical on both sides
    assert np.array_equal(rays["o"].shape, (n_prims * rpt, 3))


def test_view_factors_deep_trees_fixup():
    """view_factors! on exponentially nested geometry: some of the generated rays need more than the 32-entry short stack; the
    scheduler kernel flags them in a bitmap and k_view_factor_fixup redoes them with the deep-stack body.  The matrix must
    still equal the oracle's on identical rays."""
    from test_gpu_parity import _deep_scene

    blas = _deep_scene(20)
    g14 = [2.0 ** -i for i in range(14)]
    xf = np.stack([W.trs3x4((4 * a, 4 * b, 4 * g14[(i + j) % 14]), (1, 0, 0, 0), 1.0) for i, a in enumerate(g14) for j, b in enumerate(g14)])
    pushes = [(blas, None, xf, None)]
    o, g = engines.OracleEngine(pushes), engines.GpuEngine(pushes)
    rpt = 2
    rays = g.tlas.view_factor_rays(rpt, seed=3)
    hits = g.tlas.adapt().trace_closest(rays, counters=True)
    deep = g.tlas.counters()["max_stack"]
    assert deep > 32, f"scene no longer exercises the fix-up pass (max stack {deep})"
    vf_g = g.tlas.view_factors(rpt, seed=3)
    # the nested copies produce many exact ties, so the matrix is compared with the one accumulated from the library's own
    # closest hits of the same rays (same tie resolution), and those hits with the oracle under the usual parity classes
    n = vf_g.shape[0]
    src = np.arange(len(rays)) // rpt + 1
    ok = (hits["hit"] == 1) & (hits["meta"] != src) & (hits["meta"] >= 1) & (hits["meta"] <= n)
    want = np.zeros_like(vf_g)
    np.add.at(want, (src[ok] - 1, hits["meta"][ok] - 1), 1)
    assert np.array_equal(vf_g, want)
    assert vf_g.sum() > 0
    b = o.trace(rays)
    cls = parity.classify(hits, b, parity.make_graze_verifier(orc, rays, hits, o.instances, o.tris))
    parity.assert_parity(cls, len(rays), max_tie_frac=0.05, label="deep view-factor rays")


def test_view_factors_metadata_not_a_permutation():
    """Duplicated and out-of-range metadata (the reference indexes result[src_meta, hit_meta] unchecked, src/kernels.jl:85,95-97):
    rows fed by two triangles fall back to the scan over all primitives (no row map), out-of-range sources are skipped and
    counted, row blocks still tile the full matrix."""
    near = W.quad_mesh(z=0.0, half=1.0)  # normal +z
    far = W.quad_mesh(z=1.0, half=1.0)[:, [6, 7, 8, 3, 4, 5, 0, 1, 2]]  # flipped: normal -z, faces the first quad
    verts = np.concatenate([near, far])
    meta = np.array([1, 1, 2, 9], np.uint32)
    tl = rc.TLAS()
    tl.push(verts, None, face_meta=meta)
    tl.sync()
    rpt = 200
    vf = tl.view_factors(rpt, seed=1)
    assert vf.shape == (4, 4) and tl.last_vf_skipped == 1
    assert vf[2:].sum() == 0 and vf[:, 2:].sum() == 0 and vf[0, 0] == 0 and vf[1, 1] == 0
    assert 0 < vf[0, 1] <= 2 * rpt and 0 < vf[1, 0] <= rpt  # row 0 is fed by two triangles
    blocks = np.vstack([tl.view_factors(rpt, seed=1, row_base=0, n_rows=1), tl.view_factors(rpt, seed=1, row_base=1, n_rows=3)])
    assert np.array_equal(blocks, vf)
    assert np.array_equal(tl.view_factors(rpt, seed=1, row_base=0, row_stride=2), vf[0::2])  # strided rows through the scan fallback
    assert np.array_equal(tl.view_factors(rpt, seed=1, row_base=1, row_stride=2), vf[1::2])
    with pytest.raises(rc.RaycoreError):
        tl.view_factors(rpt, row_base=3, n_rows=2)
    # the same geometry with permutation metadata takes the row-map path; out-of-range only
    tl2 = rc.TLAS()
    tl2.push(verts, None, face_meta=np.array([2, 1, 3, 9], np.uint32))
    tl2.sync()
    vf2 = tl2.view_factors(rpt, seed=1)
    assert tl2.last_vf_skipped == 1 and vf2[3].sum() == 0 and vf2[:, 3].sum() == 0
    assert vf2[:2, 2].sum() > 0 and vf2[2, :2].sum() > 0 and vf2[:2, :2].sum() == 0
    assert np.array_equal(np.vstack([tl2.view_factors(rpt, seed=1, row_base=0, n_rows=2), tl2.view_factors(rpt, seed=1, row_base=2, n_rows=2)]), vf2)
    tl.free(); tl2.free()
