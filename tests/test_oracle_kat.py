"""Pins the CPU oracle (oracle/oracle.c) against the known-answer values of the reference's own tests.
The reference cannot be executed here (no Julia), so these KATs are what anchors the restatement."""
import numpy as np
import pytest

from oracle import oracle as orc
from raycore_b200 import workloads as W
import engines
import kat


def test_scalar_kats():
    # test/test_instanced_bvh.jl:175-184
    assert orc.expand_bits(0) == 0
    assert orc.clz32(0) == 32 and orc.clz32(1) == 31 and orc.clz32(0x80000000) == 0
    # :20-37 Morton monotonic along the diagonal
    c1, c2, c3 = orc.morton_code_30bit((0, 0, 0)), orc.morton_code_30bit((1, 1, 1)), orc.morton_code_30bit((0.5, 0.5, 0.5))
    assert c1 < c3 < c2
    assert c1 == 0 and c2 == 0x3FFFFFFF  # 1023 on every axis
    # :186-200 delta out of bounds = -1
    codes = np.array([1, 2, 4, 8], np.uint32)
    assert orc.delta(1, 10, codes) == -1 and orc.delta(0, 1, codes) == -1
    assert orc.delta(1, 2, codes) == 30 and orc.delta(2, 3, codes) == 29
    # equal codes fall back to the index tiebreak: 32 + clz(i xor j) (src/instanced-bvh.jl:1227)
    assert orc.delta(1, 2, np.array([5, 5], np.uint32)) == 32 + 30
    # NaN normalised coordinate -> 0 (x86 unsafe_trunc)
    assert orc.morton_code_30bit((np.nan, 0.0, 0.0)) == 0


def test_transform_kats():
    # test/test_instanced_bvh.jl:122-147
    ident = orc.mat4_to_mat3x4(np.eye(4))
    assert np.allclose(orc.transform_point(ident, (1, 2, 3)), (1, 2, 3))
    T = np.eye(4, dtype=np.float32)
    T[:3, 3] = (5, 10, 15)  # Julia Mat4f(1,0,0,0, 0,1,0,0, 0,0,1,0, 5,10,15,1) is column-major: translation in column 4
    t34 = orc.mat4_to_mat3x4(T)
    assert np.allclose(orc.transform_point(t34, (1, 2, 3)), (6, 12, 18))
    assert np.allclose(orc.transform_direction(t34, (1, 0, 0)), (1, 0, 0))
    inv = orc.mat3x4_inverse(t34)
    assert np.allclose(orc.transform_point(inv, (6, 12, 18)), (1, 2, 3))
    # general affine: inverse really inverts
    m = W.trs3x4((1, -2, 3), (0.5, 0.5, 0.5, 0.5), 1.7)
    p = np.array([0.3, -0.2, 0.9], np.float32)
    assert np.allclose(orc.transform_point(orc.mat3x4_inverse(m), orc.transform_point(m, p)), p, atol=1e-5)


def test_primitive_kats():
    # safe_invdir (src/instanced-bvh.jl:1742-1748)
    assert np.allclose(orc.safe_invdir((2, 0, -4)), (0.5, 1e5, -0.25))
    assert orc.safe_invdir((-0.0, 1, 1))[0] == np.float32(-1e5)  # copysign keeps the sign; check_direction removes -0 beforehand
    # slab test (test/test_intersection.jl:1-20 analogue): unit box from outside, inside, miss
    inv = orc.safe_invdir((0, 0, 1))
    tmin, tmax = orc.intersect_bbox((0.5, 0.5, -2), inv, (0, 0, 0), (1, 1, 1))
    assert tmin == 2 and tmax == 3
    tmin, tmax = orc.intersect_bbox((2.5, 0.5, -2), inv, (0, 0, 0), (1, 1, 1))
    assert tmin > tmax
    # Moeller-Trumbore: t ≈ 4 style KAT (test/test_intersection.jl:22-55 uses the watertight test; same geometry)
    hit, t, u, v = orc.intersect_triangle((0, 0, -2), (0, 0, 1), (0, 0, 2), (1, 0, 2), (0, 1, 2))
    assert hit and t == 4 and u == 0 and v == 0  # bary = (1,0,0)
    # degenerate filter: exact zero area only (src/triangle_mesh.jl:14-17)
    assert orc.is_degenerate((0, 0, 0, 1, 1, 1, 2, 2, 2))
    assert not orc.is_degenerate((0, 0, 0, 1, 0, 0, 0, 1e-6, 0))
    # coplanar ray -> NaN accepted
    hit, t, u, v = orc.intersect_triangle((0.5, 0.5, 5), (0, 0, -1), (0.5, 0, 0), (0.5, 1, 0), (0.5, 0, 1))
    assert hit and np.isnan(t)


def test_blas_structure_kats():
    # single triangle: 1 leaf node with child1 == 1 (test/test_instanced_bvh.jl:39-58)
    b = orc.OracleBLAS.from_verts(kat.TRI)
    assert len(b.nodes) == 1 and b.nodes[0]["child0"] == orc.INVALID_NODE and b.nodes[0]["child1"] == 1
    # two triangles: 3 nodes, interior root, root AABB (:60-99)
    b = orc.OracleBLAS.from_verts(W.quad_mesh(0.0, 0.5) + 0.5 * np.array([1, 1, 0] * 3, np.float32))
    assert len(b.nodes) == 3 and b.nodes[0]["child0"] != orc.INVALID_NODE
    assert np.allclose(b.root_aabb, (0, 0, 0, 1, 1, 0))
    # every internal node's boxes contain its children; parents consistent; leaves hold prims 1..n
    b = orc.OracleBLAS.from_verts(W.bumpy_sphere(12))
    n = b.n
    nodes = b.nodes
    assert sorted(nodes["child1"][n - 1:]) == list(range(1, n + 1))
    assert (np.diff(b.morton.astype(np.int64)) >= 0).all()
    for i in range(n - 1):
        for c in (nodes["child0"][i], nodes["child1"][i]):
            assert nodes["parent"][c - 1] == i + 1


def test_tlas_structure_kats():
    # node counts 1 / 3 / 161 (test/test_instanced_bvh.jl:226,262,803)
    e = engines.OracleEngine([(kat.TRI, None, [kat.I34], [1])])
    assert len(e.tlas.nodes) == 1 and e.tlas.nodes[0]["child1"] == 0
    e = engines.OracleEngine([(kat.TRI, None, [kat.I34, W.translation3x4((5, 0, 0))], [1, 2])])
    assert len(e.tlas.nodes) == 3
    xf = [W.translation3x4((((i - 1) % 9) * 1.5, ((i - 1) // 9) * 1.25, 0)) for i in range(1, 82)]
    e = engines.OracleEngine([(kat.TRI, None, xf, None)])
    assert len(e.tlas.nodes) == 161 and e.tlas.n_instances == 81
    # empty TLAS traces must miss, not crash (test/test_tlas_stress.jl:808-831)
    t = orc.OracleTLAS([], np.zeros(0, orc.INSTANCE_DTYPE))
    assert t.closest_hit(W.make_rays([(0, 0, 1)], (0, 0, -1)))["hit"][0] == 0


def test_reference_query_kats():
    kat.check_all(engines.OracleEngine)


def test_refit_matches_rebuild_boxes():
    # refit keeps topology; boxes equal a fresh build's boxes when the Morton order does not change
    xf = np.stack([W.translation3x4((3.0 * k, 0, 0)) for k in range(9)])
    e = engines.OracleEngine([(W.uv_sphere(6), None, xf, None)])
    xf2 = xf.copy()
    xf2[:, 7] += 0.25  # shift everything in y: same order
    inst = e.tlas.instances
    inst["transform"][:] = xf2
    inst["inv_transform"][:] = np.stack([orc.mat3x4_inverse(t) for t in xf2])
    e.tlas.refit()
    e2 = engines.OracleEngine([(W.uv_sphere(6), None, xf2, None)])
    assert np.array_equal(e.tlas.nodes.tobytes(), e2.tlas.nodes.tobytes())
    assert np.array_equal(e.tlas.root_aabb, e2.tlas.root_aabb)


def test_analysis_oracle_sanity():
    e = engines.OracleEngine([(W.uv_sphere(24, (0, 0, 2), 1.0), None, [kat.I34], None)])
    # grid rays along +z see the sphere: centroid ≈ (0,0,~1.3), illumination counts sum == number of hits
    hits, pts = e.tlas.hits_from_grid((0, 0, 1), 32)
    n, c = e.tlas.get_centroid((0, 0, 1), 32)
    assert n == int(hits["hit"].sum()) and n > 400
    assert abs(c[0]) < 0.05 and abs(c[1]) < 0.05 and 1.0 < c[2] < 2.0
    ill = e.tlas.get_illumination((0, 0, 1), 32)
    # default metadata = face index before filtering, which exceeds n_prims for some faces: those are dropped as in 1:length(prims) (:123)
    assert ill.sum() <= n
    # view factors between two facing quads: only cross terms, rows bounded by rays_per_triangle
    q0 = W.quad_mesh(0.0)[:, [0, 1, 2, 3, 4, 5, 6, 7, 8]]
    q1 = W.quad_mesh(1.0).reshape(-1, 3, 3)[:, ::-1, :].reshape(-1, 9)  # flipped winding: faces -z, towards q0
    e = engines.OracleEngine([(q0, np.array([1, 2], np.uint32), [kat.I34], None), (q1, np.array([3, 4], np.uint32), [kat.I34], None)])
    vf = e.tlas.view_factors(200, seed=3)
    assert vf.shape == (4, 4) and (np.diag(vf) == 0).all() and (vf.sum(1) <= 200).all()
    assert vf[:2, :2].sum() == 0 and vf[2:, 2:].sum() == 0 and vf[:2, 2:].sum() > 50 and vf[2:, :2].sum() > 50
    # RNG KAT: uniform in [0,1), reproducible, and equal to the numpy statement of the same spec
    u = [orc.rng_uniform(7, i, d) for i in range(50) for d in range(4)]
    assert min(u) >= 0 and max(u) < 1 and len(set(u)) > 190
    un = np.array([W.rng_uniform(7, np.arange(50, dtype=np.uint64), d) for d in range(4)]).T.reshape(-1)
    assert np.array_equal(np.array(u, np.float32), un)


def test_workloads_degenerate_filter_matches_oracle():
    """raycore_b200.workloads.is_degenerate (numpy, used by bench.py to number view-factor metadata) == the oracle's restatement of
    is_degenerate (src/triangle_mesh.jl:14-17) on pole faces, exact duplicates, collinear and tiny-but-valid triangles."""
    from raycore_b200 import workloads as W

    soup = np.concatenate([W.uv_sphere(24), W.bumpy_sphere(40, (1, 2, 3), 0.5), W.box_mesh()])
    rng = np.random.default_rng(0)
    extra = rng.normal(size=(200, 9)).astype(np.float32)
    extra[:50, 3:6] = extra[:50, 0:3]  # repeated vertex
    extra[50:100, 6:9] = extra[50:100, 0:3] + 2 * (extra[50:100, 3:6] - extra[50:100, 0:3])  # collinear (up to rounding)
    extra[100:150] *= 1e-18  # cross product underflows towards zero
    soup = np.concatenate([soup, extra])
    want = np.array([bool(orc.is_degenerate(v)) for v in soup])
    got = W.is_degenerate(soup)
    assert np.array_equal(got, want)
    assert 0 < got.sum() < len(soup)
