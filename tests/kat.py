"""Known-answer tests taken from the reference's own test-suite for the ray-query path, written once
and run against every engine (tests/engines.py).  Each check cites the reference test it restates."""
import numpy as np

from raycore_b200 import workloads as W

I34 = W.identity3x4()
TRI = np.array([[0, 0, 0, 1, 0, 0, 0, 1, 0]], np.float32)  # unit triangle in the XY plane (test_instanced_bvh.jl:276)
DOWN = (0, 0, -1)


def tri_at(off):
    v = TRI.copy().reshape(3, 3) + np.asarray(off, np.float32)
    return v.reshape(1, 9)


def check_all(make_engine):
    """make_engine(pushes) -> engine with .trace(rays, any_hit) and .world_bound()."""
    approx = lambda a, b, atol=1e-6: abs(float(a) - float(b)) <= atol + 1.5e-4 * abs(float(b))  # Julia isapprox default rtol = sqrt(eps(Float32))

    # -- "TLAS closest_hit - Basic" (test/test_instanced_bvh.jl:274-302): dist ≈ 1, metadata == 42, miss at (2,2)
    e = make_engine([(TRI, np.array([42], np.uint32), [I34], [1])])
    h = e.trace(W.make_rays([(0.25, 0.25, 1.0), (2, 2, 1.0)], DOWN))
    assert h["hit"][0] == 1 and approx(h["t"][0], 1.0) and h["meta"][0] == 42 and h["instance_id"][0] == 0
    assert h["instance_custom_index"][0] == 1
    assert h["hit"][1] == 0
    # miss returns the zero sentinel (test/test_intersection.jl:121-142; src/instanced-bvh.jl:2019-2022)
    assert h["t"][1] == 0 and h["bary_u"][1] == 0 and h["bary_v"][1] == 0 and h["meta"][1] == 0 and h["primitive_id"][1] == 0

    # -- "Transformed Instance" (:304-339): translated by (10,0,0): old position misses, new position hits at 1
    e = make_engine([(TRI, None, [W.translation3x4((10, 0, 0))], [1])])
    h = e.trace(W.make_rays([(0.25, 0.25, 1.0), (10.25, 0.25, 1.0)], DOWN))
    assert h["hit"][0] == 0 and h["hit"][1] == 1 and approx(h["t"][1], 1.0)

    # -- "Multiple Instances (Closest Selection)" (:341-378): nearest of two -> inst_id == 1 (1-based)
    e = make_engine([(TRI, None, [I34, W.translation3x4((0, 0, -5))], [1, 2])])
    h = e.trace(W.make_rays([(0.25, 0.25, 1.0)], DOWN))
    assert h["hit"][0] == 1 and approx(h["t"][0], 1.0) and h["instance_id"][0] + 1 == 1
    wb = e.world_bound()
    assert approx(wb[2], -5.0) and approx(wb[5], 0.0)

    # -- "TLAS any_hit - Basic" (:380-405)
    e = make_engine([(TRI, None, [I34], [1])])
    h = e.trace(W.make_rays([(0.25, 0.25, 1.0), (2, 2, 1.0)], DOWN), any_hit=True)
    assert list(h["hit"]) == [1, 0]

    # -- "TLAS Construction - Multiple Instances" (:229-268): world bound x in [0, 6]
    e = make_engine([(TRI, None, [I34, W.translation3x4((5, 0, 0))], [1, 2])])
    wb = e.world_bound()
    assert approx(wb[0], 0.0) and approx(wb[3], 6.0)

    # -- "closest_hit_kernel! - basic intersection" (:807-840): masks [T,T,F,F], t ≈ 1
    e = make_engine([(TRI, None, [I34], None)])
    h = e.trace(W.make_rays([(0.25, 0.25, 1.0), (0.5, 0.25, 1.0), (5, 5, 1.0), (-1, -1, 1.0)], DOWN))
    assert list(h["hit"]) == [1, 1, 0, 0] and approx(h["t"][0], 1.0) and approx(h["t"][1], 1.0)
    # -- "any_hit_kernel!" (:842-870): [T,T,F,F]
    h = e.trace(W.make_rays([(0.25, 0.25, 1.0), (0.1, 0.1, 1.0), (5, 5, 1.0), (0.9, 0.9, 1.0)], DOWN), any_hit=True)
    assert list(h["hit"]) == [1, 1, 0, 0]

    # -- "instance identification" (:872-916): three instances of one mesh -> 1-based positions 1,2,3
    e = make_engine([(TRI, None, [I34, W.translation3x4((5, 0, 0)), W.translation3x4((0, 5, 0))], None)])
    h = e.trace(W.make_rays([(0.25, 0.25, 1.0), (5.25, 0.25, 1.0), (0.25, 5.25, 1.0)], DOWN))
    assert list(h["hit"]) == [1, 1, 1] and list(h["instance_id"] + 1) == [1, 2, 3]

    # -- "primitive metadata" (:918-952): three meshes, 4 rays -> [T,T,T,F]
    e = make_engine([(tri_at((0, 0, 0)), None, [I34], None), (tri_at((5, 0, 0)), None, [I34], None), (tri_at((0, 5, 0)), None, [I34], None)])
    h = e.trace(W.make_rays([(0.25, 0.25, 1.0), (5.25, 0.25, 1.0), (0.25, 5.25, 1.0), (10, 10, 1.0)], DOWN))
    assert list(h["hit"]) == [1, 1, 1, 0]

    # -- "barycentric coordinates" (:954-992): bary = (w,u,v); the edge hit at (0.5, 0) must hit
    e = make_engine([(TRI, None, [I34], None)])
    h = e.trace(W.make_rays([(0.25, 0.25, 1.0), (0.1, 0.1, 1.0), (0.5, 0.0, 1.0)], DOWN))
    assert list(h["hit"]) == [1, 1, 1]
    w = 1.0 - h["bary_u"] - h["bary_v"]
    assert abs(w[0] - 0.5) < 0.01 and abs(h["bary_u"][0] - 0.25) < 0.01
    assert abs(w[1] - 0.8) < 0.01 and abs(h["bary_u"][1] - 0.1) < 0.01
    assert abs(w[2] - 0.5) < 0.01 and abs(h["bary_u"][2] - 0.5) < 0.01

    # -- "full_trace_kernel!" (:994-1042): distances 2 and 3, instance positions 1 and 2
    e = make_engine([(tri_at((0, 0, 0)), None, [I34], None), (tri_at((5, 0, 0)), None, [I34], None)])
    h = e.trace(W.make_rays([(0.25, 0.25, 2.0), (5.25, 0.25, 3.0), (10, 10, 1.0)], DOWN))
    assert list(h["hit"]) == [1, 1, 0] and approx(h["t"][0], 2.0) and approx(h["t"][1], 3.0)
    assert list(h["instance_id"][:2] + 1) == [1, 2]
    assert abs((1 - h["bary_u"][0] - h["bary_v"][0]) - 0.5) < 0.01

    # -- Mesh Update (test/test_mesh_update.jl:56,96-116): unit sphere at z, ray from (0,0,5) down: t ≈ 4 - z (atol 0.1)
    for tess, z in ((32, 0.0), (8, 0.5), (48, 1.0), (12, -0.5)):
        e = make_engine([(W.uv_sphere(tess, (0, 0, z), 1.0), None, [I34], None)])
        h = e.trace(W.make_rays([(0.01, 0.02, 5.0)], DOWN))
        assert h["hit"][0] == 1 and abs(h["t"][0] - (4 - z)) < 0.1, (tess, z, h["t"][0])

    # -- TLAS stress: 200 BLASes in a row, hit iff present (test/test_tlas_stress.jl:187-227, scaled to 24)
    pushes = [(W.uv_sphere(6, (3.0 * k, 0, 0), 1.0), None, [I34], None) for k in range(24)]
    e = make_engine(pushes)
    h = e.trace(W.make_rays([(3.0 * k + 0.01, 0.02, 5.0) for k in range(24)] + [(1.5, 0, 5.0)], DOWN))
    assert list(h["hit"][:24]) == [1] * 24 and h["hit"][24] == 0
    assert list(h["instance_id"][:24]) == list(range(24))

    # -- any_hit ignores ray.t_min, closest_hit honours it (src/instanced-bvh.jl:1907 vs :2039)
    e = make_engine([(TRI, None, [I34], None)])
    r = W.make_rays([(0.25, 0.25, 1.0)], DOWN, t_min=2.0)
    assert e.trace(r)["hit"][0] == 0 and e.trace(r, any_hit=True)["hit"][0] == 1
    # t_max clips
    r = W.make_rays([(0.25, 0.25, 1.0)], DOWN, t_max=0.5)
    assert e.trace(r)["hit"][0] == 0 and e.trace(r, any_hit=True)["hit"][0] == 0
    # two-sided (no back-face culling, :1775-1792)
    assert e.trace(W.make_rays([(0.25, 0.25, -1.0)], (0, 0, 1)))["hit"][0] == 1
