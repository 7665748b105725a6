"""RC_MODE_WATERTIGHT: the reference's watertight triangle test (intersect_triangle, src/triangle_mesh.jl:168-201) as a selectable
mode of the traversal.  CPU part: the oracle's restatement against the reference's own KATs for that function and against the library's
per-element code (host simulation), bit for bit.  GPU part (-m gpu): the CUDA kernels against the oracle in this mode, and the property
the mode exists for — no ray leaks through the shared edges of a closed mesh."""
import numpy as np
import pytest

import engines
import parity
from oracle import oracle as orc
from raycore_b200 import workloads as W

F = np.float32


def test_oracle_watertight_kats():
    # test/test_intersection.jl "Test triangle": Triangle (0,0,2),(1,0,2),(1,1,2), ray from the origin along +z through (0.5? ...) -> the
    # reference's own checks for intersect_triangle are t, the barycentric point and a miss next to the triangle
    v0, v1, v2 = (0, 0, 2), (1, 0, 2), (1, 1, 2)
    hit, t, u, v = orc.intersect_triangle_watertight((0.75, 0.25, 0), (0, 0, 1), v0, v1, v2)
    assert hit and t == 2.0
    p = (1 - u - v) * np.array(v0, F) + u * np.array(v1, F) + v * np.array(v2, F)
    assert np.allclose(p, (0.75, 0.25, 2.0), atol=1e-6)
    assert not orc.intersect_triangle_watertight((0.25, 0.75, 0), (0, 0, 1), v0, v1, v2)[0]      # other half of the quad
    assert not orc.intersect_triangle_watertight((0.75, 0.25, 3), (0, 0, 1), v0, v1, v2)[0]      # behind the origin
    assert not orc.intersect_triangle_watertight((0.75, 0.25, 0), (0, 0, 1), v0, v1, v2, t_max=1.5)[0]
    assert orc.intersect_triangle_watertight((1.0, 0.5, 0), (0, 0, 1), v0, v1, v2)[0]            # on an edge: hits (edge functions >= 0)
    assert orc.intersect_triangle_watertight((1.0, 1.0, 0), (0, 0, 1), v0, v1, v2)[0]            # on a vertex
    # agrees with Moeller-Trumbore away from the edges (t to a few ulp, same hit / miss)
    rs = np.random.RandomState(3)
    n_hit = 0
    for _ in range(2000):
        tri = rs.uniform(-1, 1, (3, 3)).astype(F)
        o, d = rs.uniform(-2, 2, 3).astype(F), rs.normal(size=3).astype(F)
        a = orc.intersect_triangle_watertight(o, d, *tri)
        b = orc.intersect_triangle(o, d, *tri)
        if b[0] and min(b[2], b[3], 1 - b[2] - b[3]) > 1e-3:
            assert a[0] and abs(a[1] - b[1]) <= 1e-4 * max(1.0, abs(b[1])) and abs(a[2] - b[2]) < 1e-3 and abs(a[3] - b[3]) < 1e-3
            n_hit += 1
        elif not b[0] and not a[0]:
            pass
    assert n_hit > 20


def _scene():
    sphere = W.bumpy_sphere(40)
    xs = W.random_trs(6, 11, extent=3.0)
    return [(sphere, None, xs, None), (W.box_mesh(), None, W.random_trs(3, 5, extent=3.0), None)]


def test_hostsim_watertight_matches_oracle_bit_for_bit():
    """the library's RC_HD watertight code compiled for the CPU: the reference-order walk is bit-identical to the oracle's, the wide walk
    agrees outside ties"""
    pushes = _scene()
    o, hr, hw = engines.OracleEngine(pushes), engines.HostsimEngine(pushes, wide=False), engines.HostsimEngine(pushes, wide=True)
    rays = np.concatenate([W.box_rays(6000, 1, half=5.0), W.interior_rays(3000, 2, radius=4.0)])
    for any_hit in (False, True):
        a = o.trace(rays, any_hit=any_hit, watertight=True)
        assert hr.trace(rays, any_hit=any_hit, watertight=True).tobytes() == a.tobytes()
    a, b = hw.trace(rays, watertight=True), o.trace(rays, watertight=True)
    assert np.array_equal(a["hit"], b["hit"]) and 0.1 < a["hit"].mean() < 0.95
    same = a["primitive_id"] == b["primitive_id"]
    assert same.mean() > 0.999 and a[same].tobytes() == b[same].tobytes()


@pytest.mark.gpu
def test_gpu_watertight_parity_and_no_leaks():
    pushes = _scene()
    o, g, gr = engines.OracleEngine(pushes), engines.GpuEngine(pushes), engines.GpuEngine(pushes, reference_order=True)
    rays = np.concatenate([W.box_rays(200_000, 1, half=5.0), W.interior_rays(100_000, 2, radius=4.0)])
    for any_hit in (False, True):
        b = o.trace(rays, any_hit=any_hit, watertight=True)
        assert gr.trace(rays, any_hit=any_hit, watertight=True).tobytes() == b.tobytes()  # reference-order walk: bit-identical
        a = g.trace(rays, any_hit=any_hit, watertight=True)
        assert np.array_equal(a["hit"], b["hit"])
    a, b = g.trace(rays, watertight=True), o.trace(rays, watertight=True)
    same = a["primitive_id"] == b["primitive_id"]
    assert same.mean() > 0.9995 and a[same].tobytes() == b[same].tobytes()
    d = np.abs(a["t"][~same] - b["t"][~same])
    assert (d <= 1e-6 * np.maximum(1.0, b["t"][~same])).all()  # a different triangle only at a tie


@pytest.mark.gpu
def test_gpu_watertight_closed_mesh_has_no_leaks():
    """C2's closed 1 M-triangle mesh, 2^22 rays from interior points: Moeller-Trumbore (the default, = the reference's traversal) loses a
    handful through shared edges; the watertight mode must lose none.  Also reports what the mode costs."""
    import raycore_b200 as rc

    tl = rc.TLAS()
    tl.push(W.bumpy_sphere(709))
    tl.sync()
    n = 1 << 22
    rays = W.interior_rays(n, seed=77, radius=0.8)
    st = tl.adapt()
    mt = st.trace_closest(rays)
    ms_mt = tl.last_kernel_ms()
    wt = st.trace_closest(rays, watertight=True)
    ms_wt = tl.last_kernel_ms()
    leaks_mt, leaks_wt = int((mt["hit"] == 0).sum()), int((wt["hit"] == 0).sum())
    print(f"interior rays: Moeller-Trumbore leaks {leaks_mt}, watertight leaks {leaks_wt}; kernel ms {ms_mt:.3f} vs {ms_wt:.3f}")
    assert leaks_wt == 0 and leaks_mt <= 64
    agree = (mt["hit"] == 1) & (wt["primitive_id"] == mt["primitive_id"])
    assert agree.mean() > 0.9999
    assert np.abs(mt["t"][agree] - wt["t"][agree]).max() <= 2e-5 * np.abs(mt["t"][agree]).max()
    assert (st.trace_any(rays, watertight=True)["hit"] == 1).all()
    tl.free()
