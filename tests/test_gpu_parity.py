"""Parity tests proper: the CUDA library through the C ABI against the CPU oracle (bit-exact ids; DESIGN.md classes)."""
import numpy as np
import pytest

from oracle import oracle as orc
from raycore_b200 import workloads as W
import raycore_b200 as rc
import engines
import kat
import parity

pytestmark = pytest.mark.gpu


def _scene_instanced(n_inst=300, tess=16, seed=11):
    xf = W.random_trs(n_inst, seed, extent=10.0)
    return [(W.bumpy_sphere(tess), None, xf, np.arange(1, n_inst + 1, dtype=np.uint32)), (W.box_mesh(), None, W.random_trs(17, seed + 1, extent=10.0), None)]


def _rays(n, seed, half=12.0):
    return np.concatenate([W.box_rays(n // 2, seed, half=half), W.interior_rays(n - n // 2, seed + 5, radius=half * 0.9)])


def test_reference_kats_cuda_wide():
    kat.check_all(lambda p: engines.GpuEngine(p))


def test_reference_kats_cuda_reference_order():
    kat.check_all(lambda p: engines.GpuEngine(p, reference_order=True))


def test_builder_bit_exact_vs_oracle():
    for verts in (kat.TRI, W.quad_mesh(), W.box_mesh(), W.uv_sphere(9), W.bumpy_sphere(40), W.bumpy_sphere(150)):
        ob = orc.OracleBLAS.from_verts(verts)
        g = engines.GpuEngine([(verts, None, [kat.I34], None)])
        assert g.tlas.read_blas_order(1).tolist() == ob.prims["input_index"].tolist()
        assert g.tlas.read_blas_nodes(1).tobytes() == ob.nodes.tobytes(), "GPU BVH2 differs from the reference restatement"
        g.tlas.free()


def _builder_stress_meshes():
    """triangle soups that exercise the builder's block structure: sizes around the 256-face blocks / sorted runs of the small-build
    kernel (<= 32,768 faces) and its upper limit, around the 512-leaf fit blocks and the 2048-key sort tiles of the five-launch path
    above it, long runs of identical Morton codes (index tie-break chains: the deepest radix trees; equal keys across sorted runs), an
    exponentially spaced comb (one-sided tree), and soups with degenerate faces sprinkled in (compaction offsets)"""
    rs = np.random.RandomState(7)
    out = {}
    for n in (1, 2, 3, 5, 255, 256, 257, 511, 512, 513, 1023, 1024, 1025, 1537, 2047, 2048, 2049, 5000, 12800, 32767, 32768, 32769, 33792, 34815, 34816, 34817):
        c = rs.uniform(-10, 10, (n, 1, 3))
        out[f"soup{n}"] = (c + rs.uniform(-0.3, 0.3, (n, 3, 3))).reshape(n, 9).astype(np.float32)
    one = np.array([[0, 0, 0, 1, 0, 0, 0, 1, 0]], np.float32)
    out["duplicates"] = np.concatenate([np.repeat(one, 3000, axis=0), out["soup511"], np.repeat(one + 5, 700, axis=0)])
    k = np.arange(1400)
    x = (1.02 ** k)[:, None].astype(np.float32)
    out["comb"] = np.concatenate([x, 0 * x, 0 * x, x + 0.5 * x, 0 * x, 0 * x, x, 0.5 * x, 0 * x], axis=1).astype(np.float32)
    holes = out["soup5000"].copy()
    holes[rs.rand(5000) < 0.3, 3:] = np.tile(holes[rs.rand(5000) < 0.3][:1, :3], 2)  # v1 = v2 = one fixed point: zero-area faces
    out["holes"] = holes
    # the same two shapes on the five-launch path (> 32,768 faces)
    out["duplicates_large"] = np.concatenate([np.repeat(one, 22000, axis=0), out["soup12800"], np.repeat(one + 5, 2100, axis=0)])
    big = np.concatenate([out["soup33792"], out["soup5000"] + 30.0])
    mask = rs.rand(len(big)) < 0.3
    big[mask, 3:] = np.tile(big[mask][:1, :3], 2)
    out["holes_large"] = big
    return out


@pytest.mark.parametrize("name", sorted(_builder_stress_meshes()))
def test_builder_block_structure_bit_exact(name):
    """BVH2 byte-identical to the oracle's and every reachable wide node byte-identical to the host collapse of the same BVH2, on meshes
    chosen to hit the fit's block boundaries, spanning-node formula, segment tables and the cooperative front end's tile boundaries"""
    import raycore_b200 as rc
    import hostsim_py as hs

    verts = _builder_stress_meshes()[name]
    ob = orc.OracleBLAS.from_verts(verts)
    g = engines.GpuEngine([(verts, None, [kat.I34], None)])
    assert g.tlas.read_blas_order(1).tolist() == ob.prims["input_index"].tolist()
    assert g.tlas.read_blas_nodes(1).tobytes() == ob.nodes.tobytes(), "GPU BVH2 differs from the reference restatement"
    g.tlas.free()
    b4 = rc.build_blas4(verts)  # default build flags: no BVH2 emission
    wide = b4.nodes().view(np.uint8).reshape(-1, 64)
    host = hs.HsBlas(verts).nodes4()
    assert wide.shape == host.shape
    child = wide.view(np.uint32).reshape(-1, 16)
    seen, todo = set(), [1]
    while todo:
        k = todo.pop()
        if k in seen:
            continue
        seen.add(k)
        assert wide[k].tobytes() == host[k].tobytes(), f"wide node {k} differs"
        for c in child[k, [10, 11, 12, 13]]:
            if not (c & 0x80000000):
                todo.append(int(c))
    assert not wide[0].any() and (len(ob.nodes) == 1 or not wide[-1].any())  # the two slots nothing references are zeroed
    b4.free()


def test_tlas_bit_exact_vs_oracle():
    pushes = _scene_instanced()
    o, g = engines.OracleEngine(pushes), engines.GpuEngine(pushes)
    for k, h in enumerate(g.handles):
        gi = g.tlas.get_instances(h)
        oi = o.instances[o.instances["blas_index"] == k + 1]
        assert gi.tobytes() == oi.tobytes(), "instance descriptors (incl. mat3x4_inverse) differ"
    assert g.tlas.read_tlas_nodes().tobytes() == o.tlas.nodes.tobytes()
    assert np.array_equal(g.world_bound(), o.tlas.root_aabb)
    s = g.tlas.sizes()
    assert s["tlas_nodes"] == 2 * 317 - 1 and s["blas_prims"] == o.tlas.c.n_blas_prims and s["blas_nodes"] == o.tlas.c.n_blas_nodes


@pytest.mark.parametrize("n_inst", [1, 2, 255, 256, 257, 1025, 5000, 32768, 32769])
def test_tlas_block_structure_bit_exact(n_inst):
    """TLAS builds and refits around the block sizes of the one-kernel path (k_tlas_small: 256 instances per block, <= 32,768 instances) and
    on the multi-launch path just above it: BVH2 + root box byte-identical to the oracle's after the build AND after a refit with new
    transforms; a run of concentric instances (equal Morton codes across sorted runs: the index tie-break) in every scene; wide-TLAS
    traversal exact against the oracle."""
    rs = np.random.RandomState(n_inst)
    ext = max(4.0, 1.5 * n_inst ** (1 / 3))

    def transforms(seed):
        xf = W.random_trs(n_inst, seed, extent=ext)
        if n_inst >= 255:  # 200 instances around one centre: equal Morton codes (the box mesh is centred), nested boxes of distinct sizes (no exact t ties)
            m = xf[40].reshape(3, 4).copy()
            for k in range(200):
                mk = m.copy()
                mk[:, :3] *= np.float32(1.0 + 0.004 * k)
                xf[40 + k] = mk.reshape(12)
        return xf

    verts = W.box_mesh()
    xf0, xf1 = transforms(3), transforms(4)
    g = engines.GpuEngine([(verts, None, xf0, None)])
    o = engines.OracleEngine([(verts, None, xf0, None)])
    for step in (0, 1):
        if step == 1:  # refit_tlas! keeps the topology of the build (:2197-2222): the oracle re-fits its own TLAS to the new transforms
            import raycore_b200 as rc
            g.tlas.update_transforms(g.handles[0], list(xf1))
            g.tlas.sync()
            assert g.tlas.last_sync_action == rc.RC_SYNC_REFIT
            moved = engines.instances_of([(verts, None, xf1, None)], orc.mat3x4_inverse)
            o.instances[:] = moved
            o.tlas.instances[:] = moved.view(orc.INSTANCE_DTYPE)
            o.tlas.refit()
        assert g.tlas.read_tlas_nodes().tobytes() == o.tlas.nodes.tobytes(), f"TLAS BVH2 differs (step {step})"
        assert np.array_equal(g.world_bound(), o.tlas.root_aabb)
        rays = W.box_rays(20000, 11 + step, half=1.1 * ext)
        a, b = g.trace(rays), o.trace(rays)
        cls = parity.classify(a, b, parity.make_graze_verifier(orc, rays, a, o.instances, o.tris))
        parity.assert_parity(cls, len(rays), label=f"tlas {n_inst} step {step}")
    g.tlas.free()
    del rs


@pytest.mark.parametrize("any_hit", [False, True])
def test_reference_order_traversal_bit_exact(any_hit):
    pushes = _scene_instanced()
    o, g = engines.OracleEngine(pushes), engines.GpuEngine(pushes, reference_order=True)
    rays = _rays(200000, 3)
    a, b = g.trace(rays, any_hit=any_hit), o.trace(rays, any_hit=any_hit)
    assert a.tobytes() == b.tobytes()


def test_wide_traversal_parity_closest():
    pushes = _scene_instanced()
    o, g = engines.OracleEngine(pushes), engines.GpuEngine(pushes)
    rays = _rays(400000, 9)
    a, b = g.trace(rays), o.trace(rays)
    cls = parity.classify(a, b, parity.make_graze_verifier(orc, rays, a, o.instances, o.tris))
    s = parity.assert_parity(cls, len(rays), label="cuda wide closest")
    assert s["exact"] >= 0.995 * len(rays), s
    # t / barycentrics within 1e-5 relative wherever ids agree (they are bit-identical: same expression, same space)
    same = (a["hit"] == 1) & (a["primitive_id"] == b["primitive_id"]) & (a["instance_id"] == b["instance_id"]) & (b["hit"] == 1) & ~np.isnan(b["t"])
    assert np.all(np.abs(a["t"][same] - b["t"][same]) <= 1e-5 * np.abs(b["t"][same]))
    assert np.all(np.abs(a["bary_u"][same] - b["bary_u"][same]) <= 1e-5) and np.all(np.abs(a["bary_v"][same] - b["bary_v"][same]) <= 1e-5)


def test_wide_traversal_parity_any():
    pushes = _scene_instanced()
    o, g = engines.OracleEngine(pushes), engines.GpuEngine(pushes)
    rays = _rays(400000, 10)
    a, b = g.trace(rays, any_hit=True), o.trace(rays, any_hit=True)
    mism = np.nonzero(a["hit"] != b["hit"])[0]
    assert len(mism) <= 4, (len(mism), mism[:5])
    ver = parity.make_graze_verifier(orc, rays, a, o.instances, o.tris)
    idx = np.nonzero(a["hit"] == 1)[0][:4000]
    assert ver(idx).all(), "any_hit reported a triangle the exact test does not accept"


def test_bumpy_sphere_250k_interior_and_primary():
    verts = W.bumpy_sphere(355)  # ~250k triangles
    pushes = [(verts, None, [kat.I34], None)]
    o, g = engines.OracleEngine(pushes), engines.GpuEngine(pushes)
    assert g.tlas.read_blas_nodes(1).tobytes() == o.blas[0].nodes.tobytes()
    for rays, label in ((W.interior_rays(1 << 20, 21), "interior"), (W.pinhole_rays(1024, 1024), "primary")):
        a, b = g.trace(rays), o.trace(rays)
        cls = parity.classify(a, b, parity.make_graze_verifier(orc, rays, a, o.instances, o.tris))
        s = parity.assert_parity(cls, len(rays), label=label)
        assert s["exact"] >= 0.999 * len(rays), s
    # the reference-order path is bit-identical on the same rays
    gr = engines.GpuEngine(pushes, reference_order=True)
    rays = W.interior_rays(1 << 18, 22)
    assert gr.trace(rays).tobytes() == o.trace(rays).tobytes()


def test_readme_sphere_c1():
    # BASELINE config 0: README sphere, single-instance TLAS, 1024x1024 primary rays (reference CPU path = oracle)
    verts = W.uv_sphere(24, (0, 0, 2), 1.0)
    pushes = [(verts, None, [kat.I34], [1])]
    o, g = engines.OracleEngine(pushes), engines.GpuEngine(pushes)
    rays = W.pinhole_rays(1024, 1024, camera_pos=(0, 0, 0))
    a, b = g.trace(rays), o.trace(rays)
    cls = parity.classify(a, b, parity.make_graze_verifier(orc, rays, a, o.instances, o.tris))
    s = parity.assert_parity(cls, len(rays), label="README sphere", max_graze=0, max_nan=0)  # C1 of BASELINE.json: the measured zeros
    assert b["hit"].mean() > 0.1 and s["exact"] >= 0.999 * len(rays)


def test_device_resident_buffers_and_counters():
    import ctypes as C

    pushes = _scene_instanced(50, 10)
    g = engines.GpuEngine(pushes)
    rays = _rays(100000, 4)
    ref = g.trace(rays)
    L, ctx = g.tlas._lib, g.tlas._ctx
    d_r, d_h = C.c_void_p(), C.c_void_p()
    assert L.rc_device_alloc(ctx, rays.nbytes, C.byref(d_r)) == 0 and L.rc_device_alloc(ctx, ref.nbytes, C.byref(d_h)) == 0
    assert L.rc_memcpy_h2d(ctx, d_r, rays.ctypes.data, rays.nbytes) == 0
    assert L.rc_trace_closest(ctx, d_r, d_h, len(rays), rc._lib.RC_RAYS_ON_DEVICE | rc._lib.RC_HITS_ON_DEVICE | rc._lib.RC_COUNTERS) == 0
    out = np.zeros_like(ref)
    assert L.rc_memcpy_d2h(ctx, out.ctypes.data, d_h, out.nbytes) == 0
    assert out.tobytes() == ref.tobytes()
    c = g.tlas.counters()
    assert c["rays"] == len(rays) and c["nodes"] > len(rays) and c["tri_tests"] > 0 and 0 < c["max_stack"] < 64
    assert g.tlas.last_kernel_ms() > 0
    L.rc_device_free(ctx, d_r), L.rc_device_free(ctx, d_h)


def _deep_scene(K):
    g = np.array([2.0 ** -i for i in range(K)], np.float32)
    pts = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    s = (pts.min(1) * 0.05)[:, None]
    a = pts + s * np.array([1, -1, 0], np.float32)
    b = pts + s * np.array([-1, 1, 0], np.float32)
    c = pts + s * np.array([0, 0, 1], np.float32)
    return np.concatenate([a, b, c], 1).astype(np.float32)


def test_deep_trees_short_stack_fixup():
    """Exponentially nested geometry (BLAS and TLAS): traversal needs more than the 32-entry shared-memory stack of the fast
    kernel (the reference's own MVector{32} stack would overflow here, src/instanced-bvh.jl:1912); the flagged rays are
    re-traced by k_trace_fixup and must still agree with the oracle."""
    blas = _deep_scene(20)
    g = [2.0 ** -i for i in range(14)]
    xf = np.stack([W.trs3x4((4 * a, 4 * b, 4 * g[(i + j) % 14]), (1, 0, 0, 0), 1.0) for i, a in enumerate(g) for j, b in enumerate(g)])
    pushes = [(blas, None, xf, None)]
    o, gq = engines.OracleEngine(pushes), engines.GpuEngine(pushes)
    rs = np.random.RandomState(0)
    n = 4096
    org = np.full((n, 3), 1e-9, np.float32)
    d = (np.array([1, 1, 1], np.float32) + rs.uniform(-0.9, 0.9, (n, 3))).astype(np.float32)
    rays = W.make_rays(org, d)
    a = gq.tlas.adapt().trace_closest(rays, counters=True)
    c = gq.tlas.counters()
    assert c["max_stack"] > 32, c
    b = o.trace(rays)
    cls = parity.classify(a, b, parity.make_graze_verifier(orc, rays, a, o.instances, o.tris))
    parity.assert_parity(cls, n, max_tie_frac=0.05, label="deep trees")
    assert (a["hit"] <= 1).all()
    # same through the non-instrumented kernel and for any_hit
    a2 = gq.trace(rays)
    assert a2.tobytes() == a.tobytes()
    assert np.array_equal(gq.trace(rays, any_hit=True)["hit"], o.trace(rays, any_hit=True)["hit"])
    # the fast kernel lists the flagged rays for the fix-up pass; with a list that is too short (or none) the pass scans the hit records instead
    import os
    for cap in ("2", "0"):
        os.environ["RC_OVF_LIST_CAP"] = cap
        try:
            g2 = engines.GpuEngine(pushes)
        finally:
            del os.environ["RC_OVF_LIST_CAP"]
        assert g2.trace(rays).tobytes() == a.tobytes(), f"fix-up pass with RC_OVF_LIST_CAP={cap}"
        g2.tlas.free()


def test_edge_case_rays_and_geometry():
    """Edge cases the domain has: zero / negative-zero direction components, axis-parallel rays lying in box faces, t_min / t_max
    windows, non-finite rays (must terminate and miss, never hang or crash), empty batches, coincident duplicate triangles (exact
    ties: the winner is traversal-order dependent in the reference, so only t is compared), rays starting on a triangle."""
    verts = np.concatenate([W.box_mesh(), W.box_mesh(), W.quad_mesh(2.0, 3.0)])  # the box twice: every box hit is an exact tie
    pushes = [(verts, None, [kat.I34, W.translation3x4((4, 0, 0))], [5, 6])]
    o, g, gr = engines.OracleEngine(pushes), engines.GpuEngine(pushes), engines.GpuEngine(pushes, reference_order=True)
    org = np.array([[0.1, 0.2, 5], [0.1, 0.2, 5], [0.5, 0.5, 5], [0.0, 0.0, 0.0], [0.1, 0.2, 0.5], [10, 10, 10], [4.1, 0.1, -5], [0.1, 0.2, 5], [0.1, 0.2, 5]], np.float32)
    d = np.array([[0, 0, -1], [-0.0, -0.0, -1], [0, 0, -1], [1, 0, 0], [0, 0, 1], [0, 0, -1], [0, 0, 1], [0, 0, -1], [0, 0, -1]], np.float32)
    rays = W.make_rays(org, d)
    rays["t_min"][7], rays["t_max"][8] = 4.7, 2.0  # window beyond the first two surfaces / before the first
    a, r, b = g.trace(rays), gr.trace(rays), o.trace(rays)
    assert r.tobytes() == b.tobytes()
    assert np.array_equal(a["hit"], b["hit"])
    ok = b["hit"] == 1
    assert np.array_equal(a["t"][ok], b["t"][ok]) or np.allclose(a["t"][ok], b["t"][ok], rtol=1e-6)
    assert b["hit"][5] == 0 and b["hit"][0] == 1 and b["hit"][8] == 0
    # non-finite rays
    bad = W.make_rays([[np.nan, 0, 0], [0, 0, 5], [np.inf, 0, 0], [0, 0, 5], [0, 0, 5]], [[0, 0, -1], [np.nan, 0, -1], [0, 0, -1], [np.inf, 0, -1], [0, 0, 0]])
    for eng in (g, gr):
        h = eng.trace(bad)
        assert (h["hit"] <= 1).all() and h["hit"][:4].sum() == 0
        assert (eng.trace(bad, any_hit=True)["hit"] <= 1).all()
    # empty batch
    assert len(g.trace(rays[:0])) == 0
    # large counts of identical rays (all lanes retire together) and a batch that is not a multiple of anything
    many = np.repeat(rays[:1], 100003)
    h = g.trace(many)
    assert (h["hit"] == 1).all() and (h["t"] == h["t"][0]).all()


def test_blas4_api():
    """BLAS4 / build_blas4 / closest_hit4 / any_hit4 (src/bvh4.jl:511-766): one geometry, no instance index in
    the result, ray.t_min ignored (`ray_mint = 0`, :610)."""
    import raycore_b200 as rc

    verts = W.uv_sphere(32, (0, 0, 2), 1.0)
    blas = rc.build_blas4(verts)
    o = engines.OracleEngine([(verts, None, [kat.I34], None)])
    rays = W.pinhole_rays(128, 128, camera_pos=(0, 0, 0))
    rays["t_min"] = 1.5  # beyond the front surface for the central rays: closest_hit4 must still report the front hit
    a = blas.trace_closest4(rays)
    r0 = rays.copy()
    r0["t_min"] = 0
    b = o.trace(r0)
    cls = parity.classify(a, b, parity.make_graze_verifier(orc, r0, a, o.instances, o.tris))
    parity.assert_parity(cls, len(rays), label="closest_hit4")
    assert np.array_equal(blas.trace_any4(rays)["hit"], o.trace(r0, any_hit=True)["hit"])
    hit, tri, t, bary = rc.closest_hit4(blas, rc.Ray((0, 0, 0), (0, 0, 1), 1.5))
    assert hit and abs(float(t) - 1.0) < 1e-3 and len(bary) == 3 and abs(float(bary.sum()) - 1) < 1e-6
    assert tri.vertices.shape == (3, 3)
    hit, _, t, _ = rc.any_hit4(blas, rc.Ray((0, 0, 0), (0, 1, 0)))
    assert not hit and t == 0
    assert blas.n_primitives == len(o.tris[1])
    with pytest.raises(rc.RaycoreError):
        rc.build_blas4(np.zeros((0, 9), np.float32))
    blas.free()


@pytest.mark.parametrize("any_hit", [False, True])
def test_single_instance_shortcut_with_transform(any_hit):
    """A TLAS with exactly one instance skips the top-level traversal in the scheduler kernel (the refill step enters the
    instance directly): rotated / scaled / translated instance, rays that miss its world box, custom instance id."""
    verts = W.uv_sphere(48)
    xf = W.random_trs(1, seed=21, extent=3.0, smin=0.4, smax=2.5)
    pushes = [(verts, None, xf, np.array([77], np.uint32))]
    o, g = engines.OracleEngine(pushes), engines.GpuEngine(pushes)
    c = xf[0].reshape(3, 4)[:, 3]
    n = 60000
    rays = W.box_rays(n, seed=4, half=6.0)
    rays["o"] += c  # around the instance
    far = W.make_rays(np.full((100, 3), 50.0, np.float32) + c, np.tile(np.array([[0, 1, 0]], np.float32), (100, 1)))  # miss the world box
    rays = np.concatenate([rays, far])
    a, b = g.trace(rays, any_hit=any_hit), o.trace(rays, any_hit=any_hit)
    if any_hit:
        assert np.array_equal(a["hit"], b["hit"])
    else:
        cls = parity.classify(a, b, parity.make_graze_verifier(orc, rays, a, o.instances, o.tris))
        s = parity.assert_parity(cls, len(rays), label="single instance")
        assert s["exact"] > 0.99 * len(rays)
    assert 0.02 < a["hit"].mean() < 0.98 and a["hit"][-100:].sum() == 0
    assert (a["instance_custom_index"][a["hit"] == 1] == 77).all() and (a["instance_id"][a["hit"] == 1] == 0).all()
