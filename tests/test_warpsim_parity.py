"""The shipped traversal kernel on the CPU.  tests/hostsim compiles `k_trace_wide` (csrc/rc_trace_fast.cuh) itself — scheduler, vote
words, shared-memory stack discipline, level changes, refill — with g++ and runs it one fibre per lane with the warp intrinsics as
lock-step exchanges (tests/hostsim/warpsim.h), so the kernel's control logic is parity-tested against the oracle without a GPU.
(The `-m gpu` tests repeat these comparisons on the real hardware through the C ABI.)"""
import numpy as np
import pytest

from oracle import oracle as orc
from raycore_b200 import workloads as W
import engines
import kat
import parity
from test_hostsim_parity import _rays_for, _scene_instanced


def _parity(eng, o, rays, label, **kw):
    a, b = eng.trace(rays), o.trace(rays)
    cls = parity.classify(a, b, parity.make_graze_verifier(orc, rays, a, o.instances, o.tris))
    return a, b, parity.assert_parity(cls, len(rays), label=label, **kw)


def test_reference_kats_on_the_shipped_kernel():
    kat.check_all(lambda p: engines.WarpsimEngine(p))


def test_instanced_scene_closest_and_any():
    pushes = _scene_instanced()
    o, w = engines.OracleEngine(pushes), engines.WarpsimEngine(pushes)
    rays = _rays_for(pushes, 20000, 9)
    a, b, s = _parity(w, o, rays, "warpsim instanced closest")
    assert s["exact"] >= 0.995 * len(rays) and 0.2 < b["hit"].mean() < 1.0, s
    # the kernel's own work counters are consistent with the batch
    _, c = w.scene.trace_warpsim(rays, counters=True)
    assert c["rays"] == len(rays) and c["inst_entries"] > 0 and c["box_tests"] == 4 * c["nodes"] and c["short_stack_overflows"] == 0
    # ... and with the simulator's step statistics: every node / triangle / instance entry is one active lane of one N / T / X step
    it, ln = c["step_iterations"], c["step_lanes"]
    assert ln["N"] == c["nodes"] and ln["T"] == c["tri_tests"] and ln["X"] >= c["inst_entries"] > 0 and ln["F"] >= len(rays)
    assert all(0 < it[k] <= ln[k] <= 32 * it[k] for k in "NTXF")
    assert ln["N"] / it["N"] > 16, "node steps run at less than half a warp: the scheduler policy regressed"
    # any_hit: the hit flag is order-independent, the reported triangle must be a genuine exact hit
    ah, bh = w.trace(rays, any_hit=True), o.trace(rays, any_hit=True)
    assert (ah["hit"] != bh["hit"]).sum() <= 1
    ver = parity.make_graze_verifier(orc, rays, ah, o.instances, o.tris)
    assert ver(np.nonzero(ah["hit"] == 1)[0][:2000]).all()
    assert np.array_equal(ah["hit"], a["hit"])


@pytest.mark.parametrize("xf", [None, "trs"])
def test_single_instance_variant(xf):
    """One instance => the launcher's SINGLE compile-time variant (no TLAS walk, no sentinel, no level change)."""
    verts = W.bumpy_sphere(48)
    t = [kat.I34] if xf is None else list(W.random_trs(1, 5, extent=2.0))
    pushes = [(verts, None, t, [9])]
    o, w = engines.OracleEngine(pushes), engines.WarpsimEngine(pushes)
    rays = np.concatenate([W.interior_rays(8000, 21, radius=0.6), W.box_rays(4000, 3, half=4.0)])
    if xf is not None:  # move the interior origins into the transformed sphere
        rays["o"][:8000] = rays["o"][:8000] * 0.3 + np.asarray(t[0], np.float32).reshape(3, 4)[:, 3]
    a, b, s = _parity(w, o, rays, "warpsim single")
    assert s["exact"] >= 0.995 * len(rays) and b["hit"][:8000].all(), s
    assert np.array_equal(w.trace(rays, any_hit=True)["hit"], b["hit"])


def test_work_distribution_is_irrelevant():
    """Any number of warps, batches smaller than a warp, a batch that is no multiple of 32: identical records."""
    pushes = _scene_instanced(n_inst=12, tess=8)
    w1, w7 = engines.WarpsimEngine(pushes, n_warps=1), engines.WarpsimEngine(pushes, n_warps=7)
    rays = _rays_for(pushes, 3001, 4)
    a = w1.trace(rays)
    assert a.tobytes() == w7.trace(rays).tobytes()
    for n in (0, 1, 31, 33):
        assert w7.trace(rays[:n]).tobytes() == a[:n].tobytes()
    perm = np.random.RandomState(1).permutation(len(rays))
    assert w7.trace(rays[perm]).tobytes() == a[perm].tobytes()
    many = np.repeat(rays[np.nonzero(a["hit"])[0][:1]], 1000)  # all lanes retire together
    h = w7.trace(many)
    assert (h["hit"] == 1).all() and (h["t"] == h["t"][0]).all()


def test_deep_trees_overflow_the_short_stack():
    """Exponentially nested geometry needs more than the kernel's 32-entry stack: those rays are flagged (RC_OVERFLOW_MARK) and
    re-traced by the deep-stack body, as k_trace_fixup does on the GPU; everything still agrees with the oracle."""
    from test_gpu_parity import _deep_scene

    blas = _deep_scene(20)
    g = [2.0 ** -i for i in range(14)]
    xf = np.stack([W.trs3x4((4 * a, 4 * b, 4 * g[(i + j) % 14]), (1, 0, 0, 0), 1.0) for i, a in enumerate(g) for j, b in enumerate(g)])
    pushes = [(blas, None, xf, None)]
    o, w = engines.OracleEngine(pushes), engines.WarpsimEngine(pushes)
    rs = np.random.RandomState(0)
    n = 1024
    d = (np.array([1, 1, 1], np.float32) + rs.uniform(-0.9, 0.9, (n, 3))).astype(np.float32)
    rays = W.make_rays(np.full((n, 3), 1e-9, np.float32), d)
    a, c = w.scene.trace_warpsim(rays, counters=True)
    assert c["short_stack_overflows"] > 0, c
    b = o.trace(rays)
    cls = parity.classify(a, b, parity.make_graze_verifier(orc, rays, a, o.instances, o.tris))
    parity.assert_parity(cls, n, max_tie_frac=0.05, label="warpsim deep trees")
    assert (a["hit"] <= 1).all()
    assert np.array_equal(w.trace(rays, any_hit=True)["hit"], o.trace(rays, any_hit=True)["hit"])


def test_edge_case_rays():
    """Zero / negative-zero direction components, t windows, non-finite rays (must terminate and miss), coincident duplicates."""
    verts = np.concatenate([W.box_mesh(), W.box_mesh(), W.quad_mesh(2.0, 3.0)])
    pushes = [(verts, None, [kat.I34, W.translation3x4((4, 0, 0))], [5, 6])]
    o, w = engines.OracleEngine(pushes), engines.WarpsimEngine(pushes)
    org = np.array([[0.1, 0.2, 5], [0.1, 0.2, 5], [0.5, 0.5, 5], [0.0, 0.0, 0.0], [0.1, 0.2, 0.5], [10, 10, 10], [4.1, 0.1, -5], [0.1, 0.2, 5], [0.1, 0.2, 5]], np.float32)
    d = np.array([[0, 0, -1], [-0.0, -0.0, -1], [0, 0, -1], [1, 0, 0], [0, 0, 1], [0, 0, -1], [0, 0, 1], [0, 0, -1], [0, 0, -1]], np.float32)
    rays = W.make_rays(org, d)
    rays["t_min"][7], rays["t_max"][8] = 4.7, 2.0
    a, b = w.trace(rays), o.trace(rays)
    assert np.array_equal(a["hit"], b["hit"])
    ok = b["hit"] == 1
    assert np.allclose(a["t"][ok], b["t"][ok], rtol=1e-6)
    bad = W.make_rays([[np.nan, 0, 0], [0, 0, 5], [np.inf, 0, 0], [0, 0, 5], [0, 0, 5]], [[0, 0, -1], [np.nan, 0, -1], [0, 0, -1], [np.inf, 0, -1], [0, 0, 0]])
    h = w.trace(bad)
    assert (h["hit"] <= 1).all() and h["hit"][:4].sum() == 0
    assert (w.trace(bad, any_hit=True)["hit"] <= 1).all()
    huge = W.make_rays([[0.1, 0.2, 5]] * 2, [[0, 0, -3e38], [1e-30, 0, -1e-30]])  # reciprocal under- and overflow: must agree on hit / miss
    assert np.array_equal(w.trace(huge)["hit"], o.trace(huge)["hit"])


def test_empty_tlas_misses():
    w = engines.WarpsimEngine([])
    h = w.trace(W.box_rays(10, 1))
    assert not h["hit"].any() and not h["t"].any()


@pytest.mark.parametrize("seed", range(12))
def test_random_scenes(seed):
    """The fuzz cases of tests/test_gpu_fuzz.py (1..40 instances, mixed BLASes, degenerate faces, t windows) through the shipped kernel on the CPU."""
    from test_gpu_fuzz import random_scene

    pushes, rays = random_scene(seed, n=6000)
    o, w = engines.OracleEngine(pushes), engines.WarpsimEngine(pushes, n_warps=2)
    a, b, _ = _parity(w, o, rays, f"warpsim fuzz {seed}", max_tie_frac=0.02)
    aa, ba = w.trace(rays, any_hit=True), o.trace(rays, any_hit=True)
    d = np.nonzero(aa["hit"] != ba["hit"])[0]
    assert len(d) <= 2 and (aa["hit"][d] == 1).all() and b["hit"].sum() > 0
