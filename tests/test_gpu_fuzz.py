"""Randomised scenes through the C ABI against the oracle: instance counts 1..40 (1 = the single-instance kernel variant), several
BLASes per scene with different triangle counts (1 triangle, a few, hundreds), degenerate faces mixed in, rotated / scaled /
overlapping instances, rays with random t_min / t_max windows, closest and any-hit, plus the reference-order mode bit for bit."""
import numpy as np
import pytest

import engines
import parity
from oracle import oracle as orc
from raycore_b200 import workloads as W


def _random_mesh(rs):
    kind = rs.randint(5)
    if kind == 0:
        m = W.uv_sphere(int(rs.randint(4, 20)))
    elif kind == 1:
        m = W.box_mesh()
    elif kind == 2:
        m = W.quad_mesh(float(rs.uniform(-1, 1)), float(rs.uniform(0.2, 2)))
    elif kind == 3:
        m = rs.uniform(-1, 1, (int(rs.randint(1, 40)), 9)).astype(np.float32)  # triangle soup
    else:
        m = W.bumpy_sphere(int(rs.randint(6, 28)))
    if rs.rand() < 0.4:  # sprinkle degenerate faces
        bad = np.repeat(rs.uniform(-1, 1, (3, 3)).astype(np.float32), 3, axis=1)
        m = np.concatenate([m[: len(m) // 2], bad, m[len(m) // 2:]])
    return np.ascontiguousarray(m, np.float32)


def random_scene(seed, n=20000):
    """(pushes, rays) of fuzz case `seed` (shared with the CPU run of the shipped kernel, tests/test_warpsim_parity.py)."""
    rs = np.random.RandomState(1000 + seed)
    n_inst_total = [1, 1, 2, 3, 5, 8, 13, 21, 40, 4, 1, 17][seed]
    n_blas = min(n_inst_total, int(rs.randint(1, 4)))
    split = np.sort(rs.choice(np.arange(1, n_inst_total), n_blas - 1, replace=False)) if n_blas > 1 else np.array([], int)
    counts = np.diff(np.concatenate([[0], split, [n_inst_total]]))
    pushes = []
    for b in range(n_blas):
        xf = W.random_trs(int(counts[b]), seed=seed * 10 + b, extent=3.0, smin=0.3, smax=2.0)
        ids = rs.randint(0, 1000, int(counts[b])).astype(np.uint32) if rs.rand() < 0.5 else None
        pushes.append((_random_mesh(rs), None, xf, ids))
    rays = W.box_rays(n, seed=seed, half=5.0)
    win = rs.rand(n) < 0.3
    rays["t_min"][win] = rs.uniform(0, 3, win.sum()).astype(np.float32)
    rays["t_max"][win] = rays["t_min"][win] + rs.uniform(0, 6, win.sum()).astype(np.float32)
    return pushes, rays


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(12))
def test_random_scene_parity(seed):
    pushes, rays = random_scene(seed)
    n = len(rays)
    o, g, gr = engines.OracleEngine(pushes), engines.GpuEngine(pushes), engines.GpuEngine(pushes, reference_order=True)
    a, r, b = g.trace(rays), gr.trace(rays), o.trace(rays)
    assert r.tobytes() == b.tobytes(), "reference-order mode differs from the oracle"
    cls = parity.classify(a, b, parity.make_graze_verifier(orc, rays, a, o.instances, o.tris))
    parity.assert_parity(cls, n, max_tie_frac=0.02, label=f"fuzz {seed}")
    # any_hit ignores t_min (src/instanced-bvh.jl:2039); compare against the oracle's any_hit and its reference-order twin
    aa, ra, ba = g.trace(rays, any_hit=True), gr.trace(rays, any_hit=True), o.trace(rays, any_hit=True)
    assert ra.tobytes() == ba.tobytes()
    d = np.nonzero(aa["hit"] != ba["hit"])[0]
    assert len(d) <= 2 and (aa["hit"][d] == 1).all()  # graze class only
    assert b["hit"].sum() > 0
