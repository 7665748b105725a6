"""collide_instances / collide_instances_any (src/collision.jl): oracle vs brute force on the CPU,
CUDA library vs oracle (byte-identical contact list, same order) on the GPU."""
import numpy as np
import pytest

from raycore_b200 import workloads as W
import engines
import kat


def _scene(n, seed, extent):
    return [(W.uv_sphere(6), None, W.random_trs(n, seed, extent=extent), None), (W.box_mesh(), None, W.random_trs(n // 3, seed + 1, extent=extent), None)]


def _brute(tlas):
    n = tlas.n_instances
    leaves = tlas.nodes[n - 1:] if n > 1 else tlas.nodes
    box = {int(l["child1"]): (l["aabb0_min"], l["aabb0_max"]) for l in leaves}
    out = set()
    for a in range(n):
        for b in range(a + 1, n):
            if (box[a][1] >= box[b][0]).all() and (box[a][0] <= box[b][1]).all():
                out.add((a + 1, b + 1))
    return out


def test_oracle_collision_vs_brute_force():
    for n, ext in ((1, 5.0), (2, 0.5), (60, 6.0), (240, 8.0)):
        e = engines.OracleEngine(_scene(n, 4, ext) if n > 2 else [(W.uv_sphere(6), None, W.random_trs(n, 4, extent=ext), None)])
        pairs, counts = e.tlas.collide_instances()
        assert set(map(tuple, pairs.tolist())) == _brute(e.tlas) and len(pairs) == len(_brute(e.tlas))
        assert (pairs[:, 0] < pairs[:, 1]).all() if len(pairs) else True
        assert counts[-1] == len(pairs)
    # collide_instances_any: intended semantics; the literal reference indexing differs when the Morton sort is not the identity
    e = engines.OracleEngine([(W.uv_sphere(6), None, [W.translation3x4((0, 0, 0))], None), (W.uv_sphere(6), None, [W.translation3x4((0.5, 0, 0))], None),
                              (W.uv_sphere(6), None, [W.translation3x4((9, 0, 0))], None)])
    assert e.tlas.collide_instances_any((0, 1), (1, 1)) and not e.tlas.collide_instances_any((0, 1), (2, 1))


@pytest.mark.gpu
def test_cuda_collision_matches_oracle():
    for n, ext in ((1, 5.0), (50, 6.0), (3000, 25.0)):
        pushes = _scene(n, 9, ext) if n > 1 else [(W.uv_sphere(6), None, W.random_trs(1, 4, extent=ext), None)]
        o, g = engines.OracleEngine(pushes), engines.GpuEngine(pushes)
        ref, _ = o.tlas.collide_instances()
        got = g.tlas.collide_instances()
        assert got.dtype == np.uint32 and got.shape == ref.shape
        assert got.tobytes() == ref.tobytes(), "contact list differs from the reference algorithm's (content or order)"
        if n > 1:
            ha, hb = g.handles
            want = any((a - 1 < n) != (b - 1 < n) for a, b in ref.tolist())  # a pair across the two handles
            assert g.tlas.collide_instances_any(ha, hb) == want
        g.tlas.free()
    # empty TLAS
    import raycore_b200 as rc

    t = rc.TLAS()
    assert len(t.collide_instances()) == 0
