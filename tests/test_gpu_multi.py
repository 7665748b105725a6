"""Multi-GPU (one process per GPU, NCCL): rays sharded over ranks, BVH rebuilt on every rank, hit records gathered to rank 0;
the gathered result must equal a single-GPU trace of the whole batch.  Skipped on boxes with fewer than 2 GPUs."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import raycore_b200 as rc
    from raycore_b200 import sharding, workloads as W

    tlas = rc.TLAS(rank)
    tlas.push(W.bumpy_sphere(40), list(W.random_trs(50, 3, extent=6.0)))
    tlas.sync()
    rays = np.concatenate([W.box_rays(100001, 1, half=8.0), W.interior_rays(100000, 2, radius=7.0)])
    full = sharding.trace_sharded(lambda r: tlas.trace_closest(r), rays, rc.HIT_DTYPE, device=torch.device("cuda", rank))
    # fused gather: every rank's kernel stores its slice of the hit records directly into rank 0's buffer (CUDA IPC peer pointer)
    import ctypes as C

    lo, hi = sharding.shard_range(len(rays), rank, world)
    buf = sharding.PeerResultBuffer(tlas, len(rays) * 32)
    d_rays = torch.from_numpy(rays[lo:hi].view(np.uint8).reshape(-1).copy()).cuda()
    L = rc._lib
    assert tlas._lib.rc_trace_closest(tlas._ctx, d_rays.data_ptr(), C.c_void_p(buf.ptr(lo * 32)), hi - lo, L.RC_RAYS_ON_DEVICE | L.RC_HITS_ON_DEVICE) == 0
    buf.fence()
    if rank == 0:
        ref = tlas.trace_closest(rays)
        fused = buf.read().view(rc.HIT_DTYPE)
        ok = full.tobytes() == ref.tobytes() and fused.tobytes() == ref.tobytes()
        open(out_path, "w").write("ok" if ok else "mismatch")
    dist.barrier()
    buf.close()
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_nccl(tmp_path):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(2, 29700 + (os.getpid() % 1000), out), nprocs=2, join=True)
    assert open(out).read() == "ok"


def test_multi_tlas_in_library_matches_single_device():
    """rc_multi_* (one process, every visible GPU, inside the library): replicated scene, sharded rays / view-factor rows; the results are
    byte-identical to a single-device TLAS.  Runs with one GPU too (one shard); the pure-C twin is tests/cabi/cabi_multi.c."""
    import raycore_b200 as rc
    from raycore_b200 import workloads as W

    mesh, xf = W.bumpy_sphere(24), list(W.random_trs(40, 3, extent=6.0))
    m, s = rc.MultiTLAS(), rc.TLAS()
    assert m.n_devices >= 1
    hm, hs = m.push(mesh, xf), s.push(mesh, xf)
    assert hm.id == hs.id and m.sync() == rc.RC_SYNC_REBUILD
    s.sync()
    rays = np.concatenate([W.box_rays(150_001, 1, half=8.0), W.interior_rays(50_000, 2, radius=7.0)])
    assert m.trace_closest(rays).tobytes() == s.trace_closest(rays).tobytes()
    assert np.array_equal(m.trace_any(rays)["hit"], s.trace_any(rays)["hit"])
    assert m.trace_closest(rays, watertight=True).tobytes() == s.trace_closest(rays, watertight=True).tobytes()
    xf2 = [t.copy() for t in xf]
    for t in xf2:
        t[3] += 0.5
    m.update_transforms(hm, xf2)
    s.update_transforms(hs, xf2)
    assert m.sync() == rc.RC_SYNC_REFIT
    s.sync()
    assert m.trace_closest(rays).tobytes() == s.trace_closest(rays).tobytes()
    assert m.delete(hm) and s.delete(hs)
    # view factors: one mesh, dense metadata, rows sharded over the devices
    keep = ~W.is_degenerate(mesh)
    meta = np.zeros(len(mesh), np.uint32)
    meta[keep] = 1 + np.arange(keep.sum())
    m.push(mesh, None, face_meta=meta); s.push(mesh, None, face_meta=meta)
    m.sync(); s.sync()
    assert np.array_equal(m.view_factors(200, seed=5), s.view_factors(200, seed=5))
    m.free(); s.free()
