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
