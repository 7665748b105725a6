"""N > 1 host logic on CPU: world_size-2 gloo processes shard rays / view-factor rows, gather to rank 0, and the result must
equal the single-process answer.  The per-rank "tracer" here is the CPU oracle (test infrastructure); on the GPU box the same
sharding functions wrap the CUDA library (bench.py, tests/test_gpu_multi.py)."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as orc
    from raycore_b200 import sharding, workloads as W
    import engines
    import kat

    pushes = [(W.bumpy_sphere(14), None, W.random_trs(9, 3, extent=4.0), None)]
    e = engines.OracleEngine(pushes)
    rays = np.concatenate([W.box_rays(2501, 1, half=6.0), W.interior_rays(2500, 2, radius=5.0)])  # odd count: uneven shards
    full = sharding.trace_sharded(lambda r: e.trace(r), rays, orc.HIT_DTYPE)
    # view-factor row blocks
    q = [(W.quad_mesh(0.0), np.array([1, 2], np.uint32), [kat.I34], None),
         (W.quad_mesh(1.0).reshape(-1, 3, 3)[:, ::-1, :].reshape(-1, 9), np.array([3, 4, ], np.uint32), [kat.I34], None),
         (W.quad_mesh(2.5, 0.5), np.array([5, 6], np.uint32), [kat.I34], None)]
    ev = engines.OracleEngine(q)
    vf = sharding.view_factor_rows_sharded(lambda lo, n: ev.tlas.view_factors(50, seed=9, row_base=lo, n_rows=n), 6)
    # interleaved shares (rows rank, rank + world, ...): ragged gather + re-interleave
    whole = ev.tlas.view_factors(50, seed=9)
    vf_i = sharding.view_factor_rows_sharded(lambda first, n, stride: whole[first::stride][:n], 6, interleaved=True)
    # a serialised geometry built on rank 0 only reaches every rank intact (the host-side import checks run without a GPU)
    from raycore_b200 import tlas as T
    from test_blob_format import make_blob

    mine = make_blob(W.bumpy_sphere(9)) if rank == 0 else None
    got = sharding.broadcast_blob(mine)
    blob_ok = T.check_exported(got) == (len(W.bumpy_sphere(9)) - int(W.is_degenerate(W.bumpy_sphere(9)).sum()), len(W.bumpy_sphere(9)), False)
    flags = [None] * world
    dist.all_gather_object(flags, bool(blob_ok) and (rank != 0 or got.tobytes() == mine.tobytes()))
    if rank == 0:
        ref = e.trace(rays)
        ref_vf = ev.tlas.view_factors(50, seed=9)
        ok = full.tobytes() == ref.tobytes() and np.array_equal(vf, ref_vf) and vf.sum() > 0 and np.array_equal(vf_i, ref_vf) and all(flags)
        open(out_path, "w").write("ok" if ok else "mismatch")
    assert sharding.shard_sizes(5001, 2) == [2500, 2501] and sharding.shard_range(10, 1, 4) == (2, 5)
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    out = str(tmp_path / "result.txt")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert open(out).read() == "ok"
