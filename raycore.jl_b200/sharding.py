"""Multi-GPU plumbing: one process per GPU, BVH replicated, rays / view-factor rows sharded, results
gathered with torch.distributed (NCCL over NVLink on GPUs; gloo in the CPU tests).  The data path itself needs no
collective — every rank traces its own contiguous block."""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of `n` units owned by `rank`: [r*n/G, (r+1)*n/G)."""
    return (rank * n) // world, ((rank + 1) * n) // world


def shard_sizes(n: int, world: int) -> List[int]:
    return [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]


def gather_records(local: torch.Tensor, total_units: int, unit_bytes: int, dst: int = 0) -> Optional[torch.Tensor]:
    """Gather per-rank blocks of fixed-size records (uint8 tensors of shard_size*unit_bytes) to `dst`, in rank order.
    Uneven shards are padded to the largest block for the collective and trimmed on arrival."""
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = shard_sizes(total_units, world)
    assert local.dtype == torch.uint8 and local.numel() == sizes[rank] * unit_bytes
    mx = max(sizes) * unit_bytes
    buf = local
    if local.numel() != mx:
        buf = torch.zeros(mx, dtype=torch.uint8, device=local.device)
        buf[: local.numel()] = local
    out = [torch.empty(mx, dtype=torch.uint8, device=local.device) for _ in range(world)] if rank == dst else None
    dist.gather(buf, out, dst=dst)
    if rank != dst:
        return None
    return torch.cat([o[: s * unit_bytes] for o, s in zip(out, sizes)])


def broadcast_blob(blob: Optional[np.ndarray], src: int = 0, device: Optional[torch.device] = None) -> np.ndarray:
    """Replicate a byte blob (e.g. `TLAS.export_geometry` bytes) from `src` to every rank: the other way to replicate the BVH
    ("rank 0 builds and broadcasts the arrays") when only one rank holds the meshes.  Ranks other than `src` pass None;
    every rank returns the bytes, ready for `TLAS.push_exported`.  Two collectives: the size, then the payload."""
    rank = dist.get_rank()
    n = torch.tensor([0 if blob is None else int(np.asarray(blob).nbytes)], dtype=torch.int64)
    if device is not None:
        n = n.to(device)
    dist.broadcast(n, src=src)
    if rank == src:
        t = torch.from_numpy(np.ascontiguousarray(blob).view(np.uint8).reshape(-1).copy())
    else:
        t = torch.empty(int(n.item()), dtype=torch.uint8)
    if device is not None:
        t = t.to(device)
    dist.broadcast(t, src=src)
    return t.cpu().numpy()


def trace_sharded(trace_fn, rays: np.ndarray, record_dtype: np.dtype, device: Optional[torch.device] = None, dst: int = 0):
    """Every rank holds the same `rays`; rank r traces rays[lo:hi] with `trace_fn` and the hit records are gathered to dst.
    Returns the full hit array on dst, None elsewhere."""
    world, rank = dist.get_world_size(), dist.get_rank()
    lo, hi = shard_range(len(rays), rank, world)
    hits = np.ascontiguousarray(trace_fn(rays[lo:hi]))
    t = torch.from_numpy(hits.view(np.uint8).reshape(-1).copy())
    if device is not None:
        t = t.to(device)
    full = gather_records(t, len(rays), record_dtype.itemsize, dst)
    if full is None:
        return None
    return full.cpu().numpy().view(record_dtype)


def view_factor_rows_sharded(vf_fn, n_prims: int, device: Optional[torch.device] = None, dst: int = 0, interleaved: bool = False):
    """Rank r computes its share of the view-factor matrix's source rows and the shares are gathered to dst (rows in order).
    contiguous (default): vf_fn(row_base, n_rows) -> uint32[n_rows, n_prims] for the block [lo, hi);
    interleaved: vf_fn(row_first, n_rows, row_stride) for the rows r, r + world, ... — equally expensive shares when the cost of a
    row varies along the matrix."""
    world, rank = dist.get_world_size(), dist.get_rank()
    if interleaved:
        counts = [len(range(r, n_prims, world)) for r in range(world)]
        block = np.ascontiguousarray(vf_fn(rank, counts[rank], world), dtype=np.uint32)
    else:
        lo, hi = shard_range(n_prims, rank, world)
        block = np.ascontiguousarray(vf_fn(lo, hi - lo), dtype=np.uint32)
    t = torch.from_numpy(block.view(np.uint8).reshape(-1).copy())
    if device is not None:
        t = t.to(device)
    if not interleaved:
        full = gather_records(t, n_prims, 4 * n_prims, dst)
        return None if full is None else full.cpu().numpy().view(np.uint32).reshape(n_prims, n_prims)
    sizes = [c * 4 * n_prims for c in counts]
    bufs = [torch.empty(sz, dtype=torch.uint8, device=t.device) for sz in sizes] if rank == dst else None
    if world == 1:
        bufs = [t]
    else:
        # ragged gather: pad to the largest share
        pad = torch.zeros(max(sizes), dtype=torch.uint8, device=t.device)
        pad[: t.numel()] = t
        recv = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
        dist.gather(pad, recv, dst=dst)
        if rank != dst:
            return None
        bufs = [recv[r][: sizes[r]] for r in range(world)]
    full = np.zeros((n_prims, n_prims), np.uint32)
    for r in range(world):
        full[r::world] = bufs[r].cpu().numpy().view(np.uint32).reshape(counts[r], n_prims)
    return full


class PeerResultBuffer:
    """A device buffer owned by rank `dst` and mapped into every other rank through CUDA IPC (rc_ipc_export / rc_ipc_open), so
    each rank's traversal kernel can store its hit records straight into its slice over NVLink: the gather is the kernel's
    epilogue, no separate collective runs.  Call `fence()` (stream sync + barrier) before the owner reads.

    Raises RuntimeError on every rank if any rank cannot map the buffer (callers then fall back to `gather_records`)."""

    def __init__(self, tlas, nbytes: int, dst: int = 0):
        import ctypes as C

        self._C, self._lib, self._ctx = C, tlas._lib, tlas._ctx
        self.rank, self.world, self.dst, self.nbytes = dist.get_rank(), dist.get_world_size(), dst, nbytes
        self.base = C.c_void_p()
        handle = (C.c_uint8 * 64)()
        ok = 1
        if self.rank == dst:
            ok = int(self._lib.rc_device_alloc(self._ctx, nbytes, C.byref(self.base)) == 0 and self._lib.rc_ipc_export(self._ctx, self.base, handle) == 0)
        obj = [bytes(handle)]
        dist.broadcast_object_list(obj, src=dst)
        if self.rank != dst:
            hb = (C.c_uint8 * 64).from_buffer_copy(obj[0])
            ok = int(self._lib.rc_ipc_open(self._ctx, hb, C.byref(self.base)) == 0)
        dev = torch.device("cuda", torch.cuda.current_device())
        t = torch.tensor([ok], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        self.ok = int(t.item()) == 1
        if not self.ok:
            self.close()
            raise RuntimeError("CUDA IPC peer mapping unavailable")

    def ptr(self, offset_bytes: int = 0) -> int:
        return self.base.value + offset_bytes

    def fence(self):
        """Everything this rank enqueued on the library context (trace stream and copy stream, on the context's own device — which need
        not be torch's current device) has completed, on every rank."""
        assert self._lib.rc_wait(self._ctx) == 0
        torch.cuda.synchronize()
        dist.barrier()

    def read(self) -> np.ndarray:
        assert self.rank == self.dst
        host = np.empty(self.nbytes, np.uint8)
        assert self._lib.rc_memcpy_d2h(self._ctx, host.ctypes.data, self.base, self.nbytes) == 0
        return host

    def close(self):
        """Collective: peers drain their work and unmap first, then the owner frees (a peer kernel or copy may still be writing into the
        mapping until its rank has passed the first barrier)."""
        if getattr(self, "base", None) is None:
            return
        mapped = bool(self.base.value)
        self._lib.rc_wait(self._ctx)
        if mapped and self.rank != self.dst:
            self._lib.rc_ipc_close(self._ctx, self.base)
            self.base = self._C.c_void_p()
        if dist.is_initialized():
            dist.barrier()
        if mapped and self.rank == self.dst:
            self._lib.rc_device_free(self._ctx, self.base)
            self.base = self._C.c_void_p()
