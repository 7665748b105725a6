"""Host-side mirror of the reference's accel API for the ray-query path, over the C ABI.

Names, argument meaning and error behaviour follow src/instanced-bvh.jl so the parity tests read like
the reference's own tests (Julia `f!(x, ...)` becomes the method `x.f(...)`; Julia's 1-based instance
index in closest_hit's tuple is kept).  Meshes are triangle soups: float32 (n_faces, 9) arrays, i.e.
what `decompose` yields in build_and_append_blas! (:581-608) — meshing itself stays with the caller.

This is what a Julia shim (`julia/RaycoreCUDA.jl`) does with ccall; the Julia toolchain
is absent from this image, so the mirror is Python + ctypes.
"""
from __future__ import annotations

import ctypes as C
from collections import namedtuple
from typing import Optional, Sequence

import numpy as np

from . import _lib as L
from ._lib import HIT_DTYPE, INSTANCE_DTYPE, NODE2_DTYPE, RAY_DTYPE, RaycoreError

TLASHandle = namedtuple("TLASHandle", ["id"])  # src/instanced-bvh.jl:180-185
INVALID_HANDLE = TLASHandle(0)
Triangle = namedtuple("Triangle", ["vertices", "metadata"])  # the part of Triangle{UInt32} the path touches
Ray = namedtuple("Ray", ["o", "d", "t_min", "t_max"], defaults=(0.0, float("inf")))  # src/ray.jl:1-7
Bounds3 = namedtuple("Bounds3", ["p_min", "p_max"])  # src/bounds.jl:6-9
RayHit = namedtuple("RayHit", ["hit", "point", "metadata"])  # src/kernels.jl:1-5

IDENTITY3x4 = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32)


def mat4_to_mat3x4(m) -> np.ndarray:
    """Mat4f (indexed m[i][j], Julia's m[i,j]) -> Mat3x4f rows (src/instanced-bvh.jl:1663-1669)."""
    m = np.asarray(m, np.float32)
    if m.shape == (4, 4):
        return np.ascontiguousarray(m[:3, :]).reshape(12)
    if m.size == 12:
        return np.ascontiguousarray(m, np.float32).reshape(12)
    raise ValueError("transform must be a 4x4 Mat4f or a 12-float Mat3x4f")


def empty_triangle() -> Triangle:
    """Zero sentinel returned on a miss (src/triangle_mesh.jl:49-57)."""
    return Triangle(np.zeros((3, 3), np.float32), np.uint32(0))


def _as_rays(rays) -> np.ndarray:
    if isinstance(rays, np.ndarray) and rays.dtype == RAY_DTYPE:
        return np.ascontiguousarray(rays)
    out = np.zeros(len(rays), RAY_DTYPE)
    for i, r in enumerate(rays):
        out[i] = (tuple(r.o), r.t_min, tuple(r.d), r.t_max)
    return out


BLOB_HEADER_DTYPE = np.dtype(
    [
        ("magic", "S8"), ("abi_version", "<u4"), ("leaf_max", "<u4"), ("hull_boxes", "<u4"), ("n", "<u4"), ("n_faces_in", "<u4"), ("has_normals", "<u4"),
        ("root_aabb", "<f4", 6), ("total_bytes", "<u8"), ("payload_hash", "<u8"),
        ("off_nodes2", "<u8"), ("off_nodes4", "<u8"), ("off_tris", "<u8"), ("off_hull", "<u8"), ("off_normals", "<u8"), ("sphere", "<f4", 4),
    ]
)  # RcBlobHeader, csrc/rc_build.cu
assert BLOB_HEADER_DTYPE.itemsize == 128
BLOB_TRI_DTYPE = np.dtype([("v0", "<f4", 3), ("prim_id", "<u4"), ("v1", "<f4", 3), ("metadata", "<u4"), ("v2", "<f4", 3), ("face_index", "<u4")])  # RcTri


def blob_header(blob) -> np.void:
    """Header record of an `export_geometry` blob."""
    return np.frombuffer(memoryview(blob)[:128], BLOB_HEADER_DTYPE)[0]


def blob_triangles(blob) -> np.ndarray:
    """Morton-sorted triangle records of an `export_geometry` blob."""
    h = blob_header(blob)
    return np.frombuffer(memoryview(blob), BLOB_TRI_DTYPE, count=int(h["n"]), offset=int(h["off_tris"]))


def blob_hash(payload) -> int:
    """Payload hash of an `export_geometry` blob (blob_hash, csrc/rc_build.cu): word-wise multiply-xorshift over everything after the header."""
    b = bytes(payload)
    m64 = (1 << 64) - 1
    h = 0x9E3779B97F4A7C15 ^ len(b)
    nw = len(b) // 8
    for w in np.frombuffer(b, "<u8", count=nw).tolist():
        h = ((h ^ w) * 0xD6E8FEB86659FD93) & m64
        h ^= h >> 32
    for x in b[8 * nw:]:
        h = ((h ^ x) * 0x100000001B3) & m64
    return h


def check_exported(blob):
    """rc_check_exported: the host-side import checks (header, section table, size, payload hash) without a context or a GPU.
    Returns (n_triangles, n_faces, has_normals); raises RaycoreError with the library's message on refusal."""
    lib = L.load()
    b = np.ascontiguousarray(np.frombuffer(blob, np.uint8) if not isinstance(blob, np.ndarray) else blob.view(np.uint8).reshape(-1))
    n, f, hn = C.c_uint32(), C.c_uint32(), C.c_uint32()
    rc = lib.rc_check_exported(b.ctypes.data if b.size else None, b.nbytes, C.byref(n), C.byref(f), C.byref(hn))
    if rc != L.RC_OK:
        raise RaycoreError(rc, lib.rc_last_error(None).decode())
    return n.value, f.value, bool(hn.value)


def blob_faces(blob) -> np.ndarray:
    """n_faces_in x 9 vertex soup in submission order (faces the degenerate filter dropped are zero)."""
    h, t = blob_header(blob), blob_triangles(blob)
    out = np.zeros((int(h["n_faces_in"]), 9), np.float32)
    out[t["face_index"]] = np.concatenate([t["v0"], t["v1"], t["v2"]], axis=1)
    return out


class DeviceQueue:
    """A device-resident work queue (the reference's SoA queues on the backend, docs/src/wavefront-renderer.jl:127-180):
    `count` records of numpy dtype `dtype` in memory owned by the TLAS's context."""

    def __init__(self, owner: "TLAS", dtype, count: int):
        self._owner = owner
        self.dtype = np.dtype(dtype)
        self.count = int(count)
        p = C.c_void_p()
        owner._ck(owner._lib.rc_device_alloc(owner._ctx, max(1, self.count * self.dtype.itemsize), C.byref(p)))
        self.ptr = p.value

    def upload(self, host) -> "DeviceQueue":
        a = np.ascontiguousarray(host, self.dtype)
        if len(a) != self.count:
            raise ValueError(f"queue holds {self.count} records, got {len(a)}")
        if self.count:
            self._owner._ck(self._owner._lib.rc_memcpy_h2d(self._owner._ctx, self.ptr, a.ctypes.data, a.nbytes))
        return self

    def download(self) -> np.ndarray:
        out = np.zeros(self.count, self.dtype)
        if self.count:
            self._owner._ck(self._owner._lib.rc_memcpy_d2h(self._owner._ctx, out.ctypes.data, self.ptr, out.nbytes))
        return out

    def free(self):
        if getattr(self, "ptr", None) and self._owner._ctx:
            self._owner._lib.rc_device_free(self._owner._ctx, self.ptr)
        self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class StaticTLAS:
    """Adapted (immutable) form, `AbstractAdaptedAccel` (src/Raycore.jl:14-49, src/instanced-bvh.jl:155-168).

    Owned by `TLAS.sync`: the same object is kept across refits and replaced on rebuilds, like
    `tlas.static_tlas` in the reference (test/test_mesh_update.jl:184-227)."""

    def __init__(self, owner: "TLAS", generation: int):
        self._owner = owner
        self._generation = generation

    def _check(self):
        if self._owner._static is not self:
            raise RaycoreError(L.RC_ERR_NOT_SYNCED, "stale StaticTLAS: the TLAS was rebuilt; re-adapt per dispatch (src/instanced-bvh.jl:221-226)")

    # queries ---------------------------------------------------------------------------------
    def trace_closest(self, rays, reference_order=False, counters=False, watertight=False) -> np.ndarray:
        """watertight=True: the reference's watertight triangle test (src/triangle_mesh.jl:168-201) instead of Moeller-Trumbore"""
        self._check()
        return self._owner._trace(rays, any_hit=False, reference_order=reference_order, counters=counters, watertight=watertight)

    def trace_any(self, rays, reference_order=False, counters=False, watertight=False) -> np.ndarray:
        self._check()
        return self._owner._trace(rays, any_hit=True, reference_order=reference_order, counters=counters, watertight=watertight)

    def closest_hit(self, ray: Ray, **kw):
        """(hit, Triangle, t, bary(w,u,v), instance_idx 1-based) — src/instanced-bvh.jl:1902-2024."""
        h = self.trace_closest(_as_rays([ray]), **kw)[0]
        return self._owner._tuple_from_hit(h, any_hit=False)

    def any_hit(self, ray: Ray, **kw):
        h = self.trace_any(_as_rays([ray]), **kw)[0]
        return self._owner._tuple_from_hit(h, any_hit=True)

    # introspection ---------------------------------------------------------------------------
    @property
    def n_instances(self) -> int:
        return self._owner._synced_instances

    @property
    def n_geometries(self) -> int:
        return self._owner._synced_geometries

    @property
    def root_aabb(self) -> Bounds3:
        return self._owner.world_bound()


class TLAS:
    """Mutable top-level acceleration structure, `AbstractAccel` (src/instanced-bvh.jl:261-358)."""

    def __init__(self, device: Optional[int] = None, *, keep_bvh2: bool = False, allow_refit: bool = False):
        """keep_bvh2: also emit the reference-layout BVH2 of every geometry (RC_BUILD_KEEP_BVH2: reference-order mode, read_blas_nodes);
        allow_refit: keep the radix tree so `update(..., refit=True)` can re-fit moved vertices (RC_BUILD_ALLOW_REFIT)."""
        self._lib = L.load()
        ctx = C.c_void_p()
        rc = self._lib.rc_create(-1 if device is None else int(device), C.byref(ctx))
        if rc != L.RC_OK:
            raise RaycoreError(rc, self._lib.rc_last_error(None).decode())
        self._ctx = ctx
        if keep_bvh2 or allow_refit:
            self._ck(self._lib.rc_set_build_flags(ctx, (L.RC_BUILD_KEEP_BVH2 if keep_bvh2 else 0) | (L.RC_BUILD_ALLOW_REFIT if allow_refit else 0)))
        self._meshes = {}  # handle id -> (verts, face_meta) kept to materialise Triangle results
        self._static: Optional[StaticTLAS] = None
        self._generation = 0
        self._synced_instances = 0
        self._synced_geometries = 0
        self._inst_handles = np.zeros(0, np.uint32)
        self._face_cache = {}
        self.last_sync_action = L.RC_SYNC_NONE

    # -- plumbing -----------------------------------------------------------------------------
    def _ck(self, rc):
        if rc != L.RC_OK:
            raise RaycoreError(rc, self._lib.rc_last_error(self._ctx).decode())

    def free(self):
        """free!(tlas) — src/instanced-bvh.jl:383-399."""
        if getattr(self, "_ctx", None):
            self._lib.rc_destroy(self._ctx)
            self._ctx = None

    def __del__(self):  # finalizer(free!, tlas), :355
        try:
            self.free()
        except Exception:
            pass

    # -- mutation -----------------------------------------------------------------------------
    @staticmethod
    def _verts(verts):
        v = np.ascontiguousarray(np.asarray(verts, np.float32).reshape(-1, 9))
        return v

    def push(self, mesh, transform=None, *, instance_id: int = 0, instance_ids: Optional[Sequence[int]] = None, face_meta=None,
             inv_transform=None) -> TLASHandle:
        """push!(tlas, mesh, transform; instance_id) / push!(tlas, mesh, transforms; instance_ids) — :639-676.

        `transform`: Mat4f (4x4), Mat3x4f (12 floats), or a list/array of them for instancing."""
        v = self._verts(mesh)
        xf, inv, ids, m = self._instance_args(transform, instance_id, instance_ids, inv_transform)
        fm = None if face_meta is None else np.ascontiguousarray(face_meta, np.uint32)
        if fm is not None and len(fm) != len(v):
            raise ValueError("face_meta length != number of faces")
        h = C.c_uint32()
        self._ck(
            self._lib.rc_push(
                self._ctx, v.ctypes.data, len(v), None if fm is None else fm.ctypes.data, xf.ctypes.data, None if inv is None else inv.ctypes.data,
                None if ids is None else ids.ctypes.data, m, 0, C.byref(h),
            )
        )
        self._meshes[h.value] = (v, fm)
        return TLASHandle(h.value)

    @staticmethod
    def _instance_args(transform, instance_id, instance_ids, inv_transform):
        """(transforms m x 12, inverse transforms or None, instance ids or None, m) of a push — the argument rules of :639-676."""
        if transform is None:
            xf = IDENTITY3x4.reshape(1, 12).copy()
            multi = False
        else:
            t = np.asarray(transform, np.float32)
            multi = not (t.shape == (4, 4) or t.shape == (12,))
            xf = np.stack([mat4_to_mat3x4(x) for x in transform]) if multi else mat4_to_mat3x4(t).reshape(1, 12)
        m = len(xf)
        if multi:
            if instance_ids is not None and len(instance_ids) != m:
                raise ValueError(f"instance_ids length {len(instance_ids)} != transforms length {m}")  # ArgumentError, :664-666
            ids = None if instance_ids is None else np.ascontiguousarray(instance_ids, np.uint32)
        else:
            ids = np.array([instance_id], np.uint32)
        inv = None if inv_transform is None else np.ascontiguousarray(np.asarray(inv_transform, np.float32).reshape(-1, 12))
        return np.ascontiguousarray(xf, np.float32), inv, ids, m

    # -- serialised geometry (to_gpu(ArrayType, blas::BLAS), src/kernel-abstractions.jl:31-36) ----
    def export_geometry(self, handle: TLASHandle) -> np.ndarray:
        """The handle's built geometry (BVH2, wide nodes, sorted triangles, hull, normals) as one byte blob."""
        size = C.c_uint64()
        self._ck(self._lib.rc_export_geometry(self._ctx, handle.id, None, 0, C.byref(size)))
        blob = np.empty(size.value, np.uint8)
        self._ck(self._lib.rc_export_geometry(self._ctx, handle.id, blob.ctypes.data, blob.nbytes, C.byref(size)))
        return blob

    def push_exported(self, blob, transform=None, *, instance_id: int = 0, instance_ids: Optional[Sequence[int]] = None, inv_transform=None) -> TLASHandle:
        """push! of a geometry restored from `export_geometry` bytes: no builder kernel runs, the device arrays are byte-identical."""
        blob = np.ascontiguousarray(np.frombuffer(blob, np.uint8) if not isinstance(blob, np.ndarray) else blob.view(np.uint8).reshape(-1))
        xf, inv, ids, m = self._instance_args(transform, instance_id, instance_ids, inv_transform)
        h = C.c_uint32()
        self._ck(
            self._lib.rc_push_exported(
                self._ctx, blob.ctypes.data, blob.nbytes, xf.ctypes.data, None if inv is None else inv.ctypes.data,
                None if ids is None else ids.ctypes.data, m, C.byref(h),
            )
        )
        self._meshes[h.value] = (blob_faces(blob), None)  # the submitted soup as far as the blob keeps it (dropped faces are zero)
        return TLASHandle(h.value)

    def delete(self, handle: TLASHandle) -> bool:
        """delete!(tlas, handle)::Bool — :690-699."""
        d = C.c_int32()
        self._ck(self._lib.rc_delete(self._ctx, handle.id, C.byref(d)))
        return bool(d.value)

    def update_transform(self, handle: TLASHandle, transform):
        """update_transform! — :755-770 (single-instance handles only)."""
        n = self._lib.rc_n_instances_of(self._ctx, handle.id)
        if self.is_valid(handle) and n != 1:
            raise RaycoreError(L.RC_ERR_INVALID_ARGUMENT, f"Handle has {n} instances, use update_transforms! for multiple")
        self.update_transforms(handle, [transform])

    def update_transforms(self, handle: TLASHandle, transforms):
        """update_transforms! — :784-797."""
        xf = np.ascontiguousarray(np.stack([mat4_to_mat3x4(x) for x in transforms]), np.float32)
        self._ck(self._lib.rc_update_transforms(self._ctx, handle.id, xf.ctypes.data, None, len(xf)))

    def update_transforms_device(self, handle: TLASHandle, transforms: "DeviceQueue"):
        """update_transforms! with the Mat3x4f array resident on the device (float32, 12 per instance), e.g. written by a
        caller's kernel — the `instance_buffer` use case (src/Raycore.jl:118-130)."""
        if transforms.dtype != np.float32 or transforms.count % 12:
            raise ValueError("device transforms must be float32 with 12 values per instance")
        self._ck(self._lib.rc_update_transforms_device(self._ctx, handle.id, transforms.ptr, None, transforms.count // 12))

    def update(self, handle: TLASHandle, mesh, face_meta=None, refit: bool = False) -> bool:
        """update!(tlas, handle, new_geometry) — :808-857.  refit=True asks for a re-fit of the kept radix tree (RC_UPDATE_REFIT: same faces,
        moved vertices, geometry built with allow_refit); returns True when the library re-fitted, False when it rebuilt."""
        v = self._verts(mesh)
        fm = None if face_meta is None else np.ascontiguousarray(face_meta, np.uint32)
        self._ck(self._lib.rc_update_geometry(self._ctx, handle.id, v.ctypes.data, len(v), None if fm is None else fm.ctypes.data, L.RC_UPDATE_REFIT if refit else 0))
        refitted = bool(self._lib.rc_last_update_refitted(self._ctx))
        old = self._meshes.get(handle.id)
        self._meshes[handle.id] = (v, old[1] if refitted and old is not None else fm)  # a refit keeps the geometry's metadata
        self._face_cache.clear()
        return refitted

    def sync(self) -> "TLAS":
        """sync!(tlas) — :894-921."""
        a = C.c_int32()
        self._ck(self._lib.rc_sync(self._ctx, C.byref(a)))
        self.last_sync_action = a.value
        if a.value == L.RC_SYNC_REBUILD or self._static is None:
            self._generation += 1
            self._static = StaticTLAS(self, self._generation)  # rebuild_static_tlas!, :930-959
            self._face_cache.clear()
            n = self._lib.rc_n_total_instances(self._ctx)
            self._inst_handles = np.zeros(n, np.uint32)
            if n:
                self._ck(self._lib.rc_get_instance_handles(self._ctx, self._inst_handles.ctypes.data, n))
            for hid in [k for k in self._meshes if not self._lib.rc_is_valid(self._ctx, k)]:
                del self._meshes[hid]
        if a.value != L.RC_SYNC_NONE:
            self._synced_instances = self._lib.rc_n_total_instances(self._ctx)
            self._synced_geometries = self._lib.rc_n_geometries(self._ctx)
        return self

    @property
    def static_tlas(self) -> Optional[StaticTLAS]:
        return self._static

    def adapt(self) -> StaticTLAS:
        """Adapt.adapt(backend, tlas): sync! then return tlas.static_tlas — :1085-1102."""
        self.sync()
        return self._static

    # -- introspection ------------------------------------------------------------------------
    def is_valid(self, handle: TLASHandle) -> bool:
        return bool(self._lib.rc_is_valid(self._ctx, handle.id))

    def n_instances(self, handle: Optional[TLASHandle] = None) -> int:
        if handle is None:
            return self._lib.rc_n_instances(self._ctx)
        return self._lib.rc_n_instances_of(self._ctx, handle.id)

    def n_total_instances(self) -> int:
        return self._lib.rc_n_total_instances(self._ctx)

    def n_geometries(self) -> int:
        return self._lib.rc_n_geometries(self._ctx)

    @property
    def dirty(self) -> bool:
        d, t = C.c_int32(), C.c_int32()
        self._lib.rc_is_dirty(self._ctx, C.byref(d), C.byref(t))
        return bool(d.value)

    @property
    def transforms_dirty(self) -> bool:
        d, t = C.c_int32(), C.c_int32()
        self._lib.rc_is_dirty(self._ctx, C.byref(d), C.byref(t))
        return bool(t.value)

    def get_instances(self, handle: TLASHandle) -> np.ndarray:
        """get_instances — :732-738 (raises for invalid / deleted handles)."""
        n = max(1, self._lib.rc_n_instances_of(self._ctx, handle.id))
        out = np.zeros(n, INSTANCE_DTYPE)
        self._ck(self._lib.rc_get_instances(self._ctx, handle.id, out.ctypes.data))
        return out

    def get_instance(self, handle: TLASHandle, instance_idx: int = 1):
        """get_instance — :714-723 (1-based instance_idx)."""
        inst = self.get_instances(handle)
        if not (1 <= instance_idx <= len(inst)):
            raise RaycoreError(L.RC_ERR_INVALID_ARGUMENT, f"Instance index {instance_idx} out of range 1:{len(inst)}")
        return inst[instance_idx - 1]

    def world_bound(self) -> Bounds3:
        b = np.zeros(6, np.float32)
        self._ck(self._lib.rc_world_bound(self._ctx, b.ctypes.data))
        return Bounds3(b[:3].copy(), b[3:].copy())

    def wait_for_gpu(self) -> "TLAS":
        """wait_for_gpu!(accel) === accel — :2418-2421."""
        self._ck(self._lib.rc_wait(self._ctx))
        return self

    def sizes(self):
        a, b, c, d = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
        self._ck(self._lib.rc_sizes(self._ctx, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return {"tlas_nodes": a.value, "blas_nodes": b.value, "blas_prims": c.value, "pending_deletes": d.value}

    def read_tlas_nodes(self) -> np.ndarray:
        n = self.sizes()["tlas_nodes"]
        out = np.zeros(n, NODE2_DTYPE)
        if n:
            self._ck(self._lib.rc_read_tlas_nodes(self._ctx, out.ctypes.data, n))
        return out

    def read_blas_nodes(self, blas_index: int) -> np.ndarray:
        n = self._lib.rc_blas_n_prims(self._ctx, blas_index)
        out = np.zeros(max(0, 2 * n - 1), NODE2_DTYPE)
        self._ck(self._lib.rc_read_blas_nodes(self._ctx, blas_index, out.ctypes.data, len(out)))
        return out

    def read_blas_order(self, blas_index: int) -> np.ndarray:
        n = self._lib.rc_blas_n_prims(self._ctx, blas_index)
        out = np.zeros(n, np.uint32)
        self._ck(self._lib.rc_read_blas_order(self._ctx, blas_index, out.ctypes.data, n))
        return out

    def read_blas_faces(self, blas_index: int) -> np.ndarray:
        n = self._lib.rc_blas_n_prims(self._ctx, blas_index)
        out = np.zeros(n, np.uint32)
        self._ck(self._lib.rc_read_blas_faces(self._ctx, blas_index, out.ctypes.data, n))
        return out

    def flat_metadata(self) -> np.ndarray:
        n = self.sizes()["blas_prims"]
        out = np.zeros(n, np.uint32)
        if n:
            self._ck(self._lib.rc_read_flat_metadata(self._ctx, out.ctypes.data, n))
        return out

    # -- queries --------------------------------------------------------------------------------
    def _trace(self, rays, any_hit, reference_order=False, counters=False, watertight=False) -> np.ndarray:
        rays = _as_rays(rays)
        hits = np.zeros(len(rays), HIT_DTYPE)
        flags = (L.RC_MODE_REFERENCE_ORDER if reference_order else 0) | (L.RC_COUNTERS if counters else 0) | (L.RC_MODE_WATERTIGHT if watertight else 0)
        fn = self._lib.rc_trace_any if any_hit else self._lib.rc_trace_closest
        if len(rays):
            self._ck(fn(self._ctx, rays.ctypes.data, hits.ctypes.data, len(rays), flags))
        return hits

    def trace_closest(self, rays, **kw) -> np.ndarray:
        return self.adapt().trace_closest(rays, **kw)

    def trace_any(self, rays, **kw) -> np.ndarray:
        return self.adapt().trace_any(rays, **kw)

    def closest_hit(self, ray: Ray, **kw):
        return self.adapt().closest_hit(ray, **kw)

    def any_hit(self, ray: Ray, **kw):
        return self.adapt().any_hit(ray, **kw)

    def counters(self, reset=True) -> dict:
        out = (C.c_uint64 * 6)()
        self._ck(self._lib.rc_get_counters(self._ctx, out, 1 if reset else 0))
        keys = ["rays", "nodes", "box_tests", "tri_tests", "inst_entries", "max_stack"]
        return dict(zip(keys, [int(x) for x in out]))

    def last_kernel_ms(self) -> float:
        return float(self._lib.rc_last_kernel_ms(self._ctx))

    def triangle_of(self, hit) -> Triangle:
        """Materialise the reference's `Triangle` for a hit from the caller-side copy of the mesh."""
        inst_pos = int(hit["instance_id"])
        hid = int(self._inst_handles[inst_pos])
        verts, _ = self._meshes[hid]
        blas_index = int(self.get_instances(TLASHandle(hid))[0]["blas_index"])
        faces = self._face_cache.get(blas_index)
        if faces is None:
            faces = self._face_cache[blas_index] = self.read_blas_faces(blas_index)
        face = int(faces[int(hit["primitive_id"])])
        return Triangle(verts[face].reshape(3, 3).copy(), np.uint32(hit["meta"]))

    def _tuple_from_hit(self, h, any_hit: bool):
        if h["hit"]:
            u, v = np.float32(h["bary_u"]), np.float32(h["bary_v"])
            w = np.float32(1.0) - u - v  # :2015
            return True, self.triangle_of(h), np.float32(h["t"]), np.array([w, u, v], np.float32), np.uint32(h["instance_id"] + 1)
        return False, empty_triangle(), np.float32(0), np.zeros(3, np.float32), np.uint32(0)

    # -- collision broad phase (src/collision.jl) --------------------------------------------------
    def collide_instances(self) -> np.ndarray:
        """collide_instances(tlas) — src/collision.jl:189-233: uint32[total, 2] of 1-based (instance_a, instance_b), a < b."""
        self.sync()
        n = C.c_uint64()
        self._ck(self._lib.rc_collide_instances(self._ctx, None, 0, C.byref(n)))
        out = np.zeros((n.value, 2), np.uint32)
        if n.value:
            self._ck(self._lib.rc_collide_instances(self._ctx, out.ctypes.data, n.value, C.byref(n)))
        return out

    def collide_instances_any(self, handle_a: TLASHandle, handle_b: TLASHandle) -> bool:
        """collide_instances_any — src/collision.jl:241-261."""
        self.sync()
        o = C.c_int32()
        self._ck(self._lib.rc_collide_instances_any(self._ctx, handle_a.id, handle_b.id, C.byref(o)))
        return bool(o.value)

    # -- analysis (src/kernels.jl) --------------------------------------------------------------
    def hits_from_grid(self, viewdir, grid_size: int = 32):
        """hits_from_grid — :58-72: (hits[grid*grid], points[grid*grid,3]) in Julia column-major cell order."""
        self.sync()
        d = np.ascontiguousarray(viewdir, np.float32)
        n = grid_size * grid_size
        hits = np.zeros(n, HIT_DTYPE)
        pts = np.zeros((n, 3), np.float32)
        self._ck(self._lib.rc_hits_from_grid(self._ctx, d.ctypes.data, grid_size, hits.ctypes.data, pts.ctypes.data))
        return hits, pts

    def get_centroid(self, viewdir, grid_size: int = 32):
        """get_centroid — :106-110: (surface_points, mean)."""
        self.sync()
        d = np.ascontiguousarray(viewdir, np.float32)
        c = np.zeros(3, np.float32)
        n = C.c_uint32()
        pts = np.zeros((grid_size * grid_size, 3), np.float32)
        self._ck(self._lib.rc_get_centroid(self._ctx, d.ctypes.data, grid_size, c.ctypes.data, C.byref(n), pts.ctypes.data))
        return pts[: n.value].copy(), c

    def get_illumination(self, viewdir, grid_size: int = 1000) -> np.ndarray:
        """get_illumination — :112-124: Float32 hit count per metadata 1..n_prims."""
        self.sync()
        d = np.ascontiguousarray(viewdir, np.float32)
        n = self.sizes()["blas_prims"]
        out = np.zeros(n, np.float32)
        if n:
            self._ck(self._lib.rc_get_illumination(self._ctx, d.ctypes.data, grid_size, out.ctypes.data, n))
        return out

    def view_factors(self, rays_per_triangle: int = 10000, seed: int = 0, row_base: int = 0, n_rows: Optional[int] = None, row_stride: int = 1) -> np.ndarray:
        """view_factors — :74-104.  Returns result[src, hit] (UInt32, indexable like Julia's result[src_meta, hit_meta]
        with 0-based numpy indices = metadata - 1); `row_base` / `row_stride` / `n_rows` select the source rows
        row_base + k * row_stride, k < n_rows (a contiguous block with stride 1, an interleaved share with stride = world size)."""
        self.sync()
        n = self.sizes()["blas_prims"]
        if n_rows is None:
            n_rows = max(0, -(-(n - row_base) // row_stride))
        out = np.zeros((n_rows, n), np.uint32)
        sk = C.c_uint64()
        if n and n_rows:
            self._ck(self._lib.rc_view_factors_strided(self._ctx, rays_per_triangle, seed, out.ctypes.data, row_base, row_stride, n_rows, 0, C.byref(sk)))
        self.last_vf_skipped = sk.value
        return out

    def view_factor_rays(self, rays_per_triangle: int, seed: int = 0, row_base: int = 0, n_rows: Optional[int] = None) -> np.ndarray:
        self.sync()
        n = self.sizes()["blas_prims"]
        n_rows = n - row_base if n_rows is None else n_rows
        out = np.zeros(n_rows * rays_per_triangle, RAY_DTYPE)
        if len(out):
            self._ck(self._lib.rc_view_factor_rays(self._ctx, rays_per_triangle, seed, row_base, n_rows, out.ctypes.data))
        return out


    # -- wavefront stages on device-resident queues (docs/src/wavefront-renderer.jl:185-362) -------------
    def queue(self, dtype, count: int) -> DeviceQueue:
        return DeviceQueue(self, dtype, count)

    def set_normals(self, handle: TLASHandle, normals):
        """Triangle.normals of the handle's geometry: float32 (n_faces, 9) = (n0, n1, n2) per submitted face."""
        n = np.ascontiguousarray(np.asarray(normals, np.float32).reshape(-1, 9))
        self._ck(self._lib.rc_set_normals(self._ctx, handle.id, n.ctypes.data, len(n), 0))

    def generate_primary_rays(self, width: int, height: int, camera_pos, focal_length: float, aspect: float, n_samples: int = 1, seed: int = 0,
                              jitter: bool = True, out: Optional[DeviceQueue] = None) -> DeviceQueue:
        """generate_primary_rays! — :185-213."""
        q = out or DeviceQueue(self, RAY_DTYPE, width * height * n_samples)
        cp = np.ascontiguousarray(camera_pos, np.float32)
        self._ck(self._lib.rc_generate_primary_rays(self._ctx, width, height, n_samples, cp.ctypes.data, focal_length, aspect, seed, q.ptr,
                                                    0 if jitter else L.RC_WAVE_NO_JITTER))
        return q

    def generate_primary_rays_lookat(self, width: int, height: int, camera_pos, right, up, forward, half_width: float, half_height: float,
                                     n_samples: int = 1, seed: int = 0, jitter: bool = True, out: Optional[DeviceQueue] = None) -> DeviceQueue:
        """generate_primary_rays_lookat! — :219-253."""
        q = out or DeviceQueue(self, RAY_DTYPE, width * height * n_samples)
        v = [np.ascontiguousarray(x, np.float32) for x in (camera_pos, right, up, forward)]
        self._ck(self._lib.rc_generate_primary_rays_lookat(self._ctx, width, height, n_samples, v[0].ctypes.data, v[1].ctypes.data, v[2].ctypes.data,
                                                           v[3].ctypes.data, half_width, half_height, seed, q.ptr, 0 if jitter else L.RC_WAVE_NO_JITTER))
        return q

    def intersect_rays(self, rays: DeviceQueue, any_hit: bool = False, out: Optional[DeviceQueue] = None) -> DeviceQueue:
        """intersect_primary_rays! — :260-275: closest_hit over a device ray queue into a device hit queue."""
        self.sync()
        q = out or DeviceQueue(self, HIT_DTYPE, rays.count)
        fn = self._lib.rc_trace_any if any_hit else self._lib.rc_trace_closest
        if rays.count:
            self._ck(fn(self._ctx, rays.ptr, q.ptr, rays.count, L.RC_RAYS_ON_DEVICE | L.RC_HITS_ON_DEVICE))
        return q

    def generate_shadow_rays(self, rays: DeviceQueue, hits: DeviceQueue, lights, shadow_bias: float = 0.01, out: Optional[DeviceQueue] = None) -> DeviceQueue:
        """generate_shadow_rays! — :277-330: queue of rays.count * n_lights shadow rays."""
        self.sync()
        lt = np.ascontiguousarray(np.asarray(lights, np.float32).reshape(-1, 3))
        q = out or DeviceQueue(self, RAY_DTYPE, rays.count * len(lt))
        self._ck(self._lib.rc_generate_shadow_rays(self._ctx, rays.ptr, hits.ptr, rays.count, lt.ctypes.data, len(lt), shadow_bias, q.ptr, 0))
        return q

    def test_shadow_rays(self, shadow_rays: DeviceQueue, out: Optional[DeviceQueue] = None) -> DeviceQueue:
        """test_shadow_rays! — :337-362: one visibility byte per shadow ray."""
        self.sync()
        q = out or DeviceQueue(self, np.uint8, shadow_rays.count)
        self._ck(self._lib.rc_test_shadow_rays(self._ctx, shadow_rays.ptr, shadow_rays.count, q.ptr, 0))
        return q

    def shadow_visibility(self, rays: DeviceQueue, hits: DeviceQueue, lights, shadow_bias: float = 0.01, out: Optional[DeviceQueue] = None) -> DeviceQueue:
        """generate_shadow_rays! + test_shadow_rays! in one kernel (no shadow-ray queue in memory)."""
        self.sync()
        lt = np.ascontiguousarray(np.asarray(lights, np.float32).reshape(-1, 3))
        q = out or DeviceQueue(self, np.uint8, rays.count * len(lt))
        self._ck(self._lib.rc_shadow_visibility(self._ctx, rays.ptr, hits.ptr, rays.count, lt.ctypes.data, len(lt), shadow_bias, q.ptr, 0))
        return q


class MultiTLAS:
    """One scene replicated on several GPUs of this node, queries sharded over them — inside the library, one process (rc_multi_*,
    SURVEY §8e).  Mutations mirror TLAS; `trace_closest` / `trace_any` / `view_factors` return exactly what a single-device TLAS returns."""

    def __init__(self, devices: Optional[Sequence[int]] = None):
        self._lib = L.load()
        m = C.c_void_p()
        dv = None if devices is None else np.ascontiguousarray(devices, np.int32)
        rc = self._lib.rc_multi_create(None if dv is None else dv.ctypes.data, 0 if dv is None else len(dv), C.byref(m))
        if rc != L.RC_OK:
            raise RaycoreError(rc, self._lib.rc_multi_last_error(None).decode())
        self._m = m

    def _ck(self, rc):
        if rc != L.RC_OK:
            raise RaycoreError(rc, self._lib.rc_multi_last_error(self._m).decode())

    @property
    def n_devices(self) -> int:
        return int(self._lib.rc_multi_device_count(self._m))

    def free(self):
        if getattr(self, "_m", None):
            self._lib.rc_multi_destroy(self._m)
            self._m = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def push(self, mesh, transform=None, *, instance_id: int = 0, instance_ids: Optional[Sequence[int]] = None, face_meta=None, inv_transform=None) -> TLASHandle:
        v = TLAS._verts(mesh)
        xf, inv, ids, m = TLAS._instance_args(transform, instance_id, instance_ids, inv_transform)
        fm = None if face_meta is None else np.ascontiguousarray(face_meta, np.uint32)
        h = C.c_uint32()
        self._ck(self._lib.rc_multi_push(self._m, v.ctypes.data, len(v), None if fm is None else fm.ctypes.data, xf.ctypes.data, None if inv is None else inv.ctypes.data,
                                         None if ids is None else ids.ctypes.data, m, 0, C.byref(h)))
        return TLASHandle(h.value)

    def delete(self, handle: TLASHandle) -> bool:
        d = C.c_int32()
        self._ck(self._lib.rc_multi_delete(self._m, handle.id, C.byref(d)))
        return bool(d.value)

    def update_transforms(self, handle: TLASHandle, transforms):
        xf = np.ascontiguousarray([mat4_to_mat3x4(t) for t in transforms], np.float32)
        self._ck(self._lib.rc_multi_update_transforms(self._m, handle.id, xf.ctypes.data, None, len(xf)))

    def update(self, handle: TLASHandle, mesh, face_meta=None, refit: bool = False):
        v = TLAS._verts(mesh)
        fm = None if face_meta is None else np.ascontiguousarray(face_meta, np.uint32)
        self._ck(self._lib.rc_multi_update_geometry(self._m, handle.id, v.ctypes.data, len(v), None if fm is None else fm.ctypes.data, L.RC_UPDATE_REFIT if refit else 0))

    def sync(self) -> int:
        a = C.c_int32()
        self._ck(self._lib.rc_multi_sync(self._m, C.byref(a)))
        return a.value

    def _trace(self, rays, any_hit, watertight=False):
        rays = _as_rays(rays)
        hits = np.zeros(len(rays), HIT_DTYPE)
        fn = self._lib.rc_multi_trace_any if any_hit else self._lib.rc_multi_trace_closest
        if len(rays):
            self._ck(fn(self._m, rays.ctypes.data, hits.ctypes.data, len(rays), L.RC_MODE_WATERTIGHT if watertight else 0))
        return hits

    def trace_closest(self, rays, **kw) -> np.ndarray:
        return self._trace(rays, False, **kw)

    def trace_any(self, rays, **kw) -> np.ndarray:
        return self._trace(rays, True, **kw)

    def view_factors(self, rays_per_triangle: int = 10000, seed: int = 0) -> np.ndarray:
        n = C.c_uint32()
        self._ck(self._lib.rc_sizes(self._lib.rc_multi_context(self._m, 0), None, None, C.byref(n), None))
        out = np.zeros((n.value, n.value), np.uint32)
        sk = C.c_uint64()
        if n.value:
            self._ck(self._lib.rc_multi_view_factors(self._m, rays_per_triangle, seed, out.ctypes.data, C.byref(sk)))
        return out


def build_static_tlas(meshes, metadata_fn=None, device: Optional[int] = None) -> StaticTLAS:
    """TLAS(meshes, metadata_fn) — src/instanced-bvh.jl:2276-2324: one BLAS + identity instance per mesh,
    instance_id = mesh index (1-based), metadata = metadata_fn(mesh_idx, face_idx) (both 1-based)."""
    tlas = TLAS(device)
    for mi, mesh in enumerate(meshes, start=1):
        v = TLAS._verts(mesh)
        fm = None
        if metadata_fn is not None:
            fm = np.array([metadata_fn(mi, fi) for fi in range(1, len(v) + 1)], np.uint32)
        tlas.push(v, None, instance_id=mi, face_meta=fm)
    return tlas.adapt()


def tlas_from_meshes(meshes, device: Optional[int] = None):
    """TLAS(meshes) -> (tlas, handles) — src/instanced-bvh.jl:2361-2378."""
    if len(meshes) == 0:
        raise RaycoreError(L.RC_ERR_INVALID_ARGUMENT, "Cannot create TLAS from empty mesh list")
    tlas = TLAS(device)
    handles = [tlas.push(m) for m in meshes]
    tlas.sync()
    return tlas, handles


class BLAS4:
    """BLAS4 / build_blas4 / closest_hit4 / any_hit4 — src/bvh4.jl:154-163, 511-522, 606-766.

    The reference collapses its LBVH into 128-byte 4-wide nodes on the host; the library's own quantised 4-wide BVH (the one
    every TLAS query already uses) backs the same API: a BLAS4 is one geometry under an identity instance, and the `*4` queries
    are rc_trace_closest / rc_trace_any on it.  As in the reference, closest_hit4 / any_hit4 ignore ray.t_min
    (`ray_mint = 0`, src/bvh4.jl:610, :700) and return (hit, Triangle, t, bary) without an instance index."""

    def __init__(self, primitives, face_meta=None, device: Optional[int] = None):
        v = np.ascontiguousarray(np.asarray(primitives, np.float32).reshape(-1, 9))
        if len(v) == 0:
            raise RaycoreError(L.RC_ERR_NO_VALID_TRIANGLES, "Cannot build BLAS4 from empty primitive list")  # src/bvh4.jl:513
        self._lib = L.load()
        fm = None if face_meta is None else np.ascontiguousarray(face_meta, np.uint32)
        b = C.c_void_p()
        rc = self._lib.rc_blas4_build(-1 if device is None else int(device), v.ctypes.data, len(v), None if fm is None else fm.ctypes.data, 0, C.byref(b))
        if rc != L.RC_OK:
            raise RaycoreError(rc, self._lib.rc_blas4_last_error(None).decode())
        self._b, self._verts, self._faces = b, v, None

    def _ck(self, rc):
        if rc != L.RC_OK:
            raise RaycoreError(rc, self._lib.rc_blas4_last_error(self._b).decode())

    def _info(self):
        n, slots, box = C.c_uint32(), C.c_uint32(), np.zeros(6, np.float32)
        self._ck(self._lib.rc_blas4_info(self._b, C.byref(n), C.byref(slots), box.ctypes.data))
        return n.value, slots.value, box

    @property
    def root_aabb(self) -> Bounds3:
        box = self._info()[2]
        return Bounds3(box[:3].copy(), box[3:].copy())

    @property
    def n_primitives(self) -> int:
        return self._info()[0]

    def nodes(self) -> np.ndarray:
        """the wide nodes (rc_wide_node records, slot 0 unused, root = slot 1)"""
        out = np.zeros(self._info()[1], L.WIDE_NODE_DTYPE)
        self._ck(self._lib.rc_blas4_read_nodes(self._b, out.ctypes.data, len(out)))
        return out

    def _trace(self, rays, any_hit, watertight=False) -> np.ndarray:
        rays = _as_rays(rays)
        hits = np.zeros(len(rays), HIT_DTYPE)
        fn = self._lib.rc_blas4_trace_any if any_hit else self._lib.rc_blas4_trace_closest
        if len(rays):
            self._ck(fn(self._b, rays.ctypes.data, hits.ctypes.data, len(rays), L.RC_MODE_WATERTIGHT if watertight else 0))
        return hits

    def trace_closest4(self, rays, **kw) -> np.ndarray:
        return self._trace(rays, False, **kw)

    def trace_any4(self, rays, **kw) -> np.ndarray:
        return self._trace(rays, True, **kw)

    def _tuple(self, h):
        if not h["hit"]:
            return False, empty_triangle(), np.float32(0), np.zeros(3, np.float32)
        if self._faces is None:
            n = self.n_primitives
            self._faces = np.zeros(n, np.uint32)
            rc = self._lib.rc_read_blas_faces(self._lib.rc_blas4_context(self._b), 1, self._faces.ctypes.data, n)
            self._ck(rc)
        u, v = np.float32(h["bary_u"]), np.float32(h["bary_v"])
        tri = Triangle(self._verts[int(self._faces[int(h["primitive_id"])])].reshape(3, 3).copy(), np.uint32(h["meta"]))
        return True, tri, np.float32(h["t"]), np.array([np.float32(1.0) - u - v, u, v], np.float32)

    def closest_hit4(self, ray: Ray):
        return self._tuple(self.trace_closest4([ray])[0])

    def any_hit4(self, ray: Ray):
        return self._tuple(self.trace_any4([ray])[0])

    def free(self):
        if getattr(self, "_b", None):
            self._lib.rc_blas4_destroy(self._b)
            self._b = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def build_blas4(primitives, face_meta=None, device: Optional[int] = None) -> BLAS4:
    """build_blas4(primitives) -> BLAS4 — src/bvh4.jl:511-522"""
    return BLAS4(primitives, face_meta, device)


def closest_hit4(blas: BLAS4, ray: Ray):
    return blas.closest_hit4(ray)


def any_hit4(blas: BLAS4, ray: Ray):
    return blas.any_hit4(ray)
