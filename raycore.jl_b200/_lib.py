"""ctypes binding of libraycore_cuda.so (include/raycore_cuda.h).

The shared library is the product; this module only declares its entry points.  There is no
fallback of any kind: if the library is missing the import fails loudly, and every call needs a
CUDA device (rc_create returns RC_ERR_CUDA without one).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RAYCORE_CUDA_LIB") or os.path.join(_HERE, "libraycore_cuda.so")

RC_OK = 0
RC_ERR_INVALID_ARGUMENT = 1
RC_ERR_INVALID_HANDLE = 2
RC_ERR_DELETED_HANDLE = 3
RC_ERR_NO_VALID_TRIANGLES = 4
RC_ERR_CUDA = 5
RC_ERR_NOT_SYNCED = 6
RC_ERR_OUT_OF_MEMORY = 7
RC_ERR_STACK_OVERFLOW = 8

RC_RAYS_ON_DEVICE = 0x1
RC_HITS_ON_DEVICE = 0x2
RC_MODE_REFERENCE_ORDER = 0x4
RC_COUNTERS = 0x8
RC_VERTS_ON_DEVICE = 0x10
RC_NO_SYNC = 0x20
RC_WAVE_NO_JITTER = 0x40
RC_BUILD_KEEP_BVH2 = 0x80
RC_BUILD_ALLOW_REFIT = 0x100
RC_UPDATE_REFIT = 0x200
RC_MODE_WATERTIGHT = 0x400
RC_IGNORE_TMIN = 0x800
WIDE_NODE_DTYPE = np.dtype([("origin", "<f4", 3), ("sx", "<f4"), ("qlo", "<u4", 3), ("qhi_x", "<u4"), ("qhi_y", "<u4"), ("qhi_z", "<u4"), ("child01", "<u4", 2),
                            ("child23", "<u4", 2), ("sy", "<f4"), ("sz", "<f4")])  # rc_wide_node
assert WIDE_NODE_DTYPE.itemsize == 64
RC_MAX_LIGHTS = 16

RC_SYNC_NONE, RC_SYNC_REFIT, RC_SYNC_REBUILD = 0, 1, 2

RAY_DTYPE = np.dtype([("o", "<f4", 3), ("t_min", "<f4"), ("d", "<f4", 3), ("t_max", "<f4")])
HIT_DTYPE = np.dtype(
    [
        ("hit", "<u4"),
        ("t", "<f4"),
        ("primitive_id", "<u4"),
        ("instance_custom_index", "<u4"),
        ("bary_u", "<f4"),
        ("bary_v", "<f4"),
        ("instance_id", "<u4"),
        ("meta", "<u4"),
    ]
)
INSTANCE_DTYPE = np.dtype(
    [("blas_index", "<u4"), ("instance_id", "<u4"), ("transform", "<f4", 12), ("inv_transform", "<f4", 12), ("flags", "<u4")]
)
NODE2_DTYPE = np.dtype(
    [
        ("aabb0_min", "<f4", 3),
        ("aabb0_max", "<f4", 3),
        ("aabb1_min", "<f4", 3),
        ("aabb1_max", "<f4", 3),
        ("child0", "<u4"),
        ("child1", "<u4"),
        ("parent", "<u4"),
    ]
)

# every symbol include/raycore_cuda.h declares (tests check the library exports all of them)
EXPORTS = [
    "rc_abi_version", "rc_create", "rc_destroy", "rc_last_error", "rc_stream", "rc_set_stream",
    "rc_push", "rc_delete", "rc_update_transforms", "rc_update_transforms_device", "rc_update_geometry", "rc_last_update_refitted", "rc_set_build_flags", "rc_sync",
    "rc_export_geometry", "rc_push_exported", "rc_check_exported",
    "rc_is_valid", "rc_n_instances", "rc_n_instances_of", "rc_n_total_instances", "rc_n_geometries", "rc_is_dirty",
    "rc_get_instances", "rc_world_bound", "rc_wait", "rc_sizes", "rc_read_tlas_nodes", "rc_read_blas_nodes",
    "rc_read_blas_order", "rc_blas_n_prims", "rc_read_blas_faces", "rc_get_instance_handles",
    "rc_trace_closest", "rc_trace_any", "rc_get_counters", "rc_last_kernel_ms", "rc_last_kernel_launches", "rc_last_build_ms",
    "rc_hits_from_grid", "rc_get_illumination", "rc_get_centroid", "rc_view_factors", "rc_view_factors_strided", "rc_view_factor_rays", "rc_read_flat_metadata",
    "rc_collide_instances", "rc_collide_instances_any",
    "rc_set_normals", "rc_generate_primary_rays", "rc_generate_primary_rays_lookat", "rc_generate_shadow_rays", "rc_test_shadow_rays",
    "rc_shadow_visibility",
    "rc_device_alloc", "rc_device_free", "rc_host_alloc", "rc_host_free", "rc_memcpy_h2d", "rc_memcpy_d2h",
    "rc_ipc_export", "rc_ipc_open", "rc_ipc_close", "rc_peer_copy_async", "rc_stream_wait_copy",
    "rc_blas4_build", "rc_blas4_destroy", "rc_blas4_last_error", "rc_blas4_info", "rc_blas4_trace_closest", "rc_blas4_trace_any", "rc_blas4_read_nodes", "rc_blas4_context",
    "rc_multi_create", "rc_multi_destroy", "rc_multi_last_error", "rc_multi_device_count", "rc_multi_context", "rc_multi_push", "rc_multi_delete",
    "rc_multi_update_transforms", "rc_multi_update_geometry", "rc_multi_sync", "rc_multi_trace_closest", "rc_multi_trace_any", "rc_multi_view_factors",
]  # fmt: skip


class RaycoreError(RuntimeError):
    """Mirror of the reference's ErrorException / ArgumentError raised by error(...)."""

    def __init__(self, code, message):
        super().__init__(message)
        self.code = code


_lib = None


def load():
    """dlopen libraycore_cuda.so and declare signatures.  Raises if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C raycore.jl_b200/csrc).  raycore_b200 has no CPU fallback."
        )
    L = C.CDLL(LIB_PATH)
    vp, u32, i32, u64 = C.c_void_p, C.c_uint32, C.c_int32, C.c_uint64
    pu32, pi32 = C.POINTER(C.c_uint32), C.POINTER(C.c_int32)
    sig = {
        "rc_abi_version": (i32, []),
        "rc_create": (i32, [i32, C.POINTER(vp)]),
        "rc_destroy": (i32, [vp]),
        "rc_last_error": (C.c_char_p, [vp]),
        "rc_stream": (vp, [vp]),
        "rc_set_stream": (i32, [vp, vp]),
        "rc_push": (i32, [vp, vp, u32, vp, vp, vp, vp, u32, u32, pu32]),
        "rc_delete": (i32, [vp, u32, pi32]),
        "rc_update_transforms": (i32, [vp, u32, vp, vp, u32]),
        "rc_update_transforms_device": (i32, [vp, u32, vp, vp, u32]),
        "rc_update_geometry": (i32, [vp, u32, vp, u32, vp, u32]),
        "rc_last_update_refitted": (i32, [vp]),
        "rc_set_build_flags": (i32, [vp, u32]),
        "rc_sync": (i32, [vp, pi32]),
        "rc_export_geometry": (i32, [vp, u32, vp, u64, C.POINTER(u64)]),
        "rc_push_exported": (i32, [vp, vp, u64, vp, vp, vp, u32, pu32]),
        "rc_check_exported": (i32, [vp, u64, pu32, pu32, pu32]),
        "rc_is_valid": (i32, [vp, u32]),
        "rc_n_instances": (u32, [vp]),
        "rc_n_instances_of": (u32, [vp, u32]),
        "rc_n_total_instances": (u32, [vp]),
        "rc_n_geometries": (u32, [vp]),
        "rc_is_dirty": (i32, [vp, pi32, pi32]),
        "rc_get_instances": (i32, [vp, u32, vp]),
        "rc_world_bound": (i32, [vp, vp]),
        "rc_wait": (i32, [vp]),
        "rc_sizes": (i32, [vp, pu32, pu32, pu32, pu32]),
        "rc_read_tlas_nodes": (i32, [vp, vp, u32]),
        "rc_read_blas_nodes": (i32, [vp, u32, vp, u32]),
        "rc_read_blas_order": (i32, [vp, u32, vp, u32]),
        "rc_blas_n_prims": (u32, [vp, u32]),
        "rc_read_blas_faces": (i32, [vp, u32, vp, u32]),
        "rc_get_instance_handles": (i32, [vp, vp, u32]),
        "rc_trace_closest": (i32, [vp, vp, vp, u64, u32]),
        "rc_trace_any": (i32, [vp, vp, vp, u64, u32]),
        "rc_get_counters": (i32, [vp, vp, i32]),
        "rc_last_kernel_ms": (C.c_float, [vp]),
        "rc_last_kernel_launches": (u32, [vp]),
        "rc_last_build_ms": (C.c_float, [vp]),
        "rc_hits_from_grid": (i32, [vp, vp, u32, vp, vp]),
        "rc_get_illumination": (i32, [vp, vp, u32, vp, u32]),
        "rc_get_centroid": (i32, [vp, vp, u32, vp, pu32, vp]),
        "rc_view_factors": (i32, [vp, u32, u64, vp, u32, u32, u32, C.POINTER(u64)]),
        "rc_view_factors_strided": (i32, [vp, u32, u64, vp, u32, u32, u32, u32, C.POINTER(u64)]),
        "rc_view_factor_rays": (i32, [vp, u32, u64, u32, u32, vp]),
        "rc_read_flat_metadata": (i32, [vp, vp, u32]),
        "rc_collide_instances": (i32, [vp, vp, u64, C.POINTER(u64)]),
        "rc_collide_instances_any": (i32, [vp, u32, u32, pi32]),
        "rc_set_normals": (i32, [vp, u32, vp, u32, u32]),
        "rc_generate_primary_rays": (i32, [vp, u32, u32, u32, vp, C.c_float, C.c_float, u64, vp, u32]),
        "rc_generate_primary_rays_lookat": (i32, [vp, u32, u32, u32, vp, vp, vp, vp, C.c_float, C.c_float, u64, vp, u32]),
        "rc_generate_shadow_rays": (i32, [vp, vp, vp, u64, vp, u32, C.c_float, vp, u32]),
        "rc_test_shadow_rays": (i32, [vp, vp, u64, vp, u32]),
        "rc_shadow_visibility": (i32, [vp, vp, vp, u64, vp, u32, C.c_float, vp, u32]),
        "rc_device_alloc": (i32, [vp, C.c_size_t, C.POINTER(vp)]),
        "rc_device_free": (i32, [vp, vp]),
        "rc_host_alloc": (i32, [vp, C.c_size_t, C.POINTER(vp)]),
        "rc_host_free": (i32, [vp, vp]),
        "rc_memcpy_h2d": (i32, [vp, vp, vp, C.c_size_t]),
        "rc_memcpy_d2h": (i32, [vp, vp, vp, C.c_size_t]),
        "rc_ipc_export": (i32, [vp, vp, vp]),
        "rc_ipc_open": (i32, [vp, vp, C.POINTER(vp)]),
        "rc_ipc_close": (i32, [vp, vp]),
        "rc_peer_copy_async": (i32, [vp, vp, vp, C.c_size_t, u32]),
        "rc_stream_wait_copy": (i32, [vp, u32]),
        "rc_blas4_build": (i32, [i32, vp, u32, vp, u32, C.POINTER(vp)]),
        "rc_blas4_destroy": (i32, [vp]),
        "rc_blas4_last_error": (C.c_char_p, [vp]),
        "rc_blas4_info": (i32, [vp, pu32, pu32, vp]),
        "rc_blas4_trace_closest": (i32, [vp, vp, vp, u64, u32]),
        "rc_blas4_trace_any": (i32, [vp, vp, vp, u64, u32]),
        "rc_blas4_read_nodes": (i32, [vp, vp, u32]),
        "rc_blas4_context": (vp, [vp]),
        "rc_multi_create": (i32, [vp, u32, C.POINTER(vp)]),
        "rc_multi_destroy": (i32, [vp]),
        "rc_multi_last_error": (C.c_char_p, [vp]),
        "rc_multi_device_count": (u32, [vp]),
        "rc_multi_context": (vp, [vp, u32]),
        "rc_multi_push": (i32, [vp, vp, u32, vp, vp, vp, vp, u32, u32, pu32]),
        "rc_multi_delete": (i32, [vp, u32, pi32]),
        "rc_multi_update_transforms": (i32, [vp, u32, vp, vp, u32]),
        "rc_multi_update_geometry": (i32, [vp, u32, vp, u32, vp, u32]),
        "rc_multi_sync": (i32, [vp, pi32]),
        "rc_multi_trace_closest": (i32, [vp, vp, vp, u64, u32]),
        "rc_multi_trace_any": (i32, [vp, vp, vp, u64, u32]),
        "rc_multi_view_factors": (i32, [vp, u32, u64, vp, C.POINTER(u64)]),
    }
    assert set(sig) == set(EXPORTS)
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L
