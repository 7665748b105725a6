"""raycore_b200 — B200-native (sm_100a) implementation of Raycore.jl's ray-query hot path.

The product is `libraycore_cuda.so` (hand-written CUDA behind the C ABI of include/raycore_cuda.h);
this package is the thin host-side mirror of the reference's accel API used by tests and bench.
There is no CPU fallback: importing works without a GPU, every call needs one.
"""
from . import _lib, workloads  # noqa: F401
from ._lib import (  # noqa: F401
    HIT_DTYPE, INSTANCE_DTYPE, NODE2_DTYPE, RAY_DTYPE, RC_SYNC_NONE, RC_SYNC_REBUILD, RC_SYNC_REFIT, RaycoreError,
)
from .tlas import (  # noqa: F401
    BLAS4, INVALID_HANDLE, Bounds3, DeviceQueue, Ray, any_hit4, build_blas4, closest_hit4, RayHit, StaticTLAS, TLAS, MultiTLAS, TLASHandle, Triangle, build_static_tlas, empty_triangle,
    mat4_to_mat3x4, tlas_from_meshes,
)


def closest_hit(accel, ray, **kw):
    """closest_hit(accel, ray) -> (hit, triangle, t, bary, instance_idx) — src/instanced-bvh.jl:1902-2024"""
    return accel.closest_hit(ray, **kw)


def any_hit(accel, ray, **kw):
    """any_hit(accel, ray) — src/instanced-bvh.jl:2034-2140"""
    return accel.any_hit(ray, **kw)


def trace_rays(accel, rays, **kw):
    """trace_rays(tlas, rays) — src/Raycore.jl:116, ext/RaycoreMakieExt.jl:81-87 (batched closest_hit)"""
    return accel.trace_closest(rays, **kw)


def world_bound(accel):
    return accel.world_bound() if hasattr(accel, "world_bound") else accel.root_aabb


def sync(tlas):
    return tlas.sync()


def get_centroid(tlas, viewdir, grid_size=32):
    return tlas.get_centroid(viewdir, grid_size)


def get_illumination(tlas, viewdir, grid_size=1000):
    return tlas.get_illumination(viewdir, grid_size)


def view_factors(tlas, rays_per_triangle=10000, **kw):
    return tlas.view_factors(rays_per_triangle, **kw)
