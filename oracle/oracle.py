"""ctypes front-end for the CPU oracle (oracle/oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product (raycore.jl_b200) never imports this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

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
TRI_DTYPE = np.dtype([("v", "<f4", 9), ("metadata", "<u4"), ("input_index", "<u4")])
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
INSTANCE_DTYPE = np.dtype(
    [("blas_index", "<u4"), ("instance_id", "<u4"), ("transform", "<f4", 12), ("inv_transform", "<f4", 12), ("flags", "<u4")]
)
assert RAY_DTYPE.itemsize == 32 and HIT_DTYPE.itemsize == 32 and NODE2_DTYPE.itemsize == 60
assert INSTANCE_DTYPE.itemsize == 108 and TRI_DTYPE.itemsize == 44

INVALID_NODE = 0xFFFFFFFF


class _Counters(C.Structure):
    _fields_ = [("nodes", C.c_uint64), ("box_tests", C.c_uint64), ("tri_tests", C.c_uint64), ("inst_entries", C.c_uint64), ("max_stack", C.c_uint32)]


class _Blas(C.Structure):
    _fields_ = [("n", C.c_uint32), ("nodes", C.c_void_p), ("prims", C.c_void_p), ("morton", C.c_void_p), ("root_aabb", C.c_float * 6)]


class _Tlas(C.Structure):
    _fields_ = [
        ("n_nodes", C.c_uint32),
        ("n_instances", C.c_uint32),
        ("n_blas", C.c_uint32),
        ("n_blas_nodes", C.c_uint32),
        ("n_blas_prims", C.c_uint32),
        ("nodes", C.c_void_p),
        ("instances", C.c_void_p),
        ("all_blas_nodes", C.c_void_p),
        ("all_blas_prims", C.c_void_p),
        ("descs", C.c_void_p),
        ("root_aabb", C.c_float * 6),
    ]


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (gcc, -ffp-contract=off)."""
    src = [os.path.join(_HERE, f) for f in ("oracle.c", "oracle.h", "Makefile")]
    if force or not os.path.exists(_LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        fp = C.POINTER(C.c_float)
        vp = C.c_void_p
        L.orc_expand_bits.restype = C.c_uint32
        L.orc_expand_bits.argtypes = [C.c_uint32]
        L.orc_morton_code_30bit.restype = C.c_uint32
        L.orc_morton_code_30bit.argtypes = [fp]
        L.orc_clz32.restype = C.c_int32
        L.orc_clz32.argtypes = [C.c_uint32]
        L.orc_delta.restype = C.c_int32
        L.orc_delta.argtypes = [C.c_int32, C.c_int32, vp, C.c_int32]
        L.orc_is_degenerate.restype = C.c_int
        L.orc_is_degenerate.argtypes = [fp]
        for name in ("orc_mat4_to_mat3x4", "orc_mat3x4_inverse", "orc_safe_invdir"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [fp, fp]
        for name in ("orc_transform_point", "orc_transform_direction"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [fp, fp, fp]
        L.orc_intersect_triangle.restype = C.c_int
        L.orc_intersect_triangle.argtypes = [fp, fp, fp, fp, fp, C.c_float, C.c_float, fp, fp, fp]
        L.orc_intersect_bbox.restype = None
        L.orc_intersect_bbox.argtypes = [fp, fp, fp, fp, C.c_float, C.c_float, fp, fp]
        L.orc_filter_triangles.restype = C.c_uint32
        L.orc_filter_triangles.argtypes = [vp, C.c_uint32, vp, vp]
        L.orc_build_blas.restype = C.POINTER(_Blas)
        L.orc_build_blas.argtypes = [vp, C.c_uint32]
        L.orc_free_blas.restype = None
        L.orc_free_blas.argtypes = [C.POINTER(_Blas)]
        L.orc_build_tlas.restype = C.POINTER(_Tlas)
        L.orc_build_tlas.argtypes = [C.POINTER(C.POINTER(_Blas)), C.c_uint32, vp, C.c_uint32]
        L.orc_free_tlas.restype = None
        L.orc_free_tlas.argtypes = [C.POINTER(_Tlas)]
        L.orc_refit_tlas.restype = None
        L.orc_refit_tlas.argtypes = [C.POINTER(_Tlas)]
        for name in ("orc_closest_hit", "orc_any_hit"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [C.POINTER(_Tlas), vp, vp, C.POINTER(_Counters)]
        for name in ("orc_trace_closest", "orc_trace_any"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [C.POINTER(_Tlas), vp, vp, C.c_uint64, C.c_int, C.POINTER(_Counters)]
        L.orc_trace_mode.restype = None
        L.orc_trace_mode.argtypes = [C.POINTER(_Tlas), vp, vp, C.c_uint64, C.c_int, C.c_int, C.POINTER(_Counters)]
        L.orc_intersect_triangle_watertight.restype = C.c_int
        L.orc_intersect_triangle_watertight.argtypes = [fp, fp, fp, fp, fp, C.c_float, C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.orc_max_threads.restype = C.c_int
        L.orc_generate_ray_grid.restype = None
        L.orc_generate_ray_grid.argtypes = [fp, fp, C.c_uint32, vp, fp]
        L.orc_hits_from_grid.restype = None
        L.orc_hits_from_grid.argtypes = [C.POINTER(_Tlas), fp, C.c_uint32, vp, vp, C.c_int]
        L.orc_get_illumination.restype = None
        L.orc_get_illumination.argtypes = [C.POINTER(_Tlas), fp, C.c_uint32, vp, C.c_int]
        L.orc_get_centroid.restype = C.c_uint32
        L.orc_get_centroid.argtypes = [C.POINTER(_Tlas), fp, C.c_uint32, fp, C.c_int]
        L.orc_view_factors.restype = None
        L.orc_view_factors.argtypes = [C.POINTER(_Tlas), C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, vp, vp, C.c_int]
        L.orc_view_factors_from_rays.restype = None
        L.orc_view_factors_from_rays.argtypes = [C.POINTER(_Tlas), vp, C.c_uint32, C.c_uint32, C.c_uint32, vp, C.c_int]
        L.orc_collide_instances.restype = C.c_uint64
        L.orc_collide_instances.argtypes = [C.POINTER(_Tlas), vp, vp]
        L.orc_collide_instances_any.restype = C.c_int
        L.orc_collide_instances_any.argtypes = [C.POINTER(_Tlas), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
        L.orc_generate_primary_rays.restype = None
        L.orc_generate_primary_rays.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, fp, C.c_float, C.c_float, C.c_uint64, C.c_int, vp]
        L.orc_generate_primary_rays_lookat.restype = None
        L.orc_generate_primary_rays_lookat.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, fp, fp, fp, fp, C.c_float, C.c_float, C.c_uint64, C.c_int, vp]
        L.orc_generate_shadow_rays.restype = None
        L.orc_generate_shadow_rays.argtypes = [C.POINTER(_Tlas), vp, vp, C.c_uint64, vp, vp, C.c_uint32, C.c_float, vp]
        L.orc_test_shadow_rays.restype = None
        L.orc_test_shadow_rays.argtypes = [C.POINTER(_Tlas), vp, C.c_uint64, vp, C.c_int]
        L.orc_rng_uniform.restype = C.c_float
        L.orc_rng_uniform.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32]
        _lib = L
    return _lib


def _f(a, n=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if n is not None:
        assert a.size == n, (a.shape, n)
    return a


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _view(ptr, dtype, count):
    if count == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (dtype.itemsize * count)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=count)


# ---------------------------------------------------------------- scalar helpers
def expand_bits(x):
    return lib().orc_expand_bits(int(x))


def morton_code_30bit(p):
    a = _f(p, 3)
    return lib().orc_morton_code_30bit(_fp(a))


def clz32(x):
    return lib().orc_clz32(int(x))


def delta(i1, i2, codes):
    c = np.ascontiguousarray(codes, dtype=np.uint32)
    return lib().orc_delta(int(i1), int(i2), c.ctypes.data, len(c))


def is_degenerate(v9):
    a = _f(v9, 9)
    return bool(lib().orc_is_degenerate(_fp(a)))


def mat4_to_mat3x4(m4):
    """m4: 4x4 array in mathematical (row, col) indexing == Julia's m[i,j]."""
    a = _f(np.asarray(m4, dtype=np.float32).T.reshape(-1), 16)  # column-major memory
    out = np.zeros(12, np.float32)
    lib().orc_mat4_to_mat3x4(_fp(a), _fp(out))
    return out


def mat3x4_inverse(m):
    a = _f(m, 12)
    out = np.zeros(12, np.float32)
    lib().orc_mat3x4_inverse(_fp(a), _fp(out))
    return out


def transform_point(m, p):
    a, b, out = _f(m, 12), _f(p, 3), np.zeros(3, np.float32)
    lib().orc_transform_point(_fp(a), _fp(b), _fp(out))
    return out


def transform_direction(m, p):
    a, b, out = _f(m, 12), _f(p, 3), np.zeros(3, np.float32)
    lib().orc_transform_direction(_fp(a), _fp(b), _fp(out))
    return out


def safe_invdir(d):
    a, out = _f(d, 3), np.zeros(3, np.float32)
    lib().orc_safe_invdir(_fp(a), _fp(out))
    return out


def intersect_triangle(o, d, v0, v1, v2, t_min=0.0, closest_t=np.inf):
    t, u, v = C.c_float(), C.c_float(), C.c_float()
    arrs = [_f(x, 3) for x in (o, d, v0, v1, v2)]
    ok = lib().orc_intersect_triangle(*[_fp(a) for a in arrs], float(t_min), float(closest_t), C.byref(t), C.byref(u), C.byref(v))
    return bool(ok), t.value, u.value, v.value


def intersect_triangle_watertight(o, d, v0, v1, v2, t_min=0.0, t_max=np.inf):
    """intersect_triangle (src/triangle_mesh.jl:168-201): (hit, t, u, v) with u, v the barycentric weights of v1, v2"""
    t, u, v = C.c_float(), C.c_float(), C.c_float()
    arrs = [_f(x, 3) for x in (o, d, v0, v1, v2)]
    ok = lib().orc_intersect_triangle_watertight(*[_fp(a) for a in arrs], float(t_min), float(t_max), C.byref(t), C.byref(u), C.byref(v))
    return bool(ok), t.value, u.value, v.value


def intersect_bbox(o, inv_d, pmin, pmax, t_min=0.0, t_max=np.inf):
    a, b = C.c_float(), C.c_float()
    arrs = [_f(x, 3) for x in (o, inv_d, pmin, pmax)]
    lib().orc_intersect_bbox(*[_fp(x) for x in arrs], float(t_min), float(t_max), C.byref(a), C.byref(b))
    return a.value, b.value


def rng_uniform(seed, index, dim):
    return lib().orc_rng_uniform(int(seed), int(index), int(dim))


def identity3x4():
    return np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32)


def make_rays(o, d, t_min=0.0, t_max=np.inf):
    o = np.asarray(o, np.float32).reshape(-1, 3)
    d = np.broadcast_to(np.asarray(d, np.float32).reshape(-1, 3), o.shape)
    r = np.zeros(len(o), RAY_DTYPE)
    r["o"], r["d"], r["t_min"], r["t_max"] = o, d, t_min, t_max
    return r


# ---------------------------------------------------------------- builders
def filter_triangles(verts, face_meta=None):
    """is_degenerate_face filter + metadata assignment (instanced-bvh.jl:593-600)."""
    v = _f(np.asarray(verts, np.float32).reshape(-1, 9))
    n = len(v)
    out = np.zeros(n, TRI_DTYPE)
    fm = None if face_meta is None else np.ascontiguousarray(face_meta, np.uint32)
    k = lib().orc_filter_triangles(v.ctypes.data, n, None if fm is None else fm.ctypes.data, out.ctypes.data)
    return out[:k].copy()


class OracleBLAS:
    def __init__(self, tris):
        tris = np.ascontiguousarray(tris, TRI_DTYPE)
        if len(tris) == 0:
            raise ValueError("Cannot build BLAS from empty primitive list")
        self._p = lib().orc_build_blas(tris.ctypes.data, len(tris))
        self.n = len(tris)

    @classmethod
    def from_verts(cls, verts, face_meta=None):
        tris = filter_triangles(verts, face_meta)
        if len(tris) == 0:
            raise ValueError("Geometry has no valid triangles")
        return cls(tris)

    @property
    def nodes(self):
        return _view(self._p.contents.nodes, NODE2_DTYPE, 2 * self.n - 1)

    @property
    def prims(self):
        return _view(self._p.contents.prims, TRI_DTYPE, self.n)

    @property
    def morton(self):
        return _view(self._p.contents.morton, np.dtype("<u4"), self.n)

    @property
    def root_aabb(self):
        return np.array(self._p.contents.root_aabb[:], np.float32)

    def __del__(self):
        if getattr(self, "_p", None):
            lib().orc_free_blas(self._p)
            self._p = None


def make_instances(blas_index, transforms, instance_ids=None, inv_transforms=None):
    transforms = np.asarray(transforms, np.float32).reshape(-1, 12)
    m = len(transforms)
    inst = np.zeros(m, INSTANCE_DTYPE)
    inst["blas_index"] = blas_index
    inst["instance_id"] = 0 if instance_ids is None else instance_ids
    inst["transform"] = transforms
    if inv_transforms is None:
        inst["inv_transform"] = np.stack([mat3x4_inverse(t) for t in transforms]) if m else np.zeros((0, 12), np.float32)
    else:
        inst["inv_transform"] = np.asarray(inv_transforms, np.float32).reshape(-1, 12)
    return inst


class OracleTLAS:
    """StaticTLAS restatement: build_tlas(blas_array, instances)."""

    def __init__(self, blas_list, instances):
        self.blas_list = list(blas_list)
        inst = np.ascontiguousarray(instances, INSTANCE_DTYPE)
        arr = (C.POINTER(_Blas) * max(1, len(self.blas_list)))(*[b._p for b in self.blas_list])
        self._p = lib().orc_build_tlas(arr, len(self.blas_list), inst.ctypes.data, len(inst))
        self.n_instances = len(inst)

    @property
    def c(self):
        return self._p.contents

    @property
    def nodes(self):
        return _view(self.c.nodes, NODE2_DTYPE, self.c.n_nodes if self.n_instances else 0)

    @property
    def instances(self):
        return _view(self.c.instances, INSTANCE_DTYPE, self.n_instances)

    @property
    def all_blas_prims(self):
        return _view(self.c.all_blas_prims, TRI_DTYPE, self.c.n_blas_prims)

    @property
    def all_blas_nodes(self):
        return _view(self.c.all_blas_nodes, NODE2_DTYPE, self.c.n_blas_nodes)

    @property
    def root_aabb(self):
        return np.array(self.c.root_aabb[:], np.float32)

    def refit(self):
        lib().orc_refit_tlas(self._p)

    def _trace(self, fn, rays, threads, counters):
        rays = np.ascontiguousarray(rays, RAY_DTYPE)
        hits = np.zeros(len(rays), HIT_DTYPE)
        cnt = _Counters()
        fn(self._p, rays.ctypes.data, hits.ctypes.data, len(rays), threads, C.byref(cnt) if counters else None)
        if counters:
            return hits, {k: getattr(cnt, k) for k, _ in _Counters._fields_}
        return hits

    def closest_hit(self, rays, threads=0, counters=False, watertight=False):
        """watertight: the reference's pbrt-style triangle test (src/triangle_mesh.jl:168-201) instead of fast_intersect_triangle"""
        if watertight:
            return self._trace(lambda *a: lib().orc_trace_mode(*a[:5], 2, a[5]), rays, threads, counters)
        return self._trace(lib().orc_trace_closest, rays, threads, counters)

    def any_hit(self, rays, threads=0, counters=False, watertight=False):
        if watertight:
            return self._trace(lambda *a: lib().orc_trace_mode(*a[:5], 3, a[5]), rays, threads, counters)
        return self._trace(lib().orc_trace_any, rays, threads, counters)

    def generate_ray_grid(self, direction, grid):
        d = _f(direction, 3)
        origins = np.zeros((grid * grid, 3), np.float32)
        dout = np.zeros(3, np.float32)
        bb = _f(self.root_aabb, 6)
        lib().orc_generate_ray_grid(_fp(bb), _fp(d), grid, origins.ctypes.data, _fp(dout))
        return origins, dout

    def hits_from_grid(self, direction, grid, threads=0):
        d = _f(direction, 3)
        hits = np.zeros(grid * grid, HIT_DTYPE)
        pts = np.zeros((grid * grid, 3), np.float32)
        lib().orc_hits_from_grid(self._p, _fp(d), grid, hits.ctypes.data, pts.ctypes.data, threads)
        return hits, pts

    def get_illumination(self, direction, grid=1000, threads=0):
        d = _f(direction, 3)
        out = np.zeros(self.c.n_blas_prims, np.float32)
        lib().orc_get_illumination(self._p, _fp(d), grid, out.ctypes.data, threads)
        return out

    def get_centroid(self, direction, grid=32, threads=0):
        d = _f(direction, 3)
        c = np.zeros(3, np.float32)
        n = lib().orc_get_centroid(self._p, _fp(d), grid, _fp(c), threads)
        return n, c

    def view_factors(self, rays_per_triangle, seed=0, row_base=0, n_rows=None, want_rays=False, threads=0):
        n = self.c.n_blas_prims
        n_rows = n - row_base if n_rows is None else n_rows
        res = np.zeros((n_rows, n), np.uint32)
        rays = np.zeros(n_rows * rays_per_triangle, RAY_DTYPE) if want_rays else None
        lib().orc_view_factors(self._p, rays_per_triangle, seed, row_base, n_rows, res.ctypes.data, None if rays is None else rays.ctypes.data, threads)
        return (res, rays) if want_rays else res

    def view_factors_from_rays(self, rays, rays_per_triangle, row_base=0, n_rows=None, threads=0):
        n = self.c.n_blas_prims
        n_rows = n - row_base if n_rows is None else n_rows
        rays = np.ascontiguousarray(rays, RAY_DTYPE)
        assert len(rays) == n_rows * rays_per_triangle
        res = np.zeros((n_rows, n), np.uint32)
        lib().orc_view_factors_from_rays(self._p, rays.ctypes.data, rays_per_triangle, row_base, n_rows, res.ctypes.data, threads)
        return res

    def collide_instances(self):
        """collide_instances (src/collision.jl:189-233): (pairs[total, 2] 1-based, inclusive count cache)."""
        n = self.n_instances
        counts = np.zeros(max(n, 1), np.uint32)
        total = lib().orc_collide_instances(self._p, counts.ctypes.data, None)
        pairs = np.zeros((total, 2), np.uint32)
        if total:
            lib().orc_collide_instances(self._p, counts.ctypes.data, pairs.ctypes.data)
        return pairs, counts[:n]

    def collide_instances_any(self, a_range, b_range, literal=False):
        return bool(lib().orc_collide_instances_any(self._p, a_range[0], a_range[1], b_range[0], b_range[1], int(literal)))

    # wavefront stages (docs/src/wavefront-renderer.jl:277-362)
    def generate_shadow_rays(self, rays, hits, lights, shadow_bias=0.01, blas_normals=None):
        """blas_normals: None or a list (one entry per BLAS) of None / float32 (n_prims, 9) arrays indexed by primitive_id."""
        rays = np.ascontiguousarray(rays, RAY_DTYPE)
        hits = np.ascontiguousarray(hits, HIT_DTYPE)
        lights = np.ascontiguousarray(lights, np.float32).reshape(-1, 3)
        out = np.zeros(len(rays) * len(lights), RAY_DTYPE)
        table = None
        keep = []
        if blas_normals is not None:
            table = (C.c_void_p * max(1, len(blas_normals)))()
            for b, a in enumerate(blas_normals):
                if a is not None:
                    a = np.ascontiguousarray(a, np.float32)
                    keep.append(a)
                    table[b] = a.ctypes.data
        lib().orc_generate_shadow_rays(self._p, rays.ctypes.data, hits.ctypes.data, len(rays), table, lights.ctypes.data, len(lights),
                                       shadow_bias, out.ctypes.data)
        return out

    def test_shadow_rays(self, shadow_rays, threads=0):
        shadow_rays = np.ascontiguousarray(shadow_rays, RAY_DTYPE)
        vis = np.zeros(len(shadow_rays), np.uint8)
        lib().orc_test_shadow_rays(self._p, shadow_rays.ctypes.data, len(shadow_rays), vis.ctypes.data, threads)
        return vis

    def __del__(self):
        if getattr(self, "_p", None):
            lib().orc_free_tlas(self._p)
            self._p = None


def max_threads():
    return lib().orc_max_threads()


def generate_primary_rays(width, height, n_samples, camera_pos, focal_length, aspect, seed=0, jitter=True):
    """generate_primary_rays! (docs/src/wavefront-renderer.jl:185-213)"""
    out = np.zeros(width * height * n_samples, RAY_DTYPE)
    cp = _f(camera_pos, 3)
    lib().orc_generate_primary_rays(width, height, n_samples, _fp(cp), focal_length, aspect, seed, int(jitter), out.ctypes.data)
    return out


def generate_primary_rays_lookat(width, height, n_samples, camera_pos, right, up, forward, half_width, half_height, seed=0, jitter=True):
    """generate_primary_rays_lookat! (docs/src/wavefront-renderer.jl:219-253)"""
    out = np.zeros(width * height * n_samples, RAY_DTYPE)
    cp, r, u, f = _f(camera_pos, 3), _f(right, 3), _f(up, 3), _f(forward, 3)
    lib().orc_generate_primary_rays_lookat(width, height, n_samples, _fp(cp), _fp(r), _fp(u), _fp(f), half_width, half_height, seed, int(jitter),
                                           out.ctypes.data)
    return out
