"""ctypes wrapper of tests/hostsim (CPU instantiation of the library's RC_HD device code).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

from raycore_b200._lib import HIT_DTYPE, INSTANCE_DTYPE, NODE2_DTYPE, RAY_DTYPE

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim")
_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-C", _DIR, "libhostsim.so"], stdout=subprocess.DEVNULL)
        L = C.CDLL(os.environ.get("RC_HOSTSIM_LIB") or os.path.join(_DIR, "libhostsim.so"))  # the override loads an experiment build (other -D flags)
        vp = C.c_void_p
        L.hs_blas_build.restype = vp
        L.hs_blas_build.argtypes = [vp, C.c_uint32, vp]
        L.hs_blas_n.restype = C.c_uint32
        L.hs_blas_n.argtypes = [vp]
        for f in ("hs_blas_nodes2", "hs_blas_root", "hs_blas_order", "hs_blas_nodes4"):
            getattr(L, f).restype = None
            getattr(L, f).argtypes = [vp, vp]
        L.hs_blas_free.argtypes = [vp]
        L.hs_mat3x4_inverse.argtypes = [vp, vp]
        L.hs_scene_build.restype = vp
        L.hs_scene_build.argtypes = [vp, C.c_uint32, vp, C.c_uint32]
        L.hs_scene_tlas_nodes2.restype = C.c_uint32
        L.hs_scene_tlas_nodes2.argtypes = [vp, vp]
        L.hs_scene_root.argtypes = [vp, vp]
        L.hs_scene_free.argtypes = [vp]
        L.hs_trace_watertight.restype = C.c_uint32
        L.hs_trace_watertight.argtypes = [vp, vp, vp, C.c_uint64, C.c_int, C.c_int]
        L.hs_trace.restype = C.c_uint32
        L.hs_trace.argtypes = [vp, vp, vp, C.c_uint64, C.c_int, C.c_int, vp]
        L.hs_primary_rays.restype = None
        L.hs_primary_rays.argtypes = [vp, vp, vp, vp, C.c_float, C.c_float, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, vp]
        L.hs_shadow_rays.restype = None
        L.hs_shadow_rays.argtypes = [vp, vp, vp, C.c_uint64, vp, vp, C.c_uint32, C.c_float, vp]
        L.hs_collapse_cached_stats.restype = None
        L.hs_collapse_cached_stats.argtypes = [vp]
        L.hs_check_wide.restype = C.c_uint32
        L.hs_check_wide.argtypes = [vp]
        L.hs_trace_warpsim.restype = C.c_uint32
        L.hs_trace_warpsim.argtypes = [vp, vp, vp, C.c_uint64, C.c_int, C.c_uint32, vp, vp]
        L.hs_validate_blas.restype = C.c_uint32
        L.hs_validate_blas.argtypes = [vp, C.c_int, C.c_uint32]
        _lib = L
    return _lib


def collapse_cached_stats():
    """(wide nodes checked, mismatches) of rc_collapse_node_cached against rc_collapse_node over every tree built in this process."""
    out = np.zeros(2, np.uint64)
    lib().hs_collapse_cached_stats(out.ctypes.data)
    return int(out[0]), int(out[1])


class HsBlas:
    def __init__(self, verts, face_meta=None):
        v = np.ascontiguousarray(np.asarray(verts, np.float32).reshape(-1, 9))
        fm = None if face_meta is None else np.ascontiguousarray(face_meta, np.uint32)
        self.p = lib().hs_blas_build(v.ctypes.data, len(v), None if fm is None else fm.ctypes.data)
        if not self.p:
            raise ValueError("Geometry has no valid triangles")
        self.n = lib().hs_blas_n(self.p)

    def nodes2(self):
        out = np.zeros(2 * self.n - 1, NODE2_DTYPE)
        lib().hs_blas_nodes2(self.p, out.ctypes.data)
        return out

    def root(self):
        out = np.zeros(6, np.float32)
        lib().hs_blas_root(self.p, out.ctypes.data)
        return out

    def order(self):
        out = np.zeros(self.n, np.uint32)
        lib().hs_blas_order(self.p, out.ctypes.data)
        return out

    def nodes4(self):
        """the wide nodes as raw 64-byte records (n + 1 slots, slot 0 unused)"""
        out = np.zeros((self.n + 1, 64), np.uint8)
        lib().hs_blas_nodes4(self.p, out.ctypes.data)
        return out

    def check_wide(self):
        return lib().hs_check_wide(self.p)

    def validate(self, corrupt=0, where=0):
        """rc_validate_blas_elem over the whole BLAS (optionally after one injected fault): number of bad references."""
        return lib().hs_validate_blas(self.p, corrupt, where)


def mat3x4_inverse(m):
    a = np.ascontiguousarray(m, np.float32)
    out = np.zeros(12, np.float32)
    lib().hs_mat3x4_inverse(a.ctypes.data, out.ctypes.data)
    return out


class HsScene:
    def __init__(self, blas_list, instances):
        self.blas_list = list(blas_list)
        inst = np.ascontiguousarray(instances, INSTANCE_DTYPE)
        arr = (C.c_void_p * max(1, len(self.blas_list)))(*[b.p for b in self.blas_list])
        self.p = lib().hs_scene_build(arr, len(self.blas_list), inst.ctypes.data, len(inst))
        self.n = len(inst)

    def tlas_nodes2(self):
        n = lib().hs_scene_tlas_nodes2(self.p, None)
        out = np.zeros(n, NODE2_DTYPE)
        if n:
            lib().hs_scene_tlas_nodes2(self.p, out.ctypes.data)
        return out

    def root(self):
        out = np.zeros(6, np.float32)
        lib().hs_scene_root(self.p, out.ctypes.data)
        return out

    def shadow_rays(self, rays, hits, lights, shadow_bias=0.01, blas_normals=None):
        rays = np.ascontiguousarray(rays, RAY_DTYPE)
        hits = np.ascontiguousarray(hits, HIT_DTYPE)
        lights = np.ascontiguousarray(lights, np.float32).reshape(-1, 3)
        out = np.zeros(len(rays) * len(lights), RAY_DTYPE)
        table, keep = None, []
        if blas_normals is not None:
            table = (C.c_void_p * max(1, len(blas_normals)))()
            for b, a in enumerate(blas_normals):
                if a is not None:
                    keep.append(np.ascontiguousarray(a, np.float32))
                    table[b] = keep[-1].ctypes.data
        lib().hs_shadow_rays(self.p, rays.ctypes.data, hits.ctypes.data, len(rays), table, lights.ctypes.data, len(lights), shadow_bias, out.ctypes.data)
        return out

    def trace(self, rays, any_hit=False, wide=True, counters=False, watertight=False):
        rays = np.ascontiguousarray(rays, RAY_DTYPE)
        hits = np.zeros(len(rays), HIT_DTYPE)
        if watertight:
            ov = lib().hs_trace_watertight(self.p, rays.ctypes.data, hits.ctypes.data, len(rays), int(any_hit), int(wide))
            assert ov == 0, f"{ov} traversal stack overflows"
            return hits
        cnt = (C.c_uint64 * 5)()
        ov = lib().hs_trace(self.p, rays.ctypes.data, hits.ctypes.data, len(rays), int(any_hit), int(wide), cnt)
        assert ov == 0, f"{ov} traversal stack overflows"
        if counters:
            return hits, dict(zip(["nodes", "box_tests", "tri_tests", "inst_entries", "max_stack"], [int(x) for x in cnt]))
        return hits

    def trace_warpsim(self, rays, any_hit=False, n_warps=3, counters=False):
        """The shipped kernel k_trace_wide (rc_trace_fast.cuh) on the CPU: n_warps warps of 32 fibres, lock step at the warp intrinsics.
        Returns hits (and, with counters, the kernel's work counters plus the run's info)."""
        rays = np.ascontiguousarray(rays, RAY_DTYPE)
        hits = np.full(len(rays), 0xAB, np.uint8).repeat(HIT_DTYPE.itemsize).view(HIT_DTYPE).copy()  # every record must be written
        cnt, info = (C.c_uint64 * 6)(), (C.c_uint64 * 18)()
        rc = lib().hs_trace_warpsim(self.p, rays.ctypes.data, hits.ctypes.data, len(rays), int(any_hit), n_warps, cnt if counters else None, info)
        assert info[3] == 0, "lanes of a warp left the kernel at different times (divergence around a warp intrinsic)"
        assert rc == 0, f"{rc} rays overflowed the deep stack"
        if counters:
            d = dict(zip(["rays", "nodes", "box_tests", "tri_tests", "inst_entries", "max_stack"], [int(x) for x in cnt]))
            d.update(short_stack_overflows=int(info[0]), exchanges=int(info[2]))
            d["step_iterations"] = dict(zip("NTXF", [int(x) for x in info[4:8]]))  # warp iterations per step kind
            d["step_lanes"] = dict(zip("NTXF", [int(x) for x in info[8:12]]))  # active lanes summed over those iterations
            d["idle_in_node_steps"] = dict(zip(["T_second_leaf", "T_at_sentinel", "T_stack_empty", "X", "F", "dead"], [int(x) for x in info[12:18]]))
            return hits, d
        return hits


def primary_rays(width, height, n_samples, camera_pos, focal_length=None, aspect=None, lookat=None, seed=0, jitter=True):
    """lookat = (right, up, forward, half_width, half_height) for generate_primary_rays_lookat!, else the pinhole form."""
    out = np.zeros(width * height * n_samples, RAY_DTYPE)
    cp = np.ascontiguousarray(camera_pos, np.float32)
    if lookat is None:
        fw = np.array([0, 0, focal_length], np.float32)
        lib().hs_primary_rays(cp.ctypes.data, None, None, fw.ctypes.data, aspect, 1.0, 0, int(jitter), width, height, n_samples, seed, out.ctypes.data)
    else:
        r, u, f = (np.ascontiguousarray(x, np.float32) for x in lookat[:3])
        lib().hs_primary_rays(cp.ctypes.data, r.ctypes.data, u.ctypes.data, f.ctypes.data, lookat[3], lookat[4], 1, int(jitter), width, height, n_samples, seed,
                              out.ctypes.data)
    return out
