"""Three ways to answer the same queries on the same scene description, so every KAT / parity test can be run
against (a) the CPU oracle, (b) the host simulation of the device code, (c) the CUDA library through the C ABI.

A scene is a list of pushes, each `(verts[n,9], face_meta|None, transforms[m,12], instance_ids[m]|None)` — exactly
the arguments of push!(tlas, mesh, transforms; instance_ids) (src/instanced-bvh.jl:661-676)."""
import numpy as np

from oracle import oracle as orc
import hostsim_py as hs
from raycore_b200._lib import INSTANCE_DTYPE


def _norm_push(p):
    verts, fm, xf, ids = p
    xf = np.asarray(xf, np.float32).reshape(-1, 12)
    return np.asarray(verts, np.float32).reshape(-1, 9), fm, xf, ids


def instances_of(pushes, inverse_fn):
    out = []
    for b, p in enumerate(pushes, start=1):
        _, _, xf, ids = _norm_push(p)
        inst = np.zeros(len(xf), INSTANCE_DTYPE)
        inst["blas_index"] = b
        inst["instance_id"] = 0 if ids is None else ids
        inst["transform"] = xf
        inst["inv_transform"] = np.stack([inverse_fn(t) for t in xf])
        out.append(inst)
    return np.concatenate(out) if out else np.zeros(0, INSTANCE_DTYPE)


class OracleEngine:
    name = "oracle"

    def __init__(self, pushes):
        self.blas = [orc.OracleBLAS.from_verts(_norm_push(p)[0], _norm_push(p)[1]) for p in pushes]
        self.tris = {b + 1: orc.filter_triangles(_norm_push(p)[0], _norm_push(p)[1]) for b, p in enumerate(pushes)}
        self.instances = instances_of(pushes, orc.mat3x4_inverse)
        self.tlas = orc.OracleTLAS(self.blas, self.instances.view(orc.INSTANCE_DTYPE))

    def trace(self, rays, any_hit=False, watertight=False, **kw):
        return self.tlas.any_hit(rays, watertight=watertight) if any_hit else self.tlas.closest_hit(rays, watertight=watertight)

    def world_bound(self):
        return self.tlas.root_aabb


class HostsimEngine:
    def __init__(self, pushes, wide=True):
        self.name = "hostsim-wide" if wide else "hostsim-ref"
        self.wide = wide
        self.blas = [hs.HsBlas(_norm_push(p)[0], _norm_push(p)[1]) for p in pushes]
        self.instances = instances_of(pushes, hs.mat3x4_inverse)
        self.scene = hs.HsScene(self.blas, self.instances)

    def trace(self, rays, any_hit=False, watertight=False, **kw):
        return self.scene.trace(rays, any_hit=any_hit, wide=self.wide, watertight=watertight)

    def world_bound(self):
        return self.scene.root()


class WarpsimEngine(HostsimEngine):
    """The shipped traversal kernel (k_trace_wide, rc_trace_fast.cuh) compiled for the CPU: 32 fibres per warp, lock step at the warp
    intrinsics (tests/hostsim/warpsim.h).  Same scene construction as HostsimEngine."""

    def __init__(self, pushes, n_warps=3):
        super().__init__(pushes, wide=True)
        self.name = "warpsim"
        self.n_warps = n_warps

    def trace(self, rays, any_hit=False, **kw):
        return self.scene.trace_warpsim(rays, any_hit=any_hit, n_warps=self.n_warps)


class GpuEngine:
    def __init__(self, pushes, reference_order=False):
        import raycore_b200 as rc

        self.name = "cuda-ref" if reference_order else "cuda-wide"
        self.reference_order = reference_order
        self.tlas = rc.TLAS(keep_bvh2=True)  # the parity tests read the BVH2 back and use the reference-order mode
        self.handles = []
        for p in pushes:
            verts, fm, xf, ids = _norm_push(p)
            self.handles.append(self.tlas.push(verts, list(xf), instance_ids=ids, face_meta=fm))
        self.tlas.sync()

    def trace(self, rays, any_hit=False, watertight=False, **kw):
        st = self.tlas.adapt()
        fn = st.trace_any if any_hit else st.trace_closest
        return fn(rays, reference_order=self.reference_order, watertight=watertight)

    def world_bound(self):
        b = self.tlas.world_bound()
        return np.concatenate([b.p_min, b.p_max])
