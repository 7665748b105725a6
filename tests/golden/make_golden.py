#!/usr/bin/env python
"""Generate tests/golden/*.npz: seeded scenes, rays, and the ORACLE's answers (closest / any hit records, BVH2 arrays' digests).

The reference is pure Julia and cannot run in this image, so these vectors are produced by oracle/ (the C restatement that
tests/test_oracle_kat.py pins to the reference's own known answers).  They freeze today's oracle behaviour: the oracle itself, the
host simulation of the device code and the CUDA path are all checked against them (tests/test_golden.py), so a change in any of the
three that alters a single bit of a hit record shows up without the other two.

    python tests/golden/make_golden.py          # rewrites the fixtures (deterministic)
    python tests/golden/make_golden.py --export-inputs DIR   # scenes + rays as <scene>_inputs.npz for tests/golden/make_ref_golden.jl
                                                              # (the real Raycore.jl, where Julia is available -> tests/golden/ref_<scene>.npz)
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from raycore_b200 import workloads as W  # noqa: E402


def _shift(rays, off):
    rays = rays.copy()
    rays["o"] += np.asarray(off, np.float32)
    return rays


def scenes():
    """name -> (pushes, rays): the recipe is code, so the fixture only stores rays and answers."""
    out = {}
    out["single_sphere"] = ([(W.uv_sphere(20, (0, 0, 2), 1.0), None, W.identity3x4()[None], np.array([1], np.uint32))],
                            np.concatenate([W.pinhole_rays(40, 40, camera_pos=(0, 0, 0)), _shift(W.interior_rays(1500, 3, radius=0.5), (0, 0, 2))]))
    xf = W.random_trs(12, seed=5, extent=4.0, smin=0.4, smax=1.6)
    out["instanced_bumpy"] = ([(W.bumpy_sphere(16), None, xf, np.arange(100, 112, dtype=np.uint32))], W.box_rays(3000, seed=6, half=6.0))
    box = np.concatenate([W.box_mesh()[:5], np.zeros((1, 9), np.float32), W.box_mesh()[5:]])  # a degenerate face inside the soup
    out["two_blas_windows"] = ([(box, None, W.random_trs(5, seed=7, extent=3.0), None), (W.quad_mesh(0.0, 2.0), np.array([11, 12], np.uint32), W.random_trs(3, seed=8, extent=3.0), None)],
                               None)
    r = W.box_rays(3000, seed=9, half=5.0)
    rs = np.random.RandomState(10)
    win = rs.rand(len(r)) < 0.5
    r["t_min"][win] = rs.uniform(0, 2, win.sum()).astype(np.float32)
    r["t_max"][win] = r["t_min"][win] + rs.uniform(0, 5, win.sum()).astype(np.float32)
    out["two_blas_windows"] = (out["two_blas_windows"][0], r)
    return out


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def export_inputs(outdir):
    os.makedirs(outdir, exist_ok=True)
    for name, (pushes, rays) in scenes().items():
        d = {"n_pushes": np.array([len(pushes)], np.int64), "rays": np.ascontiguousarray(rays).view(np.float32).reshape(len(rays), 8)}
        for k, (verts, meta, xf, ids) in enumerate(pushes):
            d[f"verts_{k}"] = np.ascontiguousarray(verts, np.float32).reshape(-1, 9)
            d[f"xf_{k}"] = np.ascontiguousarray(xf, np.float32).reshape(-1, 12)
            if meta is not None:
                d[f"meta_{k}"] = np.ascontiguousarray(meta, np.uint32)
            if ids is not None:
                d[f"ids_{k}"] = np.ascontiguousarray(ids, np.uint32)
        np.savez(os.path.join(outdir, name + "_inputs.npz"), **d)
        print("wrote", name + "_inputs.npz")


def main():
    if "--export-inputs" in sys.argv:
        return export_inputs(sys.argv[sys.argv.index("--export-inputs") + 1])
    import engines

    for name, (pushes, rays) in scenes().items():
        o = engines.OracleEngine(pushes)
        closest, anyh = o.trace(rays), o.trace(rays, any_hit=True)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), rays=rays, closest=closest, any=anyh,
                            tlas_nodes_sha256=digest(o.tlas.nodes), blas_nodes_sha256=digest(o.tlas.all_blas_nodes),
                            n_tlas_nodes=len(o.tlas.nodes), n_blas_nodes=len(o.tlas.all_blas_nodes))
        print(name, len(rays), "rays, hit rate %.3f" % closest["hit"].mean())


if __name__ == "__main__":
    main()
