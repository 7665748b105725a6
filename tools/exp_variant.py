#!/usr/bin/env python
"""Kernel-variant experiment driver: times closest_hit / any_hit on the C3 (instanced) and C2 (single mesh) scenes with the
library named by RAYCORE_CUDA_LIB and prints one JSON line with a checksum of the hit records, so variants built with different
-D flags can be compared for speed and for bit-identical results.  Usage: RAYCORE_CUDA_LIB=build/variants/x/lib.so python tools/exp_variant.py [--c2]"""
import json
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import raycore_b200 as rc  # noqa: E402
from raycore_b200 import workloads as W  # noqa: E402

L = rc._lib


def dev_trace(tlas, d_r, n, any_hit=False, reps=7, extra_flags=0):
    lib, ctx = tlas._lib, tlas._ctx
    d_h = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
    fn = lib.rc_trace_any if any_hit else lib.rc_trace_closest
    ms = []
    for _ in range(reps + 2):
        assert fn(ctx, d_r.data_ptr(), d_h.data_ptr(), n, L.RC_RAYS_ON_DEVICE | L.RC_HITS_ON_DEVICE | extra_flags) == 0, lib.rc_last_error(ctx)
        ms.append(lib.rc_last_kernel_ms(ctx))
    return sorted(ms[2:])[len(ms[2:]) // 2], zlib.crc32(d_h.cpu().numpy().tobytes())


def main():
    out = {"lib": os.environ.get("RAYCORE_CUDA_LIB", "default")}
    n = 1 << (int(sys.argv[sys.argv.index('--log2rays') + 1]) if '--log2rays' in sys.argv else 24)
    if "--vf" in sys.argv:  # C4: view_factors of 5 bumpy spheres x 1000 rays per triangle (bench.py extras)
        import ctypes as C

        vt, base = rc.TLAS(), 0
        for msh in W.viewfactor_scene(72):
            keep = ~W.is_degenerate(msh)
            meta = np.zeros(len(msh), np.uint32)
            meta[keep] = base + 1 + np.arange(keep.sum())
            base += int(keep.sum())
            vt.push(msh, None, face_meta=meta)
        vt.sync()
        npr = vt.sizes()["blas_prims"]
        d_vf = torch.empty(npr * npr, dtype=torch.int32, device="cuda")
        sk, ms = C.c_uint64(), []
        for _ in range(4):
            assert vt._lib.rc_view_factors_strided(vt._ctx, 1000, 11, d_vf.data_ptr(), 0, 1, npr, L.RC_HITS_ON_DEVICE, C.byref(sk)) == 0
            ms.append(float(vt._lib.rc_last_kernel_ms(vt._ctx)))
        out.update(scene="C4 view_factors", vf_ms=min(ms[1:]), total_hits=int(d_vf.sum().item()))
        print(json.dumps(out))
        return
    if "--c2" in sys.argv:
        tl = rc.TLAS()
        tl.push(W.bumpy_sphere(709))
        tl.sync()
        rays = W.interior_rays(n, 11, radius=0.8)
        out["scene"] = "C2 interior"
    else:
        tl = rc.TLAS()
        tl.push(W.bumpy_sphere(72), list(W.random_trs(10000, 2026, extent=40.0)))
        tl.sync()
        rays = W.box_rays(n, 7, half=44.0)
        out["scene"] = "C3"
    d_r = torch.from_numpy(rays.view(np.uint8).reshape(-1)).cuda()
    out["closest_ms"], out["closest_crc"] = dev_trace(tl, d_r, n)
    out["any_ms"], out["any_crc"] = dev_trace(tl, d_r, n, any_hit=True)
    out["closest_Mrays_s"] = n / out["closest_ms"] / 1e3
    if "--watertight" in sys.argv:  # the same launches with RC_MODE_WATERTIGHT (hit / miss can differ at edges: no CRC comparison with the default)
        out["wt_closest_ms"], _ = dev_trace(tl, d_r, n, extra_flags=L.RC_MODE_WATERTIGHT)
        out["wt_any_ms"], _ = dev_trace(tl, d_r, n, any_hit=True, extra_flags=L.RC_MODE_WATERTIGHT)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
