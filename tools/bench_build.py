#!/usr/bin/env python
"""BLAS build time (CUDA events around rc_push's build, vertices resident, pooled blocks reused): min / median of several builds for the
mesh sizes of profiles/README.md; one JSON line.  `--once TESS` builds a single mesh once more after a warm-up (for an ncu launch list)."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import raycore_b200 as rc  # noqa: E402
from raycore_b200 import workloads as W  # noqa: E402

L = rc._lib


def main():
    tl = rc.TLAS()
    lib, ctx = tl._lib, tl._ctx
    xf = W.identity3x4()
    hh, dd = C.c_uint32(), C.c_int32()
    out = {"lib": os.environ.get("RAYCORE_CUDA_LIB", "default")}
    sizes = ((65, "8k"), (96, "18k"), (128, "32k"), (160, "50k"), (355, "250k"), (709, "1M"), (1418, "4M"))
    if "--once" in sys.argv:
        sizes = ((int(sys.argv[sys.argv.index("--once") + 1]), "once"),)
    flags = L.RC_VERTS_ON_DEVICE | (L.RC_BUILD_KEEP_BVH2 if "--bvh2" in sys.argv else 0)
    for tess, label in sizes:
        v = torch.from_numpy(W.bumpy_sphere(tess)).cuda()
        ms = []
        for _ in range(3 if label == "once" else 9):
            assert lib.rc_push(ctx, v.data_ptr(), v.shape[0], None, xf.ctypes.data, None, None, 1, flags, C.byref(hh)) == 0, lib.rc_last_error(ctx)
            ms.append(float(lib.rc_last_build_ms(ctx)))
            lib.rc_delete(ctx, hh.value, C.byref(dd))
            tl.sync()
        ms = sorted(ms[1:])
        out[label] = {"faces": int(v.shape[0]), "min_ms": ms[0], "median_ms": ms[len(ms) // 2]}
        del v
    print(json.dumps(out))


if __name__ == "__main__":
    main()
