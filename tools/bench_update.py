#!/usr/bin/env python
"""Per-frame mesh (vertex) update: rc_update_geometry as a rebuild and as a refit (CUDA-event ms of the build / refit), for a 10 k- and a
32 k-triangle mesh (one-kernel paths) and a 1 M-triangle mesh (multi-launch paths); sync! after the update (host ms).  One JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import raycore_b200 as rc  # noqa: E402
from raycore_b200 import workloads as W  # noqa: E402

L = rc._lib


def main():
    out = {"lib": os.environ.get("RAYCORE_CUDA_LIB", "default")}
    for tess, label in ((72, "10k"), (128, "32k"), (709, "1M")):
        base_v = W.bumpy_sphere(tess)
        tl = rc.TLAS(allow_refit=True)
        lib, ctx = tl._lib, tl._ctx
        h = tl.push(base_v)
        tl.sync()
        fr = {"rebuild_ms": [], "refit_ms": [], "sync_after_ms": []}
        for f in range(1, 7):
            moved = (base_v.reshape(-1, 3) * (1.0 + 0.02 * f * np.sin(7.0 * base_v.reshape(-1, 3)[:, :1] + f))).astype(np.float32).reshape(-1, 9)
            d_v = torch.from_numpy(moved).cuda()
            for mode, key in ((0, "rebuild_ms"), (L.RC_UPDATE_REFIT, "refit_ms")):
                assert lib.rc_update_geometry(ctx, h.id, d_v.data_ptr(), len(moved), None, L.RC_VERTS_ON_DEVICE | mode) == 0, lib.rc_last_error(ctx)
                assert bool(lib.rc_last_update_refitted(ctx)) == bool(mode)
                fr[key].append(float(lib.rc_last_build_ms(ctx)))
                t0 = time.perf_counter()
                tl.sync()
                fr["sync_after_ms"].append(1e3 * (time.perf_counter() - t0))
            del d_v
        out[label] = {k: float(np.median(v[2:])) for k, v in fr.items()}
        tl.free()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
