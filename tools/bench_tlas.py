#!/usr/bin/env python
"""TLAS build / refit wall-clock per frame (host time of sync!, which ends with the root-box read-back): the C5 frames of tests/bench_configs.py
in isolation.  One JSON line: for each instance count the median of the refit frames (update_transforms! + sync!) and of full rebuilds
(delete! + push! of a new handle + sync!), plus sync! after a one-mesh update (one-instance TLAS)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raycore_b200 as rc  # noqa: E402
from raycore_b200 import workloads as W  # noqa: E402


def main():
    out = {"lib": os.environ.get("RAYCORE_CUDA_LIB", "default")}
    mesh = W.bumpy_sphere(12)
    for n in (1, 1000, 10000, 32768, 40000):
        xf = W.random_trs(n, 2026, extent=40.0)
        tl = rc.TLAS()
        h = tl.push(mesh, list(xf))
        tl.sync()
        rs = np.random.RandomState(5)
        refit, rebuild = [], []
        for f in range(9):
            xf2 = xf.copy()
            xf2[:, [3, 7, 11]] += rs.uniform(-0.5, 0.5, (n, 3)).astype(np.float32)
            tl.update_transforms(h, list(xf2))
            t0 = time.perf_counter()
            tl.sync()
            refit.append(1e3 * (time.perf_counter() - t0))
            assert tl.last_sync_action == rc.RC_SYNC_REFIT
        for f in range(7):  # same count, new handle: a rebuild that can keep the device arrays
            tl.delete(h)
            h = tl.push(mesh, list(xf))
            t0 = time.perf_counter()
            tl.sync()
            rebuild.append(1e3 * (time.perf_counter() - t0))
        out[str(n)] = {"refit_sync_ms": float(np.median(refit[1:])), "rebuild_sync_ms": float(np.median(rebuild[1:]))}
        tl.free()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
