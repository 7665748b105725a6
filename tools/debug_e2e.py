#!/usr/bin/env python
"""Debug: host-buffer trace vs device-resident trace on the C3 scene, byte comparison, for several ray counts."""
import ctypes as C, sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import raycore_b200 as rc
from raycore_b200 import workloads as W
import bench
L = rc._lib
tl = rc.TLAS()
tl.push(W.bumpy_sphere(72), list(W.random_trs(10000, 2026, extent=40.0)))
tl.sync()
lib, ctx = tl._lib, tl._ctx
for n in (1 << 20, 3 * (1 << 20) + 12345, 1 << 25, 100_000_000):
    h_rays = torch.empty(n * 32, dtype=torch.uint8).pin_memory()
    bench.gen_box_rays(h_rays.numpy().view(W.RAY_DTYPE), 0, 16)
    d_rays = h_rays.cuda()
    d_hits = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
    h_hits = torch.empty(n * 32, dtype=torch.uint8).pin_memory()
    assert lib.rc_trace_closest(ctx, d_rays.data_ptr(), d_hits.data_ptr(), n, L.RC_RAYS_ON_DEVICE | L.RC_HITS_ON_DEVICE) == 0
    a = d_hits.cpu().numpy().view(L.HIT_DTYPE).copy()
    assert lib.rc_trace_closest(ctx, d_rays.data_ptr(), d_hits.data_ptr(), n, L.RC_RAYS_ON_DEVICE | L.RC_HITS_ON_DEVICE) == 0
    a2 = d_hits.cpu().numpy().view(L.HIT_DTYPE)
    for rep in range(2):
        assert lib.rc_trace_closest(ctx, h_rays.data_ptr(), h_hits.data_ptr(), n, 0) == 0
        b = h_hits.numpy().view(L.HIT_DTYPE)
        bad = np.nonzero(a.view(np.uint8).reshape(-1, 32) != b.view(np.uint8).reshape(-1, 32))[0]
        bad = np.unique(bad)
        print(n, "rep", rep, "dev-vs-dev equal:", a.tobytes() == a2.tobytes(), "host-vs-dev mismatching rays:", len(bad), "hits dev", int(a["hit"].sum()), "host", int(b["hit"].sum()))
        if len(bad):
            print("  first bad idx:", bad[:10], "chunk:", bad[:10] >> 20, "max hit value host", b["hit"].max(), "dev", a["hit"].max())
            print("  dev ", a[bad[:3]])
            print("  host", b[bad[:3]])
    del h_rays, d_rays, d_hits, h_hits
