import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests'); sys.path.insert(0,'tools')
import numpy as np, torch, ctypes as C
import raycore_b200 as rc
from raycore_b200 import workloads as W
from bench_configs import dev_trace
L = rc._lib
blas = W.bumpy_sphere(72); xf = W.random_trs(10000, 2026, extent=40.0)
tl = rc.TLAS(); h = tl.push(blas, list(xf)); tl.sync()
for n in (1<<22, 4_000_000):
    rays = W.box_rays(n, 99, half=44.0)
    for tmax in (np.inf, 30.0, 5.0):
        rays["t_max"] = tmax
        for any_hit in (False, True):
            tl._lib.rc_get_counters(tl._ctx, (C.c_uint64*6)(), 1)
            mr, hits = dev_trace(tl, rays, any_hit=any_hit, reps=3)
            mr2, _ = dev_trace(tl, rays, any_hit=any_hit, reps=1, flags=L.RC_COUNTERS)
            c = tl.counters()
            print(n, tmax, 'any' if any_hit else 'closest', round(mr,1), 'hit', round(float(hits['hit'].mean()),3), {k: round(c[k]/max(c['rays'],1),2) for k in ('nodes','tri_tests','inst_entries')}, c['max_stack'], flush=True)
