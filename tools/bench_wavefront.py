#!/usr/bin/env python
"""Time the wavefront stages around the trace (docs/src/wavefront-renderer.jl:185-362) on one GPU with
device-resident queues: primary-ray generation, closest_hit, shadow-ray generation, occlusion test, and the fused
shadow-visibility kernel (stages 3+4 without a shadow-ray queue).  Scene: the C2 mesh (bumpy_sphere(709), ~1 M triangles)
seen from outside, `--size`^2 pixels, `--lights` point lights.  Prints one JSON object.
    python tools/bench_wavefront.py [--size 4096] [--lights 2] > gpurun_out/wavefront.json
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import raycore_b200 as rc  # noqa: E402
from raycore_b200 import RAY_DTYPE, HIT_DTYPE  # noqa: E402
from raycore_b200 import workloads as W  # noqa: E402


def best_ms(tl, fn, reps=5):
    ms = []
    for _ in range(reps + 2):
        fn()
        ms.append(tl._lib.rc_last_kernel_ms(tl._ctx))
    return min(ms[2:])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=4096)
    ap.add_argument("--lights", type=int, default=2)
    ap.add_argument("--mesh", type=int, default=709)
    a = ap.parse_args()
    tl = rc.TLAS()
    tl.push(W.bumpy_sphere(a.mesh), None, instance_id=1)
    tl.sync()
    n = a.size * a.size
    lights = np.array([[3, 4, -3], [-4, 1, -2], [0, -5, -1], [2, 2, -6]], np.float32)[: a.lights]
    nl = len(lights)
    rays = tl.queue(RAY_DTYPE, n)
    hits = tl.queue(HIT_DTYPE, n)
    shadow = tl.queue(RAY_DTYPE, n * nl)
    vis = tl.queue(np.uint8, n * nl)
    vis2 = tl.queue(np.uint8, n * nl)
    cam = (0.0, 0.0, -3.0)
    out = {"workload": f"bumpy_sphere({a.mesh}) {tl.sizes()['blas_prims']} triangles, {a.size}x{a.size} primary rays (jittered), {nl} lights",
           "primary_rays": n, "shadow_rays": n * nl}
    t = best_ms(tl, lambda: tl.generate_primary_rays(a.size, a.size, cam, 2.2, 1.0, seed=1, out=rays))
    out["generate_primary_rays"] = {"ms": t, "Mrays_s": n / t / 1e3, "GB_s": n * 32 / t / 1e6}
    t = best_ms(tl, lambda: tl.intersect_rays(rays, out=hits))
    out["closest_hit"] = {"ms": t, "Mrays_s": n / t / 1e3}
    h = hits.download()
    out["primary_hit_fraction"] = float(h["hit"].mean())
    t = best_ms(tl, lambda: tl.generate_shadow_rays(rays, hits, lights, out=shadow))
    out["generate_shadow_rays"] = {"ms": t, "Mrays_s": n * nl / t / 1e3, "GB_s": (n * 64 + n * nl * 32) / t / 1e6}
    t3 = t
    t = best_ms(tl, lambda: tl.test_shadow_rays(shadow, out=vis))
    out["test_shadow_rays"] = {"ms": t, "Mrays_s": n * nl / t / 1e3}
    t4 = t
    t = best_ms(tl, lambda: tl.shadow_visibility(rays, hits, lights, out=vis2))
    out["shadow_visibility_fused"] = {"ms": t, "Mrays_s": n * nl / t / 1e3, "vs_staged": (t3 + t4) / t}
    v, v2 = vis.download(), vis2.download()
    out["fused_equals_staged"] = bool(np.array_equal(v, v2))
    out["visible_fraction_of_hits"] = float(v.sum() / max(1, nl * int(h["hit"].sum())))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
