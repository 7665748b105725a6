#!/usr/bin/env python
"""(test infrastructure: compares against oracle/) Measure every BASELINE.json config (C1..C5) on one GPU and print one JSON object.
Complements bench.py (which carries the headline C2 metric and the driver contract).  Usage:
    python tests/bench_configs.py [--quick] > gpurun_out/configs.json
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import raycore_b200 as rc  # noqa: E402
from raycore_b200 import workloads as W  # noqa: E402
from oracle import oracle as orc  # noqa: E402
import engines  # noqa: E402
import parity  # noqa: E402

L = rc._lib


def dev_trace(tlas, rays, any_hit=False, reps=5, flags=0):
    """device-resident timing through the C ABI; returns (Mrays/s best, hits)"""
    lib, ctx = tlas._lib, tlas._ctx
    n = len(rays)
    d_r = torch.from_numpy(rays.view(np.uint8).reshape(-1)).cuda()
    d_h = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
    fn = lib.rc_trace_any if any_hit else lib.rc_trace_closest
    ms = []
    for _ in range(reps + 2):
        assert fn(ctx, d_r.data_ptr(), d_h.data_ptr(), n, L.RC_RAYS_ON_DEVICE | L.RC_HITS_ON_DEVICE | flags) == 0, lib.rc_last_error(ctx)
        ms.append(lib.rc_last_kernel_ms(ctx))
    return n / (min(ms[2:]) * 1e-3) / 1e6, d_h.cpu().numpy().view(L.HIT_DTYPE)


def parity_summary(a, b, rays, o):
    cls = parity.classify(a, b, parity.make_graze_verifier(orc, rays, a, o.instances, o.tris))
    return parity.summarize(cls, len(rays))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    out = {"device": torch.cuda.get_device_name(0), "cpu_threads": orc.max_threads()}
    NR = 1 << (22 if args.quick else 24)

    # ---- C1: README sphere, 1024^2 primary rays -------------------------------------------------------------------
    verts = W.uv_sphere(24, (0, 0, 2), 1.0)
    pushes = [(verts, None, [W.identity3x4()], [1])]
    g, o = engines.GpuEngine(pushes), engines.OracleEngine(pushes)
    rays = W.pinhole_rays(1024, 1024, camera_pos=(0, 0, 0))
    mr, a = dev_trace(g.tlas, rays)
    t0 = time.time(); b = o.trace(rays); cpu = len(rays) / (time.time() - t0) / 1e6
    out["C1_readme_sphere"] = {"triangles": g.tlas.sizes()["blas_prims"], "rays": len(rays), "gpu_Mrays_s": mr, "cpu_oracle_Mrays_s": cpu, "parity": parity_summary(a, b, rays, o)}
    g.tlas.free()

    # ---- C2 extras: any_hit and reference-order mode on the 1M-triangle mesh --------------------------------------------
    verts = W.bumpy_sphere(709)
    tl = rc.TLAS(keep_bvh2=True)
    tl.push(verts, None, instance_id=1)
    tl.sync()
    rays = W.interior_rays(NR, 77, radius=0.8)
    c2 = {"triangles": tl.sizes()["blas_prims"], "rays": NR}
    c2["closest_interior_Mrays_s"], hc = dev_trace(tl, rays)
    c2["any_interior_Mrays_s"], ha = dev_trace(tl, rays, any_hit=True)
    c2["closest_reference_order_Mrays_s"], hr = dev_trace(tl, rays[: NR // 4], flags=L.RC_MODE_REFERENCE_ORDER)
    prim = W.pinhole_rays(4096 if not args.quick else 2048, 4096 if not args.quick else 2048, camera_pos=(0, 0, -3))
    c2["closest_primary_Mrays_s"], hp = dev_trace(tl, prim)
    c2["any_equals_closest_hit_flag"] = bool(np.array_equal(ha["hit"], hc["hit"]))
    c2["wide_vs_reference_order_ids_equal_frac"] = float(np.mean((hc["primitive_id"][: NR // 4] == hr["primitive_id"]) & (hc["hit"][: NR // 4] == hr["hit"])))
    lib, ctx = tl._lib, tl._ctx
    d_verts = torch.from_numpy(verts).cuda()
    hh, dd, xf = C.c_uint32(), C.c_int32(), W.identity3x4()
    bms = []
    for _ in range(8):
        assert lib.rc_push(ctx, d_verts.data_ptr(), len(verts), None, xf.ctypes.data, None, None, 1, L.RC_VERTS_ON_DEVICE, C.byref(hh)) == 0
        bms.append(float(lib.rc_last_build_ms(ctx)))
        lib.rc_delete(ctx, hh.value, C.byref(dd))
        tl.sync()  # steady state: the freed BLAS blocks go back to the pool before the next build
    c2["blas_build_ms_cuda_events"] = min(bms)
    for n_t, label in ((355, "250k"), (1416, "4M")):
        v2 = torch.from_numpy(W.bumpy_sphere(n_t)).cuda()
        b2 = []
        for _ in range(4):
            assert lib.rc_push(ctx, v2.data_ptr(), v2.shape[0], None, xf.ctypes.data, None, None, 1, L.RC_VERTS_ON_DEVICE, C.byref(hh)) == 0
            b2.append(float(lib.rc_last_build_ms(ctx)))
            lib.rc_delete(ctx, hh.value, C.byref(dd))
            tl.sync()
        c2[f"blas_build_ms_{label}"] = {"faces": int(v2.shape[0]), "ms": min(b2)}
        del v2
    out["C2_1M_triangles"] = c2
    tl.free()

    # ---- C3: 10k instances of a 10k-triangle BLAS ----------------------------------------------------------------------
    blas = W.bumpy_sphere(72)
    xf = W.random_trs(10000, 2026, extent=40.0)
    tl = rc.TLAS()
    t0 = time.time()
    h = tl.push(blas, list(xf))
    tl.sync()
    c3 = {"instances": 10000, "blas_triangles": tl.sizes()["blas_prims"], "push_sync_ms": 1e3 * (time.time() - t0), "rays": NR}
    rays = W.box_rays(NR, 7, half=44.0)
    c3["closest_Mrays_s"], hc = dev_trace(tl, rays)
    c3["any_Mrays_s"], _ = dev_trace(tl, rays, any_hit=True)
    c3["hit_rate"] = float(hc["hit"].mean())
    lib, ctx = tl._lib, tl._ctx
    lib.rc_get_counters(ctx, (C.c_uint64 * 6)(), 1)
    m = 1 << 20
    tl.adapt().trace_closest(rays[:m], counters=True)
    cn = tl.counters()
    c3["per_ray"] = {k: cn[k] / m for k in ("nodes", "box_tests", "tri_tests", "inst_entries")} | {"max_stack": cn["max_stack"]}
    o = engines.OracleEngine([(blas, None, xf, None)])
    ns = 1 << 20
    t0 = time.time(); b = o.trace(rays[:ns]); c3["cpu_oracle_Mrays_s"] = ns / (time.time() - t0) / 1e6
    c3["parity"] = parity_summary(hc[:ns], b, rays[:ns], o)
    # C5 (ii): refit frames — re-randomise all transforms, update_transforms! + sync! (refit), then any_hit shadow rays
    refit_ms = []
    rs = np.random.RandomState(5)
    for f in range(5):
        xf2 = xf.copy()  # small per-frame motion (refit keeps the topology, so it is meant for coherent motion)
        xf2[:, [3, 7, 11]] += rs.uniform(-0.5, 0.5, (len(xf), 3)).astype(np.float32) * (f + 1)
        t0 = time.time()
        tl.update_transforms(h, list(xf2))
        t1 = time.time()
        tl.sync()
        refit_ms.append({"host_pack_ms": 1e3 * (t1 - t0), "sync_refit_ms": 1e3 * (time.time() - t1)})
        assert tl.last_sync_action == rc.RC_SYNC_REFIT
    c3["refit_frames"] = refit_ms
    shadow = W.box_rays(10_000_000 if not args.quick else 1_000_000, 99, half=44.0)
    shadow["t_max"] = 30.0
    c3["any_hit_shadow_10M_Mrays_s"], _ = dev_trace(tl, shadow, any_hit=True)
    out["C3_instanced_and_C5_refit"] = c3
    tl.free()

    # ---- C5 (iii): per-frame mesh (vertex) update of the test_mesh_update workload: update!(tlas, handle, mesh) as a rebuild and as a
    # refit of the kept radix tree, then sync! and 10 M any_hit shadow rays (device-resident vertices: rc_update_geometry, CUDA events) ---
    c5 = {}
    for tess, label in ((72, "10k"), (709, "1M")):
        base_v = W.bumpy_sphere(tess)
        tl = rc.TLAS(allow_refit=True)
        lib, ctx = tl._lib, tl._ctx
        h = tl.push(base_v, list(W.random_trs(64 if tess == 72 else 1, 3, extent=6.0)))
        tl.sync()
        frames = {"rebuild_ms": [], "refit_ms": [], "sync_after_ms": []}
        for f in range(1, 5):
            moved = (base_v.reshape(-1, 3) * (1.0 + 0.02 * f * np.sin(7.0 * base_v.reshape(-1, 3)[:, :1] + f))).astype(np.float32).reshape(-1, 9)
            d_v = torch.from_numpy(moved).cuda()
            for mode, key in ((0, "rebuild_ms"), (L.RC_UPDATE_REFIT, "refit_ms")):
                assert lib.rc_update_geometry(ctx, h.id, d_v.data_ptr(), len(moved), None, L.RC_VERTS_ON_DEVICE | mode) == 0, lib.rc_last_error(ctx)
                assert bool(lib.rc_last_update_refitted(ctx)) == bool(mode)
                frames[key].append(float(lib.rc_last_build_ms(ctx)))
                t0 = time.time()
                tl.sync()
                frames["sync_after_ms"].append(1e3 * (time.time() - t0))
            del d_v
        sh = W.box_rays(10_000_000 if not args.quick else 1_000_000, 98, half=8.0 if tess == 72 else 1.5)
        sh["t_max"] = 6.0
        c5[label] = {"triangles": tl.sizes()["blas_prims"], "update_rebuild_ms": min(frames["rebuild_ms"]), "update_refit_ms": min(frames["refit_ms"]),
                     "sync_after_update_ms": min(frames["sync_after_ms"]), "any_hit_shadow_10M_Mrays_s": dev_trace(tl, sh, any_hit=True)[0]}
        tl.free()
    out["C5_mesh_update"] = c5

    # ---- C4: view factors, 5 x bumpy_sphere(72), 1000 rays per triangle -----------------------------------------------------
    meshes = W.viewfactor_scene(72)
    tl = rc.TLAS()
    base = 0
    for msh in meshes:
        keep = ~W.is_degenerate(msh)
        meta = np.zeros(len(msh), np.uint32)
        meta[keep] = base + 1 + np.arange(keep.sum())
        base += int(keep.sum())
        tl.push(msh, None, face_meta=meta)
    tl.sync()
    n_prims = tl.sizes()["blas_prims"]
    rpt = 1000 if not args.quick else 100
    lib, ctx = tl._lib, tl._ctx
    d_out = torch.empty(n_prims * n_prims, dtype=torch.int32, device="cuda")
    sk = C.c_uint64()
    ms = []
    for _ in range(3):
        t0 = time.time()
        assert lib.rc_view_factors(ctx, rpt, 11, d_out.data_ptr(), 0, n_prims, L.RC_HITS_ON_DEVICE, C.byref(sk)) == 0
        ms.append((lib.rc_last_kernel_ms(ctx), 1e3 * (time.time() - t0)))
    vf = d_out.view(n_prims, n_prims)
    tot = int(vf.sum().item())
    c4 = {"triangles": n_prims, "rays_per_triangle": rpt, "rays": n_prims * rpt, "kernel_s": min(m[0] for m in ms) * 1e-3, "call_s_device_output": min(m[1] for m in ms) * 1e-3,
          "Mrays_s": n_prims * rpt / (min(m[0] for m in ms) * 1e-3) / 1e6, "total_hits": tot, "diag_zero": bool((torch.diagonal(vf) == 0).all().item()),
          "max_row_sum": int(vf.sum(1).max().item()), "skipped": int(sk.value), "matrix_bytes": int(n_prims) ** 2 * 4}
    t0 = time.time()
    host = tl.view_factors(rpt, seed=11, row_base=0, n_rows=min(n_prims, 4096))
    c4["call_s_4096_rows_host_output"] = time.time() - t0
    # oracle on a 64-row sample (same RNG spec): statistical agreement of the row totals
    o = None
    out["C4_view_factors"] = c4
    tl.free()

    # ---- C5 (i): rebuild frames: delete! + push! + sync! with the mesh-update tessellation schedule (x8) ----------------------
    tl = rc.TLAS()
    h = tl.push(W.bumpy_sphere(64))
    tl.sync()
    frames = []
    lib, ctx = tl._lib, tl._ctx
    for n_t in [256, 64, 384, 96, 512, 128, 64, 256, 768, 128]:
        mesh = W.bumpy_sphere(n_t)
        t0 = time.time()
        tl.delete(h)
        h = tl.push(mesh)
        tl.sync()
        frames.append({"tess": n_t, "faces": len(mesh), "delete_push_sync_ms": 1e3 * (time.time() - t0), "blas_build_ms_cuda_events": float(lib.rc_last_build_ms(ctx))})
    out["C5_rebuild_frames"] = frames
    tl.free()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
