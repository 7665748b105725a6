#!/usr/bin/env python
"""bench.py — closest_hit Mrays/s (incoherent) on the 1M-triangle config of BASELINE.json (configs[1]).

    python bench.py --gpus 1 --steps K --warmup W              our arm (CUDA library through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...     the reference's CPU algorithm (oracle port, OpenMP)

A step = one pass of batched closest_hit over this rank's ray batch (2^24 rays: 2^23 diffuse-bounce rays leaving
the surface + 2^23 rays from interior points, uniform directions — both incoherent).  `value` is timed with rays
and hit buffers resident in HBM; `e2e` is the same call with pinned HOST ray/hit buffers (H2D + D2H inside the
timed region).  One process per GPU; the BVH is replicated, rays are sharded (weak scaling: every rank traces its own
batch); each step ends with the NCCL gather of the hit records to rank 0 when N > 1.
Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

TESS = 709            # bumpy_sphere(709): 1,002,528 faces
RAYS_PER_RANK = 1 << 24
PRIMARY_RES = 3072    # primary rays used to seed the bounce rays
CPU_SAMPLE = 1 << 21  # rays per CPU-baseline step (bounded sample of the same ray set)


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), float(p.get("sm_max_mhz", 1965.0)), "measured"
    except Exception:
        return 6650.0, 1965.0, "fallback"


def measure_l2_gbs(torch, dev):
    """L2-resident device copy (24 MiB -> 24 MiB, both inside the 126 MB L2): read + write bytes per second, the ceiling the
    BVH fetches are compared with."""
    a = torch.empty(24 << 20, dtype=torch.uint8, device=dev)
    b = torch.empty_like(a)
    for _ in range(5):
        b.copy_(a)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    iters = 40
    for _ in range(iters):
        b.copy_(a)
    e1.record()
    torch.cuda.synchronize()
    return 2.0 * a.numel() * iters / (e0.elapsed_time(e1) * 1e-3) / 1e9


def measure_instanced(rc, W, L, torch, dev, local):
    """C3 (BASELINE configs[2], on one GPU): 10,000 instances of bumpy_sphere(72) under random T*R*S transforms, 2^23 incoherent rays through
    the scene box, device-resident closest_hit and any_hit; best of 5 launches after one warm-up (the library's CUDA-event kernel time)."""
    tl = rc.TLAS(local)
    tl.push(W.bumpy_sphere(72), list(W.random_trs(10000, 2026, extent=40.0)))
    tl.sync()
    n = 1 << 23
    rays = W.box_rays(n, 7, half=44.0)
    d_r = torch.from_numpy(rays.view(np.uint8).reshape(-1)).to(dev)
    d_h = torch.empty(n * 32, dtype=torch.uint8, device=dev)
    out = {"instanced_config": f"C3: 10000 instances x {tl.sizes()['blas_prims']} triangles, {n} rays with origins uniform in the scene box and uniform directions, rays and hits resident in HBM"}
    for name, fn in (("closest", tl._lib.rc_trace_closest), ("any", tl._lib.rc_trace_any)):
        ms = []
        for _ in range(6):
            assert fn(tl._ctx, d_r.data_ptr(), d_h.data_ptr(), n, L.RC_RAYS_ON_DEVICE | L.RC_HITS_ON_DEVICE) == 0, tl._lib.rc_last_error(tl._ctx)
            ms.append(float(tl._lib.rc_last_kernel_ms(tl._ctx)))
        out[f"instanced_{name}_hit_Mrays_s"] = n / (min(ms[1:]) * 1e-3) / 1e6
    out["instanced_hit_rate"] = float((d_h.view(torch.int32)[::8] == 1).float().mean().item())
    del d_r, d_h
    tl.free()
    return out


def measure_view_factors(rc, W, L, torch, dev, local, world, rank, dist):
    """C4: view_factors of 5 bumpy spheres (49,704 triangles) x 1000 rays per triangle into a UInt32 matrix resident in HBM.
    N > 1: every rank holds the (replicated) scene and computes its own interleaved share of the source rows, no exchange;
    the time is the max over ranks of the library's CUDA-event kernel time."""
    vt = rc.TLAS(local)
    base = 0
    for msh in W.viewfactor_scene(72):
        keep = ~W.is_degenerate(msh)
        meta = np.zeros(len(msh), np.uint32)
        meta[keep] = base + 1 + np.arange(keep.sum())
        base += int(keep.sum())
        vt.push(msh, None, face_meta=meta)
    vt.sync()
    npr = vt.sizes()["blas_prims"]
    n_mine = len(range(rank, npr, world))  # interleaved share: rows rank, rank + world, ... (equally expensive shares)
    d_vf = torch.empty(n_mine * npr, dtype=torch.int32, device=dev)
    sk = C.c_uint64()
    vms = []
    for _ in range(3):
        if dist is not None:
            dist.barrier()
        assert vt._lib.rc_view_factors_strided(vt._ctx, 1000, 11, d_vf.data_ptr(), rank, world, n_mine, L.RC_HITS_ON_DEVICE, C.byref(sk)) == 0
        vms.append(float(vt._lib.rc_last_kernel_ms(vt._ctx)))
    ms, hits = min(vms), int(d_vf.sum().item())
    if dist is not None:
        t = torch.tensor([ms, float(hits)], device=dev, dtype=torch.float64)
        mx = t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms, hits = float(mx[0].item()), int(t[1].item())
    del d_vf
    vt.free()
    return {"view_factors_s": ms * 1e-3,
            "view_factors_config": f"C4: 5 x bumpy_sphere(72) = {npr} triangles, rays_per_triangle=1000 ({npr * 1000} rays), UInt32 {npr}x{npr} matrix resident in HBM, "
                                   f"{world} GPU(s): source rows interleaved over the ranks, no exchange",
            "view_factors_total_hits": hits}


def ncu_traffic(rays_per_launch):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the committed ncu --set full
    capture of this same command (profiles/r1_traffic.json, written by tools/ncu_summary.py); None if the capture is for
    another launch size."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        return float(t["dram_bytes"]) if int(t["rays_per_launch"]) == int(rays_per_launch) else None
    except Exception:
        return None


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe): one `nvidia-smi -lms 20`
    process is read continuously; only samples stamped between mark_start() and mark_end() are summarised."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.t0, self.t1, self.proc = index, [], None, None, None
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append((time.time(), [x.strip() for x in line.strip().split(",")]))
        except Exception:
            pass

    def __enter__(self):
        self.t.start()
        deadline = time.time() + 10.0  # nvidia-smi needs a second or more to initialise on an 8-GPU box: wait for its first sample
        while not self.rows and time.time() < deadline:
            time.sleep(0.02)
        return self

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def __exit__(self, *a):
        time.sleep(0.05)
        if self.proc is not None:
            self.proc.terminate()
        self.t.join(timeout=3)

    def summary(self):
        inside = [r for ts, r in self.rows if len(r) >= 7 and self.t0 is not None and self.t0 - 0.02 <= ts <= (self.t1 or ts) + 0.02]
        rows = inside if inside else [r for _, r in self.rows if len(r) >= 7]
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "power_w_max": max(pw) if pw else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_in_timed_region": len(inside)}


def build_rays(tlas_trace, verts, faces_of_prim, n, seed):
    """2^23 diffuse-bounce rays from primary hits + 2^23 interior rays, interleaved in blocks so both kinds are in every chunk."""
    from raycore_b200 import workloads as W

    half = n // 2
    prim = W.pinhole_rays(PRIMARY_RES, PRIMARY_RES, camera_pos=(0.0, 0.0, -3.0))
    ph = tlas_trace(prim)
    normals = W.geometric_normals(verts)
    nrm = normals[faces_of_prim[ph["primitive_id"]]]
    b = W.bounce_rays(half, prim, ph, nrm, seed=0x5EED + seed)
    c = W.interior_rays(n - half, seed=77 + seed, radius=0.8)
    rays = np.empty(n, W.RAY_DTYPE)
    rays[0::2] = b
    rays[1::2] = c
    return rays, float(ph["hit"].mean())


def run_reference(args):
    """The reference's CPU algorithm (C restatement, OpenMP over rays as Threads.@threads does) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    from raycore_b200 import workloads as W

    verts = W.bumpy_sphere(TESS)
    t0 = time.time()
    blas = orc.OracleBLAS.from_verts(verts)
    inst = orc.make_instances(1, [orc.identity3x4()], [1])
    tlas = orc.OracleTLAS([blas], inst)
    build_s = time.time() - t0
    cores = host_threads()  # all the cores this process may run on (torchrun exports OMP_NUM_THREADS=1: ask for them explicitly)
    n = CPU_SAMPLE
    # the sample = the first CPU_SAMPLE rays of the benchmark's own ray set; the oracle traces the primaries itself
    prim = W.pinhole_rays(1024, 1024, camera_pos=(0.0, 0.0, -3.0))
    ph = tlas.closest_hit(prim, threads=cores)
    normals = W.geometric_normals(verts)
    order = blas.prims["input_index"]
    tris_in = orc.filter_triangles(verts)
    face_of_prim = (tris_in["metadata"] - 1).astype(np.int64)  # metadata = 1-based face index
    nrm = normals[face_of_prim[ph["primitive_id"]]]
    rays = np.empty(n, W.RAY_DTYPE)
    rays[0::2] = W.bounce_rays(n // 2, prim, ph, nrm, seed=0x5EED)
    rays[1::2] = W.interior_rays(n - n // 2, seed=77, radius=0.8)
    for _ in range(args.warmup):
        tlas.closest_hit(rays, threads=cores)
    t0 = time.time()
    for _ in range(args.steps):
        tlas.closest_hit(rays, threads=cores)
    dt = time.time() - t0
    v = n * args.steps / dt / 1e6
    line = {
        "impl": "reference", "metric": "closest_hit Mrays/s (incoherent)", "value": v, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": f"C2 bumpy_sphere({TESS}) 1,002,528 faces, 1 instance; bounded sample of {n} rays/step (bounce+interior interleaved)"},
        "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": f"{n} rays/step x {args.steps} steps; oracle/oracle.c (C restatement of the reference BVH2 path; Julia absent)", "build_s": build_s},
        "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--rays", type=int, default=RAYS_PER_RANK)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the view_factors (C4) measurement")
    ap.add_argument("--gather", default="auto", choices=["auto", "fused", "peer-copy", "nccl"], help="N > 1: how hit records reach rank 0")
    ap.add_argument("--counters", action="store_true", help="extra instrumented pass (per-ray node/triangle counts) after the timed region")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import raycore_b200 as rc
    from raycore_b200 import workloads as W

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gather == "auto":
        # measured on the 8 x B200 box (profiles/r1_scaling.md): in-kernel remote stores are free up to 4 ranks; at 8 ranks the 32-byte
        # stores of 7 senders saturate rank 0's NVLink ingress, so the copy engines push whole buffers while the next step traces
        args.gather = "fused" if world <= 4 else "peer-copy"
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.rays
    L = rc._lib

    # ---- scene: every rank builds the same BVH (replicated, deterministic builder) --------------------------
    verts = W.bumpy_sphere(TESS)
    tlas = rc.TLAS(local)
    lib, ctx = tlas._lib, tlas._ctx
    t0 = time.time()
    h = tlas.push(verts, None, instance_id=1)
    tlas.sync()
    build_wall_ms = 1e3 * (time.time() - t0)
    # device-side build time with the vertices already resident (what the reference's published build numbers measure)
    d_verts = torch.from_numpy(verts).to(dev)
    xf = W.identity3x4()
    hh = C.c_uint32()
    torch.cuda.synchronize()
    build_ms, build_dev_ms = [], []
    for _ in range(8):
        t0 = time.time()
        assert lib.rc_push(ctx, d_verts.data_ptr(), len(verts), None, xf.ctypes.data, None, None, 1, L.RC_VERTS_ON_DEVICE, C.byref(hh)) == 0
        build_ms.append(1e3 * (time.time() - t0))
        build_dev_ms.append(float(lib.rc_last_build_ms(ctx)))
        dd = C.c_int32()
        lib.rc_delete(ctx, hh.value, C.byref(dd))
        tlas.sync()  # frees the deleted BLAS so the next build reuses the pooled blocks (steady-state rebuild cost)
    n_tris = tlas.sizes()["blas_prims"]
    faces = tlas.read_blas_faces(1).astype(np.int64)

    rays, primary_hit_rate = build_rays(lambda r: tlas.trace_closest(r), verts, faces, n, seed=rank)
    d_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1)).to(dev)
    d_hits = torch.empty(n * 32, dtype=torch.uint8, device=dev)
    # ---- result gather for N > 1 --------------------------------------------------------------------------------
    # fused (default): rank 0 owns one world*n*32-byte buffer, exports it over CUDA IPC, and every rank's traversal kernel
    # stores its hit records straight into its slice of that buffer through the NVLink peer mapping — the "gather" is the
    # kernel's own epilogue, there is no separate collective.  nccl (fallback / --gather nccl): dist.gather after the trace.
    gather_mode, gather_buf, hits_ptr, g_base = "none (1 GPU)", None, d_hits.data_ptr(), C.c_void_p()
    if world > 1:
        gather_mode = "nccl"
        if args.gather in ("fused", "peer-copy"):
            try:
                from raycore_b200.sharding import PeerResultBuffer

                peer = PeerResultBuffer(tlas, world * n * 32)
                g_base = peer.base
                gather_mode = args.gather
                if gather_mode == "fused":
                    hits_ptr = peer.ptr(rank * n * 32)
                else:  # double-buffered local hit buffers, pushed to rank 0 by the copy engine while the next step traces
                    d_hits2 = [d_hits, torch.empty_like(d_hits)]
            except Exception as e:  # pragma: no cover
                print(f"[rank {rank}] fused gather unavailable ({e}); falling back to NCCL gather", file=sys.stderr)
        if gather_mode == "nccl" and rank == 0:
            gather_buf = [torch.empty_like(d_hits) for _ in range(world)]
    # all library work on torch's current stream so torch.cuda.Event brackets it
    stream = torch.cuda.Stream(dev)  # a real (non-default) stream: handle 0 would mean "private stream" to rc_set_stream
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0 and lib.rc_set_stream(ctx, C.c_void_p(stream.cuda_stream)) == 0
    flags = L.RC_RAYS_ON_DEVICE | L.RC_HITS_ON_DEVICE | L.RC_NO_SYNC

    step_no = [0]

    def step():
        if gather_mode == "peer-copy":
            b = step_no[0] & 1
            step_no[0] += 1
            assert lib.rc_stream_wait_copy(ctx, b) == 0
            assert lib.rc_trace_closest(ctx, d_rays.data_ptr(), d_hits2[b].data_ptr(), n, flags) == 0, lib.rc_last_error(ctx)
            assert lib.rc_peer_copy_async(ctx, C.c_void_p(peer.ptr(rank * n * 32)), d_hits2[b].data_ptr(), n * 32, b) == 0
            return
        rc_ = lib.rc_trace_closest(ctx, d_rays.data_ptr(), hits_ptr, n, flags)
        assert rc_ == 0, lib.rc_last_error(ctx)
        if gather_mode == "nccl":
            dist.gather(d_hits, gather_buf, dst=0)  # results gathered by NCCL over NVLink

    def barrier():
        if gather_mode == "peer-copy":
            assert lib.rc_wait(ctx) == 0  # drains the copy stream too
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kern_ms = []
    with ClockSampler(local) as clk:
        barrier()
        clk.mark_start()
        ev0.record(stream)
        for _ in range(args.steps):
            step()
        ev1.record(stream)
        barrier()
        clk.mark_end()
    ms_total = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms_total], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    assert lib.rc_wait(ctx) == 0
    if world > 1:
        # check what rank 0 received: per-rank hit counts of the gathered blocks == the counts each rank measures locally
        assert lib.rc_trace_closest(ctx, d_rays.data_ptr(), d_hits.data_ptr(), n, L.RC_RAYS_ON_DEVICE | L.RC_HITS_ON_DEVICE) == 0
        local_hits = int(d_hits.view(torch.int32).view(-1, 8)[:, 0].sum().item())
        counts = [None] * world
        dist.all_gather_object(counts, local_hits)
        if rank == 0:
            if gather_mode in ("fused", "peer-copy"):
                host = np.empty(world * n * 32, np.uint8)
                assert lib.rc_memcpy_d2h(ctx, host.ctypes.data, g_base, host.nbytes) == 0
                got = [int(host.view(np.uint32).reshape(-1, 8)[r * n:(r + 1) * n, 0].sum()) for r in range(world)]
            else:
                got = [int(b.view(torch.int32).view(-1, 8)[:, 0].sum().item()) for b in gather_buf]
            assert got == counts, (got, counts)
    ms_step = ms_total / args.steps
    value = world * n * args.steps / (ms_total * 1e-3) / 1e6

    # per-launch kernel time (CUDA events inside the library, on the launching stream), for the roofline
    for _ in range(3):
        assert lib.rc_trace_closest(ctx, d_rays.data_ptr(), d_hits.data_ptr(), n, L.RC_RAYS_ON_DEVICE | L.RC_HITS_ON_DEVICE) == 0
        kern_ms.append(lib.rc_last_kernel_ms(ctx))
    k_ms = float(np.mean(kern_ms))
    hits_np = d_hits.cpu().numpy().view(L.HIT_DTYPE)
    hit_rate = float(hits_np["hit"].mean())

    # ---- e2e: same call, pinned host buffers, H2D + D2H inside the timed region ------------------------------
    e2e = None
    if not args.no_e2e:
        assert lib.rc_set_stream(ctx, None) == 0
        h_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1)).pin_memory()
        h_hits = torch.empty(n * 32, dtype=torch.uint8).pin_memory()
        for _ in range(2):
            assert lib.rc_trace_closest(ctx, h_rays.data_ptr(), h_hits.data_ptr(), n, 0) == 0
        barrier()
        t0 = time.perf_counter()
        e_steps = max(3, min(args.steps, 10))
        for _ in range(e_steps):
            assert lib.rc_trace_closest(ctx, h_rays.data_ptr(), h_hits.data_ptr(), n, 0) == 0
        barrier()
        e_s = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([e_s], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_s = float(t.item())
        assert h_hits.numpy().view(L.HIT_DTYPE)["hit"].mean() == hits_np["hit"].mean()
        e2e = {"value": world * n * e_steps / e_s / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": n * 32, "d2h_bytes_per_step": n * 32, "steps": e_steps}

    # ---- instrumented pass: per-ray work of the shipped kernel ---------------------------------
    counters = None
    if rank == 0:
        m = min(n, 1 << 20)
        lib.rc_get_counters(ctx, (C.c_uint64 * 6)(), 1)
        assert lib.rc_trace_closest(ctx, d_rays.data_ptr(), d_hits.data_ptr(), m, L.RC_RAYS_ON_DEVICE | L.RC_HITS_ON_DEVICE | L.RC_COUNTERS) == 0
        c = tlas.counters()
        counters = {k: c[k] / m for k in ("nodes", "box_tests", "tri_tests", "inst_entries")} | {"max_stack": c["max_stack"]}

    vf = None if args.no_extras else measure_view_factors(rc, W, L, torch, dev, local, world, rank, dist if world > 1 else None)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    hbm_peak, sm_max, which = peaks()
    rays_per_s = n / (k_ms * 1e-3)
    l2_measured = measure_l2_gbs(torch, dev)
    # roofline (task definition): achieved = ALGORITHMIC bytes per launch / kernel duration, with SURVEY §8d's per-ray figure
    #   bytes_alg = 32 (RTRay) + 32 (RTHitResult) + n_node*64 + n_tri*48 + n_inst*64, counts measured by the instrumented build of
    # the shipped kernel on the first 2^20 rays.  `traffic` is the DRAM traffic ncu measured for the same launch: far BELOW the
    # algorithmic bytes, because node/triangle fetches are served by L1/L2 (the BVH working set is cache resident) — HBM only
    # carries the 64 B/ray streams plus BVH refetches.
    stream_bytes = 64.0
    c_ = counters or {"nodes": 0.0, "tri_tests": 0.0, "inst_entries": 0.0, "box_tests": 0.0}
    bvh_bytes = c_["nodes"] * 64.0 + c_["tri_tests"] * 48.0 + c_["inst_entries"] * 64.0
    bytes_alg = stream_bytes + bvh_bytes
    fp32_peak = 148 * 128 * 2 * sm_max * 1e6 / 1e12
    flops = c_["box_tests"] * 25 + c_["tri_tests"] * 58 + c_["inst_entries"] * 42
    roofline = {
        "bound": "hbm", "achieved": rays_per_s * bytes_alg / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": rays_per_s * bytes_alg / 1e9 / hbm_peak,
        "traffic": ncu_traffic(n), "peak_source": which, "kernel": "k_trace_wide<closest>", "kernel_ms": k_ms, "rays_per_launch": n,
        "bytes_alg_per_ray": bytes_alg, "bytes_alg_per_launch": bytes_alg * n,
        "note": "algorithmic bytes include the BVH node/triangle fetches, which L1/L2 serve (ncu: L2 hit 83 %, DRAM 3-4 % of peak); the kernel is "
                "bound by instruction issue / the ALU pipe with the L1 data pipe close behind (ncu: issue slots 76 %, ALU 72 %, LSU wavefronts 70 %), not by HBM and not by tensor cores — see profiles/README.md",
        "hbm_streams": {"bytes_per_ray": stream_bytes, "achieved_gbs": rays_per_s * stream_bytes / 1e9, "frac": rays_per_s * stream_bytes / 1e9 / hbm_peak},
        "l2": {"bytes_per_ray": bvh_bytes, "achieved_gbs": rays_per_s * bvh_bytes / 1e9, "peak_gbs": 6300 * sm_max * 1e6 / 1e9, "peak_source": "6300 B/clk x sm_max_mhz (B300_MICROARCH LTS cap)",
               "peak_gbs_measured": l2_measured, "peak_measured_source": "L2-resident 24 MiB device copy on this box, read + write bytes"},
        "fp32": {"flop_per_ray": flops, "achieved_tflops": rays_per_s * flops / 1e12, "peak_tflops": fp32_peak},
        "per_ray": counters,
    }

    # ---- the other two numbers of BASELINE.json's metric: BVH build ms (above) and view_factors s (C4, measured before the ranks part)
    extras = None
    if vf is not None:
        extras = dict(vf, blas_build_ms_1M_triangles=min(build_dev_ms))
        if world == 1:  # the instanced scene of BASELINE's target (a reported extra: it must never take the headline line down with it)
            try:
                extras.update(measure_instanced(rc, W, L, torch, dev, local))
            except Exception as e:  # noqa: BLE001
                extras["instanced_error"] = repr(e)

    cpu = None
    if not args.no_cpu_baseline and world == 1:  # the CPU baseline is an N = 1 figure (torchrun also pins OMP_NUM_THREADS=1)
        from oracle import oracle as orc

        t0 = time.time()
        ob = orc.OracleBLAS.from_verts(verts)
        ot = orc.OracleTLAS([ob], orc.make_instances(1, [orc.identity3x4()], [1]))
        cpu_build = time.time() - t0
        sample = rays[:CPU_SAMPLE]
        cores = host_threads()
        ot.closest_hit(sample[: 1 << 16], threads=cores)
        t0 = time.time()
        oh, oc = ot.closest_hit(sample, threads=cores, counters=True)
        dt = time.time() - t0
        # parity on the sample while we are here (ids bit-exact outside the documented classes)
        import parity

        cls = parity.classify(hits_np[: len(sample)], oh, None)
        cpu = {"value": len(sample) / dt / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
               "sample": f"first {len(sample)} rays of the benchmark ray set, 1 pass; oracle/oracle.c (C restatement of the reference BVH2 path, OpenMP over rays; Julia absent)",
               "build_s": cpu_build, "bvh2_per_ray": {k: oc[k] / len(sample) for k in ("nodes", "box_tests", "tri_tests")},
               "parity_on_sample": {k: int(len(v)) for k, v in cls.items()}}

    line = {
        "metric": "closest_hit Mrays/s (incoherent)", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": f"C2: bumpy_sphere({TESS}) {len(verts)} faces -> {n_tris} triangles, 1 instance TLAS; per rank 2^{int(np.log2(n))} rays = diffuse-bounce (hemisphere about the geometric normal, from {PRIMARY_RES}^2 primary hits) interleaved with interior-origin uniform-direction rays",
            "rays_per_rank": n, "hit_rate": hit_rate, "primary_hit_rate": primary_hit_rate, "l2_policy": "inputs larger than L2 (512 MiB rays + 512 MiB hits per step)",
            "gather": {"fused": "fused: each rank's traversal kernel stores its hit records straight into rank 0's buffer through CUDA-IPC peer pointers over NVLink (no separate collective)", "peer-copy": "each rank traces into double-buffered local hit buffers; the copy engine pushes a finished buffer into rank 0's CUDA-IPC-mapped gather buffer over NVLink while the next step traces", "nccl": "dist.gather of hit records to rank 0 (NCCL) after every trace"}.get(gather_mode, gather_mode), "parallelism": f"bvh replicated, rays sharded x{world}",
        },
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": args.steps * 2, "clocks": clk.summary(), "extras": extras,
        "build": {"blas_build_ms_cuda_events": min(build_dev_ms), "blas_build_ms_wall_device_input": min(build_ms), "push_sync_ms_host_input": build_wall_ms, "triangles": n_tris},
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
