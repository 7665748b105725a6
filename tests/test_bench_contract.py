"""The committed bench line (profiles/r1_bench_*.json, written by bench.py on a B200) carries every key of the driver contract and
the roofline / cpu_baseline / e2e objects; bench.py's argument surface is the contracted one."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _latest():
    """the newest committed line: round 2 (C3 headline, strong scaling) when present, else round 1 (C2, weak scaling)"""
    for rnd in ("r2", "r1"):
        files = sorted(glob.glob(os.path.join(ROOT, "profiles", f"{rnd}_bench_v*.json")), key=lambda f: int(f.split("_v")[-1].split(".")[0]))
        if files:
            return json.loads(open(files[-1]).read().strip().splitlines()[-1]), rnd
    raise AssertionError("no committed bench line under profiles/")


def test_bench_line_has_contract_keys():
    d, rnd = _latest()
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
              "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["unit"] == "Mrays/s" and d["higher_is_better"] is True and d["scaling"] == ("strong" if rnd == "r2" else "weak") and d["vs_baseline"] is None
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    assert d["warmup"] >= 3 and d["gpu_launches"] >= d["steps"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] in ("hbm", "tensor", "issue") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    if rnd == "r2":
        assert r["kernel_ms"] <= d["ms_per_step"] * 1.02, "the per-launch kernel time cannot exceed the step it is part of"
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and "sample" in c
    assert c["parity_on_sample"]["bad"] == 0
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    k = d["clocks"]
    assert k["samples_in_timed_region"] > 0 and not set(k["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    rays_per_step = d["config"]["total_rays"] if rnd == "r2" else d["config"]["rays_per_rank"] * d["n_gpus"]
    assert abs(d["value"] - rays_per_step / (d["ms_per_step"] * 1e-3) / 1e6) < 1e-3 * d["value"]


def test_both_arms_describe_the_same_workload():
    """bench.py's two arms print the same `config` object (the driver compares them) and generate the same ray bytes for the same indices."""
    sys.path.insert(0, ROOT)
    import numpy as np

    import bench
    from raycore_b200 import workloads as W

    assert bench.bench_config(4, 100_000_000) == bench.bench_config(4, 100_000_000)
    a = np.empty(70_000, W.RAY_DTYPE)
    bench.gen_box_rays(a, 1_000_000, 3)
    b = W.box_rays(70_000, bench.RAY_SEED, half=bench.RAY_HALF, first_index=1_000_000)
    assert a.tobytes() == b.tobytes()
    # a rank's slice of the 100 M-ray set is the same bytes as that range of the full set (strong-scaling shards)
    c = np.empty(1000, W.RAY_DTYPE)
    bench.gen_box_rays(c, 1_000_500, 2)
    assert c.tobytes() == a[500:1500].tobytes()


def test_bench_cli_surface():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True, timeout=120).stdout
    for flag in ("--gpus", "--steps", "--warmup", "--impl"):
        assert flag in out
