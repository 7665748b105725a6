"""The committed bench line (profiles/r1_bench_*.json, written by bench.py on a B200) carries every key of the driver contract and
the roofline / cpu_baseline / e2e objects; bench.py's argument surface is the contracted one."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _latest():
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r1_bench_v*.json")), key=lambda f: int(f.split("_v")[-1].split(".")[0]))
    return json.loads(open(files[-1]).read().strip().splitlines()[-1])


def test_bench_line_has_contract_keys():
    d = _latest()
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
              "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["unit"] == "Mrays/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    assert d["warmup"] >= 3 and d["gpu_launches"] >= d["steps"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and "sample" in c
    assert c["parity_on_sample"]["bad"] == 0
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    k = d["clocks"]
    assert k["samples_in_timed_region"] > 0 and not set(k["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert abs(d["value"] - d["config"]["rays_per_rank"] * d["n_gpus"] / (d["ms_per_step"] * 1e-3) / 1e6) < 1e-3 * d["value"]


def test_bench_cli_surface():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True, timeout=120).stdout
    for flag in ("--gpus", "--steps", "--warmup", "--impl"):
        assert flag in out
