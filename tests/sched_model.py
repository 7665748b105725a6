#!/usr/bin/env python
"""(test infrastructure: uses tests/hostsim) Scheduler statistics of the shipped traversal kernel from the CPU warp simulator: warp
iterations and active lanes per step kind (N node, T triangle, X instance entry, F retire + refill) on the instanced scene C3, and a
cost model  sum_k iterations_k * (body_k + overhead)  in warp instructions (body lengths from the SASS: N 180, T 110, X 60, F 150,
scheduler overhead 20).  The model tracks the B200 within 6 % over the policies measured there (profiles/README.md), so a policy
can be screened here before it costs GPU time:  RC_HOSTSIM_LIB=<experiment build of tests/hostsim> python tests/sched_model.py
usage: python tests/sched_model.py [n_instances=10000] [n_rays=40000]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402

from raycore_b200 import workloads as W  # noqa: E402
import engines  # noqa: E402

BODY = {"N": 180, "T": 110, "X": 60, "F": 150}
OVERHEAD = 20


def main():
    n_inst = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    n_rays = int(sys.argv[2]) if len(sys.argv) > 2 else 40000
    e = engines.HostsimEngine([(W.bumpy_sphere(72), None, W.random_trs(n_inst, 2026, extent=40.0), None)])
    rays = W.box_rays(n_rays, 7, half=44.0)
    _, c = e.scene.trace_warpsim(rays, n_warps=8, counters=True)
    it, ln = c["step_iterations"], c["step_lanes"]
    out = {
        "lib": os.environ.get("RC_HOSTSIM_LIB", "tests/hostsim/libhostsim.so"),
        "per_ray": {k: c[k] / n_rays for k in ("nodes", "tri_tests", "inst_entries")},
        "iterations_per_32_rays": {k: 32.0 * it[k] / n_rays for k in "NTXF"},
        "lanes_per_iteration": {k: ln[k] / max(1, it[k]) for k in "NTXF"},
        "idle_lanes_per_node_step": {k: v / max(1, it["N"]) for k, v in c["idle_in_node_steps"].items()},
        "model_warp_instructions_per_ray": sum(it[k] * (BODY[k] + OVERHEAD) for k in "NTXF") / n_rays,
    }
    print(json.dumps(out))


if __name__ == "__main__":
    main()
