#!/usr/bin/env python
"""Summarise an .ncu-rep (one line per captured launch + stall/pipe breakdown).  usage: ncu_summary.py file.ncu-rep [out.md]"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy%"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads/warp-inst"),
    ("smsp__inst_executed.sum", "warp-insts"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue-active%"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lsu-wavefronts%"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit%"),
    ("lts__t_sector_hit_rate.pct", "L2 hit%"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("dram__bytes_read.sum", "dram rd"),
    ("dram__bytes_write.sum", "dram wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 tput%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 tput%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM tput%"),
]


def num_time(v, u):
    return float(v.replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u.lower().replace("second", "s").replace("usecond", "us"), 1.0)


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    out = []
    for d in data:
        out.append(f"## {d[col['Kernel Name']][:110]}")
        for k, label in KEYS:
            if k in col:
                out.append(f"- {label}: {d[col[k]]} {units[col[k]]}  (`{k}`)")
        st = []
        for h, i in col.items():
            if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
                try:
                    st.append((float(d[i]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                except ValueError:
                    pass
        tot = sum(v for v, _ in st) or 1
        out.append("- warp-state samples: " + ", ".join(f"{n} {100 * v / tot:.1f}%" for v, n in sorted(st, reverse=True)[:9]))
    if len(sys.argv) > 3:  # ncu_summary.py rep out.md traffic.json rays_per_launch
        d = data[-1]
        def num(k):
            v, u = float(d[col[k]].replace(",", "")), units[col[k]].lower()
            return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
        import json
        json.dump({"kernel": d[col["Kernel Name"]][:80], "dram_bytes": num("dram__bytes_read.sum") + num("dram__bytes_write.sum"),
                   "dram_read_bytes": num("dram__bytes_read.sum"), "dram_write_bytes": num("dram__bytes_write.sum"),
                   "rays_per_launch": int(sys.argv[4]), "source": rep,
                   "issue_active_pct": float(d[col["smsp__issue_active.avg.pct_of_peak_sustained_active"]].replace(",", "")),
                   "alu_pct": float(d[col["sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"]].replace(",", "")),
                   "lanes_per_inst": float(d[col["smsp__thread_inst_executed_per_inst_executed.ratio"]].replace(",", "")),
                   "kernel_ms_under_ncu": num_time(d[col["gpu__time_duration.sum"]], units[col["gpu__time_duration.sum"]])}, open(sys.argv[3], "w"))
    text = "\n".join(out)
    print(text)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text + "\n")


if __name__ == "__main__":
    main()
