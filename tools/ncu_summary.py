"""Summarise ncu output for profiles/ (run in the build container; ncu reads .ncu-rep files without a GPU).

    python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/rXX_launches_summary.csv "<command line>"
    python tools/ncu_summary.py full gpurun_out/prof.ncu-rep profiles/rXX_ncu_full.txt "<command line>" [profiles/rXX_traffic.json]
"""
import csv
import json
import subprocess
import sys
from collections import OrderedDict

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct",
]


def short(name):
    name = name.replace("void ", "")
    return name.split("(")[0]


def launches(src, dst, cmd):
    rows = [r for r in csv.reader(l for l in open(src) if not l.startswith("==")) if r]
    hdr = rows[0]
    ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    ui = hdr.index("Metric Unit")
    tot = OrderedDict()
    n = 0
    for r in rows[1:]:
        if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        us = {"ns": v / 1e3, "us": v, "usecond": v, "nsecond": v / 1e3, "ms": v * 1e3, "msecond": v * 1e3}.get(r[ui], v / 1e3)
        k = short(r[ki])
        a = tot.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
        n += 1
    total = sum(a[1] for a in tot.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list ({cmd})\n# ncu --metrics gpu__time_duration.sum --clock-control none ; cold-cache, serialised: compare SHARES\n")
        f.write(f"# launches captured: {n}; total {total / 1e3:.3f} ms\nkernel,launches,total_us,share\n")
        for k, a in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k},{a[0]},{a[1]:.1f},{a[1] / total:.3f}\n")


def sequence(src, dst, cmd, last):
    """the last `last` launches in order (one proof): kernel, grid, duration"""
    rows = [r for r in csv.reader(l for l in open(src) if not l.startswith("==")) if r]
    hdr = rows[0]
    ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    gi = hdr.index("Grid Size") if "Grid Size" in hdr else None
    seq = []
    for r in rows[1:]:
        if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        us = {"ns": v / 1e3, "us": v, "usecond": v, "nsecond": v / 1e3, "ms": v * 1e3, "msecond": v * 1e3}.get(r[ui], v / 1e3)
        seq.append((short(r[ki]), r[gi] if gi is not None else "", us))
    seq = seq[-int(last):]
    with open(dst, "w") as f:
        f.write(f"# the last {last} launches of: {cmd}\n# (ncu --metrics gpu__time_duration.sum --clock-control none: cold-cache, serialised)\n# total {sum(x[2] for x in seq) / 1e3:.3f} ms\nindex,kernel,grid,us\n")
        for i, (k, g, us) in enumerate(seq):
            f.write(f"{i},{k},{g.replace(', ', 'x')},{us:.1f}\n")


def full(src, dst, cmd, traffic_json=None):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    traffic = {}
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on  ({cmd})\n# numbers are per launch\n")
        for r in rows[2:]:
            f.write(f"\n== {short(r[ki])}  grid={r[hdr.index('launch__grid_size')]}\n")
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    f.write(f"   {m:<95} {r[i]} {units[i]}\n")
            try:
                def mb(m):
                    i = hdr.index(m)
                    v = float(r[i].replace(",", ""))
                    return v * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}.get(units[i], 1)
                traffic.setdefault(short(r[ki]), []).append({"grid": int(r[hdr.index('launch__grid_size')].replace(",", "")), "dram_bytes": mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum"),
                                                             "duration_ms": float(r[hdr.index("gpu__time_duration.sum")].replace(",", "")) * {"ms": 1, "us": 1e-3, "ns": 1e-6, "msecond": 1, "usecond": 1e-3, "nsecond": 1e-6}.get(units[hdr.index("gpu__time_duration.sum")], 1)})
            except Exception:
                pass
    if traffic_json:
        json.dump({"command": cmd, "per_launch": traffic}, open(traffic_json, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(*sys.argv[2:5])
    elif sys.argv[1] == "sequence":
        sequence(*sys.argv[2:6])
    else:
        full(*sys.argv[2:6])
