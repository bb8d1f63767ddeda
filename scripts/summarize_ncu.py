#!/usr/bin/env python
"""Summaries of ncu captures for profiles/ (run here, on the CPU box, over files brought back in gpurun_out/).

  python scripts/summarize_ncu.py launches gpurun_out/launches.csv profiles/rN_launches_summary.csv "note"
  python scripts/summarize_ncu.py raw gpurun_out/prof.ncu-rep profiles/rN_kernel_metrics.csv
"""
import collections
import csv
import re
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size", "lts__t_bytes.sum",
    "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
    "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_not_selected",
    "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_branch_resolving", "smsp__pcsamp_warps_issue_stalled_barrier",
    "smsp__pcsamp_warps_issue_stalled_selected", "smsp__pcsamp_warps_issue_stalled_lg_throttle", "smsp__pcsamp_warps_issue_stalled_no_instructions",
]


def short_name(name):
    s = re.sub(r"\(.*", "", name).replace("void ", "")
    if "cub::" in s:
        s = "cub::" + re.search(r"cub::(?:\w+::)*(\w+)", s).group(1)
    if "at::" in s:
        m = re.search(r"at::(?:native::)?(\w+)", s)
        s = "torch " + (m.group(1) if m else s[:40])
    return s


def launches(src, dst, note=""):
    lines = [l for l in open(src) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    tot = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for r in rows:
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ms = v / 1e6 if unit == "ns" else v / 1e3 if unit == "us" else v if unit == "ms" else v * 1e3
        k = short_name(r["Kernel Name"])
        tot[k][0] += 1
        tot[k][1] += ms
        tot[k][2] = max(tot[k][2], ms)
    total = sum(v[1] for v in tot.values())
    out = [f"# {note}", f"# source: {src}; {len(rows)} launches, {total:.2f} ms of kernel time (cold-cache, serialised: compare SHARES)",
           "kernel,launches,total_ms,share,max_launch_ms"]
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{k},{v[0]},{v[1]:.3f},{v[1] / total:.4f},{v[2]:.3f}")
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out))


def raw(src, dst, note=""):
    text = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(text.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [i for i, h in enumerate(hdr) if h in METRICS or h in ("Kernel Name", "ID")]
    out = [f"# {note}", f"# source: ncu --set full --clock-control none --import-source on ({src})", ",".join(f"{hdr[i]} [{units[i]}]" if units[i] else hdr[i] for i in cols)]
    for r in rows[2:]:
        out.append(",".join('"' + short_name(r[i]) + '"' if hdr[i] == "Kernel Name" else r[i] for i in cols))
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](*sys.argv[2:])
