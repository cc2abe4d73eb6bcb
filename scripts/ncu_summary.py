#!/usr/bin/env python
"""Summarise ncu output into profiles/ (tracked).

    python scripts/ncu_summary.py launches <launches.csv> <out.md> [title]      # per-kernel totals + shares
    python scripts/ncu_summary.py full <report.ncu-rep> <out.md> [kernel-substr] # --set full digest of one kernel
                                                                                 # also updates profiles/ncu_traffic.json

`ncu` (the CLI that reads .ncu-rep files) must be on PATH; no GPU is needed to read a report.
"""
import csv
import io
import json
import os
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FULL_KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_xu.sum",
    "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_lsu.sum",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
]


def launches(path, out, title):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
    h = rows[0]
    ik, im, iv = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
    iu = h.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1e-3)
        a = agg.setdefault(r[ik], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# {title}\n\nper-launch times under ncu are cold-cache and serialised: compare shares, not absolutes.\n"
                f"Raw list: `{os.path.basename(path)}`\n\n| launches | total us | share | kernel |\n|---:|---:|---:|---|\n")
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {n} | {us:.1f} | {100 * us / tot:.1f}% | `{k[:110]}` |\n")


def _to_float(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return None


def full(rep, out, substr):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    h, units = rows[0], rows[1]
    ik = h.index("Kernel Name")
    traffic_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    with open(out, "w") as f:
        for r in rows[2:]:
            if substr and substr not in r[ik]:
                continue
            f.write(f"## `{r[ik]}` -- `ncu --set full --clock-control none` ({os.path.basename(rep)})\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in FULL_KEYS:
                if k in h:
                    f.write(f"| {k} | {r[h.index(k)]} | {units[h.index(k)]} |\n")
            f.write("\n")
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            def bytes_of(k):
                i = h.index(k)
                return (_to_float(r[i]) or 0.0) * scale.get(units[i], 1.0)
            name = r[ik].split("(")[0].replace("void ", "").strip()
            traffic[name] = {"dram_bytes_read": bytes_of("dram__bytes_read.sum"),
                             "dram_bytes_write": bytes_of("dram__bytes_write.sum"),
                             "grid": r[h.index("launch__grid_size")], "block": r[h.index("launch__block_size")],
                             "duration": r[h.index("gpu__time_duration.sum")] + " " + units[h.index("gpu__time_duration.sum")],
                             "source": os.path.basename(rep)}
    json.dump(traffic, open(traffic_path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "ncu launch list")
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
