#!/usr/bin/env python
"""Per-source-line executed warp-instruction totals from an ncu report captured with --import-source on.

    python scripts/ncu_source_hot.py <report.ncu-rep> [top_n]
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
per_line, per_file = defaultdict(int), defaultdict(int)
samples = defaultdict(int)
cur_file, cur_line, cur_src = None, None, None
total = 0
for row in csv.reader(io.StringIO(txt)):
    if not row:
        continue
    if row[0] == "File Path":
        cur_file = row[1].split("/")[-1]
        continue
    if row[0] in ("Function Name", "Line No"):
        continue
    if row[0] != "":                       # a CUDA source line header (aggregate row)
        cur_line, cur_src = row[0], row[1].strip()
        continue
    try:
        n = int(row[7])
        s = int(row[6])
    except (ValueError, IndexError):
        continue
    per_line[(cur_file, cur_line, cur_src)] += n
    samples[(cur_file, cur_line, cur_src)] += s
    per_file[cur_file] += n
    total += n
print(f"total warp instructions: {total}")
for f, n in sorted(per_file.items(), key=lambda kv: -kv[1]):
    print(f"  {f:24s} {n:12d} {100.0 * n / total:5.1f}%")
print()
for (f, l, s), n in sorted(per_line.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{100.0 * n / total:5.1f}% {n:10d} smp {samples[(f, l, s)]:5d} {f}:{l}  {s[:110]}")
