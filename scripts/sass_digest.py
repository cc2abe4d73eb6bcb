#!/usr/bin/env python
"""SASS digest of libbsg_b200.so for profiles/: per kernel the instruction count, registers, and the counts of the
mnemonics that show what the code is built from (packed f32x2 arithmetic, TMA bulk copies, mbarrier, MUFU, FP64, REDUX,
shared-memory atomics), plus the hot loop of the all-pairs CD kernel verbatim.

    python scripts/sass_digest.py [lib.so] > profiles/<round>_sass_digest.md      (needs cuobjdump; no GPU)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "bluesky_gym_sasha_b200", "libbsg_b200.so")
WATCH = ["FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL", "FADD", "MUFU", "UBLKCP", "SYNCS", "DFMA", "DMUL", "DADD", "REDUX", "ATOMS",
         "LDS", "STS", "LDG", "STG", "SHFL", "VOTE", "BAR", "S2R", "LDL", "STL", "HMMA", "UTCHMMA", "UTMALDG"]

sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
regs = {}
cur = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        cur = m.group(1)
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", line)
    if m and cur:
        regs[cur] = (int(m.group(1)), int(m.group(3)), int(m.group(2)))
demangle = lambda s: subprocess.run(["cu++filt", s], capture_output=True, text=True).stdout.strip() or s

kernels = collections.OrderedDict()
name = None
arch = set()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = m.group(1)
        kernels[name] = []
        continue
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        arch.add(m.group(1))
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
    if m and name:
        kernels[name].append(m.group(1))

print(f"# SASS digest of `{os.path.relpath(lib, ROOT)}` (cuobjdump -sass; architectures in the fat binary: {', '.join(sorted(arch))})\n")
print("No tensor-core (HMMA / UTCHMMA) or TMEM instruction anywhere: the path is FP32 pair arithmetic, not a contraction "
      "(north star: \"no tensor cores\").  What is Blackwell-specific: packed f32x2 arithmetic (FFMA2 / FMUL2 / FADD2, two "
      "flops per lane per issue slot) in both hot kernels, TMA 1-D bulk copies (UBLKCP) completing on mbarriers (SYNCS) "
      "in the CD kernel, REDUX warp reductions.\n")
print("| kernel | SASS instr | regs | smem B | stack B | " + " | ".join(WATCH) + " |")
print("|---|---:|---:|---:|---:|" + "---:|" * len(WATCH))
for k, ins in kernels.items():
    cnt = collections.Counter()
    for i in ins:
        op = i.split()[0] if not i.startswith("@") else i.split()[1]
        base = op.split(".")[0]
        cnt[base] += 1
    r = regs.get(k, ("", "", ""))
    print(f"| `{demangle(k)[:90]}` | {len(ins)} | {r[0]} | {r[1]} | {r[2]} | " + " | ".join(str(cnt.get(w, 0)) for w in WATCH) + " |")

# the hot loop of the brute-force CD kernel: the longest run of instructions dominated by f32x2 arithmetic
for k, ins in kernels.items():
    if "cd_tiled_kernel" in k and demangle(k).startswith("void bsg::cd_tiled_kernel<false, false, false>") or "cd_tiled_kernelILb0ELb0ELb0" in k:
        idx = [n for n, i in enumerate(ins) if re.search(r"\bF(FMA|MUL|ADD)2\b", i)]
        if idx:
            lo, hi = idx[0], idx[-1]
            # narrow to the densest window of 120 instructions
            best, bl = 0, lo
            for s in range(lo, max(lo + 1, hi - 120)):
                c = sum(1 for n in idx if s <= n < s + 120)
                if c > best:
                    best, bl = c, s
            print(f"\n## Hot loop excerpt, `{demangle(k)}` (120 instructions, {best} of them packed f32x2)\n\n```")
            print("\n".join(ins[bl:bl + 120]))
            print("```")
        break
