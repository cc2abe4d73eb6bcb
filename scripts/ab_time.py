"""Device time of the C2 step (and a few other configurations) under several library builds, interleaved so that box-to-box
and minute-to-minute drift cancels:  python scripts/ab_time.py lib_a.so lib_b.so ...   (each build runs in its own process)"""
import os
import subprocess
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, torch
sys.path.insert(0, %r)
from bluesky_gym_sasha_b200.vector_env import BlueSkyVectorEnv
def run(env_id, E, steps=200, **kw):
    v = BlueSkyVectorEnv(env_id, E, seed=0, cd_enabled=True, autoreset_mode="same_step", **kw)
    v.reset_torch()
    a = torch.rand((steps + 10, E, v.layout.act_dim), device="cuda") * 2 - 1
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
    for i in range(10):
        v.step_torch(a[i])
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for i in range(steps):
        flush.fill_(float(i))
        ev[i][0].record()
        v.step_torch(a[10 + i])
        ev[i][1].record()
    torch.cuda.synchronize()
    ts = sorted(x.elapsed_time(y) for x, y in ev)
    v.close()
    return sum(ts) / steps * 1e3
print("%%.2f %%.2f %%.2f" %% (run("HorizontalCREnv-v0", 4096, n_intruders=20), run("SectorCREnv-v0", 8192), run("MergeEnv-v0", 4096)))
''' % root
libs = sys.argv[1:]
res = {l: [] for l in libs}
for rep in range(3):
    for l in libs:
        p = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, BSG_B200_LIB=os.path.join(root, l)), capture_output=True, text=True)
        if p.returncode:
            print(l, "FAILED", p.stderr[-1500:])
            sys.exit(1)
        res[l].append([float(x) for x in p.stdout.strip().splitlines()[-1].split()])
print("mean us/step over 200 steps (L2 flushed), 3 interleaved repetitions:  HorizontalCR-20 x4096 | SectorCR x8192 | MergeEnv x4096")
for l in libs:
    print(f"{l:44s}", "  ".join("/".join(f"{r[k]:.1f}" for r in res[l]) for k in range(3)))
