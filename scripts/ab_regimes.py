"""A/B of libraries (BSG_B200_LIB) on C2 in two regimes: early episode (traffic converging: every step is one of the
first 12 of an episode) and late episode (steps 60..260 of 300-step episodes)."""
import os
import subprocess
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, os, torch
sys.path.insert(0, %r)
from bluesky_gym_sasha_b200.vector_env import BlueSkyVectorEnv
E = 4096
flush = torch.empty(64*1024*1024, dtype=torch.float32, device="cuda")
def timed(v, acts):
    ev=[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(len(acts))]
    for i, a in enumerate(acts):
        flush.fill_(float(i)); ev[i][0].record(); v.step_torch(a); ev[i][1].record()
    torch.cuda.synchronize()
    t = sorted(x.elapsed_time(y) for x, y in ev)
    return t[len(t)//2]*1e3
early = BlueSkyVectorEnv("HorizontalCREnv-v0", E, seed=0, autoreset_mode="same_step", n_intruders=20, cd_enabled=True, max_episode_steps=12)
early.reset_torch()
a = torch.rand((240, E, 1), device="cuda")*2-1
for i in range(24): early.step_torch(a[i])
te = timed(early, a[24:])
late = BlueSkyVectorEnv("HorizontalCREnv-v0", E, seed=0, autoreset_mode="same_step", n_intruders=20, cd_enabled=True)
late.reset_torch()
for i in range(60): late.step_torch(a[i])
tl = timed(late, a[20:220])
print(os.path.basename(os.environ.get("BSG_B200_LIB", "default")), "early %%.2f us   late %%.2f us" %% (te, tl))
''' % root
for lib in sys.argv[1:]:
    subprocess.run([sys.executable, "-c", code], env=dict(os.environ, BSG_B200_LIB=os.path.join(root, lib)))
