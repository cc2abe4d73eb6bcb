"""A/B of libraries (BSG_B200_LIB) on the C2 workload with the L2 flushed between steps (median device time)."""
import os
import subprocess
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, os, torch
sys.path.insert(0, %r)
from bluesky_gym_sasha_b200.vector_env import BlueSkyVectorEnv
E=4096; K=300
v = BlueSkyVectorEnv("HorizontalCREnv-v0", E, seed=0, autoreset_mode="same_step", n_intruders=20, cd_enabled=True)
v.reset_torch()
a = torch.rand((K+20, E, 1), device="cuda")*2-1
flush = torch.empty(64*1024*1024, dtype=torch.float32, device="cuda")
for i in range(20): v.step_torch(a[i])
ev=[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
for i in range(K):
    flush.fill_(float(i)); ev[i][0].record(); v.step_torch(a[20+i]); ev[i][1].record()
torch.cuda.synchronize()
t=sorted(x.elapsed_time(y) for x,y in ev)
print(os.environ.get("BSG_B200_LIB"), "median %%.2f us  mean %%.2f us" %% (t[K//2]*1e3, sum(t)/K*1e3))
''' % root
for lib in sys.argv[1:]:
    subprocess.run([sys.executable, "-c", code], env=dict(os.environ, BSG_B200_LIB=os.path.join(root, lib)))
