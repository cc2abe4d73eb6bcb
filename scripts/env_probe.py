"""Steps one env type a few times (for ncu captures of its kernel).  usage: env_probe.py ENV_ID E [cd] [steps] [n_intruders]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bluesky_gym_sasha_b200.vector_env import BlueSkyVectorEnv

env_id, E = sys.argv[1], int(sys.argv[2])
cd = len(sys.argv) > 3 and sys.argv[3] == "1"
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 12
kw = dict(n_intruders=int(sys.argv[5])) if len(sys.argv) > 5 else {}
v = BlueSkyVectorEnv(env_id, E, seed=0, cd_enabled=cd, autoreset_mode="same_step", **kw)
v.reset_torch()
a = torch.rand((steps, E, v.layout.act_dim), device="cuda") * 2 - 1
for i in range(steps):
    v.step_torch(a[i])
torch.cuda.synchronize()
print("done")
