"""How much do MergeEnv's FMS-guided intruders change their ground-speed vector per substep / per env step?  (sizing of
kCdVelTol in env_kernels.cuh: the kept candidate list of the in-group CD survives while |du| + |dv| stays below it.)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bluesky_gym_sasha_b200.vector_env import BlueSkyVectorEnv

E = 1024
v = BlueSkyVectorEnv("MergeEnv-v0", E, seed=0, cd_enabled=True, autoreset_mode="same_step")
v.reset_torch()
g = torch.Generator(device="cuda").manual_seed(1)


def uv():
    k = v.t["kin"]
    h = torch.deg2rad(k[:, 1:20, 2].double())
    t = k[:, 1:20, 1].double()
    return t * torch.sin(h), t * torch.cos(h)


for step in range(30):
    a = torch.rand((E, 2), device="cuda", generator=g) * 2 - 1
    if step in (0, 1, 2, 5, 10, 20, 29):
        sd = v.state_dict()
        u0, v0 = uv()
        out = []
        for n in range(1, 11):
            v.traf_update(1)
            u1, v1 = uv()
            d = ((u1 - u0).abs() + (v1 - v0).abs())
            out.append("%.3f/%.3f" % (d.median().item(), d.quantile(0.99).item()))
        print("step", step, "cumulative |du|+|dv| after n substeps (median/p99):", " ".join(out))
        v.load_state_dict(sd)
    v.step_torch(a)
