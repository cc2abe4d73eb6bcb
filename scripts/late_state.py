"""Steps the C2 workload N times (no build step: A/B libraries via BSG_B200_LIB); for ncu captures of a late-episode launch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bluesky_gym_sasha_b200.vector_env import BlueSkyVectorEnv

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
v = BlueSkyVectorEnv("HorizontalCREnv-v0", 4096, seed=0, cd_enabled=True, n_intruders=20, autoreset_mode="same_step")
v.reset_torch()
a = torch.rand((n, 4096, 1), device="cuda") * 2 - 1
for i in range(n):
    v.step_torch(a[i])
torch.cuda.synchronize()
print("done", os.environ.get("BSG_B200_LIB"))
