"""Device time per step of C2 in the steady state of a policy that ends episodes (near-straight flight reaches the
waypoint after 22-33 steps, so resets are spread over the steps) vs the synchronized 300-step episodes of bench.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bluesky_gym_sasha_b200.vector_env import BlueSkyVectorEnv

E = 4096
for scale, label in ((1.0, "random actions (episodes run to the 300-step cap)"), (0.02, "near-straight flight (episodes end after 22-33 steps)")):
    v = BlueSkyVectorEnv("HorizontalCREnv-v0", E, seed=0, cd_enabled=True, n_intruders=20, autoreset_mode="same_step")
    v.reset_torch()
    n = 500
    a = (torch.rand((n, E, 1), device="cuda") * 2 - 1) * scale
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
    done = 0
    for i in range(250):
        _, _, te, tr = v.step_torch(a[i])
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(250)]
    for i in range(250):
        flush.fill_(float(i))
        ev[i][0].record()
        _, _, te, tr = v.step_torch(a[250 + i])
        ev[i][1].record()
        done += int((te | tr).sum())
    torch.cuda.synchronize()
    t = sorted(x.elapsed_time(y) for x, y in ev)
    print(f"{label}: median {t[125] * 1e3:.1f} us, mean {sum(t) / 250 * 1e3:.1f} us, p95 {t[237] * 1e3:.1f} us, "
          f"{done / 250:.1f} envs finish per step")
    v.close()
