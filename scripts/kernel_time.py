"""Device time of one batched env step for a few configurations (CUDA events, L2 flushed between steps)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bluesky_gym_sasha_b200.vector_env import BlueSkyVectorEnv


def run(env_id, E, cd, steps=100, **kw):
    v = BlueSkyVectorEnv(env_id, E, seed=0, cd_enabled=cd, autoreset_mode="same_step", **kw)
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
    ms, mean = ts[steps // 2], sum(ts) / steps
    print(f"{env_id:22s} E={E:6d} cd={int(cd)} {kw}: median {ms * 1e3:8.1f} mean {mean * 1e3:8.1f} us/step  {E / mean * 1e3:.3e} env-steps/s")
    v.close()


if __name__ == "__main__":
    run("HorizontalCREnv-v0", 4096, True, n_intruders=20)
    run("HorizontalCREnv-v0", 4096, False, n_intruders=20)
    run("HorizontalCREnv-v0", 4096, False)
    run("HorizontalCREnv-v0", 65536, True, n_intruders=20)
    run("SectorCREnv-v0", 8192, True)
    run("SectorCREnv-v0", 8192, False)
    run("MergeEnv-v0", 4096, True)
    run("MergeEnv-v0", 4096, False)
    run("MergeEnv-v0", 65536, False)
    run("DescentEnv-v0", 65536, False)
    run("VerticalCREnv-v0", 16384, True)
    run("PlanWaypointEnv-v0", 65536, False)
    run("StaticObstacleEnv-v0", 16384, False)
