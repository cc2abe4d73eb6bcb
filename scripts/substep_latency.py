"""Device time of bsg_traf_update(n) against n (slope = one substep, intercept = launch + load / store) for a nearly empty
GPU and for the C2 batch: tells whether the env kernel is bound by the serial latency of a warp's substep chain."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bluesky_gym_sasha_b200.vector_env import BlueSkyVectorEnv


def run(env_id, E, cd, **kw):
    v = BlueSkyVectorEnv(env_id, E, seed=0, cd_enabled=cd, autoreset_mode="same_step", **kw)
    v.reset_torch()
    a = torch.rand((40, E, v.layout.act_dim), device="cuda") * 2 - 1
    for i in range(40):
        v.step_torch(a[i])
    sd = v.state_dict()
    out = []
    for n in (1, 2, 5, 10, 20, 40):
        ts = []
        for rep in range(30):
            v.load_state_dict(sd)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            v.traf_update(n)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        out.append((n, sorted(ts)[len(ts) // 2]))
    slope = (out[-1][1] - out[3][1]) / (out[-1][0] - out[3][0])
    print(f"{env_id} E={E} cd={int(cd)} {kw}: " + "  ".join(f"n={n}: {t:.1f}us" for n, t in out) + f"   slope {slope:.2f} us/substep")
    v.close()


if __name__ == "__main__":
    run("HorizontalCREnv-v0", 128, False)
    run("HorizontalCREnv-v0", 4096, False)
    run("HorizontalCREnv-v0", 128, True, n_intruders=20)
    run("HorizontalCREnv-v0", 4096, False, n_intruders=20)
    run("HorizontalCREnv-v0", 4096, True, n_intruders=20)
    run("HorizontalCREnv-v0", 16384, True, n_intruders=20)
    run("DescentEnv-v0", 4096, False)
