"""env-steps/s against the number of envs on one GPU (sweep 2^12 .. 2^20), back-to-back launches (L2-warm), after
150 pre-roll steps (stationary regime).  Output kept in profiles/r2t_sweep_envs.txt."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bluesky_gym_sasha_b200.vector_env import BlueSkyVectorEnv
def run(env_id, E, K=30, **kw):
    v = BlueSkyVectorEnv(env_id, E, seed=0, autoreset_mode="same_step", **kw)
    v.reset_torch()
    a = torch.rand((K+5, E, v.layout.act_dim), device="cuda")*2-1
    for i in range(150): v.step_torch(a[i % (K + 5)])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K): v.step_torch(a[5+i])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/K
    print(f"{env_id:20s} E={E:8d} {kw}  {ms*1e3:9.1f} us/step  {E/ms*1e3:.3e} env-steps/s", flush=True)
    v.close()
for E in (4096, 8192, 16384, 32768, 65536, 131072, 262144, 524288, 1048576):
    run("HorizontalCREnv-v0", E, n_intruders=20, cd_enabled=True)
for E in (4096, 65536, 1048576):
    run("HorizontalCREnv-v0", E, n_intruders=5, cd_enabled=False)
    run("HorizontalCREnv-v0", E, n_intruders=5, cd_enabled=True)
for E in (8192, 65536):
    run("SectorCREnv-v0", E, cd_enabled=False)
    run("SectorCREnv-v0", E, cd_enabled=True)
    run("MergeEnv-v0", E, cd_enabled=False)
    run("MergeEnv-v0", E, cd_enabled=True)
for E in (65536, 1048576):
    run("DescentEnv-v0", E)
