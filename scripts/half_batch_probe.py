import sys, os, torch, numpy as np
sys.path.insert(0, os.getcwd())
from bluesky_gym_sasha_b200.vector_env import BlueSkyVectorEnv
for E in (1024, 2048, 4096):
    v = BlueSkyVectorEnv("HorizontalCREnv-v0", E, seed=0, cd_enabled=True, n_intruders=20, autoreset_mode="same_step")
    v.reset_torch()
    a = torch.rand((E, 1), device="cuda") * 2 - 1
    for _ in range(100): v.step_torch(a)
    torch.cuda.synchronize()
    ts = []
    for i in range(200):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(100000)
        e0.record(); v.step_torch(a); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print(E, "L2-warm single launch: median %.1f us" % np.median(ts))
    v.close()
