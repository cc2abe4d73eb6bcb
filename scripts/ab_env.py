import sys, os, subprocess
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, os, torch
sys.path.insert(0, %r)
from bluesky_gym_sasha_b200.vector_env import BlueSkyVectorEnv
E=4096; K=200
v = BlueSkyVectorEnv("HorizontalCREnv-v0", E, seed=0, autoreset_mode="same_step", n_intruders=20, cd_enabled=True)
v.reset_torch()
a = torch.rand((K+20, E, 1), device="cuda")*2-1
for i in range(20): v.step_torch(a[i])
torch.cuda.synchronize()
best=1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K): v.step_torch(a[20+i])
    e1.record(); torch.cuda.synchronize()
    best=min(best, e0.elapsed_time(e1)/K)
print(os.environ.get("BSG_B200_LIB"), "%%.2f us/step" %% (best*1e3))
''' % root
for lib in sys.argv[1:]:
    env = dict(os.environ, BSG_B200_LIB=os.path.join(root, lib))
    subprocess.run([sys.executable, "-c", code], env=env)
