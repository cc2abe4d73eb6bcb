"""Small run of every env type (CD on, autoreset, wind, noise) + the tiled CD, for compute-sanitizer."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from bluesky_gym_sasha_b200.cd import StateBasedCD
from bluesky_gym_sasha_b200.vector_env import BlueSkyVectorEnv

W = dict(lat=np.array([51.9, 51.9, 52.1, 52.1]), lon=np.array([3.9, 4.1, 3.9, 4.1]),
         vnorth=np.array([[16.0, 12.0, 14.0, 15.0]]), veast=np.array([[3.0, 7.0, 9.0, 4.0]]))
for env_id, kw in [("HorizontalCREnv-v0", dict(n_intruders=20, cd_enabled=True)), ("HorizontalCREnv-v0", dict(cd_enabled=True)),
                   ("HorizontalCREnv-v0", dict(n_intruders=12, cd_enabled=True)),
                   ("SectorCREnv-v0", dict(cd_enabled=True, wind=W, wind_obs=True)), ("MergeEnv-v0", dict(cd_enabled=True)),
                   ("VerticalCREnv-v0", dict(cd_enabled=True, obs_noise=0.1)), ("DescentEnv-v0", dict(wind=W)),
                   ("PlanWaypointEnv-v0", {}), ("StaticObstacleEnv-v0", {})]:
    v = BlueSkyVectorEnv(env_id, 40, seed=1, autoreset_mode="same_step", max_episode_steps=6, **kw)
    v.reset()
    rng = np.random.default_rng(0)
    for _ in range(14):
        v.step(rng.uniform(-1, 1, (40, v.layout.act_dim)).astype(np.float32))
    v.close()
    print("ok", env_id, kw.keys())
rng = np.random.default_rng(0)
n = 3000
out = StateBasedCD(device=0).detect(52 + 3 * rng.random(n), 4 + 3 * rng.random(n), rng.uniform(0, 360, n), rng.uniform(150, 250, n),
                                    rng.uniform(3000, 12000, n), np.zeros(n))
print("cd ok", len(out["confpairs"]))
torch.cuda.synchronize()
