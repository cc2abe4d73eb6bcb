"""Drop-in usage: the reference's own loop (main.py:19,34 / scripts/multi_processing_example.py) on the accelerated envs.

    python examples/random_rollout.py [ENV_ID] [NUM_ENVS]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # run from a checkout
import bluesky_gym                                   # the alias package: same import as with the reference
from bluesky_gym_sasha_b200 import BlueSkyVectorEnv

env_id = sys.argv[1] if len(sys.argv) > 1 else "HorizontalCREnv-v0"
num_envs = int(sys.argv[2]) if len(sys.argv) > 2 else 4096

# 1. single env through gym.make, exactly like the reference ----------------------------------------------------------
bluesky_gym.register_envs()
env = bluesky_gym.make(env_id, render_mode=None)
obs, info = env.reset(seed=0)
done, ret = False, 0.0
while not done:
    obs, reward, terminated, truncated, info = env.step(env.action_space.sample())
    ret += reward
    done = terminated or truncated
print(f"{env_id}: one episode, return {ret:.2f}, info {info}")
env.close()

# 2. the batched simulator: thousands of envs per call ----------------------------------------------------------------
venv = BlueSkyVectorEnv(env_id, num_envs, seed=0, autoreset_mode="same_step")
obs, infos = venv.reset()
rng = np.random.default_rng(0)
t0, steps = time.perf_counter(), 200
for _ in range(steps):
    actions = rng.uniform(-1, 1, (num_envs, venv.layout.act_dim)).astype(np.float32)
    obs, rewards, terminations, truncations, infos = venv.step(actions)
dt = time.perf_counter() - t0
print(f"{num_envs} envs x {steps} steps in {dt:.3f} s = {num_envs * steps / dt:.3e} env-steps/s through the numpy API")
venv.close()
