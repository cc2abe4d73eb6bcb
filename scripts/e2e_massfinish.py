"""Cost of the steps in which EVERY env finishes at once (a fresh batch under a random policy hits the TimeLimit in lockstep):
per-step wall time of BlueSkyVectorEnv.step() with max_episode_steps=10, so every 10th step is such a step."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from bluesky_gym_sasha_b200.vector_env import BlueSkyVectorEnv

E = 4096
v = BlueSkyVectorEnv("HorizontalCREnv-v0", E, seed=0, cd_enabled=True, n_intruders=20, autoreset_mode="same_step", max_episode_steps=10)
v.reset()
a = np.random.default_rng(0).uniform(-1, 1, (E, 1)).astype(np.float32)
ts = []
for i in range(60):
    t0 = time.perf_counter()
    out = v.step(a)
    ts.append((time.perf_counter() - t0) * 1e6)
ts = np.array(ts[10:]).reshape(-1, 10)
print("per-step wall time by position in the 10-step episode [us] (median over 5 episodes):")
print(np.round(np.median(ts, axis=0), 1))
