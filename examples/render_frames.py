"""rgb_array frames: what the reference draws into its pygame window, as arrays (and PPM files you can open anywhere).

    python examples/render_frames.py [ENV_ID] [STEPS]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bluesky_gym

env_id = sys.argv[1] if len(sys.argv) > 1 else "SectorCREnv-v0"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20

bluesky_gym.register_envs()
env = bluesky_gym.make(env_id, render_mode="rgb_array")      # the mode the reference's metadata lists
env.reset(seed=0)
for t in range(steps):
    obs, reward, terminated, truncated, info = env.step(env.action_space.sample())
    if t % 10 == 0 or terminated or truncated:
        frame = env.render()                                  # (height, width, 3) uint8
        path = f"frame_{env_id}_{t:03d}.ppm"
        with open(path, "wb") as f:
            f.write(b"P6 %d %d 255\n" % (frame.shape[1], frame.shape[0]) + frame.tobytes())
        print(f"step {t}: frame {frame.shape}, {np.unique(frame.reshape(-1, 3), axis=0).shape[0]} colours -> {path}")
    if terminated or truncated:
        break
env.close()
