"""Uniform-random policy statistics per env on the CUDA simulator (compare with the first episodes of the reference's logs)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bluesky_gym_sasha_b200.policy import evaluate
from bluesky_gym_sasha_b200.vector_env import BlueSkyVectorEnv

for env_id in ("DescentEnv-v0", "VerticalCREnv-v0", "HorizontalCREnv-v0", "SectorCREnv-v0", "StaticObstacleEnv-v0", "MergeEnv-v0", "PlanWaypointEnv-v0"):
    v = BlueSkyVectorEnv(env_id, 4096, seed=9, autoreset_mode="same_step")
    r = evaluate(v, None, episodes_per_env=1)
    print(f"{env_id:22s} len {r['lengths'].mean():6.1f} return {r['returns'].mean():9.3f} " +
          " ".join(f"{k[5:]}={val.mean():.3f}" for k, val in r.items() if k.startswith("info_") and k != "info_total_reward"))
    v.close()
