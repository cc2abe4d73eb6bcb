"""Evaluate a Stable-Baselines3 model.zip trained on BlueSky-Gym (e.g. the ones the reference ships under
scripts/common/results/models_backup/<env>/<env>_<ALGO>/model.zip) on the batched simulator, on the device.

    python examples/eval_sb3_model.py HorizontalCREnv-v0 path/to/model.zip [NUM_ENVS]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # run from a checkout
from bluesky_gym_sasha_b200 import BlueSkyVectorEnv
from bluesky_gym_sasha_b200.policy import SB3Actor, evaluate

env_id, path = sys.argv[1], sys.argv[2]
num_envs = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
venv = BlueSkyVectorEnv(env_id, num_envs, seed=0, autoreset_mode="same_step")
actor = SB3Actor.from_zip(path, venv)            # deterministic actor; SB3 itself is not needed
res = evaluate(venv, actor, episodes_per_env=1)
print(f"{env_id}: {len(res['returns'])} episodes, return {res['returns'].mean():.3f} +- {res['returns'].std():.3f}, "
      f"length {res['lengths'].mean():.1f}")
for k, v in res.items():
    if k.startswith("info_"):
        print(f"  {k[5:]:>22s}: {v.mean():.4f}")
venv.close()
