"""Shipped policies in the loop under alternative values of recalled performance data (run on a GPU box)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bluesky_gym_sasha_b200.policy import SB3Actor, evaluate
from bluesky_gym_sasha_b200.vector_env import BlueSkyVectorEnv

POL = os.path.join(ROOT, "tests", "golden", "policies")
STATS = json.load(open(os.path.join(POL, "log_stats.json")))
cases = [("SectorCREnv-v0", "PPO"), ("StaticObstacleEnv-v0", "PPO"), ("VerticalCREnv-v0", "SAC"), ("DescentEnv-v0", "SAC")]
variants = [dict(), dict(vmaxer=170.0), dict(vmaxer=180.0), dict(vmaxer=190.0), dict(vmaxer=200.0), dict(vmaxer=180.0, axmax_air=2.0),
            dict(vminer=90.0), dict(vminer=110.0)]
if len(sys.argv) > 1:
    variants = [json.loads(a) for a in sys.argv[1:]]
for env_id, algo in cases:
    st = STATS[f"{env_id}_{algo}"]
    print(f"{env_id} {algo}: log return {st['total_reward_mean']:.3f} +- {st['total_reward_std']:.3f}, length {st['length_mean']:.1f}, "
          + ", ".join(f"{k[:-5]} {v:.3f}" for k, v in st.items() if k.endswith("_mean") and k not in ("total_reward_mean", "length_mean", "total_reward_first200_mean", "TimeLimit.truncated_mean")))
    for pv in variants:
        venv = BlueSkyVectorEnv(env_id, 4096, seed=123, autoreset_mode="same_step", perf=pv)
        actor = SB3Actor.from_npz(os.path.join(POL, f"{env_id}_{algo}.npz"), venv)
        res = evaluate(venv, actor, episodes_per_env=1)
        extra = ", ".join(f"{k[5:]} {v.mean():.3f}" for k, v in res.items() if k.startswith("info_") and k != "info_total_reward")
        print(f"   perf {str(pv):22s} return {res['returns'].mean():8.3f} +- {res['returns'].std():7.3f}  length {res['lengths'].mean():6.1f}  {extra}")
        venv.close()
