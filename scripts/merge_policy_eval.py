import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bluesky_gym_sasha_b200.policy import SB3Actor, evaluate
from bluesky_gym_sasha_b200.vector_env import BlueSkyVectorEnv
POL = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "policies")
for algo in ("PPO", "SAC"):
    venv = BlueSkyVectorEnv("MergeEnv-v0", 2048, seed=123, autoreset_mode="same_step")
    actor = SB3Actor.from_npz(os.path.join(POL, f"MergeEnv-v0_{algo}.npz"), venv)
    res = evaluate(venv, actor, episodes_per_env=1)
    rnd = evaluate(venv, None, episodes_per_env=1)
    print(algo, "return %.3f +- %.3f len %.1f" % (res["returns"].mean(), res["returns"].std(), res["lengths"].mean()),
          {k: round(float(v.mean()), 3) for k, v in res.items() if k.startswith("info_")}, "random %.3f" % rnd["returns"].mean())
