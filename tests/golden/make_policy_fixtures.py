#!/usr/bin/env python
"""Extracts, from the reference's shipped training artefacts (scripts/common/results/), what the
policy-in-the-loop tests need and the GPU box cannot read (it has no /root/reference):

* ``policies/<EnvId>_<ALGO>.npz`` -- the deterministic ACTOR of the shipped SB3 model.zip (weights, biases,
  activations; critics / optimisers are dropped), float32 exactly as stored;
* ``policies/log_stats.json`` -- per (env, algo) CSV training log: number of episodes, and over the last 10 % of
  the episodes (the converged policy) the mean / std of every logged column and of the episode length.

    python tests/golden/make_policy_fixtures.py        # build container only
"""
import csv
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
RES = "/root/reference/scripts/common/results"
# the models kept as fixtures (PPO actors are 64x64, SAC 256x256, DDPG 400x300: keep the repo small)
KEEP = [("DescentEnv-v0", "SAC"), ("DescentEnv-v0", "DDPG"), ("PlanWaypointEnv-v0", "SAC"), ("HorizontalCREnv-v0", "PPO"),
        ("HorizontalCREnv-v0", "SAC"), ("VerticalCREnv-v0", "PPO"), ("VerticalCREnv-v0", "SAC"), ("SectorCREnv-v0", "PPO"),
        ("StaticObstacleEnv-v0", "PPO"), ("MergeEnv-v0", "PPO"), ("MergeEnv-v0", "SAC")]


def log_stats(path):
    rows = list(csv.DictReader(open(path)))
    n = len(rows)
    k = max(50, n // 10)
    tail = rows[-k:]
    ts = np.array([float(r["timesteps"]) for r in rows[-k - 1:]])
    out = {"episodes": n, "tail_episodes": k, "length_mean": float(np.diff(ts).mean())}
    for col in rows[0]:
        if col in ("timesteps", "episodes"):
            continue
        try:
            v = np.array([float(r[col]) for r in tail])
        except ValueError:
            v = np.array([1.0 if r[col] == "True" else 0.0 for r in tail])
        out[col + "_mean"] = float(np.nanmean(v))
        out[col + "_std"] = float(np.nanstd(v))
    first = np.array([float(r["total_reward"]) for r in rows[:200]])
    out["total_reward_first200_mean"] = float(first.mean())
    return out


def main():
    if not os.path.isdir(RES):
        raise SystemExit("needs the reference at /root/reference (build container only)")
    from bluesky_gym_sasha_b200.policy import read_sb3_zip, save_actor_npz
    from bluesky_gym_sasha_b200.spec import SPECS
    os.makedirs(os.path.join(HERE, "policies"), exist_ok=True)
    stats = {}
    for env in sorted(os.listdir(os.path.join(RES, "logs_backup"))):
        for f in sorted(os.listdir(os.path.join(RES, "logs_backup", env))):
            algo = f[:-4].split("_")[-1]
            stats[f"{env}_{algo}"] = log_stats(os.path.join(RES, "logs_backup", env, f))
    json.dump(stats, open(os.path.join(HERE, "policies", "log_stats.json"), "w"), indent=1, sort_keys=True)
    for env, algo in KEEP:
        layers, meta = read_sb3_zip(os.path.join(RES, "models_backup", env, f"{env}_{algo}", "model.zip"))
        keys = sorted(k for k, *_ in SPECS[env].obs_keys)
        save_actor_npz(os.path.join(HERE, "policies", f"{env}_{algo}.npz"), layers, keys)
        print(env, algo, [w.shape for w, _, _ in layers], "trained for", meta.get("num_timesteps"), "steps")


if __name__ == "__main__":
    main()
