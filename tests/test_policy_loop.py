"""Policy-in-the-loop pins (SURVEY.md section 8c "weak pins", 8f-2): the policies the reference SHIPS were
trained on the real BlueSky simulator; their training logs record what they achieve there.  Run on this
repo's simulator (CPU oracle / CUDA kernels) they must achieve the same: a simulator whose dynamics, observation
layout or reward differed from the reference's would make them crash, wander or time out.

Fixtures (tests/golden/policies/, made by tests/golden/make_policy_fixtures.py from
/root/reference/scripts/common/results): the deterministic actor of each model.zip and, per CSV log, the mean / std
of the logged columns over the last 10 % of training episodes.  Training episodes are sampled from the
STOCHASTIC policy, evaluation here is deterministic, so the evaluation may be somewhat better than the log:
accepted band  log_mean - 0.5 sigma <= mean return <= log_mean + 1.25 sigma  (sigma = episode-to-episode std in
the log), episode length within 12 % where episodes end by reaching a goal, and the improvement over a uniform
random policy must be at least 70 % of the improvement the log shows between its first 200 and last 10 % episodes.
"""
import json
import os
import random

import numpy as np
import pytest

from bluesky_gym_sasha_b200.policy import load_actor_npz
from bluesky_gym_sasha_b200.spec import SPECS

POL = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "policies")
STATS = json.load(open(os.path.join(POL, "log_stats.json")))
# (env, algo, compare episode length?)   MergeEnv: models exist but the reference ships no PPO / SAC log for it
CASES = [("DescentEnv-v0", "SAC", True), ("DescentEnv-v0", "DDPG", True), ("PlanWaypointEnv-v0", "SAC", True),
         ("HorizontalCREnv-v0", "SAC", True), ("HorizontalCREnv-v0", "PPO", True), ("VerticalCREnv-v0", "PPO", True),
         ("VerticalCREnv-v0", "SAC", True), ("StaticObstacleEnv-v0", "PPO", False), ("SectorCREnv-v0", "PPO", False)]


def _np_actor(layers, obs_keys_sorted):
    def f(obs):
        x = np.concatenate([np.asarray(obs[k], dtype=np.float64).reshape(-1) for k in obs_keys_sorted])
        for w, b, a in layers:
            x = w.astype(np.float64) @ x + b
            x = np.maximum(x, 0) if a == "relu" else (np.tanh(x) if a == "tanh" else np.clip(x, -1, 1))
        return x
    return f


def _band(ret_mean, st):
    lo = st["total_reward_mean"] - 0.5 * st["total_reward_std"]
    hi = st["total_reward_mean"] + 1.25 * st["total_reward_std"]
    return lo <= ret_mean <= hi, (lo, hi)


def test_fixture_actor_matches_torch_module():
    """npz round trip and the sorted-key permutation: SB3Actor(flat obs in declaration order) == numpy actor(dict)."""
    import torch
    from bluesky_gym_sasha_b200.policy import SB3Actor
    layers, keys = load_actor_npz(os.path.join(POL, "HorizontalCREnv-v0_PPO.npz"))
    layout, dim = SPECS["HorizontalCREnv-v0"].obs_layout(5)
    assert keys == sorted(layout) and dim == 28
    actor = SB3Actor(layers, layout, "cpu")
    rng = np.random.default_rng(0)
    flat = rng.normal(size=(7, dim)).astype(np.float32)
    out = actor(torch.as_tensor(flat)).numpy()
    ref = _np_actor(layers, keys)
    for e in range(7):
        obs = {k: flat[e, off:off + w] for k, (off, w, _, _) in layout.items()}
        np.testing.assert_allclose(out[e], ref(obs), rtol=1e-4, atol=1e-5)


def test_shipped_descent_policy_on_oracle():
    """CPU: the shipped DescentEnv SAC actor on the float64 oracle env (12 episodes)."""
    from oracle import envs as oenvs
    layers, keys = load_actor_npz(os.path.join(POL, "DescentEnv-v0_SAC.npz"))
    pi = _np_actor(layers, keys)
    st = STATS["DescentEnv-v0_SAC"]
    np.random.seed(0)
    random.seed(0)
    rets, lens = [], []
    for _ in range(12):
        env = oenvs.DescentEnv()
        obs, _ = env.reset()
        tot = 0.0
        for t in range(300):
            obs, r, term, trunc, _ = env.step(pi(obs))
            tot += r
            if term or trunc:
                break
        rets.append(tot)
        lens.append(t + 1)
    ok, band = _band(np.mean(rets), st)
    print(f"oracle DescentEnv SAC: return {np.mean(rets):.2f} (log {st['total_reward_mean']:.2f}), length {np.mean(lens):.1f} (log {st['length_mean']:.1f})")
    assert ok, (np.mean(rets), band)
    assert abs(np.mean(lens) - st["length_mean"]) < 0.12 * st["length_mean"]
    assert max(rets) > -100 and min(rets) > -100            # never crashes (-100 would be a ground impact)


@pytest.mark.gpu
@pytest.mark.parametrize("env_id,algo,check_len", CASES)
def test_shipped_policy_on_cuda_env(cuda, env_id, algo, check_len):
    from bluesky_gym_sasha_b200.policy import SB3Actor, evaluate
    from bluesky_gym_sasha_b200.vector_env import BlueSkyVectorEnv
    E = 2048
    st = STATS[f"{env_id}_{algo}"]
    venv = BlueSkyVectorEnv(env_id, E, seed=123, autoreset_mode="same_step")
    actor = SB3Actor.from_npz(os.path.join(POL, f"{env_id}_{algo}.npz"), venv)
    res = evaluate(venv, actor, episodes_per_env=1)
    rnd = evaluate(venv, None, episodes_per_env=1)
    venv.close()
    ret, length, ret_rnd = res["returns"].mean(), res["lengths"].mean(), rnd["returns"].mean()
    print(f"{env_id} {algo}: return {ret:.3f} +- {res['returns'].std():.3f} (log {st['total_reward_mean']:.3f} +- "
          f"{st['total_reward_std']:.3f}), length {length:.1f} (log {st['length_mean']:.1f}), random policy {ret_rnd:.3f} "
          f"(log first 200 episodes {st['total_reward_first200_mean']:.3f}), {len(res['returns'])} episodes")
    assert len(res["returns"]) == E
    ok, band = _band(ret, st)
    assert ok, (ret, band)
    if check_len:
        assert abs(length - st["length_mean"]) < 0.12 * st["length_mean"], (length, st["length_mean"])
    learned = st["total_reward_mean"] - st["total_reward_first200_mean"]
    assert ret - ret_rnd > 0.7 * learned, (ret, ret_rnd, learned)


@pytest.mark.gpu
def test_shipped_merge_policy_completes_the_task(cuda):
    """MergeEnv: the reference ships PPO / SAC models but logs only for DDPG / TD3 runs that never learned (return ~ -9.2,
    faf_reach 0).  The shipped PPO actor, trained on the real BlueSky, must complete the task here: pass the FAF and reach
    the runway (faf_reach == 2: merge_env.py:246-284) in nearly every episode, far above a random policy."""
    from bluesky_gym_sasha_b200.policy import SB3Actor, evaluate
    from bluesky_gym_sasha_b200.vector_env import BlueSkyVectorEnv
    venv = BlueSkyVectorEnv("MergeEnv-v0", 2048, seed=123, autoreset_mode="same_step")
    actor = SB3Actor.from_npz(os.path.join(POL, "MergeEnv-v0_PPO.npz"), venv)
    res = evaluate(venv, actor, episodes_per_env=1)
    rnd = evaluate(venv, None, episodes_per_env=1)
    venv.close()
    print(f"MergeEnv PPO: return {res['returns'].mean():.3f}, faf_reach {res['info_faf_reach'].mean():.3f}, length "
          f"{res['lengths'].mean():.1f}; random policy return {rnd['returns'].mean():.3f}, faf_reach {rnd['info_faf_reach'].mean():.3f}")
    assert res["info_faf_reach"].mean() > 1.9 and res["lengths"].mean() < 35
    assert res["returns"].mean() > -4.0 and rnd["returns"].mean() < -8.0 and rnd["info_faf_reach"].mean() < 0.5
