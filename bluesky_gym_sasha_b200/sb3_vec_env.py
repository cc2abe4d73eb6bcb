"""Stable-Baselines3 ``VecEnv`` adapter over BlueSkyVectorEnv.

Contract mirrored (the reference trains through SB3: main.py:36-55, scripts/multi_processing_example.py:50-56):
same-step autoreset, terminal observation in ``infos[i]["terminal_observation"]``, the time-limit flag in
``infos[i]["TimeLimit.truncated"]`` (the column present in every shipped CSV log), Dict observations as
``dict[str, np.ndarray[E, k]]``, and per-env info dicts carrying the reference's info keys on every step so
``CSVLoggerCallback`` (bluesky_gym/utils/logger.py:15-35) keeps working.  SB3 is not installed in this
image; when it is, the class derives from ``stable_baselines3.common.vec_env.VecEnv``.
"""
import numpy as np

try:                                            # pragma: no cover
    from stable_baselines3.common.vec_env import VecEnv as _Base
    HAVE_SB3 = True
except ImportError:
    _Base = object
    HAVE_SB3 = False


class BlueSkySB3VecEnv(_Base):
    def __init__(self, venv):
        assert venv.autoreset_mode == "same_step", "SB3 expects same-step autoreset"
        self.venv = venv
        if HAVE_SB3:                            # pragma: no cover
            super().__init__(venv.num_envs, venv.single_observation_space, venv.single_action_space)
        else:
            self.num_envs = venv.num_envs
            self.observation_space = venv.single_observation_space
            self.action_space = venv.single_action_space
        self.render_mode = None
        self._actions = None

    def reset(self):
        obs, _ = self.venv.reset()
        return {k: v.copy() for k, v in obs.items()}

    def step_async(self, actions):
        self._actions = np.asarray(actions)

    def step_wait(self):
        obs, rew, term, trunc, infos = self.venv.step(self._actions)
        dones = term | trunc
        keys = [k for k in infos if not k.startswith("_") and k != "final_obs"]
        out = []
        for i in range(self.num_envs):
            d = {k: (float(infos[k][i]) if infos[k].dtype.kind == "f" else int(infos[k][i])) for k in keys}
            d["TimeLimit.truncated"] = bool(trunc[i] and not term[i])
            if dones[i] and "final_obs" in infos:
                d["terminal_observation"] = {k: v[i].copy() for k, v in infos["final_obs"].items()}
            out.append(d)
        return {k: v.copy() for k, v in obs.items()}, rew.astype(np.float32), dones, out

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self):
        self.venv.close()

    def seed(self, seed=None):
        return [None] * self.num_envs

    def get_attr(self, attr_name, indices=None):
        n = self.num_envs if indices is None else len(list(indices))
        return [getattr(self.venv, attr_name)] * n

    def set_attr(self, attr_name, value, indices=None):
        setattr(self.venv, attr_name, value)

    def env_method(self, method_name, *args, indices=None, **kwargs):
        return [getattr(self.venv, method_name)(*args, **kwargs)]

    def env_is_wrapped(self, wrapper_class, indices=None):
        n = self.num_envs if indices is None else len(list(indices))
        return [False] * n
