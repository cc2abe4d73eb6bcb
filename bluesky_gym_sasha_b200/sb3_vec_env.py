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


class _LazyInfos(list):
    """SB3's per-env info dicts (``infos[i]``), built on first access.  With thousands of envs a Python loop that builds
    every dict on every step costs more than the simulation; SB3 itself only looks at the envs that finished
    (``terminal_observation``, ``TimeLimit.truncated``) and callbacks such as the reference's ``CSVLoggerCallback``
    at ``infos[0]`` (bluesky_gym/utils/logger.py:18,25).  Behaves like the list of dicts it stands for."""

    def __init__(self, infos, term, trunc, dones):
        super().__init__([None] * len(dones))
        self._src, self._term, self._trunc, self._dones = infos, term, trunc, dones
        self._keys = [k for k in infos if not k.startswith("_") and k != "final_obs"]

    def _build(self, i):
        src = self._src
        d = {k: (float(src[k][i]) if src[k].dtype.kind == "f" else int(src[k][i])) for k in self._keys}
        d["TimeLimit.truncated"] = bool(self._trunc[i] and not self._term[i])
        if self._dones[i] and "final_obs" in src:
            d["terminal_observation"] = {k: v[i].copy() for k, v in src["final_obs"].items()}
        return d

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        v = super().__getitem__(i)
        if v is None:
            v = self._build(i if i >= 0 else i + len(self))
            super().__setitem__(i, v)
        return v

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]


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
        self.venv.step_async(np.asarray(actions))        # enqueued on the device; returns without waiting

    def step_wait(self):
        obs, rew, term, trunc, infos = self.venv.step_wait()
        dones = term | trunc
        # copy=True: the arrays are already fresh; copy=False: views of the pinned ring, copied here as SB3 keeps them
        obs_out = obs if self.venv.copy else {k: v.copy() for k, v in obs.items()}
        if not self.venv.copy and "final_obs" in infos:      # views of a persistent buffer: captured now
            infos["final_obs"] = {k: v.copy() for k, v in infos["final_obs"].items()}
        return obs_out, rew.astype(np.float32), dones, _LazyInfos(infos, term, trunc, dones)

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
