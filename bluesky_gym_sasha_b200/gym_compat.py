"""gymnasium when it is installed, otherwise a minimal in-repo stand-in with the same surface.

The reference depends on gymnasium (pyproject.toml:7) for ``spaces``, ``Env``, ``register`` and
``make`` (bluesky_gym/__init__.py:1-46, horizontal_cr_env.py:8-9,49-62).  This image has no
gymnasium and no network, so the front-end must run against this shim here and against the real
package wherever it exists.  Only what the reference's call sites use is provided.
"""
from collections import OrderedDict

import numpy as np

try:                                            # pragma: no cover - not installed in the build image
    import gymnasium as _gym
    from gymnasium import spaces
    from gymnasium.envs.registration import register, registry
    from gymnasium.vector import VectorEnv
    from gymnasium.vector.utils import batch_space
    Env = _gym.Env
    Wrapper = _gym.Wrapper
    make = _gym.make
    HAVE_GYMNASIUM = True
except ImportError:
    HAVE_GYMNASIUM = False

    class _Space:
        def __init__(self, shape=None, dtype=None):
            self.shape = None if shape is None else tuple(shape)
            self.dtype = None if dtype is None else np.dtype(dtype)
            self._rng = np.random.default_rng()

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)

    class Box(_Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            if shape is None:
                if isinstance(low, np.ndarray):
                    shape = low.shape
                elif isinstance(high, np.ndarray):
                    shape = high.shape
                else:
                    shape = (1,)                # gymnasium's default for scalar bounds
            super().__init__(shape, dtype)
            self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()

        def sample(self):
            lo = np.where(np.isfinite(self.low), self.low, -1.0)
            hi = np.where(np.isfinite(self.high), self.high, 1.0)
            return self._rng.uniform(lo, hi, size=self.shape).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

    class Dict(_Space):
        def __init__(self, spaces_dict):
            super().__init__(None, None)
            self.spaces = OrderedDict(spaces_dict)

        def __getitem__(self, k):
            return self.spaces[k]

        def keys(self):
            return self.spaces.keys()

        def items(self):
            return self.spaces.items()

        def sample(self):
            return OrderedDict((k, s.sample()) for k, s in self.spaces.items())

        def contains(self, x):
            return isinstance(x, dict) and all(k in x and s.contains(x[k]) for k, s in self.spaces.items())

        def seed(self, seed=None):
            for i, s in enumerate(self.spaces.values()):
                s.seed(None if seed is None else seed + i)

        def __repr__(self):
            return "Dict(" + ", ".join(f"{k!r}: {s!r}" for k, s in self.spaces.items()) + ")"

    class _SpacesModule:
        pass

    spaces = _SpacesModule()
    spaces.Box = Box
    spaces.Dict = Dict
    spaces.Space = _Space

    def batch_space(space, n=1):
        if isinstance(space, Box):
            return Box(np.broadcast_to(space.low, (n,) + space.shape).copy(),
                       np.broadcast_to(space.high, (n,) + space.shape).copy(), dtype=space.dtype)
        if isinstance(space, Dict):
            return Dict(OrderedDict((k, batch_space(s, n)) for k, s in space.spaces.items()))
        raise TypeError(f"cannot batch {space!r}")

    class Env:
        metadata = {"render_modes": []}
        render_mode = None
        observation_space = None
        action_space = None
        spec = None

        def reset(self, *, seed=None, options=None):
            if seed is not None:
                self.np_random = np.random.default_rng(seed)
            return None, {}

        def step(self, action):
            raise NotImplementedError

        def render(self):
            return None

        def close(self):
            pass

        @property
        def unwrapped(self):
            return self

    class Wrapper(Env):
        """gymnasium.Wrapper: forwards everything to ``env``; spaces can be overridden by assignment."""

        def __init__(self, env):
            self.env = env
            self._observation_space = None
            self._action_space = None

        def __getattr__(self, name):
            if name.startswith("_"):
                raise AttributeError(name)
            return getattr(self.env, name)

        @property
        def observation_space(self):
            return self._observation_space if self._observation_space is not None else self.env.observation_space

        @observation_space.setter
        def observation_space(self, space):
            self._observation_space = space

        @property
        def action_space(self):
            return self._action_space if self._action_space is not None else self.env.action_space

        @action_space.setter
        def action_space(self, space):
            self._action_space = space

        def reset(self, **kwargs):
            return self.env.reset(**kwargs)

        def step(self, action):
            return self.env.step(action)

        def close(self):
            return self.env.close()

        @property
        def unwrapped(self):
            return self.env.unwrapped

    class VectorEnv:
        metadata = {}
        num_envs = 1
        single_observation_space = None
        single_action_space = None
        observation_space = None
        action_space = None
        render_mode = None
        closed = False

        def reset(self, *, seed=None, options=None):
            raise NotImplementedError

        def step(self, actions):
            raise NotImplementedError

        def close(self, **kwargs):
            self.closed = True

        @property
        def unwrapped(self):
            return self

    class _Spec:
        def __init__(self, id, entry_point, max_episode_steps, kwargs):
            self.id, self.entry_point, self.max_episode_steps, self.kwargs = id, entry_point, max_episode_steps, kwargs

    registry = {}

    def register(id, entry_point, max_episode_steps=None, **kwargs):
        registry[id] = _Spec(id, entry_point, max_episode_steps, kwargs.get("kwargs", {}))

    class TimeLimit:
        """gymnasium.wrappers.TimeLimit: truncated once elapsed steps reach the cap; reset clears it."""

        def __init__(self, env, max_episode_steps):
            self.env, self._max, self._t = env, max_episode_steps, 0

        def __getattr__(self, name):
            return getattr(self.env, name)

        def reset(self, **kw):
            self._t = 0
            return self.env.reset(**kw)

        def step(self, action):
            obs, r, term, trunc, info = self.env.step(action)
            self._t += 1
            if self._t >= self._max:
                trunc = True
            return obs, r, term, trunc, info

        @property
        def unwrapped(self):
            return self.env.unwrapped

    def make(id, **kwargs):
        import importlib
        if id not in registry:
            raise KeyError(f"environment id {id!r} is not registered; call register_envs() first")
        spec = registry[id]
        ep = spec.entry_point
        if isinstance(ep, str):
            mod, cls = ep.split(":")
            ep = getattr(importlib.import_module(mod), cls)
        env = ep(**{**spec.kwargs, **kwargs})
        env.spec = spec
        if spec.max_episode_steps is not None:
            env = TimeLimit(env, spec.max_episode_steps)
        return env
