"""The reference's environment wrappers (bluesky_gym/wrappers/) for the accelerated envs.

``NoisyObservationWrapper`` (wrappers/uncertainty.py:4-31) adds N(0, noise_level) to every observation element.
  * Around a scalar env (``gym.make(id)``) it does exactly what the reference does, on the host, with the
    process-global ``np.random`` -- same draws for the same seed.
  * Around a ``BlueSkyVectorEnv`` the noise is generated on the device (``bsg_set_obs_noise``: one elementwise
    kernel per call, Philox stream keyed by seed / global env id / call index), so ``step_torch`` observations are
    noisy too and nothing extra crosses PCIe.
"""
import numpy as np

from .vector_env import BlueSkyVectorEnv


class _Wrapper:
    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return getattr(self.env, "unwrapped", self.env)

    def close(self):
        return self.env.close()


class NoisyObservationWrapper(_Wrapper):
    def __init__(self, env, noise_level=0.1):
        super().__init__(env)
        self.noise_level = noise_level
        self._on_device = isinstance(env, BlueSkyVectorEnv)
        if self._on_device:
            env.set_obs_noise(noise_level)

    def reset(self, **kwargs):
        observation, info = self.env.reset(**kwargs)
        return (observation if self._on_device else self.add_noise(observation)), info

    def step(self, action):
        observation, reward, done, truncated, info = self.env.step(action)
        return (observation if self._on_device else self.add_noise(observation)), reward, done, truncated, info

    def add_noise(self, observation):                       # uncertainty.py:19-31
        if isinstance(observation, np.ndarray):
            return observation + np.random.normal(0, self.noise_level, size=observation.shape)
        if isinstance(observation, dict):
            return {key: (value + np.random.normal(0, self.noise_level, size=value.shape)
                          if isinstance(value, np.ndarray) else value) for key, value in observation.items()}
        print('observation not an numpy array or dictionary, return observation unaltered')
        return observation
