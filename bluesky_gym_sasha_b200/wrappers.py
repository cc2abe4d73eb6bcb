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


class WindFieldWrapper(_Wrapper):
    """wrappers/wind.py:8-64.  The wind field is handed to the simulator (``BlueSkyVectorEnv.set_wind``): kinematics,
    autopilot heading and the optional ``wind_u`` / ``wind_v`` observations (wind along / across the ownship heading,
    divided by MAX_WIND = 50) are all computed on the device.  The reference re-adds the points after every reset
    because ``bs.traf.reset()`` clears them; here the field simply stays on.

    Around a scalar env (``gym.make(id)``) the env's one-instance simulator is rebuilt with the wind (and, for
    ``augment_obs=True``, with the two extra observation keys).  A ``BlueSkyVectorEnv`` must have been constructed with
    ``wind_obs=augment_obs`` because the observation layout is fixed at construction."""

    def __init__(self, env, lat, lon, vnorth, veast, alt=None, augment_obs=False):
        super().__init__(env)
        self.lat, self.lon, self.vnorth, self.veast, self.alt = lat, lon, vnorth, veast, alt
        self.augment_obs = augment_obs
        wind = dict(lat=lat, lon=lon, vnorth=vnorth, veast=veast, alt=alt)
        if isinstance(env, BlueSkyVectorEnv):
            if env.wind_obs != bool(augment_obs):
                raise ValueError("construct the BlueSkyVectorEnv with wind_obs=%r (the observation layout is fixed at "
                                 "construction)" % bool(augment_obs))
            env.set_wind(**wind)
        else:
            scalar = getattr(env, "unwrapped", env)
            if not hasattr(scalar, "_make"):
                raise TypeError("WindFieldWrapper needs an accelerated bluesky_gym env")
            scalar._kw.update(wind=wind, wind_obs=bool(augment_obs))
            scalar.vec.close()
            scalar._make(scalar._seed)

    @property
    def observation_space(self):
        return self.env.observation_space

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def step(self, action):
        return self.env.step(action)
