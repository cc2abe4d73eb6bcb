"""The reference's environment wrappers (bluesky_gym/wrappers/) for the accelerated envs.

``NoisyObservationWrapper`` (wrappers/uncertainty.py:4-31) adds N(0, noise_level) to every observation element.
  * Around a scalar env (``gym.make(id)``) it does exactly what the reference does, on the host, with the
    process-global ``np.random`` -- same draws for the same seed.
  * Around a ``BlueSkyVectorEnv`` the noise is generated on the device (``bsg_set_obs_noise``: one elementwise
    kernel per call, Philox stream keyed by seed / global env id / call index), so ``step_torch`` observations are
    noisy too and nothing extra crosses PCIe.
"""
import numpy as np

from .gym_compat import Wrapper, spaces
from .vector_env import BlueSkyVectorEnv


class _VecProxy:
    """Thin forwarding proxy used when the wrapped object is a ``BlueSkyVectorEnv`` (a ``VectorEnv`` is not a
    ``gymnasium.Env``, so ``gymnasium.Wrapper`` cannot hold it)."""

    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return getattr(self.env, "unwrapped", self.env)

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def step(self, action):
        return self.env.step(action)

    def close(self):
        return self.env.close()


class NoisyObservationWrapper(Wrapper):
    """wrappers/uncertainty.py:4-31.  A ``gymnasium.Wrapper`` around a scalar env (so ``Monitor``, ``DummyVecEnv``,
    ``check_env`` accept it like the reference's); around a ``BlueSkyVectorEnv`` the constructor returns the device
    form (``VecNoisyObservation``)."""

    def __new__(cls, env, *args, **kwargs):
        if isinstance(env, BlueSkyVectorEnv):
            return VecNoisyObservation(env, *args, **kwargs)
        return super().__new__(cls)

    def __init__(self, env, noise_level=0.1):
        super().__init__(env)
        self.noise_level = noise_level

    def reset(self, **kwargs):
        observation, info = self.env.reset(**kwargs)
        return self.add_noise(observation), info

    def step(self, action):
        observation, reward, done, truncated, info = self.env.step(action)
        return self.add_noise(observation), reward, done, truncated, info

    def add_noise(self, observation):                       # uncertainty.py:19-31
        if isinstance(observation, np.ndarray):
            return observation + np.random.normal(0, self.noise_level, size=observation.shape)
        if isinstance(observation, dict):
            return {key: (value + np.random.normal(0, self.noise_level, size=value.shape)
                          if isinstance(value, np.ndarray) else value) for key, value in observation.items()}
        print('observation not an numpy array or dictionary, return observation unaltered')
        return observation


class VecNoisyObservation(_VecProxy):
    """``NoisyObservationWrapper`` over the batched env: the noise is added on the device (``bsg_set_obs_noise``)."""

    def __init__(self, env, noise_level=0.1):
        super().__init__(env)
        self.noise_level = noise_level
        env.set_obs_noise(noise_level)


class WindFieldWrapper(Wrapper):
    """wrappers/wind.py:8-64.  The wind field is handed to the simulator (``BlueSkyVectorEnv.set_wind``): kinematics,
    autopilot heading and the optional ``wind_u`` / ``wind_v`` observations (wind along / across the ownship heading,
    divided by MAX_WIND = 50) are all computed on the device.  The reference re-adds the points after every reset
    because ``bs.traf.reset()`` clears them; here the field simply stays on.

    Around a scalar env (``gym.make(id)``) this is a ``gymnasium.Wrapper``: the env's one-instance simulator is rebuilt
    with the wind (and, for ``augment_obs=True``, with the two extra observation keys, which the wrapper also declares
    in its ``observation_space`` like the reference, wind.py:18-24).  Around a ``BlueSkyVectorEnv`` -- which must have
    been constructed with ``wind_obs=augment_obs`` because the observation layout is fixed at construction -- the
    constructor returns the device form (``VecWindField``)."""

    def __new__(cls, env, *args, **kwargs):
        if isinstance(env, BlueSkyVectorEnv):
            return VecWindField(env, *args, **kwargs)
        return super().__new__(cls)

    def __init__(self, env, lat, lon, vnorth, veast, alt=None, augment_obs=False):
        super().__init__(env)
        self.lat, self.lon, self.vnorth, self.veast, self.alt = lat, lon, vnorth, veast, alt
        self.augment_obs = augment_obs
        scalar = getattr(env, "unwrapped", env)
        if not hasattr(scalar, "_make"):
            raise TypeError("WindFieldWrapper needs an accelerated bluesky_gym env")
        scalar._kw.update(wind=dict(lat=lat, lon=lon, vnorth=vnorth, veast=veast, alt=alt), wind_obs=bool(augment_obs))
        scalar.vec.close()
        scalar._make(scalar._seed)
        if self.augment_obs:
            base = scalar.observation_space
            assert isinstance(base, spaces.Dict), "This wrapper only supports Dict observation spaces."
            self.observation_space = spaces.Dict({
                **{k: v for k, v in base.spaces.items() if k not in ("wind_u", "wind_v")},
                "wind_u": spaces.Box(-np.inf, np.inf, shape=(1,), dtype=np.float64),
                "wind_v": spaces.Box(-np.inf, np.inf, shape=(1,), dtype=np.float64)})

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def step(self, action):
        return self.env.step(action)


class VecWindField(_VecProxy):
    """``WindFieldWrapper`` over the batched env (``bsg_set_wind``)."""

    def __init__(self, env, lat, lon, vnorth, veast, alt=None, augment_obs=False):
        super().__init__(env)
        self.lat, self.lon, self.vnorth, self.veast, self.alt = lat, lon, vnorth, veast, alt
        self.augment_obs = augment_obs
        if env.wind_obs != bool(augment_obs):
            raise ValueError("construct the BlueSkyVectorEnv with wind_obs=%r (the observation layout is fixed at "
                             "construction)" % bool(augment_obs))
        env.set_wind(lat=lat, lon=lon, vnorth=vnorth, veast=veast, alt=alt)
