"""bluesky_gym_sasha_b200 -- B200-native batched BlueSky-Gym step path (sm_100a CUDA behind a C ABI).

Drop-in surface of the reference package: ``register_envs()`` and the gym ids; plus the batched
``BlueSkyVectorEnv``, the SB3 ``VecEnv`` adapter and the single-airspace ``StateBasedCD`` detector.
Importing this package never needs a GPU; constructing an env or detector does (no CPU fallback).
"""
from .registration import register_envs
from .spec import SPECS

__all__ = ["register_envs", "SPECS", "BlueSkyVectorEnv", "StateBasedCD", "BlueSkySB3VecEnv", "make_vec"]


def __getattr__(name):          # lazy: these import torch
    if name == "BlueSkyVectorEnv":
        from .vector_env import BlueSkyVectorEnv
        return BlueSkyVectorEnv
    if name == "StateBasedCD":
        from .cd import StateBasedCD
        return StateBasedCD
    if name == "BlueSkySB3VecEnv":
        from .sb3_vec_env import BlueSkySB3VecEnv
        return BlueSkySB3VecEnv
    if name == "make_vec":
        from .vector_env import BlueSkyVectorEnv
        return lambda env_id, num_envs, **kw: BlueSkyVectorEnv(env_id, num_envs, **kw)
    raise AttributeError(name)
