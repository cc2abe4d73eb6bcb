"""``register_envs()``: the seven gym ids of the reference (bluesky_gym/__init__.py:4-46), same caps."""
from .gym_compat import register, registry
from .spec import SPECS


def register_envs():
    for env_id, spec in SPECS.items():
        if env_id not in registry:
            register(id=env_id, entry_point=spec.entry_point, max_episode_steps=spec.max_episode_steps)
