"""Drop-in alias: ``import bluesky_gym; bluesky_gym.register_envs()`` as in the reference (main.py:19),
served by the B200-native package."""
from bluesky_gym_sasha_b200 import register_envs  # noqa: F401
from bluesky_gym_sasha_b200.gym_compat import make  # noqa: F401
from . import utils  # noqa: F401  (bluesky_gym/__init__.py:2)
