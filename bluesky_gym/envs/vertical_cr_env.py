"""Drop-in alias of the reference module path ``bluesky_gym.envs.vertical_cr_env`` (the registration's entry point,
bluesky_gym/__init__.py:6-46) -> the accelerated env class."""
from bluesky_gym_sasha_b200.envs import VerticalCREnv  # noqa: F401
