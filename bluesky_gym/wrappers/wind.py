"""Drop-in alias of the reference module path ``bluesky_gym.wrappers.wind``."""
from bluesky_gym_sasha_b200.wrappers import WindFieldWrapper  # noqa: F401

MAX_WIND = 50  # wrappers/wind.py:6
