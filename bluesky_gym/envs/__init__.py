"""``import bluesky_gym.envs`` (scripts/multi_processing_example.py:17) -> the accelerated env classes."""
from bluesky_gym_sasha_b200.envs import *  # noqa: F401,F403
