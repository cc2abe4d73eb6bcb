"""Drop-in alias of the reference module path ``bluesky_gym.wrappers.uncertainty``."""
from bluesky_gym_sasha_b200.wrappers import NoisyObservationWrapper  # noqa: F401
