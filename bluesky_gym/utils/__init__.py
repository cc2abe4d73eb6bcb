"""Drop-in alias of ``bluesky_gym.utils`` (bluesky_gym/utils/__init__.py:1): ``from bluesky_gym.utils import logger``."""
from . import logger  # noqa: F401
