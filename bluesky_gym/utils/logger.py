"""Drop-in alias of ``bluesky_gym.utils.logger`` (bluesky_gym/utils/logger.py:5-35): the CSV episode logger the
reference's training scripts hand to SB3 (main.py:17,29; scripts/multi_processing_example.py:19,32)."""
from bluesky_gym_sasha_b200.logger import CSVLoggerCallback  # noqa: F401
