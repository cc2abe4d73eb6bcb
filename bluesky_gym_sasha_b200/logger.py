"""``CSVLoggerCallback`` -- the episode logger of the reference's training scripts (bluesky_gym/utils/logger.py:5-35),
kept so that ``from bluesky_gym.utils import logger`` works unchanged on top of the accelerated envs.

Behaviour mirrored: the CSV header is ``timesteps, episodes`` followed by the keys of ``infos[0]`` at the first
callback; one row is appended whenever env 0 finishes an episode (``dones[0]``), holding the SB3 step count, the
running episode count and env 0's info values.  With SB3 installed the class is a ``BaseCallback``; without it (this
image) it is a plain object with the same ``_on_step`` contract, so the adapter tests can drive it directly.
"""
import csv
import os

try:                                            # pragma: no cover - SB3 is not installed in the build image
    from stable_baselines3.common.callbacks import BaseCallback as _Base
    HAVE_SB3 = True
except ImportError:
    HAVE_SB3 = False

    class _Base:
        def __init__(self, verbose=0):
            self.verbose, self.num_timesteps, self.locals = verbose, 0, {}

        def on_step(self):
            self.num_timesteps += 1
            return self._on_step()


class CSVLoggerCallback(_Base):
    def __init__(self, log_dir, file_name="training_log.csv", verbose=0):
        super().__init__(verbose)
        os.makedirs(log_dir, exist_ok=True)
        self.log_dir = log_dir
        self.log_file = os.path.join(log_dir, file_name)
        self.headers = ["timesteps", "episodes"]
        self.initialized = False
        self.episode_count = 0

    def _append(self, row, mode):
        with open(self.log_file, mode=mode, newline="") as f:
            csv.writer(f).writerow(row)

    def _on_step(self) -> bool:
        info0 = self.locals["infos"][0]
        if not self.initialized:                # columns = whatever env 0 reports (the reference's info keys)
            self.info_keys = list(info0.keys())
            self.headers.extend(self.info_keys)
            self._append(self.headers, "w")
            self.initialized = True
        if self.locals["dones"][0]:
            self.episode_count += 1
            self._append([self.num_timesteps, self.episode_count] + [info0.get(k, None) for k in self.info_keys], "a")
        return True
