"""Scalar ``gymnasium.Env`` views of the batched simulator: the classes ``gym.make(id)`` returns.

Same constructor keywords, spaces, return types and info keys as the reference classes
(bluesky_gym/envs/{descent,horizontal_cr,sector_cr,merge,plan_waypoint,vertical_cr,static_obstacle}_env.py); each is a ``num_envs=1``
BlueSkyVectorEnv with autoreset disabled, so ``reset`` / ``step`` follow the single-env contract
(float64 numpy observations, python scalars for reward / terminated / truncated).
"""
import numpy as np

from .gym_compat import Env
from .spec import SPECS
from .vector_env import BlueSkyVectorEnv


class _ScalarEnv(Env):
    ENV_ID = None
    metadata = {"render_modes": ["rgb_array", "human"], "render_fps": 120}   # e.g. horizontal_cr_env.py:39

    def __init__(self, render_mode=None, device=0, seed=0, cd_enabled=False, **kwargs):
        assert render_mode is None or render_mode in self.metadata["render_modes"]   # horizontal_cr_env.py:64
        if render_mode == "human":
            raise NotImplementedError("render_mode='human' needs a pygame window; use render_mode='rgb_array' (frames from render())")
        self.render_mode = render_mode
        self._seed = seed
        self._kw = dict(device=device, cd_enabled=cd_enabled, **kwargs)
        self._make(seed)

    def _make(self, seed):
        # the registration's TimeLimit is applied by gym.make's wrapper, exactly like the reference
        self.vec = BlueSkyVectorEnv(self.ENV_ID, 1, seed=seed, autoreset_mode="disabled",
                                    max_episode_steps=0, obs_dtype=np.float64, **self._kw)
        self.observation_space = self.vec.single_observation_space
        self.action_space = self.vec.single_action_space

    def _info(self, infos):
        return {k: (float(v[0]) if np.issubdtype(np.asarray(v).dtype, np.floating) else int(v[0]))
                for k, v in infos.items() if not k.startswith("_") and k != "final_obs"}

    def reset(self, seed=None, options=None):
        if seed is not None:
            self._seed = seed
        obs, infos = self.vec.reset(seed=seed)               # re-keys the Philox stream when a seed is given
        self._check_reset_flags()
        return {k: v[0].copy() for k, v in obs.items()}, self._info(infos)

    def _check_reset_flags(self):
        """The scenario generators run on the device; conditions under which the reference raises are recorded in the
        env record (BSG_I32_RESET_FLAGS) and turned back into the reference's exception here."""
        flags = int(self.vec.reset_flags()[0])
        if self.ENV_ID == "StaticObstacleEnv-v0" and flags & 4:               # static_obstacle_env.py:215-216
            raise Exception("No waypoints can be generated outside the obstacles. Check the parameters of the obstacles "
                            "in the definition of the scenario.")

    def step(self, action):
        a = np.asarray(action, dtype=np.float64).reshape(1, -1)
        obs, rew, term, trunc, infos = self.vec.step(a)
        return ({k: v[0].copy() for k, v in obs.items()}, float(rew[0]), bool(term[0]), bool(trunc[0]),
                self._info(infos))

    def render(self):
        """``render_mode="rgb_array"``: the frame the reference's ``_render_frame`` draws, as a (height, width, 3) uint8 array."""
        return self.vec.render(0) if self.render_mode == "rgb_array" else None

    def close(self):
        self.vec.close()


class DescentEnv(_ScalarEnv):
    ENV_ID = "DescentEnv-v0"


class HorizontalCREnv(_ScalarEnv):
    ENV_ID = "HorizontalCREnv-v0"


class SectorCREnv(_ScalarEnv):
    ENV_ID = "SectorCREnv-v0"

    def __init__(self, render_mode=None, ac_density_mode="normal", **kw):      # sector_cr_env.py:45
        super().__init__(render_mode=render_mode, ac_density_mode=ac_density_mode, **kw)


class MergeEnv(_ScalarEnv):
    ENV_ID = "MergeEnv-v0"


class StaticObstacleEnv(_ScalarEnv):
    ENV_ID = "StaticObstacleEnv-v0"


class PlanWaypointEnv(_ScalarEnv):                  # plan_waypoint_env.py:36-76
    ENV_ID = "PlanWaypointEnv-v0"


class VerticalCREnv(_ScalarEnv):                    # vertical_cr_env.py:49-102
    ENV_ID = "VerticalCREnv-v0"


__all__ = [s.entry_point.split(":")[1] for s in SPECS.values()]
assert all(name in globals() for name in __all__), "a registered entry point has no class in this module"
