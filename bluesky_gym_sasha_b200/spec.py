"""Static description of the accelerated environments: ids, spaces, info keys, episode caps.

Everything here mirrors a declaration in the reference (cited per entry) so that observation dicts,
action shapes and info keys are drop-in identical; the numbers the kernels need (substeps, DT,
slot counts) come from ``bsg_query_layout`` on the C side, which cites the same lines.
"""
from collections import OrderedDict
from dataclasses import dataclass, field

import numpy as np

from . import _lib

INF = np.inf


@dataclass(frozen=True)
class EnvSpec:
    env_id: str
    env_type: int
    entry_point: str
    max_episode_steps: int                      # bluesky_gym/__init__.py:7-46
    act_dim: int
    obs_keys: tuple                             # (name, width or "n", low, high) in declaration order
    info_keys: tuple                            # first entries of the device info record
    default_kwargs: dict = field(default_factory=dict)

    def obs_layout(self, n_intruders=5):
        out, off = OrderedDict(), 0
        for name, width, lo, hi in self.obs_keys:
            w = n_intruders if width == "n" else width
            out[name] = (off, w, lo, hi)
            off += w
        return out, off


SPECS = OrderedDict()


def _add(spec):
    SPECS[spec.env_id] = spec


_add(EnvSpec(                                   # descent_env.py:53-62 (Box without shape -> (1,)), :119-126
    "DescentEnv-v0", _lib.ENV_DESCENT, "bluesky_gym_sasha_b200.envs:DescentEnv", 300, 1,
    (("altitude", 1, -INF, INF), ("vz", 1, -INF, INF), ("target_altitude", 1, -INF, INF),
     ("runway_distance", 1, -INF, INF)),
    ("total_reward", "final_altitude")))

_add(EnvSpec(                                   # horizontal_cr_env.py:49-62, :215-223
    "HorizontalCREnv-v0", _lib.ENV_HORIZONTAL_CR, "bluesky_gym_sasha_b200.envs:HorizontalCREnv", 300, 1,
    (("intruder_distance", "n", -INF, INF), ("cos_difference_pos", "n", -INF, INF),
     ("sin_difference_pos", "n", -INF, INF), ("x_difference_speed", "n", -INF, INF),
     ("y_difference_speed", "n", -INF, INF), ("waypoint_distance", 1, -INF, INF),
     ("cos_drift", 1, -INF, INF), ("sin_drift", 1, -INF, INF)),
    ("total_reward", "total_intrusions", "average_drift"),
    {"n_intruders": 5}))

_add(EnvSpec(                                   # sector_cr_env.py:52-67, :219-225
    "SectorCREnv-v0", _lib.ENV_SECTOR_CR, "bluesky_gym_sasha_b200.envs:SectorCREnv", 200, 2,
    (("cos(drift)", 1, -1, 1), ("sin(drift)", 1, -1, 1), ("airspeed", 1, -1, 1),
     ("x_r", 4, -INF, INF), ("y_r", 4, -INF, INF), ("vx_r", 4, -INF, INF), ("vy_r", 4, -INF, INF),
     ("cos(track)", 4, -INF, INF), ("sin(track)", 4, -INF, INF), ("distances", 4, -INF, INF)),
    ("total_reward", "total_intrusions", "average_drift")))

_add(EnvSpec(                                   # merge_env.py:59-76, :238-244
    "MergeEnv-v0", _lib.ENV_MERGE, "bluesky_gym_sasha_b200.envs:MergeEnv", 50, 2,
    (("cos(drift)", 1, -1, 1), ("sin(drift)", 1, -1, 1), ("airspeed", 1, -INF, INF),
     ("waypoint_dist", 1, -INF, INF), ("faf_reached", 1, 0, 1),
     ("x_r", 5, -INF, INF), ("y_r", 5, -INF, INF), ("vx_r", 5, -INF, INF), ("vy_r", 5, -INF, INF),
     ("cos(track)", 5, -INF, INF), ("sin(track)", 5, -INF, INF), ("distances", 5, -INF, INF)),
    ("total_reward", "faf_reach", "average_drift", "total_intrusions")))

_add(EnvSpec(                                   # plan_waypoint_env.py:49-58, :126-133
    "PlanWaypointEnv-v0", _lib.ENV_PLAN_WAYPOINT, "bluesky_gym_sasha_b200.envs:PlanWaypointEnv", 300, 1,
    (("waypoint_distance", 5, -INF, INF), ("cos_difference", 5, -INF, INF), ("sin_difference", 5, -INF, INF),
     ("waypoint_reached", 5, 0, 1)),
    ("total_reward", "waypoints_completed")))

_add(EnvSpec(                                   # vertical_cr_env.py:64-82, :202-211
    "VerticalCREnv-v0", _lib.ENV_VERTICAL_CR, "bluesky_gym_sasha_b200.envs:VerticalCREnv", 300, 1,
    (("altitude", 1, -INF, INF), ("vz", 1, -INF, INF), ("target_altitude", 1, -INF, INF),
     ("runway_distance", 1, -INF, INF), ("intruder_distance", 5, -INF, INF), ("cos_difference_pos", 5, -INF, INF),
     ("sin_difference_pos", 5, -INF, INF), ("altitude_difference", 5, -INF, INF),
     ("x_difference_speed", 5, -INF, INF), ("y_difference_speed", 5, -INF, INF), ("z_difference_speed", 5, -INF, INF)),
    ("total_reward", "total_intrusions", "final_altitude")))

_add(EnvSpec(                                   # static_obstacle_env.py:45-55, :285-292
    "StaticObstacleEnv-v0", _lib.ENV_STATIC_OBSTACLE, "bluesky_gym_sasha_b200.envs:StaticObstacleEnv", 100, 2,
    (("destination_waypoint_distance", 1, -INF, INF), ("destination_waypoint_cos_drift", 1, -INF, INF),
     ("destination_waypoint_sin_drift", 1, -INF, INF), ("restricted_area_radius", 10, 0, 1),
     ("restricted_area_distance", 10, -INF, INF), ("cos_difference_restricted_area_pos", 10, -INF, INF),
     ("sin_difference_restricted_area_pos", 10, -INF, INF)),
    ("total_reward", "waypoint_reached", "crashed", "average_drift")))

# every id the reference registers (bluesky_gym/__init__.py:7-46) is on the accelerated path
NOT_ACCELERATED = ()

# oracle/perf.py::A320 (kept in sync by tests/test_host_logic.py)
A320_PERF = dict(vminto=73.3, vmaxic=88.5, vminer=64.0, vmaxer=163.0, vminap=64.0, vmaxap=78.0,
                 vsmin=-20.4, vsmax=18.6, hmax=12500.0, mmo=0.82, axmax_gd=2.0, axmax_air=0.5)

AUTORESET = {"disabled": _lib.AUTORESET_DISABLED, "next_step": _lib.AUTORESET_NEXT_STEP,
             "same_step": _lib.AUTORESET_SAME_STEP}
