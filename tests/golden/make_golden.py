#!/usr/bin/env python
"""Generates tests/golden/*.npz by importing the REFERENCE in this container (it cannot travel to the GPU box).

    python tests/golden/make_golden.py            # needs /root/reference; rewrites the fixtures
    python tests/golden/make_golden.py --real     # same, on a machine where bluesky-simulator itself is installed

Two kinds of fixture:

* ``ref_functions.npz`` -- ``bluesky_gym/envs/common/functions.py`` of the reference imported as is (it only
  needs numpy) and evaluated on fixed inputs.  A true pin of those helpers.
* ``ref_<EnvId>.npz`` -- the reference's own env classes (``bluesky_gym/envs/*_env.py``, unmodified) executed
  with ``oracle/bs_shim.py`` standing in for the absent ``bluesky`` package.  Reset draw order, scenario
  generators, action mapping, observations, rewards, termination / truncation and info are therefore the
  reference's code; the simulator core under ``bs.*`` is the oracle's restatement of upstream BlueSky (still
  unpinned -- see oracle/__init__.py).  Per seed the file holds a flat sequence of rows: a row is either the
  output of ``reset()`` (``is_reset`` = 1, action NaN) or of ``step(action)``; episodes run to termination /
  truncation (TimeLimit cap of the registration applied here, bluesky_gym/__init__.py:9-45) and are followed by
  a fresh ``reset()`` WITHOUT reseeding, until ``ROWS`` rows exist.  ``ref_wind_<EnvId>.npz``: the same envs inside
  the reference's own ``WindFieldWrapper`` (wrappers/wind.py, ``augment_obs=True``, the README's four-source field).  ``traf_*`` is the aircraft state right
  after the row's call (NaN padded; after a terminal step the reference has already deleted aircraft).
"""
import importlib
import os
import random
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")           # "Mean of empty slice" from the reference's average_drift at reset

SPEC = [  # env id, module, class, action dim, TimeLimit cap, rows per seed
    ("DescentEnv-v0", "descent_env", "DescentEnv", 1, 300, 100),
    ("PlanWaypointEnv-v0", "plan_waypoint_env", "PlanWaypointEnv", 1, 300, 120),
    ("HorizontalCREnv-v0", "horizontal_cr_env", "HorizontalCREnv", 1, 300, 120),
    ("VerticalCREnv-v0", "vertical_cr_env", "VerticalCREnv", 1, 300, 100),
    ("SectorCREnv-v0", "sector_cr_env", "SectorCREnv", 2, 200, 120),
    ("StaticObstacleEnv-v0", "static_obstacle_env", "StaticObstacleEnv", 2, 100, 120),
    ("MergeEnv-v0", "merge_env", "MergeEnv", 2, 50, 120),
]
SEEDS = (0, 1, 2)
TRAF_FIELDS = ("lat", "lon", "alt", "hdg", "tas", "vs", "gs", "trk")
MAX_AC = 32


def action_bank(seed, rows, adim, env_id):
    """Actions come from a private Generator so the global np.random / random streams only see the
    reference's own draws.  A third of every bank is a smooth steering pattern so that goal-reaching
    branches (waypoint reached, runway reached) occur, not only random-walk episodes."""
    rng = np.random.default_rng(10_000 + seed)
    a = rng.uniform(-1.0, 1.0, (rows, adim))
    if seed == 2:
        a[:, 0] *= 0.05                    # nearly straight flight: reaches waypoints / runway
    return a


# wrappers/README.md example: a static wind field from four sources, no altitude gradient
WIND = dict(lat=np.array([51.9, 51.9, 52.1, 52.1]), lon=np.array([3.9, 4.1, 3.9, 4.1]),
            vnorth=np.array([[16.0, 12.0, 14.0, 15.0]]), veast=np.array([[3.0, 7.0, 9.0, 4.0]]))
# augment_obs: the reference wrapper reads bs.traf.lat[id2idx('kl001')] AFTER the env's step; DescentEnv / VerticalCREnv
# delete their aircraft on the terminal step, so with augment_obs=True the reference raises IndexError there --
# those two are recorded with augment_obs=False (wind in the dynamics only)
WIND_ENVS = {"DescentEnv-v0": False, "VerticalCREnv-v0": False, "SectorCREnv-v0": True, "MergeEnv-v0": True}


def gen_env(bs, env_id, mod, cls, adim, cap, rows, wind=False):
    m = importlib.import_module("bluesky_gym.envs." + mod)
    out = {"seeds": np.array(SEEDS), "cap": np.array(cap)}
    if wind:
        wmod = importlib.import_module("bluesky_gym.wrappers.wind")     # the reference's WindFieldWrapper, as is
        for k, v in WIND.items():
            out["wind_" + k] = v
        out["augment_obs"] = np.array(WIND_ENVS[env_id])
    for seed in SEEDS:
        np.random.seed(seed)
        random.seed(seed)
        env = getattr(m, cls)()
        if wind:
            env = wmod.WindFieldWrapper(env, augment_obs=WIND_ENVS[env_id], **WIND)
        acts = action_bank(seed, rows, adim, env_id)
        rec = {"is_reset": [], "action": [], "reward": [], "terminated": [], "truncated": []}
        obs_rec, info_rec, traf_rec = {}, {}, {f: [] for f in TRAF_FIELDS}
        ntraf = []

        def push(is_reset, action, obs, reward, term, trunc, info):
            rec["is_reset"].append(is_reset)
            rec["action"].append(action)
            rec["reward"].append(float(reward))
            rec["terminated"].append(bool(term))
            rec["truncated"].append(bool(trunc))
            for k, v in obs.items():
                obs_rec.setdefault(k, []).append(np.asarray(v, dtype=np.float64).reshape(-1))
            for k, v in info.items():
                info_rec.setdefault(k, []).append(float(v))
            n = bs.traf.ntraf
            ntraf.append(n)
            for f in TRAF_FIELDS:
                row = np.full(MAX_AC, np.nan)
                row[:n] = getattr(bs.traf, f)[:n]
                traf_rec[f].append(row)

        t, need_reset = 0, True
        for r in range(rows):
            if need_reset:
                obs, info = env.reset()
                push(1, np.full(adim, np.nan), obs, 0.0, False, False, info)
                t, need_reset = 0, False
                continue
            obs, reward, term, trunc, info = env.step(acts[r].copy())
            t += 1
            trunc = bool(trunc) or t >= cap
            push(0, acts[r], obs, reward, term, trunc, info)
            need_reset = bool(term) or trunc
        p = f"s{seed}_"
        for k, v in rec.items():
            out[p + k] = np.array(v)
        out[p + "ntraf"] = np.array(ntraf)
        for k, v in obs_rec.items():
            out[p + "obs_" + k] = np.array(v)
        for k, v in info_rec.items():
            out[p + "info_" + k] = np.array(v)
        for k, v in traf_rec.items():
            out[p + "traf_" + k] = np.array(v)
        n_ep = int(np.sum(rec["is_reset"]))
        print(f"{env_id} seed {seed}: {rows} rows, {n_ep} episodes, {int(np.sum(rec['terminated']))} terminated, "
              f"{int(np.sum(rec['truncated']))} truncated, sum reward {np.sum(rec['reward']):.4f}")
    np.savez_compressed(os.path.join(HERE, f"ref_{'wind_' if wind else ''}{env_id}.npz"), **out)


def gen_functions():
    """bluesky_gym/envs/common/functions.py imported as is."""
    fn = importlib.import_module("bluesky_gym.envs.common.functions")
    rng = np.random.default_rng(7)
    ang = np.concatenate([rng.uniform(-720, 720, 200), [-540, -360, -180, 0, 180, 360, 540, 179.999, -180.001]])
    out = {"wrap_in": ang, "wrap_out": np.array([fn.bound_angle_positive_negative_180(a) for a in ang])}
    lat, lon = rng.uniform(-70, 70, 100), rng.uniform(-180, 180, 100)
    d, brg = rng.uniform(0, 500, 100), rng.uniform(-360, 720, 100)
    out["gpad_in"] = np.stack([lat, lon, d, brg], 1)
    out["gpad_out"] = np.array([fn.get_point_at_distance(*r) for r in out["gpad_in"]])
    c = np.array([51.990426702297746, 4.376124857109851])
    pts = rng.uniform(-60, 60, (100, 2))
    out["center"] = c
    out["nm_in"] = pts
    out["nm2ll_out"] = np.array([fn.nm_to_latlong(c, p) for p in pts])
    out["ll2nm_out"] = np.array([fn.latlong_to_nm(c, q) for q in out["nm2ll_out"]])
    p2 = out["nm2ll_out"][::-1].copy()
    out["hdg_in_a"], out["hdg_in_b"] = out["nm2ll_out"], p2
    out["hdg_out"] = np.array([fn.get_hdg(a, b) for a, b in zip(out["nm2ll_out"], p2)])
    polys, areas, sorted_polys = [], [], []
    for k in range(20):
        n = int(rng.integers(3, 12))
        v = rng.uniform(-40, 40, (n, 2))
        s = np.array(fn.sort_points_clockwise(v))
        polys.append(np.pad(v, ((0, 12 - n), (0, 0)), constant_values=np.nan))
        sorted_polys.append(np.pad(s, ((0, 12 - n), (0, 0)), constant_values=np.nan))
        areas.append(fn.polygon_area(s))
    out["poly_in"], out["poly_sorted"], out["poly_area"] = np.array(polys), np.array(sorted_polys), np.array(areas)
    out["eucl_out"] = np.array([fn.euclidean_distance(a, b) for a, b in zip(pts, pts[::-1])])
    np.savez_compressed(os.path.join(HERE, "ref_functions.npz"), **out)
    print("ref_functions.npz written")


def main():
    if not os.path.isdir("/root/reference/bluesky_gym"):
        raise SystemExit("make_golden.py needs the reference at /root/reference (build container only)")
    if "--real" in sys.argv:
        # With a real ``bluesky-simulator`` (+ gymnasium, pygame) installed the same script records the vectors of the
        # UNMODIFIED reference stack: that turns the partial pin into a full one and settles every [UPSTREAM-RECALL]
        # item of oracle/ (tests/test_golden.py then checks the oracle and the CUDA path against the real thing).
        import bluesky as bs                        # noqa: F401  (raises here if it is not installed)
        sys.path.insert(0, "/root/reference")
        print("recording with the REAL bluesky package:", bs.__file__)
    else:
        from oracle import bs_shim
        bs = bs_shim.install()
    gen_functions()
    for spec in SPEC:
        gen_env(bs, *spec)
    for spec in SPEC:
        if spec[0] in WIND_ENVS:
            gen_env(bs, *spec[:5], 80, wind=True)


if __name__ == "__main__":
    main()
