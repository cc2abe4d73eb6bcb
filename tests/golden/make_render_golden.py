#!/usr/bin/env python
"""Generates tests/golden/ref_render.npz: the draw calls of the REFERENCE's own ``_render_frame`` functions.

    python tests/golden/make_render_golden.py        # needs /root/reference (build container only)

pygame is not in the image, so a recording stand-in is put into ``sys.modules`` before the reference's env files
(``bluesky_gym/envs/*_env.py``, unmodified, over ``oracle/bs_shim.py`` like make_golden.py) are imported.  Every
``pygame.draw.line / circle / rect / polygon`` call the reference makes while ``render_mode="human"`` is recorded as
rows ``(kind, x0, y0, x1, y1, width, r, g, b)`` (a polygon = one row per edge).  Beside the rows of every recorded
frame the file holds the env's state at that moment in the layout ``bluesky_gym_sasha_b200/render.py::snapshot`` reads
from the device, so that ``tests/test_render.py`` can run ``render.frame_*`` on the same state and compare draw call by
draw call (CPU test), and paint both lists through ``bsg_render`` (GPU test).
"""
import importlib
import os
import random
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

LINE, RING, RECT, EDGE, EDGE_END = 1, 2, 3, 4, 5
NAMED = {"red": (255, 0, 0), "black": (0, 0, 0), "white": (255, 255, 255)}
CALLS = []


def _rgb(c):
    c = NAMED[c] if isinstance(c, str) else c
    return float(c[0]), float(c[1]), float(c[2])


def install_recording_pygame():
    pg = types.ModuleType("pygame")
    pg.init = lambda: None
    pg.display = types.SimpleNamespace(init=lambda: None, update=lambda: None,
                                       set_mode=lambda size: types.SimpleNamespace(blit=lambda *a, **k: None))
    pg.time = types.SimpleNamespace(Clock=lambda: types.SimpleNamespace(tick=lambda fps: None))
    pg.event = types.SimpleNamespace(pump=lambda: None, get=lambda: [])
    pg.Rect = lambda *a: tuple(a) if len(a) == 4 else tuple(a[0]) + tuple(a[1])      # Rect(l, t, w, h) or Rect((l, t), (w, h))

    class Surface:
        def __init__(self, size):
            self.size = size
            CALLS.clear()                                     # a new canvas = a new frame

        def fill(self, color):
            self.bg = color

        def get_rect(self):
            return (0, 0) + tuple(self.size)

    def line(canvas, color, p0, p1, width=1):
        CALLS.append((LINE, p0[0], p0[1], p1[0], p1[1], width) + _rgb(color))

    def circle(canvas, color, center, radius, width=0):
        CALLS.append((RING, center[0], center[1], radius, 0.0, width) + _rgb(color))

    def rect(canvas, color, r, width=0):
        CALLS.append((RECT, r[0], r[1], r[0] + r[2], r[1] + r[3], 0.0) + _rgb(color))

    def polygon(canvas, color, points, width=0):
        n = len(points)
        for k in range(n):
            a, b = points[k], points[(k + 1) % n]
            kind = LINE if width > 0 else (EDGE_END if k == n - 1 else EDGE)
            CALLS.append((kind, a[0], a[1], b[0], b[1], width) + _rgb(color))

    pg.Surface = Surface
    pg.draw = types.SimpleNamespace(line=line, circle=circle, rect=rect, polygon=polygon)
    sys.modules["pygame"] = pg
    return pg


SPEC = [("DescentEnv-v0", "descent_env", "DescentEnv", 1), ("PlanWaypointEnv-v0", "plan_waypoint_env", "PlanWaypointEnv", 1),
        ("HorizontalCREnv-v0", "horizontal_cr_env", "HorizontalCREnv", 1), ("VerticalCREnv-v0", "vertical_cr_env", "VerticalCREnv", 1),
        ("SectorCREnv-v0", "sector_cr_env", "SectorCREnv", 2), ("StaticObstacleEnv-v0", "static_obstacle_env", "StaticObstacleEnv", 2),
        ("MergeEnv-v0", "merge_env", "MergeEnv", 2)]
MAX_AC, FRAMES_PER_ENV, STRIDE = 32, 4, 7


def state_of(bs, env, env_id, fn):
    """The env's state in the layout render.snapshot() produces from the device tensors."""
    n = bs.traf.ntraf
    s = {"n": n}
    for f in ("lat", "lon", "alt", "hdg"):
        a = np.full(MAX_AC, np.nan)
        a[:n] = getattr(bs.traf, f)[:n]
        s[f] = a
    s["wpt_lat"] = s["wpt_lon"] = s["target_alt"] = 0.0
    s["wpts"], s["wpt_reach"], s["nvert"], s["poly"] = np.zeros(10), 0, 0, np.zeros(360)
    if env_id in ("HorizontalCREnv-v0", "StaticObstacleEnv-v0"):
        s["wpt_lat"], s["wpt_lon"], s["wpt_reach"] = env.wpt_lat[0], env.wpt_lon[0], int(env.wpt_reach[0])
    if env_id == "PlanWaypointEnv-v0":
        s["wpts"] = np.stack([env.wpt_lat, env.wpt_lon], 1).reshape(-1)
        s["wpt_reach"] = sum(int(r) << k for k, r in enumerate(env.wpt_reach))
    if env_id in ("DescentEnv-v0", "VerticalCREnv-v0"):
        s["target_alt"] = float(env.target_alt)
    if env_id == "SectorCREnv-v0":
        c = np.array([51.990426702297746, 4.376124857109851])
        ll = np.array([fn.nm_to_latlong(c, p) for p in env.poly_points])
        s["nvert"] = len(ll)
        s["poly"][:2 * len(ll)] = ll.reshape(-1)
    if env_id == "StaticObstacleEnv-v0":
        for k, verts in enumerate(env.obstacle_vertices):
            v = np.asarray(verts, dtype=np.float64)
            s["poly"][32 * k:32 * k + 2 * len(v)] = v.reshape(-1)
            s["poly"][350 + k] = len(v)
    if env_id == "MergeEnv-v0":
        s["fix_lat"], s["fix_lon"] = env.wpt_lat, env.wpt_lon
    return s


def main():
    if not os.path.isdir("/root/reference/bluesky_gym"):
        raise SystemExit("make_render_golden.py needs the reference at /root/reference (build container only)")
    install_recording_pygame()
    from oracle import bs_shim
    bs = bs_shim.install()
    fn = importlib.import_module("bluesky_gym.envs.common.functions")
    out = {}
    for env_id, mod, cls, adim in SPEC:
        m = importlib.import_module("bluesky_gym.envs." + mod)
        np.random.seed(3)
        random.seed(3)
        env = getattr(m, cls)(render_mode="human")
        rng = np.random.default_rng(5)
        env.reset()
        k = 0
        for step in range(FRAMES_PER_ENV * STRIDE):
            _, _, term, trunc, _ = env.step(rng.uniform(-1, 1, adim))
            if term or trunc:                                  # (Descent / VerticalCR delete their aircraft: frame undefined)
                env.reset()
                continue
            if step % STRIDE == STRIDE - 1:
                env._render_frame()                            # the frame of the CURRENT state (step() drew before the obs update for some envs)
                p = f"{env_id}/{k}/"
                out[p + "prims"] = np.array(CALLS, dtype=np.float64).reshape(-1, 9)
                out[p + "size"] = np.array(env.window_size)
                for key, v in state_of(bs, env, env_id, fn).items():
                    out[p + key] = np.asarray(v)
                k += 1
        out[env_id + "/frames"] = np.array(k)
        print(f"{env_id}: {k} frames, last one {len(CALLS)} draw calls")
    np.savez_compressed(os.path.join(HERE, "ref_render.npz"), **out)


if __name__ == "__main__":
    main()
