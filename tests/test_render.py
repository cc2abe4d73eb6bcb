"""rgb_array frames: the draw calls are pinned to the REFERENCE's own ``_render_frame`` functions
(tests/golden/ref_render.npz, recorded by make_render_golden.py from the reference's env files over a recording pygame
stand-in), the painter ``bsg_render`` to the coverage rules restated in oracle/raster.py."""
import os

import numpy as np
import pytest

from bluesky_gym_sasha_b200 import _lib, render
from oracle import raster

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_render.npz"))
ENV_IDS = sorted(render.FRAMES)


def golden_frames(env_id):
    for k in range(int(GOLD[env_id + "/frames"])):
        p = f"{env_id}/{k}/"
        s = {key[len(p):]: GOLD[key] for key in GOLD.files if key.startswith(p)}
        for key in ("n", "wpt_reach", "nvert"):
            s[key] = int(s[key])
        for key in ("wpt_lat", "wpt_lon", "target_alt", "fix_lat", "fix_lon"):
            if key in s:
                s[key] = float(s[key])
        yield k, s


def canvas_rows(cv):
    """Canvas records -> the golden file's (kind, x0, y0, x1, y1, w, r, g, b) rows"""
    a = cv.array()
    col = a[:, 6].copy().view(np.uint32)
    return np.concatenate([a[:, :6].astype(np.float64), np.stack([col & 255, (col >> 8) & 255, (col >> 16) & 255], 1)], 1)


@pytest.mark.parametrize("env_id", ENV_IDS)
def test_draw_calls_equal_the_reference(env_id):
    """Same state in, same draw calls out: kind, order, colours exactly; coordinates / radii / widths to float32 rounding
    (the records are float32; pixel coordinates are < 1024, so 1e-4 px is 2 ulp)."""
    n = 0
    for k, s in golden_frames(env_id):
        cv = render.FRAMES[env_id](s)
        assert (cv.width, cv.height) == tuple(s["size"])
        mine, ref = canvas_rows(cv), s["prims"]
        assert mine.shape == ref.shape, (env_id, k)
        assert np.array_equal(mine[:, 0], ref[:, 0]) and np.array_equal(mine[:, 6:], ref[:, 6:])
        assert np.abs(mine[:, 1:6] - ref[:, 1:6]).max() < 1e-4
        assert cv.background == (135, 206, 235)
        n += 1
    assert n >= 3


def test_raster_rules_on_closed_forms():
    """oracle/raster.py against pixel counts known in closed form"""
    img, _ = raster.paint([(raster.RECT, 2, 3, 10, 7, 0, 9, 8, 7)], 16, 16, (0, 0, 0))
    assert (img[..., 0] == 9).sum() == 8 * 4 and img[3, 2, 0] == 9 and img[7, 2, 0] == 0 and img[3, 10, 0] == 0
    img, _ = raster.paint([(raster.RING, 100, 100, 50, 0, 0, 255, 0, 0)], 200, 200, (0, 0, 0))
    assert abs((img[..., 0] == 255).sum() - np.pi * 2500) < 60                     # disc area
    img, _ = raster.paint([(raster.RING, 100, 100, 50, 0, 2, 255, 0, 0)], 200, 200, (0, 0, 0))
    assert abs((img[..., 0] == 255).sum() - np.pi * (2500 - 48 * 48)) < 60         # ring drawn inwards
    img, _ = raster.paint([(raster.LINE, 10, 50, 90, 50, 4, 1, 2, 3)], 100, 100, (0, 0, 0))
    assert (img[:, 50, 0] == 1).sum() == 4                                          # a width-4 line is 4 pixels thick
    tri = [(raster.EDGE, 10, 10, 90, 10, 0, 5, 5, 5), (raster.EDGE, 90, 10, 10, 90, 0, 5, 5, 5),
           (raster.EDGE_END, 10, 90, 10, 10, 0, 5, 5, 5)]
    img, _ = raster.paint(tri, 100, 100, (0, 0, 0))
    assert abs((img[..., 0] == 5).sum() - 3200) < 100 and img[20, 20, 0] == 5 and img[80, 80, 0] == 0
    img, _ = raster.paint([(raster.RECT, 0, 0, 50, 50, 0, 1, 1, 1), (raster.RECT, 25, 25, 50, 50, 0, 2, 2, 2)], 64, 64, (0, 0, 0))
    assert img[30, 30, 0] == 2 and img[10, 10, 0] == 1                              # painter's order


def _paint_device(rows, width, height, cuda):
    cv = render.Canvas(width, height)
    for r in rows:
        cv._add(r[0], r[1], r[2], r[3], r[4], r[5], (int(r[6]), int(r[7]), int(r[8])))
    return render.render_canvas(cv, cuda)


@pytest.mark.gpu
@pytest.mark.parametrize("env_id", ENV_IDS)
def test_bsg_render_paints_the_reference_draw_calls(cuda, env_id):
    """The reference's recorded draw calls through bsg_render == oracle/raster.py, pixel for pixel outside the 1e-3 px
    band around primitive edges (float32 on the device, float64 in the oracle); the band is a sliver of the frame and
    the pixels that actually differ are fewer still."""
    for k, s in golden_frames(env_id):
        w, h = (int(v) for v in s["size"])
        got = _paint_device(s["prims"], w, h, cuda)
        want, unsure = raster.paint(s["prims"], w, h, (135, 206, 235))
        assert got.shape == (h, w, 3) and got.dtype == np.uint8
        diff = np.any(got != want, axis=2)
        assert not np.any(diff & ~unsure), (env_id, k, int((diff & ~unsure).sum()))
        # (DescentEnv's integer-valued lines put whole pixel rows exactly on an edge: the band is up to 1 % there)
        assert unsure.mean() < 2e-2 and diff.mean() < 1e-3
        assert np.any(np.any(want != np.array((135, 206, 235), dtype=np.uint8), axis=2))       # something was drawn


@pytest.mark.gpu
def test_bsg_render_edge_cases(cuda):
    lib = _lib.load()
    frame = _paint_device(np.zeros((0, 9)), 33, 17, cuda)                            # no primitives, ragged size
    assert frame.shape == (17, 33, 3) and np.all(frame == np.array((135, 206, 235), dtype=np.uint8))
    rng = np.random.default_rng(0)
    rows = []
    for _ in range(1000):                                                           # close to the 1024 capacity
        kind = int(rng.integers(1, 4))
        x0, y0, x1, y1 = rng.uniform(-20, 220, 4)
        if kind == raster.RING:
            x1 = rng.uniform(1, 40)
        if kind == raster.RECT:
            x1, y1 = x0 + rng.uniform(1, 30), y0 + rng.uniform(1, 30)
        rows.append((kind, x0, y0, x1, y1, float(rng.integers(0, 6)), *rng.integers(0, 256, 3)))
    rows = np.array(rows, dtype=np.float64)
    got = _paint_device(rows, 200, 120, cuda)
    want, unsure = raster.paint(rows.astype(np.float32), 200, 120, (135, 206, 235))
    assert not np.any(np.any(got != want, axis=2) & ~unsure)
    import ctypes as C
    assert lib.bsg_render(None, 2000, 16, 16, 0, None, None) == _lib.BSG_EINVAL       # over capacity: refused, not clipped
    assert lib.bsg_render(None, 0, 0, 16, 0, None, None) == _lib.BSG_EINVAL


@pytest.mark.gpu
@pytest.mark.parametrize("env_id", ENV_IDS)
def test_rgb_array_through_gym_make(cuda, env_id):
    """``gym.make(id, render_mode="rgb_array")`` -> ``render()`` returns the frame of the env's current state: equal to the
    frame function applied to a host snapshot and painted by the oracle, and it changes as the env steps."""
    import bluesky_gym
    bluesky_gym.register_envs()
    env = bluesky_gym.make(env_id, render_mode="rgb_array")
    env.reset(seed=1)
    u = env.unwrapped
    f0 = env.render()
    cv = render.FRAMES[env_id](render.snapshot(u.vec, 0))
    assert f0.shape == (cv.height, cv.width, 3) and f0.dtype == np.uint8
    want, unsure = raster.paint(canvas_rows(cv), cv.width, cv.height, cv.background)
    assert not np.any(np.any(f0 != want, axis=2) & ~unsure)
    for _ in range(5):
        env.step(np.ones(env.action_space.shape))
    f1 = env.render()
    assert np.any(f0 != f1)
    env.close()
    with pytest.raises(NotImplementedError):
        bluesky_gym.make(env_id, render_mode="human")
    assert bluesky_gym.make(env_id).render() is None
