"""A stand-in ``bluesky`` module backed by oracle.traffic, so that the reference's OWN environment files
(/root/reference/bluesky_gym/envs/*.py, unmodified, imported from where they lie) can be executed in this
container even though ``bluesky-simulator`` is not installable here.

What it pins: the in-tree half of the path (reset draw order, scenario generators, ``_get_action``,
``_get_obs``, ``_get_reward``, termination / truncation, info dicts, the ``common/functions.py`` helpers) is
the reference's code itself, so golden vectors produced through this shim (tests/golden/make_golden.py)
check oracle/envs.py -- and through it the CUDA kernels -- against the reference's env logic rather than
against a transcription.  What it does NOT pin: everything below ``bs.*`` is still the restated upstream
BlueSky core of oracle/traffic.py ([UPSTREAM-RECALL]); parity at that boundary stays unpinned.

Only the surface the reference touches is provided (SURVEY.md section 8c lists every ``bs.*`` call site):
``bs.init``, ``bs.scr``, ``bs.stack.stack`` (DT, FF, HDG, SPD, ADDWPT, DEST, RESO), ``bs.sim.step``, ``bs.traf``
(cre, creconfs, reset, delete, id2idx, id, state arrays, ``ap.trk``, ``ap.selaltcmd``), ``bs.tools.geo``,
``bs.tools.areafilter``, ``bluesky.tools.aero.kts``, ``bluesky.simulation.ScreenIO``; plus inert ``pygame`` and
``stable_baselines3`` stubs and -- when gymnasium is not installed -- the package's own ``gym_compat`` surface
under the name ``gymnasium``.  Test infrastructure only (see oracle/__init__.py).
"""
import sys
import types

import numpy as np

from . import aero, geo
from .traffic import Traffic

REFERENCE_ROOT = "/root/reference"


class _Autopilot:
    """``bs.traf.ap``: views on the Traffic arrays (static_obstacle_env.py:123, vertical_cr_env.py:200)."""

    def __init__(self, traf):
        self._t = traf

    @property
    def trk(self):
        return self._t.ap_trk

    def selaltcmd(self, idx, alt, vspd=None):
        return self._t.selaltcmd(idx, alt, vspd)


class ShimTraffic(Traffic):
    def __init__(self, **kw):
        super().__init__(**kw)
        self.ap = _Autopilot(self)


class _Areas:
    """``bs.tools.areafilter``: named POLY shapes; even-odd point-in-polygon like matplotlib's Path
    (sector_cr_env.py:89,136,160,203; static_obstacle_env.py:171,185,212,334)."""

    def __init__(self):
        self.shapes = {}

    def defineArea(self, name, kind, coordinates, top=1e9, bottom=-1e9):
        assert kind == "POLY"
        c = np.asarray(coordinates, dtype=np.float64)
        self.shapes[name] = (c[0::2].copy(), c[1::2].copy(), top, bottom)
        return True

    def deleteArea(self, name):
        return self.shapes.pop(name, None) is not None

    def hasArea(self, name):
        return name in self.shapes

    def checkInside(self, name, lat, lon, alt):
        lat, lon, alt = (np.atleast_1d(np.asarray(v, dtype=np.float64)) for v in (lat, lon, alt))
        if name not in self.shapes:
            return np.zeros(lat.shape, dtype=bool)
        vlat, vlon, top, bottom = self.shapes[name]
        inside = np.array([bool(geo.point_in_polygon(a, o, vlat, vlon)) for a, o in zip(lat, lon)])
        return inside & (alt >= bottom) & (alt <= top)


class _Stack:
    """``bs.stack.stack(cmdline)``.  DT takes effect at once (it is issued once, in the env constructor,
    before any aircraft exists); the others are queued on the traffic object and run at the start of the
    next ``sim.step()``, which is when upstream's ``stack.process()`` runs them."""

    def __init__(self, bs):
        self.bs = bs

    def stack(self, cmdline):
        for cmd in str(cmdline).split(";"):
            tok = cmd.replace(",", " ").split()
            if not tok:
                continue
            t = self.bs.traf
            head = tok[0].upper()
            if head == "DT":
                t.simdt = float(tok[1])
                t.fms_rel_freq = max(1, int(10.5 // t.simdt))
            elif head in ("FF", "RESO", "OP", "HOLD"):
                pass                                   # fast-forward / resolution off: no effect on a detached sim
            elif head == "HDG":
                t.stack_hdg(tok[1], float(tok[2]))
            elif head == "SPD":
                t.stack_spd(tok[1], float(tok[2]))
            elif len(tok) >= 4 and tok[1].upper() == "ADDWPT":
                t.stack_addwpt(tok[0], float(tok[2]), float(tok[3]))
            elif len(tok) >= 4 and tok[1].upper() == "DEST":
                t.stack_dest(tok[0], float(tok[2]), float(tok[3]))
            else:
                raise NotImplementedError(f"bs_shim: stack command not on the reference's path: {cmd!r}")


class _Sim:
    def __init__(self, bs):
        self.bs = bs

    def step(self):
        self.bs.traf.simstep()

    def reset(self):
        self.bs.traf.reset()


def _kwikdist_matrix(lata, lona, latb, lonb):
    """[UPSTREAM-RECALL] geo.kwikdist_matrix broadcasts ``b - a.T``: with a scalar ownship and a 1-D vector of
    others (the reference's only call, merge_env.py:195) the result is a 1-D vector of NM distances."""
    lata, lona, latb, lonb = (np.asarray(v, dtype=np.float64) for v in (lata, lona, latb, lonb))
    dlat = np.radians(latb - lata.T)
    dlon = np.radians(((lonb - lona.T) + 180.0) % 360.0 - 180.0)
    cavelat = np.cos(np.radians(lata + latb.T) * 0.5)
    dangle = np.sqrt(dlat * dlat + dlon * dlon * cavelat * cavelat)
    return geo.RE_KWIK * dangle / aero.nm


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install(cd_enabled=False, default_hdg="random"):
    """Put the stand-in modules into ``sys.modules`` and make the reference package importable.
    Returns the fake ``bluesky`` module.  Call before importing ``bluesky_gym`` (the reference's)."""
    bs = _module("bluesky")
    bs.settings = types.SimpleNamespace()
    bs.scr = None

    def init(mode="sim", detached=True, **kw):
        # process-global singleton like upstream: a second env in the same process re-initialises it
        bs.traf = ShimTraffic(simdt=1.0, cd_enabled=cd_enabled, default_hdg=default_hdg)
        bs.sim = _Sim(bs)
        bs.stack = _Stack(bs)
        areas = _Areas()
        bs.tools.areafilter = areas
        sys.modules["bluesky.tools.areafilter"] = areas

    bs.init = init
    geo_mod = _module("bluesky.tools.geo", kwikqdrdist=geo.kwikqdrdist, kwikdist=geo.kwikdist,
                      kwikdist_matrix=_kwikdist_matrix,
                      kwikqdrdist_matrix=geo.kwikqdrdist_matrix, kwikpos=geo.kwikpos, qdrdist=geo.qdrdist)
    aero_mod = _module("bluesky.tools.aero", kts=aero.kts, ft=aero.ft, nm=aero.nm, fpm=aero.fpm)
    bs.tools = _module("bluesky.tools", geo=geo_mod, aero=aero_mod)
    _module("bluesky.simulation", ScreenIO=type("ScreenIO", (), {}))
    _module("bluesky.stack")
    # inert stand-ins for packages that are only used off the path (rendering, SB3 logger callback)
    if "pygame" not in sys.modules:
        _module("pygame")
    try:
        import stable_baselines3  # noqa: F401
    except Exception:
        _module("stable_baselines3")
        _module("stable_baselines3.common")
        _module("stable_baselines3.common.callbacks", BaseCallback=type("BaseCallback", (), {
            "__init__": lambda self, verbose=0: None}))
    try:
        import gymnasium  # noqa: F401
    except Exception:
        from bluesky_gym_sasha_b200 import gym_compat as gc     # spaces / Env / register surface only
        reg = _module("gymnasium.envs.registration", register=gc.register)
        envs = _module("gymnasium.envs", registration=reg)
        sp = _module("gymnasium.spaces", Box=gc.spaces.Box, Dict=gc.spaces.Dict)
        _module("gymnasium", Env=gc.Env, spaces=sp, envs=envs, Wrapper=getattr(gc, "Wrapper", object),
                ObservationWrapper=getattr(gc, "ObservationWrapper", object), make=gc.make, register=gc.register)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    return bs
