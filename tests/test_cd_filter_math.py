"""The mathematics behind K3's kept candidate list (csrc/env_kernels.cuh, group_cd), checked on the CPU against the oracle.

The kernel evaluates exactly only the pairs a conservative filter keeps, and keeps the list of the pairs among the
un-steered aircraft across the substeps of an env step while they hold their velocity.  Here the filter conditions and
their allowances are restated in float64 NumPy and run beside the oracle's StateBased.detect, substep by substep:
every conflict / LoS pair the oracle reports at substep k must be in
  (A) the per-substep filter of the pairs with aircraft 0, and
  (B) the list built at ANY earlier substep k0 of the same env step from the state at k0 with horizon (n_sub-1-k0) dt,
      for as long as no aircraft of the ring moved its velocity by more than the tolerance since k0.
(The GPU test test_in_group_cd_kept_candidates_are_a_superset checks the float32 kernel itself.)
"""
import numpy as np
import pytest

from oracle import envs
from oracle.philox import PhiloxDraws

RE = 6371000.0
R = 5.0 * 1852.0
DTLOOK = 300.0
TOL = 0.15          # kCdVelTol
EPS = 1.0           # kCdAbsEps


def _geometry(lat, lon, trk, gs):
    th = np.radians(trk)
    u, v = gs * np.sin(th), gs * np.cos(th)
    dl = lon - lon[0]
    dl = np.where(dl > 180.0, dl - 360.0, np.where(dl < -180.0, dl + 360.0, dl))
    x, y = RE * np.radians(dl), RE * np.radians(lat - lat[0])
    return x, y, u, v


def _pair_terms(lat, x, y, u, v, i, j):
    cav = np.cos(np.radians(0.5 * (lat[i] + lat[j])))
    dxl, dy = x[j] - x[i], y[j] - y[i]
    dx = dxl * cav
    du, dv = u[j] - u[i], v[j] - v[i]
    w = np.hypot(du, dv)
    return dxl, dx, dy, w, dx * dv - dy * du, du * dx + dv * dy


def filter_a(lat, lon, trk, gs):
    """pairs (0, j) kept by the per-substep filter (no allowance beyond the 2e-4 inflation)"""
    x, y, u, v = _geometry(lat, lon, trk, gs)
    keep = set()
    for j in range(1, len(lat)):
        _, dx, dy, w, crs, dot = _pair_terms(lat, x, y, u, v, 0, j)
        t0 = w * R * 1.0002 + EPS
        reach = R * 1.0002 + w * DTLOOK * 1.0002
        if abs(crs) < t0 and dot < t0 and dx * dx + dy * dy < reach * reach:
            keep.add((0, j))
    return keep


def filter_b(lat, lon, trk, gs, horizon):
    """pairs among aircraft 1.. kept by the (B) pass with the allowances for `horizon` seconds"""
    x, y, u, v = _geometry(lat, lon, trk, gs)
    n = len(lat)
    kap = d0 = dw = 0.0
    if horizon > 0.0:
        vmax = np.max(np.hypot(u, v)) + TOL
        tmax = np.max(np.tan(np.radians(np.minimum(np.abs(lat), 89.0))))
        kap = 1.5 * horizon * vmax * tmax / RE
        d0 = (4.0 * vmax + 1.0) * horizon
        dw = 2.0 * TOL
    lh = DTLOOK * 1.0002 + horizon * 1.0002
    keep = set()
    for i in range(1, n):
        for j in range(i + 1, n):
            dxl, dx, dy, w, crs, dot = _pair_terms(lat, x, y, u, v, i, j)
            l1 = abs(dxl) + abs(dy)
            rm = R * 1.0002 + kap * (l1 + d0)
            W = w + dw
            t0 = W * rm + (l1 + d0) * dw + EPS
            reach = rm + W * lh
            if abs(crs) < t0 and dot < t0 and dx * dx + dy * dy < reach * reach:
                keep.add((i, j))
    return keep


def _record_substeps(env):
    """wraps Traffic.update: per substep, the state the detection saw and the pairs it reported"""
    t = env.traf
    rec = []
    orig = t.update

    def update(fms_ready=True):
        pre = (t.lat.copy(), t.lon.copy(), t.trk.copy(), t.gs.copy())
        orig(fms_ready)
        rec.append((pre, set(t.confpairs) | set(t.lospairs)))
    t.update = update
    return rec


@pytest.mark.parametrize("make,act_dim,steps", [
    (lambda e: envs.HorizontalCREnv(n_intruders=20, cd_enabled=True, draws=PhiloxDraws(3, e, 0)), 1, 14),
    (lambda e: envs.HorizontalCREnv(n_intruders=20, cd_enabled=True, draws=PhiloxDraws(4, e, 0), init_alt=3000.0), 1, 10),
    (lambda e: envs.SectorCREnv(cd_enabled=True, draws=PhiloxDraws(5, e, 0)), 2, 16),
    (lambda e: envs.MergeEnv(cd_enabled=True, draws=PhiloxDraws(6, e, 0)), 2, 10),
])
def test_kept_list_is_a_superset_of_the_oracles_pairs(make, act_dim, steps):
    rng = np.random.default_rng(0)
    n_pairs = n_reused = 0
    for e in range(3):
        env = make(e)
        env.reset()
        rec = _record_substeps(env)
        for _ in range(steps):
            del rec[:]
            env.step(rng.uniform(-1, 1, act_dim))
            n_sub, dt = len(rec), env.SIMDT
            th = [np.radians(p[2]) for p, _ in rec]
            uv = [np.stack([p[3] * np.sin(a), p[3] * np.cos(a)]) for (p, _), a in zip(rec, th)]
            for k, (pre, found) in enumerate(rec):
                fa = filter_a(*pre)
                for (i, j) in found:
                    if min(i, j) == 0:
                        assert (0, max(i, j)) in fa, ("A", k, i, j)
            for k0 in range(n_sub):
                fb = filter_b(*rec[k0][0], horizon=(n_sub - 1 - k0) * dt)
                for k in range(k0, n_sub):
                    dev = np.abs(uv[k][0][1:] - uv[k0][0][1:]) + np.abs(uv[k][1][1:] - uv[k0][1][1:])
                    if np.any(dev > TOL):
                        break                       # the kernel would rebuild the list here
                    n_reused += k > k0
                    for (i, j) in rec[k][1]:
                        if min(i, j) >= 1:
                            n_pairs += 1
                            assert (min(i, j), max(i, j)) in fb, ("B", k0, k, i, j)
    assert n_pairs > 0 and n_reused > 0
