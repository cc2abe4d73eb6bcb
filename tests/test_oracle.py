"""CPU tests of the oracle itself: known-answer vectors, closed-form cases, self-consistency round trips,
and the distribution-level pins derived from the reference's shipped training logs.
The reference ships no golden vectors for this path ("parity unpinned"): these are the pins that exist."""
import random

import numpy as np
import pytest

from oracle import aero, cbind, envs, geo, perf, philox, statebased
from oracle.traffic import Traffic


# ---- Philox4x32-10: Random123 kat_vectors (the one true golden vector set on this path) -------------
@pytest.mark.parametrize("ctr,key,out", [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
])
def test_philox_kat(ctr, key, out):
    assert philox.philox4x32_10(ctr, key) == out


def test_philox_draw_ranges():
    d = philox.PhiloxDraws(1, 2, 3)
    xs = [d.randint(45, 315) for _ in range(2000)]
    assert min(xs) >= 45 and max(xs) < 315 and len(set(xs)) > 200
    us = [d.uniform(-15.0, 15.0) for _ in range(2000)]
    assert min(us) >= -15.0 and max(us) < 15.0
    ns = np.array([d.normal(0.005, 0.001) for _ in range(4000)])
    assert abs(ns.mean() - 0.005) < 1e-4 and abs(ns.std() - 0.001) < 1e-4


# ---- aero / geo closed forms ------------------------------------------------------------------------
def test_isa_table():
    """vcas2tas(150 m/s, h)."""
    got = aero.vcas2tas(150.0, np.array([0.0, 350.0, 1000.0, 2000.0, 3000.0, 4000.0]))
    np.testing.assert_allclose(got, [150.0, 152.4, 157.0, 164.5, 172.4, 180.8], atol=0.05)
    p, rho, T = aero.vatmos(np.array([0.0, 11000.0, 20000.0]))
    np.testing.assert_allclose(T, [288.15, 216.65, 216.65])
    np.testing.assert_allclose(p[0], 101325.0, rtol=1e-4)
    np.testing.assert_allclose(p[1], 22632.0, rtol=2e-3)


def test_cas_tas_roundtrip():
    h = np.linspace(0, 12000, 25)
    cas = np.linspace(60, 200, 25)
    np.testing.assert_allclose(aero.vtas2cas(aero.vcas2tas(cas, h), h), cas, rtol=1e-12)
    tas, c, m = aero.vcasormach(np.array([0.78, 150.0]), np.array([11000.0, 3000.0]))
    assert abs(m[0] - 0.78) < 1e-12 and abs(c[1] - 150.0) < 1e-12


def test_kwik_closed_form():
    q, d = geo.kwikqdrdist(0.0, 0.0, 1.0, 0.0)
    assert abs(q) < 1e-12 and abs(d - 6371000.0 * np.radians(1.0) / 1852.0) < 1e-9
    q, d = geo.kwikqdrdist(0.0, 0.0, 0.0, 1.0)
    assert abs(q - 90.0) < 1e-12 and abs(d - 6371000.0 * np.radians(1.0) / 1852.0) < 1e-9
    q, d = geo.kwikqdrdist(10.0, 179.5, 10.0, -179.5)          # across the antimeridian
    assert abs(q - 90.0) < 1e-9 and abs(d - 6371000.0 * np.radians(1.0) * np.cos(np.radians(10.0)) / 1852.0) < 1e-6
    la, lo = geo.kwikpos(0.0, 0.0, 90.0, 60.0)
    assert abs(la) < 1e-12 and abs(lo - 1.0) < 1e-12


def test_qdrdist_wgs84():
    q, d = geo.qdrdist(52.0, 4.0, 53.0, 4.0)
    assert abs(q) < 1e-9 and 59.9 < d < 60.2                     # one degree of latitude ~ 60 NM
    q, d = geo.qdrdist(0.0, 0.0, 0.0, 1.0)
    assert abs(q - 90.0) < 1e-9 and abs(d - 6378137.0 * np.radians(1.0) / 1852.0) < 1e-6
    lat, lon = geo.get_point_at_distance(52.0, 4.0, 100.0, 0.0)
    assert abs(lon - 4.0) < 1e-12 and abs(lat - (52.0 + np.degrees(100.0 / 6371.0))) < 1e-9


def test_wrap_and_polygon_helpers():
    assert geo.wrap180_fold(190.0) == -170.0 and geo.wrap180_fold(-190.0) == 170.0 and geo.wrap180_fold(180.0) == 180.0
    assert geo.wrap180_fold(560.0) == 200.0                      # single fold only (functions.py:4-22)
    sq = [np.array(p, float) for p in [(1, 1), (-1, 1), (-1, -1), (1, -1)]]
    assert abs(geo.polygon_area(geo.sort_points_by_angle(sq)) - 4.0) < 1e-12
    vx, vy = np.array([0, 2, 2, 0.0]), np.array([0, 0, 2, 2.0])
    assert geo.point_in_polygon(1.0, 1.0, vx, vy) and not geo.point_in_polygon(3.0, 1.0, vx, vy)
    assert abs(geo.get_hdg(np.array([0.0, 0.0]), np.array([0.0, 1.0])) - 90.0) < 1e-9


# ---- state-based CD -----------------------------------------------------------------------------------
def test_cd_head_on_closed_form():
    lat, lon = np.array([0.0, 0.0]), np.array([0.0, 1.0])
    trk, gs = np.array([90.0, 270.0]), np.array([200.0, 200.0])
    alt, vs = np.array([9000.0, 9000.0]), np.zeros(2)
    cp, lp, inconf, tcpamax, qdr, dist, dcpa, tcpa, tin = statebased.detect(lat, lon, trk, gs, alt, vs)
    d = 6371000.0 * np.radians(1.0)
    assert cp == [(0, 1), (1, 0)] and lp == [] and inconf.all()
    np.testing.assert_allclose(tcpa, d / 400.0, rtol=1e-9)
    np.testing.assert_allclose(dcpa, 0.0, atol=0.01)      # sqrt of the float64 cancellation residue of dist^2 - tcpa^2 dv2
    np.testing.assert_allclose(tin, (d - 9260.0) / 400.0, rtol=1e-9)


def test_cd_vertical_separation_and_lookahead():
    lat, lon = np.array([0.0, 0.0, 0.0]), np.array([0.0, 0.2, 5.0])
    trk, gs = np.array([90.0, 270.0, 270.0]), np.array([200.0, 200.0, 200.0])
    alt, vs = np.array([9000.0, 9000.0 + 700.0, 9000.0]), np.zeros(3)
    cp, lp, inconf, *_ = statebased.detect(lat, lon, trk, gs, alt, vs)
    assert cp == [] and not inconf.any()                         # 700 m apart; third is > 300 s away
    cp2, *_ = statebased.detect(lat, lon, trk, gs, np.array([9000.0, 9100.0, 9000.0]), vs, dtlookahead=2000.0)
    assert set(cp2) == {(0, 1), (1, 0), (0, 2), (2, 0)}


def test_creconfs_round_trip_at_equator():
    """Traffic.creconfs -> detect reproduces the requested geometry (SURVEY.md appendix A.4 check)."""
    rng = np.random.default_rng(0)
    for _ in range(200):
        t = Traffic(simdt=5.0, default_hdg=0.0)
        t.cre("OWN", "A320", aclat=0.0, aclon=0.0, achdg=float(rng.integers(0, 360)), acalt=5000.0, acspd=150.0)
        dpsi, cpa, tlos = int(rng.integers(45, 315)), int(rng.integers(0, 5)), int(rng.integers(100, 1000))
        t.creconfs("INT", "A320", 0, dpsi, cpa, tlos)
        m = statebased.detect_rows(np.array([0]), t.lat, t.lon, t.trk, t.gs, t.alt, t.vs, dtlookahead=1e9)
        assert m["swconfl"][0, 1]
        assert abs(np.sqrt(m["dcpa2"][0, 1]) / 1852.0 - cpa) < 0.03
        assert abs(m["tinconf"][0, 1] - tlos) < 1.5


def test_c_restatement_matches_numpy():
    rng = np.random.default_rng(3)
    n = 600
    lat, lon = 52 + 3 * (rng.random(n) - 0.5), 179.5 + 3 * (rng.random(n) - 0.5)
    lon = (lon + 180) % 360 - 180
    trk, gs = rng.uniform(0, 360, n), rng.uniform(150, 250, n)
    alt = np.round(rng.uniform(3000, 12000, n) / 304.8) * 304.8 + rng.uniform(-20, 20, n)
    vs = np.where(rng.random(n) < 0.7, 0.0, rng.uniform(-15, 15, n))
    cp, lp, inconf, tcpamax, *_ = statebased.detect(lat, lon, trk, gs, alt, vs)
    c = cbind.detect_rows(lat, lon, trk, gs, alt, vs, 9260.0, 304.8, 300.0, pair_cap=100000, nthreads=2)
    assert set(cp) == set(map(tuple, c["confpairs"].tolist())) and len(cp) > 50
    assert set(lp) == set(map(tuple, c["lospairs"].tolist()))
    assert np.array_equal(inconf, c["inconf"])
    np.testing.assert_allclose(tcpamax, c["tcpamax"], rtol=1e-12, atol=1e-9)
    part = cbind.detect_rows(lat, lon, trk, gs, alt, vs, 9260.0, 304.8, 300.0, row0=100, nrows=50)
    assert np.array_equal(part["nconf_row"], c["nconf_row"][100:150])


# ---- kinematics / performance --------------------------------------------------------------------------
def test_phase_and_limits():
    ph = perf.phase_fixwing(np.array([100.0, 150, 150, 150, 150, 150]), np.array([0.0, 5, -5, 0, 5, 0]),
                            np.array([0.0, 200, 200, 1000, 3000, 11000]))
    assert list(ph) == [perf.PH_GD, perf.PH_IC, perf.PH_AP, perf.PH_NA, perf.PH_CL, perf.PH_CR]
    vmin, vmax = perf.v_limits(ph, perf.A320)
    assert vmax[0] == perf.A320.vmaxic and vmin[0] == 0.0 and vmax[3] == perf.A320.vmaxer and vmax[2] == perf.A320.vmaxap


def test_ground_phase_speed_cap_pin():
    """Log-derived pin: HorizontalCREnv straight-flight episodes imply GS 87.3..90 m/s at altitude 0."""
    t = Traffic(simdt=5.0, default_hdg=0.0)
    t.cre("KL001", "A320", acspd=150.0)
    for _ in range(40):
        t.simstep()
    assert 87.3 <= t.gs[0] <= 90.0
    assert abs(t.hdg[0]) < 1e-9 and t.alt[0] == 0.0


def test_horizontal_airborne_variant_keeps_commanded_speed():
    """SURVEY 8d's second C2 input (aircraft created at 3000 m): no ground-phase cap, TAS = vcas2tas(150, 3000 m) = 172.4 m/s
    (the ISA table of SURVEY 8c), and the creconfs geometry still produces conflicts with the ownship."""
    o = envs.HorizontalCREnv(n_intruders=20, cd_enabled=True, draws=philox.PhiloxDraws(31, 0, 0), init_alt=3000.0)
    o.reset()
    for _ in range(4):
        o.step(np.array([0.0]))
    assert np.allclose(o.traf.alt, 3000.0) and np.all(np.abs(o.traf.tas - 172.39) < 0.05)
    assert any(i == 0 for i, _ in o.traf.confpairs)


def test_turn_and_climb_response():
    t = Traffic(simdt=1.0, default_hdg=0.0)
    t.cre("A", "A320", achdg=0.0, acalt=3000.0, acspd=150.0)
    t.stack_hdg("A", 90.0)
    t.selaltcmd(0, 4000.0, 10.0)
    t.simstep()
    rate = np.degrees(aero.g0 * np.tan(np.radians(25.0)) / t.tas[0])
    assert abs(t.hdg[0] - rate) < 1e-6                          # one second of a 25-degree-bank turn
    assert abs(t.vs[0] - 300 * aero.fpm) < 1e-9                 # vertical acceleration limit 300 fpm/s
    for _ in range(200):
        t.simstep()
    assert abs(t.hdg[0] - 90.0) < 1e-9 and abs(t.alt[0] - 4000.0) < 1e-6 and t.vs[0] == 0.0


# ---- env-level pins ---------------------------------------------------------------------------------------
def test_descent_episode_length_pin():
    """Reference logs: DescentEnv episodes last 40..44 steps (200 km at CAS 150 m/s, 30 s per step)."""
    np.random.seed(0)
    rng = np.random.default_rng(0)
    e = envs.DescentEnv()
    lens = []
    for _ in range(6):
        e.reset()
        n = 0
        while True:
            _, _, term, _, info = e.step(rng.uniform(-1.0, 1.0, 1))        # an untrained (random) policy
            n += 1
            if term or n > 60:
                break
        lens.append(n)
    assert all(36 <= n <= 44 for n in lens), lens


def test_env_api_shapes_and_info_keys():
    np.random.seed(1)
    random.seed(1)
    shapes = {"DescentEnv-v0": 4, "HorizontalCREnv-v0": 8, "SectorCREnv-v0": 10, "MergeEnv-v0": 12,
              "PlanWaypointEnv-v0": 4, "VerticalCREnv-v0": 11, "StaticObstacleEnv-v0": 7}
    for name, cls in envs.ENVS.items():
        e = cls()
        obs, info = e.reset()
        assert len(obs) == shapes[name] and all(v.dtype == np.float64 and v.ndim == 1 for v in obs.values())
        obs, r, term, trunc, info = e.step(np.zeros(2) if name in ("SectorCREnv-v0", "MergeEnv-v0", "StaticObstacleEnv-v0") else np.zeros(1))
        assert "total_reward" in info and np.isfinite(r)
    assert envs.MAX_EPISODE_STEPS["MergeEnv-v0"] == 50 and envs.MAX_EPISODE_STEPS["SectorCREnv-v0"] == 200


def test_merge_intruders_follow_route():
    random.seed(0)
    e = envs.MergeEnv()
    e.reset()
    for _ in range(45):
        e.step(np.zeros(2))
    t = e.traf
    assert any(i == 1 for i in t.iactwp[1:])                     # some intruder switched FIX -> RWY
    assert all(s >= 49.0 for s in t.lat)                         # nobody flew off to the far north


def test_sector_polygon_statistics():
    """5..25 vertices, num_ac 5..28 over many draws."""
    np.random.seed(2)
    e = envs.SectorCREnv()
    nv, nac = [], []
    for _ in range(40):
        e.reset()
        nv.append(len(e.poly_lat))
        nac.append(e.num_ac)
        assert e.poly_area >= 2400.0 and e.traf.ntraf == e.num_ac
        assert all(e._inside(la, lo) for la, lo in zip(e.traf.lat, e.traf.lon))
    assert 3 <= min(nv) and max(nv) <= 32 and 5 <= min(nac) and max(nac) <= 32
