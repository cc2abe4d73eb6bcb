"""Routes with altitude / speed constraints, VNAV and ASAS resolution (MVP) -- SURVEY 8f-4.

CPU: the oracle's own behaviour (oracle/traffic_ext.py is a recall of upstream: nothing in /root/reference pins it, so these
are the properties the algorithms promise) and the product's host-side Route.calcfp against the oracle's.
GPU: AirspaceTraffic (bsg_traf_pack / bsg_cd_detect / bsg_traf_substep through the C ABI) against the oracle on the same
scenario, state by state."""
import numpy as np
import pytest

from oracle import geo
from oracle.traffic_ext import TrafficExt
from tests.common import angdiff

NM, FT = 1852.0, 0.3048


def route_scenario(n, seed, nwp_max=5):
    """n aircraft around (52, 4) at cruise levels, each with a route of 2..nwp_max waypoints 40-90 km apart that bends by
    up to +-50 deg per leg; about half the waypoints carry an altitude constraint, a third a speed constraint."""
    rng = np.random.default_rng(seed)
    lat = 52.0 + 4.0 * (rng.random(n) - 0.5)
    lon = 4.0 + 6.0 * (rng.random(n) - 0.5)
    hdg = rng.uniform(0.0, 360.0, n)
    alt = rng.integers(20, 38, n) * 1000.0 * FT
    cas = rng.uniform(120.0, 150.0, n)
    nwp = rng.integers(2, nwp_max + 1, n)
    wlat, wlon = np.zeros((n, nwp_max)), np.zeros((n, nwp_max))
    walt, wspd = np.full((n, nwp_max), -999.0), np.full((n, nwp_max), -999.0)
    for i in range(n):
        la, lo, brg = lat[i], lon[i], hdg[i] + rng.uniform(-30, 30)
        level = alt[i]
        for k in range(nwp[i]):
            la, lo = geo.get_point_at_distance(la, lo, rng.uniform(40.0, 90.0), brg)
            brg += rng.uniform(-50.0, 50.0)
            wlat[i, k], wlon[i, k] = la, lo
            if rng.random() < 0.5:
                level = max(1500.0, level + rng.choice([-1.0, -1.0, 1.0]) * rng.uniform(500.0, 2500.0))
                walt[i, k] = level
            if rng.random() < 0.35:
                wspd[i, k] = rng.uniform(100.0, 150.0)
    return dict(lat=lat, lon=lon, hdg=hdg, alt=alt, cas=cas, nwp=nwp, wlat=wlat, wlon=wlon, walt=walt, wspd=wspd)


def conflict_scenario(n, seed, radius_km=150.0):
    """n aircraft spread around a circle (no two start inside each other's zone) flying to (about) its centre at nearly
    the same level: everybody in conflict with somebody; each has a two-waypoint route through the centre and out the
    other side."""
    rng = np.random.default_rng(seed)
    ang = (np.arange(n) + rng.uniform(-0.3, 0.3, n)) * (360.0 / n)
    lat, lon = np.zeros(n), np.zeros(n)
    wlat, wlon = np.zeros((n, 2)), np.zeros((n, 2))
    for i in range(n):
        lat[i], lon[i] = geo.get_point_at_distance(52.0, 4.0, radius_km * rng.uniform(0.8, 1.2), ang[i])
        wlat[i, 0], wlon[i, 0] = geo.get_point_at_distance(52.0, 4.0, rng.uniform(0.0, 8.0), rng.uniform(0, 360))
        wlat[i, 1], wlon[i, 1] = geo.get_point_at_distance(52.0, 4.0, 2.0 * radius_km, ang[i] + 180.0)
    hdg = (ang + 180.0) % 360.0
    alt = 9000.0 + rng.uniform(-120.0, 120.0, n)
    cas = rng.uniform(125.0, 145.0, n)
    return dict(lat=lat, lon=lon, hdg=hdg, alt=alt, cas=cas, nwp=np.full(n, 2), wlat=wlat, wlon=wlon,
                walt=np.full((n, 2), -999.0), wspd=np.full((n, 2), -999.0))


def make_oracle(sc, reso=None, reso_mode=0, vnav=True):
    t = TrafficExt(simdt=1.0, cd_enabled=True, reso=reso, reso_mode=reso_mode, default_hdg=0.0)
    n = len(sc["lat"])
    for i in range(n):
        t.cre(f"AC{i}", "A320", sc["lat"][i], sc["lon"][i], sc["hdg"][i], sc["alt"][i], sc["cas"][i])
    for i in range(n):
        k = sc["nwp"][i]
        t.set_route(i, sc["wlat"][i, :k], sc["wlon"][i, :k], sc["walt"][i, :k], sc["wspd"][i, :k], vnav=vnav)
    return t


# ---------------------------------------------------------------------------------------------------- CPU
def test_oracle_vnav_meets_altitude_constraints_and_follows_the_route():
    sc = route_scenario(12, 3)
    t = make_oracle(sc)
    n = len(sc["lat"])
    met = np.zeros((n, sc["wlat"].shape[1]), dtype=bool)
    passed = np.zeros_like(met)
    for _ in range(2500):
        prev = list(t.iactwp)
        lnav_prev = t.swlnav.copy()
        t.simstep()
        for i in range(n):
            k = prev[i]
            if lnav_prev[i] and (t.iactwp[i] != k or not t.swlnav[i]):          # waypoint k was just passed
                passed[i, k] = True
                d = geo.kwikdist(t.lat[i], t.lon[i], sc["wlat"][i, k], sc["wlon"][i, k]) * NM
                assert d < 12000.0, (i, k, d)                                     # within the turn distance of the fix
                if sc["walt"][i, k] >= 0.0:
                    # descents are planned to end at the fix (ToD logic); climbs start at once and may still be under way
                    if t.alt[i] >= sc["walt"][i, k] - 30.0:
                        met[i, k] = abs(t.alt[i] - sc["walt"][i, k]) < 150.0
                    else:
                        met[i, k] = t.vs[i] > 0.0
    for i in range(n):
        assert passed[i, :sc["nwp"][i]].all(), (i, passed[i])
        assert not t.swlnav[i]                                                    # route flown to its end
        con = sc["walt"][i, :sc["nwp"][i]] >= 0.0
        assert met[i, :sc["nwp"][i]][con].all(), (i, met[i], sc["walt"][i])


def test_oracle_speed_constraints_are_from_speeds():
    """A waypoint's speed holds on the leg AFTER it, and the aircraft starts decelerating before the waypoint so as to
    pass it at that speed (Autopilot.update: usenextspdcon)."""
    t = TrafficExt(simdt=1.0, cd_enabled=False, default_hdg=0.0)
    t.cre("A", "A320", 52.0, 4.0, 0.0, 6000.0, 150.0)
    la1, lo1 = geo.get_point_at_distance(52.0, 4.0, 60.0, 0.0)
    la2, lo2 = geo.get_point_at_distance(la1, lo1, 60.0, 0.0)
    t.set_route(0, [la1, la2], [lo1, lo2], alt=[-999.0, -999.0], spd=[120.0, -999.0])
    cas_at_pass = None
    for _ in range(700):
        k = t.iactwp[0]
        t.simstep()
        if k == 0 and t.iactwp[0] == 1:
            cas_at_pass = t.cas[0]
    assert cas_at_pass is not None and abs(cas_at_pass - 120.0) < 1.5, cas_at_pass
    assert abs(t.cas[0] - 120.0) < 0.1 and abs(t.selspd[0] - 120.0) < 1e-9


@pytest.mark.parametrize("reso_mode", [0, 1])
def test_oracle_mvp_keeps_separation(reso_mode):
    """Unresolved, the converging scenario produces losses of separation; with MVP every pair stays outside (or within a
    few percent of) the protected zone, ASAS hands the aircraft back afterwards, and they fly their routes again."""
    nac = 8
    sc = conflict_scenario(nac, 1)
    def run(reso):
        t = make_oracle(sc, reso=reso, reso_mode=reso_mode, vnav=False)
        worst, ever_active = 1e9, np.zeros(nac, dtype=bool)
        for _ in range(2500):
            t.simstep()
            qd = geo.kwikqdrdist_matrix(t.lat, t.lon, t.lat, t.lon)[1] * NM + 1e9 * np.eye(nac)
            dalt = np.abs(t.alt.reshape(-1, 1) - t.alt.reshape(1, -1)) + 1e9 * np.eye(nac)
            worst = min(worst, float(qd[dalt < t.hpz].min()) if (dalt < t.hpz).any() else 1e9)
            if reso:
                ever_active |= t.asas_active
        return t, worst, ever_active
    _, worst_off, _ = run(None)
    t, worst_on, ever = run("MVP")
    assert worst_off < 0.5 * t.rpz, worst_off
    assert worst_on > 0.97 * t.rpz, (worst_on, t.rpz)
    assert ever.sum() >= nac - 1
    if reso_mode == 1:                                  # (with vertical manoeuvres some pairs are still sorting themselves out)
        assert not t.asas_active.any() and not t.resopairs
        assert (np.array(t.iactwp) == 1).all()          # everybody is past the centre and on the way out


def test_route_tables_match_oracle_calcfp():
    from bluesky_gym_sasha_b200.traffic import route_tables
    sc = route_scenario(40, 11, nwp_max=6)
    t = make_oracle(sc)
    rt_pos, rt_con, rt_dir = route_tables(sc["wlat"], sc["wlon"], sc["walt"], sc["wspd"], sc["nwp"], 6)
    for i in range(40):
        k = sc["nwp"][i]
        w = t.wp[i]
        np.testing.assert_allclose(rt_con[i, :k, 2], w["toalt"], rtol=1e-6, atol=1e-3)
        np.testing.assert_allclose(rt_con[i, :k, 3], w["xtoalt"], rtol=1e-6, atol=0.5)
        want = [t._next_qdr(i, c) for c in range(k)]
        np.testing.assert_allclose(rt_dir[i, :k], want, rtol=0, atol=1e-4)
        assert np.array_equal(rt_pos[i, :k, 0], w["lat"]) and (rt_dir[i, k - 1:] == -999.0).all()


# ---------------------------------------------------------------------------------------------------- GPU
def make_device(sc, reso=None, reso_mode=0, vnav=True, **kw):
    from bluesky_gym_sasha_b200.traffic import AirspaceTraffic
    n = len(sc["lat"])
    g = AirspaceTraffic(n, simdt=1.0, reso=reso, reso_mode=reso_mode, max_wpts=sc["wlat"].shape[1], **kw)
    g.create(sc["lat"], sc["lon"], sc["hdg"], sc["alt"], sc["cas"])
    g.set_routes(np.arange(n), sc["wlat"], sc["wlon"], sc["walt"], sc["wspd"], nwp=sc["nwp"], vnav=vnav)
    return g


def compare_state(g, t, tol, step, stats, exempt, may_exempt=None):
    """Asserts every aircraft (outside ``exempt``) agrees with the oracle within ``tol`` scaled by SLACK -- one simulator
    substep of any discrete switch (top of descent, level-off, speed capture, turn start decided from float32 operands a
    hair from their thresholds) -- and counts the aircraft that agree within ``tol`` itself.  Returns (tight, compared)."""
    import torch
    torch.cuda.synchronize()
    d = dict(lat=g.lat.cpu().numpy(), lon=g.lon.cpu().numpy(), alt=g.altitude.cpu().numpy().astype(np.float64),
             tas=g.tas.cpu().numpy().astype(np.float64), hdg=g.heading.cpu().numpy().astype(np.float64),
             vs=g.vs.cpu().numpy().astype(np.float64))
    err = dict(lat=np.abs(d["lat"] - t.lat), lon=np.abs(d["lon"] - t.lon), alt=np.abs(d["alt"] - t.alt),
               tas=np.abs(d["tas"] - t.tas), hdg=angdiff(d["hdg"], t.hdg), vs=np.abs(d["vs"] - t.vs))
    if may_exempt is not None:          # aircraft whose ORACLE state sat on a knife edge earlier: exempted once they differ
        viol = np.zeros_like(exempt)
        for k, e in err.items():
            viol |= e > SLACK[k]
        exempt |= may_exempt & viol
    ok = ~exempt
    tight = ok.copy()
    for k, e in err.items():
        tight &= e <= tol[k]
        bad = ok & (e > SLACK[k])
        assert not bad.any(), (step, k, np.where(bad)[0][:5], e[bad][:5])
    for k, e in err.items():
        stats[k] = max(stats.get(k, 0.0), float(e[tight].max()) if tight.any() else 0.0)
    iw = g.iactwp.cpu().numpy()
    ln = g.swlnav.cpu().numpy()
    assert np.array_equal(iw[ok], np.array(t.iactwp)[ok]), (step, "iactwp")
    assert np.array_equal(ln[ok], t.swlnav[ok]), (step, "swlnav")
    return int(tight.sum()), int(ok.sum())


TOL = dict(lat=2e-5, lon=3e-5, alt=1.0, tas=2e-2, hdg=2e-2, vs=2e-2)
# what one substep (1 s) of a switch taken a substep apart leaves behind: 250 m along track, a climb-rate ramp step
# (300 fpm/s), one acceleration step, one second of turn at 250 kts
SLACK = dict(lat=2.5e-3, lon=4e-3, alt=25.0, tas=0.6, hdg=3.5, vs=1.6)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [3, 7])
def test_routes_and_vnav_match_oracle(cuda, seed):
    """Multi-waypoint routes with altitude / speed constraints under VNAV, detection on, resolution off: 900 substeps.
    A waypoint switch is decided at the FMS cadence from float32 bearings / distances: an aircraft whose switch test the
    oracle decides within a hair of its threshold may switch one FMS tick apart; it is exempted from then on (counted), and
    so is one that turns back the other way round after overflying a waypoint (see `flipped`)."""
    sc = route_scenario(48, seed)
    t = make_oracle(sc)
    g = make_device(sc)
    n = t.ntraf
    exempt = np.zeros(n, dtype=bool)
    flipped = np.zeros(n, dtype=bool)
    stats = {}
    tight = compared = 0
    for step in range(900):
        t.simstep()
        g.step(1)
        # an aircraft that overflies a waypoint between two FMS ticks sees the bearing flip by 180 deg: which way round it
        # turns back is decided by rounding (the oracle's included) -- a knife edge of the ORACLE's state, recorded here
        flipped |= angdiff(t.ap_trk, t.hdg) > 175.0
        if step % 10 == 9 or step < 20:
            iw = g.iactwp.cpu().numpy()
            ln = g.swlnav.cpu().numpy()
            exempt |= (iw != np.array(t.iactwp)) | (ln != t.swlnav)             # switched one tick apart
            a, b = compare_state(g, t, TOL, step, stats, exempt, may_exempt=flipped)
            tight, compared = tight + a, compared + b
    print(f"routes/VNAV seed {seed}: {int(exempt.sum())}/{n} aircraft exempted (waypoint switch one FMS tick apart, or a 180 deg turn-back); "
          f"{tight}/{compared} aircraft-checks within the tight tolerances, max |error| among them {stats}; "
          f"counters {g.counters()}")
    assert exempt.sum() <= n // 8
    assert tight >= 0.9 * compared
    assert g.counters()["wp_switches"] >= n


@pytest.mark.gpu
@pytest.mark.parametrize("reso_mode", [0, 1])
def test_mvp_resolution_matches_oracle(cuda, reso_mode):
    """Converging traffic under MVP: same conflicts, same ASAS-active flags, same trajectories as the oracle while the
    resolution manoeuvres unfold, and the same separation at the end."""
    sc = conflict_scenario(12, 5)
    t = make_oracle(sc, reso="MVP", reso_mode=reso_mode, vnav=False)
    g = make_device(sc, reso="MVP", reso_mode=reso_mode, vnav=False, cull=False, symmetric=False)
    n = t.ntraf
    if reso_mode == 0:
        # MVP's vertical part branches on `abs(vrel[2]) > 0.0`: two level aircraft are both told to descend at the SAME rate,
        # after which the test sits on rounding noise of the float64 oracle itself.  Every aircraft therefore climbs or
        # descends at its own rate (ALT acid alt vs), so that no pair ever has exactly equal vertical speeds.
        new_alt = sc["alt"] + np.where(np.arange(n) % 2 == 0, 2500.0, -2500.0)
        rates = 3.0 + 0.37 * np.arange(n)
        for i in range(n):
            t.selaltcmd(i, new_alt[i], rates[i])
        g.alt(np.arange(n), new_alt, rates)
    exempt = np.zeros(n, dtype=bool)
    stats = {}
    tol = dict(TOL, lat=2e-4, lon=3e-4, hdg=0.5, tas=0.2, vs=0.2, alt=5.0)      # closed-loop: conflict geometry feeds back
    same_active = tight = compared = 0
    for step in range(400):
        t.simstep()
        g.step(1)
        if step % 5 == 4:
            c = g.conflicts()
            gp = set(map(tuple, c["confpairs"].tolist()))
            op = set(t.confpairs)
            near = t.near_band()[0]
            assert all(near[i, j] for i, j in gp ^ op), (step, sorted(gp ^ op)[:6])
            act = g.asas_active.cpu().numpy()
            same_active += int(np.array_equal(act, t.asas_active))
            a, b = compare_state(g, t, tol, step, stats, exempt)
            tight, compared = tight + a, compared + b
    print(f"MVP mode {reso_mode}: {tight}/{compared} aircraft-checks within the tight tolerances, max |error| among them {stats}; "
          f"ASAS-active flags identical at {same_active}/80 checks; counters {g.counters()}")
    assert same_active >= 76 and tight >= 0.9 * compared
    assert g.counters()["resopair_overflow"] == 0


def dense_scenario(n, seed, box=8.0):
    """n aircraft in a box_deg x box_deg airspace on three flight levels with random tracks: a few conflicts per aircraft."""
    rng = np.random.default_rng(seed)
    lat = 52.0 + box * (rng.random(n) - 0.5)
    lon = 4.0 + box * (rng.random(n) - 0.5)
    hdg = rng.uniform(0.0, 360.0, n)
    alt = 9000.0 + 304.8 * rng.integers(0, 3, n) + rng.uniform(-20.0, 20.0, n)
    cas = rng.uniform(120.0, 150.0, n)
    wlat, wlon = np.zeros((n, 2)), np.zeros((n, 2))
    for k in range(2):
        d = (k + 1) * 150.0 / 111.0
        wlat[:, k] = lat + d * np.cos(np.radians(hdg))
        wlon[:, k] = lon + d * np.sin(np.radians(hdg)) / np.cos(np.radians(lat))
    return dict(lat=lat, lon=lon, hdg=hdg, alt=alt, cas=cas, nwp=np.full(n, 2), wlat=wlat, wlon=wlon,
                walt=np.full((n, 2), -999.0), wspd=np.full((n, 2), -999.0))


@pytest.mark.gpu
def test_mvp_many_aircraft_culled_detection(cuda):
    """1024 aircraft, a few conflicts each, through the default detection (culled + symmetric K2) and the sorted conflict
    list: the ASAS commands of every aircraft (sum of its MVP velocity changes in intruder order, caps, altitude logic) and
    the ASAS-active flags against the oracle, substep by substep while both still hold the same state."""
    import torch
    n = 1024
    sc = dense_scenario(n, 2)
    # spatially coherent storage order, as AirspaceTraffic users at scale would create the aircraft (the culling needs it)
    order = np.lexsort((sc["lon"], np.floor((sc["lat"] - 48.0) / 1.0)))
    sc = {k: (v[order] if isinstance(v, np.ndarray) and v.shape[0] == n else v) for k, v in sc.items()}
    t = make_oracle(sc, reso="MVP", reso_mode=1, vnav=False)
    g = make_device(sc, reso="MVP", reso_mode=1, vnav=False)
    checked = n_conf_seen = 0
    for step in range(6):
        t.simstep()
        g.step(1)
        c = g.conflicts()
        gp, op = set(map(tuple, c["confpairs"].tolist())), set(t.confpairs)
        near = t.near_band()[0]
        assert all(near[i, j] for i, j in gp ^ op), (step, sorted(gp ^ op)[:6])
        n_conf_seen = max(n_conf_seen, len(op))
        same = np.ones(n, dtype=bool)           # aircraft whose conflict lists are identical in both
        for i, j in gp ^ op:
            same[i] = False
        asas = g.t["asas"][:n].cpu().numpy().astype(np.float64)
        act = g.asas_active.cpu().numpy()
        inv = same & (np.bincount(np.array(sorted(op))[:, 0], minlength=n) > 0)
        assert np.array_equal(act[same], t.asas_active[same]), step
        assert np.max(angdiff(asas[inv, 0], t.asas_trk[inv])) < 0.05, (step, np.max(angdiff(asas[inv, 0], t.asas_trk[inv])))
        assert np.max(np.abs(asas[inv, 1] - t.asas_tas[inv])) < 0.05
        assert np.max(np.abs(asas[inv, 2] - t.asas_vs[inv])) < 0.02
        checked += int(inv.sum())
    per_ac = np.bincount(np.array(sorted(set(t.confpairs)))[:, 0], minlength=n)
    print(f"dense airspace: up to {n_conf_seen} conflict pairs per substep, at most {per_ac.max()} per aircraft; "
          f"{checked} aircraft-substeps of ASAS commands compared; counters {g.counters()}")
    assert n_conf_seen > 300 and checked > 600 and g.counters()["resopair_overflow"] == 0
    # and the resolution does what it is for: fewer losses of separation than the same traffic without it
    def los_after(reso, steps=240):
        a = make_device(sc, reso=reso, reso_mode=1, vnav=False)
        tot = 0
        for _ in range(steps // 20):
            a.step(20)
            tot += int(a.last["npairs"][1])
        return tot
    off, on = los_after(None), los_after("MVP")
    print(f"LoS pair-samples over 240 s: {off} without resolution, {on} with MVP")
    assert on < 0.5 * off


@pytest.mark.gpu
def test_airspace_sharded_over_two_gpus_equals_one(cuda):
    """SURVEY 8e for the traffic of 8f-4: ONE airspace with its aircraft block-partitioned over the ranks (a rank owns the
    kinematics of its block; per substep an all-gather of the CD records, detection of the own rows against the whole
    airspace, MVP on the own aircraft) gives the state of the unsharded airspace bit for bit (scripts/traf_sharded_check.py
    under torchrun).  Needs two GPUs in the box: skipped on a single-GPU one."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", os.path.join(root, "scripts", "traf_sharded_check.py"), "5000", "40"],
                       capture_output=True, text=True, timeout=300)
    assert "SHARDED_TRAFFIC_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    print([l for l in r.stdout.splitlines() if l.startswith("world")][0])
