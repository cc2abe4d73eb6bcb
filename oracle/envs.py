"""CPU restatement of the reference's four hot-path environments on top of oracle.traffic.Traffic.

Restated from the reference files directly (cited per method); all arithmetic float64 like the
reference (spaces declared np.float64, e.g. horizontal_cr_env.py:51-62).  Each env takes a ``draws``
object (oracle.philox) so the same logic can be driven either by the reference's process-global
``np.random`` / ``random`` streams in the reference's draw order, or by the Philox stream the device
reset kernels use.  The gymnasium ``TimeLimit`` wrapper (bluesky_gym/__init__.py:9-45) is not part
of these classes; ``MAX_EPISODE_STEPS`` records the registered caps.
Test infrastructure only (see oracle/__init__.py).
"""
import numpy as np

from . import aero, geo
from .philox import GlobalNumpyDraws, GlobalStdlibDraws
from .traffic import Traffic

NM2KM = 1.852
MpS2Kt = 1.94384
MAX_EPISODE_STEPS = {"DescentEnv-v0": 300, "PlanWaypointEnv-v0": 300, "HorizontalCREnv-v0": 300,
                     "VerticalCREnv-v0": 300, "SectorCREnv-v0": 200, "StaticObstacleEnv-v0": 100,
                     "MergeEnv-v0": 50}


def _a1(x):
    return np.array([x], dtype=np.float64)


class _Base:
    SIMDT = 1.0
    N_SUB = 1

    def __init__(self, draws=None, cd_enabled=False, perftab=None, default_hdg="random"):
        self.draws = draws if draws is not None else GlobalNumpyDraws()
        kw = {} if perftab is None else {"perftab": perftab}
        self.traf = Traffic(simdt=self.SIMDT, cd_enabled=cd_enabled, default_hdg=default_hdg,
                            rng_randint=lambda lo, hi: self.draws.randint(lo, hi), **kw)

    def _substeps(self):
        for _ in range(self.N_SUB):
            self.traf.simstep()


# =====================================================================================================
class DescentEnv(_Base):
    """descent_env.py.  DT 1 (:72), 30 substeps (:29,189-190)."""
    SIMDT = 1.0
    N_SUB = 30

    def reset(self):                                   # descent_env.py:162-182
        self.total_reward = 0.0
        self.final_altitude = 0.0
        alt_init = self.draws.randint(2000, 4000)
        self.target_alt = alt_init + self.draws.randint(-500, 500)
        # the reference relies on the delete loop at termination (:201-204); equivalent to a reset
        # whenever the previous episode terminated, which it always does within the 300-step cap
        self.traf.reset()
        self.traf.cre("KL001", actype="A320", acalt=float(alt_init), acspd=150.0)
        self.traf.swvnav[0] = False
        return self._get_obs(), self._get_info()

    def _get_obs(self):                                # descent_env.py:89-117
        t = self.traf
        self.altitude = float(t.alt[0])
        self.vz = float(t.vs[0])
        self.runway_distance = 200.0 - float(geo.kwikdist(52.0, 4.0, t.lat[0], t.lon[0])) * NM2KM
        return {"altitude": _a1((self.altitude - 1500.0) / 3000.0),
                "vz": _a1(self.vz / 5.0),
                "target_altitude": _a1((self.target_alt - 1500.0) / 3000.0),
                "runway_distance": _a1((self.runway_distance - 100.0) / 200.0)}

    def _get_info(self):                               # descent_env.py:119-126
        return {"total_reward": self.total_reward, "final_altitude": self.final_altitude}

    def _get_reward(self):                             # descent_env.py:128-144
        if self.runway_distance > 0 and self.altitude > 0:
            r, done = abs(self.target_alt - self.altitude) * (-5.0 / 3000.0), 0
        elif self.altitude <= 0:
            r, done = -100.0, 1
            self.final_altitude = -100.0
        else:
            r, done = self.altitude * (-50.0 / 3000.0), 1
            self.final_altitude = self.altitude
        self.total_reward += r
        return r, done

    def step(self, action):                            # descent_env.py:146-160,184-206
        vs_cmd = float(np.asarray(action).reshape(-1)[0]) * 12.5
        self.traf.selalt[0] = 1000000.0 if vs_cmd >= 0 else 0.0
        self.traf.selvs[0] = vs_cmd
        self._substeps()
        obs = self._get_obs()
        reward, terminated = self._get_reward()
        info = self._get_info()
        if terminated:
            self.traf.reset()
        return obs, reward, terminated, False, info


# =====================================================================================================
class HorizontalCREnv(_Base):
    """horizontal_cr_env.py.  DT 5 (:72), 10 substeps (:30,108-109).  ``n_intruders`` defaults to the
    reference's 5 (:17); BASELINE.json's config 2 uses 20."""
    SIMDT = 5.0
    N_SUB = 10

    def __init__(self, n_intruders=5, init_alt=0.0, **kw):
        super().__init__(**kw)
        self.n_int = n_intruders
        self.init_alt = float(init_alt)    # reference: 0 (cre without acalt); SURVEY 8d's airborne variant: 3000 m

    def reset(self):                                   # horizontal_cr_env.py:82-101
        t = self.traf
        t.reset()
        self.total_reward = 0.0
        self.total_intrusions = 0
        self.drift_hist = []
        t.cre("KL001", actype="A320", acspd=150.0, acalt=self.init_alt)
        for i in range(self.n_int):                    # :127-133
            dpsi = self.draws.randint(45, 315)
            cpa = self.draws.randint(0, 5)
            tlosh = self.draws.randint(100, 1000)
            t.creconfs(acid=str(i), actype="A320", targetidx=0, dpsi=dpsi, dcpa=cpa, tlosh=tlosh)
        wpt_dis = self.draws.randint(100, 150)         # :135-148
        self.wpt_lat, self.wpt_lon = (float(v) for v in
                                      geo.get_point_at_distance(t.lat[0], t.lon[0], wpt_dis, 0.0))
        self.wpt_reach = 0
        return self._get_obs(), self._get_info()

    def _get_obs(self):                                # horizontal_cr_env.py:150-213
        t = self.traf
        n = self.n_int
        self.ac_hdg = float(t.hdg[0])
        qdr, dis = geo.kwikqdrdist(t.lat[0], t.lon[0], t.lat[1:n + 1], t.lon[1:n + 1])
        bearing = np.radians(geo.wrap180_fold(self.ac_hdg - qdr))
        dh = np.radians(t.hdg[0] - t.hdg[1:n + 1])
        x_dif = -np.cos(dh) * t.gs[1:n + 1]
        y_dif = t.gs[0] - np.sin(dh) * t.gs[1:n + 1]
        wq, wd = geo.kwikqdrdist(t.lat[0], t.lon[0], self.wpt_lat, self.wpt_lon)
        self.waypoint_distance = float(wd) * NM2KM
        self.drift = float(geo.wrap180_fold(self.ac_hdg - float(wq)))
        return {"intruder_distance": dis * NM2KM / 150.0,
                "cos_difference_pos": np.cos(bearing),
                "sin_difference_pos": np.sin(bearing),
                "x_difference_speed": x_dif / 150.0,
                "y_difference_speed": y_dif / 150.0,
                "waypoint_distance": _a1(self.waypoint_distance / 150.0),
                "cos_drift": _a1(np.cos(np.radians(self.drift))),
                "sin_drift": _a1(np.sin(np.radians(self.drift)))}

    def _get_info(self):                               # horizontal_cr_env.py:215-223
        return {"total_reward": self.total_reward, "total_intrusions": self.total_intrusions,
                "average_drift": float(np.mean(self.drift_hist)) if self.drift_hist else float("nan")}

    def _get_reward(self):                             # horizontal_cr_env.py:225-270
        t = self.traf
        r = 0.0
        if self.waypoint_distance < 5.0 and self.wpt_reach != 1:
            self.wpt_reach = 1
            r += 1.0
        d = abs(np.radians(self.drift))
        self.drift_hist.append(d)
        r += d * -0.1
        dis = geo.kwikdist(t.lat[0], t.lon[0], t.lat[1:self.n_int + 1], t.lon[1:self.n_int + 1])
        nint = int(np.count_nonzero(dis < 5.0))
        self.total_intrusions += nint
        r += -1.0 * nint
        self.total_reward += r
        return r, (1 if self.wpt_reach else 0)

    def step(self, action):                            # horizontal_cr_env.py:103-125,272-275
        hdg_cmd = self.ac_hdg + float(np.asarray(action).reshape(-1)[0]) * 45.0
        self.traf.stack_hdg("KL001", hdg_cmd)
        self._substeps()
        obs = self._get_obs()
        reward, terminated = self._get_reward()
        info = self._get_info()
        return obs, reward, terminated, False, info


# =====================================================================================================
class PlanWaypointEnv(_Base):
    """plan_waypoint_env.py.  DT 1 (:69), 10 substeps (:22,178-179), 1 aircraft, 5 waypoints (:15)."""
    SIMDT = 1.0
    N_SUB = 10
    NUM_WAYPOINTS = 5

    def reset(self):                                   # plan_waypoint_env.py:157-172,200-213
        t = self.traf
        self.total_reward = 0.0
        self.waypoints_completed = 0
        t.reset()                                      # (the reference relies on the delete loop at :190-193)
        t.cre("KL001", actype="A320", acspd=150.0)
        self.wpt_lat, self.wpt_lon, self.wpt_reach = [], [], []
        for _ in range(self.NUM_WAYPOINTS):
            dis = self.draws.randint(0, 75)
            hdg = self.draws.randint(0, 359)
            la, lo = geo.get_point_at_distance(t.lat[0], t.lon[0], dis, hdg)
            self.wpt_lat.append(float(la))
            self.wpt_lon.append(float(lo))
            self.wpt_reach.append(0)
        return self._get_obs(), self._get_info()

    def _get_obs(self):                                # plan_waypoint_env.py:82-124
        t = self.traf
        self.ac_hdg = float(t.hdg[0])
        q, d = geo.kwikqdrdist(t.lat[0], t.lon[0], np.array(self.wpt_lat), np.array(self.wpt_lon))
        self.wpt_dis = d * NM2KM
        drift = np.radians(geo.wrap180_fold(self.ac_hdg - q))
        live = 1.0 - np.array(self.wpt_reach, dtype=np.float64)
        return {"waypoint_distance": live * self.wpt_dis / 75.0,
                "cos_difference": live * np.cos(drift),
                "sin_difference": live * np.sin(drift),
                "waypoint_reached": np.array(self.wpt_reach, dtype=np.float64)}

    def _get_info(self):                               # plan_waypoint_env.py:126-133
        return {"total_reward": self.total_reward, "waypoints_completed": self.waypoints_completed}

    def step(self, action):                            # plan_waypoint_env.py:135-155,174-195,215-228
        hdg_cmd = self.ac_hdg + float(np.asarray(action).reshape(-1)[0]) * 45.0
        self.traf.stack_hdg("KL001", hdg_cmd)
        self._substeps()
        obs = self._get_obs()                          # built with the reach flags from BEFORE this step's check
        r = 0.0
        for k in range(self.NUM_WAYPOINTS):
            if self.wpt_dis[k] < 5.0 and self.wpt_reach[k] != 1:
                self.waypoints_completed += 1
                self.wpt_reach[k] = 1
                r += 1.0
        self.total_reward += r
        terminated = 0 if 0 in self.wpt_reach else 1
        return obs, r, terminated, False, self._get_info()


# =====================================================================================================
class VerticalCREnv(_Base):
    """vertical_cr_env.py.  DT 1 (:95), 30 substeps (:40), DescentEnv + 5 creconfs intruders with dH (:42)."""
    SIMDT = 1.0
    N_SUB = 30
    NUM_INTRUDERS = 5

    def reset(self):                                   # vertical_cr_env.py:258-281, 185-200
        t = self.traf
        t.reset()
        self.total_reward = 0.0
        self.total_intrusions = 0
        self.final_altitude = 0.0
        alt_init = self.draws.randint(2000, 4000)
        self.target_alt = alt_init + self.draws.randint(-500, 500)
        t.cre("KL001", actype="A320", acalt=float(alt_init), acspd=150.0)
        t.swvnav[0] = False
        altitude, spd = float(t.alt[0]), float(t.gs[0])
        for i in range(self.NUM_INTRUDERS):
            dpsi = self.draws.randint(45, 315)
            cpa = self.draws.randint(0, 5)
            tlosh = self.draws.randint(100, int((200 * 0.9) * 1000 / spd))
            average_tod = (200 * 1000 / spd) - 2 * self.target_alt / 12.5
            if tlosh > average_tod:
                dH = self.draws.randint(int(-altitude + 500), int((self.target_alt - altitude) + 100))
            else:
                dH = self.draws.randint(int((self.target_alt - altitude) - 500), int((self.target_alt - altitude) + 500))
            t.creconfs(acid=str(i), actype="A320", targetidx=0, dpsi=dpsi, dcpa=cpa, tlosh=tlosh, dH=dH, tlosv=1e11)
            t.alt[i + 1] = t.alt[0] + dH
            t.selaltcmd(i + 1, t.alt[0] + dH, 0)
        return self._get_obs(), self._get_info()

    def _get_obs(self):                                # vertical_cr_env.py:114-183
        t = self.traf
        n = self.NUM_INTRUDERS
        self.altitude, self.vz = float(t.alt[0]), float(t.vs[0])
        hdg0 = float(t.hdg[0])
        qdr, dis = geo.kwikqdrdist(t.lat[0], t.lon[0], t.lat[1:n + 1], t.lon[1:n + 1])
        bearing = np.radians(geo.wrap180_fold(hdg0 - qdr))
        dh = np.radians(t.hdg[0] - t.hdg[1:n + 1])
        self.runway_distance = 200.0 - float(geo.kwikdist(52.0, 4.0, t.lat[0], t.lon[0])) * NM2KM
        return {"altitude": _a1((self.altitude - 1500.0) / 3000.0),
                "vz": _a1(self.vz / 5.0),
                "target_altitude": _a1((self.target_alt - 1500.0) / 3000.0),
                "runway_distance": _a1((self.runway_distance - 100.0) / 200.0),
                "intruder_distance": dis * NM2KM / 200.0,
                "cos_difference_pos": np.cos(bearing),
                "sin_difference_pos": np.sin(bearing),
                "altitude_difference": (t.alt[1:n + 1] - self.altitude) / 3000.0,
                "x_difference_speed": -np.cos(dh) * t.gs[1:n + 1] / 150.0,
                "y_difference_speed": (t.gs[0] - np.sin(dh) * t.gs[1:n + 1]) / 150.0,
                "z_difference_speed": t.vs[1:n + 1] - self.vz}

    def _get_info(self):                               # vertical_cr_env.py:202-211
        return {"total_reward": self.total_reward, "total_intrusions": self.total_intrusions,
                "final_altitude": self.final_altitude}

    def _get_reward(self):                             # vertical_cr_env.py:213-240
        t = self.traf
        n = self.NUM_INTRUDERS
        dis = geo.kwikdist(t.lat[0], t.lon[0], t.lat[1:n + 1], t.lon[1:n + 1])
        vert = np.abs(t.alt[0] - t.alt[1:n + 1])
        nint = int(np.count_nonzero((dis < 5.0) & (vert < 1000 * 0.3048)))
        self.total_intrusions += nint
        done = 0
        if self.runway_distance > 0 and self.altitude > 0:
            alt_pen = abs(self.target_alt - self.altitude) * (-5.0 / 3000.0)
        elif self.altitude <= 0:
            alt_pen, done = -100.0, 1
            self.final_altitude = -100.0
        else:
            alt_pen, done = self.altitude * (-50.0 / 3000.0), 1
            self.final_altitude = self.altitude
        r = alt_pen - 50.0 * nint
        self.total_reward += r
        return r, done

    def step(self, action):                            # vertical_cr_env.py:242-256,283-305
        vs_cmd = float(np.asarray(action).reshape(-1)[0]) * 12.5
        self.traf.selalt[0] = 1000000.0 if vs_cmd >= 0 else 0.0
        self.traf.selvs[0] = vs_cmd
        self._substeps()
        obs = self._get_obs()
        reward, terminated = self._get_reward()
        return obs, reward, terminated, False, self._get_info()


# =====================================================================================================
SECTOR_CENTER = np.array([51.990426702297746, 4.376124857109851])     # sector_cr_env.py:16


class SectorCREnv(_Base):
    """sector_cr_env.py.  DT 1 (:77), 5 substeps (:32,120-121).  ``max_ac`` mirrors the device's slot cap."""
    SIMDT = 1.0
    N_SUB = 5
    ALT = 350.0                  # passed to cre as metres (:17,211) although named FL
    NUM_AC_STATE = 4

    def __init__(self, ac_density_mode="normal", max_ac=32, max_vertices=32, **kw):
        super().__init__(**kw)
        self.density_mode = ac_density_mode
        self.max_ac, self.max_vertices = max_ac, max_vertices

    # -- scenario generation ----------------------------------------------------------------------
    def _circle_point(self, R):                        # functions.py:44-59
        a = 2.0 * np.pi * self.draws.uniform(0.0, 1.0)
        return np.array([R * np.cos(a), R * np.sin(a)])

    def _generate_polygon(self):                       # sector_cr_env.py:141-160
        R = np.sqrt(3750.0 / np.pi)
        p = geo.sort_points_by_angle([self._circle_point(R) for _ in range(3)])
        area = geo.polygon_area(p)
        while area < 2400.0 and len(p) < self.max_vertices:
            p.append(self._circle_point(R))
            p = geo.sort_points_by_angle(p)
            area = geo.polygon_area(p)
        self.poly_area = area
        self.poly_points = np.array(p)                 # NM, x north / y east
        ll = np.array([geo.nm_to_latlong(SECTOR_CENTER, q) for q in p])
        self.poly_lat, self.poly_lon = ll[:, 0], ll[:, 1]

    def _generate_waypoints(self):                     # sector_cr_env.py:162-188
        pts = self.poly_points
        nxt = np.roll(pts, -1, axis=0)
        elen = np.sqrt(np.sum((nxt - pts) ** 2, axis=1))
        perim = 0.0
        for l in elen:                                 # same left-to-right accumulation as the reference
            perim += l
        d_list = sorted(self.draws.uniform(0.0, perim) for _ in range(self.num_ac))
        self.wpts = []
        cur, k = 0.0, 0
        for d in d_list:
            while d > cur + elen[k]:
                cur += elen[k]
                k += 1
            frac = (d - cur) / elen[k]
            self.wpts.append(pts[k] + frac * (nxt[k] - pts[k]))

    def _inside(self, lat, lon):                       # sector_cr_env.py:134-139 / areafilter.checkInside
        return bool(geo.point_in_polygon(lat, lon, self.poly_lat, self.poly_lon))

    def _generate_ac(self, max_tries=100000):          # sector_cr_env.py:190-217
        pp = self.poly_points
        min_x, min_y, max_x, max_y = pp[:, 0].min(), pp[:, 1].min(), pp[:, 0].max(), pp[:, 1].max()
        pos = []
        tries = 0
        while len(pos) < self.num_ac and tries < max_tries:
            tries += 1
            p = np.array([self.draws.uniform(min_x, max_x), self.draws.uniform(min_y, max_y)])
            ll = geo.nm_to_latlong(SECTOR_CENTER, p)
            if self._inside(ll[0], ll[1]):
                pos.append(ll)
        for i, ll in enumerate(pos):
            wpt = geo.nm_to_latlong(SECTOR_CENTER, self.wpts[i])
            hdg = float(geo.get_hdg(ll, wpt))
            self.traf.cre("KL001" if i == 0 else str(i), actype="A320", aclat=float(ll[0]),
                          aclon=float(ll[1]), achdg=hdg, acspd=150.0, acalt=self.ALT)

    def reset(self):                                   # sector_cr_env.py:87-115
        self.traf.reset()
        self.total_reward = 0.0
        self.total_intrusions = 0
        self.drift_hist = []
        self._generate_polygon()
        if self.density_mode == "normal":
            rho = self.draws.normal(0.005, 0.001)
        else:
            rho = self.draws.uniform(0.003, 0.007)
        self.num_ac = int(min(max(np.ceil(rho * self.poly_area), self.NUM_AC_STATE + 1), self.max_ac))
        self._generate_waypoints()
        self._generate_ac()
        return self._get_obs(), self._get_info()

    # -- step -------------------------------------------------------------------------------------
    def _get_obs(self):                                # sector_cr_env.py:237-313
        t = self.traf
        k = self.NUM_AC_STATE
        hdg0 = float(t.hdg[0])
        w = geo.nm_to_latlong(SECTOR_CENTER, self.wpts[0])
        wq, _ = geo.kwikqdrdist(t.lat[0], t.lon[0], w[0], w[1])
        self.drift = float(geo.wrap180_fold(hdg0 - float(wq)))
        coslat0 = np.cos(np.radians(SECTOR_CENTER[0]))
        px = (t.lat - SECTOR_CENTER[0]) * 60.0 * NM2KM * 1000.0
        py = (t.lon - SECTOR_CENTER[1]) * 60.0 * coslat0 * NM2KM * 1000.0
        dist = np.sqrt((px[1:] - px[0]) ** 2 + (py[1:] - py[0]) ** 2)
        order = np.argsort(dist)[:k] + 1
        hr = np.radians(t.hdg)
        vx, vy = np.cos(hr) * t.tas, np.sin(hr) * t.tas
        dvx, dvy = vx[order] - vx[0], vy[order] - vy[0]
        trk = np.arctan2(dvy, dvx)
        return {"cos(drift)": _a1(np.cos(np.radians(self.drift))),
                "sin(drift)": _a1(np.sin(np.radians(self.drift))),
                "airspeed": _a1((t.tas[0] - 150.0) / 6.0),
                "x_r": (px[order] - px[0]) / 13000.0,
                "y_r": (py[order] - py[0]) / 13000.0,
                "vx_r": dvx / 32.0,
                "vy_r": dvy / 66.0,
                "cos(track)": np.cos(trk),
                "sin(track)": np.sin(trk),
                "distances": (dist[order - 1] - 50000.0) / 15000.0}

    def _get_info(self):                               # sector_cr_env.py:219-225
        return {"total_reward": self.total_reward, "total_intrusions": self.total_intrusions,
                "average_drift": float(np.mean(self.drift_hist)) if self.drift_hist else float("nan")}

    def _get_reward(self):                             # sector_cr_env.py:227-235,324-339
        t = self.traf
        d = abs(np.radians(self.drift))
        self.drift_hist.append(d)
        dis = geo.kwikdist(t.lat[0], t.lon[0], t.lat[1:self.num_ac], t.lon[1:self.num_ac])
        nint = int(np.count_nonzero(dis < 5.0))
        self.total_intrusions += nint
        r = d * -0.1 - 1.0 * nint
        self.total_reward += r
        return r

    def step(self, action):                            # sector_cr_env.py:117-132,315-322
        a = np.asarray(action, dtype=np.float64).reshape(-1)
        t = self.traf
        hdg_new = float(geo.wrap180_fold(t.hdg[0] + a[0] * 22.5))
        spd_new = (t.cas[0] + a[1] * (20.0 / 3.0)) * MpS2Kt
        t.stack_hdg("KL001", hdg_new)
        t.stack_spd("KL001", spd_new)
        self._substeps()
        obs = self._get_obs()
        reward = self._get_reward()
        info = self._get_info()
        truncated = not self._inside(t.lat[0], t.lon[0])
        return obs, reward, False, truncated, info


# =====================================================================================================
RWY_LAT, RWY_LON = 52.36239301495972, 4.713195734579777              # merge_env.py:40-41
FIX_LAT, FIX_LON = (float(v) for v in geo.get_point_at_distance(RWY_LAT, RWY_LON, 200.0, 0.0))  # :43-46


class MergeEnv(_Base):
    """merge_env.py.  DT 5 (:86), 10 substeps (:34,136-137), 1 + 19 aircraft (:36)."""
    SIMDT = 5.0
    N_SUB = 10
    NUM_AC = 20
    NUM_AC_STATE = 5

    def __init__(self, draws=None, **kw):
        super().__init__(draws=draws if draws is not None else GlobalStdlibDraws(), **kw)

    def reset(self):                                   # merge_env.py:103-130,148-158
        t = self.traf
        self.wpt_reach = 0
        t.reset()
        t.queue = []
        self.total_reward = 0.0
        self.drift_hist = []
        self.total_intrusions = 0
        self.faf_reached = 0
        brg = self.draws.uniform(-15.0, 15.0)
        dist = self.draws.uniform(50.0, 200.0)
        lat, lon = geo.get_point_at_distance(FIX_LAT, FIX_LON, dist, brg)
        t.cre("KL001", actype="A320", acspd=100.0, aclat=float(lat), aclon=float(lon),
              achdg=brg - 180.0, acalt=10000.0)
        for i in range(self.NUM_AC - 1):
            brg = self.draws.uniform(-15.0, 15.0)
            dist = self.draws.uniform(20.0, 500.0)
            lat, lon = geo.get_point_at_distance(FIX_LAT, FIX_LON, dist, brg)
            t.cre(f"INT{i}", actype="A320", acspd=100.0, aclat=float(lat), aclon=float(lon),
                  achdg=brg - 180.0, acalt=10000.0)
            t.stack_addwpt(f"INT{i}", FIX_LAT, FIX_LON)      # queued: runs in the first sim step
            t.stack_dest(f"INT{i}", RWY_LAT, RWY_LON)
        return self._get_obs(), self._get_info()

    def _get_obs(self):                                # merge_env.py:160-236
        t = self.traf
        k = self.NUM_AC_STATE
        hdg0 = float(t.hdg[0])
        tgt = (FIX_LAT, FIX_LON) if self.wpt_reach == 0 else (RWY_LAT, RWY_LON)
        wq, wd = geo.kwikqdrdist(t.lat[0], t.lon[0], tgt[0], tgt[1])
        self.drift = float(geo.wrap180_fold(hdg0 - float(wq)))
        self.waypoint_dist = float(wd)                 # NM
        hr = np.radians(t.hdg)
        vx, vy = np.cos(hr) * t.tas, np.sin(hr) * t.tas
        brg, dist = geo.kwikqdrdist(t.lat[0], t.lon[0], t.lat[1:], t.lon[1:])
        order = np.argsort(dist)[:k]
        dm = dist[order] * NM2KM * 1000.0
        br = np.radians(brg[order])
        dvx, dvy = vx[order + 1] - vx[0], vy[order + 1] - vy[0]
        trk = np.arctan2(dvy, dvx)
        return {"cos(drift)": _a1(np.cos(np.radians(self.drift))),
                "sin(drift)": _a1(np.sin(np.radians(self.drift))),
                "airspeed": _a1(t.tas[0]),
                "waypoint_dist": _a1(self.waypoint_dist / 250.0),
                "faf_reached": _a1(float(self.wpt_reach)),
                "x_r": dm * np.cos(br) / 1000000.0,
                "y_r": dm * np.sin(br) / 1000000.0,
                "vx_r": dvx / 150.0,
                "vy_r": dvy / 150.0,
                "cos(track)": np.cos(trk),
                "sin(track)": np.sin(trk),
                "distances": dist[order] / 250.0}

    def _get_info(self):                               # merge_env.py:238-244
        return {"total_reward": self.total_reward, "faf_reach": self.faf_reached,
                "average_drift": float(np.mean(self.drift_hist)) if self.drift_hist else float("nan"),
                "total_intrusions": self.total_intrusions}

    def _get_reward(self):                             # merge_env.py:246-284
        t = self.traf
        r, done = 0.0, 0
        if self.waypoint_dist < 10.0 and self.wpt_reach != 1:
            self.wpt_reach = 1
            self.faf_reached = 1
            r += 1.0
        elif self.waypoint_dist < 20.0 and self.wpt_reach == 1:
            self.faf_reached = 2
            done = 1
        d = abs(np.radians(self.drift))
        self.drift_hist.append(d)
        r += d * -0.1
        dis = geo.kwikdist(t.lat[0], t.lon[0], t.lat[1:], t.lon[1:])
        nint = int(np.count_nonzero(dis < 4.0))
        self.total_intrusions += nint
        r += -1.0 * nint
        self.total_reward += r
        return r, done

    def step(self, action):                            # merge_env.py:132-146,286-293
        a = np.asarray(action, dtype=np.float64).reshape(-1)
        t = self.traf
        hdg_new = float(geo.wrap180_fold(t.hdg[0] + a[0] * 15.0))
        spd_new = (t.cas[0] + a[1] * 20.0) * MpS2Kt
        t.stack_hdg("KL001", hdg_new)
        t.stack_spd("KL001", spd_new)
        self._substeps()
        obs = self._get_obs()
        reward, terminated = self._get_reward()
        info = self._get_info()
        return obs, reward, terminated, False, info


# =====================================================================================================
class StaticObstacleEnv(_Base):
    """static_obstacle_env.py.  DT 1 (:83), 10 substeps (:31), 1 aircraft, 10 polygon no-fly areas (:33).
    ``max_vertices`` mirrors the device's per-obstacle vertex cap."""
    SIMDT = 1.0
    N_SUB = 10
    NUM_OBSTACLES = 10

    def __init__(self, max_vertices=16, **kw):
        super().__init__(**kw)
        self.max_vertices = max_vertices

    def _generate_polygon(self, centre):               # static_obstacle_env.py:156-169
        poly_area = self.draws.randint(100, 1000)
        R = np.sqrt(poly_area / np.pi)

        def pt():
            al = 2.0 * np.pi * self.draws.uniform(0.0, 1.0)
            return np.array([R * np.cos(al), R * np.sin(al)])
        p = geo.sort_points_by_angle([pt() for _ in range(3)])
        area = geo.polygon_area(p)
        while area < 50.0 and len(p) < self.max_vertices:
            p.append(pt())
            p = geo.sort_points_by_angle(p)
            area = geo.polygon_area(p)
        return area, [geo.nm_to_latlong(centre, q) for q in p], R

    def _inside_any(self, lat, lon):
        return [bool(geo.point_in_polygon(lat, lon, v[:, 0], v[:, 1])) for v in self.obstacle_vertices]

    def reset(self):                                   # static_obstacle_env.py:96-131
        t = self.traf
        t.reset()
        self.total_reward = 0.0
        self.waypoint_reached = 0
        self.crashed = 0
        self.drift_hist = []
        t.cre("KL001", actype="A320", acspd=150.0, acalt=350.0)
        lat0, lon0 = float(t.lat[0]), float(t.lon[0])
        self.obstacle_centre_lat, self.obstacle_centre_lon = [], []
        for _ in range(self.NUM_OBSTACLES):            # :221-232
            dis = self.draws.randint(20, 150)
            hdg = self.draws.randint(0, 360)
            la, lo = geo.get_point_at_distance(lat0, lon0, dis, hdg)
            self.obstacle_centre_lat.append(float(la))
            self.obstacle_centre_lon.append(float(lo))
        self.obstacle_vertices, self.obstacle_radius = [], []
        for i in range(self.NUM_OBSTACLES):            # :171-196
            _, p, R = self._generate_polygon((self.obstacle_centre_lat[i], self.obstacle_centre_lon[i]))
            self.obstacle_vertices.append(np.array(p))
            self.obstacle_radius.append(R)
        loops = 0                                      # _generate_waypoint :198-219
        while True:
            loops += 1
            dis = self.draws.randint(100, 170)
            hdg = self.draws.randint(0, 360)
            la, lo = geo.get_point_at_distance(lat0, lon0, dis, hdg)
            if not any(self._inside_any(float(la), float(lo))):
                break
            if loops > 1000:
                raise Exception("No waypoints can be generated outside the obstacles.")
        self.wpt_lat, self.wpt_lon, self.wpt_reach = float(la), float(lo), 0
        q, _ = geo.kwikqdrdist(lat0, lon0, self.wpt_lat, self.wpt_lon)
        t.hdg[0] = float(q)
        t.ap_trk[0] = float(q)
        return self._get_obs(), self._get_info()

    def _get_obs(self):                                # static_obstacle_env.py:234-283
        t = self.traf
        hdg = float(t.hdg[0])
        wq, wd = geo.kwikqdrdist(t.lat[0], t.lon[0], self.wpt_lat, self.wpt_lon)
        self.wpt_dis_km = float(wd) * NM2KM
        self.drift = float(geo.wrap180_fold(hdg - float(wq)))
        oq, od = geo.kwikqdrdist(t.lat[0], t.lon[0], np.array(self.obstacle_centre_lat), np.array(self.obstacle_centre_lon))
        brg = np.radians(geo.wrap180_fold(hdg - oq))
        return {"destination_waypoint_distance": _a1(self.wpt_dis_km / 170.0),
                "destination_waypoint_cos_drift": _a1(np.cos(np.radians(self.drift))),
                "destination_waypoint_sin_drift": _a1(np.sin(np.radians(self.drift))),
                "restricted_area_radius": np.array(self.obstacle_radius) / 50.0,
                "restricted_area_distance": od * NM2KM / 170.0,
                "cos_difference_restricted_area_pos": np.cos(brg),
                "sin_difference_restricted_area_pos": np.sin(brg)}

    def _get_info(self):                               # static_obstacle_env.py:285-292
        return {"total_reward": self.total_reward, "waypoint_reached": self.waypoint_reached, "crashed": self.crashed,
                "average_drift": float(np.mean(self.drift_hist)) if self.drift_hist else float("nan")}

    def _get_reward(self):                             # static_obstacle_env.py:294-341
        # NB the waypoint distance and drift are the ones of the LAST _get_obs (not refreshed per substep)
        r = 0.0
        if self.wpt_dis_km < 5.0 and self.wpt_reach != 1:
            self.waypoint_reached = 1
            self.wpt_reach = 1
            r += 1.0
        d = abs(np.radians(self.drift))
        self.drift_hist.append(d)
        r += d * -0.01
        t = self.traf
        nin = sum(self._inside_any(float(t.lat[0]), float(t.lon[0])))
        if nin:
            r += -5.0 * nin
            self.crashed = 1
        done = 1 if (self.wpt_reach == 1 or nin) else 0
        return r, done

    def step(self, action):                            # static_obstacle_env.py:133-154,343-350
        a = np.asarray(action, dtype=np.float64).reshape(-1)
        t = self.traf
        hdg_new = float(geo.wrap180_fold(t.hdg[0] + a[0] * 45.0))
        spd_new = (t.cas[0] + a[1] * (20.0 / 3.0)) * MpS2Kt
        t.stack_hdg("KL001", hdg_new)
        t.stack_spd("KL001", spd_new)
        reward, done = 0.0, 0
        for _ in range(self.N_SUB):
            t.simstep()
            reward, done = self._get_reward()          # per substep; only the last one is returned
            if done:
                break
        obs = self._get_obs()
        self.total_reward += reward
        return obs, reward, done, False, self._get_info()


ENVS = {"DescentEnv-v0": DescentEnv, "HorizontalCREnv-v0": HorizontalCREnv,
        "SectorCREnv-v0": SectorCREnv, "MergeEnv-v0": MergeEnv,
        "PlanWaypointEnv-v0": PlanWaypointEnv, "VerticalCREnv-v0": VerticalCREnv,
        "StaticObstacleEnv-v0": StaticObstacleEnv}
