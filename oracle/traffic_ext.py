"""Traffic with the upstream features next to conflict detection: multi-waypoint routes with altitude /
speed constraints, VNAV, and ASAS conflict RESOLUTION (MVP) with the resume-navigation logic.  float64 NumPy.

[UPSTREAM-RECALL] restates, on top of oracle/traffic.py (same settings: OpenAP limits, no turn-speed / fly-over / RTA
waypoints, swtod = swtoc = True, no priority rules):
  * ``bluesky/traffic/route.py::Route`` -- addwpt with alt / spd, calcfp (next altitude constraint per waypoint: wptoalt,
    wpxtoalt), direct, getnextwp, getnextqdr
  * ``bluesky/traffic/autopilot.py::Autopilot`` -- update_fms with constraints, ComputeVNAV (top-of-descent / climb-now
    logic, steepness 3000 ft per 10 NM), the continuous VNAV / speed guidance of update(), setVNAV
  * ``bluesky/traffic/asas/resolution.py::ConflictResolution`` -- update, resumenav (resopairs, past-CPA / horizontal LoS /
    bouncing tests, waypoint recovery) and ``bluesky/traffic/asas/mvp.py::MVP`` -- resolve, MVP (per pair), the altitude
    command logic; ``bluesky/traffic/aporasas.py`` -- ASAS commands override the autopilot's while ``active``
The reference itself only ever says ``reso off`` (merge_env.py:157) and builds two-waypoint routes without constraints
(merge_env.py:155-156); nothing in /root/reference pins this module: PARITY UNPINNED (recall of upstream, like the rest of
the simulator core -- DESIGN.md section 5).  Test infrastructure only (see oracle/__init__.py).
"""
import numpy as np

from . import aero, geo, perf, statebased
from .aero import fpm, ft, g0, nm, Rearth
from .traffic import Traffic, BANKDEF, EPS

STEEPNESS = 3000.0 * ft / (10.0 * nm)     # Autopilot.steepness
VSDEF = 1500.0 * fpm
RESO_MAR = 1.01                            # settings.asas_mar (resofach = resofacv)

_EXT_F = ("nextaltco", "xtoalt", "actwp_vs", "dist2vs", "actwp_spd", "nextspd", "spdcon", "vnavvs", "axmax",
          "asas_trk", "asas_tas", "asas_vs", "asas_alt")
_EXT_B = ("swvnavspd", "swvnavvs", "asas_active", "resooff")


def distaccel(v0, v1, axabs):
    return 0.5 * np.abs(v1 * v1 - v0 * v0) / np.maximum(0.001, np.abs(axabs))


class TrafficExt(Traffic):
    """Traffic + constrained routes / VNAV + MVP resolution.  ``reso``: None (detection only) or "MVP";
    ``reso_mode``: 0 horizontal + vertical (upstream default), 1 horizontal only (RMETHH BOTH)."""

    def __init__(self, *a, reso=None, reso_mode=0, resofach=RESO_MAR, resofacv=RESO_MAR, **kw):
        self.reso, self.reso_mode = reso, int(reso_mode)
        self.resofach, self.resofacv = float(resofach), float(resofacv)
        super().__init__(*a, **kw)

    def reset(self):
        super().reset()
        for f in _EXT_F:
            setattr(self, f, np.zeros(0))
        for f in _EXT_B:
            setattr(self, f, np.zeros(0, dtype=bool))
        self.wp = []                    # per aircraft dict of route arrays (lat, lon, alt, spd, toalt, xtoalt)
        self.resopairs = set()          # ordered (own, intruder) index pairs ASAS is still working on

    def cre(self, *a, **kw):
        ok = super().cre(*a, **kw)
        if ok:
            for f in _EXT_F:
                setattr(self, f, np.append(getattr(self, f), 0.0))
            for f in _EXT_B:
                setattr(self, f, np.append(getattr(self, f), False))
            i = self.ntraf - 1
            self.nextaltco[i], self.dist2vs[i], self.nextspd[i], self.spdcon[i], self.actwp_spd[i] = -999.0, -999.0, -999.0, -999.0, -999.0
            self.phase[i] = perf.phase_fixwing(self.tas[i:i + 1], self.vs[i:i + 1], self.alt[i:i + 1])[0]      # perf.create
            self.axmax[i] = perf.axmax(self.phase[i:i + 1], self.perftab)[0]
            self.asas_trk[i], self.asas_tas[i], self.asas_vs[i], self.asas_alt[i] = self.trk[i], self.tas[i], 0.0, self.alt[i]
            self.wp.append(None)
        return ok

    def delete(self, idx):
        raise NotImplementedError("TrafficExt keeps its aircraft (index-keyed resopairs)")

    # ------------------------------------------------------------------ routes
    def set_route(self, idx, lat, lon, alt=None, spd=None, vnav=True):
        """ADDWPT x n (waypoints with optional altitude [m] / CAS [m/s] constraints; < 0 = none) + LNAV / VNAV ON: the first
        waypoint becomes active (Route.direct), the next-leg direction is known from the start."""
        lat, lon = np.asarray(lat, dtype=np.float64), np.asarray(lon, dtype=np.float64)
        n = len(lat)
        alt = np.full(n, -999.0) if alt is None else np.asarray(alt, dtype=np.float64)
        spd = np.full(n, -999.0) if spd is None else np.asarray(spd, dtype=np.float64)
        # Route.calcfp: next altitude constraint at or after each waypoint, and the distance to it
        toalt, xtoalt = np.full(n, -999.0), np.zeros(n)
        distto = np.zeros(n)
        for k in range(n - 1):
            distto[k + 1] = float(geo.qdrdist(lat[k], lon[k], lat[k + 1], lon[k + 1])[1]) * nm
        t, x = -999.0, 0.0
        for k in range(n - 1, -1, -1):
            if alt[k] >= 0.0:
                t, x = alt[k], 0.0
            else:
                x = x + distto[k + 1] if k != n - 1 else 0.0
            toalt[k], xtoalt[k] = t, x
        self.wp[idx] = dict(lat=lat, lon=lon, alt=alt, spd=spd, toalt=toalt, xtoalt=xtoalt, n=n)
        self.routes[idx] = list(zip(lat.tolist(), lon.tolist()))
        self.swvnav[idx] = bool(vnav)
        self.swvnavspd[idx] = bool(vnav)
        self._direct(idx, 0)

    def _next_qdr(self, idx, k):
        w = self.wp[idx]
        if 0 <= k < w["n"] - 1:
            return float(geo.qdrdist(w["lat"][k], w["lon"][k], w["lat"][k + 1], w["lon"][k + 1])[0])
        return -999.0

    def _direct(self, idx, k):
        """Route.direct(): waypoint k becomes the active one (also the waypoint recovery after a resolved conflict)."""
        w = self.wp[idx]
        self.iactwp[idx] = k
        self.swlastwp[idx] = k == w["n"] - 1
        self.actwp_lat[idx], self.actwp_lon[idx] = w["lat"][k], w["lon"][k]
        q, d = geo.qdrdist(self.lat[idx], self.lon[idx], w["lat"][k], w["lon"][k])
        self.curlegdir[idx] = float(q)
        self.next_qdr[idx] = self._next_qdr(idx, k)
        self.turndist[idx] = 0.0
        self.nextspd[idx] = w["spd"][k] if w["spd"][k] > 0.0 else -999.0
        if w["alt"][k] >= -0.01:
            self.nextaltco[idx], self.xtoalt[idx] = w["alt"][k], 0.0
        else:
            self.nextaltco[idx], self.xtoalt[idx] = w["toalt"][k], w["xtoalt"][k]
        self.swlnav[idx] = True
        self._compute_vnav(idx, w["toalt"][k], self.xtoalt[idx], float(d) * nm)

    def _compute_vnav(self, i, toalt, xtoalt, dist2wp):
        """Autopilot.ComputeVNAV (once per leg): dist2vs = distance to the active waypoint at which the descent starts,
        actwp_vs = vertical speed of the climb / descent."""
        if toalt < 0.0 or not self.swvnav[i]:
            self.dist2vs[i] = -999999.0
            return
        epsalt = 2.0 * ft
        gs, tas = self.gs[i], self.tas[i]
        if self.alt[i] > toalt + epsalt:
            if self.vs[i] > 0.0001:                     # stop a climb first
                self.vnavvs[i] = 0.0
                self.ap_alt[i] = self.alt[i]
                self.selalt[i] = self.alt[i]
            self.nextaltco[i], self.xtoalt[i] = toalt, xtoalt
            descdist = abs(self.alt[i] - toalt) / STEEPNESS
            self.dist2vs[i] = descdist - xtoalt
            if dist2wp - 1.02 * self.turndist[i] < self.dist2vs[i]:       # late: use what is left of the leg
                self.ap_alt[i] = self.nextaltco[i]
                t2go = dist2wp / max(0.01, gs)
                self.actwp_vs[i] = (self.nextaltco[i] - self.alt[i]) / max(0.01, t2go)
            elif xtoalt < descdist:                     # top of descent on this leg
                self.actwp_vs[i] = -abs(STEEPNESS) * (gs + (tas if gs < 0.2 * tas else 0.0))
            else:
                self.actwp_vs[i] = 0.0
        elif self.alt[i] < toalt - 10.0 * ft:           # climb as soon as possible
            if self.vs[i] < -0.0001:
                self.vnavvs[i] = 0.0
                self.ap_alt[i] = self.alt[i]
                self.selalt[i] = self.alt[i]
            self.nextaltco[i], self.xtoalt[i] = toalt, xtoalt
            self.ap_alt[i] = self.nextaltco[i]
            self.dist2vs[i] = 99999.0
            t2go = max(0.1, dist2wp + xtoalt) / max(0.01, gs)
            self.actwp_vs[i] = max(STEEPNESS * gs, (self.nextaltco[i] - self.alt[i]) / t2go)
        else:
            self.dist2vs[i] = -999.0

    # ------------------------------------------------------------------ simulation
    def update(self, fms_ready=True):
        dt = self.simdt
        tab = self.perftab
        # ---- Autopilot.update ------------------------------------------------------------
        qdr, dnm = geo.qdrdist(self.lat, self.lon, self.actwp_lat, self.actwp_lon)
        qdr = np.array(qdr, dtype=np.float64)
        dist2wp = np.array(dnm, dtype=np.float64) * nm
        if fms_ready:
            self._update_fms_ext(qdr, dist2wp)
        # VNAV: descend as late as possible, climb as soon as possible
        startdescorclimb = (self.nextaltco >= -0.1) & (
            ((self.alt > self.nextaltco) & (dist2wp < self.dist2vs + self.turndist)) | (self.alt < self.nextaltco))
        self.swvnavvs = self.swvnav & np.where(self.swlnav, startdescorclimb,
                                               dist2wp <= np.maximum(0.1 * nm, self.turndist))
        self.vnavvs = np.where(self.swvnavvs, self.actwp_vs, self.vnavvs)
        selvs_eff = np.where(np.abs(self.selvs) > 0.1, self.selvs, VSDEF)
        self.ap_vs = np.where(self.swvnavvs, self.vnavvs, selvs_eff)
        self.ap_alt = np.where(self.swvnavvs, self.nextaltco, self.selalt)
        self.selalt = np.where(self.swvnavvs, self.nextaltco, self.selalt)
        self.ap_trk = np.where(self.swlnav, qdr % 360.0, self.ap_trk)
        # FMS speed guidance: decelerate / accelerate in time for the next speed constraint
        nexttas = aero.vcasormach2tas(self.nextspd, self.alt)
        dxspdconchg = distaccel(self.tas, nexttas, self.axmax)
        usenextspdcon = (dist2wp < dxspdconchg) & (self.nextspd > -990.0) & self.swvnavspd & self.swvnav & self.swlnav
        self.selspd = np.where(usenextspdcon, self.nextspd,
                               np.where((self.spdcon >= 0.0) & self.swvnavspd, self.actwp_spd, self.selspd))
        self.ap_tas = aero.vcasormach2tas(self.selspd, self.alt)
        # ---- ASAS: detection, resolution, resume navigation --------------------------------
        if self.cd_enabled:
            self._cd_inputs = tuple(np.array(x, dtype=np.float64) for x in (self.lat, self.lon, self.trk, self.gs, self.alt, self.vs))
            (self.confpairs, self.lospairs, self.inconf, self.tcpamax, *rest) = statebased.detect(
                self.lat, self.lon, self.trk, self.gs, self.alt, self.vs, self.rpz, self.hpz, self.dtlookahead)
            self.cd_qdr, self.cd_dist, self.cd_dcpa, self.cd_tcpa, self.cd_tinconf = rest
            if self.reso == "MVP":
                if self.confpairs:
                    self._mvp_resolve()
                self._resumenav()
        # ---- APorASAS.update --------------------------------------------------------------
        act = self.asas_active
        p_trk = np.where(act, self.asas_trk, self.ap_trk)
        p_tas = np.where(act, self.asas_tas, self.ap_tas)
        p_alt = np.where(act, self.asas_alt, self.ap_alt)
        p_vs = np.abs(np.where(act, self.asas_vs, self.ap_vs))
        p_hdg = p_trk % 360.0
        # ---- perf.update + limits ---------------------------------------------------------
        self.phase = perf.phase_fixwing(self.tas, self.vs, self.alt)
        amax = perf.axmax(self.phase, tab)
        self.axmax = amax
        p_tas, p_vs, p_alt = perf.limits(p_tas, p_vs, p_alt, self.ax, self.phase, self.tas, tab)
        # ---- update_airspeed / groundspeed / pos (as oracle/traffic.py, no wind) ----------
        dspd = p_tas - self.tas
        need_ax = np.abs(dspd) > np.abs(dt * amax)
        self.ax = need_ax * np.sign(dspd) * amax
        self.tas = np.where(need_ax, self.tas + self.ax * dt, p_tas)
        self.cas = aero.vtas2cas(self.tas, self.alt)
        self.M = aero.vtas2mach(self.tas, self.alt)
        turnrate = np.degrees(g0 * np.tan(BANKDEF) / np.maximum(self.tas, EPS))
        delhdg = (p_hdg - self.hdg + 180.0) % 360.0 - 180.0
        swhdgsel = np.abs(delhdg) > np.abs(dt * turnrate)
        self.hdg = np.where(swhdgsel, self.hdg + dt * turnrate * np.sign(delhdg), p_hdg) % 360.0
        delta_alt = p_alt - self.alt
        self.swaltsel = np.abs(delta_alt) > 1.05 * np.maximum(np.abs(dt * p_vs), np.abs(dt * self.vs))
        target_vs = self.swaltsel * np.sign(delta_alt) * np.abs(p_vs)
        delta_vs = target_vs - self.vs
        need_az = np.abs(delta_vs) > 300.0 * fpm
        az = need_az * np.sign(delta_vs) * (300.0 * fpm)
        self.vs = np.where(need_az, self.vs + az * dt, target_vs)
        self.vs = np.where(np.isfinite(self.vs), self.vs, 0.0)
        hr = np.radians(self.hdg)
        self.gsnorth = self.tas * np.cos(hr)
        self.gseast = self.tas * np.sin(hr)
        self.gs = self.tas.copy()
        self.trk = self.hdg.copy()
        self.alt = np.where(self.swaltsel, np.round(self.alt + self.vs * dt, 6), p_alt)
        self.lat = self.lat + np.degrees(dt * self.gsnorth / Rearth)
        coslat = np.cos(np.radians(self.lat))
        self.lon = self.lon + np.degrees(dt * self.gseast / coslat / Rearth)
        self.distflown = self.distflown + self.gs * dt

    def _update_fms_ext(self, qdr, dist2wp):
        """Autopilot.update_fms with altitude / speed constraints (reached logic as oracle/traffic.py)."""
        next_qdr = np.where(self.next_qdr < -900.0, qdr, self.next_qdr)
        turnrad = self.tas * self.tas / (np.maximum(0.01, np.tan(BANKDEF)) * g0)
        self.turndist = np.abs(turnrad * np.tan(np.radians(0.5 * np.abs(geo.degto180(qdr % 360.0 - next_qdr % 360.0)))))
        close2wp = dist2wp / np.maximum(0.0001, np.abs(self.gs)) < 4.0
        tooclose = close2wp & (np.abs(geo.degto180(self.trk % 360.0 - qdr % 360.0)) > 90.0)
        passed = np.abs(geo.degto180(qdr - self.curlegdir)) > 90.0
        reached = self.swlnav & (tooclose | passed | (dist2wp < self.turndist))
        for i in np.where(reached)[0]:
            w = self.wp[i]
            self.actwp_spd[i] = self.nextspd[i]         # speeds are FROM-speeds: the passed waypoint's speed holds on the next leg
            self.spdcon[i] = self.nextspd[i]
            if w is None or self.swlastwp[i]:
                self.swlnav[i] = self.swvnav[i] = self.swvnavspd[i] = False
                continue
            k = self.iactwp[i] + 1                      # Route.getnextwp
            self.iactwp[i] = k
            self.swlastwp[i] = k == w["n"] - 1
            self.nextspd[i] = w["spd"][k]
            toalt = w["toalt"][k]
            self.xtoalt[i] = w["xtoalt"][k]
            self.next_qdr[i] = self._next_qdr(i, k)
            self.actwp_lat[i], self.actwp_lon[i] = w["lat"][k], w["lon"][k]
            q, d = geo.qdrdist(self.lat[i], self.lon[i], w["lat"][k], w["lon"][k])
            qdr[i], dist2wp[i] = float(q), float(d) * nm
            self.curlegdir[i] = qdr[i]
            if w["alt"][k] >= -0.01:
                self.nextaltco[i], self.xtoalt[i] = w["alt"][k], 0.0
            else:
                self.nextaltco[i] = toalt
            if self.swvnavspd[i] and self.actwp_spd[i] >= 0.0:
                self.selspd[i] = self.actwp_spd[i]
            lnq = qdr[i] if self.next_qdr[i] < -900.0 else self.next_qdr[i]
            tr = self.tas[i] * self.tas[i] / (max(0.01, np.tan(BANKDEF)) * g0)
            self.turndist[i] = abs(tr * np.tan(np.radians(0.5 * abs(geo.degto180(qdr[i] % 360.0 - lnq % 360.0)))))
            self._compute_vnav(i, toalt, self.xtoalt[i], dist2wp[i])

    # ------------------------------------------------------------------ MVP
    def _mvp_pair(self, i, j, qdr, dist, tcpa, tlos):
        """MVP.MVP(): velocity change of ``i`` that moves the closest point of approach with ``j`` to the zone's edge."""
        q = np.radians(qdr)
        drel = np.array([np.sin(q) * dist, np.cos(q) * dist, self.alt[j] - self.alt[i]])
        v1 = np.array([self.gseast[i], self.gsnorth[i], self.vs[i]])
        v2 = np.array([self.gseast[j], self.gsnorth[j], self.vs[j]])
        vrel = v2 - v1
        dcpa = drel + vrel * tcpa
        dabsH = np.sqrt(dcpa[0] * dcpa[0] + dcpa[1] * dcpa[1])
        rh = self.rpz * self.resofach
        iH = rh - dabsH
        if dabsH <= 10.0:                               # head-on: push sideways
            dabsH = 10.0
            dcpa[0] = drel[1] / dist * dabsH
            dcpa[1] = -drel[0] / dist * dabsH
        if rh < dist and dabsH < dist:                  # outside the zone: aim at the tangent, not at the CPA distance
            erratum = np.cos(np.arcsin(rh / dist) - np.arcsin(dabsH / dist))
            dv1 = ((rh / erratum - dabsH) * dcpa[0]) / (abs(tcpa) * dabsH)
            dv2 = ((rh / erratum - dabsH) * dcpa[1]) / (abs(tcpa) * dabsH)
        else:
            dv1 = (iH * dcpa[0]) / (abs(tcpa) * dabsH)
            dv2 = (iH * dcpa[1]) / (abs(tcpa) * dabsH)
        hv = self.hpz * self.resofacv
        iV = hv if abs(vrel[2]) > 0.0 else hv - abs(drel[2])
        tsolV = abs(drel[2] / vrel[2]) if abs(vrel[2]) > 0.0 else tlos
        if tsolV > self.dtlookahead:
            tsolV = tlos
            iV = hv
        dv3 = (iV / tsolV) * (-vrel[2] / abs(vrel[2])) if abs(vrel[2]) > 0.0 else iV / tsolV
        return np.array([dv1, dv2, dv3]), tsolV

    def _mvp_resolve(self):
        n = self.ntraf
        dv = np.zeros((n, 3))
        timesolveV = np.ones(n) * 1e9
        for (i, j), qdr, dist, tcpa, tlos in zip(self.confpairs, self.cd_qdr, self.cd_dist, self.cd_tcpa, self.cd_tinconf):
            dv_mvp, tsolV = self._mvp_pair(i, j, qdr, dist, tcpa, tlos)
            if tsolV < timesolveV[i]:
                timesolveV[i] = tsolV
            dv_mvp[2] = 0.5 * dv_mvp[2]                 # cooperative: half the vertical part each
            dv[i] = dv[i] - dv_mvp
            if self.resooff[i]:
                dv[i] = 0.0
        dv = dv.T
        newv = np.array([self.gseast, self.gsnorth, self.vs]) + dv
        newtrack = np.degrees(np.arctan2(newv[0], newv[1])) % 360.0
        newgs = np.sqrt(newv[0] ** 2 + newv[1] ** 2)
        newvs = self.vs if self.reso_mode == 1 else newv[2]
        vmin, vmax = perf.v_limits(self.phase, self.perftab)          # perf.vmin / vmax of the last perf.update
        self.asas_tas = np.maximum(vmin, np.minimum(vmax, newgs))
        vscapped = np.maximum(self.perftab.vsmin, np.minimum(self.perftab.vsmax, newvs))
        self.asas_trk, self.asas_vs = newtrack, vscapped
        asasalttemp = vscapped * timesolveV + self.alt
        signdvs = np.sign(vscapped - self.ap_vs * np.sign(self.selalt - self.alt))
        signalt = np.sign(asasalttemp - self.selalt)
        alt = np.where((signdvs == 0) | (signdvs == signalt), asasalttemp, self.selalt)
        cond = (timesolveV < self.dtlookahead) & (np.abs(dv[2]) > 0.0)
        alt = np.where(cond, asasalttemp, alt)
        self.asas_alt = self.selalt.copy() if self.reso_mode == 1 else alt

    def _resumenav(self):
        """ConflictResolution.resumenav: ASAS stays in command of an aircraft until each of its conflicts is past CPA, out of
        horizontal LoS and not 'bouncing'; then the route's active waypoint is flown direct again."""
        self.resopairs.update(self.confpairs)
        delpairs, change = set(), dict()
        for (i, j) in self.resopairs:
            dx = Rearth * np.radians(self.lon[j] - self.lon[i]) * np.cos(0.5 * np.radians(self.lat[j] + self.lat[i]))
            dy = Rearth * np.radians(self.lat[j] - self.lat[i])
            du, dvn = self.gseast[j] - self.gseast[i], self.gsnorth[j] - self.gsnorth[i]
            past_cpa = dx * du + dy * dvn > 0.0
            hdist = np.sqrt(dx * dx + dy * dy)
            hor_los = hdist < self.rpz
            bouncing = abs(self.trk[i] - self.trk[j]) < 30.0 and hdist < self.rpz * self.resofach
            if (not past_cpa) or hor_los or bouncing:
                change[i] = True
            else:
                change[i] = change.get(i, False)
                delpairs.add((i, j))
        for i, active in change.items():
            self.asas_active[i] = active
            if not active and self.wp[i] is not None and self.iactwp[i] >= 0:
                self._direct(i, self.iactwp[i])
        self.resopairs -= delpairs
