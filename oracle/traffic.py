"""One BlueSky traffic/simulation instance (the process-global ``bs.traf`` + ``bs.sim`` + ``bs.stack`` of the
reference), float64 NumPy, one element per aircraft.

[UPSTREAM-RECALL] restates, for the settings the reference runs under (no wind, no noise, ASAS
resolution off, OpenAP performance, ``detached=True``):
  * ``bluesky/traffic/traffic.py::Traffic`` -- cre / creconfs / reset / delete / update /
    update_airspeed / update_groundspeed / update_pos            (reference: bs.traf.* call sites,
    e.g. horizontal_cr_env.py:85,91,133; descent_env.py:173,204)
  * ``bluesky/traffic/autopilot.py::Autopilot`` -- select modes (HDG/SPD/ALT-VS), LNAV, update_fms,
    ``route.py`` direct/getnextwp, ``activewpdata.py`` reached/calcturn   (reference: bs.stack.stack
    "HDG"/"SPD"/"addwpt"/"dest", e.g. horizontal_cr_env.py:275, merge_env.py:155-156,292-293)
  * ``bluesky/traffic/aporasas.py`` and the integer-counter timers of ``bluesky/core/simtime.py``
  * ``bluesky/simulation/simulation.py::Simulation.step`` -- stack.process() then traf.update()
    (reference: bs.sim.step(), e.g. horizontal_cr_env.py:109)
Stack text commands are represented as queued callables executed at the start of the next
``simstep()``, which is when upstream's ``stack.process()`` runs them.
Test infrastructure only (see oracle/__init__.py).
"""
import numpy as np

from . import aero, geo, perf, statebased
from .windfield import Windfield
from .aero import fpm, g0, nm, ft, Rearth

FMS_DT = 10.5           # settings.fms_dt
BANKDEF = np.radians(25.0)
EPS = 0.01
ACTWP_LAT0, ACTWP_LON0 = 89.99, 0.0     # ActiveWaypoint.create defaults

_FIELDS_F = ("lat", "lon", "alt", "hdg", "trk", "tas", "gs", "cas", "M", "vs", "gsnorth", "gseast",
             "selspd", "selalt", "selvs", "ax", "ap_trk", "ap_tas", "ap_alt", "ap_vs",
             "actwp_lat", "actwp_lon", "next_qdr", "curlegdir", "turndist", "distflown")
_FIELDS_B = ("swlnav", "swvnav", "swlastwp", "swaltsel")


class Traffic:
    def __init__(self, simdt=1.0, perftab=perf.A320, cd_enabled=False, default_hdg="random",
                 rpz=statebased.RPZ_DEFAULT, hpz=statebased.HPZ_DEFAULT,
                 dtlookahead=statebased.DTLOOK_DEFAULT, rng_randint=None):
        self.simdt = float(simdt)
        self.perftab = perftab
        self.cd_enabled = cd_enabled
        self.default_hdg = default_hdg
        self.rpz, self.hpz, self.dtlookahead = rpz, hpz, dtlookahead
        self._randint = rng_randint or (lambda lo, hi: int(np.random.randint(lo, hi)))
        self.nstep = 0                  # sim steps since bs.init (never reset, like simt)
        self.queue = []                 # pending stack commands
        self.fms_rel_freq = max(1, int(FMS_DT // self.simdt))
        self.wind = Windfield()
        self.reset()

    # ------------------------------------------------------------------ bookkeeping
    def reset(self):
        """Traffic.reset(): drops all aircraft; the sim clock and timers keep running."""
        self.ntraf = 0
        self.id = []
        self.routes = []                # per aircraft list of (lat, lon)
        self.iactwp = []
        for f in _FIELDS_F:
            setattr(self, f, np.zeros(0))
        for f in _FIELDS_B:
            setattr(self, f, np.zeros(0, dtype=bool))
        self.phase = np.zeros(0, dtype=np.int32)
        self.inconf = np.zeros(0, dtype=bool)
        self.tcpamax = np.zeros(0)
        self.confpairs = []
        self.lospairs = []
        self.wind.clear()               # Traffic.reset(): self.wind.clear()

    def id2idx(self, acid):
        acid = acid.upper()             # upstream upper-cases the call sign (wrappers/wind.py:56 asks for 'kl001')
        return self.id.index(acid) if acid in self.id else -1

    def _append(self, **vals):
        for f in _FIELDS_F:
            setattr(self, f, np.append(getattr(self, f), float(vals.get(f, 0.0))))
        for f in _FIELDS_B:
            setattr(self, f, np.append(getattr(self, f), bool(vals.get(f, False))))
        self.phase = np.append(self.phase, np.int32(0))
        self.inconf = np.append(self.inconf, False)
        self.tcpamax = np.append(self.tcpamax, 0.0)

    def delete(self, idx):
        if idx < 0 or idx >= self.ntraf:
            return False
        for f in _FIELDS_F + _FIELDS_B + ("phase", "inconf", "tcpamax"):
            setattr(self, f, np.delete(getattr(self, f), idx))
        del self.id[idx], self.routes[idx], self.iactwp[idx]
        self.ntraf -= 1
        return True

    # ------------------------------------------------------------------ creation
    def cre(self, acid, actype="B744", aclat=52.0, aclon=4.0, achdg=None, acalt=0.0, acspd=0.0):
        """Traffic.cre with SI arguments (the kts/ft conversion only exists in the stack parser)."""
        if acid in self.id:
            return False                # "already exists"
        if achdg is None:
            achdg = float(self._randint(1, 360)) if self.default_hdg == "random" else float(self.default_hdg)
        aclon = aclon - 360.0 if aclon > 180.0 else (aclon + 360.0 if aclon < -180.0 else aclon)
        tas, cas, M = aero.vcasormach(acspd, acalt)
        tas, cas, M = float(tas), float(cas), float(M)
        h = np.radians(achdg)
        gsn, gse, gs0, trk0 = tas * np.cos(h), tas * np.sin(h), tas, achdg
        if self.wind.winddim > 0 and acalt > 50.0 * ft:      # Traffic.cre: wind only changes the initial gs / trk
            vn, ve = self.wind.getdata(aclat, aclon, acalt)
            gsn, gse = gsn + vn, gse + ve
            gs0, trk0 = float(np.hypot(gsn, gse)), float(np.degrees(np.arctan2(gse, gsn)))
        self._append(lat=aclat, lon=aclon, alt=acalt, hdg=achdg, trk=trk0, tas=tas, gs=gs0, cas=cas, M=M,
                     vs=0.0, gsnorth=gsn, gseast=gse, selspd=cas, selalt=acalt,
                     selvs=0.0, ap_trk=achdg, ap_tas=tas, ap_alt=acalt, ap_vs=0.0,
                     actwp_lat=ACTWP_LAT0, actwp_lon=ACTWP_LON0, next_qdr=-999.0, curlegdir=-999.0)
        self.id.append(acid)
        self.routes.append([])
        self.iactwp.append(-1)
        self.ntraf += 1
        return True

    def creconfs(self, acid, actype, targetidx, dpsi, dcpa, tlosh, dH=None, tlosv=None):
        """Traffic.creconfs: an intruder whose CPA with ``targetidx`` is ``dcpa`` NM, LoS in ``tlosh`` s."""
        latref, lonref, altref = self.lat[targetidx], self.lon[targetidx], self.alt[targetidx]
        trkref = np.radians(self.trk[targetidx])
        gsref, vsref = self.gs[targetidx], self.vs[targetidx]
        cpa = dcpa * nm
        pzr = statebased.RPZ_DEFAULT
        pzh = statebased.HPZ_DEFAULT
        trk = trkref + np.radians(dpsi)
        if dH is None:
            acalt, acvs = altref, 0.0
        else:
            acalt = altref + dH
            tlosv = tlosh if tlosv is None else tlosv
            acvs = vsref - np.sign(dH) * (abs(dH) - pzh) / tlosv
        gsn, gse = gsref * np.cos(trk), gsref * np.sin(trk)
        vreln, vrele = gsref * np.cos(trkref) - gsn, gsref * np.sin(trkref) - gse
        vrel = np.sqrt(vreln * vreln + vrele * vrele)
        drelcpa = tlosh * vrel + (0.0 if cpa > pzr else np.sqrt(pzr * pzr - cpa * cpa))
        dist = np.sqrt(drelcpa * drelcpa + cpa * cpa)
        rd, rx = drelcpa / dist, cpa / dist
        brn = np.degrees(np.arctan2(-rx * vreln + rd * vrele, rd * vreln + rx * vrele))
        aclat, aclon = geo.kwikpos(latref, lonref, brn, dist / nm)
        acspd = float(aero.vtas2cas(np.sqrt(gsn * gsn + gse * gse), acalt))
        achdg = float(np.degrees(np.arctan2(gse, gsn)))
        self.cre(acid, actype, float(aclat), float(aclon), achdg, float(acalt), acspd)
        self.selaltcmd(self.ntraf - 1, altref, acvs)
        self.vs[-1] = acvs

    # ------------------------------------------------------------------ autopilot commands
    def selhdgcmd(self, idx, hdg):          # stack "HDG acid hdg"
        if self.wind.winddim > 0 and self.alt[idx] > 50.0 * ft:
            # Autopilot.selhdgcmd: with wind the commanded HEADING is turned into the track it produces now
            vn, ve = self.wind.getdata(self.lat[idx], self.lon[idx], self.alt[idx])
            gsn = self.tas[idx] * np.cos(np.radians(hdg)) + vn
            gse = self.tas[idx] * np.sin(np.radians(hdg)) + ve
            self.ap_trk[idx] = np.degrees(np.arctan2(gse, gsn)) % 360.0
        else:
            self.ap_trk[idx] = hdg
        self.swlnav[idx] = False

    def selspdcmd(self, idx, casmach):      # stack "SPD acid spd" (spd already in m/s CAS or Mach)
        self.selspd[idx] = casmach

    def selaltcmd(self, idx, alt, vspd=None):
        self.selalt[idx] = alt
        self.swvnav[idx] = False
        if vspd is not None:
            self.selvs[idx] = vspd

    def stack_hdg(self, acid, hdg_deg):
        self.queue.append(lambda: self.selhdgcmd(self.id2idx(acid), float(hdg_deg)))

    def stack_spd(self, acid, spd_kts):
        """The SPD parser takes knots (or Mach when 0.1 < x < 1) and multiplies by aero.kts."""
        def run():
            v = float(spd_kts)
            self.selspdcmd(self.id2idx(acid), v if 0.1 < v < 1.0 else v * aero.kts)
        self.queue.append(run)

    def stack_addwpt(self, acid, lat, lon):
        self.queue.append(lambda: self._addwpt(self.id2idx(acid), float(lat), float(lon), dest=False))

    def stack_dest(self, acid, lat, lon):
        self.queue.append(lambda: self._addwpt(self.id2idx(acid), float(lat), float(lon), dest=True))

    def _addwpt(self, idx, lat, lon, dest):
        route = self.routes[idx]
        route.append((lat, lon))
        first_real = (not dest and len(route) == 1) or (dest and len(route) == 1)
        if first_real:                      # Route.direct(): make it the active waypoint, LNAV on
            self.iactwp[idx] = 0
            self.actwp_lat[idx], self.actwp_lon[idx] = lat, lon
            q, _ = geo.qdrdist(self.lat[idx], self.lon[idx], lat, lon)
            self.curlegdir[idx] = float(q)
            self.next_qdr[idx] = -999.0     # no next leg known when the wp is activated
            self.turndist[idx] = 0.0
            self.swlnav[idx] = True
            self.swlastwp[idx] = False

    # ------------------------------------------------------------------ simulation
    def near_band(self):
        """Ordered pairs whose conflict / LoS decision sits inside the comparison band of SURVEY 8c for the traffic state
        the LAST detection saw (a float32 kernel may decide them either way): (near_conf, near_los) boolean matrices."""
        m = statebased.detect_rows(np.arange(self.ntraf), *self._cd_inputs, self.rpz, self.hpz, self.dtlookahead,
                                   with_margins=True)
        return m["near_conf"], m["near_los"]

    def simstep(self):
        """Simulation.step(): stack.process(); timers step; traf.update()."""
        q, self.queue = self.queue, []
        for cmd in q:
            cmd()
        self.nstep += 1
        fms_ready = (self.nstep % self.fms_rel_freq) == 0
        if self.ntraf:
            self.update(fms_ready)

    def update(self, fms_ready=True):
        dt = self.simdt
        tab = self.perftab
        # ---- Autopilot.update ------------------------------------------------------------
        qdr, dnm = geo.qdrdist(self.lat, self.lon, self.actwp_lat, self.actwp_lon)
        qdr = np.array(qdr, dtype=np.float64)
        dist2wp = np.array(dnm, dtype=np.float64) * nm
        if fms_ready:
            self._update_fms(qdr, dist2wp)
        selvs_eff = np.where(np.abs(self.selvs) > 0.1, self.selvs, 1500.0 * fpm)
        self.ap_vs = selvs_eff                      # swvnavvs is False throughout (VNAV never armed)
        self.ap_alt = self.selalt.copy()
        self.ap_trk = np.where(self.swlnav, qdr % 360.0, self.ap_trk)
        self.ap_tas = aero.vcasormach2tas(self.selspd, self.alt)
        # ---- ASAS (detection only; resolution is off) ------------------------------------
        if self.cd_enabled:
            self._cd_inputs = tuple(np.array(x, dtype=np.float64) for x in (self.lat, self.lon, self.trk, self.gs, self.alt, self.vs))
            (self.confpairs, self.lospairs, self.inconf, self.tcpamax, *rest) = statebased.detect(
                self.lat, self.lon, self.trk, self.gs, self.alt, self.vs,
                self.rpz, self.hpz, self.dtlookahead)
            # per conflict, aligned with confpairs (upstream ConflictDetection.update keeps them as cd.qdr, cd.dist, ...)
            self.cd_qdr, self.cd_dist, self.cd_dcpa, self.cd_tcpa, self.cd_tinconf = rest
        # ---- APorASAS.update --------------------------------------------------------------
        p_trk, p_tas, p_alt = self.ap_trk, self.ap_tas, self.ap_alt
        p_vs = np.abs(self.ap_vs)
        if self.wind.winddim > 0:                   # APorASAS.update: heading that compensates for the wind
            vwn, vwe = self.wind.getdata(self.lat, self.lon, self.alt)
            Vw = np.sqrt(vwn * vwn + vwe * vwe)
            winddir = np.arctan2(vwe, vwn)
            drift = np.radians(p_trk) - winddir
            steer = np.arcsin(np.minimum(1.0, np.maximum(-1.0, Vw * np.sin(drift) / np.maximum(0.001, self.tas))))
            p_hdg = (p_trk + np.degrees(steer)) % 360.0
        else:
            p_hdg = p_trk % 360.0
        # ---- perf.update + limits ---------------------------------------------------------
        self.phase = perf.phase_fixwing(self.tas, self.vs, self.alt)
        amax = perf.axmax(self.phase, tab)
        p_tas, p_vs, p_alt = perf.limits(p_tas, p_vs, p_alt, self.ax, self.phase, self.tas, tab)
        # ---- update_airspeed --------------------------------------------------------------
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
        # ---- update_groundspeed (no wind) -------------------------------------------------
        hr = np.radians(self.hdg)
        if self.wind.winddim == 0:
            self.gsnorth = self.tas * np.cos(hr)
            self.gseast = self.tas * np.sin(hr)
            self.gs = self.tas.copy()
            self.trk = self.hdg.copy()
        else:                                       # Traffic.update_groundspeed with wind (only when airborne)
            applywind = self.alt > 50.0 * ft
            vnwnd, vewnd = self.wind.getdata(self.lat, self.lon, self.alt)
            self.gsnorth = self.tas * np.cos(hr) + vnwnd * applywind
            self.gseast = self.tas * np.sin(hr) + vewnd * applywind
            self.gs = np.where(applywind, np.sqrt(self.gsnorth ** 2 + self.gseast ** 2), self.tas)
            self.trk = np.where(applywind, np.degrees(np.arctan2(self.gseast, self.gsnorth)) % 360.0, self.hdg)
        # ---- update_pos -------------------------------------------------------------------
        self.alt = np.where(self.swaltsel, np.round(self.alt + self.vs * dt, 6), p_alt)
        self.lat = self.lat + np.degrees(dt * self.gsnorth / Rearth)
        coslat = np.cos(np.radians(self.lat))
        self.lon = self.lon + np.degrees(dt * self.gseast / coslat / Rearth)
        self.distflown = self.distflown + self.gs * dt

    def _update_fms(self, qdr, dist2wp):
        """Autopilot.update_fms + ActiveWaypoint.reached, routes without alt/spd constraints."""
        next_qdr = np.where(self.next_qdr < -900.0, qdr, self.next_qdr)
        turnrad = self.tas * self.tas / (np.maximum(0.01, np.tan(BANKDEF)) * g0)
        self.turndist = np.abs(turnrad * np.tan(np.radians(
            0.5 * np.abs(geo.degto180(qdr % 360.0 - next_qdr % 360.0)))))
        close2wp = dist2wp / np.maximum(0.0001, np.abs(self.gs)) < 4.0
        tooclose = close2wp & (np.abs(geo.degto180(self.trk % 360.0 - qdr % 360.0)) > 90.0)
        passed = np.abs(geo.degto180(qdr - self.curlegdir)) > 90.0
        reached = self.swlnav & (tooclose | passed | (dist2wp < self.turndist))
        for i in np.where(reached)[0]:
            route = self.routes[i]
            if self.swlastwp[i] or self.iactwp[i] >= len(route) - 1:
                self.swlnav[i] = False
                self.swvnav[i] = False
                continue
            self.iactwp[i] += 1
            self.swlastwp[i] = self.iactwp[i] == len(route) - 1
            lat, lon = route[self.iactwp[i]]
            self.actwp_lat[i], self.actwp_lon[i] = lat, lon
            q, d = geo.qdrdist(self.lat[i], self.lon[i], lat, lon)
            qdr[i] = float(q)
            dist2wp[i] = float(d) * nm
            self.curlegdir[i] = float(q)
            if self.iactwp[i] < len(route) - 1:
                nlat, nlon = route[self.iactwp[i] + 1]
                self.next_qdr[i] = float(geo.qdrdist(lat, lon, nlat, nlon)[0])
            else:
                self.next_qdr[i] = -999.0
