"""bluesky/traffic/windfield.py::Windfield restated ([UPSTREAM-RECALL], float64 NumPy): wind vectors defined at
lat/lon points, optionally with an altitude profile per point; ``getdata`` interpolates horizontally with
inverse-distance-squared weights on the flat-earth metric (degrees, longitude scaled by the cosine of the mean
latitude) and linearly in altitude on a 100 ft axis.

Reference call sites: ``bs.traf.wind.addpointvne(lat, lon, vnorth, veast, alt)`` (wrappers/wind.py:28) and
``bs.traf.wind.getdata(lat, lon, alt)`` (wrappers/wind.py:58).  Test infrastructure only (see oracle/__init__.py).
"""
import numpy as np

from .aero import ft


class Windfield:
    def __init__(self):
        self.altmax = 45000.0 * ft
        self.altstep = 100.0 * ft
        self.altaxis = np.arange(0.0, self.altmax + self.altstep, self.altstep)
        self.nalt = len(self.altaxis)
        self.clear()

    def clear(self):
        self.winddim = 0                    # 0 none, 1 constant, 2 2-D field, 3 field with altitude profiles
        self.lat = np.zeros(0)
        self.lon = np.zeros(0)
        self.vnorth = np.zeros((self.nalt, 0))
        self.veast = np.zeros((self.nalt, 0))
        self.nvec = 0

    def addpointvne(self, lat, lon, vnorth, veast, windalt=None):
        """Adds points (arrays) with north / east wind components [m/s]; ``vnorth[k, i]`` is point i at
        ``windalt[k]`` (one row and ``windalt=None`` = no altitude dependence)."""
        lat = np.atleast_1d(np.asarray(lat, dtype=np.float64))
        lon = np.atleast_1d(np.asarray(lon, dtype=np.float64))
        vn = np.atleast_2d(np.asarray(vnorth, dtype=np.float64))
        ve = np.atleast_2d(np.asarray(veast, dtype=np.float64))
        for i in range(len(lat)):
            if windalt is None:
                vnaxis = np.full(self.nalt, vn[0, i])
                veaxis = np.full(self.nalt, ve[0, i])
            else:
                wa = np.atleast_1d(np.asarray(windalt, dtype=np.float64))
                vnaxis = np.interp(self.altaxis, wa, vn[:, i])
                veaxis = np.interp(self.altaxis, wa, ve[:, i])
            self.lat = np.append(self.lat, lat[i])
            self.lon = np.append(self.lon, lon[i])
            self.vnorth = np.append(self.vnorth, vnaxis.reshape(-1, 1), axis=1)
            self.veast = np.append(self.veast, veaxis.reshape(-1, 1), axis=1)
            self.nvec += 1
            if self.winddim < 3:
                self.winddim = min(2, self.nvec)
            if windalt is not None:
                self.winddim = 3
        return self.nvec - 1

    def getdata(self, userlat, userlon, useralt=0.0):
        eps = 1e-20
        scalar = np.ndim(userlat) == 0
        lat = np.atleast_1d(np.asarray(userlat, dtype=np.float64)).reshape(1, -1)
        lon = np.atleast_1d(np.asarray(userlon, dtype=np.float64)).reshape(1, -1)
        npos = lat.shape[1]
        alt = np.broadcast_to(np.asarray(useralt, dtype=np.float64), (npos,)) if np.ndim(useralt) else np.full(npos, float(useralt))
        if self.winddim == 0:
            vnorth, veast = np.zeros(npos), np.zeros(npos)
        elif self.winddim == 1:
            vnorth, veast = np.full(npos, self.vnorth[0, 0]), np.full(npos, self.veast[0, 0])
        else:
            plat, plon = self.lat.reshape(-1, 1), self.lon.reshape(-1, 1)
            cavelat = np.cos(np.radians(0.5 * (lat + plat)))
            dy = lat - plat
            dx = cavelat * (lon - plon)
            invd2 = 1.0 / (eps + dx * dx + dy * dy)                     # (nvec, npos)
            horfact = invd2 / np.sum(invd2, axis=0, keepdims=True)
            if self.winddim == 2:
                vnorth = self.vnorth[0, :].dot(horfact)
                veast = self.veast[0, :].dot(horfact)
            else:
                idxalt = np.maximum(0.0, np.minimum(self.altaxis[-1] - eps, alt) / self.altstep)
                ialt = np.floor(idxalt).astype(int)
                falt = idxalt - ialt
                vn0 = np.sum(self.vnorth[ialt, :] * horfact.T, axis=1)
                vn1 = np.sum(self.vnorth[ialt + 1, :] * horfact.T, axis=1)
                ve0 = np.sum(self.veast[ialt, :] * horfact.T, axis=1)
                ve1 = np.sum(self.veast[ialt + 1, :] * horfact.T, axis=1)
                vnorth = (1.0 - falt) * vn0 + falt * vn1
                veast = (1.0 - falt) * ve0 + falt * ve1
        if scalar:
            return float(vnorth[0]), float(veast[0])
        return vnorth, veast
