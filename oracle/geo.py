"""Flat-earth ("kwik") and WGS-84 bearing/distance helpers, float64 NumPy.

[UPSTREAM-RECALL] restates ``bluesky/tools/geo.py``.  Reference call sites:
``kwikqdrdist`` horizontal_cr_env.py:169,190,265 / sector_cr_env.py:260,334 / merge_env.py:179,181,203,280;
``kwikdist`` descent_env.py:102; ``kwikdist_matrix`` merge_env.py:195; ``kwikpos`` via ``creconfs``
(horizontal_cr_env.py:133); ``qdrdist`` inside ``Autopilot.update`` (every ``bs.sim.step()``).
Also restates the in-tree helpers of bluesky_gym/envs/common/functions.py (cited per function).
Test infrastructure only (see oracle/__init__.py).
"""
import numpy as np

nm = 1852.0
RE_KWIK = 6371000.0


def wrap180_fold(angle_deg):
    """functions.py:4-22 -- a *single* +-360 fold with strict inequalities (not a modulo)."""
    a = np.asarray(angle_deg, dtype=np.float64)
    return np.where(a > 180.0, a - 360.0, np.where(a < -180.0, a + 360.0, a))


def degto180(a):
    """[UPSTREAM-RECALL] bluesky.tools.misc.degto180: modulo form, range [-180, 180)."""
    return (np.asarray(a, dtype=np.float64) + 180.0) % 360.0 - 180.0


def kwikqdrdist(lata, lona, latb, lonb):
    """Bearing [deg, 0..360) and distance [NM] a->b on the local flat earth."""
    dlat = np.radians(latb - lata)
    dlon = np.radians(((lonb - lona) + 180.0) % 360.0 - 180.0)
    cavelat = np.cos(np.radians(lata + latb) * 0.5)
    dangle = np.sqrt(dlat * dlat + dlon * dlon * cavelat * cavelat)
    dist = RE_KWIK * dangle / nm
    qdr = np.degrees(np.arctan2(dlon * cavelat, dlat)) % 360.0
    return qdr, dist


def kwikdist(lata, lona, latb, lonb):
    return kwikqdrdist(lata, lona, latb, lonb)[1]


def kwikqdrdist_matrix(lata, lona, latb, lonb):
    """(N,M) matrices; row i = aircraft a_i, column j = aircraft b_j."""
    lata = np.asarray(lata, dtype=np.float64).reshape(-1, 1)
    lona = np.asarray(lona, dtype=np.float64).reshape(-1, 1)
    latb = np.asarray(latb, dtype=np.float64).reshape(1, -1)
    lonb = np.asarray(lonb, dtype=np.float64).reshape(1, -1)
    return kwikqdrdist(lata, lona, latb, lonb)


def kwikpos(latd1, lond1, qdr, dist_nm):
    dx = dist_nm * np.sin(np.radians(qdr))
    dy = dist_nm * np.cos(np.radians(qdr))
    dlat = dy / 60.0
    dlon = dx / np.maximum(0.01, 60.0 * np.cos(np.radians(latd1)))
    latd2 = latd1 + dlat
    lond2 = ((lond1 + dlon) + 180.0) % 360.0 - 180.0
    return latd2, lond2


def rwgs84(latd):
    lat = np.radians(latd)
    a = 6378137.0
    b = 6356752.314245
    coslat = np.cos(lat)
    sinlat = np.sin(lat)
    an = a * a * coslat
    bn = b * b * sinlat
    ad = a * coslat
    bd = b * sinlat
    return np.sqrt((an * an + bn * bn) / (ad * ad + bd * bd))


def qdrdist(latd1, lond1, latd2, lond2):
    """WGS-84-radius haversine.  Bearing [deg, -180..180] and distance [NM]."""
    latd1 = np.asarray(latd1, dtype=np.float64)
    latd2 = np.asarray(latd2, dtype=np.float64)
    res1 = rwgs84(0.5 * (latd1 + latd2))
    a = 6378137.0
    r1 = rwgs84(latd1)
    r2 = rwgs84(latd2)
    res2 = 0.5 * (np.abs(latd1) * (r1 + a) + np.abs(latd2) * (r2 + a)) / \
        np.maximum(0.000001, np.abs(latd1) + np.abs(latd2))
    sw = (latd1 * latd2 >= 0.0)
    r = np.where(sw, res1, res2)
    lat1 = np.radians(latd1)
    lon1 = np.radians(lond1)
    lat2 = np.radians(latd2)
    lon2 = np.radians(lond2)
    sin1 = np.sin(0.5 * (lat2 - lat1))
    sin2 = np.sin(0.5 * (lon2 - lon1))
    coslat1 = np.cos(lat1)
    coslat2 = np.cos(lat2)
    root = sin1 * sin1 + coslat1 * coslat2 * sin2 * sin2
    d = 2.0 * r * np.arctan2(np.sqrt(root), np.sqrt(1.0 - root))
    qdr = np.degrees(np.arctan2(np.sin(lon2 - lon1) * coslat2,
                                coslat1 * np.sin(lat2) - np.sin(lat1) * coslat2 * np.cos(lon2 - lon1)))
    return qdr, d / nm


# ---- in-tree helpers: bluesky_gym/envs/common/functions.py -------------------------------------

def get_point_at_distance(lat1, lon1, d_km, bearing, R=6371.0):
    """functions.py:24-42 -- spherical direct problem, distance in km."""
    lat1 = np.radians(lat1)
    lon1 = np.radians(lon1)
    a = np.radians(bearing)
    ang = d_km / R
    lat2 = np.arcsin(np.sin(lat1) * np.cos(ang) + np.cos(lat1) * np.sin(ang) * np.cos(a))
    lon2 = lon1 + np.arctan2(np.sin(a) * np.sin(ang) * np.cos(lat1),
                             np.cos(ang) - np.sin(lat1) * np.sin(lat2))
    return np.degrees(lat2), np.degrees(lon2)


def nm_to_latlong(center, point):
    """functions.py:98-114 -- x is north [NM], y is east [NM]."""
    lat = center[0] + point[0] / 60.0
    lon = center[1] + point[1] / (60.0 * np.cos(np.radians(center[0])))
    return np.array([lat, lon])


def latlong_to_nm(center, point):
    """functions.py:116-132."""
    x = (point[0] - center[0]) * 60.0
    y = (point[1] - center[1]) * 60.0 * np.cos(np.radians(center[0]))
    return np.array([x, y])


def get_hdg(point1, point2):
    """functions.py:150-178 -- great-circle initial bearing, [0, 360)."""
    lat1, lon1 = np.radians(point1)
    lat2, lon2 = np.radians(point2)
    dl = lon2 - lon1
    x = np.sin(dl) * np.cos(lat2)
    y = np.cos(lat1) * np.sin(lat2) - np.sin(lat1) * np.cos(lat2) * np.cos(dl)
    return (np.degrees(np.arctan2(x, y)) + 360.0) % 360.0


def polygon_area(vertices):
    """functions.py:77-96 -- shoelace."""
    v = np.asarray(vertices, dtype=np.float64)
    x, y = v[:, 0], v[:, 1]
    return abs(float(np.sum(x * np.roll(y, -1) - y * np.roll(x, -1)))) / 2.0


def sort_points_by_angle(vertices):
    """functions.py:61-75 -- ascending atan2(y, x) (stable argsort)."""
    v = [np.asarray(p, dtype=np.float64) for p in vertices]
    order = np.argsort([np.arctan2(p[1], p[0]) for p in v])
    return [v[i] for i in order]


def point_in_polygon(px, py, vx, vy):
    """[UPSTREAM-RECALL] areafilter.Poly.checkInside -> matplotlib Path.contains_points.

    Even-odd ray crossing in the (x=lat, y=lon) plane; boundary points are ambiguous upstream and are
    treated as an epsilon band by the parity tests.  Reached from sector_cr_env.py:136,203.
    """
    vx = np.asarray(vx, dtype=np.float64)
    vy = np.asarray(vy, dtype=np.float64)
    x2 = np.roll(vx, -1)
    y2 = np.roll(vy, -1)
    cond = (vy > py) != (y2 > py)
    with np.errstate(divide="ignore", invalid="ignore"):
        xint = vx + (py - vy) * (x2 - vx) / (y2 - vy)
    crossings = np.count_nonzero(cond & (px < xint))
    return (crossings & 1) == 1
