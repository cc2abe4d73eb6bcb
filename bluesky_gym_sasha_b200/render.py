"""rgb_array frames of one env of a BlueSkyVectorEnv.

Each ``frame_*`` function below mirrors the draw calls of the corresponding reference ``_render_frame`` -- same canvas
size, scale, colours, marker lengths and the reference's own screen convention (x grows with cos(bearing), y shrinks with
sin(bearing): north points right) -- from a host snapshot of that ONE env's state (a few hundred bytes; rendering is not on
the step path), and records them on a ``Canvas``.  ``bsg_render`` (csrc/render.cu) paints the list on the device and
the frame comes back as a ``(height, width, 3) uint8`` array: what ``render_mode="rgb_array"`` promises in the reference's
metadata but never delivers (its frames only ever reach a pygame window).  References:
horizontal_cr_env.py:277-395, sector_cr_env.py:341-440, merge_env.py:295-447, descent_env.py:211-275,
plan_waypoint_env.py:224-290, vertical_cr_env.py:311-420, static_obstacle_env.py:350-432.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

NM2KM = 1.852
SKY = (135, 206, 235)
RED, GREY, BLACK, WHITE = (220, 20, 60), (80, 80, 80), (0, 0, 0), (255, 255, 255)
SECTOR_CENTER = (51.990426702297746, 4.376124857109851)          # sector_cr_env.py:16
RWY = (52.36239301495972, 4.713195734579777)                     # merge_env.py:41-42


def kwikqdrdist(lata, lona, latb, lonb):
    """bluesky.tools.geo.kwikqdrdist: bearing [deg 0..360) and distance [NM], flat earth."""
    re = 6371000.0
    dlat = np.radians(latb - lata)
    dlon = np.radians(((lonb - lona) + 180.0) % 360.0 - 180.0)
    cavelat = np.cos(np.radians(lata + latb) * 0.5)
    dangle = np.sqrt(dlat * dlat + dlon * dlon * cavelat * cavelat)
    return np.degrees(np.arctan2(dlon * cavelat, dlat)) % 360.0, re * dangle / 1852.0


def kwikpos(latd1, lond1, qdr, dist_nm):
    dx = dist_nm * np.sin(np.radians(qdr))
    dy = dist_nm * np.cos(np.radians(qdr))
    return latd1 + dy / 60.0, ((lond1 + dx / max(0.01, 60.0 * np.cos(np.radians(latd1)))) + 180.0) % 360.0 - 180.0


class Canvas:
    """Records pygame-style draw calls as bsg_render primitive records."""

    def __init__(self, width, height, background=SKY):
        self.width, self.height, self.background = int(width), int(height), background
        self.prims = []

    @staticmethod
    def _col(c):
        c = {"red": (255, 0, 0), "black": (0, 0, 0)}.get(c, c) if isinstance(c, str) else c
        return np.array([c[0] | (c[1] << 8) | (c[2] << 16)], dtype=np.uint32).view(np.float32)[0]

    def _add(self, kind, x0, y0, x1, y1, w, color):
        self.prims.append((float(kind), float(x0), float(y0), float(x1), float(y1), float(w), self._col(color), 0.0))

    def line(self, color, p0, p1, width=1):
        self._add(_lib.PRIM_LINE, p0[0], p0[1], p1[0], p1[1], width, color)

    def circle(self, color, center, radius, width=0):
        self._add(_lib.PRIM_RING, center[0], center[1], radius, 0.0, width, color)

    def rect(self, color, left, top, w, h):
        self._add(_lib.PRIM_RECT, left, top, left + w, top + h, 0.0, color)

    def polygon(self, color, points, width=0):
        n = len(points)
        if n < 3:
            return
        for k in range(n):
            a, b = points[k], points[(k + 1) % n]
            if width > 0:
                self.line(color, a, b, width)
            else:
                self._add(_lib.PRIM_EDGE_END if k == n - 1 else _lib.PRIM_EDGE, a[0], a[1], b[0], b[1], 0.0, color)

    def array(self):
        return np.asarray(self.prims, dtype=np.float32).reshape(-1, _lib.PRIM_FLOATS)

    def background_u32(self):
        b = self.background
        return int(b[0] | (b[1] << 8) | (b[2] << 16))


# ---- per-env frames ------------------------------------------------------------------------------------------------
def _marker(cv, color, x, y, hdg, ac_px, hd_px, ac_width=4):
    """aircraft symbol: a thick stub and a thin heading line from (x, y) along the reference's screen heading"""
    c, s = np.cos(np.radians(hdg)), np.sin(np.radians(hdg))
    cv.line(color, (x, y), (x + c * ac_px, y - s * ac_px), ac_width)
    cv.line(color, (x, y), (x + c * hd_px, y - s * hd_px), 1)


def frame_horizontal(s):                                   # horizontal_cr_env.py:277-395
    W = H = 512
    md = 200.0
    cv = Canvas(W, H)
    hdg = s["hdg"][0]
    c, si = np.cos(np.radians(hdg)), np.sin(np.radians(hdg))
    ex, ey = c * 8 / md * W, si * 8 / md * W
    cv.line(BLACK, (W / 2 - ex / 2, H / 2 + ey / 2), (W / 2 + ex / 2, H / 2 - ey / 2), 4)
    cv.line(BLACK, (W / 2, H / 2), (W / 2 + c * 50 / md * W, H / 2 - si * 50 / md * W), 1)
    for i in range(1, s["n"]):
        q, d = kwikqdrdist(s["lat"][0], s["lon"][0], s["lat"][i], s["lon"][i])
        col = RED if d < 5.0 else GREY
        x = W / 2 + np.cos(np.radians(q)) * d * NM2KM / md * W
        y = H / 2 - np.sin(np.radians(q)) * d * NM2KM / md * H
        _marker(cv, col, x, y, s["hdg"][i], 3 / md * W, 10 / md * W)
        cv.circle(col, (x, y), 5.0 * NM2KM / md * W, 2)
    q, d = kwikqdrdist(s["lat"][0], s["lon"][0], s["wpt_lat"], s["wpt_lon"])
    col = (155, 155, 155) if s["wpt_reach"] else WHITE
    x, y = W / 2 + np.cos(np.radians(q)) * d * NM2KM / md * W, H / 2 - np.sin(np.radians(q)) * d * NM2KM / md * W
    cv.circle(col, (x, y), 4, 0)
    cv.circle(col, (x, y), 5.0 / md * W, 2)
    return cv


def frame_plan_waypoint(s):                                # plan_waypoint_env.py:224-290
    W = H = 512
    md = 200.0
    cv = Canvas(W, H)
    c, si = np.cos(np.radians(s["hdg"][0])), np.sin(np.radians(s["hdg"][0]))
    cv.line(BLACK, (W / 2, H / 2), (W / 2 + c * 8 / md * W, H / 2 - si * 8 / md * W), 4)
    cv.line(BLACK, (W / 2, H / 2), (W / 2 + c * 50 / md * W, H / 2 - si * 50 / md * W), 1)
    for k in range(5):
        q, d = kwikqdrdist(s["lat"][0], s["lon"][0], s["wpts"][2 * k], s["wpts"][2 * k + 1])
        col = (155, 155, 155) if (s["wpt_reach"] >> k) & 1 else WHITE
        x, y = W / 2 + np.cos(np.radians(q)) * d * NM2KM / md * W, H / 2 - np.sin(np.radians(q)) * d * NM2KM / md * W
        cv.circle(col, (x, y), 4, 0)
        cv.circle(col, (x, y), 5.0 / md * W, 2)
    return cv


def frame_sector(s):                                       # sector_cr_env.py:341-440
    W = H = 512
    cv = Canvas(W, H)
    nv = s["nvert"]
    plat, plon = s["poly"][0:2 * nv:2], s["poly"][1:2 * nv:2]
    px = (plat - SECTOR_CENTER[0]) * 60.0                                   # latlong_to_nm: x north, y east [NM]
    py = (plon - SECTOR_CENTER[1]) * 60.0 * np.cos(np.radians(SECTOR_CENTER[0]))
    pts = np.stack([px, py], axis=1)
    md = max(np.linalg.norm(a - b) for a in pts for b in pts) * NM2KM
    ppk = W / md
    cv.polygon((255, 0, 0), [(W / 2 + p[0] * NM2KM * ppk, H / 2 - p[1] * NM2KM * ppk) for p in pts], width=2)
    pos = []
    for i in range(s["n"]):
        q, d = kwikqdrdist(SECTOR_CENTER[0], SECTOR_CENTER[1], s["lat"][i], s["lon"][i])
        pos.append((W / 2 + np.cos(np.radians(q)) * d * NM2KM * ppk, H / 2 - np.sin(np.radians(q)) * d * NM2KM * ppk))
    _marker(cv, BLACK, pos[0][0], pos[0][1], s["hdg"][0], 10, 20)
    for i in range(1, s["n"]):
        sep = kwikqdrdist(s["lat"][0], s["lon"][0], s["lat"][i], s["lon"][i])[1]
        col = RED if sep < 5.0 else GREY
        _marker(cv, col, pos[i][0], pos[i][1], s["hdg"][i], 3, 20)
        cv.circle(col, pos[i], 5.0 * NM2KM * ppk, 2)
    return cv


def frame_merge(s):                                        # merge_env.py:295-447
    W, H = 750, 500
    md = 500.0
    cv = Canvas(W, H)
    cx, cy = W / 2, H / 2
    cv.circle(WHITE, (cx, cy), 4, 0)
    cv.circle(WHITE, (cx, cy), 10.0 / md * W, 2)
    L = 5000.0 / md * W
    for ang, col, wd in ((180.0, BLACK, 2), (315.0, (3, 252, 11), 4), (45.0, (3, 252, 11), 4)):
        cv.line(col, (cx, cy), (cx + np.cos(np.radians(ang)) * L / 2, cy - np.sin(np.radians(ang)) * L / 2), wd)
    q, d = kwikqdrdist(s["fix_lat"], s["fix_lon"], RWY[0], RWY[1])
    cv.line(WHITE, (cx + np.cos(np.radians(q)) * d * NM2KM / md * W, cy - np.sin(np.radians(q)) * d * NM2KM / md * H),
            (cx + np.cos(np.radians(180.0)) * L / 2, cy - np.sin(np.radians(180.0)) * L / 2), 4)
    for i in range(s["n"]):
        q, d = kwikqdrdist(s["fix_lat"], s["fix_lon"], s["lat"][i], s["lon"][i])
        x, y = cx + np.cos(np.radians(q)) * d * NM2KM / md * W, cy - np.sin(np.radians(q)) * d * NM2KM / md * H
        c, si = np.cos(np.radians(s["hdg"][i])), np.sin(np.radians(s["hdg"][i]))
        if i == 0:
            cv.line(BLACK, (x, y), (x + c * 8 / md * W / 2, y - si * 8 / md * W / 2), 4)
            cv.line(BLACK, (x, y), (x + c * 10 / md * W, y - si * 10 / md * W), 1)
        else:
            col = RED if d < 4.0 else GREY                                  # (distance from the FIX, as the reference has it)
            _marker(cv, col, x, y, s["hdg"][i], 3 / md * W, 10 / md * W)
            cv.circle(col, (x, y), 4.0 * NM2KM / md * W, 2)
    return cv


def _profile_frame(s):
    """side view shared by DescentEnv and VerticalCREnv: ground, target altitude, runway, ownship"""
    W, H = 512, 256
    zero, md, max_alt = 25, 180.0, 5000.0
    cv = Canvas(W, H)
    cv.rect((154, 205, 50), 0, H - 50, W, 50)
    ty = int((-1 * (s["target_alt"] - max_alt) / max_alt) * (H - 50))
    cv.line(WHITE, (0, ty), (W, ty), 1)
    rwy_dist = 200.0 - kwikqdrdist(52.0, 4.0, s["lat"][0], s["lon"][0])[1] * NM2KM
    r0 = int(((rwy_dist + zero) / md) * W)
    cv.line((119, 136, 153), (r0, H - 50), (int(r0 + (30 / md) * W), H - 50), 3)
    ay = int((-1 * (s["alt"][0] - max_alt) / max_alt) * (H - 50))
    a0 = int((zero / md) * W)
    cv.line(BLACK, (a0, ay), (int(a0 + (4 / md) * W), ay), 5)
    return cv, (W, H, zero, md, max_alt)


def frame_descent(s):                                      # descent_env.py:211-275
    return _profile_frame(s)[0]


def frame_vertical(s):                                     # vertical_cr_env.py:311-420
    cv, (W, H, zero, md, max_alt) = _profile_frame(s)
    for i in range(1, s["n"]):
        q, d = kwikqdrdist(s["lat"][0], s["lon"][0], s["lat"][i], s["lon"][i])
        b = s["hdg"][0] - q
        b = b + 360.0 if b < -180.0 else (b - 360.0 if b > 180.0 else b)  # fn.bound_angle_positive_negative_180
        xd, yd = d * NM2KM * np.cos(np.radians(b)), d * NM2KM * np.sin(np.radians(b))
        iy = int((-1 * (s["alt"][i] - max_alt) / max_alt) * (H - 50))
        a0 = int(((zero + xd) / md) * W)
        a1 = int(a0 + (4 / md) * W)
        cv.line(WHITE if abs(yd) > 5.0 else "red", (a0, iy), (a1, iy), int(5 + yd / 20))
        hm, vm = (5.0 * NM2KM / md) * W, (1000 * 0.3048 / max_alt) * H
        for p0, p1 in (((a0 - hm / 2, iy - vm), (a1 + hm / 2, iy - vm)), ((a0 - hm / 2, iy + vm), (a1 + hm / 2, iy + vm)),
                       ((a0 - hm / 2, iy - vm), (a0 - hm / 2, iy + vm)), ((a1 + hm / 2, iy - vm), (a1 + hm / 2, iy + vm))):
            cv.line("black", p0, p1, 1)
    return cv


def frame_static_obstacle(s):                              # static_obstacle_env.py:350-432
    W = H = 512
    md = 350.0
    cv = Canvas(W, H)
    # screen origin: 315 deg, half a screen diagonal from the ownship's initial position (static_obstacle_env.py:112-115)
    slat, slon = kwikpos(52.0, 4.0, 315.0, np.sqrt(2 * (md / 2) ** 2) / NM2KM)
    def to_px(lat, lon):
        q, d = kwikqdrdist(slat, slon, lat, lon)
        return np.sin(np.radians(q)) * d * NM2KM / md * W, -np.cos(np.radians(q)) * d * NM2KM / md * W
    xa, ya = to_px(s["lat"][0], s["lon"][0])
    c, si = np.cos(np.radians(s["hdg"][0])), np.sin(np.radians(s["hdg"][0]))
    cv.line((235, 52, 52), (xa, ya), (xa + si * 8 / md * W, ya - c * 8 / md * W), 5)
    cv.line(BLACK, (xa, ya), (xa + si * 50 / md * W, ya - c * 50 / md * W), 1)
    for k in range(10):
        nv = int(s["poly"][350 + k])
        cv.polygon(BLACK, [to_px(s["poly"][32 * k + 2 * v], s["poly"][32 * k + 2 * v + 1]) for v in range(nv)])
    x, y = to_px(s["wpt_lat"], s["wpt_lon"])
    cv.circle(WHITE, (x, y), 4, 0)
    cv.circle(WHITE, (x, y), 5.0 / md * W, 2)
    return cv


FRAMES = {"HorizontalCREnv-v0": frame_horizontal, "PlanWaypointEnv-v0": frame_plan_waypoint, "SectorCREnv-v0": frame_sector,
          "MergeEnv-v0": frame_merge, "DescentEnv-v0": frame_descent, "VerticalCREnv-v0": frame_vertical,
          "StaticObstacleEnv-v0": frame_static_obstacle}


def snapshot(venv, e):
    """Host copy of env ``e``'s state: what the frame functions read (one small D2H per tensor)."""
    t = venv.t
    pos = t["pos"][e].cpu().numpy()
    kin = t["kin"][e].cpu().numpy().astype(np.float64)
    i32 = t["env_i32"][e].cpu().numpy()
    f64 = t["env_f64"][e].cpu().numpy()
    s = dict(lat=pos[:, 0], lon=pos[:, 1], alt=kin[:, 0], hdg=kin[:, 2], n=int(i32[_lib.I32_NUM_AC]),
             wpt_lat=float(f64[_lib.F64_WPT_LAT]), wpt_lon=float(f64[_lib.F64_WPT_LON]),
             target_alt=float(f64[_lib.F64_TARGET_ALT]), wpts=f64[_lib.F64_WPTS:_lib.F64_WPTS + 10],
             wpt_reach=int(i32[_lib.I32_WPT_REACH]), nvert=int(i32[_lib.I32_NVERT]),
             poly=t["poly"][e].cpu().numpy() if t.get("poly") is not None else None)
    if venv.env_id == "MergeEnv-v0":
        # FIX = get_point_at_distance(RWY, 200 km, bearing 0) (merge_env.py:43-46): due north on the sphere
        s["fix_lat"], s["fix_lon"] = RWY[0] + np.degrees(200.0 / 6371.0), RWY[1]
    return s


def render_canvas(cv, device):
    """Canvas -> (height, width, 3) uint8 frame through bsg_render."""
    lib = _lib.load()
    prims = cv.array()
    dev = torch.device("cuda", device) if not isinstance(device, torch.device) else device
    d_prims = torch.as_tensor(prims, device=dev).contiguous()
    out = torch.empty((cv.height, cv.width, 3), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(lib.bsg_render(C.c_void_p(d_prims.data_ptr()), prims.shape[0], cv.width, cv.height, cv.background_u32(),
                                  C.c_void_p(out.data_ptr()), st))
    return out.cpu().numpy()


def render_env(venv, env_index=0):
    """rgb_array frame of env ``env_index`` of a BlueSkyVectorEnv."""
    if not 0 <= env_index < venv.num_envs:
        raise IndexError("env_index out of range")
    return render_canvas(FRAMES[venv.env_id](snapshot(venv, env_index)), venv.device)
