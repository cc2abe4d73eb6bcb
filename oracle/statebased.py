"""State-based conflict detection (all-pairs CPA), dense float64 NumPy.

[UPSTREAM-RECALL] restates ``bluesky/traffic/asas/statebased.py::StateBased.detect`` and the pair-list
part of ``detection.py::ConflictDetection.update``.  The reference never switches ASAS on (the only
ASAS command it issues is ``reso off``, merge_env.py:157); BASELINE.json's north_star adds
"StateBased CD every step" on top, and this function is what the CUDA kernels are checked against.
Memory is O(N^2) (about 15 dense temporaries), so use it for N up to a few thousand; for sampled
rows at N = 100k use ``detect_rows`` below or the C restatement oracle/cd_statebased.c.
Test infrastructure only (see oracle/__init__.py).
"""
import numpy as np

from .geo import kwikqdrdist_matrix

nm = 1852.0
ft = 0.3048
RPZ_DEFAULT = 5.0 * nm        # settings.asas_pzr [NM]
HPZ_DEFAULT = 1000.0 * ft     # settings.asas_pzh [ft]
DTLOOK_DEFAULT = 300.0        # settings.asas_dtlookahead [s]


def _as_vec(x, n):
    x = np.asarray(x, dtype=np.float64)
    return np.full(n, float(x)) if x.ndim == 0 else x


def detect_rows(rows, lat, lon, trk, gs, alt, vs, rpz=RPZ_DEFAULT, hpz=HPZ_DEFAULT,
                dtlookahead=DTLOOK_DEFAULT, with_margins=False):
    """CPA quantities of ordered pairs (i, j) for i in ``rows`` and every j.

    Returns a dict of (len(rows), N) matrices: swconfl, swlos, tcpa, dcpa2, tinconf, toutconf, dist,
    qdr, and -- when ``with_margins`` -- ``near``: True where any predicate operand sits within the
    comparison band of its threshold, i.e. where a float32 kernel may
    legitimately decide the other way.
    """
    lat = np.asarray(lat, dtype=np.float64)
    lon = np.asarray(lon, dtype=np.float64)
    trk = np.asarray(trk, dtype=np.float64)
    gs = np.asarray(gs, dtype=np.float64)
    alt = np.asarray(alt, dtype=np.float64)
    vs = np.asarray(vs, dtype=np.float64)
    n = lat.shape[0]
    rows = np.asarray(rows, dtype=np.int64)
    rpz = _as_vec(rpz, n)
    hpz = _as_vec(hpz, n)
    dtl = _as_vec(dtlookahead, n)

    eye = (rows.reshape(-1, 1) == np.arange(n).reshape(1, -1))
    I = eye.astype(np.float64)

    qdr, dist = kwikqdrdist_matrix(lat[rows], lon[rows], lat, lon)
    dist = dist * nm + 1e9 * I
    qdrrad = np.radians(qdr)
    dx = dist * np.sin(qdrrad)
    dy = dist * np.cos(qdrrad)

    trkrad = np.radians(trk)
    u = gs * np.sin(trkrad)
    v = gs * np.cos(trkrad)
    du = u.reshape(1, -1) - u[rows].reshape(-1, 1)      # velocity of j relative to i
    dv = v.reshape(1, -1) - v[rows].reshape(-1, 1)

    dv2 = du * du + dv * dv
    dv2 = np.where(np.abs(dv2) < 1e-6, 1e-6, dv2)
    vrel = np.sqrt(dv2)

    tcpa = -(du * dx + dv * dy) / dv2 + 1e9 * I
    dcpa2 = np.abs(dist * dist - tcpa * tcpa * dv2)

    rpzm = np.maximum(rpz[rows].reshape(-1, 1), rpz.reshape(1, -1))
    R2 = rpzm * rpzm
    swhorconf = dcpa2 < R2
    dxinhor = np.sqrt(np.maximum(0.0, R2 - dcpa2))
    dtinhor = dxinhor / vrel
    tinhor = np.where(swhorconf, tcpa - dtinhor, 1e8)
    touthor = np.where(swhorconf, tcpa + dtinhor, -1e8)

    dalt = alt.reshape(1, -1) - alt[rows].reshape(-1, 1) + 1e9 * I
    dvs = vs.reshape(1, -1) - vs[rows].reshape(-1, 1)
    dvs = np.where(np.abs(dvs) < 1e-6, 1e-6, dvs)
    hpzm = np.maximum(hpz[rows].reshape(-1, 1), hpz.reshape(1, -1))
    tcrosshi = (dalt + hpzm) / -dvs
    tcrosslo = (dalt - hpzm) / -dvs
    tinver = np.minimum(tcrosshi, tcrosslo)
    toutver = np.maximum(tcrosshi, tcrosslo)

    tinconf = np.maximum(tinver, tinhor)
    toutconf = np.minimum(toutver, touthor)
    dtl_i = dtl[rows].reshape(-1, 1)
    swconfl = swhorconf & (tinconf <= toutconf) & (toutconf > 0.0) & (tinconf < dtl_i) & (~eye)
    swlos = (dist < rpzm) & (np.abs(dalt) < hpzm)

    out = dict(swconfl=swconfl, swlos=swlos, tcpa=tcpa, dcpa2=dcpa2, tinconf=tinconf,
               toutconf=toutconf, dist=dist, qdr=qdr)
    if with_margins:
        rel = 1e-4
        tt = 1e-2
        near_conf = (np.abs(dcpa2 - R2) / R2 < rel) | \
                    (swhorconf & ((np.abs(tinconf - toutconf) < tt) | (np.abs(toutconf) < tt) |
                                  (np.abs(tinconf - dtl_i) < tt)))
        # the vertical window collapses to a knife edge when |dalt| ~ hpz
        near_vert = np.abs(np.abs(dalt) - hpzm) < 0.05
        near_los = (np.abs(dist - rpzm) / rpzm < rel) | near_vert
        out["near_conf"] = (near_conf | (swhorconf & near_vert)) & (~eye)
        out["near_los"] = near_los & (~eye)
    return out


def detect(lat, lon, trk, gs, alt, vs, rpz=RPZ_DEFAULT, hpz=HPZ_DEFAULT, dtlookahead=DTLOOK_DEFAULT):
    """Full N x N detection.  Returns the tuple upstream's ``detect`` returns, with index pairs.

    (confpairs, lospairs, inconf, tcpamax, qdr, dist, dcpa, tcpa, tinconf); pair lists are row-major
    ``(i, j)`` index tuples (upstream maps them to aircraft ids), both directions present.
    """
    n = len(lat)
    m = detect_rows(np.arange(n), lat, lon, trk, gs, alt, vs, rpz, hpz, dtlookahead)
    sw = m["swconfl"]
    inconf = np.any(sw, axis=1)
    tcpamax = np.max(m["tcpa"] * sw, axis=1) if n else np.zeros(0)
    ci, cj = np.where(sw)
    li, lj = np.where(m["swlos"])
    confpairs = list(zip(ci.tolist(), cj.tolist()))
    lospairs = list(zip(li.tolist(), lj.tolist()))
    return (confpairs, lospairs, inconf, tcpamax, m["qdr"][sw], m["dist"][sw],
            np.sqrt(m["dcpa2"][sw]), m["tcpa"][sw], m["tinconf"][sw])
