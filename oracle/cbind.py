"""ctypes binding of oracle/libcd_oracle.so (the C restatement).  Test infrastructure only."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libcd_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.cd_detect_rows.restype = ctypes.c_int
        _LIB.cd_pair_eval.restype = ctypes.c_int
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t)) if a is not None else None


def _vec(x, n):
    x = np.asarray(x, dtype=np.float64)
    return np.ascontiguousarray(np.full(n, float(x)) if x.ndim == 0 else x)


def detect_rows(lat, lon, trk, gs, alt, vs, rpz, hpz, dtlookahead, row0=0, nrows=None, pair_cap=0,
                nthreads=None):
    """Rows [row0, row0+nrows) x all columns.  Returns dict(inconf, tcpamax, nconf_row, nlos_row,
    n_conf, n_los[, confpairs, lospairs])."""
    arrs = [np.ascontiguousarray(np.asarray(a, dtype=np.float64)) for a in (lat, lon, trk, gs, alt, vs)]
    n = arrs[0].shape[0]
    nthreads = nthreads or (os.cpu_count() or 1)
    nrows = n - row0 if nrows is None else nrows
    rpz, hpz, dtl = _vec(rpz, n), _vec(hpz, n), _vec(dtlookahead, n)
    inconf = np.zeros(nrows, dtype=np.uint8)
    tcpamax = np.zeros(nrows, dtype=np.float64)
    ncr = np.zeros(nrows, dtype=np.uint32)
    nlr = np.zeros(nrows, dtype=np.uint32)
    totals = np.zeros(2, dtype=np.uint64)
    pc = np.zeros((pair_cap, 2), dtype=np.int32) if pair_cap else None
    pl = np.zeros((pair_cap, 2), dtype=np.int32) if pair_cap else None
    d = ctypes.c_double
    rc = lib().cd_detect_rows(*[_p(a, d) for a in arrs], _p(rpz, d), _p(hpz, d), _p(dtl, d),
                              ctypes.c_long(n), ctypes.c_long(row0), ctypes.c_long(nrows),
                              _p(inconf, ctypes.c_uint8), _p(tcpamax, d), _p(ncr, ctypes.c_uint32),
                              _p(nlr, ctypes.c_uint32), _p(pc, ctypes.c_int32), _p(pl, ctypes.c_int32),
                              ctypes.c_long(pair_cap), _p(totals, ctypes.c_uint64), ctypes.c_int(nthreads))
    if rc != 0:
        raise MemoryError("cd_detect_rows failed")
    out = dict(inconf=inconf.astype(bool), tcpamax=tcpamax, nconf_row=ncr, nlos_row=nlr,
               n_conf=int(totals[0]), n_los=int(totals[1]))
    if pair_cap:
        out["confpairs"] = pc[:min(pair_cap, int(totals[0]))]
        out["lospairs"] = pl[:min(pair_cap, int(totals[1]))]
    return out
