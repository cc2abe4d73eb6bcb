"""ISA atmosphere and airspeed conversions, float64 NumPy.

[UPSTREAM-RECALL] restates ``bluesky/tools/aero.py`` (vectorised ``v*`` functions).  Reached from the
reference through every ``bs.traf.cre`` (e.g. horizontal_cr_env.py:91, descent_env.py:173) and every
``bs.sim.step()`` (e.g. horizontal_cr_env.py:109); ``kts`` is imported at static_obstacle_env.py:7.
Test infrastructure only (see oracle/__init__.py).
"""
import numpy as np

kts = 0.514444
ft = 0.3048
fpm = ft / 60.0
nm = 1852.0
g0 = 9.80665
R = 287.05287
p0 = 101325.0
rho0 = 1.225
T0 = 288.15
Tstrat = 216.65
gamma = 1.40
beta = -0.0065
Rearth = 6371000.0


def vtemp(h):
    return np.maximum(T0 + beta * h, Tstrat)


def vatmos(h):
    """(p, rho, T) of the ISA at geometric altitude h [m]."""
    T = vtemp(h)
    rhotrop = rho0 * (T / T0) ** 4.256848030018761
    dhstrat = np.maximum(0.0, h - 11000.0)
    rho = rhotrop * np.exp(-dhstrat / 6341.552161)
    p = rho * R * T
    return p, rho, T


def vvsound(h):
    return np.sqrt(gamma * R * vtemp(h))


def vtas2mach(tas, h):
    return tas / vvsound(h)


def vmach2tas(M, h):
    return M * vvsound(h)


def vtas2cas(tas, h):
    p, rho, _ = vatmos(h)
    qdyn = p * ((1.0 + rho * tas * tas / (7.0 * p)) ** 3.5 - 1.0)
    cas = np.sqrt(7.0 * p0 / rho0 * ((qdyn / p0 + 1.0) ** (2.0 / 7.0) - 1.0))
    return np.where(tas < 0, -cas, cas)


def vcas2tas(cas, h):
    p, rho, _ = vatmos(h)
    qdyn = p0 * ((1.0 + rho0 * cas * cas / (7.0 * p0)) ** 3.5 - 1.0)
    tas = np.sqrt(7.0 * p / rho * ((qdyn / p + 1.0) ** (2.0 / 7.0) - 1.0))
    return np.where(cas < 0, -tas, tas)


def vcasormach(spd, h):
    """(tas, cas, M) from a speed that is Mach when 0.1 < spd < 1, else CAS [m/s]."""
    spd = np.asarray(spd, dtype=np.float64)
    ismach = np.logical_and(0.1 < spd, spd < 1.0)
    tas = np.where(ismach, vmach2tas(spd, h), vcas2tas(spd, h))
    cas = np.where(ismach, vtas2cas(tas, h), spd)
    M = np.where(ismach, spd, vtas2mach(tas, h))
    return tas, cas, M


def vcasormach2tas(spd, h):
    spd = np.asarray(spd, dtype=np.float64)
    return np.where(np.abs(spd) < 1.0, vmach2tas(spd, h), vcas2tas(spd, h))
