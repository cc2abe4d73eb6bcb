"""OpenAP-lite performance envelope (flight phase, speed/VS/altitude limits, axmax), float64 NumPy.

[UPSTREAM-RECALL] restates the parts of ``bluesky/traffic/performance/openap/{perfoap,phase,coeff}.py``
that change the aircraft state under the reference's settings (default ``performance_model='openap'``;
every ``bs.traf.cre`` in the reference uses actype "A320", e.g. horizontal_cr_env.py:91).  Drag, thrust
and fuel only feed outputs the reference never reads, so they are not restated; ``axmax`` uses the
fixed-value rule (2 m/s^2 on the ground, ``axmax_air`` airborne).

The A320 envelope below is DATA, not code: OpenAP's WRAP table is not available in this image, so the
numbers are recalled / derived from the reference's shipped training logs (SURVEY.md section 8c:
HorizontalCREnv straight-flight episodes imply a ground-phase CAS cap of about 88-89 m/s, DescentEnv
episodes imply CAS 150 m/s is *not* clamped airborne).  They are overridable per field, and the same
struct is what the CUDA side receives (``bsg_perf`` in include/bsg.h), so oracle and kernels always
agree on them.  Test infrastructure only (see oracle/__init__.py).
"""
from dataclasses import dataclass, asdict

import numpy as np

from . import aero

# bluesky.traffic.performance.openap.phase constants
PH_NA, PH_TO, PH_IC, PH_CL, PH_CR, PH_DE, PH_AP, PH_LD, PH_GD = 0, 1, 2, 3, 4, 5, 6, 7, 8


@dataclass(frozen=True)
class PerfTable:
    vminto: float = 73.3      # to_v_lof min          [m/s CAS]
    vmaxic: float = 88.5      # ic_va_avg max (log-derived 87.3..90.0)
    vminer: float = 64.0      # min over ic/cl/cr/de/fa tables
    vmaxer: float = 163.0     # max over ic/cl/cr/de/fa tables (>= 150: Descent/Sector fly CAS 150)
    vminap: float = 64.0      # fa_va_avg min
    vmaxap: float = 78.0      # fa_va_avg max
    vsmin: float = -20.4      # [m/s]
    vsmax: float = 18.6       # [m/s]
    hmax: float = 12500.0     # cr_h_max [m]
    mmo: float = 0.82
    axmax_gd: float = 2.0     # [m/s^2]
    axmax_air: float = 0.5    # [m/s^2] (fixed-value rule)

    def as_dict(self):
        return asdict(self)


A320 = PerfTable()


def phase_fixwing(tas, vs, alt):
    """phase.get(lifttype=FIXWING, unit='SI'): later assignments overwrite earlier ones."""
    alt_ft = np.asarray(alt, dtype=np.float64) / aero.ft
    roc = np.asarray(vs, dtype=np.float64) / aero.fpm
    ph = np.zeros(alt_ft.shape, dtype=np.int32)
    ph[alt_ft <= 75.0] = PH_GD
    ph[(alt_ft >= 75.0) & (alt_ft <= 1000.0) & (roc >= 150.0)] = PH_IC
    ph[(alt_ft >= 75.0) & (alt_ft <= 1000.0) & (roc <= -150.0)] = PH_AP
    ph[(alt_ft >= 1000.0) & (roc >= 150.0)] = PH_CL
    ph[(alt_ft >= 1000.0) & (roc <= -150.0)] = PH_DE
    ph[(alt_ft >= 10000.0) & (roc <= 150.0) & (roc >= -150.0)] = PH_CR
    return ph


def v_limits(phase, tab: PerfTable):
    """perfoap._construct_v_limits: the chain of np.where's contains the always-true
    ``(ph >= CL) | (ph <= DE)``, so NA/IC/CL/CR/DE all end on the en-route pair; AP and GD override."""
    phase = np.asarray(phase)
    vmin = np.full(phase.shape, tab.vminer, dtype=np.float64)
    vmax = np.full(phase.shape, tab.vmaxer, dtype=np.float64)
    vmin = np.where(phase == PH_AP, tab.vminap, vmin)
    vmax = np.where(phase == PH_AP, tab.vmaxap, vmax)
    vmin = np.where(phase == PH_GD, 0.0, vmin)
    vmax = np.where(phase == PH_GD, tab.vmaxic, vmax)
    return vmin, vmax


def axmax(phase, tab: PerfTable):
    return np.where(np.asarray(phase) == PH_GD, tab.axmax_gd, tab.axmax_air)


def limits(intent_tas, intent_vs, intent_h, ax, phase, cur_tas, tab: PerfTable):
    """perfoap.limits -- NB the CAS round trip is evaluated at the *allowed commanded altitude*."""
    vmin, vmax = v_limits(phase, tab)
    amax = axmax(phase, tab)
    allow_h = np.where(intent_h > tab.hmax, tab.hmax, intent_h)
    intent_cas = aero.vtas2cas(intent_tas, allow_h)
    allow_cas = np.where(intent_cas < vmin, vmin, intent_cas)
    allow_cas = np.where(intent_cas > vmax, vmax, allow_cas)
    allow_tas = aero.vcas2tas(allow_cas, allow_h)
    allow_tas = np.where(aero.vtas2mach(allow_tas, allow_h) > tab.mmo,
                         aero.vmach2tas(tab.mmo, allow_h), allow_tas)
    vs_max_with_acc = (1.0 - ax / amax) * tab.vsmax
    allow_vs = np.where((intent_vs > 0) & (intent_vs > tab.vsmax), vs_max_with_acc, intent_vs)
    allow_vs = np.where((intent_vs < 0) & (intent_vs < tab.vsmin), vs_max_with_acc, allow_vs)
    allow_vs = np.where((np.asarray(phase) == PH_GD) & (cur_tas < tab.vminto), 0.0, allow_vs)
    return allow_tas, allow_vs, allow_h
