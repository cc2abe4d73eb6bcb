"""Shared helpers of the parity tests: synthetic airspaces and oracle <-> device state transfer."""
import numpy as np

from bluesky_gym_sasha_b200 import _lib


def synth_airspace(n, box_deg=40.0, seed=0, lat0=52.0, lon0=4.0, alt_jitter=20.0):
    """uniform lat/lon box, flight-level altitudes, 80 % level flight.

    Levels are 1000 ft apart and hpz is 1000 ft, so aircraft snapped *exactly* to levels put every
    adjacent-level pair on the |dalt| == hpz knife edge, where the float64 reference itself decides by
    rounding noise.  The parity tests therefore jitter altitudes by +-alt_jitter metres (altimetry
    scatter); bench.py keeps the exact snapping (it does not change the work)."""
    rng = np.random.default_rng(seed)
    lat = lat0 + box_deg * (rng.random(n) - 0.5)
    lon = lon0 + box_deg * (rng.random(n) - 0.5)
    alt = np.round(rng.uniform(3000.0, 12000.0, n) / 304.8) * 304.8 + rng.uniform(-alt_jitter, alt_jitter, n)
    gs = rng.uniform(150.0, 250.0, n)
    trk = rng.uniform(0.0, 360.0, n)
    vs = np.where(rng.random(n) < 0.8, 0.0, rng.choice([-1.0, 1.0], n) * rng.uniform(5.0, 15.0, n))
    return lat, lon, trk, gs, alt, vs


def inject_oracle_env(venv, e, oenv, extra_f64=None, extra_i32=None, poly=None, extra_f32=None):
    """Copies an oracle env's post-reset traffic into slot e of a BlueSkyVectorEnv (bsg_load_state)."""
    t = oenv.traf
    lnav = t.swlnav.copy()
    i32 = {_lib.I32_SIMK: int(t.nstep), _lib.I32_NUM_AC: int(t.ntraf), _lib.I32_STEP: 0,
           _lib.I32_WPT_REACH: 0, _lib.I32_DRIFT_N: 0, _lib.I32_INTRUSIONS: 0, _lib.I32_NEEDS_RESET: 0,
           _lib.I32_FAF: 0}
    i32.update(extra_i32 or {})
    f32 = {_lib.F32_TOTAL_REWARD: 0.0, _lib.F32_DRIFT_SUM: 0.0, _lib.F32_FINAL_ALT: 0.0}
    f32.update(extra_f32 or {})
    venv.load_state(e, t.lat, t.lon, t.alt, t.tas, t.hdg, t.vs, t.selspd, t.selalt, t.selvs, t.ap_trk, t.cas,
                    ax=t.ax, lnav=lnav, iactwp=np.array(t.iactwp), curlegdir=t.curlegdir,
                    env_f64=extra_f64, env_i32=i32, env_f32=f32, poly=poly)


def device_traffic(venv):
    """Host copies of the device aircraft state: dict of [E, G] arrays."""
    pos = venv.t["pos"].cpu().numpy()
    kin = venv.t["kin"].cpu().numpy().astype(np.float64)
    cmd = venv.t["cmd"].cpu().numpy().astype(np.float64)
    aux = venv.t["aux"].cpu().numpy().astype(np.float64)
    fl = venv.t["flags"].cpu().numpy()
    return dict(lat=pos[..., 0], lon=pos[..., 1], alt=kin[..., 0], tas=kin[..., 1], hdg=kin[..., 2], vs=kin[..., 3],
                selspd=cmd[..., 0], selalt=cmd[..., 1], selvs=cmd[..., 2], ap_trk=cmd[..., 3],
                ax=aux[..., 0], curlegdir=aux[..., 1], cas=aux[..., 2], flags=fl,
                inconf=venv.t["inconf"].cpu().numpy().astype(bool), tcpamax=venv.t["tcpamax"].cpu().numpy())


def angdiff(a, b):
    return np.abs((np.asarray(a) - np.asarray(b) + 180.0) % 360.0 - 180.0)
