"""Where does AirspaceTraffic first leave the oracle?  python scripts/traf_debug.py routes|mvp SEED_OR_MODE [aircraft]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.test_traffic_ext import route_scenario, conflict_scenario, make_oracle, make_device, TOL, angdiff
kind, arg = sys.argv[1], int(sys.argv[2])
watch = int(sys.argv[3]) if len(sys.argv) > 3 else None
if kind == "routes":
    sc = route_scenario(48, arg); t = make_oracle(sc); g = make_device(sc); steps = 900
else:
    sc = conflict_scenario(12, 5); t = make_oracle(sc, reso="MVP", reso_mode=arg, vnav=False)
    g = make_device(sc, reso="MVP", reso_mode=arg, vnav=False, cull=False, symmetric=False); steps = 400
n = t.ntraf
seen = set()
def dev():
    torch.cuda.synchronize()
    f = g.t["flags"][:n].cpu().numpy()
    return dict(lat=g.lat.cpu().numpy(), lon=g.lon.cpu().numpy(), alt=g.altitude.cpu().numpy().astype(float), tas=g.tas.cpu().numpy().astype(float),
                hdg=g.heading.cpu().numpy().astype(float), vs=g.vs.cpu().numpy().astype(float), selspd=g.selspd.cpu().numpy().astype(float),
                selalt=g.selalt.cpu().numpy().astype(float), aptrk=g.ap_trk.cpu().numpy().astype(float),
                vn1=g.t["vnav1"][:n].cpu().numpy().astype(float), vn2=g.t["vnav2"][:n].cpu().numpy().astype(float), aux=g.t["aux"][:n].cpu().numpy().astype(float),
                asas=g.t["asas"][:n].cpu().numpy().astype(float), fl=f, iw=(f >> 16) & 0xff)
def show(i, d, step):
    print(f"  step {step} ac {i}: ORACLE lat {t.lat[i]:.6f} lon {t.lon[i]:.6f} alt {t.alt[i]:.2f} tas {t.tas[i]:.3f} hdg {t.hdg[i]:.3f} vs {t.vs[i]:.3f} selspd {t.selspd[i]:.2f} selalt {t.selalt[i]:.1f} aptrk {t.ap_trk[i]:.3f} "
          f"iwp {t.iactwp[i]} lnav {int(t.swlnav[i])} vnav {int(t.swvnav[i])} nextaltco {t.nextaltco[i]:.1f} xtoalt {t.xtoalt[i]:.0f} actwp_vs {t.actwp_vs[i]:.3f} dist2vs {t.dist2vs[i]:.0f} turndist {t.turndist[i]:.0f} "
          f"spd {t.actwp_spd[i]:.1f} nextspd {t.nextspd[i]:.1f} spdcon {t.spdcon[i]:.1f} vnavvs {t.vnavvs[i]:.3f} act {int(t.asas_active[i])} asas {t.asas_trk[i]:.2f} {t.asas_tas[i]:.2f} {t.asas_vs[i]:.3f} {t.asas_alt[i]:.1f}")
    print(f"               DEVICE lat {d['lat'][i]:.6f} lon {d['lon'][i]:.6f} alt {d['alt'][i]:.2f} tas {d['tas'][i]:.3f} hdg {d['hdg'][i]:.3f} vs {d['vs'][i]:.3f} selspd {d['selspd'][i]:.2f} selalt {d['selalt'][i]:.1f} aptrk {d['aptrk'][i]:.3f} "
          f"iwp {d['iw'][i]} lnav {int(d['fl'][i] & 2 > 0)} vnav {int(d['fl'][i] & 4 > 0)} nextaltco {d['vn1'][i,0]:.1f} xtoalt {d['vn1'][i,1]:.0f} actwp_vs {d['vn1'][i,2]:.3f} dist2vs {d['vn1'][i,3]:.0f} turndist {d['aux'][i,3]:.0f} "
          f"spd {d['vn2'][i,0]:.1f} nextspd {d['vn2'][i,1]:.1f} spdcon {d['vn2'][i,2]:.1f} vnavvs {d['vn2'][i,3]:.3f} act {int(d['fl'][i] & 32 > 0)} asas {d['asas'][i,0]:.2f} {d['asas'][i,1]:.2f} {d['asas'][i,2]:.3f} {d['asas'][i,3]:.1f}")
for step in range(steps):
    t.simstep(); g.step(1)
    d = dev()
    err = dict(lat=np.abs(d["lat"] - t.lat), lon=np.abs(d["lon"] - t.lon), alt=np.abs(d["alt"] - t.alt), tas=np.abs(d["tas"] - t.tas),
               hdg=angdiff(d["hdg"], t.hdg), vs=np.abs(d["vs"] - t.vs))
    if watch is not None:
        if step % 10 == 9 or max(e[watch] / TOL[k] for k, e in err.items()) > 1: show(watch, d, step)
        continue
    for i in range(n):
        if i in seen: continue
        bad = [k for k, e in err.items() if e[i] > TOL[k] * (10 if kind == "mvp" else 1)]
        if bad:
            seen.add(i)
            print(f"step {step}: aircraft {i} leaves the tight tolerance in {bad}: " + ", ".join(f"{k} {err[k][i]:.3g}" for k in bad))
            show(i, d, step)
print("aircraft that left:", sorted(seen))
