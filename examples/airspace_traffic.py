"""One airspace of N aircraft on routes with VNAV and MVP conflict resolution: the upstream calls a scenario would make
(cre, ADDWPT, LNAV / VNAV ON, RESO MVP, bs.sim.step) on the device.

    python examples/airspace_traffic.py [N] [SECONDS]
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bluesky_gym_sasha_b200.traffic import AirspaceTraffic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000
seconds = int(sys.argv[2]) if len(sys.argv) > 2 else 120
rng = np.random.default_rng(0)

traf = AirspaceTraffic(n_max=n, simdt=1.0, reso="MVP")
lat, lon = 52 + 30 * (rng.random(n) - 0.5), 4 + 30 * (rng.random(n) - 0.5)
hdg = rng.uniform(0, 360, n)
alt = np.round(rng.uniform(3000, 11000, n) / 304.8) * 304.8
idx = traf.create(lat, lon, hdg, alt, rng.uniform(120, 160, n))              # bs.traf.cre
# three waypoints ahead of every aircraft, the last one 600 m lower at a slower speed (altitude / speed constraints)
d = np.array([0.5, 1.0, 1.5])[None, :]
wlat = lat[:, None] + d * np.cos(np.radians(hdg))[:, None]
wlon = lon[:, None] + d * np.sin(np.radians(hdg))[:, None] / np.cos(np.radians(lat))[:, None]
walt = np.stack([np.full(n, -999.0), np.full(n, -999.0), alt - 600.0], axis=1)
wspd = np.stack([np.full(n, -999.0), np.full(n, -999.0), np.full(n, 110.0)], axis=1)
traf.set_routes(np.arange(n), wlat, wlon, alt=walt, spd=wspd)               # ADDWPT x 3, LNAV / VNAV ON

traf.step(5)
torch.cuda.synchronize()
t0 = time.perf_counter()
traf.step(seconds)                                                           # `seconds` x bs.sim.step()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
c = traf.conflicts()
print(f"{n} aircraft, {seconds} s simulated in {dt * 1e3:.1f} ms ({seconds / dt:.0f} x real time); last substep: "
      f"{c['n_conf']} conflict pairs, {c['n_los']} LoS pairs, {int(traf.asas_active.sum())} aircraft following a resolution; "
      f"counters {traf.counters()}")
