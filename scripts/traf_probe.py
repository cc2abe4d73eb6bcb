"""A few AirspaceTraffic substeps at N aircraft (for ncu captures of traf_substep_kernel).  usage: traf_probe.py [N] [reso 0|1]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from bluesky_gym_sasha_b200.traffic import AirspaceTraffic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
reso = len(sys.argv) > 2 and sys.argv[2] == "1"
rng = np.random.default_rng(0)
box = 40.0 if n <= 200_000 else 120.0
lat, lon = 30 + min(box, 100.0) * (rng.random(n) - 0.5), 4 + box * (rng.random(n) - 0.5)
hdg = rng.uniform(0, 360, n)
alt = np.round(rng.uniform(3000, 12000, n) / 304.8) * 304.8
tr = AirspaceTraffic(n, simdt=1.0, reso="MVP" if reso else None, max_wpts=4)
tr.create(lat, lon, hdg, alt, rng.uniform(120, 150, n))
d = np.array([0.5, 1.0, 1.5, 2.0])[None, :]
tr.set_routes(np.arange(n), lat[:, None] + d * np.cos(np.radians(hdg))[:, None],
              lon[:, None] + d * np.sin(np.radians(hdg))[:, None] / np.cos(np.radians(lat))[:, None])
tr.step(6, detect=reso)
torch.cuda.synchronize()
print("done", tr.counters())
