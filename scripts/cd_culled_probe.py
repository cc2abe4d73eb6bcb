"""The detection forms at N = 100k on strip-sorted records, for ncu captures: default = three culled launches
(`<0,1,0>`); with the argument `forms` one launch each of culled, culled + symmetric (`<0,1,1>`, the default of
StateBasedCD.detect) and every ordered pair (`<0,0,0>`):

    ncu --set full -k regex:cd_tiled_kernel -c 3 -o out python scripts/cd_culled_probe.py 100000 forms
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from bluesky_gym_sasha_b200.cd import StateBasedCD

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
rng = np.random.default_rng(1)
lat, lon = 52 + 40 * (rng.random(n) - 0.5), 4 + 40 * (rng.random(n) - 0.5)
alt = np.round(rng.uniform(3000, 12000, n) / 304.8) * 304.8
gs, trk = rng.uniform(150, 250, n), rng.uniform(0, 360, n)
vs = np.where(rng.random(n) < 0.8, 0.0, rng.choice([-1.0, 1.0], n) * rng.uniform(5, 15, n))
cd = StateBasedCD(device=0)
d = [cd._as_dev(x) for x in (lat, lon, trk, gs, alt, vs)]
perm = cd.spatial_order(d[0], d[1])
rec, _ = cd.pack(*[x[perm] for x in d], 52.0, 4.0)
if len(sys.argv) > 2 and sys.argv[2] == "forms":
    out = cd.detect_packed(rec, n, cull=True)
    out = cd.detect_packed(rec, n, cull=True, symmetric=True)
    out = cd.detect_packed(rec, n, cull=False)
else:
    for _ in range(3):
        out = cd.detect_packed(rec, n, cull=True)
torch.cuda.synchronize()
print("conflicts", int(out["npairs"][0]))
