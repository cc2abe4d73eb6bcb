"""A/B of libraries (BSG_B200_LIB) on the culled N = 100k detection."""
import os
import subprocess
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, os, numpy as np, torch
sys.path.insert(0, %r)
from bluesky_gym_sasha_b200.cd import StateBasedCD
n = 100000
rng = np.random.default_rng(1)
lat, lon = 52 + 40 * (rng.random(n) - 0.5), 4 + 40 * (rng.random(n) - 0.5)
alt = np.round(rng.uniform(3000, 12000, n) / 304.8) * 304.8
gs, trk = rng.uniform(150, 250, n), rng.uniform(0, 360, n)
vs = np.where(rng.random(n) < 0.8, 0.0, rng.choice([-1.0, 1.0], n) * rng.uniform(5, 15, n))
cd = StateBasedCD(device=0)
d = [cd._as_dev(x) for x in (lat, lon, trk, gs, alt, vs)]
perm = cd.spatial_order(d[0], d[1])
rec, _ = cd.pack(*[x[perm] for x in d], 52.0, 4.0)
for cull in (True, False):
    for _ in range(3): cd.detect_packed(rec, n, cull=cull)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); o = cd.detect_packed(rec, n, cull=cull); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(os.path.basename(os.environ.get("BSG_B200_LIB", "default")), "culled" if cull else "plain ", "%%.3f ms" %% best, int(o["npairs"][0]))
''' % root
for lib in sys.argv[1:]:
    subprocess.run([sys.executable, "-c", code], env=dict(os.environ, BSG_B200_LIB=os.path.join(root, lib)))
