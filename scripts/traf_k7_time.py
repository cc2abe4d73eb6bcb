import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from bluesky_gym_sasha_b200.traffic import AirspaceTraffic
for n in (1 << 20, 100_000):
    rng = np.random.default_rng(0)
    lat, lon = 30 + 100.0 * (rng.random(n) - 0.5), 4 + 120.0 * (rng.random(n) - 0.5)
    hdg = rng.uniform(0, 360, n); alt = np.round(rng.uniform(3000, 12000, n) / 304.8) * 304.8
    tr = AirspaceTraffic(n, simdt=1.0, reso=None, max_wpts=4)
    tr.create(lat, lon, hdg, alt, rng.uniform(120, 150, n))
    d = np.array([0.5, 1.0, 1.5, 2.0])[None, :]
    tr.set_routes(np.arange(n), lat[:, None] + d * np.cos(np.radians(hdg))[:, None], lon[:, None] + d * np.sin(np.radians(hdg))[:, None] / np.cos(np.radians(lat))[:, None])
    tr.step(5, detect=False); torch.cuda.synchronize()
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
    ts = []
    for i in range(30):
        flush.fill_(float(i))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); tr.step(1, detect=False); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print(os.environ.get("BSG_B200_LIB", "product"), n, "K7 median %.1f us (memset + kernel)" % np.median(ts))
