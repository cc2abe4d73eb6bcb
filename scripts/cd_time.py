"""Device time of the single-airspace CD forms at N = 100k (BASELINE configs[4] on one GPU): every ordered pair, symmetric,
culled, culled + symmetric.  Works with any build of the package on PYTHONPATH (A/B of library versions):

    python scripts/cd_time.py                       # this checkout
    PYTHONPATH=ab_libs/r1 python scripts/cd_time.py # a saved copy of another build
"""
import os
import sys

import numpy as np
import torch

if not os.environ.get("PYTHONPATH"):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bluesky_gym_sasha_b200.cd import StateBasedCD  # noqa: E402


def main(n=100_000, reps=7):
    rng = np.random.default_rng(1)
    lat = 52 + 40 * (rng.random(n) - 0.5)
    lon = 4 + 40 * (rng.random(n) - 0.5)
    alt = np.round(rng.uniform(3000, 12000, n) / 304.8) * 304.8
    gs = rng.uniform(150, 250, n)
    trk = rng.uniform(0, 360, n)
    vs = np.where(rng.random(n) < 0.8, 0.0, rng.choice([-1.0, 1.0], n) * rng.uniform(5, 15, n))
    cd = StateBasedCD(device=0)
    d = [cd._as_dev(x) for x in (lat, lon, trk, gs, alt, vs)]
    perm = cd.spatial_order(d[0], d[1])
    rec, _ = cd.pack(*d, 52.0, 4.0)
    rec_s, _ = cd.pack(*[x[perm] for x in d], 52.0, 4.0)
    import bluesky_gym_sasha_b200
    print("package:", os.path.dirname(bluesky_gym_sasha_b200.__file__))
    for name, r, kw in (("every ordered pair", rec, {}), ("symmetric", rec_s, dict(symmetric=True)),
                        ("culled", rec_s, dict(cull=True)), ("culled + symmetric", rec_s, dict(cull=True, symmetric=True)),
                        ("every ordered pair, no lists", rec, dict(want_pairs=False)),
                        ("row shard 1/8 (12544 rows x all)", rec, dict(row0=12544 * 3, n_rows=12544)),
                        ("row shard 1/8, culled", rec_s, dict(row0=12544 * 3, n_rows=12544, cull=True))):
        for _ in range(2):
            out = cd.detect_packed(r, n, **kw)
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = cd.detect_packed(r, n, **kw)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(f"{name:32s} best {min(ts):8.3f} ms  median {sorted(ts)[len(ts) // 2]:8.3f} ms   conflicts {int(out['npairs'][0])}  LoS {int(out['npairs'][1])}")


if __name__ == "__main__":
    main()
