"""One airspace sharded over the ranks (AirspaceTraffic(group=True)) against the same airspace on one GPU: every rank
builds the WHOLE scenario from the same seed, rank 0's GPU also runs it unsharded, the ranks run their blocks, and after
every few substeps the sharded state (gathered) is compared with the unsharded one.

    torchrun --nproc-per-node 2 scripts/traf_sharded_check.py [N] [SUBSTEPS]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from bluesky_gym_sasha_b200.traffic import AirspaceTraffic

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
per = -(-n // world)
per = -(-per // 256) * 256                                  # whole tiles per rank; the last rank may hold fewer aircraft
rng = np.random.default_rng(11)
box = 6.0 if n <= 20000 else 30.0
lat, lon = 52 + box * (rng.random(n) - 0.5), 4 + box * (rng.random(n) - 0.5)
hdg = rng.uniform(0, 360, n)
alt = np.round(rng.uniform(3000, 9000, n) / 304.8) * 304.8 + rng.uniform(-30, 30, n)
spd = rng.uniform(120, 160, n)
d = np.array([0.4, 0.8, 1.2])[None, :]
wlat = lat[:, None] + d * np.cos(np.radians(hdg))[:, None]
wlon = lon[:, None] + d * np.sin(np.radians(hdg))[:, None] / np.cos(np.radians(lat))[:, None]
walt = np.stack([np.full(n, -999.0), np.full(n, -999.0), alt - 500.0], axis=1)
wspd = np.stack([np.full(n, -999.0), np.full(n, -999.0), np.full(n, 115.0)], axis=1)


def build(sl, **kw):
    t = AirspaceTraffic(n_max=max(sl.stop - sl.start, 1) if kw else n, device=local, simdt=1.0, reso="MVP", **kw)
    k = sl.stop - sl.start
    if k > 0:
        t.create(lat[sl], lon[sl], hdg[sl], alt[sl], spd[sl])
        t.set_routes(np.arange(k), wlat[sl], wlon[sl], alt=walt[sl], spd=wspd[sl])
    return t


lo, hi = min(rank * per, n), min((rank + 1) * per, n)
shard = AirspaceTraffic(n_max=per, device=local, simdt=1.0, reso="MVP", group=True)
if hi > lo:
    sl = slice(lo, hi)
    shard.create(lat[sl], lon[sl], hdg[sl], alt[sl], spd[sl])
    shard.set_routes(np.arange(hi - lo), wlat[sl], wlon[sl], alt=walt[sl], spd=wspd[sl])
whole = build(slice(0, n)) if rank == 0 else None

worst = {}
n_conf_seen = 0
for s in range(0, steps, 5):
    shard.step(5)
    if whole is not None:
        whole.step(5)
    torch.cuda.synchronize()
    state = torch.zeros((per, 8), dtype=torch.float64, device="cuda")
    k = hi - lo
    if k > 0:
        state[:k, 0:2] = shard.t["pos"][:k]
        state[:k, 2:6] = shard.t["kin"][:k].double()
        state[:k, 6] = (shard.t["flags"][:k] & 32).double()            # BSG_TF_ASAS
        state[:k, 7] = shard.t["asas"][:k, 0].double()
    allst = torch.zeros((per * world, 8), dtype=torch.float64, device="cuda")
    dist.all_gather_into_tensor(allst, state)
    if rank == 0:
        a = allst[:n].cpu().numpy()
        w = whole.t
        b = np.concatenate([w["pos"][:n].cpu().numpy(), w["kin"][:n].double().cpu().numpy(),
                            (w["flags"][:n] & 32).double().cpu().numpy()[:, None], w["asas"][:n, 0].double().cpu().numpy()[:, None]], axis=1)
        for c, name in enumerate(("lat", "lon", "alt", "tas", "hdg", "vs", "asas flag")):
            dlt = np.abs(a[:, c] - b[:, c])
            if name == "hdg":
                dlt = np.minimum(dlt, 360.0 - dlt)
            worst[name] = max(worst.get(name, 0.0), float(dlt.max()))
        n_conf_seen = max(n_conf_seen, int(whole.last["npairs"][0]))
ok = True
if rank == 0:
    print(f"world {world}, N = {n}, {steps} substeps with MVP + VNAV routes: largest |sharded - unsharded| {worst}; "
          f"up to {n_conf_seen} conflict pairs per substep, {int((whole.t['flags'][:n] & 32).ne(0).sum())} aircraft under ASAS at the end")
    ok = worst["asas flag"] == 0.0 and worst["lat"] < 1e-9 and worst["lon"] < 1e-9 and worst["alt"] < 1e-3 and worst["tas"] < 1e-4 \
        and worst["hdg"] < 1e-3 and worst["vs"] < 1e-4 and n_conf_seen > 0
    # timing of the sharded substep (max over ranks below)
dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
shard.step(20)
e1.record()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / 20.0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"sharded substep: {t.item():.3f} ms (max over ranks)")
    print("SHARDED_TRAFFIC_OK" if ok else "SHARDED_TRAFFIC_MISMATCH")
dist.destroy_process_group()
sys.exit(0 if ok else 1)
