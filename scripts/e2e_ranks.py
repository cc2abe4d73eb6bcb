"""End-to-end BlueSkyVectorEnv.step() under torchrun, one rank per GPU, and what limits it when the ranks share a host:
per rank the device->host transfer of one output block alone on the node and with every rank transferring at once (GPUs
behind a shared PCIe switch uplink halve each other's rate), the host cost of an asynchronous step_torch(), and the full
step() -- whole-node env-steps/s = max over ranks.

    torchrun --nproc-per-node 8 scripts/e2e_ranks.py
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from bluesky_gym_sasha_b200.vector_env import BlueSkyVectorEnv

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("gloo")
E, K = 4096, 300
v = BlueSkyVectorEnv("HorizontalCREnv-v0", E, device=local, seed=0, env_id_offset=rank * E, cd_enabled=True, n_intruders=20,
                     autoreset_mode="same_step")
v.reset()
a = np.random.default_rng(rank).uniform(-1, 1, (K, E, 1)).astype(np.float32)
at = torch.from_numpy(a[0]).cuda()
for i in range(100):
    v.step_torch(at)


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


d = torch.empty((v._out_bytes,), dtype=torch.uint8, device="cuda")
hp = torch.empty((v._out_bytes,), dtype=torch.uint8).pin_memory()


def d2h_us(n=200):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        hp.copy_(d, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


d2h_us(20)
alone = 0.0
for r in range(world):                          # one rank at a time
    barrier()
    if r == rank:
        alone = d2h_us()
barrier()
together = d2h_us()                             # every rank at once
barrier()
t0 = time.perf_counter()
for i in range(K):
    v.step_torch(at)
host_torch = (time.perf_counter() - t0) / K * 1e6
barrier()
for i in range(20):
    v.step(a[i])
barrier()
t0 = time.perf_counter()
for i in range(K):
    v.step(a[i])
full = (time.perf_counter() - t0) / K * 1e6
barrier()
row = torch.tensor([alone, together, host_torch, full], dtype=torch.float64)
rows = [torch.zeros(4, dtype=torch.float64) for _ in range(world)]
if world > 1:
    dist.all_gather(rows, row)
else:
    rows = [row]
if rank == 0:
    print(f"world {world}, {os.cpu_count()} host cores, affinity of rank 0: {len(os.sched_getaffinity(0))} cores")
    print("rank   D2H alone   D2H all ranks   step_torch host   step() full    [us]")
    for r, x in enumerate(rows):
        print(f"{r:4d}   {x[0]:9.1f}   {x[1]:13.1f}   {x[2]:15.1f}   {x[3]:11.1f}")
    worst = max(float(x[3]) for x in rows)
    print(f"whole node: {E * world / worst * 1e6:.3e} env-steps/s ({worst:.1f} us/step on the slowest rank)", flush=True)
    os.system("nvidia-smi topo -m 2>/dev/null | head -14")
v.close()
if world > 1:
    dist.destroy_process_group()
