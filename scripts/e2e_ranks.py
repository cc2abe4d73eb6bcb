"""End-to-end BlueSkyVectorEnv.step() under torchrun, one rank per GPU: whole-node env-steps/s (max over ranks) for the
host-copy pool size in BSG_HOST_THREADS (unset = the library's share-of-the-cores default).

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
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
E, K = 4096, 300
v = BlueSkyVectorEnv("HorizontalCREnv-v0", E, device=local, seed=0, env_id_offset=rank * E, cd_enabled=True, n_intruders=20,
                     autoreset_mode="same_step")
v.reset()
a = np.random.default_rng(rank).uniform(-1, 1, (K, E, 1)).astype(np.float32)
for i in range(20):
    v.step(a[i])
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
for i in range(K):
    v.step(a[i])
dt = torch.tensor([time.perf_counter() - t0], device="cuda")
if world > 1:
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"world {world} BSG_HOST_THREADS={os.environ.get('BSG_HOST_THREADS', 'auto')} cores={os.cpu_count()}: "
          f"{E * world * K / dt.item():.3e} env-steps/s, {dt.item() / K * 1e6:.1f} us/step (max over ranks)", flush=True)
v.close()
if world > 1:
    dist.destroy_process_group()
