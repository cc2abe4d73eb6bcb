"""Experiment: what does the env kernel cost when its observation stores go straight to pinned host memory (zero copy over
PCIe) instead of HBM + a device->host copy afterwards?  The bound obs pointer is swapped for a mapped pinned block; nothing
else changes (next_step autoreset: the same-step path reads the observation back).  Run on a GPU box."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from bluesky_gym_sasha_b200 import _lib
from bluesky_gym_sasha_b200.vector_env import BlueSkyVectorEnv, _ptr

E = 4096
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")


def run(host_obs, steps=200):
    v = BlueSkyVectorEnv("HorizontalCREnv-v0", E, seed=0, cd_enabled=True, n_intruders=20, autoreset_mode="next_step")
    hobs = None
    if host_obs:
        hobs = torch.zeros((E, v.layout.obs_dim), dtype=torch.float32).pin_memory()
        t = dict(v.t)
        t["obs"] = hobs
        tt = _lib.TensorTable(**{k: _ptr(x) for k, x in t.items()})
        _lib.check(v._lib.bsg_bind_state(v._h, C.byref(tt)))
    v.reset_torch()
    a = torch.rand((E, 1), device="cuda") * 2 - 1
    for _ in range(30):
        v.step_torch(a)
    torch.cuda.synchronize()
    ts = []
    for i in range(steps):
        flush.fill_(float(i))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        v.step_torch(a)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ref = v.t["obs"].cpu().numpy() if not host_obs else hobs.numpy().copy()
    v.close()
    return float(np.median(ts)), float(np.mean(ts)), ref


m0, a0, r0 = run(False)
m1, a1, r1 = run(True)
print(f"obs in HBM        : median {m0:.1f} mean {a0:.1f} us per step (+ D2H of the obs afterwards)")
print(f"obs in pinned host: median {m1:.1f} mean {a1:.1f} us per step; same observations: {np.array_equal(r0, r1)}")
