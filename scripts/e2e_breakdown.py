"""Where the time of one BlueSkyVectorEnv.step() goes (HorizontalCR-20, E = 4096).  Run on a GPU box."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from bluesky_gym_sasha_b200 import _lib
from bluesky_gym_sasha_b200.vector_env import BlueSkyVectorEnv

E = 4096
v = BlueSkyVectorEnv("HorizontalCREnv-v0", E, seed=0, cd_enabled=True, n_intruders=20, autoreset_mode="same_step")
v.reset()
at0 = torch.rand((E, 1), device="cuda") * 2 - 1
for _ in range(400):                    # episodes of different lengths: the finished envs spread over the steps, as in bench.py
    v.step_torch(at0)
a = np.random.default_rng(0).uniform(-1, 1, (E, 1)).astype(np.float32)


def t(f, n=300):
    for _ in range(20):
        f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


print("full step()                 %.1f us" % t(lambda: v.step(a)))
blk = v._blocks[0]
h = blk.v
print("bsg_step_host_block only    %.1f us" % t(lambda: _lib.check(v._lib.bsg_step_host_block(v._h, v._act_ptr, blk.ptr, v._out_bytes, v._stream()))))
print("_acquire                    %.1f us" % t(lambda: v._acquire()))
print("_step_results               %.1f us" % t(lambda: v._step_results(blk)))
v.obs_dtype = np.dtype(np.float64)
print("full step() float64 obs     %.1f us" % t(lambda: v.step(a)))
print("_step_results float64       %.1f us" % t(lambda: v._step_results(blk)))
v.obs_dtype = np.dtype(np.float32)
at = torch.from_numpy(a).cuda()
print("step_torch (async)          %.1f us" % t(lambda: v.step_torch(at)))
o32 = h["obs"]
print("obs dict (copy)             %.1f us" % t(lambda: v._obs_dict_np(o32)))
print("obs f32 copy                %.1f us" % t(lambda: o32.copy()))
print("infos                       %.1f us" % t(lambda: v._infos_np(h["info"])))
print("rew/term/trunc astype       %.1f us" % t(lambda: (h["reward"].astype(np.float64), h["terminated"].astype(bool), h["truncated"].astype(bool))))
print("actions into pinned         %.1f us" % t(lambda: v._act_np.__setitem__(Ellipsis, np.asarray(a, dtype=np.float32).reshape(E, 1))))
v.copy = False
print("full step() copy=False      %.1f us" % t(lambda: v.step(a)))
d = torch.empty((v._out_bytes,), dtype=torch.uint8, device="cuda")
hp = torch.empty((v._out_bytes,), dtype=torch.uint8).pin_memory()


def d2h():
    hp.copy_(d, non_blocking=True)
    torch.cuda.synchronize()


print("D2H %.2f MB pinned+sync      %.1f us" % (v._out_bytes / 1e6, t(d2h)))
