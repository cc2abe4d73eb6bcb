"""Experiment: would a CUDA graph of the e2e step's three stream operations (H2D actions, env kernel, D2H output block) beat
issuing them one by one?  The graph is captured around bsg_step_host_begin and replayed (the replays reuse one final_count
slot, so only the timing is meaningful).  Run on a GPU box."""
import ctypes as C
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
for _ in range(300):
    v.step_torch(at0)
a = np.random.default_rng(0).uniform(-1, 1, (E, 1)).astype(np.float32)
v._act_np[...] = a
blk = v._blocks[0]
lib = v._lib
s = torch.cuda.Stream()


def plain():
    _lib.check(lib.bsg_step_host_block(v._h, v._act_ptr, blk.ptr, v._out_bytes, s.cuda_stream))


def timeit(f, n=500):
    for _ in range(50):
        f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


print("three stream operations + synchronize: %.1f us per step" % timeit(plain))
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=s):
    _lib.check(lib.bsg_step_host_begin(v._h, v._act_ptr, blk.ptr, v._out_bytes, s.cuda_stream))


def graphed():
    with torch.cuda.stream(s):
        g.replay()
    s.synchronize()


print("one graph launch + synchronize       : %.1f us per step" % timeit(graphed))
