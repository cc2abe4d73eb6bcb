import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes as C
from bluesky_gym_sasha_b200.vector_env import BlueSkyVectorEnv, _ptr
from bluesky_gym_sasha_b200 import _lib
E=4096
v = BlueSkyVectorEnv("HorizontalCREnv-v0", E, seed=0, cd_enabled=True, n_intruders=20, autoreset_mode="same_step")
v.reset()
a = np.random.default_rng(0).uniform(-1,1,(E,1)).astype(np.float32)
def t(f, n=200):
    for _ in range(10): f()
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/n*1e6
print("full step()            %.1f us" % t(lambda: v.step(a)))
h=v._hbuf[0]
def raw():
    _lib.check(v._lib.bsg_step_host(v._h, _ptr(v.h["actions"]), _ptr(h["obs"]), _ptr(h["reward"]), _ptr(h["terminated"]), _ptr(h["truncated"]), _ptr(h["info"]), _ptr(h["final_count"]), v._stream()))
print("bsg_step_host only     %.1f us" % t(raw))
at = torch.from_numpy(a).cuda()
print("step_torch (async)     %.1f us" % t(lambda: v.step_torch(at)))
o32 = h["obs"].numpy()
print("obs astype f64         %.1f us" % t(lambda: o32.astype(np.float64)))
dst = np.empty(o32.shape, np.float64)
print("np.copyto f64 prealloc %.1f us" % t(lambda: np.copyto(dst, o32)))
print("obs dict slicing       %.1f us" % t(lambda: v._obs_dict_np(o32)))
print("infos                  %.1f us" % t(lambda: v._infos_np(h["info"].numpy())))
print("obs f32 copy            %.1f us" % t(lambda: o32.copy()))
v.copy=False
print("full step() copy=False %.1f us" % t(lambda: v.step(a)))
d = torch.empty((E, 103), dtype=torch.float32, device="cuda"); hp = torch.empty((E,103), dtype=torch.float32).pin_memory()
def d2h(): hp.copy_(d, non_blocking=True); torch.cuda.synchronize()
print("D2H 1.7MB pinned+sync  %.1f us" % t(d2h))
