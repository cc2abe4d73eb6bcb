"""Device time of the stages of one AirspaceTraffic substep at N = 100k (CUDA events, median of 20)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
import bench
from bluesky_gym_sasha_b200 import _lib
from bluesky_gym_sasha_b200.cd import _ptr
from bluesky_gym_sasha_b200 import traffic as T

dev = torch.device("cuda", 0)
stages = {}
orig_step = T.AirspaceTraffic.step
def timed(name, fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); r = fn(); b.record()
    stages.setdefault(name, []).append((a, b))
    return r
def step(self, n_sub=1, detect=True):
    cd = self.cd
    with torch.cuda.device(self.device):
        st = self._stream()
        for _ in range(n_sub):
            self.nstep += 1
            fms_ready = (self.nstep % self.fms_rel_freq) == 0
            timed("pack", lambda: _lib.check(self.lib.bsg_traf_pack(C.byref(self.cfg), C.byref(self.tt), _ptr(self.rec), st)))
            out = timed("detect", lambda: cd.detect_packed(self.rec, self.n, want_pairs=True, cull=self.cull, symmetric=self.symmetric))
            self.last = out
            timed("substep", lambda: _lib.check(self.lib.bsg_traf_substep(C.byref(self.cfg), C.byref(self.tt), _ptr(self.rec), int(fms_ready),
                  _ptr(out["pairs"]), _ptr(out["attr"]), _ptr(out["npairs"]), None, cd.pair_capacity, _ptr(self.work), self.work.numel(), st)))
T.AirspaceTraffic.step = step
r = bench.bench_traffic(torch, dev, 6543.4)
torch.cuda.synchronize()
for k, v in stages.items():
    ts = sorted(a.elapsed_time(b) * 1e3 for a, b in v[-20:])
    print(f"{k:12s} median {ts[len(ts)//2]:8.1f} us   min {ts[0]:8.1f}")
print({k: r[k] for k in ("ms_per_substep", "n_conf", "n_los", "counters", "substep_call_us")}, r["substep_kernel"])
