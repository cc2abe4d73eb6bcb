import sys; import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, time, ctypes as C
from bluesky_gym_sasha_b200 import _lib
lib=_lib.load()
a = np.random.rand(4096,103).astype(np.float32)
def t(f,n=2000):
    for _ in range(50): f()
    t0=time.perf_counter()
    for _ in range(n): f()
    return (time.perf_counter()-t0)/n*1e6
def mt():
    d=np.empty_like(a); lib.bsg_host_copy(C.c_void_p(d.ctypes.data), C.c_void_p(a.ctypes.data), a.nbytes); return d
pre=np.empty_like(a)
def mtpre():
    lib.bsg_host_copy(C.c_void_p(pre.ctypes.data), C.c_void_p(a.ctypes.data), a.nbytes)
print("mt copy fresh %.1f us" % t(mt))
print("mt copy prealloc %.1f us" % t(mtpre))
libc=C.CDLL("libc.so.6")
print(libc.mallopt(-3, 1<<25), libc.mallopt(-1, 1<<28))
print("mt copy fresh, tuned malloc %.1f us" % t(mt))
held=[]
def mthold():
    held.append(mt())
    if len(held)>4: held.pop(0)
print("mt copy fresh, tuned, 4 held %.1f us" % t(mthold))
