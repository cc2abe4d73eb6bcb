import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bluesky_gym_sasha_b200.vector_env import BlueSkyVectorEnv
from bluesky_gym_sasha_b200.cd import StateBasedCD
from oracle import envs as oenvs, statebased
from tests.common import inject_oracle_env, device_traffic
from tests.test_gpu_env import _inject
np.random.seed(1234)
o = oenvs.HorizontalCREnv(n_intruders=20, cd_enabled=True)
o.reset()
venv = BlueSkyVectorEnv("HorizontalCREnv-v0", 1, seed=7, cd_enabled=True, autoreset_mode="disabled", max_episode_steps=0, n_intruders=20)
venv.reset()
_inject(venv, 0, o, "HorizontalCREnv-v0")
t = o.traf
for k in range(9): t.simstep()
venv.traf_update(9)
d = device_traffic(venv)
n = t.ntraf
S = (t.lat, t.lon, t.trk, t.gs, t.alt, t.vs)
D = (d["lat"][0,:n], d["lon"][0,:n], d["hdg"][0,:n], d["tas"][0,:n], d["alt"][0,:n], d["vs"][0,:n])
print("state diffs", [float(np.max(np.abs(a-b))) for a,b in zip(S,D)])
mS = statebased.detect_rows(np.arange(n), *S, with_margins=True)
mD = statebased.detect_rows(np.arange(n), *D, with_margins=True)
pS = set(zip(*map(lambda x:x.tolist(), np.where(mS["swconfl"])))); pD = set(zip(*map(lambda x:x.tolist(), np.where(mD["swconfl"]))))
g = StateBasedCD().detect(*D, lat0=float(D[0][0]), lon0=float(D[1][0]))
pG = set(map(tuple, g["confpairs"].tolist()))
print("oracle(S9)", sorted(pS)); print("oracle(D9)", sorted(pD)); print("K2(D9)", sorted(pG))
venv.traf_update(1)
d2 = device_traffic(venv)
print("K3 inconf", np.where(d2["inconf"][0])[0], "nconf", venv.t["env_i32"][0, 10].item())
t.simstep()
print("oracle inconf", np.where(t.inconf)[0], len(t.confpairs))
for p in sorted(pG ^ pS):
    i,j = p
    print(p, {k: float(mS[k][i,j]) for k in ("tcpa","dcpa2","tinconf","toutconf","dist")}, "near", bool(mS["near_conf"][i,j]))
