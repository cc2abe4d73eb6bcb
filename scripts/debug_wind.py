import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bluesky_gym_sasha_b200.vector_env import BlueSkyVectorEnv
from tests.test_golden import _OracleWind, _oracle, _load
from tests.test_gpu_env import _inject
from tests.common import device_traffic
env_id = sys.argv[1] if len(sys.argv) > 1 else "SectorCREnv-v0"
g = _load(env_id, True)
kw = dict(wind=dict(lat=g["wind_lat"], lon=g["wind_lon"], vnorth=g["wind_vnorth"], veast=g["wind_veast"]), wind_obs=bool(g["augment_obs"]))
venv = BlueSkyVectorEnv(env_id, 1, seed=7, autoreset_mode="disabled", max_episode_steps=0, **kw)
venv.reset()
np.random.seed(0); random.seed(0)
o = _OracleWind(_oracle(env_id), g)
o.reset()
_inject(venv, 0, o.env, env_id)
n0 = o.traf.ntraf
venv._wind_t["gs"][0, :n0] = torch.as_tensor(np.stack([o.traf.gsnorth, o.traf.gseast], 1), dtype=torch.float32, device=venv.device)
a = g["s0_action"][1]
print("hdg0", np.round(o.traf.hdg[:6], 3))
for sub in range(1, 6):
    venv.traf_update(1); o.traf.simstep() if sub > 0 else None
    d = device_traffic(venv)
    t = o.traf
    print("sub", sub, "dhdg", np.round(d["hdg"][0, :6] - t.hdg[:6], 4), "hdg", np.round(t.hdg[:6], 3), "dlat*1e6", np.round((d["lat"][0, :6] - t.lat[:6]) * 1e6, 2))
print("---- step mode")
venv2 = BlueSkyVectorEnv(env_id, 1, seed=7, autoreset_mode="disabled", max_episode_steps=0, **kw)
venv2.reset()
np.random.seed(0); random.seed(0)
o2 = _OracleWind(_oracle(env_id), g)
o2.reset()
_inject(venv2, 0, o2.env, env_id)
venv2._wind_t["gs"][0, :n0] = torch.as_tensor(np.stack([o2.traf.gsnorth, o2.traf.gseast], 1), dtype=torch.float32, device=venv2.device)
gobs, grew, _, _, _ = venv2.step(a.reshape(1, -1).astype(np.float32))
oobs, orew, _, _, _ = o2.step(a.copy())
d = device_traffic(venv2); t = o2.traf
print("dhdg", np.round(d["hdg"][0, :8] - t.hdg[:8], 4), "dtas", np.round(d["tas"][0, :8] - t.tas[:8], 4))
for k in oobs:
    print(k, np.round(gobs[k][0], 4), np.round(oobs[k], 4))
print("golden vx_r", g["s0_obs_vx_r"][1])
