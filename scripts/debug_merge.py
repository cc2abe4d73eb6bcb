import sys, os, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bluesky_gym_sasha_b200.vector_env import BlueSkyVectorEnv
from oracle import envs as oenvs
from tests.common import device_traffic, angdiff
from tests.test_gpu_env import _inject
env_id = sys.argv[1] if len(sys.argv) > 1 else "MergeEnv-v0"
random.seed(1234); np.random.seed(1234)
E=4
mk = {"MergeEnv-v0": oenvs.MergeEnv, "SectorCREnv-v0": oenvs.SectorCREnv, "DescentEnv-v0": oenvs.DescentEnv, "HorizontalCREnv-v0": oenvs.HorizontalCREnv}[env_id]
os_ = [mk() for _ in range(E)]
[o.reset() for o in os_]
venv = BlueSkyVectorEnv(env_id, E, seed=7, autoreset_mode="disabled", max_episode_steps=0)
venv.reset()
for e,o in enumerate(os_): _inject(venv, e, o, env_id)
rng = np.random.default_rng(0)
for step in range(int(sys.argv[2]) if len(sys.argv)>2 else 12):
    a = rng.uniform(-1,1,size=(E,venv.layout.act_dim)).astype(np.float32)
    gobs, *_ = venv.step(a)
    d = device_traffic(venv)
    worst = {}
    for e,o in enumerate(os_):
        oobs, *_ = o.step(a[e].astype(np.float64)); t=o.traf; n=t.ntraf
        for k in ("lat","lon","alt","tas","vs","cas","selspd"):
            worst[k] = max(worst.get(k,0), float(np.max(np.abs(d[k][e,:n]-getattr(t,k)))))
        worst["hdg"] = max(worst.get("hdg",0), float(np.max(angdiff(d["hdg"][e,:n], t.hdg))))
        worst["aptrk"] = max(worst.get("aptrk",0), float(np.max(angdiff(d["ap_trk"][e,:n], t.ap_trk))))
        for k,v in oobs.items():
            worst["obs:"+k] = max(worst.get("obs:"+k,0), float(np.max(np.abs(gobs[k][e]-v))))
    print(step, {k: float(f"{v:.2e}") for k,v in worst.items()})
