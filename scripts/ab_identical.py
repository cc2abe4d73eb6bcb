"""Do two library builds (BSG_B200_LIB) produce bit-identical results?  Each library runs the same seeded rollouts in its own
process; every `probe_every` steps the state is saved and the substep loop is run for n = 1 .. n_sub substeps from it
(bsg_traf_update), so the in-sim CD outputs after EVERY substep count are compared, not only the env step's last one.

    python scripts/ab_identical.py ab_libs/base.so ab_libs/reuse1.so
"""
import hashlib
import json
import os
import subprocess
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, os, json, hashlib, torch
sys.path.insert(0, %r)
from bluesky_gym_sasha_b200.vector_env import BlueSkyVectorEnv
CASES = [("HorizontalCREnv-v0", 1024, dict(n_intruders=20), 10), ("HorizontalCREnv-v0", 1024, dict(n_intruders=12), 10),
         ("SectorCREnv-v0", 1024, {}, 5), ("MergeEnv-v0", 1024, {}, 10), ("VerticalCREnv-v0", 1024, {}, 10),
         ("StaticObstacleEnv-v0", 1024, {}, 5)]
out = {}
for env_id, E, kw, nsub in CASES:
    v = BlueSkyVectorEnv(env_id, E, seed=3, cd_enabled=True, autoreset_mode="same_step", **kw)
    v.reset_torch()
    g = torch.Generator(device="cuda"); g.manual_seed(7)
    h = hashlib.sha256()
    nconf_total = 0
    def eat(keys):
        for k in keys:
            h.update(v.t[k].cpu().numpy().tobytes())
    for step in range(120):
        a = torch.rand((E, v.layout.act_dim), device="cuda", generator=g) * 2 - 1
        if step %% 8 == 0:
            sd = v.state_dict()
            for n in range(1, nsub + 1):
                v.load_state_dict(sd)
                v.traf_update(n)
                eat(("tcpamax", "inconf", "env_i32", "pos", "kin"))
            v.load_state_dict(sd)
        v.step_torch(a)
        eat(("tcpamax", "inconf", "env_i32", "pos", "kin", "obs", "reward", "terminated", "truncated", "info"))
        nconf_total += int(v.t["info"][4].sum().item())
    out["%%s %%s" %% (env_id, kw)] = [h.hexdigest()[:16], nconf_total]
    v.close()
print(json.dumps(out))
''' % root

res = {}
for lib in sys.argv[1:]:
    p = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, BSG_B200_LIB=os.path.join(root, lib)),
                       capture_output=True, text=True)
    if p.returncode:
        print(lib, "FAILED\n", p.stderr[-2000:])
        sys.exit(1)
    res[lib] = json.loads(p.stdout.strip().splitlines()[-1])
    print(lib, res[lib])
libs = list(res)
same = all(res[l] == res[libs[0]] for l in libs[1:])
print("IDENTICAL" if same else "DIFFERENT")
sys.exit(0 if same else 2)
