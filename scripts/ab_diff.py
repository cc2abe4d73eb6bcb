"""Where do two library builds start to differ?  Runs the same seeded rollout under each library (own process), hashes every
state / output tensor after every step, reports the first (case, step, tensor) that differs and how large the difference is.

    python scripts/ab_diff.py bluesky_gym_sasha_b200/libbsg_b200.so ab_libs/base.so
"""
import json
import os
import subprocess
import sys

import numpy as np

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ("pos", "kin", "cmd", "tcpamax", "inconf", "env_i32", "env_f32", "obs", "reward", "terminated", "truncated")
code = r'''
import sys, json, hashlib, numpy as np, torch
sys.path.insert(0, %r)
from bluesky_gym_sasha_b200.vector_env import BlueSkyVectorEnv
CASES = [("HorizontalCREnv-v0", 256, dict(n_intruders=20)), ("SectorCREnv-v0", 256, {}), ("MergeEnv-v0", 256, {})]
KEYS = %r
dump = json.loads(sys.argv[1]) if len(sys.argv) > 1 else None          # [case index, step]: save the tensors of that step
out = []
for ci, (env_id, E, kw) in enumerate(CASES):
    v = BlueSkyVectorEnv(env_id, E, seed=3, cd_enabled=True, autoreset_mode="same_step", **kw)
    v.reset_torch()
    g = torch.Generator(device="cuda"); g.manual_seed(7)
    rows = []
    for step in range(40):
        a = torch.rand((E, v.layout.act_dim), device="cuda", generator=g) * 2 - 1
        v.step_torch(a)
        torch.cuda.synchronize()
        rows.append({k: hashlib.sha1(v.t[k].cpu().numpy().tobytes()).hexdigest()[:12] for k in KEYS})
        if dump and dump[0] == ci and dump[1] == step:
            np.savez(sys.argv[2], **{k: v.t[k].cpu().numpy() for k in KEYS})
    out.append(rows)
    v.close()
print(json.dumps(out))
''' % (root, KEYS)


def run(lib, *extra):
    p = subprocess.run([sys.executable, "-c", code, *extra], env=dict(os.environ, BSG_B200_LIB=os.path.join(root, lib)),
                       capture_output=True, text=True)
    if p.returncode:
        print(lib, "FAILED\n", p.stderr[-2000:])
        sys.exit(1)
    return json.loads(p.stdout.strip().splitlines()[-1])


la, lb = sys.argv[1:3]
ra, rb = run(la), run(lb)
names = ["HorizontalCR-20", "SectorCR", "MergeEnv"]
for ci, (xa, xb) in enumerate(zip(ra, rb)):
    first = None
    for step, (ha, hb) in enumerate(zip(xa, xb)):
        bad = [k for k in KEYS if ha[k] != hb[k]]
        if bad:
            first = (step, bad)
            break
    if first is None:
        print(f"{names[ci]}: identical over {len(xa)} steps")
        continue
    step, bad = first
    print(f"{names[ci]}: first difference at step {step} in {bad}")
    run(la, json.dumps([ci, step]), "/tmp/ab_a.npz")
    run(lb, json.dumps([ci, step]), "/tmp/ab_b.npz")
    A, B = np.load("/tmp/ab_a.npz"), np.load("/tmp/ab_b.npz")
    for k in bad:
        a, b = A[k].astype(np.float64), B[k].astype(np.float64)
        d = np.abs(a - b)
        idx = np.unravel_index(np.nanargmax(d), d.shape)
        print(f"   {k}: {int((a != b).sum())} of {a.size} elements differ, max |diff| {np.nanmax(d):.3g} at {idx}: {a[idx]!r} vs {b[idx]!r}")
