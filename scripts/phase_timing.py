"""Where does the time of one env-step launch go?  Needs a profiling build of the library:

    BSG_LIB_OUT=ab_libs/libbsg_phase.so BSG_EXTRA_NVCC_FLAGS=-DBSG_PHASE_TIMING python -m bluesky_gym_sasha_b200.build --force
    BSG_B200_LIB=ab_libs/libbsg_phase.so python scripts/phase_timing.py [env_id] [E]

Every env's lane 0 stamps %globaltimer at 8 phase boundaries (env_step.cu, BSG_STAMP); the stamps land in the unused
tail of the final_obs buffer.  Printed: per phase the median / p95 / max duration over the envs, when the first and the
last warp started and ended relative to the first start, and the CUDA-event time of the same launch (L2 flushed)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bluesky_gym_sasha_b200.vector_env import BlueSkyVectorEnv  # noqa: E402

NAMES = ["load ei32 + state + action", "compute_targets", "substep 0", "substeps 1..n-1", "load env record + obs/reward",
         "autoreset (generator + obs)", "stores"]


def main():
    env_id = sys.argv[1] if len(sys.argv) > 1 else "HorizontalCREnv-v0"
    E = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    kw = dict(cd_enabled=True, n_intruders=20) if env_id == "HorizontalCREnv-v0" else dict(cd_enabled=True)
    v = BlueSkyVectorEnv(env_id, E, seed=0, autoreset_mode="same_step", **kw)
    v.reset_torch()
    g = torch.Generator(device="cuda").manual_seed(0)
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
    rows = []
    for i in range(60):
        a = torch.rand((E, v.layout.act_dim), device="cuda", generator=g) * 2 - 1
        flush.fill_(float(i))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        v.step_torch(a)
        e1.record()
        torch.cuda.synchronize()
        if i < 20:
            continue
        raw = v.t["final_obs"].view(-1).view(torch.int64)[-8 * E:].view(E, 8).cpu().numpy()
        smid = (raw[:, 7] & 0xff).astype(np.int64)
        st = raw.astype(np.float64)
        good = np.all(np.diff(raw[:, :7], axis=1) >= 0, axis=1) & (raw[:, 0] > 0) & (raw[:, 6] - raw[:, 0] < 10_000_000)
        t0 = st[good, 0].min()
        st = (st - t0) * 1e-3                                    # us since the first warp started
        st[~good] = st[good].mean(axis=0)                        # (envs whose stamps are incomplete take the average: not counted as outliers)
        i32 = v.t["env_i32"].cpu().numpy()
        kin, cmd = v.t["kin"][:, 0].cpu().numpy(), v.t["cmd"][:, 0].cpu().numpy()
        turning = np.abs((kin[:, 2] - cmd[:, 3] + 180.0) % 360.0 - 180.0) > 0.01        # ownship still turning after the step
        fl = v.t["flags"].cpu().numpy()
        kin_all, cmd_all = v.t["kin"].cpu().numpy(), v.t["cmd"].cpu().numpy()
        alive = (fl & 1) != 0
        nturn = (alive & (np.abs((kin_all[:, :, 2] - cmd_all[:, :, 3] + 180.0) % 360.0 - 180.0) > 0.01)).sum(axis=1)   # aircraft turning
        nlnav = (alive & ((fl & 2) != 0)).sum(axis=1)
        rows.append((e0.elapsed_time(e1) * 1e3, st, int(v.t["final_count"][int(v.t["final_count"][2])]), smid,
                     i32[:, 10].astype(np.float64), i32[:, 11].astype(np.float64), turning, nturn.astype(np.float64),
                     nlnav.astype(np.float64), alive.sum(axis=1).astype(np.float64), i32[:, 0].astype(np.float64)))
    ev = np.median([r[0] for r in rows])
    print(f"{env_id} E={E}: CUDA-event time per launch {ev:.1f} us (median of {len(rows)}), envs finishing per step "
          f"{np.mean([r[2] for r in rows]):.0f}")
    st = np.stack([r[1] for r in rows])                          # [launch, env, stamp]
    d = np.diff(st, axis=2)
    print(f"{'phase':34s} {'median':>8s} {'p95':>8s} {'max':>8s}   [us per env]")
    for k, name in enumerate(NAMES):
        x = d[:, :, k].reshape(-1)
        if k == 5:
            x = x[x > 0.5] if (x > 0.5).any() else x              # only the envs that reset
        print(f"{name:34s} {np.median(x):8.2f} {np.percentile(x, 95):8.2f} {x.max():8.2f}")
    start, end = st[:, :, 0], st[:, :, 7]
    print(f"warp start: median {np.median(start):.2f}  p95 {np.percentile(start, 95):.2f}  max {start.max(axis=1).mean():.2f} us after the first")
    print(f"warp end  : median {np.median(end):.2f}  p95 {np.percentile(end, 95):.2f}  max {end.max(axis=1).mean():.2f} us after the first start")
    print(f"lifetime of a warp: median {np.median(end - start):.2f}  p95 {np.percentile(end - start, 95):.2f}  max {(end - start).max():.2f} us")
    # who are the stragglers?  phase breakdown of the slowest 2 % of the warps, and whether they sit on particular SMs
    life = end - start
    thr = np.percentile(life, 98)
    slow = life >= thr
    print(f"slowest 2 % of the warps (lifetime >= {thr:.1f} us): mean phase durations "
          + ", ".join(f"{NAMES[k].split()[0]} {d[:, :, k][slow].mean():.1f}" for k in range(7))
          + "  | all warps: " + ", ".join(f"{d[:, :, k].mean():.1f}" for k in range(7)))
    sm = np.stack([r[3] for r in rows])
    per_sm_load = np.array([[np.sum(sm[l] == k) for k in range(160)] for l in range(sm.shape[0])])      # envs (warps) per SM
    ld = per_sm_load[0][per_sm_load[0] > 0]
    print(f"warps per SM: min {ld.min()} max {ld.max()} (over {len(ld)} SMs)")
    for cnt in sorted(set(ld.tolist())):
        sel = np.isin(sm[0], np.where(per_sm_load[0] == cnt)[0])
        print(f"   SMs holding {cnt} warps: mean lifetime {life[0][sel].mean():.2f} us, max {life[0][sel].max():.2f} us, share of slow warps "
              f"{(slow[0] & sel).sum() / max(slow[0].sum(), 1) * 100:.0f} %")
    sm_slow = np.bincount(sm[slow], minlength=160)
    print("slow warps per SM (top 8 SMs):", sorted(sm_slow.tolist(), reverse=True)[:8], "of", int(slow.sum()))
    # what makes a warp slow?  conflicts / LoS found in the env's last substep (= exact pair evaluations), ownship turning
    nconf, nlos, turning = (np.stack([r[k] for r in rows]) for k in (4, 5, 6))
    print(f"correlation of warp lifetime with nconf {np.corrcoef(life.ravel(), nconf.ravel())[0, 1]:.2f}, nlos "
          f"{np.corrcoef(life.ravel(), nlos.ravel())[0, 1]:.2f}, ownship turning {np.corrcoef(life.ravel(), turning.ravel().astype(float))[0, 1]:.2f}")
    for lo, hi in ((0, 1), (1, 5), (5, 10), (10, 20), (20, 40), (40, 1000)):
        sel = (nconf >= lo) & (nconf < hi)
        if sel.any():
            print(f"   nconf in [{lo}, {hi}): {sel.mean() * 100:5.1f} % of the envs, lifetime mean {life[sel].mean():.2f} p95 {np.percentile(life[sel], 95):.2f} us, "
                  f"substeps 1..n-1 mean {d[:, :, 3][sel].mean():.2f} us")
    print(f"   ownship turning: {turning.mean() * 100:.1f} % of the envs, lifetime mean {life[turning].mean():.2f} vs {life[~turning].mean():.2f} us")
    for name, k in (("aircraft turning after the step", 7), ("aircraft under LNAV", 8), ("aircraft alive", 9), ("env step counter", 10)):
        x = np.stack([r[k] for r in rows])
        if x.std() > 0:
            print(f"   correlation of warp lifetime with {name}: {np.corrcoef(life.ravel(), x.ravel())[0, 1]:.2f}; "
                  + ", ".join(f"{int(q)}: {life[x == q].mean():.1f} us ({(x == q).mean() * 100:.0f} %)" for q in np.unique(x)[:12]))
    sm_n = np.array([[nconf[l][sm[l] == k].sum() for k in range(160)] for l in range(sm.shape[0])])
    sm_life = np.array([[life[l][sm[l] == k].max() if (sm[l] == k).any() else np.nan for k in range(160)] for l in range(sm.shape[0])])
    ok = ~np.isnan(sm_life)
    print(f"per SM: correlation of the SM's last warp end with the SM's total nconf {np.corrcoef(sm_life[ok], sm_n[ok])[0, 1]:.2f}")
    fin = d[:, :, 5] > 0.5
    print(f"envs that reset: {fin.mean() * 100:.1f} %; their lifetime median {np.median((end - start)[fin]) if fin.any() else 0:.2f} us")


if __name__ == "__main__":
    main()
