"""Parity of the TIMED path: device Philox resets + same-step autoreset + continued stepping, across episode
boundaries, against the float64 oracle envs driven by the same Philox streams (seed, global env id, episode).

tests/test_gpu_env.py checks a step from an injected state and a device reset separately; here the two run together
the way bench.py and SB3 drive the env: every env is created by the device's scenario generator, stepped with random
actions, finishes (termination or the TimeLimit cap), has its terminal observation set aside, is regenerated inside the
same launch and keeps stepping.  The oracle follows with `PhiloxDraws(seed, gid, episode)` per episode; the sim clock
(FMS timer phase) runs on across episodes on both sides.  Tolerances: those of tests/test_gpu_env.py.

An env whose `done` flag differs from the oracle's at some step has crossed a discrete threshold within float32
rounding (waypoint reach at 5 km, polygon exit, ...): it is counted, reported and dropped; the count
must stay a rarity.
"""
import numpy as np
import pytest

from oracle.philox import PhiloxDraws
from tests.common import device_traffic
from tests.test_gpu_env import TOL, ErrStats, _compare_asas_pairs, _compare_obs, _make_oracle, pos_tol, reward_tol

pytestmark = pytest.mark.gpu

VNORM = {"SectorCREnv-v0": (32.0, 66.0), "MergeEnv-v0": (150.0, 150.0)}


def _run(env_id, E, steps, cap, seed, sample=None, n_int=0, cd=False, off=0, density="normal"):
    from bluesky_gym_sasha_b200.vector_env import BlueSkyVectorEnv
    kw = dict(n_intruders=n_int) if n_int else {}
    venv = BlueSkyVectorEnv(env_id, E, seed=seed, env_id_offset=off, cd_enabled=cd, cd_pairs=cd, autoreset_mode="same_step",
                            max_episode_steps=cap, **kw)
    sample = list(range(E)) if sample is None else list(sample)
    gobs, _ = venv.reset()
    orc, episode, t_ep = {}, {}, {}
    for e in sample:
        o = _make_oracle(env_id, draws=PhiloxDraws(seed, off + e, 0), cd=cd, n_int=n_int, density=density)
        oobs, _ = o.reset()
        _compare_obs(gobs, oobs, e, -1)
        orc[e], episode[e], t_ep[e] = o, 0, 0
    rng = np.random.default_rng(seed)
    act_dim = venv.layout.act_dim
    stats = ErrStats()
    dropped, exempt = {}, {e: False for e in sample}
    n_boundaries = n_full = n_pairs = 0
    for step in range(steps):
        a = rng.uniform(-1.0, 1.0, size=(E, act_dim)).astype(np.float32)
        gobs, grew, gterm, gtrunc, ginfo = venv.step(a)
        d = device_traffic(venv)
        for e in sample:
            if e in dropped:
                continue
            o = orc[e]
            lnav_before = o.traf.swlnav.copy()
            oobs, orew, oterm, otrunc, oinfo = o.step(a[e].astype(np.float64))
            # MergeEnv: an intruder that overflies its LAST waypoint freezes ap.trk at a bearing to a point metres away
            # (unbounded sensitivity; tests/test_gpu_env.py): from then on only the ownship's quantities are compared
            exempt[e] = exempt[e] or bool((lnav_before & ~o.traf.swlnav).any())
            t_ep[e] += 1
            otrunc = bool(otrunc) or (cap > 0 and t_ep[e] >= cap)              # gymnasium TimeLimit
            odone = bool(oterm) or otrunc
            if (bool(gterm[e]) or bool(gtrunc[e])) != odone:
                dropped[e] = (step, "done flag")
                continue
            assert bool(gterm[e]) == bool(oterm) and bool(gtrunc[e]) == otrunc, (step, e)
            if not exempt[e]:
                assert abs(grew[e] - orew) < reward_tol(env_id), (step, e, grew[e], orew)
                stats.add("reward", abs(grew[e] - orew))
                for k, v in oinfo.items():
                    if not (isinstance(v, float) and np.isnan(v)):
                        assert abs(ginfo[k][e] - v) < 1e-2 + 1e-4 * abs(v), (step, e, k, ginfo[k][e], v)
                n_full += 1
            if cd and not exempt[e]:
                t = o.traf
                assert ginfo["asas_nconf"][e] == len(t.confpairs) and ginfo["asas_nlos"][e] == len(t.lospairs), (step, e)
            if odone:
                # terminal observation of the finished episode, then the next episode's first observation
                n_boundaries += 1
                fo = {k: v for k, v in ginfo["final_obs"].items()}
                assert ginfo["_final_obs"][e]
                _compare_obs(fo, oobs, e, step, VNORM.get(env_id), ownship_only=exempt[e], stats=stats)
                episode[e] += 1
                t_ep[e] = 0
                exempt[e] = False
                o.draws = PhiloxDraws(seed, off + e, episode[e])
                oobs, _ = o.reset()
                _compare_obs(gobs, oobs, e, step, VNORM.get(env_id), stats=stats)
                t = o.traf
                n = t.ntraf
                assert np.max(np.abs(d["lat"][e, :n] - t.lat)) < 1e-9 and np.max(np.abs(d["lon"][e, :n] - t.lon)) < 1e-9, (step, e)
            else:
                _compare_obs(gobs, oobs, e, step, VNORM.get(env_id), ownship_only=exempt[e], stats=stats)
                if cd and not exempt[e]:
                    n_pairs += _compare_asas_pairs(venv.asas_pairs(e), o.traf, (step, e))
                if not exempt[e]:
                    t = o.traf
                    n = t.ntraf
                    for name, dv, tol in (("pos", np.maximum(np.abs(d["lat"][e, :n] - t.lat), np.abs(d["lon"][e, :n] - t.lon)), pos_tol(env_id)),
                                          ("alt", np.abs(d["alt"][e, :n] - t.alt), TOL["alt"]), ("tas", np.abs(d["tas"][e, :n] - t.tas), TOL["tas"]),
                                          ("vs", np.abs(d["vs"][e, :n] - t.vs), TOL["vs"])):
                        # (FMS-guided aircraft past a waypoint keep a lateral offset of a few metres: test_gpu_env.py)
                        lim = np.where(np.asarray(t.iactwp) >= 1, 5.0 * tol, tol) if name == "pos" else tol
                        assert np.all(dv < lim), (step, e, name, dv.max())
                        stats.add(name, dv.max())
    venv.close()
    print(f"{env_id}: {len(sample)} envs x {steps} steps, {n_boundaries} episode boundaries crossed, {n_full} env-steps compared in "
          f"full, {n_pairs} ASAS conflict pairs compared, dropped {dropped}; max |error|: {stats}")
    return n_boundaries, dropped


@pytest.mark.parametrize("env_id,kw,cap,steps", [
    ("DescentEnv-v0", {}, 300, 130),                                   # ~41-step episodes (runway reached / crash)
    ("PlanWaypointEnv-v0", {}, 35, 130),                               # a random policy never finishes: TimeLimit boundaries
    ("HorizontalCREnv-v0", dict(n_int=20, cd=True), 35, 120),          # BASELINE configs[1] (a random policy rarely reaches the
    ("HorizontalCREnv-v0", dict(n_int=5), 35, 120),                    #  waypoint: TimeLimit boundaries); the reference's default
    ("VerticalCREnv-v0", dict(cd=True), 300, 130),
    ("SectorCREnv-v0", dict(cd=True), 60, 130),                        # polygon exits + TimeLimit
    ("StaticObstacleEnv-v0", {}, 100, 160),
    ("MergeEnv-v0", dict(cd=True), 50, 120),                           # registered cap 50
])
def test_multi_episode_same_step_autoreset_matches_oracle(cuda, env_id, kw, cap, steps):
    E = 8
    n_boundaries, dropped = _run(env_id, E, steps, cap, seed=41, off=5000, **kw)
    assert n_boundaries >= E                                           # every env crossed at least one boundary on average
    assert len(dropped) <= 1, dropped


def test_c2_at_baseline_size_sampled_envs(cuda):
    """BASELINE configs[1] at its real size: 4096 envs x 21 aircraft, CD in every substep, same-step autoreset -- 64
    sampled env ids (first / last CTA, first / last env of a CTA, the last env) against the oracle for 30 steps."""
    E = 4096
    rs = np.random.default_rng(3)
    sample = sorted(set([0, 1, 2, 3, 4, 5, 6, 7, E - 4, E - 3, E - 2, E - 1, 2047, 2048]) | set(rs.choice(E, 50, replace=False).tolist()))
    n_boundaries, dropped = _run("HorizontalCREnv-v0", E, 36, 300, seed=0, sample=sample, n_int=20, cd=True)
    assert len(dropped) <= 1, dropped


def test_sector_at_baseline_size_sampled_envs(cuda):
    """BASELINE configs[2] per-GPU share: 8192 SectorCR envs (polygon, 5..32 aircraft, CD on), 24 sampled env ids, 25 steps,
    registered cap."""
    E = 8192
    rs = np.random.default_rng(4)
    sample = sorted(set([0, 1, E - 1, 4095, 4096]) | set(rs.choice(E, 19, replace=False).tolist()))
    n_boundaries, dropped = _run("SectorCREnv-v0", E, 25, 200, seed=2, sample=sample, cd=True)
    assert len(dropped) <= 1, dropped
