"""Wrappers of the reference (bluesky_gym/wrappers/) on the accelerated envs."""
import numpy as np
import pytest

from oracle.philox import noise_normals


@pytest.mark.gpu
def test_noisy_observation_vector_env_matches_philox_oracle(cuda):
    """Device noise = clean observation + sigma * (the oracle's Philox normal sequence), element by element, for
    reset, step and the terminal observations of same-step autoreset; reward / flags / state are untouched."""
    import bluesky_gym  # noqa: F401  (alias package)
    from bluesky_gym.wrappers.uncertainty import NoisyObservationWrapper
    from bluesky_gym_sasha_b200.vector_env import BlueSkyVectorEnv
    E, seed, off, sigma = 24, 5, 300, 0.2
    kw = dict(seed=seed, env_id_offset=off, autoreset_mode="same_step", max_episode_steps=3)
    clean = BlueSkyVectorEnv("HorizontalCREnv-v0", E, **kw)
    noisy = NoisyObservationWrapper(BlueSkyVectorEnv("HorizontalCREnv-v0", E, **kw), noise_level=sigma)
    dim = clean.layout.obs_dim

    def flat(o):
        return np.concatenate([o[k] for k in clean.obs_layout], axis=1).astype(np.float64)

    oc, _ = clean.reset()
    on, _ = noisy.reset()
    call = 0
    z = np.stack([noise_normals(seed, off + e, call, dim) for e in range(E)])
    np.testing.assert_allclose(flat(on) - flat(oc), sigma * z, rtol=0, atol=2e-5)
    rng = np.random.default_rng(0)
    for step in range(1, 7):
        a = rng.uniform(-1, 1, (E, 1)).astype(np.float32)
        oc, rc, tc, uc, ic = clean.step(a)
        on, rn, tn, un, inn = noisy.step(a)
        call += 1
        z = np.stack([noise_normals(seed, off + e, call, dim) for e in range(E)])
        np.testing.assert_allclose(flat(on) - flat(oc), sigma * z, rtol=0, atol=2e-5)
        assert np.array_equal(rc, rn) and np.array_equal(tc, tn) and np.array_equal(uc, un)
        if "final_obs" in ic:
            done = ic["_final_obs"]
            assert done.any() and np.array_equal(done, inn["_final_obs"])
            zt = np.stack([noise_normals(seed, off + e, call, dim, terminal=True) for e in range(E)])
            d = flat(inn["final_obs"]) - flat(ic["final_obs"])
            np.testing.assert_allclose(d[done], sigma * zt[done], rtol=0, atol=2e-5)
    assert step == 6 and call == 6
    # statistics over a larger batch: zero mean, the requested standard deviation
    big = BlueSkyVectorEnv("HorizontalCREnv-v0", 4096, seed=1, n_intruders=20, obs_noise=0.1)
    ref = BlueSkyVectorEnv("HorizontalCREnv-v0", 4096, seed=1, n_intruders=20)
    d = flat2 = np.concatenate([v for v in big.reset()[0].values()], axis=1) - np.concatenate([v for v in ref.reset()[0].values()], axis=1)
    assert abs(d.mean()) < 1e-3 and abs(d.std() - 0.1) < 1e-3
    for v in (clean, noisy, big, ref):
        v.close()


@pytest.mark.gpu
def test_noisy_observation_scalar_env_follows_reference_draws(cuda):
    """Around a scalar env the wrapper is the reference's host code: same np.random draws for the same seed."""
    import bluesky_gym
    from bluesky_gym.wrappers.uncertainty import NoisyObservationWrapper
    bluesky_gym.register_envs()
    env = NoisyObservationWrapper(bluesky_gym.make("DescentEnv-v0"), noise_level=0.3)
    ref = bluesky_gym.make("DescentEnv-v0")
    np.random.seed(4)
    o, _ = env.reset()
    oc, _ = ref.reset()
    np.random.seed(4)
    for k in oc:                                  # dict order = declaration order, one draw call per key
        np.testing.assert_allclose(o[k], oc[k] + np.random.normal(0, 0.3, size=oc[k].shape))
    o, r, te, tr, info = env.step(np.array([0.1]))
    assert set(o) == set(oc) and isinstance(r, float)
    from bluesky_gym_sasha_b200 import gym_compat
    assert isinstance(env, gym_compat.Wrapper) and isinstance(env, gym_compat.Env)      # what Monitor / check_env need
    assert env.observation_space is ref.observation_space or list(env.observation_space.keys()) == list(ref.observation_space.keys())
    assert env.unwrapped.ENV_ID == "DescentEnv-v0"
    env.close()
    ref.close()


WIND4 = dict(lat=np.array([51.9, 51.9, 52.1, 52.1]), lon=np.array([3.9, 4.1, 3.9, 4.1]),
             vnorth=np.array([[16.0, 12.0, 14.0, 15.0]]), veast=np.array([[3.0, 7.0, 9.0, 4.0]]))


@pytest.mark.gpu
def test_wind_field_wrapper_api(cuda):
    """wrappers/README.md usage, unchanged: scalar env from gym.make inside WindFieldWrapper(augment_obs=True)."""
    import bluesky_gym
    from bluesky_gym.wrappers.wind import WindFieldWrapper
    from bluesky_gym_sasha_b200.vector_env import BlueSkyVectorEnv
    bluesky_gym.register_envs()
    env = bluesky_gym.make("MergeEnv-v0")
    windy = WindFieldWrapper(env, augment_obs=True, **WIND4)
    assert "wind_u" in windy.observation_space.spaces and windy.observation_space["wind_v"].shape == (1,)
    from bluesky_gym_sasha_b200 import gym_compat
    assert isinstance(windy, gym_compat.Wrapper) and list(windy.observation_space.keys())[-2:] == ["wind_u", "wind_v"]
    obs, info = windy.reset()
    assert list(obs)[-2:] == ["wind_u", "wind_v"] and obs["wind_u"].dtype == np.float64
    tas0 = obs["airspeed"][0]
    for _ in range(3):
        obs, r, te, tr, info = windy.step(np.zeros(2))
    assert 0.05 < np.hypot(obs["wind_u"][0], obs["wind_v"][0]) < 0.5          # 12-18 m/s of wind / MAX_WIND
    windy.close()
    # vector env: the observation layout is fixed at construction
    v = BlueSkyVectorEnv("MergeEnv-v0", 4, seed=0)
    with pytest.raises(ValueError):
        WindFieldWrapper(v, augment_obs=True, **WIND4)
    WindFieldWrapper(v, augment_obs=False, **WIND4)
    v.reset()
    kin0 = v.t["kin"].cpu().numpy().copy()
    v.step(np.zeros((4, 2), dtype=np.float32))
    gs = v._wind_t["gs"].cpu().numpy()
    tas = v.t["kin"].cpu().numpy()[..., 1]
    assert np.all(np.abs(np.hypot(gs[..., 0], gs[..., 1])[:, :20] - tas[:, :20]) > 1.0)   # ground speed != airspeed now
    v.set_wind()                                                                            # off again
    v.close()


@pytest.mark.gpu
def test_wind_altitude_profile_matches_oracle_windfield(cuda):
    """3-D field (per-point altitude profile): the device's wind_u / wind_v against oracle/windfield.py evaluated at
    the device's own aircraft state, over a descent through the profile."""
    from bluesky_gym_sasha_b200.vector_env import BlueSkyVectorEnv
    from oracle.windfield import Windfield
    alt = np.array([0.0, 1500.0, 3000.0, 6000.0])
    vn = np.array([[2.0, 4.0, -3.0], [8.0, 10.0, 5.0], [15.0, 12.0, 9.0], [30.0, 25.0, 20.0]])
    ve = np.array([[1.0, -2.0, 3.0], [-4.0, 6.0, 2.0], [7.0, -9.0, 5.0], [10.0, 12.0, -14.0]])
    lat, lon = np.array([51.8, 52.3, 52.0]), np.array([3.8, 4.2, 4.6])
    wf = Windfield()
    wf.addpointvne(lat, lon, vn, ve, alt)
    E = 16
    v = BlueSkyVectorEnv("DescentEnv-v0", E, seed=2, autoreset_mode="disabled", max_episode_steps=0,
                         wind=dict(lat=lat, lon=lon, vnorth=vn, veast=ve, alt=alt), wind_obs=True)
    v.reset()
    rng = np.random.default_rng(0)
    for step in range(25):
        obs, *_ = v.step(rng.uniform(-1, 0.2, (E, 1)).astype(np.float32))
        pos = v.t["pos"].cpu().numpy()[:, 0]
        kin = v.t["kin"].cpu().numpy()[:, 0].astype(np.float64)
        wn, we = wf.getdata(pos[:, 0], pos[:, 1], kin[:, 0])
        h = np.radians(kin[:, 2])
        np.testing.assert_allclose(obs["wind_u"][:, 0], (wn * np.cos(h) + we * np.sin(h)) / 50.0, atol=2e-4)
        np.testing.assert_allclose(obs["wind_v"][:, 0], (-wn * np.sin(h) + we * np.cos(h)) / 50.0, atol=2e-4)
    assert kin[:, 0].min() < 2500.0 and kin[:, 0].max() > 500.0          # the profile was actually traversed
    v.close()
