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
    env.close()
    ref.close()
