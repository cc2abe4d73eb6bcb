"""Philox4x32-10 counter RNG and the draw-stream convention shared with the CUDA reset kernels.

The reference draws its scenarios from the process-global ``np.random`` (e.g. horizontal_cr_env.py:130-132)
or stdlib ``random`` (merge_env.py:115-116); a batched device simulator cannot share one global stream,
so the product keys a Philox stream by (seed, global env id, episode).  This module is the CPU mirror of
``bluesky_gym_sasha_b200/csrc/rng.cuh`` so that an oracle env driven by ``PhiloxDraws`` makes exactly
the integer draws the device makes.  Known-answer vectors: Random123 ``kat_vectors`` (tests/test_oracle_philox.py).
Test infrastructure only (see oracle/__init__.py).
"""
import math
import random as _pyrandom

import numpy as np

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(ctr, key):
    c0, c1, c2, c3 = [int(x) & MASK for x in ctr]
    k0, k1 = [int(x) & MASK for x in key]
    for r in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c3 ^ k1) & MASK, p0 & MASK
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return c0, c1, c2, c3


class PhiloxDraws:
    """Sequential draws from the stream (seed, env_gid, episode); draw d = word d&3 of block d>>2.

    counter = (block, episode, stream_tag, seed_hi), key = (seed_lo, env_gid).
    """

    def __init__(self, seed, env_gid, episode, stream_tag=0):
        self.key = (seed & MASK, env_gid & MASK)
        self.seed_hi = (seed >> 32) & MASK
        self.episode = episode & MASK
        self.tag = stream_tag & MASK
        self.d = 0
        self._blk = None

    def u32(self):
        if (self.d & 3) == 0:
            self._blk = philox4x32_10((self.d >> 2, self.episode, self.tag, self.seed_hi), self.key)
        w = self._blk[self.d & 3]
        self.d += 1
        return w

    def randint(self, lo, hi):
        """Integer in [lo, hi): lo + mulhi(u32, hi - lo)."""
        return lo + ((self.u32() * (hi - lo)) >> 32)

    def u01(self):
        return (self.u32() >> 8) * (1.0 / 16777216.0)

    def uniform(self, a, b):
        return a + (b - a) * self.u01()

    def normal(self, mu, sigma):
        u1 = ((self.u32() >> 8) + 1) * (1.0 / 16777216.0)
        u2 = (self.u32() >> 8) * (1.0 / 16777216.0)
        return mu + sigma * math.sqrt(-2.0 * math.log(u1)) * math.cos(2.0 * math.pi * u2)


class GlobalNumpyDraws:
    """The reference's own source: the process-global ``np.random`` (seed it with np.random.seed)."""

    def randint(self, lo, hi):
        return int(np.random.randint(lo, hi))

    def uniform(self, a, b):
        return float(np.random.uniform(a, b))

    def normal(self, mu, sigma):
        return float(np.random.normal(mu, sigma))


class GlobalStdlibDraws(GlobalNumpyDraws):
    """merge_env.py:115-116,150-151 draw from stdlib ``random`` instead."""

    def uniform(self, a, b):
        return _pyrandom.uniform(a, b)


def noise_normals(seed, env_gid, call_index, n, terminal=False):
    """The N(0, 1) sequence the device observation-noise kernel (csrc/obs_noise.cu) adds to the n elements of env
    ``env_gid``'s observation at reset/step call number ``call_index``: block m of the Philox stream tagged 'NOIS'
    ('NOIT' for terminal observations), Box-Muller on words 0 and 1 -> elements 2m (cos) and 2m+1 (sin)."""
    tag = 0x4e4f4954 if terminal else 0x4e4f4953
    key = (seed & MASK, env_gid & MASK)
    out = []
    for m in range((n + 1) // 2):
        w = philox4x32_10((m, call_index & MASK, tag, (seed >> 32) & MASK), key)
        u1 = ((w[0] >> 8) + 1) * (1.0 / 16777216.0)
        u2 = (w[1] >> 8) * (1.0 / 16777216.0)
        r = math.sqrt(-2.0 * math.log(u1))
        out += [r * math.cos(2.0 * math.pi * u2), r * math.sin(2.0 * math.pi * u2)]
    return np.array(out[:n])
