// rng.cuh -- Philox4x32-10 counter RNG for the device reset kernels.
// The reference draws scenarios from process-global streams (np.random at horizontal_cr_env.py:130-132,
// stdlib random at merge_env.py:115-116); a batched simulator keys one stream per
// (seed, global env id, episode) instead, so results do not depend on how envs are sharded over GPUs.
// CPU mirror with the same draw convention: oracle/philox.py (known-answer tested in tests/).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bsg {

struct Philox {
    uint32_t k0, k1;        // key  = (seed_lo, env_gid)
    uint32_t ep, tag, hi;   // counter words 1..3 = (episode, stream_tag, seed_hi)

    __device__ __forceinline__ void block(uint32_t blk, uint32_t out[4]) const {
        uint32_t c0 = blk, c1 = ep, c2 = tag, c3 = hi, a = k0, b = k1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
            uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
            uint32_t n0 = h1 ^ c1 ^ a, n2 = h0 ^ c3 ^ b;
            c0 = n0; c1 = l1; c2 = n2; c3 = l0;
            a += 0x9E3779B9u; b += 0xBB67AE85u;
        }
        out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
    }
    // draw d of the stream = word d&3 of block d>>2 (random access: lanes fetch their own draws)
    __device__ __forceinline__ uint32_t u32(uint32_t d) const {
        uint32_t w[4];
        block(d >> 2, w);
        uint32_t s = d & 3u;
        return s == 0 ? w[0] : (s == 1 ? w[1] : (s == 2 ? w[2] : w[3]));
    }
    __device__ __forceinline__ int randint(uint32_t d, int lo, int hi_) const {
        return lo + (int)__umulhi(u32(d), (uint32_t)(hi_ - lo));
    }
    __device__ __forceinline__ double u01(uint32_t d) const { return (double)(u32(d) >> 8) * (1.0 / 16777216.0); }
    __device__ __forceinline__ double uniform(uint32_t d, double a, double b) const { return a + (b - a) * u01(d); }
    // Box-Muller on draws d, d+1
    __device__ __forceinline__ double normal(uint32_t d, double mu, double sigma) const {
        double u1 = (double)((u32(d) >> 8) + 1u) * (1.0 / 16777216.0);
        double u2 = (double)(u32(d + 1) >> 8) * (1.0 / 16777216.0);
        return mu + sigma * sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
    }
};

__device__ __forceinline__ Philox make_philox(uint64_t seed, int64_t env_gid, uint32_t episode, uint32_t tag = 0) {
    Philox p;
    p.k0 = (uint32_t)seed; p.k1 = (uint32_t)env_gid;
    p.ep = episode; p.tag = tag; p.hi = (uint32_t)(seed >> 32);
    return p;
}

}  // namespace bsg
