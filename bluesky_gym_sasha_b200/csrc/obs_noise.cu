// obs_noise.cu -- NoisyObservationWrapper on the device (reference: bluesky_gym/wrappers/uncertainty.py:4-31):
// every element of the observation returned by reset() / step() gets independent Gaussian noise
// N(0, noise_level); rewards, termination and the simulator state are untouched.
//
// The reference draws from the process-global np.random; a batched simulator keys a Philox4x32-10 stream per
// (seed, global env id, call index) instead (tag 'NOIS'), so the noise does not depend on how envs are sharded.
// Element pair (2m, 2m+1) of an env's observation uses block m of that stream: Box-Muller on words 0 and 1 gives
// two normals (cos and sin branch).  The terminal observations kept aside by same-step autoreset are noised
// from the stream with tag 'NOIT' (SB3 stores the wrapper's noisy terminal observation too).
// One thread per element pair; HBM-bound (reads and writes the observation block once).
#include <cuda_runtime.h>
#include <stdint.h>

#include "env_kernels.cuh"

namespace bsg {

constexpr uint32_t kTagNoise = 0x4e4f4953u, kTagNoiseTerminal = 0x4e4f4954u;

__device__ __forceinline__ void normal_pair(const Philox& ph, uint32_t m, float sigma, float& z0, float& z1) {
    uint32_t w[4];
    ph.block(m, w);
    float u1 = (float)((w[0] >> 8) + 1u) * (1.0f / 16777216.0f);
    float u2 = (float)(w[1] >> 8) * (1.0f / 16777216.0f);
    float r = sigma * sqrtf(-2.0f * logf(u1));
    float s, c;
    sincospif(2.0f * u2, &s, &c);
    z0 = r * c; z1 = r * s;
}

__global__ void __launch_bounds__(256) obs_noise_kernel(float* __restrict__ obs, const int32_t* __restrict__ row_env,
                                                         const int32_t* __restrict__ n_rows_dev, long long n_rows, int obs_dim,
                                                         const uint8_t* __restrict__ mask, float sigma, uint64_t seed,
                                                         long long gid0, uint32_t call, uint32_t tag) {
    const int half = (obs_dim + 1) >> 1;
    if (n_rows_dev) n_rows = *n_rows_dev;                 // compacted terminal observations: count lives on the device
    const long long n = n_rows * half;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const long long row = t / half;
        const int m = (int)(t - row * half);
        const long long e = row_env ? row_env[row] : row;
        if (mask && !mask[e]) continue;
        Philox ph = make_philox(seed, gid0 + e, call, tag);
        float z0, z1;
        normal_pair(ph, (uint32_t)m, sigma, z0, z1);
        float* o = obs + row * obs_dim + 2 * m;
        o[0] += z0;
        if (2 * m + 1 < obs_dim) o[1] += z1;
    }
}

}  // namespace bsg

// obs: [E, obs_dim] in place.  final_obs / final_ids / final_count: the compacted terminal rows of this step (may be null).
int bsg_launch_obs_noise(const bsg::EnvParams& P, float sigma, uint32_t call, bool with_final, cudaStream_t st) {
    const int half = (P.obs_dim + 1) / 2;
    long long n = (long long)P.E * half;
    if (n == 0) return BSG_OK;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    bsg::obs_noise_kernel<<<blocks, 256, 0, st>>>(P.obs, nullptr, nullptr, P.E, P.obs_dim, P.reset_mask, sigma, P.seed, P.gid0,
                                                   call, bsg::kTagNoise);
    if (with_final && P.final_obs && P.final_ids && P.final_count)
        bsg::obs_noise_kernel<<<blocks, 256, 0, st>>>(P.final_obs, P.final_ids, P.final_count + P.fc_slot, 0, P.obs_dim, nullptr, sigma,
                                                       P.seed, P.gid0, call, bsg::kTagNoiseTerminal);
    return bsg_cuda_check(cudaGetLastError(), "obs_noise_kernel launch");
}
