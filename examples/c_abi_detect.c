/* The C ABI from a C host (no Python, no torch): state-based conflict detection for one airspace.
 *
 *   gcc examples/c_abi_detect.c -Iinclude -I/usr/local/cuda/include -Lbluesky_gym_sasha_b200 -lbsg_b200 \
 *       -L/usr/local/cuda/lib64 -lcudart -lm -Wl,-rpath,$PWD/bluesky_gym_sasha_b200 -o /tmp/c_abi_detect && /tmp/c_abi_detect 20000
 *
 * Packs N synthetic aircraft in a spatially coherent order (bsg_cd_pack_ordered), runs the culled + symmetric detection
 * (bsg_cd_detect_culled) and the plain all-pairs form (bsg_cd_detect) and checks that both find the same conflicts:
 * what a BlueSky ConflictDetection plug-in written in C / C++ would call (include/bsg.h). */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "bsg.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)
#define BK(x) do { int r_ = (x); if (r_ != BSG_OK) { fprintf(stderr, "%s: %s\n", #x, bsg_last_error()); return 3; } } while (0)

static double urand(unsigned long long *s) {            /* xorshift64*: uniform in [0, 1) */
    *s ^= *s >> 12; *s ^= *s << 25; *s ^= *s >> 27;
    return (double)((*s * 2685821657736338717ULL) >> 11) / 9007199254740992.0;
}

int main(int argc, char **argv) {
    const long long n = argc > 1 ? atoll(argv[1]) : 20000;
    if (bsg_device_count() < 1) { fprintf(stderr, "no CUDA device: %s\n", bsg_last_error()); return 1; }
    const size_t nb = (size_t)n * sizeof(double);
    double *h[6], *d[6];
    unsigned long long seed = 88172645463325252ULL;
    for (int k = 0; k < 6; ++k) h[k] = (double *)malloc(nb);
    for (long long i = 0; i < n; ++i) {
        h[0][i] = 52.0 + 20.0 * (urand(&seed) - 0.5);                       /* lat  */
        h[1][i] = 4.0 + 20.0 * (urand(&seed) - 0.5);                        /* lon  */
        h[2][i] = 360.0 * urand(&seed);                                     /* trk  */
        h[3][i] = 150.0 + 100.0 * urand(&seed);                             /* gs   */
        h[4][i] = 304.8 * floor(10.0 + 30.0 * urand(&seed)) + 40.0 * (urand(&seed) - 0.5);   /* alt */
        h[5][i] = urand(&seed) < 0.8 ? 0.0 : 20.0 * (urand(&seed) - 0.5);   /* vs   */
    }
    for (int k = 0; k < 6; ++k) { CK(cudaMalloc((void **)&d[k], nb)); CK(cudaMemcpy(d[k], h[k], nb, cudaMemcpyHostToDevice)); }

    const long long n_pad = bsg_cd_padded(n);
    float *rec, *rec_plain, *tcpamax;
    int32_t *perm;
    uint32_t *nconf, *nlos;
    uint8_t *inconf;
    unsigned long long *npairs, counts[2][2];
    void *work_order, *work_cull;
    const long long wo = bsg_cd_order_workspace(n), wc = bsg_cd_cull_workspace(n, n);
    CK(cudaMalloc((void **)&rec, (size_t)n_pad * 8 * sizeof(float)));
    CK(cudaMalloc((void **)&rec_plain, (size_t)n_pad * 8 * sizeof(float)));
    CK(cudaMalloc((void **)&perm, (size_t)n * sizeof(int32_t)));
    CK(cudaMalloc((void **)&nconf, (size_t)n * 4)); CK(cudaMalloc((void **)&nlos, (size_t)n * 4));
    CK(cudaMalloc((void **)&tcpamax, (size_t)n * 4)); CK(cudaMalloc((void **)&inconf, (size_t)n));
    CK(cudaMalloc((void **)&npairs, 16)); CK(cudaMalloc(&work_order, (size_t)wo)); CK(cudaMalloc(&work_cull, (size_t)wc));
    bsg_cd_lists lists = {NULL, NULL, 0, NULL, 0, npairs};                  /* counts only */

    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float ms[2];
    /* (a) spatial order + pack, then culled + symmetric detection (defaults: rpz 5 NM, hpz 1000 ft, look-ahead 300 s) */
    for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(e0, 0));
        BK(bsg_cd_pack_ordered(d[0], d[1], d[2], d[3], d[4], d[5], n, 52.0, 4.0, rec, perm, work_order, wo, NULL));
        BK(bsg_cd_detect_culled(rec, n, 0, n, 0.0f, 0.0f, 0.0f, BSG_CD_SYMMETRIC, nconf, nlos, tcpamax, inconf, &lists, work_cull, wc, NULL));
        CK(cudaEventRecord(e1, 0));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms[0], e0, e1));
    }
    CK(cudaMemcpy(counts[0], npairs, 16, cudaMemcpyDeviceToHost));
    /* (b) caller's order, every ordered pair */
    for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(e0, 0));
        BK(bsg_cd_pack(d[0], d[1], d[2], d[3], d[4], d[5], n, 52.0, 4.0, rec_plain, NULL));
        BK(bsg_cd_detect(rec_plain, n, 0, n, 0.0f, 0.0f, 0.0f, 0, nconf, nlos, tcpamax, inconf, &lists, NULL));
        CK(cudaEventRecord(e1, 0));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms[1], e0, e1));
    }
    CK(cudaMemcpy(counts[1], npairs, 16, cudaMemcpyDeviceToHost));
    printf("N = %lld (ABI v%d): culled + symmetric %.3f ms -> %llu conflicts, %llu LoS pairs; every ordered pair %.3f ms -> %llu, %llu\n",
           n, bsg_abi_version(), ms[0], counts[0][0], counts[0][1], ms[1], counts[1][0], counts[1][1]);
    if (counts[0][0] != counts[1][0] || counts[0][1] != counts[1][1]) { fprintf(stderr, "the two forms disagree\n"); return 4; }
    printf("ok\n");
    return 0;
}
