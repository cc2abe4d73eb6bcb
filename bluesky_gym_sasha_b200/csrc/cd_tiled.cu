// cd_tiled.cu -- K2: single-airspace state-based conflict detection, tiled all-pairs (n-body style).
// Replaces the dense N x N float64 matrices of bluesky/traffic/asas/statebased.py::StateBased.detect
// (upstream's optional single-threaded C++ twin is `cstatebased`); O(N) memory instead of ~15 N^2.
//
// Decomposition: a work item = (row block of 512 aircraft) x (group of 8 column tiles of 256 aircraft).
// A persistent grid of (SM count x resident CTAs) strides over the items, so the tail is < 1 item.
// Each thread keeps R = 2 own-rows in registers; the intruder tile (256 x 32 B = 8 KB) is staged in
// shared memory by the TMA engine (cp.async.bulk + mbarrier complete_tx, two stages) and read back
// with broadcast LDS.128, so every byte of column data is fetched from L2 once per CTA-tile and the
// inner loop is pure FP32-pipe work (bound: FP32 issue rate; MUFU is the co-limiter at 3 per pair).
// Rare events (conflict / LoS found) leave the hot loop through one predicated branch.
#include <cuda_runtime.h>
#include <stdint.h>

#include "bsg_internal.h"
#include "bsg_math.cuh"
#include "cd_pair.cuh"

namespace bsg {

constexpr int kTJ = 256;                 // columns per tile
constexpr int kNT = 256;                 // threads per CTA
constexpr int kR = 2;                    // rows per thread
constexpr int kRowsPerCta = kNT * kR;    // 512
constexpr int kTilesPerItem = 8;
constexpr uint32_t kTileBytes = kTJ * 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

struct CdArgs {
    const float4* rec;     // [n_pad * 2] (A, B) interleaved
    long long n_all, n_pad, row0, n_rows;
    float R2, hpz, dtlook;
    uint32_t* nconf_row;
    uint32_t* nlos_row;
    float* tcpamax;
    int32_t* pairs;
    long long cap;
    unsigned long long* npairs;
    int n_rowblocks, n_colgroups, n_tiles;
};

template <bool WRAP, bool DIAG>
__device__ __forceinline__ void cd_tile(const float4* __restrict__ tile, long long c0, const float4 (&Ai)[kR],
                                        const float4 (&Bi)[kR], const long long (&ri)[kR],
                                        const bool (&rvalid)[kR], const CdArgs& a, uint32_t (&nconf)[kR],
                                        uint32_t (&nlos)[kR], float (&tmax)[kR]) {
#pragma unroll 4
    for (int j = 0; j < kTJ; ++j) {
        const float4 Aj = tile[2 * j];
        const float4 Bj = tile[2 * j + 1];
#pragma unroll
        for (int k = 0; k < kR; ++k) {
            const bool same = DIAG ? (ri[k] == c0 + j) : false;
            CdPair p = cd_pair_eval<WRAP>(Ai[k], Bi[k], Aj, Bj, a.R2, a.hpz, a.dtlook, same);
            if ((p.conf | p.los) && rvalid[k]) {           // rare path
                if (p.los) nlos[k]++;
                if (p.conf) {
                    nconf[k]++;
                    tmax[k] = fmaxf(tmax[k], p.tcpa);
                    if (a.pairs) {
                        unsigned long long s = atomicAdd(a.npairs, 1ULL);
                        if ((long long)s < a.cap) {
                            a.pairs[2 * s] = (int32_t)ri[k];
                            a.pairs[2 * s + 1] = (int32_t)(c0 + j);
                        }
                    } else if (a.npairs) {
                        atomicAdd(a.npairs, 1ULL);
                    }
                }
                if (p.los && a.npairs) atomicAdd(a.npairs + 1, 1ULL);
            }
        }
    }
}

template <bool WRAP>
__global__ void __launch_bounds__(kNT, 2) cd_tiled_kernel(const CdArgs a) {
    __shared__ __align__(128) float4 s_tile[2][kTJ * 2];
    __shared__ __align__(8) uint64_t s_full[2];

    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(&s_full[0], 1);
        mbar_init(&s_full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t parity[2] = {0u, 0u};

    const long long n_items = (long long)a.n_rowblocks * a.n_colgroups;
    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int rb = (int)(item / a.n_colgroups);
        const int cg = (int)(item % a.n_colgroups);
        const int t_begin = cg * kTilesPerItem;
        const int t_end = min(t_begin + kTilesPerItem, a.n_tiles);

        // own rows -> registers (clamped loads for the ragged last block; results discarded)
        float4 Ai[kR], Bi[kR];
        long long ri[kR];
        bool rvalid[kR];
        uint32_t nconf[kR], nlos[kR];
        float tmax[kR];
        const long long rbase = a.row0 + (long long)rb * kRowsPerCta;
#pragma unroll
        for (int k = 0; k < kR; ++k) {
            long long r = rbase + tid + k * kNT;
            rvalid[k] = r < a.row0 + a.n_rows;
            ri[k] = rvalid[k] ? r : (a.row0 + a.n_rows - 1);
            Ai[k] = __ldg(&a.rec[2 * ri[k]]);
            Bi[k] = __ldg(&a.rec[2 * ri[k] + 1]);
            nconf[k] = 0; nlos[k] = 0; tmax[k] = 0.0f;
        }

        if (tid == 0) {        // prologue: first tile of the item
            mbar_expect_tx(&s_full[t_begin & 1], kTileBytes);
            tma_load_1d(&s_tile[t_begin & 1][0], a.rec + 2LL * t_begin * kTJ, kTileBytes, &s_full[t_begin & 1]);
        }
        for (int t = t_begin; t < t_end; ++t) {
            const int s = t & 1;
            if (tid == 0 && t + 1 < t_end) {     // prefetch the next tile into the other stage
                mbar_expect_tx(&s_full[s ^ 1], kTileBytes);
                tma_load_1d(&s_tile[s ^ 1][0], a.rec + 2LL * (t + 1) * kTJ, kTileBytes, &s_full[s ^ 1]);
            }
            mbar_wait(&s_full[s], parity[s]);
            parity[s] ^= 1u;
            const long long c0 = (long long)t * kTJ;
            const bool diag = (c0 < rbase + kRowsPerCta) && (c0 + kTJ > rbase);
            if (diag)
                cd_tile<WRAP, true>(s_tile[s], c0, Ai, Bi, ri, rvalid, a, nconf, nlos, tmax);
            else
                cd_tile<WRAP, false>(s_tile[s], c0, Ai, Bi, ri, rvalid, a, nconf, nlos, tmax);
            __syncthreads();     // stage s may be overwritten by the prefetch of iteration t+1
        }
#pragma unroll
        for (int k = 0; k < kR; ++k) {
            if (rvalid[k]) {
                const long long o = ri[k] - a.row0;
                if (nconf[k]) {
                    atomicAdd(&a.nconf_row[o], nconf[k]);
                    if (a.tcpamax) atomicMax((int*)&a.tcpamax[o], __float_as_int(tmax[k]));
                }
                if (nlos[k] && a.nlos_row) atomicAdd(&a.nlos_row[o], nlos[k]);
            }
        }
    }
}

__global__ void cd_finalize_kernel(const uint32_t* __restrict__ nconf_row, uint8_t* __restrict__ inconf, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) inconf[i] = nconf_row[i] > 0;
}

// float64 SoA -> 32-byte float record (see cd_pair.cuh); padding rows are inert aircraft.
__global__ void cd_pack_kernel(const double* __restrict__ lat, const double* __restrict__ lon,
                               const double* __restrict__ trk, const double* __restrict__ gs,
                               const double* __restrict__ alt, const double* __restrict__ vs, long long n,
                               long long n_pad, double lat0, double lon0, float4* __restrict__ rec) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    float4 A, B;
    if (i < n) {
        double la = lat[i];
        double dl = fmod((lon[i] - lon0) + 180.0, 360.0);
        if (dl < 0.0) dl += 360.0;
        dl -= 180.0;
        double s, c, st, ct;
        sincos(la * (0.5 * kDeg2RadD), &s, &c);
        sincos(trk[i] * kDeg2RadD, &st, &ct);
        A = make_float4((float)(kRearthD * kDeg2RadD * dl), (float)(kRearthD * kDeg2RadD * (la - lat0)), (float)c, (float)s);
        B = make_float4((float)(gs[i] * st), (float)(gs[i] * ct), (float)alt[i], (float)vs[i]);
    } else {
        A = make_float4(0.0f, 0.0f, 1.0f, 0.0f);
        B = make_float4(0.0f, 0.0f, 3.0e9f, 0.0f);     // |dalt| ~ 3e9: never in conflict nor in LoS
    }
    rec[2 * i] = A;
    rec[2 * i + 1] = B;
}

}  // namespace bsg

using namespace bsg;

extern "C" int64_t bsg_cd_padded(int64_t n) { return ((n + kTJ - 1) / kTJ) * kTJ; }

extern "C" int bsg_cd_pack(const double* d_lat, const double* d_lon, const double* d_trk, const double* d_gs,
                           const double* d_alt, const double* d_vs, int64_t n, double lat0, double lon0,
                           float* d_rec, void* stream) {
    if (n < 0 || (n > 0 && (!d_lat || !d_lon || !d_trk || !d_gs || !d_alt || !d_vs)) || !d_rec)
        return bsg_fail(BSG_EINVAL, "bsg_cd_pack: null pointer or negative n");
    int64_t n_pad = bsg_cd_padded(n);
    if (n_pad == 0) return BSG_OK;
    int blocks = (int)((n_pad + 255) / 256);
    cd_pack_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_lat, d_lon, d_trk, d_gs, d_alt, d_vs, n, n_pad,
                                                              lat0, lon0, (float4*)d_rec);
    return bsg_cuda_check(cudaGetLastError(), "bsg_cd_pack launch");
}

extern "C" int bsg_cd_detect(const float* d_rec, int64_t n_all, int64_t row0, int64_t n_rows, float rpz, float hpz,
                             float dtlookahead, uint32_t flags, uint32_t* d_nconf_row, uint32_t* d_nlos_row,
                             float* d_tcpamax, uint8_t* d_inconf, int32_t* d_pairs, int64_t cap,
                             unsigned long long* d_npairs, void* stream) {
    if (n_all < 0 || row0 < 0 || n_rows < 0 || row0 + n_rows > n_all)
        return bsg_fail(BSG_EINVAL, "bsg_cd_detect: row range outside [0, n_all)");
    if (n_all > 0x7fffffffLL) return bsg_fail(BSG_EINVAL, "bsg_cd_detect: n_all exceeds int32 pair indices");
    if (n_rows > 0 && (!d_rec || !d_nconf_row)) return bsg_fail(BSG_EINVAL, "bsg_cd_detect: null d_rec / d_nconf_row");
    if (d_pairs && (!d_npairs || cap < 0)) return bsg_fail(BSG_EINVAL, "bsg_cd_detect: pair list needs d_npairs and cap >= 0");
    if (flags & BSG_CD_SYMMETRIC) return bsg_fail(BSG_EINVAL, "bsg_cd_detect: BSG_CD_SYMMETRIC not implemented yet");
    cudaStream_t st = (cudaStream_t)stream;
    if (d_npairs) BSG_CUDA(cudaMemsetAsync(d_npairs, 0, 2 * sizeof(unsigned long long), st));
    if (n_rows == 0) return BSG_OK;
    BSG_CUDA(cudaMemsetAsync(d_nconf_row, 0, sizeof(uint32_t) * n_rows, st));
    if (d_nlos_row) BSG_CUDA(cudaMemsetAsync(d_nlos_row, 0, sizeof(uint32_t) * n_rows, st));
    if (d_tcpamax) BSG_CUDA(cudaMemsetAsync(d_tcpamax, 0, sizeof(float) * n_rows, st));

    CdArgs a;
    a.rec = (const float4*)d_rec;
    a.n_all = n_all; a.n_pad = bsg_cd_padded(n_all); a.row0 = row0; a.n_rows = n_rows;
    if (rpz <= 0.0f) rpz = 5.0f * 1852.0f;
    if (hpz <= 0.0f) hpz = 1000.0f * 0.3048f;
    if (dtlookahead <= 0.0f) dtlookahead = 300.0f;
    a.R2 = rpz * rpz; a.hpz = hpz; a.dtlook = dtlookahead;
    a.nconf_row = d_nconf_row; a.nlos_row = d_nlos_row; a.tcpamax = d_tcpamax;
    a.pairs = d_pairs; a.cap = cap; a.npairs = d_npairs;
    a.n_tiles = (int)(a.n_pad / kTJ);
    a.n_rowblocks = (int)((n_rows + kRowsPerCta - 1) / kRowsPerCta);
    a.n_colgroups = (a.n_tiles + kTilesPerItem - 1) / kTilesPerItem;

    int dev = 0, sms = 0, occ = 0;
    BSG_CUDA(cudaGetDevice(&dev));
    BSG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const bool wrap = (flags & BSG_CD_LON_WRAP) != 0;
    if (wrap) BSG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, cd_tiled_kernel<true>, kNT, 0));
    else BSG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, cd_tiled_kernel<false>, kNT, 0));
    if (occ < 1) occ = 1;
    long long n_items = (long long)a.n_rowblocks * a.n_colgroups;
    int grid = (int)((n_items < (long long)sms * occ) ? n_items : (long long)sms * occ);
    if (wrap) cd_tiled_kernel<true><<<grid, kNT, 0, st>>>(a);
    else cd_tiled_kernel<false><<<grid, kNT, 0, st>>>(a);
    BSG_CUDA(cudaGetLastError());
    if (d_inconf) {
        cd_finalize_kernel<<<(int)((n_rows + 255) / 256), 256, 0, st>>>(d_nconf_row, d_inconf, n_rows);
        BSG_CUDA(cudaGetLastError());
    }
    return BSG_OK;
}
