// host_pool.cu -- the host half of the end-to-end step: a small persistent thread pool that copies the
// observation block out of the pinned mirror into the caller's (pageable) result array.  One core moves
// ~18 GB/s, so a 1.7 MB batched observation costs more than the env kernel itself when copied by the
// Python thread alone.  Pure host code (no device work happens here).
//
// Design notes (measured on the B200 box): a worker that slept on a condition variable for the ~120 us of
// a step needs 50-100 us to wake, longer than the copy it is wanted for.  So (1) bsg_step_host_copy calls
// host_pool_prewake() BEFORE it launches the kernel: the workers wake while the GPU is busy and poll for
// work for a bounded time (kSpinUs), then go back to sleep; (2) a copy is cut into 64 KB pieces handed out
// through an atomic counter to whoever is awake -- the calling thread included -- so nobody ever waits for
// a late sleeper, and the call returns as soon as every piece is done.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "bsg_internal.h"

namespace bsg {

namespace {
constexpr size_t kPiece = 64u << 10;
constexpr int kSlots = 8;
constexpr int kSpinUs = 400;

struct Job {
    char* dst = nullptr;
    const char* src = nullptr;
    size_t n = 0;               // source bytes
    int pieces = 0;
    bool widen = false;         // float32 -> float64 conversion instead of a plain copy
    std::atomic<int> next{0}, done{0}, active{0};
};

inline void work_on(Job& j) {
    for (;;) {
        int c = j.next.fetch_add(1, std::memory_order_acq_rel);
        if (c >= j.pieces) return;
        size_t lo = (size_t)c * kPiece, len = j.n - lo < kPiece ? j.n - lo : kPiece;
        if (j.widen) {
            const float* s = (const float*)(j.src + lo);
            double* d = (double*)(j.dst + 2 * lo);
            for (size_t i = 0, m = len / sizeof(float); i < m; ++i) d[i] = (double)s[i];
        } else {
            memcpy(j.dst + lo, j.src + lo, len);
        }
        j.done.fetch_add(1, std::memory_order_acq_rel);
    }
}

inline int64_t now_us() {
    return std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct Pool {
    std::mutex mu;
    std::condition_variable cv;
    std::vector<std::thread> workers;
    bool stop = false;
    std::atomic<uint64_t> wake_gen{0};      // bumped by prewake() and by every job
    std::atomic<uint64_t> job_gen{0};       // job g lives in slot[g % kSlots]
    Job slot[kSlots];

    void run() {
        uint64_t seen_wake = 0, seen_job = job_gen.load();
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return stop || wake_gen.load(std::memory_order_acquire) != seen_wake; });
                if (stop) return;
                seen_wake = wake_gen.load(std::memory_order_acquire);
            }
            int64_t deadline = now_us() + kSpinUs;
            int polls = 0;
            for (;;) {
                uint64_t g = job_gen.load(std::memory_order_acquire);
                if (g != seen_job) {
                    seen_job = g;
                    Job& j = slot[g % kSlots];
                    j.active.fetch_add(1, std::memory_order_acq_rel);
                    if (job_gen.load(std::memory_order_acquire) == g) work_on(j);   // (slot not recycled meanwhile)
                    j.active.fetch_sub(1, std::memory_order_acq_rel);
                    deadline = now_us() + kSpinUs;
                    continue;
                }
#if defined(__x86_64__)
                __builtin_ia32_pause();
#endif
                if ((++polls & 63) == 0 && now_us() > deadline) break;
            }
        }
    }
    explicit Pool(int w) {
        for (int k = 0; k < w; ++k) workers.emplace_back([this] { run(); });
    }
    ~Pool() {
        { std::lock_guard<std::mutex> lk(mu); stop = true; }
        cv.notify_all();
        for (auto& t : workers) t.join();
    }
    void wake() {
        { std::lock_guard<std::mutex> lk(mu); wake_gen.fetch_add(1, std::memory_order_acq_rel); }
        cv.notify_all();
    }
};

Pool* g_pool = nullptr;
std::once_flag g_once;
std::atomic<bool> g_busy{false};

void init_pool() {
    // Copiers (workers + the calling thread).  Measured at E = 4096 on a 16-core host (scripts/e2e_variants.py): 4 copiers
    // 160 us per step, 8: 151 us, 12: 147 us -- the tail of the copy that cannot overlap the transfer shrinks with the
    // number of copiers.  Default: three quarters of this rank's share of the cores, between 2 and 12.  One process per
    // GPU (torchrun): the ranks of a node share its cores and the workers poll while they wait, so the share is
    // cores / LOCAL_WORLD_SIZE (8 ranks on a 32-core host: 3 copiers each; 2 measured 5 % better than 4 there,
    // scripts/e2e_ranks.py).  BSG_HOST_THREADS overrides.
    unsigned hc = std::thread::hardware_concurrency();
    int ranks = 1;
    if (const char* e = getenv("LOCAL_WORLD_SIZE")) ranks = atoi(e);
    if (ranks < 1) ranks = 1;
    int copiers = hc ? (int)(hc * 3u / 4u) / ranks : 4;
    if (copiers < 2) copiers = 2;
    if (copiers > 12) copiers = 12;
    if (const char* e = getenv("BSG_HOST_THREADS")) copiers = atoi(e);
    if (hc && copiers > (int)hc) copiers = (int)hc;
    int w = copiers - 1;
    if (w < 0) w = 0;
    if (w > 15) w = 15;
    if (w > 0) g_pool = new Pool(w);             // lives for the process; idle workers sleep on a condition variable
}
}  // namespace

// Wakes the workers so that they are polling by the time the next host_copy_mt() arrives.
void host_pool_prewake() {
    std::call_once(g_once, init_pool);
    if (g_pool) g_pool->wake();
}

// memcpy(dst, src, n) -- or, with widen, dst[i] = (double)src[i] over n bytes of float32 -- shared between the calling
// thread and whichever pool workers are awake.
void host_copy_mt(void* dst, const void* src, size_t n, bool widen) {
    std::call_once(g_once, init_pool);
    bool expected = false;
    if (!g_pool || n < 4 * kPiece || !g_busy.compare_exchange_strong(expected, true)) {
        if (widen) {                             // small, or another handle's thread is using the pool
            const float* s = (const float*)src;
            double* d = (double*)dst;
            for (size_t i = 0, m = n / sizeof(float); i < m; ++i) d[i] = (double)s[i];
        } else {
            memcpy(dst, src, n);
        }
        return;
    }
    Pool& p = *g_pool;
    const uint64_t g = p.job_gen.load(std::memory_order_acquire) + 1;
    Job& j = p.slot[g % kSlots];
    while (j.active.load(std::memory_order_acquire) != 0) { /* a straggler from kSlots jobs ago (never in practice) */ }
    j.dst = (char*)dst; j.src = (const char*)src; j.n = n; j.pieces = (int)((n + kPiece - 1) / kPiece); j.widen = widen;
    j.done.store(0, std::memory_order_relaxed);
    j.next.store(0, std::memory_order_release);
    p.job_gen.store(g, std::memory_order_release);
    work_on(j);
    while (j.done.load(std::memory_order_acquire) < j.pieces) { /* pieces in flight on other cores */ }
    g_busy.store(false, std::memory_order_release);
}

}  // namespace bsg
