#include "host_pack.h"

#include <immintrin.h>
#include <sched.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <mutex>
#include <thread>
#include <vector>

namespace ldpc535 {
namespace {

void pack_scalar(const float *__restrict src, float *__restrict dst, size_t n)
{
    for (size_t i = 0; i < n; i++) dst[i] = src[2 * i];
}

// 8 reals per step with non-temporal stores: the staging buffer is written once and read only
// by the copy engine, so the read-for-ownership of every destination line is skipped.
// Compiled for AVX2 by attribute (the rest of the file is baseline x86-64) and only called
// after __builtin_cpu_supports("avx2") said yes.
__attribute__((target("avx2"))) void pack_avx2(const float *__restrict src, float *__restrict dst, size_t n)
{
    size_t i = 0;
    while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 31)) { dst[i] = src[2 * i]; i++; }
    for (; i + 8 <= n; i += 8) {
        const __m256 a = _mm256_loadu_ps(src + 2 * i);        // r0 i0 r1 i1 | r2 i2 r3 i3
        const __m256 b = _mm256_loadu_ps(src + 2 * i + 8);    // r4 i4 r5 i5 | r6 i6 r7 i7
        const __m256 s = _mm256_shuffle_ps(a, b, 0x88);       // r0 r1 r4 r5 | r2 r3 r6 r7
        const __m256d q = _mm256_permute4x64_pd(_mm256_castps_pd(s), 0xD8);
        _mm256_stream_ps(dst + i, _mm256_castpd_ps(q));
    }
    _mm_sfence();
    for (; i < n; i++) dst[i] = src[2 * i];
}

using pack_fn = void (*)(const float *, float *, size_t);
pack_fn pick_pack()
{
    __builtin_cpu_init();
    return __builtin_cpu_supports("avx2") ? pack_avx2 : pack_scalar;
}
const pack_fn g_pack = pick_pack();

}  // namespace

void pack_real_parts_serial(const float *src, float *dst, size_t n) { g_pack(src, dst, n); }

int default_pack_threads()
{
    if (const char *s = std::getenv("LDPC535_PACK_THREADS")) {
        const int v = std::atoi(s);
        if (v >= 1) return std::min(v, 64);
    }
    unsigned hc = std::thread::hardware_concurrency();
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) hc = (unsigned)CPU_COUNT(&set);
    return (int)std::max(1u, std::min(32u, hc));
}

// The team: workers take blocks of 64 Ki symbols from a shared counter (no thread finishes long
// after the others), spin for ~100 us between jobs before they sleep on the condition variable --
// decode_batch packs a 128 MiB chunk every ~1.8 ms, and a sleeping worker's wake-up (tens of
// microseconds, 38 times per 5.12 GB) was part of every chunk -- and the caller spins the same way
// for the last block to finish.
struct PackPool::Impl {
    static constexpr size_t kBlock = (size_t)1 << 16;       // symbols per block
    static constexpr int kSpin = 4000;                       // pause iterations before sleeping (~50-150 us)
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::vector<std::thread> workers;
    // one job at a time, published by `generation` (release) and read after it (acquire)
    const float *src = nullptr;
    float *dst = nullptr;
    size_t n = 0;
    int parts = 0;
    std::atomic<size_t> next{0};                             // next block to be taken
    std::atomic<unsigned long> generation{0};
    std::atomic<int> pending{0};                             // workers still on the current job
    std::atomic<bool> stop{false};

    void run_blocks()
    {
        for (;;) {
            const size_t lo = next.fetch_add(kBlock, std::memory_order_relaxed);
            if (lo >= n) return;
            const size_t hi = std::min(n, lo + kBlock);
            g_pack(src + 2 * lo, dst + lo, hi - lo);
        }
    }

    void worker(int id)
    {
        unsigned long seen = 0;
        for (;;) {
            unsigned long g = seen;
            for (int i = 0; i < kSpin; i++) {
                g = generation.load(std::memory_order_acquire);
                if (g != seen || stop.load(std::memory_order_relaxed)) break;
                _mm_pause();
            }
            if (g == seen && !stop.load(std::memory_order_relaxed)) {
                std::unique_lock<std::mutex> lk(mu);
                cv_work.wait(lk, [&] { return stop.load() || generation.load(std::memory_order_acquire) != seen; });
                g = generation.load(std::memory_order_acquire);
            }
            if (stop.load()) return;
            seen = g;
            if (id < parts) {
                run_blocks();
                if (pending.fetch_sub(1, std::memory_order_acq_rel) == 1) {
                    std::lock_guard<std::mutex> lk(mu);
                    cv_done.notify_one();
                }
            }
        }
    }
};

PackPool::PackPool(int threads) : impl_(new Impl), n_threads_(std::max(1, threads))
{
    for (int id = 1; id < n_threads_; id++) impl_->workers.emplace_back(&Impl::worker, impl_, id);
}

PackPool::~PackPool()
{
    {
        std::lock_guard<std::mutex> lk(impl_->mu);
        impl_->stop.store(true);
    }
    impl_->cv_work.notify_all();
    for (auto &t : impl_->workers) t.join();
    delete impl_;
}

void PackPool::pack(const float *src, float *dst, size_t n)
{
    const size_t kMinPerThread = 1u << 16;
    const int parts = (int)std::min<size_t>((size_t)n_threads_, std::max<size_t>(1, n / kMinPerThread));
    if (parts <= 1) { g_pack(src, dst, n); return; }
    Impl &s = *impl_;
    {
        std::lock_guard<std::mutex> lk(s.mu);
        s.src = src; s.dst = dst; s.n = n;
        s.parts = parts;
        s.next.store(0, std::memory_order_relaxed);
        s.pending.store(parts - 1, std::memory_order_relaxed);
        s.generation.fetch_add(1, std::memory_order_release);
    }
    s.cv_work.notify_all();
    s.run_blocks();
    for (int i = 0; i < Impl::kSpin; i++) {
        if (s.pending.load(std::memory_order_acquire) == 0) return;
        _mm_pause();
    }
    std::unique_lock<std::mutex> lk(s.mu);
    s.cv_done.wait(lk, [&] { return s.pending.load(std::memory_order_acquire) == 0; });
}

}  // namespace ldpc535
