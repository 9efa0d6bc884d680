#include "host_pack.h"

#include <immintrin.h>
#include <sched.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <thread>
#include <vector>

namespace ldpc535 {
namespace {
void pack_range(const float *__restrict src, float *__restrict dst, size_t n)
{
    size_t i = 0;
#if defined(__AVX2__)
    // head up to a 32-byte boundary of dst, then 8 reals per step with non-temporal stores
    // (the staging buffer is written once and read only by the copy engine: skip the
    // read-for-ownership of every destination line)
    while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 31)) { dst[i] = src[2 * i]; i++; }
    for (; i + 8 <= n; i += 8) {
        const __m256 a = _mm256_loadu_ps(src + 2 * i);        // r0 i0 r1 i1 | r2 i2 r3 i3
        const __m256 b = _mm256_loadu_ps(src + 2 * i + 8);    // r4 i4 r5 i5 | r6 i6 r7 i7
        const __m256 s = _mm256_shuffle_ps(a, b, 0x88);       // r0 r1 r4 r5 | r2 r3 r6 r7
        const __m256d q = _mm256_permute4x64_pd(_mm256_castps_pd(s), 0xD8);
        _mm256_stream_ps(dst + i, _mm256_castpd_ps(q));
    }
    _mm_sfence();
#endif
    for (; i < n; i++) dst[i] = src[2 * i];
}
}  // namespace

int host_sharing_ranks()
{
    for (const char *name : {"LOCAL_WORLD_SIZE", "WORLD_SIZE"})
        if (const char *s = std::getenv(name)) return std::max(1, std::atoi(s));
    return 1;
}

int default_pack_threads()
{
    if (const char *s = std::getenv("LDPC535_PACK_THREADS")) {
        const int v = std::atoi(s);
        if (v >= 1) return std::min(v, 64);
    }
    // all the cores this process may run on, shared between the ranks of a multi-GPU job
    // (one process per GPU: LOCAL_WORLD_SIZE / WORLD_SIZE as torch.distributed.run sets them)
    unsigned hc = std::thread::hardware_concurrency();
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) hc = (unsigned)CPU_COUNT(&set);
    const unsigned ranks = (unsigned)host_sharing_ranks();
    return (int)std::max(1u, std::min(32u, hc / ranks));
}

void pack_real_parts(const float *src, float *dst, size_t n, int threads)
{
    const size_t kMinPerThread = 1u << 16;
    int t = (int)std::min<size_t>((size_t)std::max(threads, 1), std::max<size_t>(1, n / kMinPerThread));
    if (t <= 1) { pack_range(src, dst, n); return; }
    std::vector<std::thread> pool;
    pool.reserve(t - 1);
    const size_t per = (((n + t - 1) / t) + 15) & ~(size_t)15;
    for (int k = 1; k < t; k++) {
        const size_t lo = std::min(n, per * k), hi = std::min(n, per * (k + 1));
        if (hi > lo) pool.emplace_back(pack_range, src + 2 * lo, dst + lo, hi - lo);
    }
    pack_range(src, dst, std::min(n, per));
    for (auto &th : pool) th.join();
}
}  // namespace ldpc535
