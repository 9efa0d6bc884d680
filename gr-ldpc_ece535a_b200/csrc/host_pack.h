// Host-side staging helper of the host-buffer decode path: copy the REAL parts of a span of
// gr_complex symbols into (pinned) staging memory with a few worker threads.  The decoder
// never reads the imaginary parts (lib/ldpc_decoder_cb_impl.cc:151 of the reference), so
// packing while staging halves the bytes that cross PCIe -- and a GNU Radio buffer is
// pageable, so it has to be copied into pinned memory anyway.
#pragma once
#include <cstddef>

namespace ldpc535 {

// A persistent team of worker threads (created once per handle, parked on a condition variable
// between calls): pack() splits the span over the team and returns when every part is done.
class PackPool
{
public:
    explicit PackPool(int threads);       // threads >= 1 including the calling thread
    ~PackPool();
    PackPool(const PackPool &) = delete;
    PackPool &operator=(const PackPool &) = delete;
    int threads() const { return n_threads_; }
    // dst[i] = src[2 * i] for i < n
    void pack(const float *src_interleaved, float *dst, size_t n);

private:
    struct Impl;
    Impl *impl_;
    int n_threads_;
};

// Cores this process may run on (sched_getaffinity), capped at 32; LDPC535_PACK_THREADS overrides.
int default_pack_threads();
// dst[i] = src[2 * i], on the calling thread (AVX2 with streaming stores when the CPU has it,
// checked at run time; scalar otherwise)
void pack_real_parts_serial(const float *src_interleaved, float *dst, size_t n);
}  // namespace ldpc535
