// Host-side staging helper of the host-buffer decode path: copy the REAL parts of a span of
// gr_complex symbols into (pinned) staging memory with a few worker threads.  The decoder
// never reads the imaginary parts (lib/ldpc_decoder_cb_impl.cc:151 of the reference), so
// packing while staging halves the bytes that cross PCIe -- and a GNU Radio buffer is
// pageable, so it has to be copied into pinned memory anyway.
#pragma once
#include <cstddef>

namespace ldpc535 {
// dst[i] = src[2 * i] for i < n, split over `threads` workers (>= 1).
void pack_real_parts(const float *src_interleaved, float *dst, size_t n, int threads);
int default_pack_threads();
// ranks of a one-process-per-GPU job sharing this host (LOCAL_WORLD_SIZE / WORLD_SIZE), >= 1
int host_sharing_ranks();
}  // namespace ldpc535
