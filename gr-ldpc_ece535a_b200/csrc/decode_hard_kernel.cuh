// Hard-decision decoder for 64-symbol codes (method 3, decodeHard,
// lib/ldpc_decoder_cb_impl.cc:559-572, plus checkFrame :236-253 and the byte packing :207-219).
// Purely HBM-bound: 512 B in, 6 B out per codeword and ~30 warp instructions.  Lane l takes
// symbols l and l + 32 (two 8-byte loads, each 256 contiguous bytes per warp), so the two ballot
// words are the codeword's bits 0..31 and 32..63 in natural order: the parity of check `lane` is
// one popc against its row masks, and the data bytes are a funnel shift + bit reverse of the
// ballot words -- no per-bit work.  Four codewords are in flight per warp to cover HBM latency.
// (First version: one 16-byte load per lane and even/odd ballot words; assembling the bytes from
// the split words made the kernel integer-pipe bound at 3.5 TB/s -- profiles/r1_microbench.txt.)
#pragma once
#include "decode_kernels.cuh"

namespace ldpc535 {

constexpr int kHardThreads = 256;
#ifndef HARD_UNROLL
#define HARD_UNROLL 4
#endif
#ifndef HARD_MIN_BLOCKS
#define HARD_MIN_BLOCKS 8
#endif
constexpr int kHardUnroll = HARD_UNROLL;

template <int DC>
__global__ void __launch_bounds__(kHardThreads, HARD_MIN_BLOCKS)
decode_hard64_kernel(const DecodeParams p)
{
    const int lane = threadIdx.x & 31;
    const int M = p.M;
    uint32_t row_lo = 0, row_hi = 0;                      // row of check `lane`
#pragma unroll
    for (int s = 0; s < DC; s++) {
        const int v = (lane < M) ? p.chk_var[s * M + lane] : 0xFFFF;
        if (v != 0xFFFF) {
            if (v < 32) row_lo |= 1u << v; else row_hi |= 1u << (v - 32);
        }
    }
    const long long warps_total = (long long)gridDim.x * (kHardThreads / 32);
    const long long warp0 = (long long)blockIdx.x * (kHardThreads / 32) + (threadIdx.x >> 5);
    const bool word_out = (p.nbytes == 4);

    // a warp owns 32 consecutive codewords per trip: it decodes them kHardUnroll at a time and
    // lane l keeps the results of the l-th, so the outputs leave as coalesced warp-wide stores
    for (long long base = warp0 * 32; base < p.n_win; base += warps_total * 32) {
        uint32_t my_bytes = 0, my_synd = 255, my_iters = 255;
#pragma unroll 1
        for (int r = 0; r < 32; r += kHardUnroll) {
            if (base + r >= p.n_win) break;              // warp-uniform
            float x0[kHardUnroll], x1[kHardUnroll];
            bool ok[kHardUnroll];
#pragma unroll
            for (int u = 0; u < kHardUnroll; u++) {
                const long long w = base + r + u;
                ok[u] = false;
                x0[u] = x1[u] = 0.f;
                if (w < p.n_win) {
                    const long long off = p.win_offset ? p.win_offset[w] : w * 64LL;
                    const bool neg = p.polarity && p.polarity[w] < 0;
                    ok[u] = off >= 0 && off + 64 <= p.n_sym;
                    if (ok[u]) {
                        if (p.sym_re) {
                            x0[u] = __ldcs(p.sym_re + off + lane);
                            x1[u] = __ldcs(p.sym_re + off + lane + 32);
                        } else {
                            x0[u] = __ldcs(p.sym + off + lane).x;      // 8-byte loads, imaginary part unused
                            x1[u] = __ldcs(p.sym + off + lane + 32).x;
                        }
                        if (neg) { x0[u] = -x0[u]; x1[u] = -x1[u]; }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < kHardUnroll; u++) {
                // rx < 0 -> 0 else 1
                const uint32_t h_lo = __ballot_sync(0xffffffffu, !(x0[u] < 0.f));
                const uint32_t h_hi = __ballot_sync(0xffffffffu, !(x1[u] < 0.f));
                const uint32_t bad = __ballot_sync(0xffffffffu, __popc((row_lo & h_lo) ^ (row_hi & h_hi)) & 1);
                if (lane == r + u && ok[u]) {
                    // data bits = columns M .. 63, first bit of a byte is its MSB
                    const uint32_t data = (M >= 32) ? (h_hi >> (M - 32)) : __funnelshift_r(h_lo, h_hi, M);
                    my_bytes = __byte_perm(__brev(data), 0, 0x0123);
                    my_synd = (uint32_t)min(__popc(bad), p.thr + 1);
                    my_iters = 0;
                }
            }
        }
        const long long w = base + lane;
        if (w < p.n_win) {
            if (word_out) {
                reinterpret_cast<uint32_t *>(p.out_bytes)[w] = my_bytes;
            } else {
                for (int b = 0; b < p.nbytes; b++) p.out_bytes[w * p.nbytes + b] = (uint8_t)(my_bytes >> (8 * b));
            }
            if (p.out_synd) p.out_synd[w] = (uint8_t)my_synd;
            if (p.out_iters) p.out_iters[w] = (uint8_t)my_iters;
        }
    }
}

}  // namespace ldpc535
