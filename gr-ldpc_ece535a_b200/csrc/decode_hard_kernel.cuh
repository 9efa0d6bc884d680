// Hard-decision decoder for 64-symbol codes (method 3, decodeHard,
// lib/ldpc_decoder_cb_impl.cc:559-572, plus checkFrame :236-253 and the byte packing :207-219).
// Purely HBM-bound: 512 B in, 6 B out per codeword and a handful of integer instructions.
// One warp instruction loads one whole codeword (lane l: symbols 2l, 2l+1 as one 16-byte load,
// 512 contiguous bytes per warp); the even / odd symbol decisions come back as two ballot words
// and stay in that split form: lane j's parity-check row is split the same way once per kernel,
// and the four data bytes are assembled from the split words.  Four codewords are in flight per
// warp to cover the HBM latency.
#pragma once
#include "decode_kernels.cuh"

namespace ldpc535 {

constexpr int kHardThreads = 256;
constexpr int kHardUnroll = 4;

template <int DC>
__global__ void __launch_bounds__(kHardThreads)
decode_hard64_kernel(const DecodeParams p)
{
    const int lane = threadIdx.x & 31;
    const int M = p.M;
    // row of check `lane`, split into even and odd columns: bit l of row_e <-> column 2l
    uint32_t row_e = 0, row_o = 0;
#pragma unroll
    for (int s = 0; s < DC; s++) {
        const int v = (lane < M) ? p.chk_var[s * M + lane] : 0xFFFF;
        if (v != 0xFFFF) {
            if (v & 1) row_o |= 1u << (v >> 1); else row_e |= 1u << (v >> 1);
        }
    }
    const long long warps_total = (long long)gridDim.x * (kHardThreads / 32);
    const long long warp0 = (long long)blockIdx.x * (kHardThreads / 32) + (threadIdx.x >> 5);
    const int dshift = M >> 1;                           // data bits start at column M (M even)

    for (long long base = warp0 * kHardUnroll; base < p.n_win; base += warps_total * kHardUnroll) {
        float4 v[kHardUnroll];
        float pol[kHardUnroll];
        bool ok[kHardUnroll];
#pragma unroll
        for (int u = 0; u < kHardUnroll; u++) {
            const long long w = base + u;
            ok[u] = false;
            pol[u] = 1.f;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (w < p.n_win) {
                const long long off = p.win_offset ? p.win_offset[w] : w * 64LL;
                pol[u] = p.polarity ? (float)p.polarity[w] : 1.f;
                ok[u] = off >= 0 && off + 64 <= p.n_sym;
                if (ok[u]) {
                    if (p.sym_re) {
                        if ((off & 1) == 0) {
                            const float2 t = __ldcs(reinterpret_cast<const float2 *>(p.sym_re + off) + lane);
                            v[u].x = t.x; v[u].z = t.y;
                        } else {
                            v[u].x = __ldcs(p.sym_re + off + 2 * lane);
                            v[u].z = __ldcs(p.sym_re + off + 2 * lane + 1);
                        }
                    } else if ((off & 1) == 0) {
                        v[u] = __ldcs(reinterpret_cast<const float4 *>(p.sym + off) + lane);
                    } else {
                        v[u].x = __ldcs(&p.sym[off + 2 * lane].x);
                        v[u].z = __ldcs(&p.sym[off + 2 * lane + 1].x);
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kHardUnroll; u++) {
            const long long w = base + u;
            if (w >= p.n_win) break;                     // warp-uniform
            // rx = pol * Re;  rx < 0 -> 0 else 1
            const uint32_t be = __ballot_sync(0xffffffffu, !(pol[u] * v[u].x < 0.f));
            const uint32_t bo = __ballot_sync(0xffffffffu, !(pol[u] * v[u].z < 0.f));
            const uint32_t bad = __ballot_sync(0xffffffffu, (__popc((row_e & be) ^ (row_o & bo))) & 1);
            if (lane < p.nbytes) {
                // byte `lane`: columns M + 8 lane .. + 7, MSB first; column c sits at bit c/2 of be / bo
                const int c0 = dshift + 4 * lane;
                uint32_t byte = 0;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    byte |= ((be >> (c0 + q)) & 1u) << (7 - 2 * q);
                    byte |= ((bo >> (c0 + q)) & 1u) << (6 - 2 * q);
                }
                p.out_bytes[w * p.nbytes + lane] = ok[u] ? (uint8_t)byte : (uint8_t)0;
            }
            if (lane == 0) {
                if (p.out_synd) p.out_synd[w] = ok[u] ? (uint8_t)min(__popc(bad), p.thr + 1) : (uint8_t)255;
                if (p.out_iters) p.out_iters[w] = ok[u] ? (uint8_t)0 : (uint8_t)255;
            }
        }
    }
}

}  // namespace ldpc535
