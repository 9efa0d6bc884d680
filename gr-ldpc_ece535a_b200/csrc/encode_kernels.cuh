// Encoder kernels: parity c = P d over GF(2) on bit-packed words (AND/XOR + popcount),
// then the BPSK map bit 1 -> +1+0j, bit 0 -> -1+0j, parity symbols first, data symbols
// second (lib/ldpc_encoder_bc_impl.cc:138-165).  P = (L U)^-1 B replaces the two dense
// real-valued LAPACK solves the reference runs per frame (:275-294, :180-223).
// Output-bandwidth bound: K/8 bytes in, 8 N bytes out per frame; stores are 128-bit and
// warp-coalesced.  Tensor cores are deliberately unused: this is not a float contraction.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ldpc535 {

struct EncodeParams {
    const uint8_t *in;       // n_frames * nbytes
    long long n_frames;
    float2 *out;             // n_frames * N complex
    int M, N, K, nbytes, kwords, mwords;
    const uint32_t *Pt;      // [K][mwords]  column masks (small codes)
    const uint32_t *Pw;      // [kwords][M]  word-major rows (generic)
    int row_splits;          // generic kernel: a frame tile is shared by this many CTAs (parity rows and data half split), >= 1
};

// bytes (MSB first) -> word with bit k = data bit k
__device__ __forceinline__ uint32_t msb_bytes_to_bits(uint32_t le_word)
{
    return __byte_perm(__brev(le_word), 0, 0x0123);
}

__device__ __forceinline__ float bpsk(uint32_t bit) { return bit ? 1.f : -1.f; }

// ---- small codes: M <= 32, K <= 32 (the shipped 32x64 code).  One frame per lane for the
// GF(2) product, then the warp writes its 32 frames cooperatively: 512 B per instruction.
__global__ void __launch_bounds__(256)
encode_small_kernel(const EncodeParams p)
{
    __shared__ uint32_t sPt[32];
    if (threadIdx.x < 32) sPt[threadIdx.x] = (threadIdx.x < (unsigned)p.K) ? p.Pt[threadIdx.x] : 0u;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int M = p.M, N = p.N, K = p.K;
    const long long warps_total = (long long)gridDim.x * (blockDim.x >> 5);
    const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const bool aligned4 = (p.nbytes == 4);
    const bool vec4 = ((N & 1) == 0);
    const uint32_t kmask = (K >= 32) ? 0xffffffffu : ((1u << K) - 1u);

    for (long long base = warp_global * 32; base < p.n_frames; base += warps_total * 32) {
        const long long f = base + lane;
        uint32_t d = 0, c = 0;
        if (f < p.n_frames) {
            if (aligned4) {
                d = msb_bytes_to_bits(__ldg(reinterpret_cast<const uint32_t *>(p.in) + f));
            } else {
                for (int b = 0; b < p.nbytes; b++)
                    d |= (__brev((uint32_t)__ldg(p.in + f * p.nbytes + b)) >> 24) << (8 * b);
            }
            d &= kmask;
#pragma unroll 8
            for (int k = 0; k < 32; k++) c ^= sPt[k] & (0u - ((d >> k) & 1u));
        }
        const unsigned long long fw = (unsigned long long)c | ((unsigned long long)d << M);
        const uint32_t flo = (uint32_t)fw, fhi = (uint32_t)(fw >> 32);
        const int nf = (int)min((long long)32, p.n_frames - base);
        for (int t = 0; t < nf; t++) {
            const uint32_t lo = __shfl_sync(0xffffffffu, flo, t);
            const uint32_t hi = __shfl_sync(0xffffffffu, fhi, t);
            float2 *dst = p.out + (base + t) * (long long)N;
            if (vec4) {
                const int i = 2 * lane;
                if (i < N) {
                    const uint32_t wsel = (i < 32) ? lo : hi;
                    const uint32_t two = (wsel >> (i & 31)) & 3u;
                    __stcs(reinterpret_cast<float4 *>(dst) + lane,
                           make_float4(bpsk(two & 1u), 0.f, bpsk(two >> 1), 0.f));
                }
            } else {
                if (lane < N) __stcs(dst + lane, make_float2(bpsk((lo >> lane) & 1u), 0.f));
                if (lane + 32 < N) __stcs(dst + lane + 32, make_float2(bpsk((hi >> lane) & 1u), 0.f));
            }
        }
    }
}

// ---- generic: a CTA encodes a tile of kEncTile frames.  Data words sit in shared memory
// (broadcast reads); a thread owns kEncRows parity rows and streams their bit-packed rows of P
// once per tile (word-major layout -> coalesced, served by L2), accumulating
// acc[t][r] ^= P[row_r][w] & d[t][w] with one LOP3 each.  Every 16-byte broadcast load of data
// feeds 4 * kEncRows LOP3s, so the integer pipe, not the shared-memory pipe, sets the pace:
// M*K/32 AND-XOR words per frame (524 288 for the n = 8192 code) at 64 lanes/clk/SM bound this
// kernel at ~36 % of the HBM write roofline.
// Small batches (fewer tiles than twice the SMs) would leave most SMs idle for the 67 us a tile of
// the n = 8192 code takes: p.row_splits CTAs then share a tile, each with M / row_splits parity
// rows (a multiple of the 1024 rows a CTA pass covers) and its share of the data half.
#ifndef ENC_TILE
#define ENC_TILE 16
#endif
#ifndef ENC_ROWS
#define ENC_ROWS 4
#endif
constexpr int kEncTile = ENC_TILE;
constexpr int kEncRows = ENC_ROWS;
constexpr int kEncThreads = 256;

__host__ __device__ inline size_t encode_generic_smem_bytes(int kwords, int mwords)
{
    return sizeof(uint32_t) * (size_t)kEncTile * ((size_t)((kwords + 3) & ~3) + (size_t)mwords);
}

__global__ void __launch_bounds__(kEncThreads)
encode_generic_kernel(const EncodeParams p)
{
    extern __shared__ __align__(16) uint32_t enc_smem[];
    const int M = p.M, N = p.N, K = p.K;
    const int kw4 = (p.kwords + 3) & ~3;                 // padded to uint4
    uint32_t *dsm = enc_smem;                            // [kEncTile][kw4]
    uint32_t *csm = enc_smem + (size_t)kEncTile * kw4;   // [kEncTile][mwords]
    const int tid = threadIdx.x, lane = tid & 31;
    const long long n_tiles = (p.n_frames + kEncTile - 1) / kEncTile;
    const int RS = p.row_splits;

    for (long long unit = blockIdx.x; unit < n_tiles * RS; unit += gridDim.x) {
        const long long tile = unit / RS;
        const int rs = (int)(unit - tile * RS);
        const int row_lo = (int)((long long)M * rs / RS), row_hi = (int)((long long)M * (rs + 1) / RS);   // parity rows of this CTA
        const int dat_lo = M + (int)((long long)K * rs / RS) / 2 * 2, dat_hi = rs + 1 == RS ? N : M + (int)((long long)K * (rs + 1) / RS) / 2 * 2;
        const long long f0 = tile * kEncTile;
        const int nf = (int)min((long long)kEncTile, p.n_frames - f0);
        __syncthreads();
        for (int idx = tid; idx < kEncTile * kw4; idx += kEncThreads) {
            const int t = idx / kw4, w = idx % kw4;
            uint32_t word = 0;
            if (t < nf && w < p.kwords) {
                const uint8_t *src = p.in + (f0 + t) * p.nbytes;
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const int byte = 4 * w + b;
                    if (byte < p.nbytes) word |= (__brev((uint32_t)__ldg(src + byte)) >> 24) << (8 * b);
                }
                const int rem = K - 32 * w;
                if (rem < 32) word &= (1u << rem) - 1u;
            }
            dsm[idx] = word;
        }
        __syncthreads();
        for (int jb = row_lo; jb < row_hi; jb += kEncThreads * kEncRows) {
            uint32_t acc[kEncRows][kEncTile];
#pragma unroll
            for (int r = 0; r < kEncRows; r++)
#pragma unroll
                for (int t = 0; t < kEncTile; t++) acc[r][t] = 0;
            for (int w = 0; w < kw4; w += 4) {
                uint32_t pw[kEncRows][4];
#pragma unroll
                for (int r = 0; r < kEncRows; r++) {
                    const int j = jb + r * kEncThreads + tid;
#pragma unroll
                    for (int q = 0; q < 4; q++)
                        pw[r][q] = (j < row_hi && w + q < p.kwords) ? __ldg(p.Pw + (size_t)(w + q) * M + j) : 0u;
                }
#pragma unroll
                for (int t = 0; t < kEncTile; t++) {
                    const uint4 dd = *reinterpret_cast<const uint4 *>(dsm + (size_t)t * kw4 + w);
#pragma unroll
                    for (int r = 0; r < kEncRows; r++)
                        acc[r][t] ^= (pw[r][0] & dd.x) ^ (pw[r][1] & dd.y) ^ (pw[r][2] & dd.z) ^ (pw[r][3] & dd.w);
                }
            }
#pragma unroll
            for (int r = 0; r < kEncRows; r++) {
                const int j = jb + r * kEncThreads + tid;
#pragma unroll
                for (int t = 0; t < kEncTile; t++) {
                    const uint32_t wd = __ballot_sync(0xffffffffu, __popc(acc[r][t]) & 1);
                    if (lane == 0 && j < row_hi) csm[(size_t)t * p.mwords + (j >> 5)] = wd;
                }
            }
        }
        __syncthreads();
        for (int t = 0; t < nf; t++) {
            float2 *dst = p.out + (f0 + t) * (long long)N;
            const uint32_t *cw = csm + (size_t)t * p.mwords;
            const uint32_t *dw = dsm + (size_t)t * kw4;
            // this CTA's parity rows, then its share of the data half (two loops: a filter inside one loop over
            // all N symbols cost 15 % at 2 000 frames)
            if ((N & 1) == 0 && (M & 1) == 0) {
                for (int i = row_lo + 2 * tid; i < row_hi; i += 2 * kEncThreads) {
                    const uint32_t two = (cw[i >> 5] >> (i & 31)) & 3u;
                    __stcs(reinterpret_cast<float4 *>(dst) + (i >> 1), make_float4(bpsk(two & 1u), 0.f, bpsk(two >> 1), 0.f));
                }
                for (int i = dat_lo + 2 * tid; i < dat_hi; i += 2 * kEncThreads) {
                    const int k = i - M;
                    const uint32_t two = (dw[k >> 5] >> (k & 31)) & 3u;
                    __stcs(reinterpret_cast<float4 *>(dst) + (i >> 1), make_float4(bpsk(two & 1u), 0.f, bpsk(two >> 1), 0.f));
                }
            } else {
                for (int i = row_lo + tid; i < row_hi; i += kEncThreads)
                    __stcs(dst + i, make_float2(bpsk((cw[i >> 5] >> (i & 31)) & 1u), 0.f));
                for (int i = dat_lo + tid; i < dat_hi; i += kEncThreads) {
                    const int k = i - M;
                    __stcs(dst + i, make_float2(bpsk((dw[k >> 5] >> (k & 31)) & 1u), 0.f));
                }
            }
        }
    }
}

}  // namespace ldpc535
