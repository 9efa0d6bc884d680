// Measurement utilities of the library (not on the decode/encode path):
//   * counter-based synthetic inputs -- Philox4x32-10 keyed by a seed, counter = GLOBAL frame
//     index, so a shard of a multi-GPU run holds exactly the frames the same indices have in a
//     one-GPU run ("seed 535, counter = global codeword index").  Data bytes uniform,
//     channel noise by Box-Muller on the real axis with the reference simulator's convention
//     sigma^2 = N0 = 10^(-EbN0/10) (apps/ldpc_lapack.cpp:635-642); the imaginary parts stay 0
//     (the decoder never reads them, lib/ldpc_decoder_cb_impl.cc:151).
//   * pipe-ceiling probes: the SM special-function (MUFU) pipe with the decoder's 1 ex2 : 2 lg2
//     mix and the fp64 pipe with min-sum's add / compare / select mix, timed with CUDA events --
//     the denominators of the compute rooflines bench.py reports.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ldpc535 {

struct Philox {
    static constexpr uint32_t kM0 = 0xD2511F53u, kM1 = 0xCD9E8D57u, kW0 = 0x9E3779B9u, kW1 = 0xBB67AE85u;
    // Philox4x32-10 (Salmon et al., SC'11): ten rounds, key bumped by the Weyl constants
    __host__ __device__ static inline void mulhilo(uint32_t a, uint32_t b, uint32_t &hi, uint32_t &lo)
    {
        const uint64_t p = (uint64_t)a * b;
        hi = (uint32_t)(p >> 32);
        lo = (uint32_t)p;
    }
    __host__ __device__ static inline void run(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1, uint32_t (&out)[4])
    {
        for (int r = 0; r < 10; r++) {
            uint32_t hi0, lo0, hi1, lo1;
            mulhilo(kM0, c0, hi0, lo0);
            mulhilo(kM1, c2, hi1, lo1);
            const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
            k0 += kW0; k1 += kW1;
        }
        out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
    }
};

// byte b of frame f (global index) = byte (b & 15), little endian, of Philox(ctr = (f_lo, f_hi, b >> 4, 0))
__global__ void synth_bytes_kernel(uint8_t *out, long long first_frame, long long n_frames, int nbytes,
                                   uint32_t seed_lo, uint32_t seed_hi)
{
    const int blocks_per_frame = (nbytes + 15) >> 4;
    const long long total = n_frames * blocks_per_frame;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const long long fl = t / blocks_per_frame;
        const int blk = (int)(t - fl * blocks_per_frame);
        const unsigned long long f = (unsigned long long)(first_frame + fl);
        uint32_t x[4];
        Philox::run((uint32_t)f, (uint32_t)(f >> 32), (uint32_t)blk, 0u, seed_lo, seed_hi, x);
        uint8_t *dst = out + fl * nbytes + blk * 16;
        const int n = min(16, nbytes - blk * 16);
        for (int b = 0; b < n; b++) dst[b] = (uint8_t)(x[b >> 2] >> (8 * (b & 3)));
    }
}

// re(symbol i of frame f) += sigma * n, n ~ N(0,1): symbols 4q .. 4q+3 take the two Box-Muller pairs
// of Philox(ctr = (f_lo, f_hi, q, 1)); u = (x + 0.5) * 2^-32 in (0, 1)
__global__ void synth_awgn_kernel(float2 *sym, long long first_frame, long long n_frames, int N, float sigma,
                                  uint32_t seed_lo, uint32_t seed_hi)
{
    const int quads = N >> 2;
    const long long total = n_frames * quads;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const long long fl = t / quads;
        const int q = (int)(t - fl * quads);
        const unsigned long long f = (unsigned long long)(first_frame + fl);
        uint32_t x[4];
        Philox::run((uint32_t)f, (uint32_t)(f >> 32), (uint32_t)q, 1u, seed_lo, seed_hi, x);
        float nrm[4];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const float u1 = ((float)x[2 * h] + 0.5f) * 2.3283064365386963e-10f;
            const float u2 = ((float)x[2 * h + 1] + 0.5f) * 2.3283064365386963e-10f;
            const float rad = sqrtf(-2.f * logf(fminf(fmaxf(u1, 1e-10f), 0.99999994f)));
            float sn, cs;
            sincospif(2.f * u2, &sn, &cs);
            nrm[2 * h] = rad * cs;
            nrm[2 * h + 1] = rad * sn;
        }
        float4 *p = reinterpret_cast<float4 *>(sym + fl * N + 4 * q);      // 4 symbols = 32 bytes
        float4 a = p[0], b = p[1];
        a.x += sigma * nrm[0]; a.z += sigma * nrm[1];
        b.x += sigma * nrm[2]; b.z += sigma * nrm[3];
        p[0] = a; p[1] = b;
    }
}

// ---- pipe-ceiling probes -----------------------------------------------------------------------
__device__ __forceinline__ float probe_ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float probe_lg2(float x) { float y; asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// 8 independent chains per thread, 3 MUFU (1 ex2 : 2 lg2, the sum-product mix) per chain step
__global__ void __launch_bounds__(256) probe_mufu_kernel(float *out, int iters, float seed)
{
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = seed + 0.001f * (threadIdx.x + i);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = probe_lg2(probe_ex2(a[i])) - probe_lg2(a[i] + 2.f);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// fp64 pipe: 8 independent chains, per step 1 DADD + 1 compare-select pair (fmin lowers to
// DSETP + 2 SEL on the 32-bit halves) -- what min-sum's check and bit updates are made of.
// Counted as 2 fp64-pipe instructions (DADD, DSETP) per chain step.
__global__ void __launch_bounds__(256) probe_fp64_kernel(double *out, int iters, double seed)
{
    double a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = seed + 0.001 * (threadIdx.x + i); b[i] = seed * 3.0 + i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            a[i] = a[i] + b[i];
            b[i] = fmin(a[i], b[i] + 0.5);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i] + b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace ldpc535
