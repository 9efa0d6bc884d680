// Large-code encoder: parity c = P d over GF(2) by table look-up (the "Method of the Four
// Russians" applied to the generator), for codes whose dense P no longer fits a per-frame
// AND/XOR scan (BASELINE config 4's n = 8192 code: P is 4096 x 4096 bits = 2 MB, 524 288
// AND-XOR words per frame, which capped encode_generic_kernel at ~36 % of the HBM write
// roofline on the integer pipe).
//
// One input BYTE selects 8 columns of P at once.  For every byte position g (K/8 of them) and
// byte value v a table row holds the XOR of the selected columns, cut into blocks of 1024 parity
// rows:  T[rb][g][v] = 128 bytes.  A frame's parity block is the XOR of K/8 table rows -- 8x
// fewer word operations than the scan, all of them 128-bit.  The table (M/1024 * K/8 * 32 KB =
// 64 MB for the n = 8192 code) lives in HBM/L2 and is built once per code on the GPU
// (encode_m4r_build_kernel).
//
// encode_m4r_kernel: a CTA of 1024 threads encodes a tile of 128 * TPF frames.  For each row
// block the CTA streams the K/8 tables T[rb][g][*] (32 KB each) through a ring of shared-memory
// stages with cp.async; 8 lanes share a frame, lane c owns the 16-byte chunk c of the running
// XOR, so every shared-memory wavefront is one whole table row (conflict-free), and each table
// byte fetched from L2 is used by all frames of the tile.  The bit order inside a table row is
// chosen so that the 8 lanes of a frame write 128 contiguous bytes of BPSK symbols per store:
//     chunk c, word i, bit b   <->   parity row  rb*1024 + 16*(16 i + b/2) + 2 c + (b & 1).
// The data half of the codeword (symbols M .. N-1) is written by the same kernel, spread over its steps.
// Output-bandwidth bound by design: 8 N bytes out per frame; what limits it in practice is the
// shared-memory read of 128 bytes per (frame, byte position, row block).
#pragma once
#include "encode_kernels.cuh"
#include "decode_kernels.cuh"   // smem_u32

namespace ldpc535 {

constexpr int kM4rRows = 1024;          // parity rows per block  (128-byte table rows)
constexpr int kM4rThreads = 1024;
constexpr int kM4rSlots = kM4rThreads / 8;
constexpr int kM4rStages = 4;           // 4 x 32 KB ring
constexpr size_t kM4rStageBytes = 256 * 128;

__host__ __device__ inline size_t encode_m4r_table_bytes(int M, int K)
{
    return (size_t)(M / kM4rRows) * (size_t)(K / 8) * kM4rStageBytes;
}

// One warp per (row block, byte position): lane = word (c, i) of the 128-byte row.
__global__ void __launch_bounds__(32)
encode_m4r_build_kernel(const uint32_t *__restrict__ Pt, uint32_t *__restrict__ T, int M, int K, int mwords)
{
    const int G = K / 8;
    const int rb = blockIdx.x / G, g = blockIdx.x % G;
    const int wd = threadIdx.x, c = wd >> 2, i = wd & 3;
    (void)M;
    uint32_t pcol[8];                                   // permuted column words of data bits 8g .. 8g+7
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const uint32_t *col = Pt + (size_t)(8 * g + j) * mwords;
        uint32_t w = 0;
        for (int b = 0; b < 32; b++) {
            const int r = rb * kM4rRows + 16 * (16 * i + (b >> 1)) + 2 * c + (b & 1);
            w |= ((__ldg(col + (r >> 5)) >> (r & 31)) & 1u) << b;
        }
        pcol[j] = w;
    }
    uint32_t *dst = T + ((size_t)blockIdx.x * 256) * 32 + wd;
    for (int v = 0; v < 256; v++) {
        uint32_t e = 0;
#pragma unroll
        for (int j = 0; j < 8; j++)                     // data bit 8g + j is bit 7-j of the byte (MSB first)
            e ^= pcol[j] & (0u - ((v >> (7 - j)) & 1u));
        dst[(size_t)v * 32] = e;
    }
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;"
                 :: "r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N_) : "memory"); }

__device__ __forceinline__ float bpsk_bit(uint32_t word, int b)
{
    return __uint_as_float(0xbf800000u ^ (((word >> b) & 1u) << 31));     // 1 -> +1.0, 0 -> -1.0
}

// TPF frames per thread slot: a tile is 128 * TPF frames; the unit of work is (tile, row block), so
// a CTA's accumulators are 128 bytes per frame and the units spread evenly over the SMs.  A unit
// also writes its share of the data half of the codeword (K/8 / RB input bytes per frame: one
// 4-byte word = 256 bytes of symbols per barrier interval), so those stores are spread over its
// steps: an SM drains stores at ~32 B/clk and a burst stalls every warp behind the next barrier.
//
// Data movement: the table stages T[rb][g][*] (32 KB each) arrive by 1-D TMA bulk copies issued
// by one thread, each completing on its own mbarrier (4-slot ring, two stages consumed per CTA
// barrier, the next two requested right after it); the frames' input bytes are staged 16 bytes
// per frame at a time with cp.async, one chunk ahead, so that no global-load latency sits in the
// look-up loop.  All per-frame addresses are 32-bit offsets from the kernel's base pointers (the
// host splits batches so that they fit); shared memory is addressed with 32-bit offsets too.
// Requires K % 128 == 0, M % 1024 == 0, M >= 4096 (TPF <= 2 M/1024), (K/32) % (M/1024) == 0, p.in and p.out 16-byte aligned,
// n_frames * K/32 < 2^32 and n_frames * N / 2 < 2^32.
// Shared memory: ring[4][32 KB] | input chunks [2][TPF][128][16 B] | 4 mbarriers.
template <int TPF>
__host__ __device__ constexpr size_t encode_m4r_smem_bytes()
{
    return (size_t)kM4rStages * kM4rStageBytes + 2 * (size_t)TPF * kM4rSlots * 16 + 64;
}

__device__ __forceinline__ uint4 lds128(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
}

template <int TPF>
__global__ void __launch_bounds__(kM4rThreads, 1)
encode_m4r_kernel(const EncodeParams p, const uint4 *__restrict__ T)
{
    extern __shared__ __align__(128) unsigned char m4r_smem[];
    constexpr uint32_t kStage = (uint32_t)kM4rStageBytes;
    constexpr uint32_t kInBuf = (uint32_t)TPF * kM4rSlots * 16;
    const int tid = threadIdx.x, slot = tid >> 3, c = tid & 7;
    const int G = p.K >> 3, RB = p.M / kM4rRows;
    const uint32_t smem0 = smem_u32(m4r_smem);
    const uint32_t ring_c = smem0 + 16u * c;                       // + stage * 32 KB + v * 128
    const uint32_t in_s = smem0 + kM4rStages * kStage;             // + buf * kInBuf + (t * 128 + slot) * 16
    const uint32_t bars = in_s + 2 * kInBuf;                       // 4 x 8 bytes
    const uint32_t sys_words = (uint32_t)(G / RB) >> 2;            // 4-byte data words of a frame this unit writes out
    const uint32_t tile_frames = (uint32_t)kM4rSlots * TPF;
    const uint32_t n_frames = (uint32_t)p.n_frames;
    const uint32_t n_tiles = (n_frames + tile_frames - 1) / tile_frames;
    const uint32_t n_units = n_tiles * (uint32_t)RB;
    const uint32_t in_wstride = (uint32_t)p.nbytes >> 2;           // input words per frame
    const uint32_t out_qstride = (uint32_t)p.N >> 1;               // float4 (symbol pairs) per frame
    const uint32_t *__restrict__ inw = reinterpret_cast<const uint32_t *>(p.in);
    float4 *__restrict__ outq = reinterpret_cast<float4 *>(p.out);

    if (tid == 0) {
#pragma unroll
        for (int b = 0; b < kM4rStages; b++)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bars + 8u * b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t uses = 0;                                             // completed phases of every ring slot
    const bool early = ((tid >> 7) & 1) != 0;                      // warps 4-7, 12-15, ..: stores before the look-ups

    for (uint32_t unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const uint32_t tile = unit / (uint32_t)RB;
        const uint32_t rb = unit - tile * (uint32_t)RB;
        const uint32_t f0 = tile * tile_frames + slot;             // frame of t = 0; frame t is f0 + 128 t
        const uint32_t in0 = f0 * in_wstride;                      // word offset of frame f0
        const uint32_t out0 = f0 * out_qstride;                    // float4 offset of frame f0
        const int tmax = f0 >= n_frames ? 0 : (int)min((uint32_t)TPF, (n_frames - f0 + kM4rSlots - 1) / kM4rSlots);
        const unsigned char *Trb = reinterpret_cast<const unsigned char *>(T) + (size_t)rb * G * kStage;
        auto request_stage = [&](int g) {                          // thread 0: T[rb][g][*] -> ring slot g % 4
            if (g < G) {
                const uint32_t bar = bars + 8u * (g & (kM4rStages - 1));
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(kStage) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             :: "r"(smem0 + (uint32_t)(g & (kM4rStages - 1)) * kStage), "l"(Trb + (size_t)g * kStage),
                                "r"(kStage), "r"(bar) : "memory");
            }
        };
        auto request_input = [&](int chunk) {                      // lane c stages 16 bytes of frame t = c
            if (c < tmax) {
                const uint32_t dst = in_s + (uint32_t)(chunk & 1) * kInBuf + ((uint32_t)c * kM4rSlots + slot) * 16u;
                const uint32_t *src = inw + (in0 + (uint32_t)c * kM4rSlots * in_wstride + 4u * chunk);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
            }
            cp_async_commit();
        };
        __syncthreads();                                           // the previous unit has been read to the end
        if (tid == 0) { request_stage(0); request_stage(1); }
        request_input(0);
        uint4 acc[TPF];
#pragma unroll
        for (int t = 0; t < TPF; t++) acc[t] = make_uint4(0u, 0u, 0u, 0u);
        const uint32_t n_pairs = sys_words * TPF;                  // (frame, data word) pairs of the unit, <= G / 2
        uint32_t sysw = tmax > 0 ? __ldg(inw + (in0 + rb * sys_words)) : 0u;   // data word written out next
        for (int g4 = 0; g4 < G; g4 += 4) {
            uint32_t vw[TPF];                                      // input bytes g4 .. g4+3 of this slot's frames
#pragma unroll
            for (int half = 0; half < 2; half++) {
                const int g = g4 + 2 * half;
                if (half == 0 && (g4 & 15) == 0) cp_async_wait<0>();   // this chunk of input bytes (requested 16 steps ago)
                __syncthreads();                                   // ring slots of g-2, g-1 and the other input buffer are free
                if (tid == 0) { request_stage(g + 2); request_stage(g + 3); }
                if (half == 0) {
                    if ((g4 & 15) == 0 && g4 + 16 < G) request_input((g4 >> 4) + 1);
                    const uint32_t ia = in_s + (uint32_t)((g4 >> 4) & 1) * kInBuf + (uint32_t)slot * 16u + (uint32_t)(g4 & 12);
#pragma unroll
                    for (int t = 0; t < TPF; t++) vw[t] = lds32(ia + (uint32_t)t * kM4rSlots * 16u);
                }
                // data half, spread thinly: one (frame, word) pair = 32 symbols = 16 float4 per barrier interval
                // (lane c writes float4 c and c + 8); pair k of the unit is frame t = k % TPF, word k / TPF.
                // The word of the next pair is fetched one interval ahead of its use.
                auto store_data_half = [&]() {
                    const uint32_t k = (uint32_t)g >> 1;
                    if (k < n_pairs) {
                        const int t = (int)(k % TPF);
                        if (t < tmax) {
                            const uint32_t sw = rb * sys_words + k / TPF;
                            const int b0 = 8 * (c >> 2) + 7 - 2 * (c & 3);         // symbol j of the word is bit 8 (j / 8) + 7 - j % 8
                            float4 *dst = outq + (out0 + (uint32_t)t * kM4rSlots * out_qstride + ((uint32_t)p.M >> 1) + 16u * sw + c);
                            __stcs(dst, make_float4(bpsk_bit(sysw, b0), 0.f, bpsk_bit(sysw, b0 - 1), 0.f));
                            __stcs(dst + 8, make_float4(bpsk_bit(sysw, b0 + 16), 0.f, bpsk_bit(sysw, b0 + 15), 0.f));
                        }
                    }
                    if (k + 1 < n_pairs) {
                        const int t2 = (int)((k + 1) % TPF);
                        if (t2 < tmax)
                            sysw = __ldg(inw + (in0 + (uint32_t)t2 * kM4rSlots * in_wstride + rb * sys_words + (k + 1) / TPF));
                    }
                };
                // One SM drains stores at 32 B/clk (tools/microbench/store_rate.cu), and a warp whose store waits
                // for that path also holds back its own shared-memory reads behind it (in-order MIO queue).  After
                // a CTA barrier all warps sit at the same program point, so half of the warps of every
                // sub-partition issue the stores of the interval before its look-ups, the other half after them.
                if (early) store_data_half();
                const uint32_t parity = (uses + ((uint32_t)g >> 2)) & 1u;
                mbar_wait(bars + 8u * (2 * half), parity);
                mbar_wait(bars + 8u * (2 * half + 1), parity);
#pragma unroll
                for (int gg = 0; gg < 2; gg++) {
                    const uint32_t stage = ring_c + (uint32_t)(2 * half + gg) * kStage;    // (g + gg) % 4, g4 % 4 == 0
#pragma unroll
                    for (int t = 0; t < TPF; t++) {
                        const int sh = 8 * (2 * half + gg);
                        const uint32_t voff = sh >= 7 ? (vw[t] >> (sh - 7)) & 0x7f80u : (vw[t] << (7 - sh)) & 0x7f80u;   // byte * 128
                        const uint4 e = lds128(stage + voff);
                        acc[t].x ^= e.x; acc[t].y ^= e.y; acc[t].z ^= e.z; acc[t].w ^= e.w;
                    }
                }
                if (!early) store_data_half();
            }
        }
        uses += (uint32_t)G >> 2;
        // parity symbols of this row block: 8 lanes of a frame write 128 contiguous bytes per store
#pragma unroll
        for (int t = 0; t < TPF; t++) {
            if (t >= tmax) continue;
            float4 *dst = outq + (out0 + (uint32_t)t * kM4rSlots * out_qstride + rb * (kM4rRows / 2) + c);
            const uint32_t w4[4] = {acc[t].x, acc[t].y, acc[t].z, acc[t].w};
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int s2 = 0; s2 < 16; s2++)
                    __stcs(dst + 8 * (16 * i + s2),
                           make_float4(bpsk_bit(w4[i], 2 * s2), 0.f, bpsk_bit(w4[i], 2 * s2 + 1), 0.f));
        }
    }
}

}  // namespace ldpc535
