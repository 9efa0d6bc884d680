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
// encode_m4r_kernel: a CTA of 1024 threads works on units of (tile of up to 128 * TPF frames, row
// block).  The K/8 tables T[rb][g][*] of the unit (32 KB each) arrive in a ring of shared-memory
// stages by 1-D TMA bulk copies on mbarriers; 8 lanes share a frame, lane c owns the 16-byte chunk c
// of the running XOR, so every shared-memory wavefront is one whole table row (conflict-free), and
// each table byte fetched from L2 is used by all frames of the tile.  The bit order inside a table
// row is chosen so that the 8 lanes of a frame write 128 contiguous bytes of BPSK symbols per store:
//     chunk c, word i, bit b   <->   parity row  rb*1024 + 16*(16 i + b/2) + 2 c + (b & 1).
// The data half of the codeword (symbols M .. N-1) is written by the same kernel, spread over its steps.
// Output-bandwidth bound by design: 8 N bytes out per frame; what limits it in practice is the
// shared-memory read of 128 bytes per (frame, byte position, row block) plus the arrival of the
// table stages, and the stores that share the SM's load/store path with both.
#pragma once
#include <type_traits>
#include "encode_kernels.cuh"
#include "decode_kernels.cuh"   // smem_u32

namespace ldpc535 {

constexpr int kM4rRows = 1024;          // parity rows per block  (128-byte table rows)
constexpr int kM4rThreads = 1024;
constexpr int kM4rSlots = kM4rThreads / 8;
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

// A CTA of 1024 threads = 128 frame slots x 8 lanes; slot s of a tile holds frames s, s + 128, ..
// (at most TPF of them: the running XOR of a frame is 4 registers of each of its 8 lanes).  The unit
// of work is (tile, row block); tiles have run.tile_frames <= 128 * TPF frames, a size the host
// picks so that the units fill whole waves of the GPU (launch_encode: 200 000 frames are 196 tiles of
// 1024 = 5.3 units per SM, i.e. 6 on the slowest SM, or 222 tiles of 901 = 6.0 on every SM); a warp
// skips the look-ups of frames beyond its tile share.  A unit also writes its share of the data half
// of the codeword (K/8 / RB input bytes per frame), one half (frame, word) pair per stage, so those
// stores are spread over its steps: an SM drains stores at ~32 B/clk and a burst stalls every warp
// behind the next barrier.
//
// Data movement: the table stages T[rb][g][*] (32 KB each) arrive by 1-D TMA bulk copies issued by
// one thread, each completing on the mbarrier of its ring slot.  The ring has NSLOT slots; a CTA
// barrier every SPB stages frees the SPB slots just consumed, which are re-requested right after it,
// so NSLOT - SPB stages are always in flight or waiting.  The ring position and phase run on across
// units.  The frames' input bytes are staged 16 bytes per frame at a time with cp.async, one chunk
// ahead (written and read by the lanes of one warp: no CTA barrier involved).  All per-frame
// addresses are 32-bit offsets from the kernel's base pointers (the host splits batches so that they
// fit); shared memory is addressed with 32-bit offsets too.
// Requires K % 128 == 0, M % 1024 == 0, (K/32) % (M/1024) == 0, p.in and p.out 16-byte aligned,
// n_frames * K/32 < 2^32 and n_frames * N / 2 < 2^32.
// Shared memory: ring[NSLOT][32 KB] | input chunks [2][TPF][128][16 B] | NSLOT mbarriers.
struct M4rRun {
    uint32_t tile_frames;    // frames per tile, 1 .. 128 * TPF
    uint32_t flags;          // timing experiments only (results are wrong): see kM4rNo*
};
constexpr uint32_t kM4rNoParityStores = 1, kM4rNoDataStores = 2, kM4rNoFill = 4, kM4rNoLookups = 8;

template <int TPF, int NSLOT>
__host__ __device__ constexpr size_t encode_m4r_smem_bytes()
{
    return (size_t)NSLOT * kM4rStageBytes + 2 * (size_t)TPF * kM4rSlots * 16 + 64;
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

template <int TPF, int NSLOT, int SPB>
__global__ void __launch_bounds__(kM4rThreads, 1)
encode_m4r_kernel(const EncodeParams p, const uint4 *__restrict__ T, const M4rRun run)
{
    static_assert((SPB == 2 || SPB == 4) && SPB < NSLOT && NSLOT <= 8 && TPF <= 8, "ring: SPB stages consumed per barrier, NSLOT - SPB ahead");
    extern __shared__ __align__(128) unsigned char m4r_smem[];
    constexpr uint32_t kStage = (uint32_t)kM4rStageBytes;
    constexpr uint32_t kInBuf = (uint32_t)TPF * kM4rSlots * 16;
    constexpr int LEAD = NSLOT - SPB;
    const int tid = threadIdx.x, slot = tid >> 3, c = tid & 7;
    const int G = p.K >> 3, RB = p.M / kM4rRows;
    const uint32_t smem0 = smem_u32(m4r_smem);
    const uint32_t ring_c = smem0 + 16u * c;                       // + ring slot * 32 KB + v * 128
    const uint32_t in_s = smem0 + NSLOT * kStage;                  // + buf * kInBuf + (t * 128 + slot) * 16
    const uint32_t bars = in_s + 2 * kInBuf;                       // NSLOT x 8 bytes
    const uint32_t sys_words = (uint32_t)(G / RB) >> 2;            // 4-byte data words of a frame this unit writes out
    const uint32_t tile_frames = run.tile_frames;
    const uint32_t n_frames = (uint32_t)p.n_frames;
    const uint32_t n_tiles = (n_frames + tile_frames - 1) / tile_frames;
    const uint32_t n_units = n_tiles * (uint32_t)RB;
    const uint32_t in_wstride = (uint32_t)p.nbytes >> 2;           // input words per frame
    const uint32_t out_qstride = (uint32_t)p.N >> 1;               // float4 (symbol pairs) per frame
    const uint32_t *__restrict__ inw = reinterpret_cast<const uint32_t *>(p.in);
    float4 *__restrict__ outq = reinterpret_cast<float4 *>(p.out);
    const bool fill = !(run.flags & kM4rNoFill), look = !(run.flags & kM4rNoLookups);
    const bool st_par = !(run.flags & kM4rNoParityStores), st_dat = !(run.flags & kM4rNoDataStores);

    if (tid == 0) {
#pragma unroll
        for (int b = 0; b < NSLOT; b++)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bars + 8u * b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t ring_pos = 0, ring_par = 0;                           // slot and phase of the next stage to be consumed
    const bool early = ((tid >> 7) & 1) != 0;                      // warps 4-7, 12-15, ..: stores before the look-ups

    for (uint32_t unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const uint32_t tile = unit / (uint32_t)RB;
        const uint32_t rb = unit - tile * (uint32_t)RB;
        const uint32_t tile_f0 = tile * tile_frames;
        const uint32_t tile_n = min(tile_frames, n_frames - tile_f0);         // frames of this tile
        const uint32_t f0 = tile_f0 + slot;                        // frame of t = 0; frame t is f0 + 128 t
        const uint32_t in0 = f0 * in_wstride;                      // word offset of frame f0
        const uint32_t out0 = f0 * out_qstride;                    // float4 offset of frame f0
        const int tmax = (uint32_t)slot >= tile_n ? 0 : (int)min((uint32_t)TPF, (tile_n - slot + kM4rSlots - 1) / kM4rSlots);
        const int tmax_w = __shfl_sync(0xffffffffu, tmax, 0);      // the warp's first slot has the most frames
        const unsigned char *Trb = reinterpret_cast<const unsigned char *>(T) + (size_t)rb * G * kStage;
        auto request_stage = [&](int g, uint32_t rs) {             // thread 0: T[rb][g][*] -> ring slot rs
            if (g < G && fill) {
                const uint32_t bar = bars + 8u * rs;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(kStage) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             :: "r"(smem0 + rs * kStage), "l"(Trb + (size_t)g * kStage), "r"(kStage), "r"(bar) : "memory");
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
        if (tid == 0) {
#pragma unroll
            for (int i = 0; i < LEAD; i++) {
                const uint32_t rs = ring_pos + i;
                request_stage(i, rs >= NSLOT ? rs - NSLOT : rs);
            }
        }
        request_input(0);
        uint4 acc[TPF];
#pragma unroll
        for (int t = 0; t < TPF; t++) acc[t] = make_uint4(0u, 0u, 0u, 0u);
        const uint32_t n_pairs = sys_words * TPF;                  // (frame, data word) pairs of the unit, <= G / 2
        uint32_t sysw = tmax > 0 ? __ldg(inw + (in0 + rb * sys_words)) : 0u;   // data word written out next
        uint32_t vw[TPF];                                          // input bytes g & ~3 .. +3 of this slot's frames
#pragma unroll
        for (int t = 0; t < TPF; t++) vw[t] = 0u;
        // data half, spread thinly: stage g writes one half of (frame, word) pair g / 2 = 16 symbols = 8 float4
        // (lane c writes float4 c on even, c + 8 on odd stages); pair k of the unit is frame t = k % TPF,
        // word k / TPF.  The word of the next pair is fetched when the current one has been written.
        auto store_data_half = [&](int g) {
            const uint32_t k = (uint32_t)g >> 1;
            if (k >= n_pairs || !st_dat) return;
            const int t = (int)(k % TPF);
            const int odd = g & 1;
            if (t < tmax) {
                const uint32_t sw = rb * sys_words + k / TPF;
                const int b0 = 8 * (c >> 2) + 7 - 2 * (c & 3) + 16 * odd;      // symbol j of the word is bit 8 (j / 8) + 7 - j % 8
                float4 *dst = outq + (out0 + (uint32_t)t * kM4rSlots * out_qstride + ((uint32_t)p.M >> 1) + 16u * sw + c + 8 * odd);
                __stcs(dst, make_float4(bpsk_bit(sysw, b0), 0.f, bpsk_bit(sysw, b0 - 1), 0.f));
            }
            if (odd && k + 1 < n_pairs) {
                const int t2 = (int)((k + 1) % TPF);
                if (t2 < tmax)
                    sysw = __ldg(inw + (in0 + (uint32_t)t2 * kM4rSlots * in_wstride + rb * sys_words + (k + 1) / TPF));
            }
        };
        for (int g0 = 0; g0 < G; g0 += SPB) {
            __syncthreads();                                       // the SPB ring slots consumed last are free
            if (tid == 0) {
#pragma unroll
                for (int i = 0; i < SPB; i++) {
                    const uint32_t rs = ring_pos + LEAD + i;
                    request_stage(g0 + LEAD + i, rs >= NSLOT ? rs - NSLOT : rs);
                }
            }
            // One SM drains stores at 32 B/clk (tools/microbench/store_rate.cu), and a warp whose store waits
            // for that path also holds back its own shared-memory reads behind it (in-order MIO queue).  After
            // a CTA barrier all warps sit at the same program point, so half of the warps of every
            // sub-partition issue the stores of the interval before its look-ups, the other half after them.
            if (early) {
#pragma unroll
                for (int i = 0; i < SPB; i++) if (g0 + i < G) store_data_half(g0 + i);
            }
            if ((g0 & 15) == 0) {                                  // next 16 input bytes per frame (requested 16 stages ago)
                cp_async_wait<0>();
                __syncwarp();
                if (g0 + 16 < G) request_input((g0 >> 4) + 1);
            }
            if ((g0 & 3) == 0) {
                const uint32_t ia = in_s + (uint32_t)((g0 >> 4) & 1) * kInBuf + (uint32_t)slot * 16u + (uint32_t)(g0 & 12);
#pragma unroll
                for (int t = 0; t < TPF; t++) vw[t] = lds32(ia + (uint32_t)t * kM4rSlots * 16u);
            }
            uint32_t stage[SPB];
#pragma unroll
            for (int i = 0; i < SPB; i++) {
                const uint32_t rs = ring_pos + i >= NSLOT ? ring_pos + i - NSLOT : ring_pos + i;
                const uint32_t rp = ring_pos + i >= NSLOT ? ring_par ^ 1u : ring_par;
                if (fill) mbar_wait(bars + 8u * rs, rp);
                stage[i] = ring_c + rs * kStage;
            }
            // the look-ups of the interval as one straight-line block per frame count of the warp (the loads
            // of a block issue back to back; a branch per frame made each wait for its predecessor's XOR)
            auto look_up = [&](auto nt_c) {
                constexpr int NT = decltype(nt_c)::value;
#pragma unroll
                for (int i = 0; i < SPB; i++) {
                    const uint32_t sel = 0x4440u | (uint32_t)((g0 + i) & 3);   // byte (g0 + i) & 3 of vw, zero-extended
#pragma unroll
                    for (int t = 0; t < NT; t++) {
                        const uint4 e = lds128(stage[i] + (__byte_perm(vw[t], 0u, sel) << 7));   // row = byte * 128
                        acc[t].x ^= e.x; acc[t].y ^= e.y; acc[t].z ^= e.z; acc[t].w ^= e.w;
                    }
                }
            };
            if (look) {
                switch (tmax_w) {
                case 8: look_up(std::integral_constant<int, (TPF < 8 ? TPF : 8)>()); break;
                case 7: look_up(std::integral_constant<int, (TPF < 7 ? TPF : 7)>()); break;
                case 6: look_up(std::integral_constant<int, (TPF < 6 ? TPF : 6)>()); break;
                case 5: look_up(std::integral_constant<int, (TPF < 5 ? TPF : 5)>()); break;
                case 4: look_up(std::integral_constant<int, (TPF < 4 ? TPF : 4)>()); break;
                case 3: look_up(std::integral_constant<int, (TPF < 3 ? TPF : 3)>()); break;
                case 2: look_up(std::integral_constant<int, (TPF < 2 ? TPF : 2)>()); break;
                case 1: look_up(std::integral_constant<int, 1>()); break;
                default: break;
                }
            }
            ring_pos += SPB;
            if (ring_pos >= NSLOT) { ring_pos -= NSLOT; ring_par ^= 1u; }
            if (!early) {
#pragma unroll
                for (int i = 0; i < SPB; i++) if (g0 + i < G) store_data_half(g0 + i);
            }
        }
        // parity symbols of this row block: 8 lanes of a frame write 128 contiguous bytes per store
        if (st_par) {
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
}

// ---------------------------------------------------------------------------------------------
// Free-running variant: no CTA barrier in the steady state.
//
// What the barrier kernel above spends (B200, 200 000 frames, parts switched off one by one,
// profiles/r2_encoder_parts.txt): an interval with everything but the barriers, the input staging
// and the loop removed costs 0.6 of its 3.8 ms; the table stages alone arrive at 13 TB/s (not a
// limit); the look-ups run at the shared-memory rate in between; and the stores -- 16 MB per unit
// and SM at the SM's 32 B/clk -- add almost exactly their own time: after every barrier all warps
// issue the same instruction mix at once, so loads and stores queue behind each other instead of
// overlapping the way they do in a free-running loop (tools/microbench/store_rate.cu: 8 ld.shared.v4
// per 2 st.global.v4 sustain 24.8 B/clk of stores).
//
// Here warp 31 is a producer: one lane issues the TMA bulk copy of stage n as soon as the 31
// consumer warps have released the ring slot (an `empty` mbarrier per slot, one arrival per warp);
// consumers wait on the slot's `full` mbarrier, look up, release, and drift apart by up to NSLOT
// stages, so at any time some warps load while others store.  124 frame slots x 8 lanes; the frames'
// input bytes are staged per warp as before.  The parity symbols of a unit no longer leave as one
// burst: the running XOR (16 bytes per frame and lane) is parked in a scratch buffer in L2 when the
// unit ends and drained during the next unit, one 4-register word at a time, two stores per
// two-stage step -- the same thin stream as the data half.  Only a CTA's last unit bursts.
constexpr int kM4rFrWarps = 31;
constexpr int kM4rFrSlots = kM4rFrWarps * 4;
constexpr int kM4rFrPitch = kM4rFrSlots + 1;     // input staging rows: lane c writes row c, and c * 125 mod 8 differ -> no bank conflict

template <int TPF, int NSLOT>
__host__ __device__ constexpr size_t encode_m4r_fr_smem_bytes()
{
    return (size_t)NSLOT * kM4rStageBytes + 2 * (size_t)TPF * kM4rFrPitch * 16 + 16 * NSLOT;
}
__host__ __device__ inline size_t encode_m4r_fr_park_bytes(int ctas) { return (size_t)ctas * 8 * 1024 * 16; }

__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}

// Uniform quantities of a launch, precomputed on the host: as kernel parameters they are read
// straight from the constant bank by the instructions that use them and cost no register (the first
// version kept them in per-thread registers, spilled five of them, and spent 45 % of its warp
// samples waiting for the reloads behind its own stores in the load/store queue).
struct M4rFrRun {
    uint32_t tile_frames, n_frames, n_units;
    uint32_t G, RB;                  // byte positions (stages per unit), row blocks
    uint32_t sys_words;              // 4-byte data words of a frame written out by one unit
    uint32_t n_pairs;                // sys_words * frames per slot
    uint32_t in_wstride;             // input words per frame
    uint32_t out_qstride;            // float4 per frame
    uint32_t in_tstride;             // input words between frames t and t + 1 of a slot (124 frames)
    uint32_t out_tstride;            // float4 between frames t and t + 1 of a slot
    uint32_t half_m;                 // float4 offset of the data half inside a frame
    uint32_t grid_tiles, grid_rbs;   // gridDim.x / RB, gridDim.x % RB: one CTA step in (tile, row block)
    uint32_t drain;                  // 1: park + drain the parity (G == 64 * TPF), 0: burst at the end of every unit
};

template <int TPF, int NSLOT>
__global__ void __launch_bounds__(kM4rThreads, 1)
encode_m4r_fr_kernel(const uint32_t *__restrict__ inw, float4 *__restrict__ outq, const unsigned char *__restrict__ T,
                     const M4rFrRun run, uint32_t *__restrict__ park)
{
    static_assert(NSLOT % 2 == 0 && NSLOT <= 6 && TPF <= 8, "two stages per step; frame masks are bytes");
    extern __shared__ __align__(128) unsigned char m4r_smem[];
    constexpr uint32_t kStage = (uint32_t)kM4rStageBytes;
    constexpr uint32_t kInBuf = (uint32_t)TPF * kM4rFrPitch * 16;
    const uint32_t smem0 = smem_u32(m4r_smem);
    const uint32_t in_s = smem0 + NSLOT * kStage;
    const uint32_t full = in_s + 2 * kInBuf, empty = full + 8u * NSLOT;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int b = 0; b < NSLOT; b++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(full + 8u * b));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(empty + 8u * b), "r"(kM4rFrWarps));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if ((threadIdx.x >> 5) == kM4rFrWarps) {                       // ---- producer: table stages, in order, as slots free up
        if ((threadIdx.x & 31) == 0) {
            uint32_t pos = 0, use = 0;                             // ring slot and how often it has been filled
            for (uint32_t unit = blockIdx.x; unit < run.n_units; unit += gridDim.x) {
                const uint32_t rb = unit % run.RB;
                const unsigned char *Trb = T + (size_t)rb * run.G * kStage;
                for (uint32_t g = 0; g < run.G; g++) {
                    if (use > 0) mbar_wait(empty + 8u * pos, (use - 1u) & 1u);
                    const uint32_t bar = full + 8u * pos;
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(kStage) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 :: "r"(smem0 + pos * kStage), "l"(Trb + (size_t)g * kStage), "r"(kStage), "r"(bar) : "memory");
                    if (++pos == NSLOT) { pos = 0; use++; }
                }
            }
        }
        return;
    }

    // ---- consumers: slot = tid / 8 (124 of them), lane c = tid % 8 owns 16 bytes of the slot's running XORs.
    // Register budget (64 with 1024 threads): 32 for the running XORs, 8 for the input bytes, 8 for two
    // table rows in flight.  Per-frame offsets are therefore not kept but recomputed from the unit index
    // where they are used (a few integer multiplies per step; `opaque` stops the compiler from hoisting
    // them back into registers that it then spills).
    const uint32_t c = threadIdx.x & 7u;
    uint32_t pos = 0, par = 0;
    uint32_t tm = 0;                                               // byte 0: frames of this slot (tmax), byte 1: of the parked unit, byte 2: most in the warp
    uint32_t dw = 0, dnext = 0;                                    // parity word being drained, the one after it
    auto opaque = [](uint32_t v) { asm volatile("" : "+r"(v)); return v; };
    auto frame0 = [&](uint32_t tile) {                             // first frame of this slot in a tile
        return tile * run.tile_frames + (opaque(threadIdx.x) >> 3);
    };
    // (tile, row block) of the unit and of the CTA's previous unit, advanced without a division
    uint32_t tile = blockIdx.x / run.RB, rb = blockIdx.x % run.RB, ptile = 0, prb = 0;

    for (uint32_t unit = blockIdx.x; unit < run.n_units; unit += gridDim.x) {
        {
            const uint32_t slot = threadIdx.x >> 3;
            const uint32_t tile_f0 = tile * run.tile_frames;
            const uint32_t tile_n = min(run.tile_frames, run.n_frames - tile_f0);
            const uint32_t tmax = slot >= tile_n ? 0u : min((uint32_t)TPF, (tile_n - slot + kM4rFrSlots - 1) / kM4rFrSlots);
            const uint32_t tmax_w = __shfl_sync(0xffffffffu, tmax, 0);
            tm = (tm & 0xff00u) | tmax | (tmax_w << 16);
        }
        auto request_input = [&](uint32_t chunk) {                 // lane c stages 16 bytes of frame t = c
            if (c < (tm & 0xffu)) {
                const uint32_t dst = in_s + (chunk & 1u) * kInBuf + (c * kM4rFrPitch + (threadIdx.x >> 3)) * 16u;
                const uint32_t *src = inw + (frame0(tile) * run.in_wstride + c * run.in_tstride + 4u * chunk);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
            }
            cp_async_commit();
        };
        __syncwarp();                                              // the warp has read the previous unit's last input chunk
        request_input(0);
        uint4 acc[TPF];
#pragma unroll
        for (int t = 0; t < TPF; t++) acc[t] = make_uint4(0u, 0u, 0u, 0u);
        uint32_t sysw = (tm & 0xffu) ? __ldg(inw + (frame0(tile) * run.in_wstride + rb * run.sys_words)) : 0u;   // data word written out next
        uint32_t vw[TPF];
#pragma unroll
        for (int t = 0; t < TPF; t++) vw[t] = 0u;
        const uint32_t *mypark = park + (size_t)blockIdx.x * (8 * 1024 * 4) + threadIdx.x;   // word (t, i) at + (4 t + i) * 1024
        if (tm & 0xff00u) dnext = __ldcg(mypark);

        for (uint32_t g0 = 0; g0 < run.G; g0 += 2) {
            if ((g0 & 15u) == 0) {
                cp_async_wait<0>();
                __syncwarp();
                if (g0 + 16 < run.G) request_input((g0 >> 4) + 1);
            }
            if ((g0 & 3u) == 0) {
                const uint32_t ia = in_s + ((g0 >> 4) & 1u) * kInBuf + (threadIdx.x >> 3) * 16u + (g0 & 12u);
#pragma unroll
                for (int t = 0; t < TPF; t++) vw[t] = lds32(ia + (uint32_t)t * kM4rFrPitch * 16u);
            }
            mbar_wait(full + 8u * pos, par);
            mbar_wait(full + 8u * pos + 8u, par);
            const uint32_t stage = smem0 + 16u * c + pos * kStage;
            auto look_up = [&](auto nt_c) {
                constexpr int NT = decltype(nt_c)::value;
#pragma unroll
                for (int i = 0; i < 2; i++) {
                    const uint32_t sel = 0x4440u | ((g0 + i) & 3u);
#pragma unroll
                    for (int t = 0; t < NT; t++) {
                        const uint4 e = lds128(stage + i * kStage + (__byte_perm(vw[t], 0u, sel) << 7));
                        acc[t].x ^= e.x; acc[t].y ^= e.y; acc[t].z ^= e.z; acc[t].w ^= e.w;
                    }
                }
            };
            switch (tm >> 16) {
            case 8: look_up(std::integral_constant<int, (TPF < 8 ? TPF : 8)>()); break;
            case 7: look_up(std::integral_constant<int, (TPF < 7 ? TPF : 7)>()); break;
            case 6: look_up(std::integral_constant<int, 6>()); break;
            case 5: look_up(std::integral_constant<int, 5>()); break;
            case 4: look_up(std::integral_constant<int, 4>()); break;
            case 3: look_up(std::integral_constant<int, 3>()); break;
            case 2: look_up(std::integral_constant<int, 2>()); break;
            case 1: look_up(std::integral_constant<int, 1>()); break;
            default: break;
            }
            __syncwarp();                                          // every lane's loads of the two slots have returned
            if ((threadIdx.x & 31) == 0) { mbar_arrive(empty + 8u * pos); mbar_arrive(empty + 8u * pos + 8u); }
            pos += 2;
            if (pos == NSLOT) { pos = 0; par ^= 1u; }

            // data half: (frame, word) pair g0 / 2 = 32 symbols, lane c writes float4 c and c + 8
            {
                const uint32_t k = g0 >> 1;
                if (k < run.n_pairs) {
                    const uint32_t t = k % TPF;
                    if (t < (tm & 0xffu)) {
                        const uint32_t sw = rb * run.sys_words + k / TPF;
                        const uint32_t b0 = 8 * (c >> 2) + 7 - 2 * (c & 3);    // symbol j of the word is bit 8 (j / 8) + 7 - j % 8
                        float4 *dst = outq + (frame0(tile) * run.out_qstride + t * run.out_tstride + run.half_m + 16u * sw + c);
                        __stcs(dst, make_float4(bpsk_bit(sysw, b0), 0.f, bpsk_bit(sysw, b0 - 1), 0.f));
                        __stcs(dst + 8, make_float4(bpsk_bit(sysw, b0 + 16), 0.f, bpsk_bit(sysw, b0 + 15), 0.f));
                    }
                    if (k + 1 < run.n_pairs) {
                        const uint32_t t2 = (k + 1) % TPF;
                        if (t2 < (tm & 0xffu))
                            sysw = __ldg(inw + (frame0(tile) * run.in_wstride + t2 * run.in_tstride + rb * run.sys_words + (k + 1) / TPF));
                    }
                }
            }
            // parity of the previous unit: frame g0 / 64 of the slot, pieces g0 % 64 and + 1 of its 64 (piece j =
            // word j / 16, bits 2 (j % 16) and + 1).  The word is shifted down as it is used; the next one is
            // fetched from the parking area in L2 eight steps ahead of its first use.
            {
                const uint32_t td = g0 >> 6;
                if (td < ((tm >> 8) & 0xffu)) {
                    if ((g0 & 15u) == 0) {
                        dw = dnext;
                        const uint32_t nx = (g0 >> 4) + 1;                   // next word: (frame, word) = (nx / 4, nx % 4)
                        if ((nx >> 2) < ((tm >> 8) & 0xffu)) dnext = __ldcg(mypark + nx * 1024);
                    }
                    float4 *dst = outq + (frame0(ptile) * run.out_qstride + td * run.out_tstride +
                                          prb * (kM4rRows / 2) + c + 8u * (g0 & 63u));
                    __stcs(dst, make_float4(bpsk_bit(dw, 0), 0.f, bpsk_bit(dw, 1), 0.f));
                    __stcs(dst + 8, make_float4(bpsk_bit(dw, 2), 0.f, bpsk_bit(dw, 3), 0.f));
                    dw >>= 4;
                }
            }
        }
        const bool last = unit + gridDim.x >= run.n_units;
        if (!last && run.drain) {
            // park the running XOR; it is drained by this thread during the next unit
            uint32_t *dstp = park + (size_t)blockIdx.x * (8 * 1024 * 4) + threadIdx.x;
#pragma unroll
            for (int t = 0; t < TPF; t++) {
                if (t < (int)(tm & 0xffu)) {
                    __stcg(dstp + (4 * t + 0) * 1024, acc[t].x); __stcg(dstp + (4 * t + 1) * 1024, acc[t].y);
                    __stcg(dstp + (4 * t + 2) * 1024, acc[t].z); __stcg(dstp + (4 * t + 3) * 1024, acc[t].w);
                }
            }
            tm = (tm & 0xffu) << 8;
        } else {
#pragma unroll
            for (int t = 0; t < TPF; t++) {
                if (t >= (int)(tm & 0xffu)) continue;
                float4 *dst = outq + (frame0(tile) * run.out_qstride + (uint32_t)t * run.out_tstride + rb * (kM4rRows / 2) + c);
                const uint32_t w4[4] = {acc[t].x, acc[t].y, acc[t].z, acc[t].w};
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int s2 = 0; s2 < 16; s2++)
                        __stcs(dst + 8 * (16 * i + s2),
                               make_float4(bpsk_bit(w4[i], 2 * s2), 0.f, bpsk_bit(w4[i], 2 * s2 + 1), 0.f));
            }
            tm = 0;
        }
        ptile = tile; prb = rb;
        tile += run.grid_tiles; rb += run.grid_rbs;
        if (rb >= run.RB) { rb -= run.RB; tile++; }
    }
}

}  // namespace ldpc535
