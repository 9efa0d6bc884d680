// Generic decoder kernels (any parity-check matrix):
//   decode_warp_kernel   one warp per codeword, M <= 32 checks, N <= 64 bits.  Lane j is
//                        check j (and bits j, j+32); hard decisions and syndromes travel
//                        by warp ballot, messages in a 768-byte shared-memory strip.
//   decode_block_kernel  one CTA per codeword (n = 8192 class codes): all messages stay
//                        in shared memory for the whole decode; adjacency tables are
//                        staged into shared memory once per CTA with a TMA bulk copy.
// Both fuse: complex symbol -> LLR (r = -pol * Re), the iterations, the syndrome test for
// early termination, the saturating syndrome weight and MSB-first byte packing
// (lib/ldpc_decoder_cb_impl.cc:149-153, :478-557, :236-253, :207-219).
#pragma once
#include <type_traits>

#include "spa_math.cuh"

namespace ldpc535 {

enum { kMethodMinSum = 0, kMethodSpa = 1, kMethodBitFlip = 2, kMethodHard = 3 };

struct DecodeParams {
    const float2 *sym;            // complex symbols (re, im) ...
    const float *sym_re;          // ... or, when non-null, real parts only (host-packed staging)
    long long n_sym;
    const long long *win_offset;  // or nullptr: window w starts at w * N
    const signed char *polarity;  // or nullptr: +1
    long long n_win;
    int M, N, K, nbytes, E;
    int max_iters, early_stop, thr;
    const uint16_t *chk_var;      // [DC][M]
    const uint16_t *var_slot;     // [DV][N]
    const uint8_t *chk_deg;       // [M]
    uint8_t *out_bytes, *out_synd, *out_iters;
    // message dumps (DEBUG instantiations only)
    float *dbgL, *dbgE, *dbgM;
    float *dbgS;                  // scratch [n_win][E]: latest bit->check messages M (the kernels
                                  // store them as copysign(exp(-|M|), M), see spa_math.cuh)
    const int32_t *slot_edge;     // [DC][M] -> CSR edge id or -1
    // warp kernel: conflict-free strip layout (code_tables.cpp: color_warp_layout)
    const uint16_t *w_chk_pos;    // [DC][32]
    const uint16_t *w_var_pos;    // [DV][64]
    const int32_t *w_pos_edge;    // [DC*32] -> CSR edge id or -1
    // block kernel: table staging
    int stage_tables;             // 1: copy chk_var / var_slot into shared memory
    int tabA_bytes, tabB_bytes;   // padded to 16 B
    // register-table regular kernel: windows are handed out by an atomic cursor (starts at 0)
    unsigned int *cursor;
    // "c4-adaptive": two kernel families are queued for the same windows and a device-side word
    // written by pick_family_kernel says which one runs; the other returns at once
    const int *select;            // or nullptr: run
    int select_want;
};

// Device-side choice between the lock-step thread-per-codeword kernel and the warp-per-codeword
// kernel for an early-stop batch (both give identical outputs; only the time differs).  A lock-step
// CTA runs until the slowest of its 256 codewords stops, a warp until its own codeword stops.  From
// the iteration counts of a decoded sample: q_k = share of codewords that ran >= k iterations;
//   lock step:  t0 + t1 * sum_k (1 - (1 - q_k)^256)      warp:  u0 + u1 * sum_k q_k
// with the per-batch constants measured on B200 (us per 256 codewords per SM; lock step: 12.71 ms per 2 M codewords at
// 50 iterations and 1.455 ms at 5, profiles/r2_node_rate2.txt;
// warp: fit to 35.6 / 70.4 / 104.1 Gbit/s at 4.73 / 2.19 / 1.27 mean iterations, profiles/r2_warp_layout.txt).
__global__ void __launch_bounds__(256) pick_family_kernel(const uint8_t *iters, int n, int max_iters, int *select)
{
    __shared__ int hist[256];
    hist[threadIdx.x] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += 256) atomicAdd(&hist[iters[i]], 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        int valid = 0;
        for (int k = 1; k <= max_iters && k < 255; k++) valid += hist[k];      // 255 marks a refused window
        float lock = 0.f, mean = 0.f;
        int ge = valid;                                                        // codewords with iters >= k
        for (int k = 1; k <= max_iters && k < 255; k++) {
            const float q = valid ? (float)ge / (float)valid : 1.f;
            mean += q;
            lock += 1.f - __powf(1.f - q, 256.f);
            ge -= hist[k];
        }
        const float t_lock = 3.6f + 4.77f * lock, t_warp = 3.4f + 6.46f * mean;
        *select = (t_lock < t_warp) ? 1 : 0;
    }
}

constexpr int kWarpKernelThreads = 128;
#ifndef WARP_MIN_BLOCKS
#define WARP_MIN_BLOCKS 8
#endif

// decode_block_kernel shared-memory layout (host and device must agree):
//   [0,16) mbarrier | msg[DC*M] f32 | r[N] f32 | hard[ceil(N/32)] | par[ceil(M/32)] | red[4]
//   | (16-aligned) chk_var copy | var_slot copy
__host__ __device__ inline size_t block_smem_fixed_bytes(int dc, int M, int N, int msg_size = 4)
{
    size_t b = 16 + (size_t)msg_size * dc * M + 4 * (size_t)N + 4 * (size_t)((N + 31) >> 5) +
               4 * (size_t)((M + 31) >> 5) + 16;
    return (b + 15) & ~(size_t)15;
}
constexpr float kInf = __builtin_huge_valf();

// stored form of a bit->check message: sum-product keeps copysign(exp(-|M|), M) (spa_math.cuh),
// the other methods the message itself; a padded slot holds the identity of the check update
template <int METHOD, typename T>
__device__ __forceinline__ T enc_msg(T M)
{
    if constexpr (METHOD == kMethodSpa) return to_check_msg(M); else return M;
}
template <int METHOD, typename T>
__device__ __forceinline__ T pad_msg()
{
    if constexpr (METHOD == kMethodSpa) return (T)0; else return (T)kInf;
}

// Re(symbol i) of either input layout
__device__ __forceinline__ void cp_async_f32(float *smem_dst, const float *gmem_src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;"
                 :: "r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}

__device__ __forceinline__ float load_re(const DecodeParams &p, long long i)
{
    return p.sym_re ? __ldg(p.sym_re + i) : __ldg(&p.sym[i].x);
}

// message type of a method: fp64 for min-sum (exact, see spa_math.cuh), fp32 otherwise
template <int METHOD>
using msg_t = std::conditional_t<METHOD == kMethodMinSum, double, float>;

__device__ __forceinline__ uint8_t pack_msb_first(uint32_t bits8)
{
    return (uint8_t)(__brev(bits8 & 0xffu) >> 24);
}

// ---------------------------------------------------------------------------------
// warp per codeword
// ---------------------------------------------------------------------------------
// 8 CTAs (32 warps) per SM for the sparse instantiation; the dense one (check degree <= 16, bit degree
// <= 8) holds 16 + 2 x 8 messages per lane and spilled ~500 bytes at 64 registers: 4 CTAs per SM there.
template <int METHOD, int DC, int DV, bool DEBUG>
__global__ void __launch_bounds__(kWarpKernelThreads, (DC <= 6 ? WARP_MIN_BLOCKS : WARP_MIN_BLOCKS / 2))
decode_warp_kernel(const DecodeParams p)
{
    using T = msg_t<METHOD>;
    extern __shared__ __align__(16) unsigned char smem_w[];
    if (p.select && *p.select != p.select_want) return;      // the other family was picked for these windows
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    // Strip of this warp (layout: code_tables.cpp color_warp_layout): DC rows of 32 messages, one row
    // of dummies (writes of unused slots), one row holding the check update's identity and one row of
    // zeros (what unused check / bit slots read).  An unused slot reads and writes the word of ITS row
    // in a bank the access's real edges leave free, so every access of the loop is one wavefront
    // (round 1 read a single identity / zero word: a 2-way conflict on almost every access of the
    // shipped code, whose degrees are uneven).  fp64 messages (min-sum) are kept as two planes of
    // 32-bit words, low halves then high halves: an 8-byte word per lane made every access 4
    // wavefronts with the 32-bank colouring, two 4-byte accesses are 2.
    constexpr int kRows = DC + 3, kDummy = DC * 32, kPadRow = (DC + 1) * 32, kZeroRow = (DC + 2) * 32;
    constexpr bool k64 = sizeof(T) == 8;
    uint32_t *strip = reinterpret_cast<uint32_t *>(smem_w) + warp * ((k64 ? 2 : 1) * kRows * 32);
    auto ld = [&](int pos) -> T {
        if constexpr (k64) return (T)__hiloint2double((int)strip[kRows * 32 + pos], (int)strip[pos]);
        else return (T)__uint_as_float(strip[pos]);
    };
    auto st = [&](int pos, T v) {
        if constexpr (k64) { strip[pos] = (uint32_t)__double2loint((double)v); strip[kRows * 32 + pos] = (uint32_t)__double2hiint((double)v); }
        else strip[pos] = __float_as_uint((float)v);
    };
    const int M = p.M, N = p.N;

    // ---- this lane's rows/columns of H (loop invariant, registers) ----
    uint32_t row_lo = 0, row_hi = 0;                 // bits of check `lane`
    int cdeg = 0, cpos[DC];                          // strip position of slot s of this lane's check
#pragma unroll
    for (int s = 0; s < DC; s++) {
        const int v = (lane < M) ? p.chk_var[s * M + lane] : 0xFFFF;
        cpos[s] = p.w_chk_pos[s * 32 + lane];
        if (v != 0xFFFF) {
            cdeg = s + 1;
            if (v < 32) row_lo |= 1u << v; else row_hi |= 1u << (v - 32);
        }
    }
    int vpos[2][DV], vchk[2][DV], vdeg[2];
#pragma unroll
    for (int t = 0; t < 2; t++) {
        const int v = lane + 32 * t;
        vdeg[t] = 0;
#pragma unroll
        for (int k = 0; k < DV; k++) {
            const int idx = (v < N) ? p.var_slot[k * N + v] : 0xFFFF;
            vpos[t][k] = p.w_var_pos[k * 64 + v];       // real edge: strip position; unused slot: 0x8000 | free bank
            vchk[t][k] = 0;
            if (idx != 0xFFFF) {
                vchk[t][k] = idx % M;
                vdeg[t] = k + 1;
            }
        }
    }
    // Unused slots cost no predicates in the loop: a padded check slot READS the identity row and
    // writes its own (never read) position; an unused bit slot reads the zero row and writes the dummy row.
    int crd[DC], vrd[2][DV], vwr[2][DV];
#pragma unroll
    for (int s = 0; s < DC; s++) crd[s] = (s < cdeg) ? cpos[s] : kPadRow + (cpos[s] & 31);
#pragma unroll
    for (int t = 0; t < 2; t++)
#pragma unroll
        for (int k = 0; k < DV; k++) {
            vrd[t][k] = (k < vdeg[t]) ? vpos[t][k] : kZeroRow + (vpos[t][k] & 31);
            vwr[t][k] = (k < vdeg[t]) ? vpos[t][k] : kDummy + (vpos[t][k] & 31);
        }
    st(kPadRow + lane, pad_msg<METHOD, T>());
    st(kZeroRow + lane, (T)0);
    __syncwarp();

    const long long warps_total = (long long)gridDim.x * (blockDim.x >> 5);
    // The symbols of a window are fetched one window AHEAD with cp.async into a 64-word strip of the
    // warp (no registers held across the iterations): a warp that loaded and then used them sat out
    // the whole HBM latency once per codeword -- a fifth of all warp samples at 5 iterations (ncu,
    // profiles/r2c_*).
    //
    // A warp takes BLOCKS of 32 consecutive windows (round 2; it used to take every warps_total-th
    // one).  With early stop a codeword of the shipped code runs 1 to 5 iterations of ~120
    // instructions and paid ~200 more for its window offset, bounds test, 64-bit addresses and three
    // single-byte stores (half of all instructions at 6 dB).  Per block, lane j now loads offset and
    // polarity of window j once (coalesced), keeps the results of window j, and the block's outputs
    // leave as one coalesced store per array.
    // Addresses are kept, not recomputed: the lane's word of the staging strip as a 32-bit shared
    // address, the lane's symbol of the window to fetch next as a pointer that advances by one frame
    // (aligned frames) or is rebuilt from the shuffled offset (search windows).  `opaque` keeps the
    // compiler from rematerialising them from tid and the kernel parameters at every use (it did:
    // ~50 instructions per prefetch).
    auto opaque = [](uint32_t v) { asm volatile("" : "+r"(v)); return v; };
    const uint32_t stage_a = opaque((uint32_t)__cvta_generic_to_shared(
        reinterpret_cast<float *>(smem_w + sizeof(T) * (blockDim.x >> 5) * (kRows * 32)) + warp * 64 + lane));
    // result strip of the warp: 32 x (bytes lo, bytes hi, syndrome weight, iterations), behind the staging strips
    const uint32_t res_a = opaque((uint32_t)__cvta_generic_to_shared(
        smem_w + sizeof(T) * (blockDim.x >> 5) * (kRows * 32) + sizeof(float) * 64 * (blockDim.x >> 5) + 512 * warp));
    const bool packed = p.sym_re != nullptr;                  // real parts only (4 bytes per symbol) or complex (8)
    const char *gbase = packed ? reinterpret_cast<const char *>(p.sym_re) : reinterpret_cast<const char *>(p.sym);
    const uint32_t esz = packed ? 4u : 8u;
    auto fetch_at = [&](const char *g, bool okq) {           // g = this lane's first symbol of the window
        if (okq) {
            if (lane < N)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(stage_a), "l"(g) : "memory");
            if (lane + 32 < N)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(stage_a + 128u), "l"(g + 32u * esz) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto window_of = [&](long long wq, long long &offq, bool &okq) {
        okq = wq < p.n_win;
        offq = !okq ? 0 : p.win_offset ? p.win_offset[wq] : wq * (long long)N;
        okq = okq && offq >= 0 && offq + N <= p.n_sym;
    };
    const long long n_blocks = (p.n_win + 31) >> 5;
    const long long b0 = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    const char *gnext;                                        // this lane's first symbol of the window fetched next
    {
        long long o; bool k;
        window_of(b0 << 5, o, k);
        gnext = gbase + (o + lane) * (long long)esz;
        fetch_at(gnext, k);
    }
    for (long long blk = b0; blk < n_blocks; blk += warps_total) {
      const long long w0 = blk << 5;
      const int cnt = (int)min(32ll, p.n_win - w0);
      long long off_l; bool ok_l;                       // of window w0 + lane
      window_of(lane < cnt ? w0 + lane : p.n_win, off_l, ok_l);
      const float pol_l = (p.polarity && lane < cnt) ? (float)p.polarity[w0 + lane] : 1.f;
      const uint32_t okmask = __ballot_sync(0xffffffffu, ok_l);
      for (int j = 0; j < cnt; j++) {
        const long long w = w0 + j;
        const bool ok = (okmask >> j) & 1u;
        constexpr float kIn = (METHOD == kMethodSpa) ? kSpaScale : 1.f;      // SPA works in units of ln 2
        constexpr float kOut = (METHOD == kMethodSpa) ? kSpaUnscale : 1.f;   // (message dumps only)
        const float scale = p.polarity ? -kIn * __shfl_sync(0xffffffffu, pol_l, j) : -kIn;
        asm volatile("cp.async.wait_group 0;" ::: "memory");    // each lane reads back its own two words
        T r[2];
        {
            float s0, s1;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(s0) : "r"(stage_a));
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(s1) : "r"(stage_a + 128u));
            r[0] = (T)((ok && lane < N) ? __fmul_rn(scale, s0) : 0.f);
            r[1] = (T)((ok && lane + 32 < N) ? __fmul_rn(scale, s1) : 0.f);
        }
        if (j + 1 < cnt) {
            if (p.win_offset) gnext = gbase + (__shfl_sync(0xffffffffu, off_l, j + 1) + lane) * (long long)esz;
            else gnext += (unsigned)N * esz;
            fetch_at(gnext, (okmask >> (j + 1)) & 1u);
        } else {
            long long o; bool k;
            window_of((blk + warps_total) << 5, o, k);
            gnext = gbase + (o + lane) * (long long)esz;
            fetch_at(gnext, k);
        }
        float rk[2][DV];                               // r on a real edge of the bit, 0 on an unused slot
#pragma unroll
        for (int t = 0; t < 2; t++)
#pragma unroll
            for (int k = 0; k < DV; k++) rk[t][k] = (k < vdeg[t]) ? (float)r[t] : 0.f;
        uint32_t hard0 = 0, hard1 = 0, bad = 0;
        int iters = 0;
        bool broke = false;

        if (METHOD == kMethodHard || METHOD == kMethodBitFlip) {
            // decodeHard / decodeBitFlipping's prior: rx < 0 -> 0 else 1, rx = -r
            const bool y0 = (lane < N) && !(r[0] > (T)0);
            const bool y1 = (lane + 32 < N) && !(r[1] > (T)0);
            hard0 = __ballot_sync(0xffffffffu, y0);
            hard1 = __ballot_sync(0xffffffffu, y1);
            bad = __ballot_sync(0xffffffffu, (__popc(row_lo & hard0) + __popc(row_hi & hard1)) & 1);
            if (METHOD == kMethodBitFlip) {
                // lib/ldpc_decoder_cb_impl.cc:439-473.  E(i,j) = parity_i ^ ci_j on an edge;
                // a bit is set to !y when more than M/2 of its checks disagree with y.
                bool c0 = y0, c1 = y1;
                for (int h = 0; h < p.max_iters; h++) {
                    int d0 = 0, d1 = 0;
#pragma unroll
                    for (int k = 0; k < DV; k++) {
                        if (k < vdeg[0]) d0 += (int)((((bad >> vchk[0][k]) & 1u) ^ (uint32_t)c0) != (uint32_t)y0);
                        if (k < vdeg[1]) d1 += (int)((((bad >> vchk[1][k]) & 1u) ^ (uint32_t)c1) != (uint32_t)y1);
                    }
                    if (d0 > M / 2) c0 = !y0;
                    if (d1 > M / 2) c1 = !y1;
                    hard0 = __ballot_sync(0xffffffffu, c0 && lane < N);
                    hard1 = __ballot_sync(0xffffffffu, c1 && lane + 32 < N);
                    bad = __ballot_sync(0xffffffffu, (__popc(row_lo & hard0) + __popc(row_hi & hard1)) & 1);
                    if (h + 1 < p.max_iters && bad == 0) break;
                }
            }
        } else {
            __syncwarp();
            // M_ji = r_i on every edge (lib/ldpc_decoder_cb_impl.cc:489-496); unused slots write the dummy row
#pragma unroll
            for (int t = 0; t < 2; t++) {
                const T e0 = enc_msg<METHOD, T>(r[t]);
#pragma unroll
                for (int k = 0; k < DV; k++) {
                    if (DEBUG && p.dbgS && k < vdeg[t])
                        p.dbgS[w * p.E + p.w_pos_edge[vpos[t][k]]] = (float)r[t] * kOut;
                    st(vwr[t][k], e0);
                }
            }
            iters = p.max_iters;
            for (int h = 0; h < p.max_iters; h++) {
                __syncwarp();
                T m[DC];
#pragma unroll
                for (int s = 0; s < DC; s++) m[s] = ld(crd[s]);
                if (DEBUG && p.dbgM && lane < M) {
#pragma unroll
                    for (int s = 0; s < DC; s++)
                        if (s < cdeg) {
                            const long long ei = w * p.E + p.slot_edge[s * M + lane];
                            p.dbgM[ei] = p.dbgS[ei];             // M entering this check step
                        }
                }
                if constexpr (METHOD == kMethodSpa) check_node_spa<DC>(m); else check_node_minsum<DC, T>(m);
#pragma unroll
                for (int s = 0; s < DC; s++) st(cpos[s], m[s]);
                if (DEBUG && p.dbgE && lane < M) {
#pragma unroll
                    for (int s = 0; s < DC; s++)
                        if (s < cdeg) p.dbgE[w * p.E + p.slot_edge[s * M + lane]] = (float)m[s] * kOut;
                }
                __syncwarp();
                T x[2][DV], L[2];
#pragma unroll
                for (int t = 0; t < 2; t++) {
#pragma unroll
                    for (int k = 0; k < DV; k++) x[t][k] = ld(vrd[t][k]);
                    if constexpr (METHOD == kMethodSpa) L[t] = var_node_spa_rk<DV>(x[t], rk[t]);
                    else L[t] = var_node_minsum<DV, T>(x[t], DV, r[t]);
                }
                // SPA decides 1 on L <= 0 (:527), min-sum on LQ < 0 (:398)
                const bool b0 = (lane < N) && (METHOD == kMethodSpa ? (L[0] <= (T)0) : (L[0] < (T)0));
                const bool b1 = (lane + 32 < N) && (METHOD == kMethodSpa ? (L[1] <= (T)0) : (L[1] < (T)0));
                hard0 = __ballot_sync(0xffffffffu, b0);
                hard1 = __ballot_sync(0xffffffffu, b1);
                bad = __ballot_sync(0xffffffffu, (__popc(row_lo & hard0) + __popc(row_hi & hard1)) & 1);
                if (DEBUG && p.dbgL) {
                    if (lane < N) p.dbgL[w * N + lane] = (float)L[0] * kOut;
                    if (lane + 32 < N) p.dbgL[w * N + lane + 32] = (float)L[1] * kOut;
                }
                // SPA tests every iteration, the last included (:535); min-sum skips the
                // test on the last one (:406)
                const bool test = p.early_stop && (METHOD == kMethodSpa || h + 1 < p.max_iters);
                if (test && bad == 0) { iters = h + 1; broke = true; break; }
#pragma unroll
                for (int t = 0; t < 2; t++)
#pragma unroll
                    for (int k = 0; k < DV; k++)
                    {
                        if (DEBUG && p.dbgS && k < vdeg[t])
                            p.dbgS[w * p.E + p.w_pos_edge[vpos[t][k]]] = (float)x[t][k] * kOut;
                        st(vwr[t][k], enc_msg<METHOD, T>(x[t][k]));
                    }
            }
            // dbgM so far holds the M that ENTERED the last check step -- what the reference
            // still holds after a successful test; without one the last Step 2 counts too.
            if (DEBUG && p.dbgM && !broke) {
                __syncwarp();
                if (lane < M) {
#pragma unroll
                    for (int s = 0; s < DC; s++)
                        if (s < cdeg) {
                            const long long ei = w * p.E + p.slot_edge[s * M + lane];
                            p.dbgM[ei] = p.dbgS[ei];
                        }
                }
            }
        }

        // ---- results of window j: warp-uniform, parked in the warp's result strip by one lane ----
        if (!ok) { hard0 = hard1 = 0; bad = 0xffffffffu; iters = 255; }
        if (lane == 0) {
            const unsigned long long data = ((unsigned long long)hard0 | ((unsigned long long)hard1 << 32)) >> M;   // bits M .. N-1
            // byte i of the output = bits 8 i .. 8 i + 7 of `data`, MSB first: reverse the bits of every byte
            uint4 rv;
            rv.x = __byte_perm(__brev((uint32_t)data), 0u, 0x0123);
            rv.y = __byte_perm(__brev((uint32_t)(data >> 32)), 0u, 0x0123);
            rv.z = (uint32_t)min(__popc(bad), p.thr + 1);
            rv.w = (uint32_t)min(iters, 255);
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(res_a + 16u * j), "r"(rv.x), "r"(rv.y), "r"(rv.z), "r"(rv.w) : "memory");
        }
      }
      // ---- outputs of the block: lane l stores window w0 + l ----
      __syncwarp();
      if (lane < cnt) {
          uint4 rv;
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(rv.x), "=r"(rv.y), "=r"(rv.z), "=r"(rv.w) : "r"(res_a + 16u * lane));
          if (p.nbytes == 4) {
              reinterpret_cast<uint32_t *>(p.out_bytes)[w0 + lane] = rv.x;           // out_bytes is 4-byte aligned (ldpc535.h)
          } else {
              const unsigned long long rb = (unsigned long long)rv.x | ((unsigned long long)rv.y << 32);
              for (int i = 0; i < p.nbytes; i++) p.out_bytes[(w0 + lane) * p.nbytes + i] = (uint8_t)(rb >> (8 * i));
          }
          if (p.out_synd) p.out_synd[w0 + lane] = (uint8_t)rv.z;
          if (p.out_iters) p.out_iters[w0 + lane] = (uint8_t)rv.w;
      }
      __syncwarp();
    }
}

// ---------------------------------------------------------------------------------
// CTA per codeword
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

// 1-D TMA bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int METHOD, int DC, int DV, bool DEBUG>
__global__ void __launch_bounds__(1024, 1)
decode_block_kernel(const DecodeParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int M = p.M, N = p.N;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
    const int nwords = (N + 31) >> 5;

    // shared layout: bar | msg[DC*M] | r[N] | hard[nwords] | par[(M+31)/32] | red | tabA | tabB
    using T = msg_t<METHOD>;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    T *msg = reinterpret_cast<T *>(smem_raw + 16);
    float *r = reinterpret_cast<float *>(msg + (size_t)DC * M);
    uint32_t *hard = reinterpret_cast<uint32_t *>(r + N);
    uint32_t *par = hard + nwords;
    int *red = reinterpret_cast<int *>(par + ((M + 31) >> 5));
    const size_t tab_off = block_smem_fixed_bytes(DC, M, N, (int)sizeof(T));
    const uint16_t *chk_var = p.chk_var;
    const uint16_t *var_slot = p.var_slot;

    if (p.stage_tables) {
        uint16_t *sA = reinterpret_cast<uint16_t *>(smem_raw + tab_off);
        uint16_t *sB = reinterpret_cast<uint16_t *>(smem_raw + tab_off + p.tabA_bytes);
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(bar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                         :: "r"(smem_u32(bar)), "r"((uint32_t)(p.tabA_bytes + p.tabB_bytes)) : "memory");
            tma_bulk_g2s(sA, p.chk_var, (uint32_t)p.tabA_bytes, bar);
            tma_bulk_g2s(sB, p.var_slot, (uint32_t)p.tabB_bytes, bar);
        }
        // everyone waits for the bytes (phase 0)
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done) : "r"(smem_u32(bar)) : "memory");
        }
        chk_var = sA;
        var_slot = sB;
    }

    for (long long w = blockIdx.x; w < p.n_win; w += gridDim.x) {
        const long long off = p.win_offset ? p.win_offset[w] : w * (long long)N;
        const float pol = p.polarity ? (float)p.polarity[w] : 1.f;
        const bool ok = off >= 0 && off + N <= p.n_sym;
        constexpr float kIn = (METHOD == kMethodSpa) ? kSpaScale : 1.f;
        constexpr float kOut = (METHOD == kMethodSpa) ? kSpaUnscale : 1.f;
        __syncthreads();                                   // previous window fully drained
        for (int i = tid; i < N; i += nt) r[i] = ok ? __fmul_rn(-pol * kIn, load_re(p, off + i)) : 0.f;
        if (tid == 0) red[0] = 0;
        __syncthreads();

        int iters = 0;
        bool clean = false;                                // last syndrome test saw all zeros
        if (METHOD == kMethodHard || METHOD == kMethodBitFlip) {
            for (int base = 0; base < N; base += nt) {
                const int i = base + tid;
                const bool y = (i < N) && !(r[i] > 0.f);
                const uint32_t wd = __ballot_sync(0xffffffffu, y);
                if (lane == 0 && i < N) hard[i >> 5] = wd;
            }
            __syncthreads();
            if (METHOD == kMethodBitFlip) {
                // y stays in r's sign; ci lives in hard[].  One sweep per iteration:
                // parities of the pre-sweep ci, then every bit counts its disagreeing checks.
                for (int h = 0; h < p.max_iters; h++) {
                    int anybad = 0;
                    for (int base = 0; base < M; base += nt) {
                        const int j = base + tid;
                        uint32_t pj = 0;
                        if (j < M) {
#pragma unroll
                            for (int s = 0; s < DC; s++) {
                                const int v = chk_var[s * M + j];
                                if (v != 0xFFFF) pj ^= hard[v >> 5] >> (v & 31);
                            }
                        }
                        const uint32_t wd = __ballot_sync(0xffffffffu, pj & 1u);
                        if (lane == 0 && j < M) par[j >> 5] = wd;
                    }
                    __syncthreads();
                    for (int base = 0; base < N; base += nt) {
                        const int i = base + tid;
                        bool c = false;
                        if (i < N) {
                            const uint32_t y = !(r[i] > 0.f);
                            c = (hard[i >> 5] >> (i & 31)) & 1u;
                            int dis = 0;
#pragma unroll
                            for (int k = 0; k < DV; k++) {
                                const int idx = var_slot[k * N + i];
                                if (idx != 0xFFFF) {
                                    const int j = idx % M;
                                    dis += (int)((((par[j >> 5] >> (j & 31)) & 1u) ^ (uint32_t)c) != y);
                                }
                            }
                            if (dis > M / 2) c = !y;
                        }
                        __syncwarp();
                        const uint32_t wd = __ballot_sync(0xffffffffu, c);
                        // all reads of hard[] by this warp's lanes are of their own word only
                        if (lane == 0 && i < N) hard[i >> 5] = wd;
                    }
                    __syncthreads();
                    // Finished? (not on the last iteration, :470)
                    for (int j = tid; j < M; j += nt) {
                        uint32_t pj = 0;
#pragma unroll
                        for (int s = 0; s < DC; s++) {
                            const int v = chk_var[s * M + j];
                            if (v != 0xFFFF) pj ^= hard[v >> 5] >> (v & 31);
                        }
                        anybad |= (int)(pj & 1u);
                    }
                    anybad = __syncthreads_or(anybad);
                    if (h + 1 < p.max_iters && !anybad) { clean = true; break; }
                }
            }
        } else {
            for (int idx = tid; idx < DC * M; idx += nt) {
                const int v = chk_var[idx];
                if (DEBUG && p.dbgS && v != 0xFFFF) p.dbgS[w * p.E + p.slot_edge[idx]] = r[v] * kOut;
                msg[idx] = (v != 0xFFFF) ? enc_msg<METHOD, T>((T)r[v]) : pad_msg<METHOD, T>();
            }
            __syncthreads();
            iters = p.max_iters;
            // Early stop without a separate syndrome pass (sum-product, non-debug): a bit->check
            // message is stored as t = copysign(2^-|M|, M) with |t| <= 1, so bit 30 of its
            // pattern is always 0 -- the variable phase parks the bit's hard decision there,
            // and the NEXT check phase XORs the six tags of a check into its parity while it
            // strips them.  "Iteration h passed the test" is therefore known one check phase
            // late (its results are discarded), which costs ~0.4 iteration per codeword and
            // saves a pass over the adjacency table and a barrier in EVERY iteration.
            constexpr bool kTag = (METHOD == kMethodSpa) && !DEBUG;
            const bool tag = kTag && p.early_stop;
            for (int h = 0; h < p.max_iters; h++) {
                // ---- check nodes ----
                int tagbad = 0;
                for (int j = tid; j < M; j += nt) {
                    T m[DC];
#pragma unroll
                    for (int s = 0; s < DC; s++) m[s] = msg[s * M + j];
                    if constexpr (kTag) {
                        if (tag) {
                            uint32_t par = 0;
#pragma unroll
                            for (int s = 0; s < DC; s++) {
                                const uint32_t bits = __float_as_uint((float)m[s]);
                                par ^= bits;
                                m[s] = (T)__uint_as_float(bits & ~0x40000000u);
                            }
                            tagbad |= (int)((par >> 30) & 1u);
                        }
                    }
                    const int deg = p.chk_deg[j];
                    if (DEBUG && p.dbgM) {
#pragma unroll
                        for (int s = 0; s < DC; s++)
                            if (s < deg) {
                                const long long ei = w * p.E + p.slot_edge[s * M + j];
                                p.dbgM[ei] = p.dbgS[ei];         // M entering this check step
                            }
                    }
                    if constexpr (METHOD == kMethodSpa) check_node_spa<DC>(m); else check_node_minsum<DC, T>(m);
#pragma unroll
                    for (int s = 0; s < DC; s++)
                        if (s < deg) msg[s * M + j] = m[s];
                    if (DEBUG && p.dbgE) {
#pragma unroll
                        for (int s = 0; s < DC; s++)
                            if (s < deg) p.dbgE[w * p.E + p.slot_edge[s * M + j]] = (float)m[s] * kOut;
                    }
                }
                if (tag) {
                    // tags carry the decisions of iteration h-1 (none before the first one)
                    const int bad = __syncthreads_or(tagbad);
                    if (h > 0 && !bad) { iters = h; clean = true; break; }
                } else {
                    __syncthreads();
                }
                // ---- variable nodes: L, hard decision, next bit->check messages ----
                for (int base = 0; base < N; base += nt) {
                    const int i = base + tid;
                    bool b = false;
                    if (i < N) {
                        T x[DV];
                        int idx[DV], dv = 0;
#pragma unroll
                        for (int k = 0; k < DV; k++) {
                            idx[k] = var_slot[k * N + i];
                            x[k] = (T)0;
                            if (idx[k] != 0xFFFF) { x[k] = msg[idx[k]]; dv = k + 1; }
                        }
                        T L;
                        if constexpr (METHOD == kMethodSpa) L = var_node_spa<DV>(x, dv, r[i]);
                        else L = var_node_minsum<DV, T>(x, dv, (T)r[i]);
                        b = (METHOD == kMethodSpa) ? (L <= (T)0) : (L < (T)0);
                        if (DEBUG && p.dbgL) p.dbgL[w * N + i] = (float)L * kOut;
                        const uint32_t tagbit = (tag && b) ? 0x40000000u : 0u;
#pragma unroll
                        for (int k = 0; k < DV; k++)
                            if (k < dv) {
                                if (DEBUG && p.dbgS) p.dbgS[w * p.E + p.slot_edge[idx[k]]] = (float)x[k] * kOut;
                                T t = enc_msg<METHOD, T>(x[k]);
                                if constexpr (kTag) t = (T)__uint_as_float(__float_as_uint((float)t) | tagbit);
                                msg[idx[k]] = t;
                            }
                    }
                    const uint32_t wd = __ballot_sync(0xffffffffu, b);
                    if (lane == 0 && i < N) hard[i >> 5] = wd;
                }
                __syncthreads();
                // ---- Finished? ----
                const bool test = !tag && p.early_stop && (METHOD == kMethodSpa || h + 1 < p.max_iters);
                if (test) {
                    int anybad = 0;
                    for (int j = tid; j < M; j += nt) {
                        uint32_t pj = 0;
#pragma unroll
                        for (int s = 0; s < DC; s++) {
                            const int v = chk_var[s * M + j];
                            if (v != 0xFFFF) pj ^= hard[v >> 5] >> (v & 31);
                        }
                        anybad |= (int)(pj & 1u);
                    }
                    anybad = __syncthreads_or(anybad);
                    if (!anybad) { iters = h + 1; clean = true; break; }
                }
            }
            // The bit->check messages of a passing iteration have already been formed in
            // place (the reference breaks before its Step 2); outputs cannot tell, and the
            // dump below keeps the reference's view: dbgM = M entering the last check step
            // unless no test passed.
            if (DEBUG && p.dbgM && !clean) {
                for (int j = tid; j < M; j += nt) {
                    const int deg = p.chk_deg[j];
#pragma unroll
                    for (int s = 0; s < DC; s++)
                        if (s < deg) {
                            const long long ei = w * p.E + p.slot_edge[s * M + j];
                            p.dbgM[ei] = p.dbgS[ei];
                        }
                }
            }
        }

        // ---- saturating syndrome weight of the final decision (checkFrame, :236-253) ----
        int cnt = 0;
        if (!clean) {
            for (int j = tid; j < M; j += nt) {
                uint32_t pj = 0;
#pragma unroll
                for (int s = 0; s < DC; s++) {
                    const int v = chk_var[s * M + j];
                    if (v != 0xFFFF) pj ^= hard[v >> 5] >> (v & 31);
                }
                cnt += (int)(pj & 1u);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
            if (lane == 0 && cnt) atomicAdd(&red[0], cnt);
            __syncthreads();
            cnt = red[0];
        }
        // ---- outputs: data bits M .. N-1, MSB first ----
        for (int b = tid; b < p.nbytes; b += nt) {
            uint32_t bits = 0;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int v = M + 8 * b + q;
                if (v < N) bits |= ((hard[v >> 5] >> (v & 31)) & 1u) << q;
            }
            p.out_bytes[w * p.nbytes + b] = ok ? pack_msb_first(bits) : (uint8_t)0;
        }
        if (tid == 0) {
            if (p.out_synd) p.out_synd[w] = ok ? (uint8_t)min(cnt, p.thr + 1) : (uint8_t)255;
            if (p.out_iters) p.out_iters[w] = ok ? (uint8_t)min(iters, 255) : (uint8_t)255;
        }
    }
}

}  // namespace ldpc535
