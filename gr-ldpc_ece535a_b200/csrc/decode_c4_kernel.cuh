// Thread-per-codeword sum-product kernel specialised for the 32x64 code the reference blocks
// are hard-wired to (lib/ldpc_decoder_cb_impl.cc:39-40, :60-96).
//
// The re-ordered H (the constructor's reorderHMatrix, :255-307) is evaluated at COMPILE TIME
// from the literal in ldpc535_default_code.h, and every edge of the Tanner graph becomes a
// template constant: the iteration body is straight-line code with exact node degrees (no
// padded slots, no adjacency-table loads, no shuffles or ballots).  One thread owns one
// codeword: its 64 intrinsic values live in registers, its 168 edge messages in a private
// shared-memory column (message e of thread t at ms[e * NT + t]: every warp access is
// conflict-free and needs no synchronisation).  The arithmetic is spa_math.cuh's, in the
// same operation order as the generic kernels, so all kernel families agree to the bit.
// A handle dispatches here only after checking its run-time tables against these
// compile-time ones (tables_match_c4).
#pragma once
#include <utility>

#include "code_tables.h"
#include "decode_kernels.cuh"
#include "ldpc535_default_code.h"

namespace ldpc535 {
namespace c4 {

constexpr int kM = LDPC535_DEFAULT_M, kN = LDPC535_DEFAULT_N, kE = LDPC535_DEFAULT_E;
static_assert(kM == 32 && kN == 64, "the thread kernel packs hard decisions in two words");

struct Tables {
    int pivots[kM];
    int row_ptr[kM + 1];
    int col_idx[kE];        // re-ordered columns, ascending within a row (CSR edge order)
    int col_ptr[kN + 1];
    int edge_of_col[kE];    // CSR edge ids of a column, rows ascending
    unsigned row_lo[kM], row_hi[kM];
    unsigned col_mask[kN];        // bit j = check j contains the bit (kM <= 32)
    bool ok;
};

// reorderHMatrix ("First" pivot strategy) on 64-bit rows, then CSR/CSC of the permuted H.
constexpr Tables make_tables()
{
    Tables t = {};
    unsigned long long F[kM] = {}, H[kM] = {};
    for (int j = 0; j < kM; j++)
        for (int e = ldpc535_default_row_ptr[j]; e < ldpc535_default_row_ptr[j + 1]; e++)
            F[j] |= 1ull << ldpc535_default_col_idx[e];
    for (int j = 0; j < kM; j++) H[j] = F[j];
    t.ok = true;
    for (int i = 0; i < kM; i++) {
        int chosen = -1;
        for (int c = i; c < kN; c++)
            if ((F[i] >> c) & 1ull) { chosen = c; break; }
        if (chosen < 0) { chosen = 0; t.ok = false; }
        t.pivots[i] = chosen;
        if (chosen != i) {
            for (int r = 0; r < kM; r++) {
                if (((F[r] >> i) ^ (F[r] >> chosen)) & 1ull) F[r] ^= (1ull << i) | (1ull << chosen);
                if (((H[r] >> i) ^ (H[r] >> chosen)) & 1ull) H[r] ^= (1ull << i) | (1ull << chosen);
            }
        }
        for (int k = i + 1; k < kM; k++)
            if ((F[k] >> i) & 1ull) F[k] ^= F[i];
    }
    int e = 0;
    for (int j = 0; j < kM; j++) {
        t.row_ptr[j] = e;
        for (int c = 0; c < kN; c++)
            if ((H[j] >> c) & 1ull) t.col_idx[e++] = c;
        t.row_lo[j] = (unsigned)(H[j] & 0xffffffffull);
        t.row_hi[j] = (unsigned)(H[j] >> 32);
    }
    t.row_ptr[kM] = e;
    if (e != kE) t.ok = false;
    for (int c = 0; c < kN; c++) {
        t.col_mask[c] = 0;
        for (int j = 0; j < kM; j++)
            if ((H[j] >> c) & 1ull) t.col_mask[c] |= 1u << j;
    }
    int q = 0;
    for (int c = 0; c < kN; c++) {
        t.col_ptr[c] = q;
        for (int j = 0; j < kM; j++)
            for (int k = t.row_ptr[j]; k < t.row_ptr[j + 1]; k++)
                if (t.col_idx[k] == c) t.edge_of_col[q++] = k;
    }
    t.col_ptr[kN] = q;
    return t;
}

constexpr Tables kT = make_tables();
static_assert(kT.ok, "the shipped code must have an LU pivot order");

template <int... Is, class F>
__device__ __forceinline__ void static_for_impl(std::integer_sequence<int, Is...>, F &&f)
{
    (f(std::integral_constant<int, Is>{}), ...);
}
template <int N, class F>
__device__ __forceinline__ void static_for(F &&f)
{
    static_for_impl(std::make_integer_sequence<int, N>{}, static_cast<F &&>(f));
}

// Occupancy and storage (measured on B200, profiles/): ONE 256-thread CTA per SM, 2 warps per
// scheduler.  A thread keeps its 64 intrinsic values and kRegEdges of its 168 messages in
// registers (255 per thread) and the other kSmemEdges in its private shared-memory column.
// Shared-memory accesses and MUFU share the MIO instruction queue, which was the top stall
// with all messages in shared memory (384 threads: 1.03e12 edge-iterations/s; this split:
// 1.13e12).  Phase fences (empty asm with a memory clobber) make the compiler re-read the
// shared-memory messages at phase boundaries instead of carrying all 168 in registers and
// spilling them to local memory.
#ifndef C4_THREADS
#define C4_THREADS 256
#endif
constexpr int kThreads = C4_THREADS;
#ifndef C4_CTAS
#define C4_CTAS 1
#endif
constexpr int kCtasPerSm = C4_CTAS;
// The unrolled iteration body is ~55 KB of SASS, more than the 32 KB L1.5 instruction cache:
// twelve warps streaming it independently stall on instruction fetch (ncu: no_instruction was
// the top stall).  A CTA barrier every few nodes keeps all warps of the SM within one
// cache-resident window of the code, so each line is fetched once per SM per pass.
// Round 1's body (2 990 instructions per iteration) was fastest with a barrier every 8 checks / 16
// bits; the shorter round-2 body (2 846) with one after the 32 checks and one every 32 bits (A/B on
// one box, ms per 2 M codewords at 50 iterations: 8/16 12.97, 16/32 12.82, 11/22 12.73, 32/32 12.71,
// 32/22 12.92, 32/16 12.96, 16/64 12.93, 4/8 13.39).
#ifndef C4_CHK_GROUP
#define C4_CHK_GROUP 32
#endif
#ifndef C4_VAR_GROUP
#define C4_VAR_GROUP 32
#endif
constexpr int kChkGroup = C4_CHK_GROUP;     // checks between barriers
constexpr int kVarGroup = C4_VAR_GROUP;     // variables between barriers
#ifndef C4_REG_EDGES
#define C4_REG_EDGES 106
#endif
#ifndef C4_GROUPS
#define C4_GROUPS 1
#endif
constexpr int kGroups = C4_GROUPS;          // 1: all warps in lock step; 2, 3: rotated groups
constexpr int kRegEdges = C4_REG_EDGES;
constexpr int kSmemEdges = kE - kRegEdges;
// Intrinsic values: 64 shared-memory columns like the messages (same base register and constant
// offsets in the iteration, so the compiler can order these loads freely against the message
// stores), filled once per batch from a staging buffer of one 64-value row per thread into which
// the NEXT batch's symbols are fetched asynchronously.  The row pitch of 65 words makes both
// staging access patterns bank-conflict free: the fetch writes 32 consecutive words of one row,
// the owner reads word i of 32 consecutive rows (banks row + i).
constexpr int kRPitch = kN + 1;
constexpr size_t kSmemBytes = (size_t)kThreads * (kSmemEdges + kN + kRPitch) * sizeof(float);

__device__ __forceinline__ void cp_async_4(float *smem_dst, const float *gmem_src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;"
                 :: "r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}

template <bool DEBUG>
__global__ void __launch_bounds__(kThreads, kCtasPerSm)
decode_c4_thread_kernel(const DecodeParams p)
{
    extern __shared__ float c4_smem[];
    constexpr int NT = kThreads;
    if (p.select && *p.select != p.select_want) return;      // the other family was picked for these windows
    float *ms = c4_smem + threadIdx.x;
    float mreg[kRegEdges];
    auto msg_ld = [&](auto ec) -> float {
        constexpr int e = decltype(ec)::value;
        if constexpr (e < kSmemEdges) return ms[e * NT]; else return mreg[e - kSmemEdges];
    };
    auto msg_st = [&](auto ec, float v) {
        constexpr int e = decltype(ec)::value;
        if constexpr (e < kSmemEdges) ms[e * NT] = v; else mreg[e - kSmemEdges] = v;
    };
    const long long stride = (long long)gridDim.x * NT;
    // group of this warp: warps [g * W, (g+1) * W) with W = warps / kGroups a multiple of 4, so
    // every SM sub-partition (warp id mod 4) holds one warp of each group
    const int grp = (kGroups > 1) ? (int)(threadIdx.x / (NT / kGroups)) : 0;
    auto group_sync = [&]() {
        if constexpr (kGroups == 1) __syncthreads();
        else asm volatile("bar.sync %0, %1;" :: "r"(1 + grp), "n"(NT / kGroups) : "memory");
    };

    // Fetch of one batch's real parts into the staging rows, warp-cooperative and asynchronous: for
    // codeword j of the warp's 32 every lane copies symbols `lane` and `lane + 32` with 4-byte
    // cp.async, so one instruction reads 256 contiguous bytes of gr_complex (128 of packed reals)
    // and writes 32 consecutive words of row j.  Issued one whole batch ahead: the load phase used
    // to be 64 scalar loads per thread at a 512-byte stride between lanes with every warp of the SM
    // waiting on them -- 10.5 us of a 36 us batch at 5 iterations (profiles/r2_iter_sweep.txt).
    float *const rstage = c4_smem + (kSmemEdges + kN) * NT;
    const int lane = threadIdx.x & 31;
    const int elt = p.sym_re ? 1 : 2;               // floats per symbol
    const float *const gsrc = (p.sym_re ? p.sym_re : reinterpret_cast<const float *>(p.sym)) + lane * elt;
    auto fetch_batch = [&](long long base_b) {
        const long long wb = base_b + threadIdx.x;
        long long offb = -1;
        if (wb < p.n_win) offb = p.win_offset ? p.win_offset[wb] : wb * (long long)kN;
        const bool good = offb >= 0 && offb + kN <= p.n_sym;
        const unsigned m = __ballot_sync(0xffffffffu, good);
        float *dst = rstage + (threadIdx.x - lane) * kRPitch + lane;
#pragma unroll 4
        for (int j = 0; j < 32; j++) {
            const long long oj = __shfl_sync(0xffffffffu, offb, j);
            if ((m >> j) & 1u) {
                const float *s = gsrc + oj * elt;
                cp_async_4(dst + j * kRPitch, s);
                cp_async_4(dst + j * kRPitch + 32, s + 32 * elt);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    fetch_batch((long long)blockIdx.x * NT);

    // every thread of the CTA makes the same number of trips (barriers inside); threads past
    // the end of the batch compute on zeros and store nothing
    for (long long base = (long long)blockIdx.x * NT; base < p.n_win; base += stride) {
        const long long w = base + threadIdx.x;
        const bool live = w < p.n_win;
        const long long off = !live ? -1 : (p.win_offset ? p.win_offset[w] : w * (long long)kN);
        const float npol = ((live && p.polarity) ? -(float)p.polarity[w] : -1.f) * kSpaScale;   // units of ln 2
        const bool ok = live && off >= 0 && off + kN <= p.n_sym;

        // ---- r_i = -pol * Re(sym_i)  (lib/ldpc_decoder_cb_impl.cc:149-153, :486) ----
        // The real parts of this batch were fetched into the staging rows while the previous batch
        // iterated (fetch_batch); scale them into the r columns, then start the next batch's fetch.
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        float r[kN];
        {
            const float *row = rstage + threadIdx.x * kRPitch;
#pragma unroll
            for (int i = 0; i < kN; i++) r[i] = ok ? __fmul_rn(npol, row[i]) : 0.f;
        }
        static_for<kN>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            ms[(kSmemEdges + i) * NT] = r[i];
        });
        __syncwarp();                      // every row of the warp has been read: the rows may be overwritten
        if (base + stride < p.n_win) fetch_batch(base + stride);
        // M_ji = r_i on every edge (:489-496), stored as t = copysign(2^-|M|, M)
        static_for<kE>([&](auto ec) {
            constexpr int e = decltype(ec)::value;
            constexpr int c = kT.col_idx[e];
            if (DEBUG && p.dbgS && live) p.dbgS[w * kE + e] = r[c] * kSpaUnscale;
            msg_st(ec, to_check_msg(r[c]));
        });

        asm volatile("" ::: "memory");
        unsigned out_h1 = 0;
        int out_cnt = 0, iters = p.max_iters;
        bool broke = false;          // this codeword's test passed: its outputs are frozen

        // ---- Step 1: check messages of checks [J0, J1), in place (:503-516) ----
        auto check_range = [&](auto j0c, auto j1c) {
            constexpr int J0 = decltype(j0c)::value, J1 = decltype(j1c)::value;
            static_for<J1 - J0>([&](auto jc) {
                constexpr int j = J0 + decltype(jc)::value;
                constexpr int a = kT.row_ptr[j];
                constexpr int d = kT.row_ptr[j + 1] - a;
                float m[d];
                static_for<d>([&](auto sc) {
                    constexpr int s = decltype(sc)::value;
                    m[s] = msg_ld(std::integral_constant<int, a + s>{});
                });
                if (DEBUG && p.dbgM && live && !broke) {     // M entering this check step
#pragma unroll
                    for (int s = 0; s < d; s++) p.dbgM[w * kE + a + s] = p.dbgS[w * kE + a + s];
                }
                check_node_spa<d>(m);
                static_for<d>([&](auto sc) {
                    constexpr int s = decltype(sc)::value;
                    msg_st(std::integral_constant<int, a + s>{}, m[s]);
                });
                if (DEBUG && p.dbgE && live && !broke) {
#pragma unroll
                    for (int s = 0; s < d; s++) p.dbgE[w * kE + a + s] = m[s] * kSpaUnscale;
                }
                if constexpr (j % kChkGroup == kChkGroup - 1) group_sync();
            });
        };
        // ---- Test (:519-532) fused with Step 2 (:540-553), then Finished? (:535-537) ----
        auto var_all = [&](int h) {
            // decisions: bits 32..63 are the output; the parity word `syn` (bit j = check j unsatisfied) is the XOR of
            // the column masks of the bits decided 1 -- one predicated XOR per bit instead of a mask-and-popc per
            // check over two packed words (324 -> ~160 instructions per iteration, same weight)
            unsigned h1 = 0, syn = 0;
            static_for<kN>([&](auto ic) {
                constexpr int i = decltype(ic)::value;
                constexpr int a = kT.col_ptr[i];
                constexpr int dv = kT.col_ptr[i + 1] - a;
                float L = 0.f;
                if constexpr (dv > 0) {
                    float x[dv];
                    static_for<dv>([&](auto kc) {
                        constexpr int k = decltype(kc)::value;
                        constexpr int e = kT.edge_of_col[a + k];
                        x[k] = msg_ld(std::integral_constant<int, e>{});
                    });
                    L = var_node_spa<dv>(x, dv, ms[(kSmemEdges + i) * NT]);
                    static_for<dv>([&](auto kc) {
                        constexpr int k = decltype(kc)::value;
                        constexpr int e = kT.edge_of_col[a + k];
                        if (DEBUG && p.dbgS && live && !broke) p.dbgS[w * kE + e] = x[k] * kSpaUnscale;
                        msg_st(std::integral_constant<int, e>{}, to_check_msg(x[k]));
                    });
                }
                if (DEBUG && p.dbgL && live && !broke) p.dbgL[w * kN + i] = L * kSpaUnscale;
                constexpr unsigned cmask = kT.col_mask[i];
                if (L <= 0.f) {                                     // :527
                    syn ^= cmask;
                    if constexpr (i >= 32) h1 |= 1u << (i - 32);
                }
                if constexpr (i % kVarGroup == kVarGroup - 1) group_sync();
            });
            // the syndrome is tested every iteration, the last included
            if (p.early_stop || h + 1 == p.max_iters) {
                const int cnt = __popc(syn);
                if (!broke) { out_h1 = h1; out_cnt = cnt; }
                // A converged codeword stops here in the reference; its thread keeps pace with
                // the CTA (barriers) but its outputs no longer change.
                if (p.early_stop && !broke && cnt == 0) { iters = h + 1; broke = true; }
            }
        };
        using I0 = std::integral_constant<int, 0>;
        using IH = std::integral_constant<int, kM / 2>;
        using IM = std::integral_constant<int, kM>;

        // Rotation: the warps of an SM sub-partition belong to different groups, and group g runs
        // one sub-step behind group g-1, so that while one group is in the MUFU-free variable
        // phase the others keep the XU pipe busy with check nodes (in lock step all warps hit
        // the variable phase together and the XU pipe idles ~20 % of the time).
        const int n_sub = kGroups * p.max_iters;
        for (int t = 0; t < n_sub + kGroups - 1; t++) {
            const int u = t - grp;
            if (u >= 0 && u < n_sub) {
                const int h = u / kGroups, ph = u - h * kGroups;
                if constexpr (kGroups == 1) {
                    check_range(I0{}, IM{});
                    asm volatile("" ::: "memory");
                    var_all(h);
                } else if constexpr (kGroups == 2) {
                    if (ph == 0) check_range(I0{}, IM{}); else var_all(h);
                } else {
                    if (ph == 0) check_range(I0{}, IH{});
                    else if (ph == 1) check_range(IH{}, IM{});
                    else var_all(h);
                }
            }
            asm volatile("" ::: "memory");
            if (p.early_stop) {
                if (__syncthreads_and(broke || !live)) break;
            } else if (kGroups > 1) {
                __syncthreads();
            }
        }
        if (DEBUG && p.dbgM && live && !broke) {
            static_for<kE>([&](auto ec) {
                constexpr int e = decltype(ec)::value;
                p.dbgM[w * kE + e] = p.dbgS[w * kE + e];
            });
        }

        // ---- outputs: bits 32..63 MSB first (:207-219), checkFrame weight (:236-253) ----
        if (ok) {
            const unsigned bytes = __byte_perm(__brev(out_h1), 0, 0x0123);
            *reinterpret_cast<unsigned *>(p.out_bytes + w * 4) = bytes;
            if (p.out_synd) p.out_synd[w] = (uint8_t)min(out_cnt, p.thr + 1);
            if (p.out_iters) p.out_iters[w] = (uint8_t)min(iters, 255);
        } else if (live) {
            *reinterpret_cast<unsigned *>(p.out_bytes + w * 4) = 0u;
            if (p.out_synd) p.out_synd[w] = 255;
            if (p.out_iters) p.out_iters[w] = 255;
        }
    }
}

}  // namespace c4

inline cudaError_t launch_c4_thread(const DecodeParams &p, bool dbg, int sm_count, cudaStream_t st)
{
    auto kern = dbg ? c4::decode_c4_thread_kernel<true> : c4::decode_c4_thread_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)c4::kSmemBytes);
    if (e != cudaSuccess) return e;
    const long long ctas = (p.n_win + c4::kThreads - 1) / c4::kThreads;
    long long grid = (long long)sm_count * c4::kCtasPerSm;
    if (ctas < grid) grid = ctas;
    kern<<<(int)grid, c4::kThreads, c4::kSmemBytes, st>>>(p);
    return cudaGetLastError();
}

}  // namespace ldpc535

namespace {
// The handle's run-time tables equal the compile-time ones of the specialised kernel.
inline bool tables_match_c4(const ldpc535::CodeTables &t)
{
    using namespace ldpc535::c4;
    if (t.M != kM || t.N != kN || t.E != kE) return false;
    for (int j = 0; j <= kM; j++) if (t.row_ptr[j] != kT.row_ptr[j]) return false;
    for (int e = 0; e < kE; e++) if (t.col_idx[e] != kT.col_idx[e]) return false;
    for (int c = 0; c <= kN; c++) if (t.col_ptr[c] != kT.col_ptr[c]) return false;
    for (int e = 0; e < kE; e++) if (t.edge_of_col[e] != kT.edge_of_col[e]) return false;
    for (int j = 0; j < kM; j++) if (t.pivots[j] != kT.pivots[j]) return false;
    return true;
}
}  // namespace
