// C ABI of the B200 LDPC hot path (include/ldpc535.h): handle management, table upload,
// kernel dispatch, and the pinned-buffer H2D -> kernel -> D2H pipeline behind the
// host-buffer entry points.  No CPU fallback anywhere: without a usable sm_100 device
// every compute entry point returns an error.
#include "../../include/ldpc535.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "code_tables.h"
#include "decode_kernels.cuh"
#include "decode_c4_kernel.cuh"
#include "decode_c4_refill_kernel.cuh"
#include "decode_regular_kernel.cuh"
#include "decode_hard_kernel.cuh"
#include "encode_kernels.cuh"
#include "encode_m4r_kernel.cuh"
#include "encode_m4r_tiles.h"
#include "host_pack.h"
#include "synth_kernels.cuh"
#include "ldpc535_default_code.h"

using namespace ldpc535;

namespace {

thread_local std::string g_last_error;

int fail(int status, const std::string &msg)
{
    g_last_error = msg;
    return status;
}

#define CU(call)                                                                          \
    do {                                                                                  \
        cudaError_t e__ = (call);                                                         \
        if (e__ != cudaSuccess)                                                           \
            return fail(LDPC535_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)

constexpr long long kAdaptiveSample = 16384;     // windows decoded first to pick the family ("c4-adaptive")
enum KernelFamily { kAuto = 0, kWarp = 1, kBlock = 2, kC4Thread = 3, kRegular = 4, kHard64 = 5, kC4Refill = 6, kC4Adaptive = 7 };

constexpr int kSlots = 3;                       // pipeline depth of the host-buffer API
constexpr size_t kChunkSymBytes = 128u << 20;   // symbol bytes staged per slot (upper bound; slots grow on demand)
constexpr int kPackAdaptive = 2;                // pack_pinned: 0 never, 1 always, 2 chunk by chunk (default)
constexpr int kH2dRing = 8;                     // timing events of the last chunks' host-to-device copies

struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    // staging grows on demand (a block's first general_work() with a few frames must not stall on
    // 3 x 128 MB of cudaMalloc): capacities in bytes / windows, 0 = not allocated yet
    size_t cap_sym = 0, cap_win = 0, cap_stage = 0;
    void *d_sym = nullptr;        // <= kChunkSymBytes
    long long *d_off = nullptr;   // cap_win windows
    signed char *d_pol = nullptr;
    uint8_t *d_bytes = nullptr, *d_synd = nullptr, *d_iters = nullptr;
    long long *h_off = nullptr;   // pinned, rebased offsets
    float *h_stage = nullptr;     // pinned, <= kChunkSymBytes / 2: real parts packed by the host
};

}  // namespace

struct ldpc535_code {
    int device = 0;
    int sm_count = 0;
    size_t smem_optin = 0;
    CodeTables t;
    // device tables
    uint16_t *d_chk_var = nullptr, *d_var_slot = nullptr;
    uint8_t *d_chk_deg = nullptr;
    int32_t *d_slot_edge = nullptr;
    uint16_t *d_w_chk_pos = nullptr, *d_w_var_pos = nullptr;   // warp kernel strip layout
    int32_t *d_w_pos_edge = nullptr;
    uint16_t *d_var_row4 = nullptr;   // [N][4] message addresses of a bit (regular codes, dv <= 4)
    uint16_t *d_var_row4_rt = nullptr; // the same under the coloured-row layout of the register-table kernel
    uint32_t *d_chk_color = nullptr;  // [M] colours of a storage column's six slots inside their rows: slot s at bits 5s+2 .. 5s+6
    unsigned int *d_cursor = nullptr; // window cursors of the register-table kernel (one per launch in flight, round-robin)
    unsigned int cursor_launch = 0;
    int c4_refill_min = 1;            // "c4-refill": lanes of a warp that must be waiting before its slots are refilled
    int *d_select = nullptr;          // "c4-adaptive": family words (one per launch in flight, round-robin) + sample iteration counts
    unsigned int select_launch = 0;
    bool fits_regular = false;
    int regular_variant = 1;          // 1: 512-thread register-table kernel (fixed sizes), 0: 1024-thread kernel
    uint32_t *d_Pt = nullptr, *d_Pw = nullptr;
    uint32_t *d_m4r = nullptr;        // table of the look-up encoder (large codes), see encode_m4r_kernel.cuh
    bool use_m4r = true;              // LDPC535_ENCODER=generic switches it off (A/B measurements)
    int m4r_ring = 60;                // look-up encoder variant (LDPC535_M4R_RING): 40 | 60 = free-running kernel with a 4- / 6-slot
                                      // ring (60: default); 42 | 62 = barrier kernel, ring slots x stages per barrier
    std::vector<std::pair<cudaStream_t, uint4 *>> m4r_parks;   // free-running kernel: parity scratch, one per stream in use
    int m4r_tpf = 7;                  // free-running kernel: frames per slot, 7 | 8 (LDPC535_M4R_TPF; 8 spills)
    int m4r_tile = 0;                 // frames per tile, 0 = balanced over the SMs (LDPC535_M4R_TILE)
    uint32_t m4r_flags = 0;           // timing experiments (LDPC535_M4R_FLAGS, kM4rNo*): wrong results
    int tabA_bytes = 0, tabB_bytes = 0;
    int dc_t = 0, dv_t = 0;           // template sizes used (6/3 or 16/8), 0 = unsupported degrees
    bool fits_warp = false, fits_block = false, is_c4 = false;
    int stage_tables = 0;
    size_t block_smem = 0;
    int block_threads = 0;
    int forced = kAuto;
    cudaStream_t stream = nullptr;
    // host-API pipeline
    int pack_threads = 1;
    int pack_pinned = kPackAdaptive;  // what happens to PINNED input (pageable input is always packed through staging):
                                  // 0 the copy engine reads it as it is (8 bytes per symbol over PCIe), 1 the host team packs
                                  // the real parts first (4 bytes per symbol, but 12 bytes of host memory traffic), 2 decided
                                  // from the measured rates of the team and of the copy engine.  Measured per 5.12 GB on B200
                                  // boxes: one rank, 16 threads: packed 70 ms, raw 94 ms; two ranks, 12 threads each: packed
                                  // 109 ms, raw 94 ms; four ranks, 8 threads each: packed 220 ms, raw 101 ms.
    double pack_ns_per_byte = 0, h2d_ns_per_byte = 0;      // running estimates (input bytes packed / bytes copied), 0 = not measured yet
    cudaEvent_t h2d_t0[kH2dRing] = {}, h2d_t1[kH2dRing] = {};
    size_t h2d_len[kH2dRing] = {};                          // bytes of the copy of chunk k % kH2dRing, 0 = none
    bool h2d_alone[kH2dRing] = {};                          // no earlier copy was pending when it was issued
    uint64_t chunks_packed = 0, chunks_raw = 0, h2d_bytes = 0, host_calls = 0;
    bool packing = false;                                   // current regime of the measured mode
    PackPool *pool = nullptr;     // persistent packing team, created on first use
    size_t max_win_per_chunk = 0;
    Slot slots[kSlots];
    uint64_t launches = 0;
};

namespace {

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev)
    {
        if (dev < 0) { ok = false; return; }
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

#define NEED_DEVICE(c, g)                                                                  \
    do {                                                                                   \
        if ((c)->device < 0)                                                               \
            return fail(LDPC535_ERR_NO_DEVICE, "handle was created with LDPC535_DEVICE_NONE (tables only; there is no CPU compute path)"); \
        if (!(g).ok) return fail(LDPC535_ERR_CUDA, "cudaSetDevice failed");                \
    } while (0)

template <typename T>
int upload(T **dptr, const void *src, size_t bytes, size_t alloc_bytes)
{
    CU(cudaMalloc(reinterpret_cast<void **>(dptr), alloc_bytes));
    CU(cudaMemset(*dptr, 0xFF, alloc_bytes));
    CU(cudaMemcpy(*dptr, src, bytes, cudaMemcpyHostToDevice));
    return LDPC535_OK;
}

bool tables_match_c4(const CodeTables &t);

int finish_create(ldpc535_code *c)
{
    const CodeTables &t = c->t;
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, c->device));
    if (prop.major < 10)
        return fail(LDPC535_ERR_NO_DEVICE, std::string("device is sm_") + std::to_string(prop.major) +
                                               std::to_string(prop.minor) + ", kernels are built for sm_100a");
    c->sm_count = prop.multiProcessorCount;
    c->pack_threads = default_pack_threads();
    if (const char *e = getenv("LDPC535_PACK_PINNED")) c->pack_pinned = std::max(0, std::min(2, atoi(e)));
    c->smem_optin = prop.sharedMemPerBlockOptin;
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));

    if (t.dc_max <= 6 && t.dv_max <= 3) { c->dc_t = 6; c->dv_t = 3; }
    else if (t.dc_max <= 16 && t.dv_max <= 8) { c->dc_t = 16; c->dv_t = 8; }
    else return fail(LDPC535_ERR_UNSUPPORTED, "check degree > 16 or bit degree > 8");
    if (t.chk_var.empty())
        return fail(LDPC535_ERR_UNSUPPORTED, "dc_max * M exceeds 65535 message slots");

    // decoder tables, padded to the template slot counts (extra slots = 0xFFFF)
    const size_t a_bytes = (size_t)c->dc_t * t.M * sizeof(uint16_t);
    const size_t b_bytes = (size_t)c->dv_t * t.N * sizeof(uint16_t);
    c->tabA_bytes = (int)((a_bytes + 15) & ~(size_t)15);
    c->tabB_bytes = (int)((b_bytes + 15) & ~(size_t)15);
    int st;
    if ((st = upload(&c->d_chk_var, t.chk_var.data(), t.chk_var.size() * 2, c->tabA_bytes))) return st;
    if ((st = upload(&c->d_var_slot, t.var_slot.data(), t.var_slot.size() * 2, c->tabB_bytes))) return st;
    if ((st = upload(&c->d_chk_deg, t.chk_deg_slot.data(), t.chk_deg_slot.size(), t.chk_deg_slot.size()))) return st;
    {
        std::vector<int32_t> slot_edge((size_t)c->dc_t * t.M, -1);
        for (int e = 0; e < t.E; e++) slot_edge[t.edge_slot[e]] = e;
        if ((st = upload(&c->d_slot_edge, slot_edge.data(), slot_edge.size() * 4, slot_edge.size() * 4))) return st;
    }
    // encoder tables
    if ((st = upload(&c->d_Pt, t.Pt.data(), t.Pt.size() * 4, std::max<size_t>(t.Pt.size() * 4, 128)))) return st;
    {
        std::vector<uint32_t> Pw((size_t)t.kwords * t.M);
        for (int j = 0; j < t.M; j++)
            for (int w = 0; w < t.kwords; w++) Pw[(size_t)w * t.M + j] = t.P[(size_t)j * t.kwords + w];
        if ((st = upload(&c->d_Pw, Pw.data(), Pw.size() * 4, Pw.size() * 4))) return st;
    }

    // large codes: the look-up encoder's table, built on the device from the column masks
    if (const char *e = getenv("LDPC535_ENCODER")) c->use_m4r = strcmp(e, "generic") != 0;
    if (const char *e = getenv("LDPC535_M4R_RING")) { const int r = atoi(e); if (r == 42 || r == 62 || r == 40 || r == 60) c->m4r_ring = r; }
    if (const char *e = getenv("LDPC535_M4R_TPF")) c->m4r_tpf = atoi(e) == 8 ? 8 : 7;
    if (const char *e = getenv("LDPC535_M4R_TILE")) c->m4r_tile = std::max(0, std::min(1024, atoi(e)));
#ifdef LDPC535_TIMING_EXPERIMENTS      // builds under tools/ab/ only: parts of the kernel switched off, wrong results
    if (const char *e = getenv("LDPC535_M4R_FLAGS")) c->m4r_flags = (uint32_t)atoi(e);
#endif
    if (t.M % kM4rRows == 0 && t.K % 128 == 0 && t.M >= 4 * kM4rRows && (t.K / 32) % (t.M / kM4rRows) == 0 &&
        encode_m4r_table_bytes(t.M, t.K) <= ((size_t)512 << 20) &&
        encode_m4r_smem_bytes<8, 6>() <= c->smem_optin) {
        CU(cudaMalloc(reinterpret_cast<void **>(&c->d_m4r), encode_m4r_table_bytes(t.M, t.K)));
        encode_m4r_build_kernel<<<(t.M / kM4rRows) * (t.K / 8), 32>>>(c->d_Pt, c->d_m4r, t.M, t.K, t.mwords);
        CU(cudaGetLastError());
        CU(cudaDeviceSynchronize());
    }

    // regular code -> the specialised CTA-per-codeword kernel
    {
        bool regular = t.dv_max <= 4;
        for (int j = 0; j < t.M && regular; j++) regular = t.chk_deg[j] == t.dc_max;
        for (int v = 0; v < t.N && regular; v++) regular = t.var_deg[v] == t.dv_max;
        regular = regular && t.dc_max == 6 && t.dv_max == 3 && t.M > 32;
        if (regular && regular_smem_bytes(6, t.M, t.N) <= c->smem_optin) {
            std::vector<uint16_t> row4((size_t)t.N * 4, 0);
            for (int v = 0; v < t.N; v++)
                for (int k = 0; k < t.dv_max; k++) row4[(size_t)v * 4 + k] = t.var_slot[(size_t)k * t.N + v];
            if ((st = upload(&c->d_var_row4, row4.data(), row4.size() * 2, row4.size() * 2))) return st;
            if (!t.row_color.empty()) {
                for (int v = 0; v < t.N; v++)
                    for (int k = 0; k < t.dv_max; k++) row4[(size_t)v * 4 + k] = t.var_slot_colored[(size_t)k * t.N + v];
                if ((st = upload(&c->d_var_row4_rt, row4.data(), row4.size() * 2, row4.size() * 2))) return st;
                std::vector<uint32_t> cc(t.M, 0);
                for (int j = 0; j < t.M; j++)
                    for (int sl = 0; sl < 6; sl++) cc[j] |= (uint32_t)t.row_color[(size_t)j * 6 + sl] << (5 * sl + 2);
                if ((st = upload(&c->d_chk_color, cc.data(), cc.size() * 4, cc.size() * 4))) return st;
            }
            c->fits_regular = true;
        }
    }
    c->fits_warp = (t.M <= 32 && t.N <= 64);
    if (c->fits_warp) {
        // padded to the template slot counts (rows beyond dc_max / dv_max are never used)
        std::vector<uint16_t> cp((size_t)c->dc_t * 32, 0), vp((size_t)c->dv_t * 64, 0);
        for (size_t i = 0; i < vp.size(); i++) vp[i] = (uint16_t)(0x8000u | (i & 31u));    // slots beyond dv_max: unused, own bank
        std::vector<int32_t> pe((size_t)c->dc_t * 32, -1);
        std::copy(t.w_chk_pos.begin(), t.w_chk_pos.end(), cp.begin());
        std::copy(t.w_var_pos.begin(), t.w_var_pos.end(), vp.begin());
        std::copy(t.w_pos_edge.begin(), t.w_pos_edge.end(), pe.begin());
        for (int s = t.dc_max; s < c->dc_t; s++)
            for (int l = 0; l < 32; l++) cp[(size_t)s * 32 + l] = (uint16_t)(s * 32 + l);
        if ((st = upload(&c->d_w_chk_pos, cp.data(), cp.size() * 2, cp.size() * 2))) return st;
        if ((st = upload(&c->d_w_var_pos, vp.data(), vp.size() * 2, vp.size() * 2))) return st;
        if ((st = upload(&c->d_w_pos_edge, pe.data(), pe.size() * 4, pe.size() * 4))) return st;
    }
    c->is_c4 = tables_match_c4(t);
    const size_t fixed = block_smem_fixed_bytes(c->dc_t, t.M, t.N);
    const size_t staged = fixed + c->tabA_bytes + c->tabB_bytes;
    if (staged <= c->smem_optin) { c->stage_tables = 1; c->block_smem = staged; c->fits_block = true; }
    else if (fixed <= c->smem_optin) { c->stage_tables = 0; c->block_smem = fixed; c->fits_block = true; }
    int nt = std::max(t.M, (t.N + 1) / 2);
    nt = std::min(1024, ((nt + 31) / 32) * 32);
    c->block_threads = std::max(nt, 64);
    if (const char *e = getenv("LDPC535_REGULAR_VARIANT")) c->regular_variant = atoi(e) ? 1 : 0;
    if (const char *e = getenv("LDPC535_C4_REFILL_MIN")) c->c4_refill_min = std::max(1, std::min(32, atoi(e)));
    if (const char *e = getenv("LDPC535_BLOCK_THREADS")) c->block_threads = std::max(64, std::min(1024, atoi(e) / 32 * 32));
    if (!c->fits_warp && !c->fits_block)
        return fail(LDPC535_ERR_UNSUPPORTED, "code does not fit the shared-memory resident decoder");
    return LDPC535_OK;
}

void release(ldpc535_code *c)
{
    if (!c) return;
    if (c->device < 0) { delete c->pool; delete c; return; }
    DeviceGuard g(c->device);
    delete c->pool;
    c->pool = nullptr;
    for (auto &s : c->slots) {
        if (s.stream) cudaStreamSynchronize(s.stream);
        cudaFree(s.d_sym); cudaFree(s.d_off); cudaFree(s.d_pol);
        cudaFree(s.d_bytes); cudaFree(s.d_synd); cudaFree(s.d_iters);
        if (s.h_off) cudaFreeHost(s.h_off);
        if (s.h_stage) cudaFreeHost(s.h_stage);
        if (s.done) cudaEventDestroy(s.done);
        if (s.stream) cudaStreamDestroy(s.stream);
    }
    for (int i = 0; i < kH2dRing; i++) {
        if (c->h2d_t0[i]) cudaEventDestroy(c->h2d_t0[i]);
        if (c->h2d_t1[i]) cudaEventDestroy(c->h2d_t1[i]);
    }
    if (c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    cudaFree(c->d_chk_var); cudaFree(c->d_var_slot); cudaFree(c->d_chk_deg);
    cudaFree(c->d_slot_edge); cudaFree(c->d_Pt); cudaFree(c->d_Pw); cudaFree(c->d_m4r); for (auto &pk : c->m4r_parks) cudaFree(pk.second); cudaFree(c->d_var_row4); cudaFree(c->d_var_row4_rt); cudaFree(c->d_chk_color); cudaFree(c->d_cursor); cudaFree(c->d_select); cudaFree(c->d_w_chk_pos); cudaFree(c->d_w_var_pos); cudaFree(c->d_w_pos_edge);
    delete c;
}

int family_from_name(const char *name, int *out)
{
    if (!name || !*name || !strcmp(name, "auto")) { *out = kAuto; return 0; }
    if (!strcmp(name, "warp")) { *out = kWarp; return 0; }
    if (!strcmp(name, "block")) { *out = kBlock; return 0; }
    if (!strcmp(name, "c4-thread")) { *out = kC4Thread; return 0; }
    if (!strcmp(name, "c4-refill")) { *out = kC4Refill; return 0; }
    if (!strcmp(name, "c4-adaptive")) { *out = kC4Adaptive; return 0; }
    if (!strcmp(name, "regular")) { *out = kRegular; return 0; }
    return 1;
}

// early_stop: the thread-per-codeword kernel stops a whole 256-codeword CTA together, so with
// early stop its time does not drop when frames converge early; the warp-per-codeword kernel
// stops each codeword on its own (measured on B200, 5 iterations max: equal at 2 dB, warp kernel
// 1.1x / 1.5x / 2x faster at 4 / 6 / 8 dB; profiles/r1_microbench.txt).
int resolve_family(const ldpc535_code *c, int forced, int method, int early_stop = 0, long long n_win = 0)
{
    int f = forced;
    if (f == kAuto) {
        // bit-flip flips a bit only when MORE than M/2 of its checks disagree (reference :464): with
        // every bit degree <= M/2 it can never flip and equals the hard decision
        const bool flipless = method == LDPC535_METHOD_BITFLIP && c->t.dv_max <= c->t.M / 2;
        // decode_hard64_kernel keeps one 32-bit word of data bits per codeword: K <= 32, i.e. M = 32
        if ((method == LDPC535_METHOD_HARD || flipless) && c->fits_warp && c->t.N == 64 &&
            c->t.M == 32) f = kHard64;
        // Thread per codeword at fixed iterations.  With early stop: warp per codeword.  The persistent-slot
        // thread kernel (decode_c4_refill_kernel, "c4-refill") is opt-in: measured against the warp kernel
        // at 5 iterations max it is 1.02x / 0.89x / 1.00x / 1.23x as fast at Eb/N0 = 2 / 4 / 6 / 8 dB and
        // 0.77x at 50 iterations max, 2 dB (profiles/r2_refill_sweep.txt) -- a win only on clean channels.
        else if (c->is_c4 && method == LDPC535_METHOD_SUMPRODUCT && !early_stop) f = kC4Thread;
        // Early stop on a batch that fills the GPU several times over: a sample is decoded first and
        // the device picks lock-step thread-per-codeword (frames that mostly run to max_iters: +17 % at
        // 2 dB, 5 iterations) or warp-per-codeword (frames that stop early) for the rest.
        else if (c->is_c4 && method == LDPC535_METHOD_SUMPRODUCT && early_stop) f = kC4Adaptive;
        else if (c->fits_regular && method == LDPC535_METHOD_SUMPRODUCT) f = kRegular;
        else if (c->fits_warp) f = kWarp;
        else f = kBlock;
    }
    if (f == kC4Thread && !(c->is_c4 && method == LDPC535_METHOD_SUMPRODUCT)) return -1;
    if (f == kC4Adaptive) {
        if (!(c->is_c4 && method == LDPC535_METHOD_SUMPRODUCT)) return -1;
        if (!early_stop) f = kC4Thread;
        else if (n_win < (long long)c->sm_count * c4::kThreads * 4 + kAdaptiveSample) f = kWarp;
    }
    if (f == kC4Refill) {
        if (!(c->is_c4 && method == LDPC535_METHOD_SUMPRODUCT)) return -1;
        // early stop and batches that fill the GPU twice over (and fit its 32-bit offsets); else its siblings
        if (!early_stop) f = kC4Thread;
        else if (n_win < (long long)c->sm_count * c4::kRfThreads * 2 || n_win >= (1ll << 26)) f = kWarp;
    }
    if (f == kRegular && !(c->fits_regular && method == LDPC535_METHOD_SUMPRODUCT)) return -1;
    if (f == kWarp && !c->fits_warp) return -1;
    if (f == kBlock && !c->fits_block) return -1;
    return f;
}

int norm_method(int method) { return (method >= 1 && method <= 3) ? method : 0; }

// ---- launches --------------------------------------------------------------------

template <int METHOD, int DC, int DV, bool DBG>
cudaError_t launch_warp(const ldpc535_code *c, const DecodeParams &p, cudaStream_t st)
{
    const int wpb = kWarpKernelThreads / 32;
    const size_t smem = sizeof(msg_t<METHOD>) * wpb * (DC + 3) * 32 + sizeof(float) * wpb * 64 + 512 * wpb;   // + symbol staging + results
    long long blocks = (p.n_win + wpb - 1) / wpb;
    auto kern = decode_warp_kernel<METHOD, DC, DV, DBG>;
    int per_sm = 8;                        // grid = the resident CTAs: one wave of persistent warps
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kWarpKernelThreads, smem);
    // more CTAs than are resident: the block scheduler then evens out the sub-partitions' pace
    // (measured: 1x resident 8.97e11, 4x 9.82e11, 16x 1.005e12 edge-iterations/s)
    static const int mult = getenv("LDPC535_WARP_GRID_MULT") ? std::max(1, atoi(getenv("LDPC535_WARP_GRID_MULT"))) : 16;
    const int grid = (int)std::min<long long>(blocks, (long long)c->sm_count * std::max(per_sm, 1) * mult);
    kern<<<grid, kWarpKernelThreads, smem, st>>>(p);
    return cudaGetLastError();
}

template <int METHOD, int DC, int DV, bool DBG>
cudaError_t launch_block(const ldpc535_code *c, DecodeParams p, cudaStream_t st)
{
    auto kern = decode_block_kernel<METHOD, DC, DV, DBG>;
    // shared memory of this method's message type; adjacency tables are staged when they fit
    const size_t fixed = block_smem_fixed_bytes(DC, c->t.M, c->t.N, (int)sizeof(msg_t<METHOD>));
    const size_t staged = fixed + c->tabA_bytes + c->tabB_bytes;
    size_t smem;
    if (staged <= c->smem_optin) { p.stage_tables = 1; smem = staged; }
    else if (fixed <= c->smem_optin) { p.stage_tables = 0; smem = fixed; }
    else return cudaErrorInvalidConfiguration;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, c->block_threads, smem);
    if (e != cudaSuccess) return e;
    per_sm = std::max(per_sm, 1);
    const int grid = (int)std::min<long long>(p.n_win, (long long)c->sm_count * per_sm);
    kern<<<grid, c->block_threads, smem, st>>>(p);
    return cudaGetLastError();
}

template <int DC, int DV>
cudaError_t launch_generic(const ldpc535_code *c, int family, int method, bool dbg,
                           const DecodeParams &p, cudaStream_t st)
{
    if (family == kWarp) {
        if (dbg) return launch_warp<kMethodSpa, DC, DV, true>(c, p, st);
        switch (method) {
        case 1: return launch_warp<kMethodSpa, DC, DV, false>(c, p, st);
        case 2: return launch_warp<kMethodBitFlip, DC, DV, false>(c, p, st);
        case 3: return launch_warp<kMethodHard, DC, DV, false>(c, p, st);
        default: return launch_warp<kMethodMinSum, DC, DV, false>(c, p, st);
        }
    }
    if (dbg) return launch_block<kMethodSpa, DC, DV, true>(c, p, st);
    switch (method) {
    case 1: return launch_block<kMethodSpa, DC, DV, false>(c, p, st);
    case 2: return launch_block<kMethodBitFlip, DC, DV, false>(c, p, st);
    case 3: return launch_block<kMethodHard, DC, DV, false>(c, p, st);
    default: return launch_block<kMethodMinSum, DC, DV, false>(c, p, st);
    }
}

int launch_decode(ldpc535_code *c, int forced, int method, bool dbg, DecodeParams p, cudaStream_t st)
{
    method = norm_method(method);
    int family = resolve_family(c, forced, method, p.early_stop, p.n_win);
    if (family < 0) return fail(LDPC535_ERR_UNSUPPORTED, "kernel family not available for this code/method");
    if (p.n_win == 0) return LDPC535_OK;
    p.M = c->t.M; p.N = c->t.N; p.K = c->t.K; p.E = c->t.E;
    p.nbytes = (c->t.K + 7) / 8;
    p.chk_var = c->d_chk_var; p.var_slot = c->d_var_slot; p.chk_deg = c->d_chk_deg;
    p.slot_edge = c->d_slot_edge;
    p.w_chk_pos = c->d_w_chk_pos; p.w_var_pos = c->d_w_var_pos; p.w_pos_edge = c->d_w_pos_edge;
    p.stage_tables = c->stage_tables; p.tabA_bytes = c->tabA_bytes; p.tabB_bytes = c->tabB_bytes;
    cudaError_t e;
    if (family == kRegular && dbg) family = kBlock;            // message dumps live in the generic kernel
    if (family == kHard64) {
        const long long warps = (p.n_win + 31) / 32;
        auto kern = (c->dc_t == 6) ? decode_hard64_kernel<6> : decode_hard64_kernel<16>;
        int per_sm = 8;                    // grid = exactly the resident CTAs: one wave, no tail
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kHardThreads, 0);
        const int grid = (int)std::min<long long>((warps + kHardThreads / 32 - 1) / (kHardThreads / 32),
                                                  (long long)c->sm_count * std::max(per_sm, 1));
        kern<<<grid, kHardThreads, 0, st>>>(p);
        e = cudaGetLastError();
    } else
    if (family == kRegular) {
        // BASELINE config 4's size gets the fully unrolled instantiation
        // BASELINE config 4's size gets the compile-time-size instantiations: by default the
        // register-table kernel (512 threads); LDPC535_REGULAR_VARIANT=0 selects the 1024-thread
        // kernel with the tables in shared memory (kept for A/B measurements and other sizes)
        const bool fixed8k = c->t.M == 4096 && c->t.N == 8192 && c->block_threads == 1024;
        const int grid = (int)std::min<long long>(p.n_win, (long long)c->sm_count);
        if (fixed8k && c->regular_variant == 1 && c->d_chk_color) {
            if (!c->d_cursor && cudaMalloc(reinterpret_cast<void **>(&c->d_cursor), 64 * sizeof(unsigned int)) != cudaSuccess)
                return fail(LDPC535_ERR_CUDA, "cursor allocation");
            p.cursor = c->d_cursor + (c->cursor_launch++ & 63u);          // zeroed on the stream just before its launch
            if ((e = cudaMemsetAsync(p.cursor, 0, sizeof(unsigned int), st)) != cudaSuccess)
                return fail(LDPC535_ERR_CUDA, cudaGetErrorString(e));
            auto kern = decode_regular_rt_kernel<6, 3, 4096, 8192>;
            const size_t smem = regular_rt_smem_bytes<6, 4096, 8192>();
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) {
                kern<<<grid, 512, smem, st>>>(p, c->d_var_row4_rt, c->d_chk_color);
                e = cudaGetLastError();
            }
        } else {
            auto kern = fixed8k ? decode_regular_kernel<6, 3, 4096, 8192> : decode_regular_kernel<6, 3>;
            const size_t smem = regular_smem_bytes(6, c->t.M, c->t.N);
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) {
                kern<<<grid, c->block_threads, smem, st>>>(p, c->d_var_row4);
                e = cudaGetLastError();
            }
        }
    } else if (family == kC4Adaptive && !dbg) {
        // 1. the first kAdaptiveSample windows on the warp kernel; 2. pick_family_kernel turns their
        // iteration counts into the family word; 3. + 4. both families are queued for the remaining
        // windows and the one not picked returns at once.  Everything stays on the stream.
        if (!c->d_select && cudaMalloc(reinterpret_cast<void **>(&c->d_select), 64 * (sizeof(int) + kAdaptiveSample)) != cudaSuccess)
            return fail(LDPC535_ERR_CUDA, "family word allocation");
        const unsigned slot = c->select_launch++ & 63u;
        int *select = c->d_select + slot;
        uint8_t *sample_iters = reinterpret_cast<uint8_t *>(c->d_select + 64) + (size_t)slot * kAdaptiveSample;
        DecodeParams head = p;
        head.n_win = kAdaptiveSample;
        if (!head.out_iters) head.out_iters = sample_iters;
        e = launch_generic<6, 3>(c, kWarp, method, false, head, st);
        if (e == cudaSuccess) {
            pick_family_kernel<<<1, 256, 0, st>>>(head.out_iters, (int)kAdaptiveSample, p.max_iters, select);
            e = cudaGetLastError();
        }
        DecodeParams rest = p;
        const long long S = kAdaptiveSample;
        rest.n_win = p.n_win - S;
        if (p.win_offset) rest.win_offset = p.win_offset + S;
        else {                                            // aligned frames: shift the symbol base instead
            if (p.sym_re) rest.sym_re = p.sym_re + S * p.N; else rest.sym = p.sym + S * p.N;
            rest.n_sym = p.n_sym - S * p.N;
        }
        if (p.polarity) rest.polarity = p.polarity + S;
        rest.out_bytes = p.out_bytes + S * p.nbytes;
        if (p.out_synd) rest.out_synd = p.out_synd + S;
        if (p.out_iters) rest.out_iters = p.out_iters + S;
        rest.select = select;
        if (e == cudaSuccess) { rest.select_want = 1; e = launch_c4_thread(rest, false, c->sm_count, st); }
        if (e == cudaSuccess) { rest.select_want = 0; e = launch_generic<6, 3>(c, kWarp, method, false, rest, st); }
        if (e == cudaSuccess) c->launches += 3;
    } else if (family == kC4Refill && !dbg && p.n_sym < (1ll << 32)) {
        if (!c->d_cursor && cudaMalloc(reinterpret_cast<void **>(&c->d_cursor), 64 * sizeof(unsigned int)) != cudaSuccess)
            return fail(LDPC535_ERR_CUDA, "cursor allocation");
        unsigned int *cursor = c->d_cursor + (c->cursor_launch++ & 63u);      // zeroed on the stream just before its launch
        if ((e = cudaMemsetAsync(cursor, 0, sizeof(unsigned int), st)) != cudaSuccess)
            return fail(LDPC535_ERR_CUDA, cudaGetErrorString(e));
        e = launch_c4_refill(p, cursor, c->c4_refill_min, c->sm_count, st);
    } else if (family == kC4Thread || family == kC4Refill || family == kC4Adaptive) e = launch_c4_thread(p, dbg, c->sm_count, st);
    else if (c->dc_t == 6) e = launch_generic<6, 3>(c, family, method, dbg, p, st);
    else e = launch_generic<16, 8>(c, family, method, dbg, p, st);
    if (e != cudaSuccess) return fail(LDPC535_ERR_CUDA, std::string("decode launch: ") + cudaGetErrorString(e));
    c->launches++;
    return LDPC535_OK;
}

int launch_encode(ldpc535_code *c, const uint8_t *d_in, size_t n_frames, float *d_out, cudaStream_t st)
{
    if (n_frames == 0) return LDPC535_OK;
    const CodeTables &t = c->t;
    EncodeParams p;
    p.in = d_in; p.n_frames = (long long)n_frames; p.out = reinterpret_cast<float2 *>(d_out);
    p.M = t.M; p.N = t.N; p.K = t.K; p.nbytes = (t.K + 7) / 8; p.kwords = t.kwords; p.mwords = t.mwords;
    p.Pt = c->d_Pt; p.Pw = c->d_Pw; p.row_splits = 1;
    if (t.M <= 32 && t.K <= 32) {
        const long long warps = ((long long)n_frames + 31) / 32;
        const int grid = (int)std::min<long long>((warps + 7) / 8, (long long)c->sm_count * 8);
        encode_small_kernel<<<grid, 256, 0, st>>>(p);
    // look-up encoder from 2 048 frames on (measured: scan 0.081 / 0.151 / 0.284 ms at 1 000 / 2 000 / 3 000 frames,
    // look-up 0.155 ms at 3 072 and 0.223 ms at 8 000: the curves cross at ~1 800 frames)
    } else if (c->d_m4r && c->use_m4r && n_frames >= 2048 && (reinterpret_cast<uintptr_t>(d_in) & 15) == 0 &&
               (reinterpret_cast<uintptr_t>(d_out) & 15) == 0) {
        // look-up encoder; batches are cut so that the kernel's 32-bit frame offsets hold
        const long long sms = c->sm_count, rbs = t.M / kM4rRows;
        const long long chunk_max = std::max<long long>(1024, ((1ll << 32) / std::max(t.N / 2, p.nbytes / 4) - 1) / 1024 * 1024);
        for (long long done = 0; done < (long long)n_frames; done += chunk_max) {
            const long long nf = std::min<long long>(chunk_max, (long long)n_frames - done);
            EncodeParams q = p;
            q.in = p.in + done * p.nbytes; q.out = p.out + done * t.N; q.n_frames = nf;
            // Frames per tile: a unit costs a + b * frames / 128 (table stages of the row block, look-ups and
            // stores of its frames; measured a : b = 2.85) and the slowest SM works on ceil(units / SMs) of
            // them, so take the tile count that minimises ceil(tiles * RB / SMs) * (2.85 + frames / 128).
            const bool fr = c->m4r_ring == 40 || c->m4r_ring == 60;
            const int tpf = fr ? c->m4r_tpf : 8;
            const long long slots = fr ? kM4rFrSlots : kM4rSlots, cap = slots * tpf;
            const M4rTiling tiling = m4r_choose_tiling(nf, rbs, sms, slots, cap, c->m4r_tile);
            M4rRun run;
            run.tile_frames = (uint32_t)tiling.tile_frames;
            run.flags = c->m4r_flags;
            const long long units = tiling.units;
            const int grid = (int)std::min<long long>(units, sms);
            cudaError_t e;
            if (fr) {
                uint4 *park = nullptr;
                for (auto &pk : c->m4r_parks) if (pk.first == st) park = pk.second;
                if (!park) {
                    if (cudaMalloc(reinterpret_cast<void **>(&park), encode_m4r_fr_park_bytes((int)sms)) != cudaSuccess)
                        return fail(LDPC535_ERR_CUDA, "look-up encoder scratch allocation");
                    c->m4r_parks.emplace_back(st, park);
                }
                M4rFrRun fr_run;
                fr_run.tile_frames = run.tile_frames; fr_run.n_frames = (uint32_t)nf; fr_run.n_units = (uint32_t)units;
                fr_run.G = (uint32_t)(t.K / 8); fr_run.RB = (uint32_t)rbs;
                fr_run.sys_words = (fr_run.G / fr_run.RB) >> 2; fr_run.n_pairs = fr_run.sys_words * tpf;
                fr_run.in_wstride = (uint32_t)p.nbytes >> 2; fr_run.out_qstride = (uint32_t)t.N >> 1;
                fr_run.in_tstride = fr_run.in_wstride * kM4rFrSlots; fr_run.out_tstride = fr_run.out_qstride * kM4rFrSlots;
                fr_run.half_m = (uint32_t)t.M >> 1;
                fr_run.grid_tiles = (uint32_t)grid / fr_run.RB; fr_run.grid_rbs = (uint32_t)grid % fr_run.RB;
                fr_run.drain = fr_run.G >= 64u * tpf ? 1u : 0u;
                void (*kern)(const uint32_t *, float4 *, const unsigned char *, const M4rFrRun, uint32_t *) =
                    c->m4r_ring == 60 ? (tpf == 7 ? encode_m4r_fr_kernel<7, 6> : encode_m4r_fr_kernel<8, 6>)
                                      : (tpf == 7 ? encode_m4r_fr_kernel<7, 4> : encode_m4r_fr_kernel<8, 4>);
                const size_t smem = c->m4r_ring == 60 ? encode_m4r_fr_smem_bytes<8, 6>() : encode_m4r_fr_smem_bytes<8, 4>();
                if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
                    return fail(LDPC535_ERR_CUDA, cudaGetErrorString(e));
                kern<<<grid, kM4rThreads, smem, st>>>(reinterpret_cast<const uint32_t *>(q.in), reinterpret_cast<float4 *>(q.out),
                                                      reinterpret_cast<const unsigned char *>(c->d_m4r), fr_run,
                                                      reinterpret_cast<uint32_t *>(park));
            } else {
                void (*kern)(const EncodeParams, const uint4 *, const M4rRun) =
                    c->m4r_ring == 62 ? encode_m4r_kernel<8, 6, 2> : encode_m4r_kernel<8, 4, 2>;
                const size_t smem = c->m4r_ring == 42 ? encode_m4r_smem_bytes<8, 4>() : encode_m4r_smem_bytes<8, 6>();
                if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
                    return fail(LDPC535_ERR_CUDA, cudaGetErrorString(e));
                kern<<<grid, kM4rThreads, smem, st>>>(q, reinterpret_cast<const uint4 *>(c->d_m4r), run);
            }
            if (done + chunk_max < (long long)n_frames) {
                if ((e = cudaGetLastError()) != cudaSuccess) return fail(LDPC535_ERR_CUDA, std::string("encode launch: ") + cudaGetErrorString(e));
                c->launches++;
            }
        }
    } else {
        const size_t smem = encode_generic_smem_bytes(t.kwords, t.mwords);
        if (smem > c->smem_optin) return fail(LDPC535_ERR_UNSUPPORTED, "code too large for the encoder tile");
        cudaError_t e = cudaFuncSetAttribute(encode_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(LDPC535_ERR_CUDA, cudaGetErrorString(e));
        const long long tiles = ((long long)n_frames + kEncTile - 1) / kEncTile;
        // fewer tiles than SMs: several CTAs share one, each with a whole number of 1024-row passes (even row counts).
        // Measured on the n = 8192 code: 500 frames 0.145 -> 0.054 ms, 1 000 frames 0.143 -> 0.086 ms; splitting beyond one
        // CTA per SM only adds overhead (2 000 frames: 0.136 -> 0.162 ms with 4 CTAs per tile).
        const int passes = std::max(1, t.M / (kEncThreads * kEncRows));
        p.row_splits = 1;
        if ((t.M % (kEncThreads * kEncRows)) == 0 && (t.K % 2) == 0)
            while (p.row_splits * 2 <= passes && passes % (p.row_splits * 2) == 0 && tiles * p.row_splits * 2 <= (long long)c->sm_count) p.row_splits *= 2;
        const int grid = (int)std::min<long long>(tiles * p.row_splits, (long long)c->sm_count * 4);
        encode_generic_kernel<<<grid, kEncThreads, smem, st>>>(p);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(LDPC535_ERR_CUDA, std::string("encode launch: ") + cudaGetErrorString(e));
    c->launches++;
    return LDPC535_OK;
}

// Streams and events of the pipeline slots; idempotent per slot, so a failure half way leaves
// nothing that a retry would allocate twice.
int ensure_slots(ldpc535_code *c)
{
    const size_t frame_bytes = (size_t)c->t.N * 8;
    c->max_win_per_chunk = std::max<size_t>(1, kChunkSymBytes / frame_bytes);
    // the thread-per-codeword kernel works in waves of sm_count x 256 windows: a chunk that is a whole
    // number of waves leaves no SM idle at its end (262 144 windows were 6.9 waves on 148 SMs)
    if (c->is_c4) {
        const size_t wave = (size_t)c->sm_count * c4::kThreads;
        if (c->max_win_per_chunk >= wave) c->max_win_per_chunk -= c->max_win_per_chunk % wave;
    }
    for (auto &s : c->slots) {
        if (!s.stream) CU(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        if (!s.done) CU(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    }
    for (int i = 0; i < kH2dRing; i++) {
        if (!c->h2d_t0[i]) CU(cudaEventCreate(&c->h2d_t0[i]));
        if (!c->h2d_t1[i]) CU(cudaEventCreate(&c->h2d_t1[i]));
    }
    return LDPC535_OK;
}

size_t grow_to(size_t need, size_t floor_, size_t cap)
{
    size_t v = floor_;
    while (v < need) v <<= 1;
    return std::max(need, std::min(v, std::max(cap, need)));
}

// Make slot `s` hold sym_bytes of symbols, n_win windows and stage_bytes of pinned staging.  The
// slot's stream must be idle (callers synchronise it before re-using the slot).  Buffers are
// released before they are re-allocated and their pointers cleared, so an allocation failure
// leaves the slot consistent (release() and a later retry both work).
int slot_reserve(ldpc535_code *c, Slot &s, size_t sym_bytes, size_t n_win, size_t stage_bytes)
{
    const size_t frame_bytes = (size_t)c->t.N * 8;
    const size_t nb = (size_t)(c->t.K + 7) / 8;
    if (sym_bytes > s.cap_sym) {
        const size_t want = grow_to(sym_bytes, (size_t)1 << 20, std::max(kChunkSymBytes, frame_bytes));
        cudaFree(s.d_sym); s.d_sym = nullptr; s.cap_sym = 0;
        CU(cudaMalloc(&s.d_sym, want));
        s.cap_sym = want;
    }
    if (n_win > s.cap_win) {
        const size_t want = grow_to(n_win, 4096, c->max_win_per_chunk);
        cudaFree(s.d_off); cudaFree(s.d_pol); cudaFree(s.d_bytes); cudaFree(s.d_synd); cudaFree(s.d_iters);
        if (s.h_off) cudaFreeHost(s.h_off);
        s.d_off = nullptr; s.d_pol = nullptr; s.d_bytes = s.d_synd = s.d_iters = nullptr; s.h_off = nullptr;
        s.cap_win = 0;
        CU(cudaMalloc(reinterpret_cast<void **>(&s.d_off), want * sizeof(long long)));
        CU(cudaMalloc(reinterpret_cast<void **>(&s.d_pol), want));
        CU(cudaMalloc(reinterpret_cast<void **>(&s.d_bytes), want * nb));
        CU(cudaMalloc(reinterpret_cast<void **>(&s.d_synd), want));
        CU(cudaMalloc(reinterpret_cast<void **>(&s.d_iters), want));
        CU(cudaMallocHost(reinterpret_cast<void **>(&s.h_off), want * sizeof(long long)));
        s.cap_win = want;
    }
    if (stage_bytes > s.cap_stage) {
        const size_t want = grow_to(stage_bytes, (size_t)1 << 20, std::max(kChunkSymBytes / 2, frame_bytes / 2));
        if (s.h_stage) cudaFreeHost(s.h_stage);
        s.h_stage = nullptr; s.cap_stage = 0;
        CU(cudaMallocHost(reinterpret_cast<void **>(&s.h_stage), want));
        s.cap_stage = want;
    }
    return LDPC535_OK;
}

// On ANY exit of a host-buffer entry point -- error returns included -- wait for every slot stream:
// async copies into the caller's output arrays and out of its (pinned) input must not outlive the
// call, the caller is free to release those buffers as soon as it has the status.
struct SlotDrain {
    ldpc535_code *c;
    ~SlotDrain()
    {
        for (auto &s : c->slots)
            if (s.stream) cudaStreamSynchronize(s.stream);
    }
};

// true when `p` is page-locked host memory the copy engine can read directly
bool host_pointer_is_pinned(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

int create_common(const int32_t *row_ptr, const int32_t *col_idx, int M, int N, int device,
                  ldpc535_code **out)
{
    if (!out) return fail(LDPC535_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (device != LDPC535_DEVICE_NONE) {
        int count = 0;
        if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
            cudaGetLastError();
            return fail(LDPC535_ERR_NO_DEVICE, "no CUDA device (there is no CPU fallback)");
        }
        if (device < 0 || device >= count) return fail(LDPC535_ERR_NO_DEVICE, "device index out of range");
    }
    ldpc535_code *c = new (std::nothrow) ldpc535_code();
    if (!c) return fail(LDPC535_ERR_NOMEM, "out of host memory");
    c->device = device;
    int st = build_code_tables(row_ptr, col_idx, M, N, c->t);
    if (st) {
        delete c;
        return fail(st, st == LDPC535_ERR_SINGULAR ? "reorderHMatrix: a row has no pivot (singular L/U)"
                                                   : "invalid parity-check matrix");
    }
    if (getenv("LDPC535_VERBOSE"))
        fprintf(stderr, "ldpc535: code %dx%d E=%d dc<=%d dv<=%d; message-array placement leaves %d extra "
                        "shared-memory wavefronts over %d variable-phase access groups\n",
                c->t.M, c->t.N, c->t.E, c->t.dc_max, c->t.dv_max, c->t.bank_extra_wavefronts, c->t.bank_groups);
    if (device == LDPC535_DEVICE_NONE) {     // tables only: introspection works, compute refuses
        *out = c;
        return LDPC535_OK;
    }
    DeviceGuard g(device);
    if (!g.ok) { delete c; return fail(LDPC535_ERR_CUDA, "cudaSetDevice failed"); }
    st = finish_create(c);
    if (st) { release(c); return st; }
    *out = c;
    return LDPC535_OK;
}

}  // namespace

// ===================================================================================
extern "C" {

const char *ldpc535_version(void) { return "ldpc535 0.1 (sm_100a)"; }

const char *ldpc535_strerror(int status)
{
    switch (status) {
    case LDPC535_OK: return "ok";
    case LDPC535_ERR_INVALID: return "invalid argument";
    case LDPC535_ERR_NO_DEVICE: return "no usable CUDA device";
    case LDPC535_ERR_CUDA: return "CUDA error";
    case LDPC535_ERR_SINGULAR: return "parity-check matrix has no LU pivot order (singular)";
    case LDPC535_ERR_UNSUPPORTED: return "unsupported code or kernel";
    case LDPC535_ERR_NOMEM: return "out of memory";
    default: return "unknown status";
    }
}

const char *ldpc535_last_error(void) { return g_last_error.c_str(); }

int ldpc535_device_count(int *count)
{
    if (!count) return fail(LDPC535_ERR_INVALID, "count is NULL");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); n = 0; }
    *count = n;
    return LDPC535_OK;
}

int ldpc535_device_info(int device, char *name, int *sm_major, int *sm_minor, int *sm_count,
                        int *sm_clock_khz)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) {
        cudaGetLastError();
        return fail(LDPC535_ERR_NO_DEVICE, "no such device");
    }
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (name) { strncpy(name, prop.name, 127); name[127] = 0; }
    if (sm_major) *sm_major = prop.major;
    if (sm_minor) *sm_minor = prop.minor;
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (sm_clock_khz) {
        int khz = 0;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
        *sm_clock_khz = khz;
    }
    return LDPC535_OK;
}

int ldpc535_code_create_sparse(const int32_t *row_ptr, const int32_t *col_idx, int M, int N,
                               int device, ldpc535_code **out)
{
    return create_common(row_ptr, col_idx, M, N, device, out);
}

int ldpc535_code_create(const int32_t *H, int M, int N, int device, ldpc535_code **out)
{
    if (!H || M < 1 || N < 1) return fail(LDPC535_ERR_INVALID, "bad H");
    std::vector<int32_t> row_ptr(M + 1, 0), col_idx;
    for (int j = 0; j < M; j++) {
        for (int i = 0; i < N; i++)
            if (H[(size_t)j * N + i] != 0) col_idx.push_back(i);
        row_ptr[j + 1] = (int32_t)col_idx.size();
    }
    return create_common(row_ptr.data(), col_idx.data(), M, N, device, out);
}

int ldpc535_code_create_default(int device, ldpc535_code **out)
{
    return create_common(ldpc535_default_row_ptr, ldpc535_default_col_idx, LDPC535_DEFAULT_M,
                         LDPC535_DEFAULT_N, device, out);
}

void ldpc535_code_destroy(ldpc535_code *code) { release(code); }

int ldpc535_code_info(const ldpc535_code *c, int *M, int *N, int *K, int *E, int *device)
{
    if (!c) return fail(LDPC535_ERR_INVALID, "code is NULL");
    if (M) *M = c->t.M;
    if (N) *N = c->t.N;
    if (K) *K = c->t.K;
    if (E) *E = c->t.E;
    if (device) *device = c->device;
    return LDPC535_OK;
}

int ldpc535_code_get_pivots(const ldpc535_code *c, int32_t *chosen)
{
    if (!c || !chosen) return fail(LDPC535_ERR_INVALID, "NULL argument");
    memcpy(chosen, c->t.pivots.data(), sizeof(int32_t) * c->t.M);
    return LDPC535_OK;
}

int ldpc535_code_get_h(const ldpc535_code *c, int32_t *row_ptr, int32_t *col_idx)
{
    if (!c || !row_ptr || !col_idx) return fail(LDPC535_ERR_INVALID, "NULL argument");
    memcpy(row_ptr, c->t.row_ptr.data(), sizeof(int32_t) * (c->t.M + 1));
    memcpy(col_idx, c->t.col_idx.data(), sizeof(int32_t) * c->t.E);
    return LDPC535_OK;
}

int ldpc535_code_get_generator(const ldpc535_code *c, uint32_t *P)
{
    if (!c || !P) return fail(LDPC535_ERR_INVALID, "NULL argument");
    memcpy(P, c->t.P.data(), sizeof(uint32_t) * c->t.P.size());
    return LDPC535_OK;
}

const char *ldpc535_code_kernel_name(const ldpc535_code *c, int method)
{
    if (!c) return "";
    switch (resolve_family(c, c->forced, norm_method(method))) {
    case kWarp: return "warp";
    case kBlock: return "block";
    case kC4Thread: return "c4-thread";
    case kC4Refill: return "c4-refill";
    case kC4Adaptive: return "c4-adaptive";
    case kRegular: return "regular";
    case kHard64: return "hard64";
    default: return "unsupported";
    }
}

const char *ldpc535_code_kernel_for(const ldpc535_code *c, int method, int early_stop, size_t n_win)
{
    if (!c) return "";
    const int m = norm_method(method);
    const int f = resolve_family(c, c->forced, m, early_stop ? 1 : 0, (long long)n_win);
    switch (f) {
    case kC4Refill: return "c4-refill";
    case kC4Adaptive: return "c4-adaptive";
    case kWarp: return "warp";
    case kBlock: return "block";
    case kC4Thread: return "c4-thread";
    case kRegular: return "regular";
    case kHard64: return "hard64";
    default: return "unsupported";
    }
}

int ldpc535_code_set_kernel(ldpc535_code *c, const char *kernel)
{
    if (!c) return fail(LDPC535_ERR_INVALID, "code is NULL");
    int f;
    if (family_from_name(kernel, &f)) return fail(LDPC535_ERR_INVALID, "unknown kernel family");
    if (f == kWarp && !c->fits_warp) return fail(LDPC535_ERR_UNSUPPORTED, "code does not fit the warp kernel");
    if (f == kBlock && !c->fits_block) return fail(LDPC535_ERR_UNSUPPORTED, "code does not fit the block kernel");
    if ((f == kC4Thread || f == kC4Refill || f == kC4Adaptive) && !c->is_c4) return fail(LDPC535_ERR_UNSUPPORTED, "not the shipped 32x64 code");
    if (f == kRegular && !c->fits_regular) return fail(LDPC535_ERR_UNSUPPORTED, "not a (3,6)-regular code that fits shared memory");
    c->forced = f;
    return LDPC535_OK;
}

uint64_t ldpc535_launch_count(const ldpc535_code *c) { return c ? c->launches : 0; }

int ldpc535_code_host_path(const ldpc535_code *c, int *pack_pinned, int *pack_threads)
{
    if (!c) return fail(LDPC535_ERR_INVALID, "code is NULL");
    if (pack_pinned) *pack_pinned = c->pack_pinned;
    if (pack_threads) *pack_threads = c->pack_threads;
    return LDPC535_OK;
}

int ldpc535_code_host_stats(const ldpc535_code *c, uint64_t *chunks_packed, uint64_t *chunks_raw, uint64_t *h2d_bytes)
{
    if (!c) return fail(LDPC535_ERR_INVALID, "code is NULL");
    if (chunks_packed) *chunks_packed = c->chunks_packed;
    if (chunks_raw) *chunks_raw = c->chunks_raw;
    if (h2d_bytes) *h2d_bytes = c->h2d_bytes;
    return LDPC535_OK;
}

int ldpc535_code_set_host_path(ldpc535_code *c, int pack_pinned, int pack_threads)
{
    if (!c) return fail(LDPC535_ERR_INVALID, "code is NULL");
    if (pack_threads < 0 || pack_threads > 256) return fail(LDPC535_ERR_INVALID, "pack_threads must be 0 (keep) .. 256");
    if (pack_threads > 0 && pack_threads != c->pack_threads) {
        c->pack_threads = pack_threads;
        delete c->pool;                     // the team is re-created with the new size on next use
        c->pool = nullptr;
    }
    if (pack_pinned > 2) return fail(LDPC535_ERR_INVALID, "pack_pinned must be -1 (default), 0, 1 or 2");
    c->pack_pinned = pack_pinned < 0 ? kPackAdaptive : pack_pinned;
    c->pack_ns_per_byte = 0;                // measured again with the new team
    return LDPC535_OK;
}

// ---- measurement utilities (synth_kernels.cuh) -------------------------------------------------
int ldpc535_synth_bytes_dev(ldpc535_code *c, uint64_t seed, uint64_t first_frame, size_t n_frames,
                            uint8_t *d_bytes, void *stream)
{
    if (!c || (n_frames && !d_bytes)) return fail(LDPC535_ERR_INVALID, "NULL argument");
    DeviceGuard g(c->device);
    NEED_DEVICE(c, g);
    if (!n_frames) return LDPC535_OK;
    const int nb = (c->t.K + 7) / 8;
    const long long total = (long long)n_frames * ((nb + 15) / 16);
    const int grid = (int)std::min<long long>((total + 255) / 256, (long long)c->sm_count * 16);
    synth_bytes_kernel<<<grid, 256, 0, stream ? (cudaStream_t)stream : c->stream>>>(
        d_bytes, (long long)first_frame, (long long)n_frames, nb, (uint32_t)seed, (uint32_t)(seed >> 32));
    CU(cudaGetLastError());
    return LDPC535_OK;
}

int ldpc535_synth_awgn_dev(ldpc535_code *c, uint64_t seed, uint64_t first_frame, size_t n_frames,
                           float sigma, float *d_sym, void *stream)
{
    if (!c || (n_frames && !d_sym)) return fail(LDPC535_ERR_INVALID, "NULL argument");
    if (c->t.N % 4 || (reinterpret_cast<uintptr_t>(d_sym) & 15))
        return fail(LDPC535_ERR_INVALID, "N must be a multiple of 4 and d_sym 16-byte aligned");
    DeviceGuard g(c->device);
    NEED_DEVICE(c, g);
    if (!n_frames) return LDPC535_OK;
    const long long total = (long long)n_frames * (c->t.N / 4);
    const int grid = (int)std::min<long long>((total + 255) / 256, (long long)c->sm_count * 16);
    synth_awgn_kernel<<<grid, 256, 0, stream ? (cudaStream_t)stream : c->stream>>>(
        reinterpret_cast<float2 *>(d_sym), (long long)first_frame, (long long)n_frames, c->t.N, sigma,
        (uint32_t)seed, (uint32_t)(seed >> 32));
    CU(cudaGetLastError());
    return LDPC535_OK;
}

int ldpc535_probe_pipe_peak(ldpc535_code *c, int which, double *ops_per_s)
{
    if (!c || !ops_per_s || (which != LDPC535_PIPE_MUFU && which != LDPC535_PIPE_FP64))
        return fail(LDPC535_ERR_INVALID, "bad probe arguments");
    DeviceGuard g(c->device);
    NEED_DEVICE(c, g);
    void *out = nullptr;
    cudaEvent_t a = nullptr, b = nullptr;
    double best = 0;
    cudaError_t e = cudaMalloc(&out, (size_t)c->sm_count * 8 * 256 * 8);
    if (e == cudaSuccess) e = cudaEventCreate(&a);
    if (e == cudaSuccess) e = cudaEventCreate(&b);
    const int iters = 4000;
    for (int ctas = 2; ctas <= 8 && e == cudaSuccess; ctas *= 2) {          // 16, 32, 64 warps per SM
        const int grid = c->sm_count * ctas;
        for (int rep = 0; rep < 3 && e == cudaSuccess; rep++) {             // rep 0 warms up
            cudaEventRecord(a, c->stream);
            if (which == LDPC535_PIPE_MUFU) probe_mufu_kernel<<<grid, 256, 0, c->stream>>>((float *)out, iters, 0.5f);
            else probe_fp64_kernel<<<grid, 256, 0, c->stream>>>((double *)out, iters, 0.5);
            cudaEventRecord(b, c->stream);
            e = cudaEventSynchronize(b);
            float ms = 0;
            if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, a, b);
            // per chain step: 3 MUFU, or 3 fp64-pipe instructions (2 DADD + 1 DSETP)
            const double ops = (double)grid * 256 * iters * 8 * 3;
            if (rep && ms > 0) best = std::max(best, ops / (ms * 1e-3));
        }
    }
    if (a) cudaEventDestroy(a);
    if (b) cudaEventDestroy(b);
    cudaFree(out);
    if (e != cudaSuccess) return fail(LDPC535_ERR_CUDA, std::string("pipe probe: ") + cudaGetErrorString(e));
    *ops_per_s = best;
    return LDPC535_OK;
}

int ldpc535_host_alloc(size_t bytes, void **ptr)
{
    if (!ptr) return fail(LDPC535_ERR_INVALID, "ptr is NULL");
    CU(cudaMallocHost(ptr, bytes ? bytes : 1));
    return LDPC535_OK;
}

int ldpc535_host_free(void *ptr)
{
    if (ptr) CU(cudaFreeHost(ptr));
    return LDPC535_OK;
}

int ldpc535_dev_alloc(ldpc535_code *c, size_t bytes, void **dptr)
{
    if (!c || !dptr) return fail(LDPC535_ERR_INVALID, "NULL argument");
    DeviceGuard g(c->device);
    NEED_DEVICE(c, g);
    CU(cudaMalloc(dptr, bytes ? bytes : 1));
    return LDPC535_OK;
}

int ldpc535_dev_free(ldpc535_code *c, void *dptr)
{
    if (!c) return fail(LDPC535_ERR_INVALID, "code is NULL");
    DeviceGuard g(c->device);
    NEED_DEVICE(c, g);
    if (dptr) CU(cudaFree(dptr));
    return LDPC535_OK;
}

int ldpc535_memcpy_h2d(ldpc535_code *c, void *dptr, const void *hptr, size_t bytes)
{
    if (!c) return fail(LDPC535_ERR_INVALID, "code is NULL");
    DeviceGuard g(c->device);
    NEED_DEVICE(c, g);
    CU(cudaMemcpyAsync(dptr, hptr, bytes, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return LDPC535_OK;
}

int ldpc535_memcpy_d2h(ldpc535_code *c, void *hptr, const void *dptr, size_t bytes)
{
    if (!c) return fail(LDPC535_ERR_INVALID, "code is NULL");
    DeviceGuard g(c->device);
    NEED_DEVICE(c, g);
    CU(cudaMemcpyAsync(hptr, dptr, bytes, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return LDPC535_OK;
}

int ldpc535_stream_sync(ldpc535_code *c, void *stream)
{
    if (!c) return fail(LDPC535_ERR_INVALID, "code is NULL");
    DeviceGuard g(c->device);
    NEED_DEVICE(c, g);
    CU(cudaStreamSynchronize(stream ? (cudaStream_t)stream : c->stream));
    return LDPC535_OK;
}

int ldpc535_encode_batch_dev(ldpc535_code *c, const uint8_t *d_in, size_t n_frames, float *d_out,
                             void *stream)
{
    if (!c || (n_frames && (!d_in || !d_out))) return fail(LDPC535_ERR_INVALID, "NULL argument");
    // the kernels read the input as 32-bit words and store symbols 16 bytes at a time
    if ((reinterpret_cast<uintptr_t>(d_in) & 3) || (reinterpret_cast<uintptr_t>(d_out) & 15))
        return fail(LDPC535_ERR_INVALID, "device pointers must be aligned: d_in to 4 bytes, d_out to 16 bytes");
    DeviceGuard g(c->device);
    NEED_DEVICE(c, g);
    return launch_encode(c, d_in, n_frames, d_out, stream ? (cudaStream_t)stream : c->stream);
}

int ldpc535_decode_batch_dev(ldpc535_code *c, const float *d_sym, size_t n_sym,
                             const int64_t *d_win_offset, const int8_t *d_polarity, size_t n_win,
                             int method, int max_iters, int early_stop, int synd_threshold,
                             uint8_t *d_out_bytes, uint8_t *d_out_synd, uint8_t *d_out_iters,
                             void *stream)
{
    if (!c || (n_win && (!d_sym || !d_out_bytes))) return fail(LDPC535_ERR_INVALID, "NULL argument");
    if (max_iters < 1 || max_iters > 255) return fail(LDPC535_ERR_INVALID, "max_iters must be 1..255");
    if (synd_threshold < 0) return fail(LDPC535_ERR_INVALID, "synd_threshold < 0");
    // 128-bit symbol loads, 32-bit stores of packed bytes
    if ((reinterpret_cast<uintptr_t>(d_sym) & 15) || (reinterpret_cast<uintptr_t>(d_out_bytes) & 3) ||
        (reinterpret_cast<uintptr_t>(d_win_offset) & 7))
        return fail(LDPC535_ERR_INVALID, "device pointers must be aligned: d_sym to 16 bytes, d_out_bytes to 4 bytes, d_win_offset to 8 bytes");
    DeviceGuard g(c->device);
    NEED_DEVICE(c, g);
    DecodeParams p = {};
    p.sym = reinterpret_cast<const float2 *>(d_sym);
    p.n_sym = (long long)n_sym;
    p.win_offset = reinterpret_cast<const long long *>(d_win_offset);
    p.polarity = reinterpret_cast<const signed char *>(d_polarity);
    p.n_win = (long long)n_win;
    p.max_iters = max_iters; p.early_stop = early_stop ? 1 : 0; p.thr = std::min(synd_threshold, 253);
    p.out_bytes = d_out_bytes; p.out_synd = d_out_synd; p.out_iters = d_out_iters;
    return launch_decode(c, c->forced, method, false, p, stream ? (cudaStream_t)stream : c->stream);
}

int ldpc535_encode_batch(ldpc535_code *c, const uint8_t *in, size_t n_frames, float *out)
{
    if (!c || (n_frames && (!in || !out))) return fail(LDPC535_ERR_INVALID, "NULL argument");
    DeviceGuard g(c->device);
    NEED_DEVICE(c, g);
    int st = ensure_slots(c);
    if (st) return st;
    SlotDrain drain{c};
    const size_t nb = (size_t)(c->t.K + 7) / 8;
    const size_t frame_bytes = (size_t)c->t.N * 8;
    // the symbol staging buffer holds the OUTPUT here; input bytes ride in d_bytes
    const size_t chunk = c->max_win_per_chunk;
    size_t done = 0;
    int k = 0;
    while (done < n_frames) {
        Slot &s = c->slots[k % kSlots];
        const size_t n = std::min(chunk, n_frames - done);
        CU(cudaStreamSynchronize(s.stream));
        if ((st = slot_reserve(c, s, n * frame_bytes, n, 0))) return st;
        CU(cudaMemcpyAsync(s.d_bytes, in + done * nb, n * nb, cudaMemcpyHostToDevice, s.stream));
        st = launch_encode(c, s.d_bytes, n, reinterpret_cast<float *>(s.d_sym), s.stream);
        if (st) return st;
        CU(cudaMemcpyAsync(out + done * c->t.N * 2, s.d_sym, n * frame_bytes, cudaMemcpyDeviceToHost, s.stream));
        done += n;
        k++;
    }
    for (auto &s : c->slots) CU(cudaStreamSynchronize(s.stream));
    return LDPC535_OK;
}

int ldpc535_decode_batch(ldpc535_code *c, const float *sym, size_t n_sym, const int64_t *win_offset,
                         const int8_t *polarity, size_t n_win, int method, int max_iters,
                         int early_stop, int synd_threshold, uint8_t *out_bytes, uint8_t *out_synd,
                         uint8_t *out_iters)
{
    if (!c || (n_win && (!sym || !out_bytes))) return fail(LDPC535_ERR_INVALID, "NULL argument");
    if (max_iters < 1 || max_iters > 255) return fail(LDPC535_ERR_INVALID, "max_iters must be 1..255");
    if (synd_threshold < 0) return fail(LDPC535_ERR_INVALID, "synd_threshold < 0");
    const size_t N = (size_t)c->t.N;
    const size_t nb = (size_t)(c->t.K + 7) / 8;
    if (!win_offset) {
        if (n_win * N > n_sym) return fail(LDPC535_ERR_INVALID, "n_win * N exceeds n_sym");
    } else {
        for (size_t w = 0; w < n_win; w++)
            if (win_offset[w] < 0 || (size_t)win_offset[w] + N > n_sym)
                return fail(LDPC535_ERR_INVALID, "window offset out of range");
    }
    DeviceGuard g(c->device);
    NEED_DEVICE(c, g);
    int st = ensure_slots(c);
    if (st) return st;
    SlotDrain drain{c};
    // Pageable input (a GNU Radio buffer) must be staged through pinned memory anyway: stage only
    // the real parts.  Pinned input goes to the copy engine as it is unless the handle packs pinned
    // input too (ldpc535_code_set_host_path).
    const int pack_mode = !n_win ? 0 : !host_pointer_is_pinned(sym) ? 1 : c->pack_pinned;
    if (pack_mode && !c->pool) {
        c->pool = new (std::nothrow) PackPool(c->pack_threads);
        if (!c->pool) return fail(LDPC535_ERR_NOMEM, "out of host memory");
    }
    for (int i = 0; i < kH2dRing; i++) c->h2d_len[i] = 0;   // events of an earlier call say nothing about this one's queue
    c->host_calls++;
    const size_t max_sym = kChunkSymBytes / 8;     // symbols per staging buffer
    size_t done = 0;
    int k = 0;
    while (done < n_win) {
        Slot &s = c->slots[k % kSlots];
        CU(cudaStreamSynchronize(s.stream));
        size_t n = 0, sym_lo = 0, sym_hi = 0;
        if (!win_offset) {
            n = std::min(c->max_win_per_chunk, n_win - done);
            sym_lo = done * N;
            sym_hi = sym_lo + n * N;
        } else {
            // grow the chunk while the span of symbols it touches fits the staging buffer
            size_t lo = (size_t)win_offset[done], hi = lo + N;
            n = 1;
            while (done + n < n_win && n < c->max_win_per_chunk) {
                const size_t o = (size_t)win_offset[done + n];
                const size_t nlo = std::min(lo, o), nhi = std::max(hi, o + N);
                if (nhi - nlo > max_sym) break;
                lo = nlo; hi = nhi; n++;
            }
            sym_lo = lo; sym_hi = hi;
        }
        // the copy of chunk k - kSlots (this slot's previous one) is over: its time per byte, if it had the copy
        // engines to itself (issued with no earlier copy pending -- two engines share the link, and a queued copy's
        // start event fires long before its first byte moves)
        const int ring = k % kH2dRing;
        if (k >= kSlots) {
            const int j = (k - kSlots) % kH2dRing;
            float ms = 0.f;
            if (c->h2d_len[j] && c->h2d_alone[j] && cudaEventElapsedTime(&ms, c->h2d_t0[j], c->h2d_t1[j]) == cudaSuccess && ms > 0.f) {
                const double v = (double)ms * 1e6 / (double)c->h2d_len[j];
                c->h2d_ns_per_byte = c->h2d_ns_per_byte > 0 ? 0.5 * c->h2d_ns_per_byte + 0.5 * v : v;
            }
            cudaGetLastError();
        }
        bool alone = true;
        for (int b = 1; b < kSlots && b <= k; b++) {
            const int j = (k - b) % kH2dRing;
            if (c->h2d_len[j] && cudaEventQuery(c->h2d_t1[j]) == cudaErrorNotReady) alone = false;
        }
        cudaGetLastError();
        c->h2d_alone[ring] = alone;
        const size_t span = sym_hi - sym_lo;
        bool pack = pack_mode == 1;
        if (pack_mode == kPackAdaptive) {
            // Packing S symbols holds this thread for 8 S p and leaves a 4 S h copy; the raw copy takes 8 S h (p, h =
            // measured ns per byte of the team and of the copy engine): packing wins when p < h.  One regime at a
            // time, with hysteresis -- mixing the two was measured and loses: packing only the chunks that fit under
            // the copies already queued moved 20 % fewer PCIe bytes in the same time with two ranks on a 24-core
            // host (both paths share the host's memory system), and flipping between the regimes chunk by chunk
            // was slower than either.  While the input goes raw, the second chunk of every 8th call is packed to
            // keep p current.
            if (c->pack_ns_per_byte <= 0 || (!c->packing && (c->host_calls & 7u) == 0)) pack = k == 1;
            if (c->pack_ns_per_byte > 0 && c->h2d_ns_per_byte > 0) {
                c->packing = c->pack_ns_per_byte * (c->packing ? 1.05 : 1.25) < c->h2d_ns_per_byte;
                pack = pack || c->packing;
            }
        }
        if ((st = slot_reserve(c, s, span * (pack ? 4 : 8), n, pack ? span * 4 : 0))) return st;
        if (win_offset) {
            for (size_t i = 0; i < n; i++) s.h_off[i] = (long long)((size_t)win_offset[done + i] - sym_lo);
            CU(cudaMemcpyAsync(s.d_off, s.h_off, n * sizeof(long long), cudaMemcpyHostToDevice, s.stream));
        }
        if (pack) {
            const auto t0 = std::chrono::steady_clock::now();
            c->pool->pack(sym + sym_lo * 2, s.h_stage, span);
            const double v = std::chrono::duration<double, std::nano>(std::chrono::steady_clock::now() - t0).count() / ((double)span * 8.0);
            c->pack_ns_per_byte = c->pack_ns_per_byte > 0 ? 0.75 * c->pack_ns_per_byte + 0.25 * v : v;
        }
        const void *h2d_src = pack ? static_cast<const void *>(s.h_stage) : static_cast<const void *>(sym + sym_lo * 2);
        c->h2d_len[ring] = span * (pack ? 4 : 8);
        CU(cudaEventRecord(c->h2d_t0[ring], s.stream));
        CU(cudaMemcpyAsync(s.d_sym, h2d_src, c->h2d_len[ring], cudaMemcpyHostToDevice, s.stream));
        CU(cudaEventRecord(c->h2d_t1[ring], s.stream));
        c->h2d_bytes += c->h2d_len[ring];
        (pack ? c->chunks_packed : c->chunks_raw)++;
        if (polarity)
            CU(cudaMemcpyAsync(s.d_pol, polarity + done, n, cudaMemcpyHostToDevice, s.stream));
        DecodeParams p = {};
        p.sym = reinterpret_cast<const float2 *>(s.d_sym);
        p.sym_re = pack ? reinterpret_cast<const float *>(s.d_sym) : nullptr;
        p.n_sym = (long long)(sym_hi - sym_lo);
        p.win_offset = win_offset ? s.d_off : nullptr;
        p.polarity = polarity ? s.d_pol : nullptr;
        p.n_win = (long long)n;
        p.max_iters = max_iters; p.early_stop = early_stop ? 1 : 0; p.thr = std::min(synd_threshold, 253);
        p.out_bytes = s.d_bytes; p.out_synd = s.d_synd; p.out_iters = s.d_iters;
        st = launch_decode(c, c->forced, method, false, p, s.stream);
        if (st) return st;
        CU(cudaMemcpyAsync(out_bytes + done * nb, s.d_bytes, n * nb, cudaMemcpyDeviceToHost, s.stream));
        if (out_synd) CU(cudaMemcpyAsync(out_synd + done, s.d_synd, n, cudaMemcpyDeviceToHost, s.stream));
        if (out_iters) CU(cudaMemcpyAsync(out_iters + done, s.d_iters, n, cudaMemcpyDeviceToHost, s.stream));
        done += n;
        k++;
    }
    for (auto &s : c->slots) CU(cudaStreamSynchronize(s.stream));
    return LDPC535_OK;
}

int ldpc535_decode_debug(ldpc535_code *c, const float *sym, size_t n_win, int max_iters,
                         int early_stop, const char *kernel, float *out_L, float *out_E,
                         float *out_M, uint8_t *out_bytes, uint8_t *out_iters)
{
    if (!c || !sym) return fail(LDPC535_ERR_INVALID, "NULL argument");
    if (max_iters < 1 || max_iters > 255) return fail(LDPC535_ERR_INVALID, "max_iters must be 1..255");
    int fam;
    if (family_from_name(kernel, &fam)) return fail(LDPC535_ERR_INVALID, "unknown kernel family");
    if (fam == kAuto) fam = c->forced;
    DeviceGuard g(c->device);
    NEED_DEVICE(c, g);
    const size_t N = (size_t)c->t.N, E = (size_t)c->t.E, nb = (size_t)(c->t.K + 7) / 8;
    float *d_sym = nullptr, *dL = nullptr, *dE = nullptr, *dM = nullptr, *dS = nullptr;
    uint8_t *d_bytes = nullptr, *d_iters = nullptr;
    int st = LDPC535_OK;
    auto cleanup = [&]() {
        cudaFree(d_sym); cudaFree(dL); cudaFree(dE); cudaFree(dM); cudaFree(dS); cudaFree(d_bytes); cudaFree(d_iters);
    };
#define CUX(call)                                                                          \
    do {                                                                                   \
        cudaError_t e__ = (call);                                                          \
        if (e__ != cudaSuccess) {                                                          \
            cleanup();                                                                     \
            return fail(LDPC535_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
        }                                                                                  \
    } while (0)
    CUX(cudaMalloc(reinterpret_cast<void **>(&d_sym), std::max<size_t>(n_win * N * 8, 8)));
    CUX(cudaMalloc(reinterpret_cast<void **>(&dL), std::max<size_t>(n_win * N * 4, 4)));
    CUX(cudaMalloc(reinterpret_cast<void **>(&dE), std::max<size_t>(n_win * E * 4, 4)));
    CUX(cudaMalloc(reinterpret_cast<void **>(&dM), std::max<size_t>(n_win * E * 4, 4)));
    CUX(cudaMalloc(reinterpret_cast<void **>(&dS), std::max<size_t>(n_win * E * 4, 4)));
    CUX(cudaMalloc(reinterpret_cast<void **>(&d_bytes), std::max<size_t>(n_win * nb, 1)));
    CUX(cudaMalloc(reinterpret_cast<void **>(&d_iters), std::max<size_t>(n_win, 1)));
    CUX(cudaMemcpyAsync(d_sym, sym, n_win * N * 8, cudaMemcpyHostToDevice, c->stream));
    CUX(cudaMemsetAsync(dL, 0, std::max<size_t>(n_win * N * 4, 4), c->stream));
    CUX(cudaMemsetAsync(dE, 0, std::max<size_t>(n_win * E * 4, 4), c->stream));
    CUX(cudaMemsetAsync(dM, 0, std::max<size_t>(n_win * E * 4, 4), c->stream));
    CUX(cudaMemsetAsync(dS, 0, std::max<size_t>(n_win * E * 4, 4), c->stream));
    DecodeParams p = {};
    p.sym = reinterpret_cast<const float2 *>(d_sym);
    p.n_sym = (long long)(n_win * N);
    p.n_win = (long long)n_win;
    p.max_iters = max_iters; p.early_stop = early_stop ? 1 : 0; p.thr = c->t.M / 8;
    p.out_bytes = d_bytes; p.out_iters = d_iters;
    p.dbgL = dL; p.dbgE = dE; p.dbgM = dM; p.dbgS = dS;
    st = launch_decode(c, fam, LDPC535_METHOD_SUMPRODUCT, true, p, c->stream);
    if (st) { cleanup(); return st; }
    if (out_L) CUX(cudaMemcpyAsync(out_L, dL, n_win * N * 4, cudaMemcpyDeviceToHost, c->stream));
    if (out_E) CUX(cudaMemcpyAsync(out_E, dE, n_win * E * 4, cudaMemcpyDeviceToHost, c->stream));
    if (out_M) CUX(cudaMemcpyAsync(out_M, dM, n_win * E * 4, cudaMemcpyDeviceToHost, c->stream));
    if (out_bytes) CUX(cudaMemcpyAsync(out_bytes, d_bytes, n_win * nb, cudaMemcpyDeviceToHost, c->stream));
    if (out_iters) CUX(cudaMemcpyAsync(out_iters, d_iters, n_win, cudaMemcpyDeviceToHost, c->stream));
    CUX(cudaStreamSynchronize(c->stream));
#undef CUX
    cleanup();
    return LDPC535_OK;
}

}  // extern "C"

// ===================================================================================
// several GPUs from one process: contiguous shards, one host thread per device
struct ldpc535_pool {
    std::vector<ldpc535_code *> codes;
};

extern "C" {

int ldpc535_pool_create(const int32_t *H, int M, int N, const int *devices, int n_devices, ldpc535_pool **out)
{
    if (!out || !devices || n_devices < 1) return fail(LDPC535_ERR_INVALID, "bad pool arguments");
    *out = nullptr;
    ldpc535_pool *p = new (std::nothrow) ldpc535_pool();
    if (!p) return fail(LDPC535_ERR_NOMEM, "out of host memory");
    for (int i = 0; i < n_devices; i++) {
        ldpc535_code *c = nullptr;
        const int st = H ? ldpc535_code_create(H, M, N, devices[i], &c) : ldpc535_code_create_default(devices[i], &c);
        if (st) {
            const std::string msg = g_last_error;
            ldpc535_pool_destroy(p);
            return fail(st, msg);
        }
        p->codes.push_back(c);
    }
    if (n_devices > 1)                      // the handles share this host: split the packing threads
        for (ldpc535_code *c : p->codes)
            ldpc535_code_set_host_path(c, -1, std::max(1, c->pack_threads / n_devices));
    *out = p;
    return LDPC535_OK;
}

void ldpc535_pool_destroy(ldpc535_pool *p)
{
    if (!p) return;
    for (ldpc535_code *c : p->codes) ldpc535_code_destroy(c);
    delete p;
}

int ldpc535_pool_size(const ldpc535_pool *p) { return p ? (int)p->codes.size() : 0; }

ldpc535_code *ldpc535_pool_code(ldpc535_pool *p, int i)
{
    return (p && i >= 0 && i < (int)p->codes.size()) ? p->codes[i] : nullptr;
}

int ldpc535_pool_decode_batch(ldpc535_pool *p, const float *sym, size_t n_sym, const int64_t *win_offset,
                              const int8_t *polarity, size_t n_win, int method, int max_iters, int early_stop,
                              int synd_threshold, uint8_t *out_bytes, uint8_t *out_synd, uint8_t *out_iters)
{
    if (!p || p->codes.empty()) return fail(LDPC535_ERR_INVALID, "pool is NULL");
    const size_t g = p->codes.size();
    const size_t N = (size_t)p->codes[0]->t.N, nb = (size_t)(p->codes[0]->t.K + 7) / 8;
    if (!win_offset && n_win * N > n_sym) return fail(LDPC535_ERR_INVALID, "n_win * N exceeds n_sym");
    const size_t per = (n_win + g - 1) / g;
    std::vector<int> status(g, LDPC535_OK);
    std::vector<std::string> errors(g);
    std::vector<std::thread> threads;
    for (size_t i = 0; i < g; i++) {
        const size_t lo = std::min(n_win, i * per), hi = std::min(n_win, (i + 1) * per);
        if (hi <= lo) continue;
        threads.emplace_back([=, &status, &errors]() {
            // aligned frames: hand the shard its own slice of the symbol buffer; window lists keep
            // absolute offsets into the whole buffer
            const float *s = win_offset ? sym : sym + lo * N * 2;
            const size_t ns = win_offset ? n_sym : (hi - lo) * N;
            status[i] = ldpc535_decode_batch(p->codes[i], s, ns, win_offset ? win_offset + lo : nullptr,
                                             polarity ? polarity + lo : nullptr, hi - lo, method, max_iters,
                                             early_stop, synd_threshold, out_bytes + lo * nb,
                                             out_synd ? out_synd + lo : nullptr, out_iters ? out_iters + lo : nullptr);
            if (status[i]) errors[i] = g_last_error;          // thread-local: copy it out
        });
    }
    for (auto &t : threads) t.join();
    for (size_t i = 0; i < g; i++)
        if (status[i]) return fail(status[i], errors[i]);
    return LDPC535_OK;
}

int ldpc535_pool_encode_batch(ldpc535_pool *p, const uint8_t *in, size_t n_frames, float *out)
{
    if (!p || p->codes.empty()) return fail(LDPC535_ERR_INVALID, "pool is NULL");
    const size_t g = p->codes.size();
    const size_t N = (size_t)p->codes[0]->t.N, nb = (size_t)(p->codes[0]->t.K + 7) / 8;
    const size_t per = (n_frames + g - 1) / g;
    std::vector<int> status(g, LDPC535_OK);
    std::vector<std::string> errors(g);
    std::vector<std::thread> threads;
    for (size_t i = 0; i < g; i++) {
        const size_t lo = std::min(n_frames, i * per), hi = std::min(n_frames, (i + 1) * per);
        if (hi <= lo) continue;
        threads.emplace_back([=, &status, &errors]() {
            status[i] = ldpc535_encode_batch(p->codes[i], in + lo * nb, hi - lo, out + lo * N * 2);
            if (status[i]) errors[i] = g_last_error;
        });
    }
    for (auto &t : threads) t.join();
    for (size_t i = 0; i < g; i++)
        if (status[i]) return fail(status[i], errors[i]);
    return LDPC535_OK;
}

}  // extern "C"
