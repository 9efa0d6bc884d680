/*
 * ldpc535.h -- C ABI of the B200 (sm_100a) LDPC encode/decode hot path.
 *
 * This is the drop-in boundary: the GNU Radio blocks' work() methods (and any
 * other host: the Python binding, the benchmark harness) call these entry
 * points; nothing above this line knows about CUDA.  Plain pointers and sizes,
 * `int` status returns, no exceptions, no globals, no torch types.  There is no
 * CPU fallback: every compute entry point fails with LDPC535_ERR_NO_DEVICE /
 * LDPC535_ERR_CUDA when no sm_100 GPU is usable.
 *
 * What each entry point replaces in the reference (paths relative to the
 * reference checkout):
 *
 *   ldpc535_code_create*        the constructors' H literal + reorderHMatrix
 *                               lib/ldpc_decoder_cb_impl.cc:35-117 (:255-307),
 *                               lib/ldpc_encoder_bc_impl.cc:33-102 (:225-273)
 *   ldpc535_encode_batch[_dev]  byte unpack + makeParityCheck/solve/mod2 + BPSK map
 *                               lib/ldpc_encoder_bc_impl.cc:133-170, :180-223, :275-311
 *   ldpc535_decode_batch[_dev]  LLR formation + decode{SumProductSoft,LogDomainSimple,
 *                               BitFlipping,Hard} + checkFrame + bit packing
 *                               lib/ldpc_decoder_cb_impl.cc:149-166, :207-219, :236-253,
 *                               :309-572
 *   (the sync state machine, :168-205, stays on the host in the block and replays
 *    over the per-window results these calls return)
 *
 * Layouts
 *   symbols   interleaved complex float32 (re, im) == gr_complex; only `re` is used
 *             (lib/ldpc_decoder_cb_impl.cc:151).
 *   frame     N symbols: M parity symbols then K = N - M data symbols, +1+0j for
 *             bit 1 and -1+0j for bit 0 (lib/ldpc_encoder_bc_impl.cc:154-165).
 *   bytes     K/8 bytes per frame (ceil), data bits MSB first
 *             (lib/ldpc_encoder_bc_impl.cc:138-147, lib/ldpc_decoder_cb_impl.cc:209-219).
 *   window    a decode unit: N consecutive symbols starting at symbol offset
 *             win_offset[w] of the symbol buffer, multiplied by polarity[w] (+1/-1).
 *
 * Threading: a handle is bound to one device and owns its streams and staging
 * buffers; calls on one handle must not overlap (GNU Radio's thread-per-block
 * model); different handles may be used concurrently.
 */
#ifndef LDPC535_H
#define LDPC535_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define LDPC535_API
#else
#define LDPC535_API __attribute__((visibility("default")))
#endif

/* status codes */
enum {
    LDPC535_OK = 0,
    LDPC535_ERR_INVALID = 1,      /* bad argument */
    LDPC535_ERR_NO_DEVICE = 2,    /* no CUDA device / device index out of range / not sm_100 */
    LDPC535_ERR_CUDA = 3,         /* a CUDA call failed; see ldpc535_last_error() */
    LDPC535_ERR_SINGULAR = 4,     /* reorderHMatrix found no pivot on some row (the reference
                                     would silently build a singular L/U and emit garbage) */
    LDPC535_ERR_UNSUPPORTED = 5,  /* code too large for the shared-memory resident decoder */
    LDPC535_ERR_NOMEM = 6
};

/* decoder methods: the reference block's constructor argument
 * (lib/ldpc_decoder_cb_impl.cc:155-164, grc/ldpc_ece535a_ldpc_decoder_cb.xml:9-30);
 * any other value means LogDomain, like the reference's final `else`. */
enum {
    LDPC535_METHOD_LOGDOMAIN = 0,   /* min-sum, decodeLogDomainSimple */
    LDPC535_METHOD_SUMPRODUCT = 1,  /* decodeSumProductSoft */
    LDPC535_METHOD_BITFLIP = 2,     /* decodeBitFlipping */
    LDPC535_METHOD_HARD = 3         /* decodeHard */
};

/* Pass as `device` to ldpc535_code_create*: build the code tables only (pivots, re-ordered
 * H, generator are readable; every compute entry point returns LDPC535_ERR_NO_DEVICE).
 * Used by host-only tests; it is NOT a CPU compute path. */
#define LDPC535_DEVICE_NONE (-1)

/* reference defaults (lib/ldpc_decoder_cb_impl.cc:39-40, :141-142) */
#define LDPC535_REF_ITERATIONS 5
#define LDPC535_REF_EARLY_STOP 1

typedef struct ldpc535_code ldpc535_code;   /* opaque: one code on one device */

/* ---- library / device ---------------------------------------------------- */
LDPC535_API const char *ldpc535_version(void);
LDPC535_API const char *ldpc535_strerror(int status);
/* Thread-local text of the last failure on this thread (CUDA error string etc.). */
LDPC535_API const char *ldpc535_last_error(void);
LDPC535_API int ldpc535_device_count(int *count);
/* name: >= 128 bytes.  sm_major/minor, sm_count, sm_clock_khz may be NULL. */
LDPC535_API int ldpc535_device_info(int device, char *name, int *sm_major, int *sm_minor,
                                    int *sm_count, int *sm_clock_khz);

/* ---- code handles -------------------------------------------------------- */
/* Dense row-major 0/1 matrix, as the reference constructors hold it.  Runs the
 * reference's column re-ordering ("First" pivot strategy), builds the decoder's
 * adjacency tables and the encoder's bit-packed generator P = (L U)^-1 B over GF(2),
 * and uploads them to `device`. */
LDPC535_API int ldpc535_code_create(const int32_t *H, int M, int N, int device,
                                    ldpc535_code **out);
/* Same, H given sparse: CSR over check rows (row_ptr[M+1], col_idx[row_ptr[M]]). */
LDPC535_API int ldpc535_code_create_sparse(const int32_t *row_ptr, const int32_t *col_idx,
                                           int M, int N, int device, ldpc535_code **out);
/* The 32x64 code the reference blocks are hard-wired to. */
LDPC535_API int ldpc535_code_create_default(int device, ldpc535_code **out);
LDPC535_API void ldpc535_code_destroy(ldpc535_code *code);

/* Any out pointer may be NULL. E = number of ones of H. */
LDPC535_API int ldpc535_code_info(const ldpc535_code *code, int *M, int *N, int *K, int *E,
                                  int *device);
/* chosen[M]: pivot column of every re-ordering step (reorderHMatrix's chosenCol). */
LDPC535_API int ldpc535_code_get_pivots(const ldpc535_code *code, int32_t *chosen);
/* Re-ordered H as CSR: row_ptr[M+1], col_idx[E] (columns ascending per row). */
LDPC535_API int ldpc535_code_get_h(const ldpc535_code *code, int32_t *row_ptr, int32_t *col_idx);
/* Generator rows, bit-packed: P[j * words + (k >> 5)] bit (k & 31) = P(j, k),
 * words = (K + 31) / 32;  parity c_j = <P_j, d> mod 2. */
LDPC535_API int ldpc535_code_get_generator(const ldpc535_code *code, uint32_t *P);
/* Which decoder kernel family this code dispatches to: "c4-thread", "warp", "block". */
LDPC535_API const char *ldpc535_code_kernel_name(const ldpc535_code *code, int method);
/* The same for a particular call (early stop and the batch size take part in the dispatch): e.g. sum-product
 * on the shipped code runs on "c4-thread" at fixed iterations and on "warp" with early stop. */
LDPC535_API const char *ldpc535_code_kernel_for(const ldpc535_code *code, int method, int early_stop,
                                                size_t n_win);

/* ---- pinned host memory (optional; pageable pointers work, slower) -------- */
LDPC535_API int ldpc535_host_alloc(size_t bytes, void **ptr);
LDPC535_API int ldpc535_host_free(void *ptr);

/* ---- host-buffer API (what the blocks call): H2D + kernel + D2H inside ---- */
/* in: n_frames * ceil(K/8) bytes; out: n_frames * N complex (2 floats each).
 * Synchronous: `out` is complete on return. */
LDPC535_API int ldpc535_encode_batch(ldpc535_code *code, const uint8_t *in, size_t n_frames,
                                     float *out);

/*
 * Decode n_win windows of the host symbol buffer sym[0 .. n_sym) (complex).
 *   win_offset  symbol offset of each window, or NULL for aligned frames
 *               (window w starts at w * N)
 *   polarity    +1 / -1 per window, or NULL for +1 (the IN_SYNC_INVERTED / -tx
 *               retry of lib/ldpc_decoder_cb_impl.cc:151-152, :178-187)
 *   method      LDPC535_METHOD_*
 *   max_iters   reference: 5
 *   early_stop  reference: 1 (stop at the first iteration whose syndrome is zero)
 *   synd_threshold  the saturating count of checkFrame: out_synd = min(#unsatisfied
 *               checks of the final decision, threshold + 1); reference: M / 8
 *   out_bytes   n_win * ceil(K/8) packed data bytes (never NULL)
 *   out_synd    n_win (may be NULL)
 *   out_iters   n_win (may be NULL): 1-based iteration whose test passed, else
 *               max_iters; 0 for the non-iterative methods
 * Synchronous.
 */
LDPC535_API int ldpc535_decode_batch(ldpc535_code *code, const float *sym, size_t n_sym,
                                     const int64_t *win_offset, const int8_t *polarity,
                                     size_t n_win, int method, int max_iters, int early_stop,
                                     int synd_threshold, uint8_t *out_bytes, uint8_t *out_synd,
                                     uint8_t *out_iters);

/* ---- device-resident API (benchmarks, pipelines that already live in HBM) - */
/* All pointers are device pointers on the handle's device.  `stream` is a
 * cudaStream_t (NULL = the handle's own stream).  Asynchronous: returns after
 * enqueueing; use ldpc535_stream_sync or your own stream/event calls.
 * Alignment (cudaMalloc'ed buffers satisfy all of it; a misaligned pointer is refused with
 * LDPC535_ERR_INVALID instead of faulting on the device): d_sym and the encoder's d_out to
 * 16 bytes, d_out_bytes and the encoder's d_in to 4 bytes, d_win_offset to 8 bytes. */
LDPC535_API int ldpc535_encode_batch_dev(ldpc535_code *code, const uint8_t *d_in,
                                         size_t n_frames, float *d_out, void *stream);
LDPC535_API int ldpc535_decode_batch_dev(ldpc535_code *code, const float *d_sym, size_t n_sym,
                                         const int64_t *d_win_offset, const int8_t *d_polarity,
                                         size_t n_win, int method, int max_iters,
                                         int early_stop, int synd_threshold,
                                         uint8_t *d_out_bytes, uint8_t *d_out_synd,
                                         uint8_t *d_out_iters, void *stream);
LDPC535_API int ldpc535_stream_sync(ldpc535_code *code, void *stream);
LDPC535_API int ldpc535_dev_alloc(ldpc535_code *code, size_t bytes, void **dptr);
LDPC535_API int ldpc535_dev_free(ldpc535_code *code, void *dptr);
LDPC535_API int ldpc535_memcpy_h2d(ldpc535_code *code, void *dptr, const void *hptr, size_t bytes);
LDPC535_API int ldpc535_memcpy_d2h(ldpc535_code *code, void *hptr, const void *dptr, size_t bytes);

/*
 * Message dump for parity tests (sum-product only): decodes n_win aligned windows
 * of the HOST buffer with early_stop as given and writes, per window,
 *   out_L[N]   the last "Test" sums          (lib/ldpc_decoder_cb_impl.cc:519-525)
 *   out_E[E]   check->bit messages           (:503-516)
 *   out_M[E]   bit->check messages           (:540-553)
 * E and M in CSR edge order of the re-ordered H (ldpc535_code_get_h).  Any of the
 * three may be NULL.  out_bytes/out_iters as in ldpc535_decode_batch (may be NULL).
 * kernel: NULL = the code's default dispatch, else "warp" / "block" / "c4-thread".
 */
LDPC535_API int ldpc535_decode_debug(ldpc535_code *code, const float *sym, size_t n_win,
                                     int max_iters, int early_stop, const char *kernel,
                                     float *out_L, float *out_E, float *out_M,
                                     uint8_t *out_bytes, uint8_t *out_iters);

/* Force a kernel family for subsequent decode calls ("warp", "block", "c4-thread", "c4-refill", "regular",
 * NULL/"auto" = default).  "c4-refill": thread-per-codeword kernel with per-codeword early stop and slot
 * refill for the shipped code -- faster than "warp" on clean channels (Eb/N0 >= 6 dB), slower below; it
 * applies to early-stop calls of >= 2 x 256 x SM-count windows and falls back to its siblings otherwise.  Returns LDPC535_ERR_UNSUPPORTED if the code cannot run on it.
 * Measurement knobs read from the environment at ldpc535_code_create* (results never change):
 * LDPC535_REGULAR_VARIANT=0 runs the (3,6)-regular n = 8192 code on the 1024-thread kernel instead
 * of the register-table one; LDPC535_ENCODER=generic keeps large codes on the AND/XOR scan encoder
 * instead of the table look-up one; LDPC535_M4R_RING=40|60|42|62 picks the look-up encoder (free-running
 * with a 4- / 6-slot ring, default 60; barrier kernel with a 4- / 6-slot ring), LDPC535_M4R_TPF=7|8 its
 * frames per slot, LDPC535_M4R_TILE its frames per tile. */
LDPC535_API int ldpc535_code_set_kernel(ldpc535_code *code, const char *kernel);

/* How ldpc535_decode_batch moves host symbols to the device.  Pageable input is always staged
 * through pinned memory by a persistent team of `pack_threads` host threads that copy only the
 * REAL parts (the decoder never reads the imaginary ones), halving the PCIe bytes.  For PINNED
 * input *pack_pinned says: 0 = the copy engine reads the caller's buffer as it is, 1 = the team
 * packs it the same way first, 2 (default) = decided from running measurements of the team's and
 * the copy engine's rates: chunks are packed while the team packs a byte faster than the copy
 * engine moves one (packing costs 12 bytes of host memory traffic per symbol to save 4 on PCIe: it
 * pays with many cores per GPU and not with few).  Default pack_threads = the cores this process
 * may run on (at most 32). */
LDPC535_API int ldpc535_code_host_path(const ldpc535_code *code, int *pack_pinned, int *pack_threads);
/* What the host path has done so far on this handle: 128 MiB chunks packed / handed over as they
 * were, and the bytes copied host -> device for symbols. */
LDPC535_API int ldpc535_code_host_stats(const ldpc535_code *code, uint64_t *chunks_packed, uint64_t *chunks_raw,
                                        uint64_t *h2d_bytes);
/* Override them, e.g. from a launcher that runs one process per GPU and knows how many host cores
 * each process gets: pack_threads >= 1 (0 keeps the current team size); pack_pinned 0 / 1 / 2, or
 * -1 for the default.  The library itself reads no launcher environment. */
LDPC535_API int ldpc535_code_set_host_path(ldpc535_code *code, int pack_pinned, int pack_threads);

/* ---- several GPUs from one host process -------------------------------------- */
/* Codewords are independent, so a batch is cut into contiguous shards, one per device, each
 * decoded / encoded by that device's own handle from its own host thread, every shard writing
 * its slice of the caller's output arrays (a host-side gather; there is no inter-GPU traffic
 * and no collective).  devices[i] may repeat (two handles on one GPU). */
typedef struct ldpc535_pool ldpc535_pool;
LDPC535_API int ldpc535_pool_create(const int32_t *H, int M, int N, const int *devices, int n_devices,
                                    ldpc535_pool **out);            /* H == NULL: the 32x64 default code */
LDPC535_API void ldpc535_pool_destroy(ldpc535_pool *pool);
LDPC535_API int ldpc535_pool_size(const ldpc535_pool *pool);
LDPC535_API ldpc535_code *ldpc535_pool_code(ldpc535_pool *pool, int i);
/* Same arguments and results as ldpc535_decode_batch / ldpc535_encode_batch. */
LDPC535_API int ldpc535_pool_decode_batch(ldpc535_pool *pool, const float *sym, size_t n_sym,
                                          const int64_t *win_offset, const int8_t *polarity,
                                          size_t n_win, int method, int max_iters, int early_stop,
                                          int synd_threshold, uint8_t *out_bytes, uint8_t *out_synd,
                                          uint8_t *out_iters);
LDPC535_API int ldpc535_pool_encode_batch(ldpc535_pool *pool, const uint8_t *in, size_t n_frames,
                                          float *out);

/* ---- measurement utilities (no reference counterpart; used by bench.py and the tests) --------- */
/* Counter-based synthetic inputs: Philox4x32-10 keyed by `seed`, counter = GLOBAL frame index
 * first_frame + f, so any shard of a multi-GPU run holds the frames a one-GPU run has at the same
 * indices.  synth_bytes: n_frames * ceil(K/8) uniform data bytes.  synth_awgn: adds sigma * N(0,1)
 * to the REAL part of each of the n_frames * N symbols in d_sym (the reference simulator's AWGN
 * convention is sigma^2 = 10^(-EbN0_dB/10), apps/ldpc_lapack.cpp:635-642).  Asynchronous. */
LDPC535_API int ldpc535_synth_bytes_dev(ldpc535_code *code, uint64_t seed, uint64_t first_frame,
                                        size_t n_frames, uint8_t *d_bytes, void *stream);
LDPC535_API int ldpc535_synth_awgn_dev(ldpc535_code *code, uint64_t seed, uint64_t first_frame,
                                       size_t n_frames, float sigma, float *d_sym, void *stream);
/* Measured ceiling of an SM pipe on the handle's device, in thread-level operations per second
 * (synchronous, ~50 ms): the denominators of the compute rooflines bench.py reports.
 * LDPC535_PIPE_MUFU: special-function pipe, the sum-product mix of 1 ex2 : 2 lg2;
 * LDPC535_PIPE_FP64: fp64 pipe, min-sum's mix of 2 DADD : 1 DSETP. */
enum { LDPC535_PIPE_MUFU = 0, LDPC535_PIPE_FP64 = 1 };
LDPC535_API int ldpc535_probe_pipe_peak(ldpc535_code *code, int which, double *ops_per_s);

/* Number of kernel launches this handle has issued (bench.py's gpu_launches). */
LDPC535_API uint64_t ldpc535_launch_count(const ldpc535_code *code);

#ifdef __cplusplus
}
#endif
#endif /* LDPC535_H */
