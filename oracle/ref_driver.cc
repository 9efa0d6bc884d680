// ref_driver.cc -- C entry points around the REFERENCE's own LDPC blocks (oracle/_ref build).
//
// TEST INFRASTRUCTURE ONLY.  This file is compiled together with the reference's unmodified
// sources, taken where they lie (never copied into this repository):
//
//     /root/reference/lib/ldpc_decoder_cb_impl.cc   /root/reference/lib/ldpc_encoder_bc_impl.cc
//     + their headers under /root/reference/lib and /root/reference/include
//
// against the dependency stand-ins in oracle/refshim/ (uBLAS containers, gr::block base class,
// LAPACKE_dgesv, boost::shared_ptr / boost::random names) into oracle/_ref/libldpc_ref.so.
// Everything on the path -- reorderHMatrix, makeParityCheck/solve/mod2, the four decode
// methods, checkFrame, both general_work bodies with the sync state machine and the bit
// packing -- is the reference's code; the stand-ins only provide containers, the block base
// class and a textbook dgesv (see refshim/lapacke_dgesv.c for why any dgesv gives the same bits).
//
// Two views are exported:
//   * the scheduler's view: make(), forecast(), general_work() + consume_each() read-back,
//     with whatever the block printed to std::cout (its only diagnostics channel);
//   * the private per-codeword members (decode*, checkFrame, makeParityCheck, reorderHMatrix),
//     reached by compiling THIS translation unit with `private` spelled `public` while the two
//     _impl headers are parsed (the reference's own .cc files are compiled untouched), so that
//     codes other than the pasted-in 32x64 literal and other iteration counts can be checked.
#include <cstdint>
#include <cstring>
#include <sstream>
#include <string>

#include <boost/numeric/ublas/matrix.hpp>
#include <boost/numeric/ublas/vector.hpp>
#include <boost/random.hpp>
#include <boost/shared_ptr.hpp>
#include <gnuradio/block.h>
#include <ldpc_ece535a/api.h>

#define private public
#include "ldpc_decoder_cb_impl.h"
#include "ldpc_encoder_bc_impl.h"
#undef private

using namespace boost::numeric;
using gr::ldpc_ece535a::ldpc_decoder_cb;
using gr::ldpc_ece535a::ldpc_decoder_cb_impl;
using gr::ldpc_ece535a::ldpc_encoder_bc;
using gr::ldpc_ece535a::ldpc_encoder_bc_impl;

namespace {

struct dec_handle {
    ldpc_decoder_cb::sptr blk;
    ldpc_decoder_cb_impl *impl;
};
struct enc_handle {
    ldpc_encoder_bc::sptr blk;
    ldpc_encoder_bc_impl *impl;
};

// std::cout capture (process-wide; the blocks print with std::cout only)
std::ostringstream *g_capture = nullptr;
std::streambuf *g_saved = nullptr;

void fill(ublas::matrix<int> &m, const int *src, int r, int c)
{
    m = ublas::matrix<int>(r, c);
    for (int i = 0; i < r; i++)
        for (int j = 0; j < c; j++) m(i, j) = src[(size_t)i * c + j];
}
void dump(const ublas::matrix<int> &m, int *dst)
{
    if (!dst) return;
    for (size_t i = 0; i < m.size1(); i++)
        for (size_t j = 0; j < m.size2(); j++) dst[i * m.size2() + j] = m(i, j);
}

}  // namespace

extern "C" {

const char *ref_version(void)
{
    return "gr-ldpc_ece535a reference blocks (lib/ldpc_decoder_cb_impl.cc, "
           "lib/ldpc_encoder_bc_impl.cc) + oracle/refshim stand-ins";
}

// ---- std::cout capture ---------------------------------------------------------------------
void ref_cout_begin(void)
{
    if (g_capture) return;
    g_capture = new std::ostringstream();
    g_saved = std::cout.rdbuf(g_capture->rdbuf());
}
// copies what was printed since ref_cout_begin (NUL-terminated, truncated to cap), restores
// std::cout, returns the untruncated length
long ref_cout_end(char *buf, long cap)
{
    if (!g_capture) return 0;
    std::cout.rdbuf(g_saved);
    const std::string s = g_capture->str();
    delete g_capture;
    g_capture = nullptr;
    if (buf && cap > 0) {
        const size_t n = std::min<size_t>(s.size(), (size_t)cap - 1);
        memcpy(buf, s.data(), n);
        buf[n] = 0;
    }
    return (long)s.size();
}

// ---- decoder block ---------------------------------------------------------------------------
void *ref_decoder_new(int method)
{
    dec_handle *h = new dec_handle;
    h->blk = ldpc_decoder_cb::make(method);          // lib/ldpc_decoder_cb_impl.cc:25-30
    h->impl = dynamic_cast<ldpc_decoder_cb_impl *>(h->blk.get());
    return h;
}
void ref_decoder_free(void *p) { delete (dec_handle *)p; }

int ref_decoder_forecast(void *p, int noutput_items)
{
    gr_vector_int req(1, 0);
    ((dec_handle *)p)->blk->forecast(noutput_items, req);
    return req[0];
}

// one scheduler call: in = ninput_items gr_complex (interleaved fp32), out = room for
// noutput_items bytes; returns items produced, *consumed = what consume_each() was told
int ref_decoder_work(void *p, const float *in, int ninput_items, uint8_t *out, int noutput_items,
                     long *consumed)
{
    dec_handle *h = (dec_handle *)p;
    gr_vector_int nin(1, ninput_items);
    gr_vector_const_void_star ins(1, (const void *)in);
    gr_vector_void_star outs(1, (void *)out);
    const int produced = h->blk->general_work(noutput_items, nin, ins, outs);
    if (consumed) *consumed = h->blk->ref_take_consumed();
    return produced;
}

void ref_decoder_get_state(void *p, int *state, unsigned *errors, unsigned *M, unsigned *N,
                           unsigned *iterations)
{
    ldpc_decoder_cb_impl *d = ((dec_handle *)p)->impl;
    if (state) *state = d->d_state;
    if (errors) *errors = d->d_errors;
    if (M) *M = d->d_M;
    if (N) *N = d->d_N;
    if (iterations) *iterations = d->d_iterations;
}
void ref_decoder_set_state(void *p, int state, unsigned errors)
{
    ldpc_decoder_cb_impl *d = ((dec_handle *)p)->impl;
    d->d_state = state;
    d->d_errors = errors;
}
void ref_decoder_get_h(void *p, int *H) { dump(((dec_handle *)p)->impl->d_H, H); }

// Replace the pasted-in 32x64 literal: d_H = H (raw, M x N), then what the constructor does
// with it (lib/ldpc_decoder_cb_impl.cc:98-106): reorderHMatrix(d_H, L, U).  Outputs (any may be
// NULL): the permuted H and the L, U factors (M x (N-M)).  iterations > 0 replaces d_iterations.
void ref_decoder_set_code(void *p, const int *H, int M, int N, int iterations, int *Hp, int *L,
                          int *U)
{
    ldpc_decoder_cb_impl *d = ((dec_handle *)p)->impl;
    d->d_M = M;
    d->d_N = N;
    fill(d->d_H, H, M, N);
    ublas::matrix<int> Lm(ublas::zero_matrix<int>(M, N - M));
    ublas::matrix<int> Um(ublas::zero_matrix<int>(M, N - M));
    d->reorderHMatrix(d->d_H, Lm, Um);
    if (iterations > 0) d->d_iterations = iterations;
    dump(d->d_H, Hp);
    dump(Lm, L);
    dump(Um, U);
}

// one codeword through a private decode member with the block's d_H:
// method 0 decodeLogDomainSimple, 1 decodeSumProductSoft, 2 decodeBitFlipping, 3 decodeHard
void ref_decoder_decode(void *p, int method, const double *rx, int iterations, int *vhat)
{
    ldpc_decoder_cb_impl *d = ((dec_handle *)p)->impl;
    const unsigned N = d->d_N;
    ublas::vector<double> tx(N);
    for (unsigned i = 0; i < N; i++) tx(i) = rx[i];
    ublas::vector<int> v;
    if (method == 3) v = d->decodeHard(tx);
    else if (method == 2) v = d->decodeBitFlipping(tx, d->d_H, iterations);
    else if (method == 1) v = d->decodeSumProductSoft(tx, d->d_H, iterations);
    else v = d->decodeLogDomainSimple(tx, d->d_H, iterations);
    for (unsigned i = 0; i < N; i++) vhat[i] = v(i);
}

// n aligned frames through one private decode member, framed the way general_work frames them
// (lib/ldpc_decoder_cb_impl.cc:149-153 tx from the real parts, :209-219 data bits MSB first):
// the throughput loop of bench.py's reference arm and of the bulk parity tests.  The sync machine
// is not involved.  out_bytes: (N-M)/8 per frame; out_synd (may be NULL): checkFrame(vhat, M/8).
void ref_decoder_decode_frames(void *p, int method, const float *sym, long n, int iterations,
                               uint8_t *out_bytes, uint8_t *out_synd)
{
    ldpc_decoder_cb_impl *d = ((dec_handle *)p)->impl;
    const unsigned N = d->d_N, M = d->d_M, nb = (N - M) / 8;
    const gr_complex *in = (const gr_complex *)sym;
    for (long f = 0; f < n; f++, in += N) {
        ublas::vector<double> tx(N);
        for (unsigned i = 0; i < N; i++) tx(i) = (in + i)->real();
        ublas::vector<int> v;
        if (method == 3) v = d->decodeHard(tx);
        else if (method == 2) v = d->decodeBitFlipping(tx, d->d_H, iterations);
        else if (method == 1) v = d->decodeSumProductSoft(tx, d->d_H, iterations);
        else v = d->decodeLogDomainSimple(tx, d->d_H, iterations);
        for (unsigned i = 0; i < nb; i++) {
            uint8_t b = 0;
            for (unsigned j = 0; j < 8; j++)
                if (v(M + i * 8 + j) == 1) b |= 1 << (7 - j);
            out_bytes[f * nb + i] = b;
        }
        if (out_synd) out_synd[f] = (uint8_t)d->checkFrame(v, M / 8);
    }
}

int ref_decoder_check_frame(void *p, const int *u, int threshold)
{
    ldpc_decoder_cb_impl *d = ((dec_handle *)p)->impl;
    ublas::vector<int> v(d->d_N);
    for (unsigned i = 0; i < d->d_N; i++) v(i) = u[i];
    return d->checkFrame(v, threshold);
}

// ---- encoder block ---------------------------------------------------------------------------
void *ref_encoder_new(void)
{
    enc_handle *h = new enc_handle;
    h->blk = ldpc_encoder_bc::make();                // lib/ldpc_encoder_bc_impl.cc:23-28
    h->impl = dynamic_cast<ldpc_encoder_bc_impl *>(h->blk.get());
    return h;
}
void ref_encoder_free(void *p) { delete (enc_handle *)p; }

int ref_encoder_forecast(void *p, int noutput_items)
{
    gr_vector_int req(1, 0);
    ((enc_handle *)p)->blk->forecast(noutput_items, req);
    return req[0];
}

// in = ninput_items bytes, out = room for noutput_items gr_complex (interleaved fp32)
int ref_encoder_work(void *p, const uint8_t *in, int ninput_items, float *out, int noutput_items,
                     long *consumed)
{
    enc_handle *h = (enc_handle *)p;
    gr_vector_int nin(1, ninput_items);
    gr_vector_const_void_star ins(1, (const void *)in);
    gr_vector_void_star outs(1, (void *)out);
    const int produced = h->blk->general_work(noutput_items, nin, ins, outs);
    if (consumed) *consumed = h->blk->ref_take_consumed();
    return produced;
}

void ref_encoder_get_tables(void *p, int *H, int *L, int *U)
{
    ldpc_encoder_bc_impl *e = ((enc_handle *)p)->impl;
    dump(e->d_H, H);
    dump(e->d_L, L);
    dump(e->d_U, U);
}

// what the constructor does (lib/ldpc_encoder_bc_impl.cc:95-101) for another code
void ref_encoder_set_code(void *p, const int *H, int M, int N, int *Hp, int *L, int *U)
{
    ldpc_encoder_bc_impl *e = ((enc_handle *)p)->impl;
    e->d_M = M;
    e->d_N = N;
    fill(e->d_H, H, M, N);
    e->d_L = ublas::zero_matrix<int>(M, N - M);
    e->d_U = ublas::zero_matrix<int>(M, N - M);
    e->reorderHMatrix(e->d_H, e->d_L, e->d_U);
    dump(e->d_H, Hp);
    dump(e->d_L, L);
    dump(e->d_U, U);
}

// makeParityCheck(d, d_H, d_L, d_U): d = N-M data bits, c = M parity bits
void ref_encoder_make_parity(void *p, const int *d, int *c)
{
    ldpc_encoder_bc_impl *e = ((enc_handle *)p)->impl;
    const unsigned K = e->d_N - e->d_M;
    ublas::vector<int> data(K);
    for (unsigned i = 0; i < K; i++) data(i) = d[i];
    const ublas::vector<int> par = e->makeParityCheck(data, e->d_H, e->d_L, e->d_U);
    for (unsigned i = 0; i < par.size(); i++) c[i] = par(i);
}

}  // extern "C"
