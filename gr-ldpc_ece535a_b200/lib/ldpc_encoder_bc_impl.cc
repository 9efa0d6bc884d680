/* -*- c++ -*- */
/*
 * ldpc_encoder_bc: GNU Radio block, unsigned char in, gr_complex out.
 *
 * Behaviour follows the reference block (lib/ldpc_encoder_bc_impl.cc:33-178): every 4 input
 * bytes (MSB first) are one 32-bit data word; the frame is 32 parity symbols followed by the
 * 32 data symbols, bit 1 -> +1+0j, bit 0 -> -1+0j; a tail shorter than 4 bytes or an output
 * space shorter than 64 items is left for the next call.  All whole frames of a call go to the
 * GPU as ONE batch (ldpc535_encode_batch).
 */
#ifdef HAVE_CONFIG_H
#include "config.h"
#endif

#include "ldpc_encoder_bc_impl.h"

#include <gnuradio/io_signature.h>

#include <algorithm>
#include <cstdlib>
#include <cstdio>
#include <stdexcept>
#include <string>

namespace gr {
namespace ldpc_ece535a {

ldpc_encoder_bc::sptr ldpc_encoder_bc::make()
{
    return gnuradio::get_initial_sptr(new ldpc_encoder_bc_impl());
}

ldpc_encoder_bc_impl::ldpc_encoder_bc_impl()
    : gr::block("ldpc_encoder_bc", gr::io_signature::make(1, 1, sizeof(unsigned char)),
                gr::io_signature::make(1, 1, sizeof(gr_complex))),
      d_M(0), d_N(0), d_nbytes(0), d_code(NULL)
{
    const char *dev = std::getenv("LDPC535_DEVICE");
    const int st = ldpc535_code_create_default(dev ? std::atoi(dev) : 0, &d_code);
    if (st != LDPC535_OK)
        throw std::runtime_error(std::string("ldpc_encoder_bc: cannot set up the GPU encoder: ") +
                                 ldpc535_strerror(st) + " (" + ldpc535_last_error() + ")");
    int K = 0;
    ldpc535_code_info(d_code, &d_M, &d_N, &K, NULL, NULL);
    d_nbytes = K / 8;
    set_output_multiple(d_N);
    set_relative_rate((double)d_N / (double)d_nbytes);
}

ldpc_encoder_bc_impl::~ldpc_encoder_bc_impl() { ldpc535_code_destroy(d_code); }

void ldpc_encoder_bc_impl::forecast(int noutput_items, gr_vector_int &ninput_items_required)
{
    // ceil(noutput_items / 16): one input byte becomes 16 output symbols (reference :111-116)
    const int per_byte = d_N / d_nbytes;
    ninput_items_required[0] = (noutput_items + per_byte - 1) / per_byte;
}

int ldpc_encoder_bc_impl::general_work(int noutput_items, gr_vector_int &ninput_items,
                                       gr_vector_const_void_star &input_items,
                                       gr_vector_void_star &output_items)
{
    const unsigned char *in = (const unsigned char *)input_items[0];
    gr_complex *out = (gr_complex *)output_items[0];

    const long frames = std::min((long)noutput_items / d_N, (long)ninput_items[0] / d_nbytes);
    if (frames > 0) {
        const int st = ldpc535_encode_batch(d_code, in, (size_t)frames, reinterpret_cast<float *>(out));
        if (st != LDPC535_OK) {
            std::fprintf(stderr, "ldpc_encoder_bc: GPU encode failed: %s (%s)\n", ldpc535_strerror(st),
                         ldpc535_last_error());
            return WORK_DONE;                       // no CPU fallback
        }
    }
    consume_each((int)(frames * d_nbytes));
    return (int)(frames * d_N);
}

}  // namespace ldpc_ece535a
}  // namespace gr
