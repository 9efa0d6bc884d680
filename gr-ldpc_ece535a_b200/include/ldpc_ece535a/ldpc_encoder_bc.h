/* -*- c++ -*- */
/*
 * gr::ldpc_ece535a::ldpc_encoder_bc -- data bytes in, BPSK codeword symbols out.
 *
 * Public surface of the reference block (include/ldpc_ece535a/ldpc_encoder_bc.h:22-36): same
 * namespace, class name, base class, sptr typedef and argument-less make() factory.  The
 * parity computation runs as a bit-packed GF(2) product on the GPU (include/ldpc535.h).
 */
#ifndef INCLUDED_LDPC_ECE535A_LDPC_ENCODER_BC_H
#define INCLUDED_LDPC_ECE535A_LDPC_ENCODER_BC_H

#include <ldpc_ece535a/api.h>
#include <gnuradio/block.h>

namespace gr {
namespace ldpc_ece535a {

class LDPC_ECE535A_API ldpc_encoder_bc : virtual public gr::block
{
public:
    typedef boost::shared_ptr<ldpc_encoder_bc> sptr;

    static sptr make();
};

}  // namespace ldpc_ece535a
}  // namespace gr

#endif /* INCLUDED_LDPC_ECE535A_LDPC_ENCODER_BC_H */
