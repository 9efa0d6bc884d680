/* -*- c++ -*- */
#ifndef INCLUDED_LDPC_ECE535A_LDPC_ENCODER_BC_IMPL_H
#define INCLUDED_LDPC_ECE535A_LDPC_ENCODER_BC_IMPL_H

#include <ldpc_ece535a/ldpc_encoder_bc.h>

#include "ldpc535.h"

namespace gr {
namespace ldpc_ece535a {

// Replaces the reference's ldpc_encoder_bc_impl (lib/ldpc_encoder_bc_impl.h:34-67): d_H, d_L,
// d_U and the per-frame LAPACK solves are folded, once, into the generator the code handle
// holds on the GPU.
class ldpc_encoder_bc_impl : public ldpc_encoder_bc
{
private:
    int d_M, d_N, d_nbytes;
    ldpc535_code *d_code;

public:
    ldpc_encoder_bc_impl();
    ~ldpc_encoder_bc_impl();

    void forecast(int noutput_items, gr_vector_int &ninput_items_required);

    int general_work(int noutput_items, gr_vector_int &ninput_items,
                     gr_vector_const_void_star &input_items, gr_vector_void_star &output_items);
};

}  // namespace ldpc_ece535a
}  // namespace gr

#endif /* INCLUDED_LDPC_ECE535A_LDPC_ENCODER_BC_IMPL_H */
