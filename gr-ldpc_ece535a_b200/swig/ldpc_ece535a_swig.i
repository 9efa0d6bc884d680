/* -*- c++ -*- */
/* Python binding of the B200-backed blocks inside a GNU Radio 3.7 install: the same module name
 * and the same two block names the upstream interface file exports
 * (swig/ldpc_ece535a_swig.i:17-22 of the reference), so `import ldpc_ece535a` followed by
 * ldpc_ece535a.ldpc_decoder_cb(method) / ldpc_ece535a.ldpc_encoder_bc() keeps working.
 * Needs GNU Radio's gnuradio.i; it cannot be built in the GNU-Radio-less CI image, where
 * python/ldpc_ece535a/blocks.py exposes the same constructors over lib/block_harness.cc. */

#define LDPC_ECE535A_API

%include "gnuradio.i"

%{
#include "ldpc_ece535a/ldpc_decoder_cb.h"
#include "ldpc_ece535a/ldpc_encoder_bc.h"
#include "ldpc_ece535a/image_sink.h"
%}

%include "ldpc_ece535a/ldpc_decoder_cb.h"
GR_SWIG_BLOCK_MAGIC2(ldpc_ece535a, ldpc_decoder_cb);

%include "ldpc_ece535a/ldpc_encoder_bc.h"
GR_SWIG_BLOCK_MAGIC2(ldpc_ece535a, ldpc_encoder_bc);

%include "ldpc_ece535a/image_sink.h"
GR_SWIG_BLOCK_MAGIC2(ldpc_ece535a, image_sink);
