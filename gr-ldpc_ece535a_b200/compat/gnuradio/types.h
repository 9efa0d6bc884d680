// Minimal stand-in for <gnuradio/types.h> / <gnuradio/gr_complex.h> (see block.h).
#pragma once
#include <complex>
#include <vector>
typedef std::complex<float> gr_complex;
typedef std::vector<int> gr_vector_int;
typedef std::vector<unsigned int> gr_vector_uint;
typedef std::vector<const void *> gr_vector_const_void_star;
typedef std::vector<void *> gr_vector_void_star;
