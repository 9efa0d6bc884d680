// Stand-in for <boost/random.hpp> (oracle/_ref build only).  The decoder block only declares and
// seeds a generator and a uniform distribution (lib/ldpc_decoder_cb_impl.h:36-37,
// lib/ldpc_decoder_cb_impl.cc:40-42); the draw itself is commented out (:194), so no random
// value ever reaches the data path.
#pragma once
#include <random>
namespace boost { namespace random {
using std::mt19937;
template <class T = int> using uniform_int_distribution = std::uniform_int_distribution<T>;
}}
