// Stand-in for <boost/numeric/ublas/matrix_proxy.hpp> (oracle/_ref build only): see ublas_min.hpp.
#pragma once
#include <boost/numeric/ublas/ublas_min.hpp>
