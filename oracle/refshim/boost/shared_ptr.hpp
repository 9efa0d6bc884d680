// Stand-in for <boost/shared_ptr.hpp> (oracle/_ref build only; Boost is not in this image).
#pragma once
#include <memory>
namespace boost {
using std::shared_ptr;
}
