// Stand-in for <boost/shared_ptr.hpp> used ONLY when the module is built without GNU Radio /
// Boost (this image has neither): GNU Radio 3.7's public block API spells its smart pointer
// boost::shared_ptr, so the compat build maps that name onto std::shared_ptr.
#pragma once
#include <memory>
namespace boost {
using std::shared_ptr;
using std::dynamic_pointer_cast;
using std::enable_shared_from_this;
}
