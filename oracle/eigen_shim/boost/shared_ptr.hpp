// oracle/eigen_shim/boost/shared_ptr.hpp — TEST INFRASTRUCTURE.  Declaration-only stand-in (see function.hpp).
#ifndef HMPC_ORACLE_STUB_BOOST_SHARED_PTR
#define HMPC_ORACLE_STUB_BOOST_SHARED_PTR
#include <memory>
namespace boost {
template <class T>
using shared_ptr = std::shared_ptr<T>;
}
#endif
