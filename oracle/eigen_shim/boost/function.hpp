// oracle/eigen_shim/boost/function.hpp — TEST INFRASTRUCTURE.  Declaration-only stand-in: the unitree SDK's loop.h
// (reached from the reference's ControlFSMData.h) stores a boost::function<void()> in a class that is never used here.
#ifndef HMPC_ORACLE_STUB_BOOST_FUNCTION
#define HMPC_ORACLE_STUB_BOOST_FUNCTION
#include <functional>
namespace boost {
template <class Sig>
using function = std::function<Sig>;
}
#endif
