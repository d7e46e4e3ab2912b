// oracle/eigen_shim/boost/bind.hpp — TEST INFRASTRUCTURE.  Empty stand-in: unitree_legged_sdk.h includes it, nothing
// on the compiled path uses it.
