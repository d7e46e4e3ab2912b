// oracle/eigen_shim/lcm/lcm-cpp.hpp — TEST INFRASTRUCTURE.
// Declaration-only stand-in for the LCM C++ header: the reference's ControlFSMData.h reaches
// include/sdk/include/unitree_legged_sdk/lcm.h (through interface/CmdPanel.h), which names these three
// types in class declarations.  Nothing of LCM is called on the controller path compiled for the oracle.
#ifndef HMPC_ORACLE_STUB_LCM
#define HMPC_ORACLE_STUB_LCM
#include <string>
namespace lcm {
struct ReceiveBuffer {
  void* data;
  unsigned int data_size;
};
class Subscription {};
class LCM {};
}  // namespace lcm
#endif
