// locomotion_host.h — host-side C++ mirror of the hot path's CALLER.
//
// Keeps the names and call shapes of the reference's locomotion controller for the part that feeds and
// consumes the QP boundary:
//
//   reference class      hector_control/ConvexMPC/ConvexMPCLocomotion.h:38-98
//   ctor                 ConvexMPCLocomotion(double _dt, int _iterations_between_mpc)         (.h:41)
//   run(ControlFSMData&) ConvexMPCLocomotion.cpp:31-268   (gait selection, pFoot, MPC gate, f_ff hand-over)
//   setGaitNum(int)      .h:45
//   updateMPCIfNeeded(int* mpcTable, ControlFSMData&, bool omniMode)   ConvexMPCLocomotion.cpp:273-441
//   Gait                 hector_control/ConvexMPC/GaitGenerator.{h,cpp}
//
// The reference's ControlFSMData drags in Eigen, ROS and the whole controller; only the fields this path
// reads/writes are kept, as plain arrays with the reference's member names.  Swing-leg control, foot
// placement and joint PD gains (ConvexMPCLocomotion.cpp:119-168, 205-266) are out of scope.
// Everything numerical behind update_problem_data() runs on the GPU (libhector_mpc_b200.so); this file is
// the thin double-precision data preparation the reference also does on the host.
#ifndef HECTOR_LOCOMOTION_HOST_H
#define HECTOR_LOCOMOTION_HOST_H

#include "../../include/hector_mpc_b200.h"

// ---- the subset of the reference's data model that the path touches -----------------------------------
struct StateEstimate {        // include/common/StateEstimatorContainer.h:47-59
  double position[3];
  double orientation[4];      // (w,x,y,z)
  double rBody[9];            // world -> body, row-major
  double rpy[3];
  double omegaWorld[3];
  double vWorld[3];
  double vBody[3];
};
struct LegControllerData {    // include/common/LegController.h (q after updateData, p from forward kinematics)
  double q[5];
  double p[3];
};
struct LegControllerCommand { // only the member this path writes (ConvexMPCLocomotion.cpp:259)
  double feedforwardForce[6];
};
struct DesiredStateData {     // include/common/DesiredCommand.h
  double stateDes[12];
};
struct ControlFSMData {       // include/common/ControlFSMData.h, reduced
  StateEstimate* _stateEstimate;      // reference: _stateEstimator->getResult()
  LegControllerData* _legData;        // [2]   reference: _legController->data
  LegControllerCommand* _legCommands; // [2]   reference: _legController->commands
  DesiredStateData* _desiredStateCommand;
};

// Gait::mpc_gait / setIterations / getContactSubPhase / getSwingSubPhase — GaitGenerator.cpp:6-17, 28-82, 85-113
class Gait {
 public:
  Gait(int nMPC_segments, int offset0, int offset1, int duration0, int duration1);
  ~Gait();
  int* mpc_gait();
  void setIterations(int iterationsPerMPC, int currentIteration);
  void getContactSubPhase(double out[2]) const;  // 0 outside stance, else progress through it (continuous _phase)
  void getSwingSubPhase(double out[2]) const;    // 0 outside swing, else progress through it
  int _stance, _swing;

 private:
  int* _mpc_table;
  int _offsets[2], _durations[2];
  double _offsetsPhase[2], _durationsPhase[2];
  int _iteration, _nIterations;
  double _phase;
};

class ConvexMPCLocomotion {
 public:
  ConvexMPCLocomotion(double _dt, int _iterations_between_mpc);
  void run(ControlFSMData& data);
  void setGaitNum(int gaitNum) { gaitNumber = gaitNum; }
  bool firstRun = true;

  // inspection hooks for tests (not in the reference)
  const double* trajectory() const { return trajAll; }
  const double* footForce(int leg) const { return f_ff[leg]; }
  const int* lastGaitTable() const { return lastTable; }
  int iteration() const { return iterationCounter; }

 private:
  void updateMPCIfNeeded(int* mpcTable, ControlFSMData& data, bool omniMode);

  int iterationsBetweenMPC;
  int horizonLength;
  double dt;
  double dtMPC;
  int iterationCounter = 0;
  double f_ff[2][6];
  Gait walking, standing;
  int gaitNumber;
  double world_position_desired[3];
  double pFoot[2][3];
  double trajAll[12 * 10];
  int lastTable[20];
};

// ---- batched caller (row f-1): the same data preparation for B robots, one GPU launch ------------------
// Fills `records[i]` exactly as updateMPCIfNeeded + update_problem_data would for robot i.
extern "C" void hmpc_prepare_record(const StateEstimate* se, const LegControllerData legs[2], const DesiredStateData* cmd,
                                    const double world_position_desired_xy[2], const int* mpcTable, int horizon,
                                    double dtMPC, struct update_data_t* out, double* trajAll_out /* [12*horizon] or NULL */);
// f_ff[leg] = -rBody * [GRF; GRM] from the first-step wrench (ConvexMPCLocomotion.cpp:419-440)
extern "C" void hmpc_wrench_to_feedforward(const double* rBody, const double* wrench12, double f_ff_out[2][6]);

#endif
