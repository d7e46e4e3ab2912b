/*
 * ref_tick_probe.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * One control tick of the reference's walking controller, driven from C, for pinning the restatements of the
 * widening rows of SURVEY.md §8(f).  oracle/Makefile compiles the reference's own translation units
 *   ConvexMPC/{ConvexMPCLocomotion,GaitGenerator,SolverMPC,RobotState,convexMPC_interface}.cpp
 *   src/common/{LegController,SwingLegController,FootSwingTrajectory,DesiredCommand}.cpp
 * UNCHANGED from /root/reference against oracle/eigen_shim (stand-ins for Eigen and, declaration-only, for the
 * LCM/boost headers the unitree SDK headers name) and the reference's qpOASES -> oracle/_ref/libref_tick.so.
 *
 * Drop-in build (-DREFTICK_DROP_IN -> oracle/_ref/libref_tick_b200.so): the same controller objects, but the three MPC
 * files and qpOASES are LEFT OUT and the link is completed by hector_simulation_b200/libhector_mpc_b200.so — the
 * maintainer's change of INTEGRATION.md carried out on the reference's own objects.  Used by the -m gpu suite to tick the
 * reference's controller with the GPU library behind its unchanged setup_problem/update_problem_data/get_solution calls.
 *
 * A tick here is what FSMState_Walking::run does (src/FSM/FSMState_Walking.cpp:25-40):
 *   LegController::updateData -> [state estimate] -> DesiredStateCommand::setStateCommands ->
 *   ConvexMPCLocomotion::run (gait, swing-leg controller, updateMPCIfNeeded -> setup_problem/update_problem_data)
 *   -> LegController::updateCommand.
 * The state estimate is written the way CheaterEstimator does it (include/common/CheaterEstimator.h:10-40: rBody and
 * rpy from the quaternion with ori::quaternionToRotationMatrix / ori::quatToRPY) from the pose the caller passes in.
 *
 * The controller's working memory the restatements must reproduce is private in the reference's classes
 * (world_position_desired, swingTimes, firstSwing, the swing trajectories' end points); this file reads it by
 * including the reference's HEADERS with `private` spelled `public` — the reference's own translation units are
 * compiled untouched, and access specifiers do not change the object layout under the Itanium C++ ABI.
 */
#include <fcntl.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include <eigen3/Eigen/Dense>

#define private public
#define protected public
#include "../include/common/Math/orientation_tools.h"
#include "ConvexMPCLocomotion.h"
#include "convexMPC_interface.h"
#undef private
#undef protected

// file-scope objects of convexMPC_interface.cpp:12-17 (external linkage there).  Not available in the drop-in build
// (REFTICK_DROP_IN: the controller objects linked against libhector_mpc_b200.so instead of the reference's MPC files).
#ifndef REFTICK_DROP_IN
extern update_data_t update;
extern problem_setup problem_configuration;
#endif

namespace {
struct QuietStdout {
  int saved;
  QuietStdout()
  {
    fflush(stdout);
    std::cout.flush();
    saved = dup(1);
    int nul = open("/dev/null", O_WRONLY);
    dup2(nul, 1);
    close(nul);
  }
  ~QuietStdout()
  {
    fflush(stdout);
    std::cout.flush();
    dup2(saved, 1);
    close(saved);
  }
};

struct NullInterface : IOInterface {
  void sendRecv(const LowlevelCmd*, LowlevelState*) override {}
};

struct Tick {
  Biped biped;
  LowlevelState lowState;
  LowlevelCmd lowCmd;
  LegController legs;
  StateEstimate est;
  StateEstimatorContainer estimators;
  DesiredStateCommand desired;
  NullInterface io;
  ControlFSMData data;
  ConvexMPCLocomotion cmpc;
  Tick(double dt, int iterations_between_mpc)
      : legs(biped), estimators(&lowState, legs.data, &est), desired(&est, dt), cmpc(dt, iterations_between_mpc)
  {
    data._biped = &biped;
    data._stateEstimator = &estimators;
    data._legController = &legs;
    data._desiredStateCommand = &desired;
    data._interface = &io;
    data._lowCmd = &lowCmd;
    data._lowState = &lowState;
  }
};
}  // namespace

extern "C" {

/* everything the tests compare, plain doubles/ints (numpy mirror: oracle/oracle_py.py REFTICK_DTYPE) */
struct reftick_out_t {
  /* after updateData + the state estimate (inputs of the restatements) */
  double rBody[9];          /* row-major */
  double rpy[3];
  double leg_q[10];         /* _legController->data[leg].q, first joint offset applied (LegController.cpp:111-113) */
  double leg_p[6];          /* data[leg].p */
  double J[60];             /* data[leg].J_force_moment, [leg][6][5] row-major */
  double wpd_before[3];     /* world_position_desired before ConvexMPCLocomotion::run */
  double wpd_entry[3];      /* ... as updateMPCIfNeeded will find it: integrated by run() (ConvexMPCLocomotion.cpp:63-65),
                               or reset to the position on the first run (:73-75); computed here with the same two lines */
  /* after ConvexMPCLocomotion::run */
  double wpd[3];
  double phase;             /* Gait::_phase */
  int gait_iteration;       /* Gait::_iteration */
  int mpc_table[20];
  int mpc_ran;              /* (iterationCounter % 5) == 0 on entry */
  int iteration_counter;    /* after the tick */
  double q_soln[120];       /* get_solution(i); stale values when mpc_ran == 0, like the reference */
  double f_ff[12];          /* cmpc.f_ff[leg] */
  double swing_states[2];   /* swing.swingStates */
  double swing_times[2];    /* swing.swingTimes */
  int first_swing[2];       /* swing.firstSwing */
  double p0[6], pf[6];      /* swing.footSwingTrajectory[leg]._p0 / _pf */
  double q_des[10], p_des[6], v_des[6];   /* commands[leg].qDes / pDes / vDes after run */
  double ff_cmd[12];        /* commands[leg].feedforwardForce after run */
  double cmpc_pf[6];        /* cmpc.footSwingTrajectories[leg]._pf: run()'s own touch-down heuristic (ConvexMPCLocomotion.cpp:119-160) */
  /* after LegController::updateCommand */
  double tau[10];           /* lowCmd.motorCmd[].tau */
  unsigned char update_record[sizeof(update_data_t)]; /* the global `update` record (valid when mpc_ran) */
};

size_t reftick_sizeof_out(void) { return sizeof(reftick_out_t); }

void* reftick_create(double dt, int iterations_between_mpc)
{
  QuietStdout quiet;
  // the reference's constructor opens "foot_pos.txt" in the working directory (ConvexMPCLocomotion.cpp:28):
  // let that land in /tmp, not in the repository
  char cwd[4096];
  const bool have_cwd = getcwd(cwd, sizeof(cwd)) != NULL;
  if (chdir("/tmp") != 0) return NULL;
  Tick* t = new Tick(dt, iterations_between_mpc);
  if (have_cwd && chdir(cwd) != 0) { delete t; return NULL; }
  return t;
}
void reftick_destroy(void* h) { delete static_cast<Tick*>(h); }

/* gait_number: 1 standing, 2 walking (ConvexMPCLocomotion.cpp:52-55).  motor_q/motor_dq: the ten raw joint readings
 * as LowlevelState holds them (floats).  v_des_body[2], yaw_rate, roll, pitch -> setStateCommands. */
void reftick_run(void* h, int gait_number, const double* position, const double* vWorld, const double* orientation,
                 const double* omegaWorld, const float* motor_q, const float* motor_dq, const double* v_des_body,
                 double yaw_rate, double roll, double pitch, reftick_out_t* out)
{
  Tick& t = *static_cast<Tick*>(h);
  memset(out, 0, sizeof(*out));
  QuietStdout quiet;
  for (int i = 0; i < 10; i++) {
    t.lowState.motorState[i].q = motor_q[i];
    t.lowState.motorState[i].dq = motor_dq[i];
  }
  t.legs.updateData(&t.lowState);
  // the state estimate, written as CheaterEstimator::run does
  for (int i = 0; i < 4; i++) t.est.orientation[i] = orientation[i];
  t.est.rBody = ori::quaternionToRotationMatrix(t.est.orientation);
  for (int i = 0; i < 3; i++) {
    t.est.omegaWorld[i] = omegaWorld[i];
    t.est.vWorld[i] = vWorld[i];
    t.est.position[i] = position[i];
  }
  t.est.omegaBody = t.est.rBody * t.est.omegaWorld;
  t.est.rpy = ori::quatToRPY(t.est.orientation);
  t.est.vBody = t.est.rBody * t.est.vWorld;
  Vec3<double> vdb(v_des_body[0], v_des_body[1], 0.0);
  t.desired.setStateCommands(roll, pitch, vdb, yaw_rate);

  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) out->rBody[i * 3 + j] = t.est.rBody(i, j);
  for (int i = 0; i < 3; i++) out->rpy[i] = t.est.rpy[i];
  for (int leg = 0; leg < 2; leg++) {
    for (int k = 0; k < 5; k++) out->leg_q[5 * leg + k] = t.legs.data[leg].q(k);
    for (int k = 0; k < 3; k++) out->leg_p[3 * leg + k] = t.legs.data[leg].p(k);
    for (int r = 0; r < 6; r++)
      for (int c = 0; c < 5; c++) out->J[leg * 30 + r * 5 + c] = t.legs.data[leg].J_force_moment(r, c);
  }
  for (int i = 0; i < 3; i++) out->wpd_before[i] = t.cmpc.world_position_desired[i];
  out->mpc_ran = (t.cmpc.iterationCounter % 5) == 0;
  {
    Vec3<double> v_des_robot(t.desired.data.stateDes[6], t.desired.data.stateDes[7], 0);
    Vec3<double> v_des_world = t.est.rBody.transpose() * v_des_robot;
    double w[3] = {t.cmpc.world_position_desired[0], t.cmpc.world_position_desired[1], t.cmpc.world_position_desired[2]};
    w[0] += t.cmpc.dt * v_des_world[0];
    w[1] += t.cmpc.dt * v_des_world[1];
    w[2] = 0.55;
    if (t.cmpc.firstRun)
      for (int i = 0; i < 3; i++) w[i] = t.est.position[i];
    for (int i = 0; i < 3; i++) out->wpd_entry[i] = w[i];
  }

  t.cmpc.setGaitNum(gait_number);
  t.cmpc.run(t.data);

  Gait* gait = (gait_number == 2) ? &t.cmpc.walking : &t.cmpc.standing;
  for (int i = 0; i < 3; i++) out->wpd[i] = t.cmpc.world_position_desired[i];
  out->phase = gait->_phase;
  out->gait_iteration = gait->_iteration;
  int* table = gait->mpc_gait();
  for (int i = 0; i < 20; i++) out->mpc_table[i] = table[i];
  out->iteration_counter = t.cmpc.iterationCounter;
  for (int i = 0; i < 120; i++) out->q_soln[i] = get_solution(i);
  for (int leg = 0; leg < 2; leg++) {
    for (int k = 0; k < 6; k++) {
      out->f_ff[6 * leg + k] = t.cmpc.f_ff[leg](k);
      out->ff_cmd[6 * leg + k] = t.legs.commands[leg].feedforwardForce(k);
    }
    out->swing_states[leg] = t.cmpc.swing.swingStates[leg];
    out->swing_times[leg] = t.cmpc.swing.swingTimes[leg];
    out->first_swing[leg] = t.cmpc.swing.firstSwing[leg] ? 1 : 0;
    for (int k = 0; k < 3; k++) {
      out->p0[3 * leg + k] = t.cmpc.swing.footSwingTrajectory[leg]._p0[k];
      out->pf[3 * leg + k] = t.cmpc.swing.footSwingTrajectory[leg]._pf[k];
      out->cmpc_pf[3 * leg + k] = t.cmpc.footSwingTrajectories[leg]._pf[k];
      out->p_des[3 * leg + k] = t.legs.commands[leg].pDes(k);
      out->v_des[3 * leg + k] = t.legs.commands[leg].vDes(k);
    }
    for (int k = 0; k < 5; k++) out->q_des[5 * leg + k] = t.legs.commands[leg].qDes(k);
  }
#ifndef REFTICK_DROP_IN
  memcpy(out->update_record, &update, sizeof(update_data_t));
#endif

  t.legs.updateCommand(&t.lowCmd);
  for (int i = 0; i < 10; i++) out->tau[i] = t.lowCmd.motorCmd[i].tau;
}

}  // extern "C"
