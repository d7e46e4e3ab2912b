// locomotion_host.cpp — see locomotion_host.h.  Double-precision host data preparation around the GPU QP,
// following hector_control/ConvexMPC/ConvexMPCLocomotion.cpp:31-62,171-190,273-441 and GaitGenerator.cpp.
#include "locomotion_host.h"

#include <cmath>
#include <cstring>

namespace {
const double kHipYaw[2][3] = {{-0.005, -0.057, -0.126}, {-0.005, 0.057, -0.126}};  // Biped.h:11-13,19-22

// y = M^T x for a row-major 3x3 (rBody.transpose() * v)
inline void mulT(const double* M, const double* x, double* y)
{
  for (int i = 0; i < 3; i++) y[i] = M[0 * 3 + i] * x[0] + M[1 * 3 + i] * x[1] + M[2 * 3 + i] * x[2];
}
inline void mul(const double* M, const double* x, double* y)
{
  for (int i = 0; i < 3; i++) y[i] = M[i * 3 + 0] * x[0] + M[i * 3 + 1] * x[1] + M[i * 3 + 2] * x[2];
}
}  // namespace

// ---------------------------------------------------------------------------------------------------
// Gait (GaitGenerator.cpp:6-17, 85-113)
// ---------------------------------------------------------------------------------------------------
Gait::Gait(int nMPC_segments, int offset0, int offset1, int duration0, int duration1) : _nIterations(nMPC_segments)
{
  _offsets[0] = offset0;
  _offsets[1] = offset1;
  _durations[0] = duration0;
  _durations[1] = duration1;
  for (int i = 0; i < 2; i++) {  // GaitGenerator.cpp:9-10
    _offsetsPhase[i] = (double)_offsets[i] / (double)nMPC_segments;
    _durationsPhase[i] = (double)_durations[i] / (double)nMPC_segments;
  }
  _mpc_table = new int[nMPC_segments * 2];
  _stance = duration0;
  _swing = nMPC_segments - duration0;
  _iteration = 0;
  _phase = 0;
}
Gait::~Gait() { delete[] _mpc_table; }

int* Gait::mpc_gait()
{
  for (int i = 0; i < _nIterations; i++) {
    int iter = (i + _iteration) % _nIterations;
    for (int j = 0; j < 2; j++) {
      int progress = iter - _offsets[j];
      if (progress < 0) progress += _nIterations;
      _mpc_table[i * 2 + j] = (progress < _durations[j]) ? 1 : 0;
    }
  }
  return _mpc_table;
}

void Gait::setIterations(int iterationsPerMPC, int currentIteration)
{
  _iteration = (currentIteration / iterationsPerMPC) % _nIterations;
  _phase = (double)(currentIteration % (iterationsPerMPC * _nIterations)) / (double)(iterationsPerMPC * _nIterations);
}

// GaitGenerator.cpp:28-47: progress through the stance phase from the continuous _phase; exactly 0 outside it
void Gait::getContactSubPhase(double out[2]) const
{
  for (int i = 0; i < 2; i++) {
    double progress = _phase - _offsetsPhase[i];
    if (progress < 0) progress += 1.;
    out[i] = (progress > _durationsPhase[i]) ? 0. : progress / _durationsPhase[i];
  }
}

// GaitGenerator.cpp:53-79: the same for the swing phase, which starts where stance ends (wrapped into [0, 1])
void Gait::getSwingSubPhase(double out[2]) const
{
  for (int i = 0; i < 2; i++) {
    double swing_offset = _offsetsPhase[i] + _durationsPhase[i];
    if (swing_offset > 1) swing_offset -= 1.;
    const double swing_duration = 1. - _durationsPhase[i];
    double progress = _phase - swing_offset;
    if (progress < 0) progress += 1.;
    out[i] = (progress > swing_duration) ? 0. : progress / swing_duration;
  }
}

// ---------------------------------------------------------------------------------------------------
// data preparation shared by the single-robot class and the batched caller
// ---------------------------------------------------------------------------------------------------
extern "C" void hmpc_prepare_record(const StateEstimate* se, const LegControllerData legs[2], const DesiredStateData* cmd,
                                    const double wpd_xy[2], const int* mpcTable, int horizon, double dtMPC,
                                    update_data_t* out, double* trajAll_out)
{
  const double* p = se->position;
  // joint angles: second offset + fmod (ConvexMPCLocomotion.cpp:289-313)
  double q[10];
  for (int i = 0; i < 2; i++)
    for (int k = 0; k < 5; k++) q[i * 5 + k] = legs[i].q[k];
  const double PI = 3.14159265359;
  q[2] += 0.3 * PI; q[3] -= 0.6 * PI; q[4] += 0.3 * PI;
  q[7] += 0.3 * PI; q[8] -= 0.6 * PI; q[9] += 0.3 * PI;
  const double PI2 = 2 * PI;
  for (int i = 0; i < 10; i++) q[i] = fmod(q[i], PI2);
  // foot positions and r (ConvexMPCLocomotion.cpp:58-62, 315-319)
  double pFoot[2][3];
  for (int i = 0; i < 2; i++) {
    double hp[3] = {kHipYaw[i][0] + legs[i].p[0], kHipYaw[i][1] + legs[i].p[1], kHipYaw[i][2] + legs[i].p[2]};
    double w[3];
    mulT(se->rBody, hp, w);
    for (int a = 0; a < 3; a++) pFoot[i][a] = p[a] + w[a];
  }
  double r[6];
  for (int i = 0; i < 6; i++) r[i] = pFoot[i % 2][i / 2] - p[i / 2];
  const double Q[12] = {100, 100, 250, 200, 200, 300, 1, 1, 1, 1, 1, 1};           // :321
  const double Alpha[12] = {1e-4, 1e-4, 5e-4, 1e-4, 1e-4, 5e-4, 1e-2, 1e-2, 1e-2, 1e-2, 1e-2, 1e-2};  // :322
  const double yaw = se->rpy[2];
  // reference trajectory (ConvexMPCLocomotion.cpp:331-399)
  double v_des_robot[3] = {cmd->stateDes[6], cmd->stateDes[7], 0};
  double v_des_world[3];
  mulT(se->rBody, v_des_robot, v_des_world);
  const double max_pos_error = .05;
  double xStart = wpd_xy[0], yStart = wpd_xy[1];
  if (xStart - p[0] > max_pos_error) xStart = p[0] + max_pos_error;
  if (p[0] - xStart > max_pos_error) xStart = p[0] - max_pos_error;
  if (yStart - p[1] > max_pos_error) yStart = p[1] + max_pos_error;
  if (p[1] - yStart > max_pos_error) yStart = p[1] - max_pos_error;
  const double trajInitial[12] = {cmd->stateDes[3], cmd->stateDes[4], 0.0, xStart, yStart, 0.55,
                                  0, 0, cmd->stateDes[11], v_des_world[0], v_des_world[1], 0};
  double trajLocal[12 * K_MAX_GAIT_SEGMENTS];
  double* trajAll = trajAll_out ? trajAll_out : trajLocal;
  for (int i = 0; i < horizon; i++) {
    for (int j = 0; j < 12; j++) trajAll[12 * i + j] = trajInitial[j];
    if (i == 0) {
      trajAll[0] = se->rpy[0]; trajAll[1] = se->rpy[1]; trajAll[2] = se->rpy[2];
      trajAll[3] = p[0]; trajAll[4] = p[1]; trajAll[5] = p[2];
    } else {
      trajAll[12 * i + 3] = (v_des_world[0] == 0 ? trajInitial[3] : p[0]) + i * dtMPC * v_des_world[0];
      trajAll[12 * i + 4] = (v_des_world[1] == 0 ? trajInitial[4] : p[1]) + i * dtMPC * v_des_world[1];
      trajAll[12 * i + 2] = (cmd->stateDes[11] == 0) ? trajInitial[2] : yaw + i * dtMPC * cmd->stateDes[11];
    }
  }
  // double -> float narrowing of update_problem_data (convexMPC_interface.cpp:87-99)
  memset(out, 0, sizeof(*out));
  for (int i = 0; i < 3; i++) { out->p[i] = (float)p[i]; out->v[i] = (float)se->vWorld[i]; out->w[i] = (float)se->omegaWorld[i]; }
  for (int i = 0; i < 4; i++) out->q[i] = (float)se->orientation[i];
  for (int i = 0; i < 6; i++) out->r[i] = (float)r[i];
  for (int i = 0; i < 10; i++) out->joint_angles[i] = (float)q[i];
  out->yaw = (float)yaw;
  for (int i = 0; i < 12; i++) { out->weights[i] = (float)Q[i]; out->Alpha_K[i] = (float)Alpha[i]; }
  for (int i = 0; i < 12 * horizon; i++) out->traj[i] = (float)trajAll[i];
  for (int i = 0; i < 2 * horizon; i++) out->gait[i] = (unsigned char)mpcTable[i];
}

extern "C" void hmpc_wrench_to_feedforward(const double* rBody, const double* w12, double f_ff[2][6])
{
  for (int leg = 0; leg < 2; leg++) {
    double GRF[3], GRM[3], a[3], b[3];
    for (int axis = 0; axis < 3; axis++) {
      GRF[axis] = w12[leg * 3 + axis];
      GRM[axis] = w12[leg * 3 + axis + 6];
    }
    mul(rBody, GRF, a);
    mul(rBody, GRM, b);
    for (int i = 0; i < 3; i++) { f_ff[leg][i] = -a[i]; f_ff[leg][i + 3] = -b[i]; }
  }
}

// ---------------------------------------------------------------------------------------------------
// ConvexMPCLocomotion
// ---------------------------------------------------------------------------------------------------
ConvexMPCLocomotion::ConvexMPCLocomotion(double _dt, int _iterations_between_mpc)
    : iterationsBetweenMPC(_iterations_between_mpc), horizonLength(10), dt(_dt),
      walking(10, 0, 5, 5, 5), standing(10, 0, 0, 10, 10)
{
  gaitNumber = 1;
  dtMPC = dt * iterationsBetweenMPC;
  memset(f_ff, 0, sizeof(f_ff));
  memset(trajAll, 0, sizeof(trajAll));
  memset(lastTable, 0, sizeof(lastTable));
  world_position_desired[0] = world_position_desired[1] = world_position_desired[2] = 0;
}

void ConvexMPCLocomotion::run(ControlFSMData& data)
{
  bool omniMode = false;
  StateEstimate& seResult = *data._stateEstimate;
  Gait* gait = &standing;
  if (gaitNumber == 1) gait = &standing;
  else if (gaitNumber == 2) gait = &walking;
  // integrate position setpoint (ConvexMPCLocomotion.cpp:46-56)
  double v_des_robot[3] = {data._desiredStateCommand->stateDes[6], data._desiredStateCommand->stateDes[7], 0};
  double v_des_world[3];
  mulT(seResult.rBody, v_des_robot, v_des_world);
  world_position_desired[0] += dt * v_des_world[0];
  world_position_desired[1] += dt * v_des_world[1];
  world_position_desired[2] = 0.55;
  if (firstRun) {  // :65-70
    world_position_desired[0] = seResult.position[0];
    world_position_desired[1] = seResult.position[1];
    world_position_desired[2] = seResult.position[2];
    firstRun = false;
  }
  gait->setIterations(iterationsBetweenMPC, iterationCounter);  // :171
  double contactStates[2], swingStates[2];                       // :184-185, this tick's continuous gait phase
  gait->getContactSubPhase(contactStates);
  gait->getSwingSubPhase(swingStates);
  int* mpcTable = gait->mpc_gait();                              // :187
  updateMPCIfNeeded(mpcTable, data, omniMode);                   // :190
  iterationCounter++;
  // :199-266 — a foot whose swing sub-phase is positive belongs to the swing controller; otherwise, if its contact
  // sub-phase is positive, it receives the MPC wrench as feed-forward force; on a tick where both are 0 nothing is written
  for (int foot = 0; foot < 2; foot++) {
    if (swingStates[foot] > 0) continue;
    if (contactStates[foot] > 0)
      for (int i = 0; i < 6; i++) data._legCommands[foot].feedforwardForce[i] = f_ff[foot][i];
  }
}

void ConvexMPCLocomotion::updateMPCIfNeeded(int* mpcTable, ControlFSMData& data, bool omniMode)
{
  (void)omniMode;
  if ((iterationCounter % 5) == 0) {  // hard-coded gate, quirk Q11 (:277)
    StateEstimate& seResult = *data._stateEstimate;
    const double* p = seResult.position;
    update_data_t rec;
    hmpc_prepare_record(&seResult, data._legData, data._desiredStateCommand, world_position_desired, mpcTable,
                        horizonLength, dtMPC, &rec, trajAll);
    // the clamp of :338-346 is written back to the member
    const double max_pos_error = .05;
    if (world_position_desired[0] - p[0] > max_pos_error) world_position_desired[0] = p[0] + max_pos_error;
    if (p[0] - world_position_desired[0] > max_pos_error) world_position_desired[0] = p[0] - max_pos_error;
    if (world_position_desired[1] - p[1] > max_pos_error) world_position_desired[1] = p[1] + max_pos_error;
    if (p[1] - world_position_desired[1] > max_pos_error) world_position_desired[1] = p[1] - max_pos_error;
    for (int i = 0; i < 2 * horizonLength; i++) lastTable[i] = mpcTable[i];

    // the boundary, called exactly as the reference does (:410-415) with double arrays
    double dp[3], dv[3], dw[3], dq[4], dr[6], dj[10], dwt[12], dal[12];
    for (int i = 0; i < 3; i++) { dp[i] = seResult.position[i]; dv[i] = seResult.vWorld[i]; dw[i] = seResult.omegaWorld[i]; }
    for (int i = 0; i < 4; i++) dq[i] = seResult.orientation[i];
    // r / joint angles / weights were prepared in double inside hmpc_prepare_record; recompute the doubles
    // the same way so that update_problem_data performs the (only) double->float narrowing, as in the reference
    {
      double q[10];
      for (int i = 0; i < 2; i++)
        for (int k = 0; k < 5; k++) q[i * 5 + k] = data._legData[i].q[k];
      const double PI = 3.14159265359;
      q[2] += 0.3 * PI; q[3] -= 0.6 * PI; q[4] += 0.3 * PI;
      q[7] += 0.3 * PI; q[8] -= 0.6 * PI; q[9] += 0.3 * PI;
      for (int i = 0; i < 10; i++) dj[i] = fmod(q[i], 2 * PI);
      for (int i = 0; i < 2; i++) {
        double hp[3] = {kHipYaw[i][0] + data._legData[i].p[0], kHipYaw[i][1] + data._legData[i].p[1],
                        kHipYaw[i][2] + data._legData[i].p[2]};
        double w[3];
        mulT(seResult.rBody, hp, w);
        for (int a = 0; a < 3; a++) pFoot[i][a] = p[a] + w[a];
      }
      for (int i = 0; i < 6; i++) dr[i] = pFoot[i % 2][i / 2] - p[i / 2];
    }
    const double Q[12] = {100, 100, 250, 200, 200, 300, 1, 1, 1, 1, 1, 1};
    const double Alpha[12] = {1e-4, 1e-4, 5e-4, 1e-4, 1e-4, 5e-4, 1e-2, 1e-2, 1e-2, 1e-2, 1e-2, 1e-2};
    for (int i = 0; i < 12; i++) { dwt[i] = Q[i]; dal[i] = Alpha[i]; }
    setup_problem(dtMPC, horizonLength, 0.25, 500);
    update_problem_data(dp, dv, dq, dw, dr, dj, seResult.rpy[2], dwt, trajAll, dal, mpcTable);
    double w12[12];
    for (int i = 0; i < 12; i++) w12[i] = get_solution(i);
    hmpc_wrench_to_feedforward(seResult.rBody, w12, f_ff);  // :419-440
  }
}

// ---------------------------------------------------------------------------------------------------
// flat C handles so that tests (ctypes) and C callers can drive the class
// ---------------------------------------------------------------------------------------------------
extern "C" {
void* hloco_create(double dt, int iterations_between_mpc) { return new ConvexMPCLocomotion(dt, iterations_between_mpc); }
void hloco_destroy(void* h) { delete static_cast<ConvexMPCLocomotion*>(h); }
void hloco_set_gait(void* h, int gaitNum) { static_cast<ConvexMPCLocomotion*>(h)->setGaitNum(gaitNum); }
void hloco_run(void* h, StateEstimate* se, LegControllerData* legs, DesiredStateData* cmd, LegControllerCommand* out)
{
  ControlFSMData d;
  d._stateEstimate = se;
  d._legData = legs;
  d._desiredStateCommand = cmd;
  d._legCommands = out;
  static_cast<ConvexMPCLocomotion*>(h)->run(d);
}
// the gait's sub-phases at control iteration `iteration` (what run() gates the feed-forward force with): a test hook
void hloco_gait_subphases(int nseg, const int* offsets, const int* durations, int iterationsPerMPC, int iteration,
                          double contact[2], double swing[2])
{
  Gait g(nseg, offsets[0], offsets[1], durations[0], durations[1]);
  g.setIterations(iterationsPerMPC, iteration);
  g.getContactSubPhase(contact);
  g.getSwingSubPhase(swing);
}
const double* hloco_trajectory(void* h) { return static_cast<ConvexMPCLocomotion*>(h)->trajectory(); }
const double* hloco_foot_force(void* h, int leg) { return static_cast<ConvexMPCLocomotion*>(h)->footForce(leg); }
const int* hloco_gait_table(void* h) { return static_cast<ConvexMPCLocomotion*>(h)->lastGaitTable(); }
int hloco_iteration(void* h) { return static_cast<ConvexMPCLocomotion*>(h)->iteration(); }
}
