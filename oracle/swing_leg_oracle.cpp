/*
 * swing_leg_oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as solve_mpc_oracle.cpp).
 *
 * CPU restatement (Eigen-free, double precision, no FMA contraction) of the reference's swing-leg controller,
 *   hector_control/src/common/SwingLegController.cpp:46-56    updateSwingLeg (order of the steps)
 *   hector_control/src/common/SwingLegController.cpp:61-70    updateFootPosition
 *   hector_control/ConvexMPC/GaitGenerator.cpp:54-80          Gait::getSwingSubPhase
 *   hector_control/src/common/SwingLegController.cpp:82-93    updateSwingTimes
 *   hector_control/src/common/SwingLegController.cpp:98-128   computeFootPlacement
 *   hector_control/src/common/SwingLegController.cpp:134-155  computeFootDesiredPosition
 *   hector_control/src/common/FootSwingTrajectory.cpp:17-36, include/common/Math/Interpolation.h:53-74
 *   hector_control/src/common/SwingLegController.cpp:160-193  computeIK
 * on the record layouts of include/hector_mpc_b200.h (hmpc_state_t, hmpc_rollout_t, hmpc_swing_t, hmpc_swing_cmd_t).
 *
 * PARITY STATUS: pinned against the reference's own sources.  The reference has no tests or vectors for this
 * controller, and its sources need Eigen (absent here); oracle/Makefile compiles SwingLegController.cpp,
 * FootSwingTrajectory.cpp, GaitGenerator.cpp, LegController.cpp and ConvexMPCLocomotion.cpp UNCHANGED against
 * oracle/eigen_shim into oracle/_ref/libref_tick.so, and tests/test_reference_tick.py requires this restatement to
 * reproduce that controller tick by tick (520 ticks of walking, two updateSwingLeg calls per tick as
 * ConvexMPCLocomotion::run makes them): controller memory, touch-down point, pDes and vDes bit for bit, the IK joint
 * targets within 1e-12 rad.  The arithmetic is restated expression by expression (including the float clamp of the
 * placement offsets, fminf/fmaxf at :117-118, and M_PI in the joint offsets at :190-192).
 */
#include <algorithm>
#include <cmath>
#include <cstring>

#include "../include/hector_mpc_b200.h"

namespace {
const double kHipYaw[2][3] = {{-0.005, -0.057, -0.126}, {-0.005, 0.057, -0.126}};  // Biped.h:11-13, 21-24
const double kHipRoll0[3] = {0.0465, 0.015, -0.0705};                               // Biped.h:14-16 (leg 0)

void quat_to_rbody(const double* q, double* rB)  // orientation_tools.h:182-200 (R, then transposed)
{
  const double e0 = q[0], e1 = q[1], e2 = q[2], e3 = q[3];
  const double R[9] = {1 - 2 * (e2 * e2 + e3 * e3), 2 * (e1 * e2 - e0 * e3), 2 * (e1 * e3 + e0 * e2),
                       2 * (e1 * e2 + e0 * e3), 1 - 2 * (e1 * e1 + e3 * e3), 2 * (e2 * e3 - e0 * e1),
                       2 * (e1 * e3 - e0 * e2), 2 * (e2 * e3 + e0 * e1), 1 - 2 * (e1 * e1 + e2 * e2)};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) rB[i * 3 + j] = R[j * 3 + i];
}
double clampd(double v, double lo, double hi) { return std::max(lo, std::min(v, hi)); }  // SwingLegController.h:83-85
double bezier(double y0, double yf, double x) { return y0 + (x * x * x + 3.0 * (x * x * (1.0 - x))) * (yf - y0); }
}  // namespace

extern "C" void oracle_swing_update(const hmpc_state_t* states, const hmpc_rollout_t* loop, const double* phase, hmpc_swing_t* swing,
                                    int n, int n_iterations, double dt, double dtSwing, hmpc_swing_cmd_t* cmd)
{
  for (int i = 0; i < n; i++) {
    const hmpc_state_t& st = states[i];
    const hmpc_rollout_t& lo = loop[i];
    hmpc_swing_t& sw = swing[i];
    hmpc_swing_cmd_t& out = cmd[i];
    memset(&out, 0, sizeof(out));
    double rB[9];
    quat_to_rbody(st.orientation, rB);
    // updateFootPosition (:61-70): pFoot_w = position + rBody^T (hipYaw + leg.p), z forced to 0
    double pFoot_w[2][3];
    for (int leg = 0; leg < 2; leg++) {
      const double hp[3] = {kHipYaw[leg][0] + st.leg_p[3 * leg], kHipYaw[leg][1] + st.leg_p[3 * leg + 1], kHipYaw[leg][2] + st.leg_p[3 * leg + 2]};
      for (int a = 0; a < 3; a++) pFoot_w[leg][a] = st.position[a] + (rB[0 * 3 + a] * hp[0] + rB[1 * 3 + a] * hp[1] + rB[2 * 3 + a] * hp[2]);
      pFoot_w[leg][2] = 0.0;
    }
    // Gait::getSwingSubPhase (GaitGenerator.cpp:54-80)
    double swingStates[2];
    for (int leg = 0; leg < 2; leg++) {
      const double offsetsPhase = (double)lo.gait_offset[leg] / (double)n_iterations;
      const double durationsPhase = (double)lo.gait_duration[leg] / (double)n_iterations;
      double swing_offset = offsetsPhase + durationsPhase;
      if (swing_offset > 1) swing_offset -= 1.;
      const double swing_duration = 1. - durationsPhase;
      double progress = phase[i] - swing_offset;
      if (progress < 0) progress += 1.;
      if (progress > swing_duration) progress = 0.;
      else progress = progress / swing_duration;
      swingStates[leg] = progress;
    }
    const int g_stance = lo.gait_duration[0], g_swing = n_iterations - lo.gait_duration[0];  // GaitGenerator.cpp:13-14
    // updateSwingTimes (:82-93)
    for (int leg = 0; leg < 2; leg++) {
      if (sw.first_swing[leg]) {
        sw.swing_time[leg] = dtSwing * g_swing;
      } else {
        sw.swing_time[leg] -= dt;
        if (sw.swing_time[leg] <= 0) sw.first_swing[leg] = 1;
      }
    }
    // computeFootPlacement (:98-128)
    const double vdr[3] = {st.state_des[2], st.state_des[3], 0.0};
    double vdw[3];
    for (int a = 0; a < 3; a++) vdw[a] = rB[0 * 3 + a] * vdr[0] + rB[1 * 3 + a] * vdr[1] + rB[2 * 3 + a] * vdr[2];
    double Pf[2][3];
    for (int leg = 0; leg < 2; leg++) {
      for (int a = 0; a < 3; a++)
        Pf[leg][a] = st.position[a] + (rB[0 * 3 + a] * kHipYaw[leg][0] + rB[1 * 3 + a] * kHipYaw[leg][1] + rB[2 * 3 + a] * kHipYaw[leg][2]) +
                     st.vWorld[a] * sw.swing_time[leg];
      const double p_rel_max = 0.3;
      double pfx_rel = 1.75 * st.vWorld[0] * 0.5 * g_stance * dtSwing + 0.1 * (st.vWorld[0] - vdw[0]);
      double pfy_rel = 1.75 * st.vWorld[1] * 0.5 * g_stance * dtSwing + 0.1 * (st.vWorld[1] - vdw[1]);
      pfx_rel = fminf(fmaxf(pfx_rel, -p_rel_max), p_rel_max);  // float functions on doubles, as written (:117-118)
      pfy_rel = fminf(fmaxf(pfy_rel, -p_rel_max), p_rel_max);
      Pf[leg][0] += pfx_rel;
      Pf[leg][1] += pfy_rel;
      Pf[leg][2] = 0.0;
      for (int a = 0; a < 3; a++) out.pf[3 * leg + a] = Pf[leg][a];
    }
    // computeFootDesiredPosition (:134-155) + computeIK (:160-193) for the legs in swing
    for (int leg = 0; leg < 2; leg++) {
      if (!(swingStates[leg] > 0)) continue;
      out.swing[leg] = 1;
      if (sw.first_swing[leg]) {
        sw.first_swing[leg] = 0;
        for (int a = 0; a < 3; a++) sw.p0[3 * leg + a] = pFoot_w[leg][a];
      }
      const double ph = swingStates[leg], height = 0.15;  // setHeight(0.15) at :107
      const double* p0 = sw.p0 + 3 * leg;
      double pDes[3];
      for (int a = 0; a < 2; a++) pDes[a] = bezier(p0[a], Pf[leg][a], ph);
      pDes[2] = (ph < 0.5) ? bezier(p0[2], p0[2] + height, ph * 2) : bezier(p0[2] + height, Pf[leg][2], ph * 2 - 1);
      const double side_w = (leg == 1) ? 1.0 : -1.0;
      const double hipWidthOffSet[3] = {-0.015, side_w * -0.055, 0.0};
      double pb[3], vb[3];
      const double d[3] = {pDes[0] - st.position[0], pDes[1] - st.position[1], pDes[2] - st.position[2]};
      for (int a = 0; a < 3; a++) {
        pb[a] = (rB[a * 3] * d[0] + rB[a * 3 + 1] * d[1] + rB[a * 3 + 2] * d[2]) + hipWidthOffSet[a];
        vb[a] = rB[a * 3] * (0.0 - st.vWorld[0]) + rB[a * 3 + 1] * (0.0 - st.vWorld[1]) + rB[a * 3 + 2] * (0.0 - st.vWorld[2]);
        out.p_des[3 * leg + a] = pb[a];
        out.v_des[3 * leg + a] = vb[a];
      }
      // computeIK
      const double side = (leg == 0) ? -1.0 : 1.0;
      const double hip_roll[3] = {kHipRoll0[0] - 0.06, 0.0, kHipYaw[0][2] + kHipRoll0[2] * 2};
      const double f[3] = {pb[0] - hip_roll[0], pb[1] - hip_roll[1], pb[2] - hip_roll[2]};
      const double distance_3D = std::sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
      const double distance_2D_yOz = std::sqrt(f[1] * f[1] + f[2] * f[2]);
      const double distance_horizontal = 0.0205;
      const double distance_vertical = std::sqrt(std::max(0.00001, distance_2D_yOz * distance_2D_yOz - distance_horizontal * distance_horizontal));
      const double distance_2D_xOz = std::pow(distance_3D * distance_3D - distance_horizontal * distance_horizontal, 0.5);  // pow(.,0.5) as written (:171)
      const double acosArg1 = clampd(distance_2D_xOz / (2.0 * 0.22), -1.0, 1.0);
      const double acosArg2 = clampd(distance_vertical / distance_2D_xOz, -1.0, 1.0);
      double divisor = std::fabs(f[0]);
      divisor = (divisor == 0.0) ? 1e-6 : divisor;
      double* q = out.q_des + 5 * leg;
      q[0] = 0.0;
      q[1] = std::asin(clampd(f[1] / distance_2D_yOz, -1.0, 1.0)) + std::asin(clampd(distance_horizontal * side / distance_2D_yOz, -1.0, 1.0));
      q[2] = std::acos(acosArg1) - std::acos(acosArg2) * (f[0]) / divisor;
      q[3] = 2.0 * std::asin(clampd(distance_2D_xOz / 2.0 / 0.22, -1.0, 1.0)) - M_PI;
      q[4] = -st.leg_q[5 * leg + 3] - st.leg_q[5 * leg + 2];
      q[2] -= 0.3 * M_PI;
      q[3] += 0.6 * M_PI;
      q[4] -= 0.3 * M_PI;
    }
  }
}
