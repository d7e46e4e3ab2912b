/*
 * ref_shim_probe.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * C entry points around the reference's OWN translation units
 *   hector_control/ConvexMPC/SolverMPC.cpp, RobotState.cpp, convexMPC_interface.cpp
 * which oracle/Makefile compiles unchanged from /root/reference against oracle/eigen_shim (a stand-in
 * for the absent Eigen, see its header) and the reference's qpOASES -> oracle/_ref/libref_mpc.so.
 * This file only (1) calls `resize_qp_mats` + `solve_mpc` the way `setup_problem` /
 * `update_problem_data` do (convexMPC_interface.cpp:42-66, :83-103), (2) copies the reference's
 * file-scope QP matrices out for comparison with the restatement, (3) silences the reference's
 * per-solve stdout prints (SolverMPC.cpp:639-640, :717) by pointing fd 1 at /dev/null for the call.
 */
#include <fcntl.h>
#include <unistd.h>

#include <cstdio>
#include <iostream>

#include "SolverMPC.h"  // the reference's header (include path: its ConvexMPC directory)

using Eigen::Dynamic;

// file-scope objects of SolverMPC.cpp:26-45, :365-368 (external linkage there)
extern Matrix<fpt, Dynamic, 13> A_qp;
extern Matrix<fpt, Dynamic, Dynamic> B_qp;
extern Matrix<fpt, Dynamic, Dynamic> fmat;
extern Matrix<fpt, Dynamic, Dynamic> qH;
extern Matrix<fpt, Dynamic, 1> qg;
extern Matrix<fpt, Dynamic, 1> U_b;
extern Matrix<fpt, Dynamic, 1> L_b;
extern Matrix<fpt, 13, 1> x_0;
extern Matrix<fpt, 13, 13> A_ct;
extern Matrix<fpt, 13, 12> B_ct_r;

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
}  // namespace

extern "C" {

size_t refshim_sizeof_update_data(void) { return sizeof(update_data_t); }

/* One solve_mpc on a float record.  Optional outputs (row-major, NULL to skip):
 * H[12N*12N], g[12N], A[16N*12N], lb[16N], ub[16N], x0[13], Aqp[13N*13], Act[169], Bct[156]. */
void refshim_solve(const update_data_t* u, const problem_setup* s, double* q_soln, float* H, float* g, float* A,
                   float* lb, float* ub, float* x0, float* Aqp, float* Act, float* Bct)
{
  update_data_t upd = *u;
  problem_setup ps = *s;
  const int N = ps.horizon, nv = 12 * N, nc = 16 * N;
  {
    QuietStdout quiet;
    resize_qp_mats((s16)N);  // setup_problem does this before every solve (ConvexMPCLocomotion.cpp:410)
    solve_mpc(&upd, &ps);
  }
  const mfp* q = get_q_soln();
  for (int i = 0; i < nv; i++) q_soln[i] = q[i];
  if (H) for (int i = 0; i < nv; i++) for (int j = 0; j < nv; j++) H[i * nv + j] = qH(i, j);
  if (g) for (int i = 0; i < nv; i++) g[i] = qg(i);
  if (A) for (int i = 0; i < nc; i++) for (int j = 0; j < nv; j++) A[i * nv + j] = fmat(i, j);
  if (lb) for (int i = 0; i < nc; i++) lb[i] = L_b(i);
  if (ub) for (int i = 0; i < nc; i++) ub[i] = U_b(i);
  if (x0) for (int i = 0; i < 13; i++) x0[i] = x_0(i);
  if (Aqp) for (int i = 0; i < 13 * N; i++) for (int j = 0; j < 13; j++) Aqp[i * 13 + j] = A_qp(i, j);
  if (Act) for (int i = 0; i < 13; i++) for (int j = 0; j < 13; j++) Act[i * 13 + j] = A_ct(i, j);
  if (Bct) for (int i = 0; i < 13; i++) for (int j = 0; j < 12; j++) Bct[i * 12 + j] = B_ct_r(i, j);
}

/* The reference boundary itself, doubles in (convexMPC_interface.h:39-43). */
void refshim_boundary_solve(double dt, int horizon, double mu, double f_max, double* p, double* v, double* q,
                            double* w, double* r, double* joint_angles, double yaw, double* weights,
                            double* state_trajectory, double* Alpha_K, int* gait, double* q_soln)
{
  {
    QuietStdout quiet;
    setup_problem(dt, horizon, mu, f_max);
    update_problem_data(p, v, q, w, r, joint_angles, yaw, weights, state_trajectory, Alpha_K, gait);
  }
  for (int i = 0; i < 12 * horizon; i++) q_soln[i] = get_solution(i);
}

}  // extern "C"
