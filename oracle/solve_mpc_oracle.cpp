/*
 * solve_mpc_oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (Eigen-free) of the reference's per-tick convex-MPC QP formulation,
 *   hector_control/ConvexMPC/SolverMPC.cpp:371-732   (solve_mpc)
 *   hector_control/ConvexMPC/SolverMPC.cpp:65-89     (euler_to_rotation)
 *   hector_control/ConvexMPC/SolverMPC.cpp:133-193   (c2qp)
 *   hector_control/ConvexMPC/SolverMPC.cpp:302-342   (cross_mat, ct_ss_mats, quat_to_rpy)
 *   hector_control/ConvexMPC/RobotState.cpp:9-53     (RobotState::set)
 * handing the reduced QP to the reference's own vendored qpOASES 3.2 (compiled unchanged from
 * /root/reference by oracle/Makefile into oracle/_ref/) exactly as SolverMPC.cpp:702-712 does.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this.  The product library (libhector_mpc_b200.so) never links or calls it.
 *
 * PARITY STATUS: pinned against the reference's own source text; bit-level arithmetic of the absent
 * third-party library (Eigen, version not pinned by the reference) restated.  The reference ships no
 * tests or golden vectors and its formulation needs Eigen, which is neither vendored
 * nor installed here.  Two things stand in for that:
 *   1. oracle/_ref/libref_mpc.so — SolverMPC.cpp, RobotState.cpp and convexMPC_interface.cpp compiled
 *      UNCHANGED from /root/reference against oracle/eigen_shim (an eager stand-in for the Eigen API
 *      subset those files use; sequential-sum products, cofactor 3x3 inverse) and the reference's
 *      qpOASES.  This restatement, in TRIG-AS-COMPILED mode, reproduces it bit for bit — qH, qg, fmat,
 *      L_b, U_b, x_0, A_qp and the returned solution (tests/test_reference_compiled.py) — so every
 *      index, sign, literal and quirk below is checked against the reference's text by the compiler,
 *      not by reading.  Committed vectors: tests/golden/ref_compiled_h10.npz.
 *   2. What remains a restatement is Eigen's own arithmetic: `float` wherever the reference uses
 *      `fpt`, products as plain sequential sums with separately rounded multiply and add (the
 *      reference is built with -O3 and no -march, i.e. SSE2 without FMA, hector_control/CMakeLists.txt:7;
 *      Eigen's coefficient-based products accumulate sequentially in k, its blocked GEBP kernel may
 *      group differently — a freedom the reference itself has, its Eigen version being unpinned),
 *      3x3 inverses by the cofactor formula Eigen uses.  The QP-solve half IS the reference (qpOASES
 *      built from its own sources).  Measured sensitivity to that freedom (mode bit 2, the two gemv's of
 *      :570 grouped four columns at a time like Eigen 3.3's kernel): first-step wrench median 1e-6,
 *      worst 3.7e-5 relative on 256 robots of configs[2] — inside the contract.
 *
 * TRIG RESOLUTION (found by compiling the reference, not visible from reading it): SolverMPC.cpp
 * includes <cmath> (:6) and then qpOASES.hpp (:8), whose Utils.ipp:36 includes <math.h>.  With
 * libstdc++ (GCC >= 6) <math.h> brings std::sin/cos/asin(float) into the global namespace, so the
 * unqualified calls on `fpt` arguments in euler_to_rotation (:73-85), quat_to_rpy (:340) and the
 * foot-rotation literals (:428-433) resolve to libm's FLOAT functions, and products of two such
 * values are rounded to float (terms led by the literal `1.0*` stay double); atan2(float, double)
 * (:339,:341) and fmod(float, double) (:392) promote to double.  glibc's sinf/cosf/asinf differ from
 * the correctly rounded value in 1.3 % / 1.3 % / 7 % of arguments and (sinf, cosf) come in CPU-dispatched FMA variants,
 * so those last bits are not reproducible across machines — let alone on a GPU.  Hence two modes:
 *   mode 0, CANONICAL (default; what the CUDA kernel reproduces bit for bit and what the golden
 *           cfg*.npz fixtures hold): every trig call in double, narrowed to float — the correctly
 *           rounded version of the same expression;
 *   mode 2, TRIG AS COMPILED: the overloads the reference's TU gets, for the bit-for-bit check
 *           against libref_mpc.so.
 * The two differ by last-bit perturbations of Rb and R_foot; measured on 256 robots of configs[2]:
 * first-step wrench median 4e-9, worst 3.5e-5 relative (an fp64 referee on the canonical QP sits
 * 2.3e-5 from the compiled reference at worst) — inside the 1e-4 contract, and the scale of the
 * reference's own sensitivity to its libm.
 *
 * Build with -ffp-contract=off and WITHOUT -march=native so no FMA contraction can occur.
 *
 * Extension: c2qp's loops are hard-coded to 10 (SolverMPC.cpp:148,161,180);
 * here they run to `horizon`, which is identical at horizon == 10 and is the documented 10->N
 * generalisation for the horizon-sweep configs.
 */
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <vector>

#include "../include/hector_mpc_b200.h"

#ifdef ORACLE_WITH_QPOASES
#include <qpOASES.hpp>
#endif

namespace {

const double kBigNumber = 5e10;  // SolverMPC.cpp:16

// ---- tiny dense helpers, sequential accumulation, no FMA (file is built -ffp-contract=off) ----

// C(rxc) = A(rxk) * B(kxc), row-major.  Every coefficient is the sum over the inner index in increasing order
// starting from the first product (Eigen: res = a0*b0; res += a1*b1; ...).  The loops run row of C by row of B so that
// the compiler can keep a whole row of accumulators in SSE registers — the order of additions per coefficient, hence
// every bit of the result, is that of the textbook triple loop.
template <class T>
void matmul(const T* A, const T* B, T* C, int r, int k, int c)
{
  for (int i = 0; i < r; i++) {
    T* Ci = C + (size_t)i * c;
    const T a0 = A[(size_t)i * k];
    for (int j = 0; j < c; j++) Ci[j] = a0 * B[j];
    for (int t = 1; t < k; t++) {
      const T at = A[(size_t)i * k + t];
      const T* Bt = B + (size_t)t * c;
      for (int j = 0; j < c; j++) Ci[j] = Ci[j] + at * Bt[j];
    }
  }
}

// The same product with the k-sum cut into depth blocks the way a blocked GEBP kernel accumulates — every block summed from
// zero in a register, then added to the result, C = ((0 + s_0) + s_1) + ... — sensitivity probe of the freedom Eigen's
// cache blocking has (a single block, kc >= k, is the sequential sum above).
template <typename T>
void matmul_kblocked(const T* A, const T* B, T* C, int r, int k, int c, int kc)
{
  for (int i = 0; i < r; i++)
    for (int j = 0; j < c; j++) {
      T res = (T)0;
      for (int k0 = 0; k0 < k; k0 += kc) {
        const int k1 = k0 + kc < k ? k0 + kc : k;
        T acc = A[(size_t)i * k + k0] * B[(size_t)k0 * c + j];
        for (int t = k0 + 1; t < k1; t++) acc = acc + A[(size_t)i * k + t] * B[(size_t)t * c + j];
        res = res + acc;
      }
      C[(size_t)i * c + j] = res;
    }
}

// Eigen's 3x3 inverse (Eigen/src/LU/InverseImpl.h, compute_inverse<...,3>): cofactors, det from
// the first cofactor column dotted with the first matrix column, multiply by 1/det.
template <class T>
T cofactor3(const T* m, int i, int j)
{
  int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
  return m[i1 * 3 + j1] * m[i2 * 3 + j2] - m[i1 * 3 + j2] * m[i2 * 3 + j1];
}
template <class T>
void inverse3(const T* m, T* inv)
{
  T c00 = cofactor3(m, 0, 0), c10 = cofactor3(m, 1, 0), c20 = cofactor3(m, 2, 0);
  T det = (c00 * m[0] + c10 * m[3]) + c20 * m[6];
  T invdet = T(1) / det;
  // inverse(j,i) = cofactor(i,j) / det
  inv[0] = c00 * invdet;
  inv[1] = c10 * invdet;
  inv[2] = c20 * invdet;
  inv[3] = cofactor3(m, 0, 1) * invdet;
  inv[4] = cofactor3(m, 1, 1) * invdet;
  inv[5] = cofactor3(m, 2, 1) * invdet;
  inv[6] = cofactor3(m, 0, 2) * invdet;
  inv[7] = cofactor3(m, 1, 2) * invdet;
  inv[8] = cofactor3(m, 2, 2) * invdet;
}

// Quaternionf::toRotationMatrix() (Eigen/src/Geometry/Quaternion.h); q = (w,x,y,z).
template <class T>
void quat_to_R(const T* q, T* R)
{
  T w = q[0], x = q[1], y = q[2], z = q[3];
  T tx = T(2) * x, ty = T(2) * y, tz = T(2) * z;
  T twx = tx * w, twy = ty * w, twz = tz * w;
  T txx = tx * x, txy = ty * x, txz = tz * x;
  T tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = T(1) - (tyy + tzz);
  R[1] = txy - twz;
  R[2] = txz + twy;
  R[3] = txy + twz;
  R[4] = T(1) - (txx + tzz);
  R[5] = tyz - twx;
  R[6] = txz - twy;
  R[7] = tyz + twx;
  R[8] = T(1) - (txx + tyy);
}

// Foot rotation of one leg from its five (offset-corrected) joint angles: SolverMPC.cpp:428-433.
// Same expression tree as the reference, written with named sub-terms instead of one literal.
// TR is the type the trig calls resolve to (see "trig resolution" in the header comment): with
// TR = double every sub-term is double; with TR = float the C++ usual arithmetic conversions give
// exactly the mixed float/double evaluation of the reference's literal (a product of two float
// trig values is rounded to float; a term that starts with the literal `1.0*` is double).
template <class T, class TR>
void foot_rotation(const T* q, T* Rf)
{
  const TR s0 = std::sin((TR)q[0]), c0 = std::cos((TR)q[0]);
  const TR s1 = std::sin((TR)q[1]), c1 = std::cos((TR)q[1]);
  const TR s2 = std::sin((TR)q[2]), c2 = std::cos((TR)q[2]);
  const TR s3 = std::sin((TR)q[3]), c3 = std::cos((TR)q[3]);
  const TR s4 = std::sin((TR)q[4]), c4 = std::cos((TR)q[4]);
  // recurring brackets of the reference expression
  const auto a = c0 * s2 + c2 * s0 * s1;        // (cos(q0)*sin(q2) + cos(q2)*sin(q0)*sin(q1))
  const auto b = c0 * c2 - 1.0 * s0 * s1 * s2;  // (cos(q0)*cos(q2) - 1.0*sin(q0)*sin(q1)*sin(q2))
  const auto c = c2 * s0 + c0 * s1 * s2;        // (cos(q2)*sin(q0) + cos(q0)*sin(q1)*sin(q2))
  const auto d = s0 * s2 - 1.0 * c0 * c2 * s1;  // (sin(q0)*sin(q2) - 1.0*cos(q0)*cos(q2)*sin(q1))
  // the sum q2+q3+q4 is formed in the matrix scalar type (floats in the reference)
  const T q234 = q[2] + q[3] + q[4];
  Rf[0] = (T)(-1.0 * s4 * (c3 * a + s3 * b) - c4 * (1.0 * s3 * a - c3 * b));
  Rf[1] = (T)(-1.0 * c1 * s0);
  Rf[2] = (T)(c4 * (c3 * a + s3 * b) - s4 * (1.0 * s3 * a - c3 * b));
  Rf[3] = (T)(c4 * (c3 * c - 1.0 * s3 * d) - 1.0 * s4 * (s3 * c + c3 * d));
  Rf[4] = (T)(c0 * c1);
  Rf[5] = (T)(c4 * (s3 * c + c3 * d) + s4 * (c3 * c - 1.0 * s3 * d));
  Rf[6] = (T)(-1.0 * std::sin((TR)q234) * c1);
  Rf[7] = (T)(s1);
  Rf[8] = (T)(std::cos((TR)q234) * c1);
}

// Pitch from its clamped sine (SolverMPC.cpp:340) in the trig type TR.
template <class T, class TR>
T pitch_from_sine(T as)
{
  return (T)std::asin((TR)as);
}
// The Euler-rate map whose inverse is Rb (SolverMPC.cpp:83-85) in the trig type TR: with TR = float,
// `cos(y)*cos(p)` is a float product of two float cosines.
template <class T, class TR>
void euler_rate_map(T pitch, T yaw, T* Rbm)
{
  const TR p = (TR)pitch, y = (TR)yaw;
  Rbm[0] = (T)(std::cos(y) * std::cos(p)); Rbm[1] = (T)(-std::sin(y)); Rbm[2] = (T)0;
  Rbm[3] = (T)(std::sin(y) * std::cos(p)); Rbm[4] = (T)(std::cos(y));  Rbm[5] = (T)0;
  Rbm[6] = (T)(-std::sin(p));              Rbm[7] = (T)0;              Rbm[8] = (T)1;
}

}  // namespace

// Everything the formulation produces, for parity checks of intermediate stages.
template <class T>
struct Formulation {
  int N = 0;
  T q[10];
  T R[9];
  T rpy[3];
  T Rb[9];
  T x0[13];
  T I_world[9];
  T A_ct[169];
  T B_ct[156];
  T Rfoot[2][9];
  T Acd[169];
  T Bcd[156];
  std::vector<T> A_qp;  // 13N x 13
  std::vector<T> B_qp;  // 13N x 12N
  T Fblk[192];          // F_control 16 x 12
  std::vector<T> lb, ub;  // 16N
  std::vector<T> H;       // 12N x 12N
  std::vector<T> g;       // 12N
};

template <class T>
static void formulate(const update_data_t* u, const problem_setup* setup, Formulation<T>& F, bool trig_as_compiled = false,
                      bool gemv_by4 = false, int gemm_kc = 0)
{
  const int N = setup->horizon;
  const int nx = 13 * N, nu = 12 * N, nc = 16 * N;
  F.N = N;

  // ---- SolverMPC.cpp:374-393 joint angles, offsets, fmod --------------------------------------
  const double PI = 3.14159265359;
  for (int i = 0; i < 10; i++) F.q[i] = (T)u->joint_angles[i];
  F.q[2] = (T)((double)F.q[2] + 0.3 * PI);
  F.q[3] = (T)((double)F.q[3] - 0.6 * PI);
  F.q[4] = (T)((double)F.q[4] + 0.3 * PI);
  F.q[7] = (T)((double)F.q[7] + 0.3 * PI);
  F.q[8] = (T)((double)F.q[8] - 0.6 * PI);
  F.q[9] = (T)((double)F.q[9] + 0.3 * PI);
  const double PI2 = 2 * PI;
  for (int i = 0; i < 10; i++) F.q[i] = (T)fmod((double)F.q[i], PI2);

  // ---- RobotState.cpp:9-53 ---------------------------------------------------------------------
  T quat[4] = {(T)u->q[0], (T)u->q[1], (T)u->q[2], (T)u->q[3]};
  quat_to_R(quat, F.R);
  T r_feet[3][2];
  for (int rs = 0; rs < 3; rs++)
    for (int c = 0; c < 2; c++) r_feet[rs][c] = (T)u->r[rs * 2 + c];
  const T I_body[3] = {(T)0.5413, (T)0.5200, (T)0.0691};

  // ---- SolverMPC.cpp:333-342 quat_to_rpy --------------------------------------------------------
  {
    T qw = quat[0], qx = quat[1], qy = quat[2], qz = quat[3];
    double as_d = 2. * (double)(qw * qy - qx * qz);
    if (!(as_d < .99999)) as_d = .99999;  // t_min(a,b): a<b ? a : b
    T as = (T)as_d;
    F.rpy[0] = (T)atan2((double)((T)2 * (qw * qx + qy * qz)), 1. - (double)((T)2 * (qx * qx + qy * qy)));
    F.rpy[1] = trig_as_compiled ? pitch_from_sine<T, T>(as) : pitch_from_sine<T, double>(as);
    F.rpy[2] = (T)atan2((double)((T)2 * (qw * qz + qx * qy)), 1. - (double)((T)2 * (qy * qy + qz * qz)));
  }
  // ---- SolverMPC.cpp:65-89 euler_to_rotation (Rb = inverse of the rate map) --------------------
  {
    T Rbm[9];
    if (trig_as_compiled) euler_rate_map<T, T>(F.rpy[1], F.rpy[2], Rbm);
    else euler_rate_map<T, double>(F.rpy[1], F.rpy[2], Rbm);
    inverse3(Rbm, F.Rb);
  }
  // ---- SolverMPC.cpp:420-421 --------------------------------------------------------------------
  for (int i = 0; i < 3; i++) {
    F.x0[i] = F.rpy[i];
    F.x0[3 + i] = (T)u->p[i];
    F.x0[6 + i] = (T)u->w[i];
    F.x0[9 + i] = (T)u->v[i];
  }
  F.x0[12] = (T)9.81f;
  {
    // I_world = (R * I_body) * R^T ; I_body diagonal so (R*I_body)(i,k) = R(i,k)*Id[k] exactly
    T RI[9], Rt[9];
    for (int i = 0; i < 3; i++)
      for (int k = 0; k < 3; k++) {
        RI[i * 3 + k] = F.R[i * 3 + k] * I_body[k];
        Rt[i * 3 + k] = F.R[k * 3 + i];
      }
    matmul(RI, Rt, F.I_world, 3, 3, 3);
  }
  // ---- SolverMPC.cpp:312-331 ct_ss_mats, m = 9.0 (SolverMPC.cpp:423) ---------------------------
  {
    memset(F.A_ct, 0, sizeof(F.A_ct));
    memset(F.B_ct, 0, sizeof(F.B_ct));
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) F.A_ct[i * 13 + 6 + j] = F.Rb[i * 3 + j];
    for (int i = 0; i < 3; i++) F.A_ct[(3 + i) * 13 + 9 + i] = (T)1;
    F.A_ct[11 * 13 + 12] = (T)-1;
    T I_inv[9];
    inverse3(F.I_world, I_inv);
    const T m = (T)9.0;
    for (int b = 0; b < 2; b++) {
      T rx = r_feet[0][b], ry = r_feet[1][b], rz = r_feet[2][b];
      T cm[9] = {(T)0, -rz, ry, rz, (T)0, -rx, -ry, rx, (T)0};
      T blk[9];
      matmul(I_inv, cm, blk, 3, 3, 3);
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) F.B_ct[(6 + i) * 12 + b * 3 + j] = blk[i * 3 + j];
    }
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        F.B_ct[(6 + i) * 12 + 6 + j] = I_inv[i * 3 + j];
        F.B_ct[(6 + i) * 12 + 9 + j] = I_inv[i * 3 + j];
      }
    for (int i = 0; i < 3; i++) {
      F.B_ct[(9 + i) * 12 + 0 + i] = (T)1 / m;
      F.B_ct[(9 + i) * 12 + 3 + i] = (T)1 / m;
    }
  }
  // ---- SolverMPC.cpp:426-433 foot rotations ----------------------------------------------------
  for (int leg = 0; leg < 2; leg++) {
    if (trig_as_compiled) foot_rotation<T, T>(&F.q[5 * leg], F.Rfoot[leg]);
    else foot_rotation<T, double>(&F.q[5 * leg], F.Rfoot[leg]);
  }

  // ---- SolverMPC.cpp:133-193 c2qp (forward Euler, re-powered blocks) ---------------------------
  {
    const T dt = (T)setup->dt;
    for (int i = 0; i < 169; i++) F.Acd[i] = ((i / 13 == i % 13) ? (T)1 : (T)0) + dt * F.A_ct[i];
    for (int i = 0; i < 156; i++) F.Bcd[i] = dt * F.B_ct[i];
    F.A_qp.assign((size_t)nx * 13, (T)0);
    F.B_qp.assign((size_t)nx * nu, (T)0);
    // The reference re-powers from the identity for every block (SolverMPC.cpp:148-177: `Acdm *= Acd` i+1 times per
    // A_qp block, i-j times per B_qp block — 220 13x13 products at N = 10); so does this restatement, so that the CPU
    // baseline timed from it does the reference's floating-point work, not less.  P_k = (((I*Acd)*Acd)*...).
    T Ident[169], Pa[169], Pb[169];
    for (int i = 0; i < 169; i++) Ident[i] = (i / 13 == i % 13) ? (T)1 : (T)0;
    auto power = [&](int k, T* out) {  // out = I * Acd^k by k successive right-multiplications
      memcpy(out, Ident, sizeof(Ident));
      for (int t = 0; t < k; t++) {
        matmul(out, F.Acd, Pb, 13, 13, 13);
        memcpy(out, Pb, sizeof(Pb));
      }
    };
    for (int i = 0; i < N; i++) {
      power(i + 1, Pa);
      memcpy(&F.A_qp[(size_t)i * 169], Pa, 169 * sizeof(T));
    }
    T blk[156];
    for (int i = 0; i < N; i++)
      for (int j = 0; j <= i; j++) {
        power(i - j, Pa);
        matmul(Pa, F.Bcd, blk, 13, 13, 12);
        for (int r = 0; r < 13; r++) memcpy(&F.B_qp[(size_t)(13 * i + r) * nu + 12 * j], &blk[r * 12], 12 * sizeof(T));
      }
  }

  // ---- SolverMPC.cpp:466-482 bounds -------------------------------------------------------------
  F.lb.assign(nc, (T)0);
  F.ub.assign(nc, (T)0);
  for (int leg = 0; leg < 2; leg++)
    for (int i = 0; i < N; i++) {
      for (int j = 0; j < 4; j++) {
        F.ub[8 * leg + j + 16 * i] = (T)kBigNumber;
        F.lb[8 * leg + j + 16 * i] = (T)0.0f;
      }
      F.ub[8 * leg + 4 + 16 * i] = (T)0.01f;
      F.ub[8 * leg + 5 + 16 * i] = (T)0.0f;
      F.ub[8 * leg + 6 + 16 * i] = (T)0.0f;
      F.ub[8 * leg + 7 + 16 * i] = (T)setup->f_max * (T)u->gait[2 * i + leg];
      F.lb[8 * leg + 4 + 16 * i] = (T)0.0f;
      F.lb[8 * leg + 5 + 16 * i] = (T)(-kBigNumber);
      F.lb[8 * leg + 6 + 16 * i] = (T)(-kBigNumber);
      F.lb[8 * leg + 7 + 16 * i] = (T)0.0f;
    }

  // ---- SolverMPC.cpp:488-548 F_control ----------------------------------------------------------
  {
    const T mu = (T)2.0;  // local literal shadows setup->mu (quirk Q3)
    const T lt = (T)0.09, lh = (T)0.06;
    memset(F.Fblk, 0, sizeof(F.Fblk));
    T* Fc = F.Fblk;
    for (int leg = 0; leg < 2; leg++) {
      const int r0 = 8 * leg, cF = 3 * leg, cM = 6 + 3 * leg;
      const T* Rf = F.Rfoot[leg];
      Fc[(r0 + 0) * 12 + cF + 0] = -mu; Fc[(r0 + 0) * 12 + cF + 2] = (T)1;
      Fc[(r0 + 1) * 12 + cF + 0] = mu;  Fc[(r0 + 1) * 12 + cF + 2] = (T)1;
      Fc[(r0 + 2) * 12 + cF + 1] = -mu; Fc[(r0 + 2) * 12 + cF + 2] = (T)1;
      Fc[(r0 + 3) * 12 + cF + 1] = mu;  Fc[(r0 + 3) * 12 + cF + 2] = (T)1;
      // v * R_foot^T * R^T evaluated left to right; the selector vectors have a single non-zero
      // so the first product is exact: (e_k * s) * R_foot^T = s * R_foot(:,k)^T.
      T xw[3], yw[3], zlt[3], zlh[3];
      for (int j = 0; j < 3; j++) {
        // second product: sum_k v1(k) * R(j,k), k = 0,1,2
        T v1x[3] = {Rf[0], Rf[3], Rf[6]};                       // Moment_selection -> column 0
        T v1y[3] = {Rf[1], Rf[4], Rf[7]};                       // M_vec -> column 1
        T v1t[3] = {-lt * Rf[2], -lt * Rf[5], -lt * Rf[8]};     // -lt_vec -> column 2 scaled
        T v1h[3] = {-lh * Rf[2], -lh * Rf[5], -lh * Rf[8]};     // -lh_vec
        xw[j] = (v1x[0] * F.R[j * 3 + 0] + v1x[1] * F.R[j * 3 + 1]) + v1x[2] * F.R[j * 3 + 2];
        yw[j] = (v1y[0] * F.R[j * 3 + 0] + v1y[1] * F.R[j * 3 + 1]) + v1y[2] * F.R[j * 3 + 2];
        zlt[j] = (v1t[0] * F.R[j * 3 + 0] + v1t[1] * F.R[j * 3 + 1]) + v1t[2] * F.R[j * 3 + 2];
        zlh[j] = (v1h[0] * F.R[j * 3 + 0] + v1h[1] * F.R[j * 3 + 1]) + v1h[2] * F.R[j * 3 + 2];
      }
      for (int j = 0; j < 3; j++) {
        Fc[(r0 + 4) * 12 + cM + j] = xw[j];
        Fc[(r0 + 5) * 12 + cF + j] = zlt[j];
        Fc[(r0 + 5) * 12 + cM + j] = yw[j];
        Fc[(r0 + 6) * 12 + cF + j] = zlh[j];
        // left foot uses -M_vec, right foot +M_vec (quirk Q5, SolverMPC.cpp:525-526 vs 545-546)
        Fc[(r0 + 6) * 12 + cM + j] = (leg == 0) ? -yw[j] : yw[j];
      }
      Fc[(r0 + 7) * 12 + cF + 2] = (T)2;
    }
  }

  // ---- SolverMPC.cpp:450-461, 557-570 cost ------------------------------------------------------
  {
    std::vector<T> wdiag(nx), Xd(nx, (T)0);
    for (int i = 0; i < N; i++) {
      for (int j = 0; j < 12; j++) {
        wdiag[13 * i + j] = (T)u->weights[j];
        Xd[13 * i + j] = (T)u->traj[12 * i + j];
      }
      wdiag[13 * i + 12] = (T)0;
    }
    // qH = 2 * ((B^T * S) * B + Alpha_rep) with S and Alpha_rep DENSE like the reference's (SolverMPC.cpp:31,45,454,566):
    // the products with their zeros are carried out (exact, so the bits equal a diagonal scaling), because this function
    // is also what the CPU baseline times.
    std::vector<T> Bt((size_t)nu * nx), Sd((size_t)nx * nx, (T)0), T1((size_t)nu * nx);
    for (int k = 0; k < nx; k++) {
      Sd[(size_t)k * nx + k] = wdiag[k];
      for (int i = 0; i < nu; i++) Bt[(size_t)i * nx + k] = F.B_qp[(size_t)k * nu + i];
    }
    matmul(Bt.data(), Sd.data(), T1.data(), nu, nx, nx);
    std::vector<T> BSB((size_t)nu * nu);
    if (gemm_kc > 0) matmul_kblocked(T1.data(), F.B_qp.data(), BSB.data(), nu, nx, nu, gemm_kc);
    else matmul(T1.data(), F.B_qp.data(), BSB.data(), nu, nx, nu);
    F.H.assign((size_t)nu * nu, (T)0);
    for (int i = 0; i < nu; i++)
      for (int j = 0; j < nu; j++) {
        const T alpha = (i == j) ? (T)u->Alpha_K[i % 12] : (T)0;
        F.H[(size_t)i * nu + j] = (T)2 * (BSB[(size_t)i * nu + j] + alpha);
      }
    // d = A_qp * x0 - X_d ; g = (2*B^T*S) * d
    // gemv_by4 (sensitivity probe, see "what stays a restatement" in the header): the two matrix-vector products summed
    // the way Eigen 3.3's column-major gemv kernel groups them — four columns at a time, pairwise inside the group,
    // res += (a0 x0 + a1 x1) + (a2 x2 + a3 x3), leftover columns one by one — instead of one sequential sum.
    auto gemv_row = [gemv_by4](int n, auto coef, auto vec) {
      if (!gemv_by4) {
        T acc = coef(0) * vec(0);
        for (int k = 1; k < n; k++) acc = acc + coef(k) * vec(k);
        return acc;
      }
      T res = (T)0;
      int k = 0;
      for (; k + 4 <= n; k += 4)
        res = res + ((coef(k) * vec(k) + coef(k + 1) * vec(k + 1)) + (coef(k + 2) * vec(k + 2) + coef(k + 3) * vec(k + 3)));
      for (; k < n; k++) res = res + coef(k) * vec(k);
      return res;
    };
    std::vector<T> d(nx);
    for (int r = 0; r < nx; r++)
      d[r] = gemv_row(13, [&](int c) { return F.A_qp[(size_t)r * 13 + c]; }, [&](int c) { return F.x0[c]; }) - Xd[r];
    F.g.assign(nu, (T)0);
    for (int i = 0; i < nu; i++)
      F.g[i] = gemv_row(nx, [&](int k) { return T1[(size_t)i * nx + k] * (T)2; }, [&](int k) { return d[k]; });
  }
}

// ---- SolverMPC.cpp:589-637 swing-leg elimination -------------------------------------------------
static inline bool near_zero_f(float a) { return (a < 0.0001 && a > -.0001); }
static inline bool near_two_f(float a) { return near_zero_f(a - 2); }

struct ReducedQP {
  int nv = 0, nc = 0;
  std::vector<int> var_ind, con_ind;
  std::vector<double> H, g, A, lb, ub;
};

template <class T>
static void eliminate(const Formulation<T>& F, ReducedQP& Q)
{
  const int N = F.N, nu = 12 * N, ncon = 16 * N;
  std::vector<char> var_elim(nu + 16, 0), con_elim(ncon + 16, 0);
  for (int i = 0; i < ncon; i++) {
    if (!(near_zero_f((float)(double)F.lb[i]) && near_zero_f((float)(double)F.ub[i]))) continue;
    // row i of the block-diagonal constraint matrix
    const int step = i / 16, lr = i % 16;
    for (int j = 0; j < nu; j++) {
      double c = (j / 12 == step) ? (double)F.Fblk[lr * 12 + j % 12] : 0.0;
      if (near_two_f((float)c)) {
        int cs = (j % 2 == 0) ? (j + 4) / 6 * 8 - 1 : (j + 1) / 6 * 8 + 7;
        var_elim[j + 6] = var_elim[j + 5] = var_elim[j + 4] = 1;
        var_elim[j - 2] = var_elim[j - 1] = var_elim[j] = 1;
        for (int k = 0; k < 8; k++) con_elim[cs - k] = 1;
      }
    }
  }
  Q.var_ind.clear();
  Q.con_ind.clear();
  for (int i = 0; i < nu; i++)
    if (!var_elim[i]) Q.var_ind.push_back(i);
  for (int i = 0; i < ncon; i++)
    if (!con_elim[i]) Q.con_ind.push_back(i);
  Q.nv = (int)Q.var_ind.size();
  Q.nc = (int)Q.con_ind.size();
  Q.H.resize((size_t)Q.nv * Q.nv);
  Q.g.resize(Q.nv);
  Q.A.resize((size_t)Q.nc * Q.nv);
  Q.lb.resize(Q.nc);
  Q.ub.resize(Q.nc);
  for (int i = 0; i < Q.nv; i++) {
    int oa = Q.var_ind[i];
    Q.g[i] = (double)F.g[oa];
    for (int j = 0; j < Q.nv; j++) Q.H[(size_t)i * Q.nv + j] = (double)F.H[(size_t)oa * nu + Q.var_ind[j]];
  }
  for (int c = 0; c < Q.nc; c++) {
    int oc = Q.con_ind[c];
    for (int s = 0; s < Q.nv; s++) {
      int ov = Q.var_ind[s];
      double v = (ov / 12 == oc / 16) ? (double)F.Fblk[(oc % 16) * 12 + ov % 12] : 0.0;
      Q.A[(size_t)c * Q.nv + s] = (double)(float)v;  // `float cval = ...` SolverMPC.cpp:687
    }
    Q.lb[c] = (double)F.lb[oc];
    Q.ub[c] = (double)F.ub[oc];
  }
}

// ---- SolverMPC.cpp:702-712 : the reference's own solver, called the reference's way --------------
static int solve_reduced(ReducedQP& Q, std::vector<double>& x, int* nwsr_out)
{
#ifdef ORACLE_WITH_QPOASES
  x.assign(Q.nv > 0 ? Q.nv : 1, 0.0);
  if (Q.nv == 0) { if (nwsr_out) *nwsr_out = 0; return 0; }
  qpOASES::int_t nWSR = 500;
  qpOASES::QProblem problem_red(Q.nv, Q.nc);
  qpOASES::Options op;
  op.setToMPC();
  op.printLevel = qpOASES::PL_NONE;
  problem_red.setOptions(op);
  int rval = problem_red.init(Q.H.data(), Q.g.data(), Q.A.data(), NULL, NULL, Q.lb.data(), Q.ub.data(), nWSR);
  int rval2 = problem_red.getPrimalSolution(x.data());
  if (nwsr_out) *nwsr_out = (int)nWSR;
  return (rval2 != qpOASES::SUCCESSFUL_RETURN) ? 1000 + rval2 : (rval != qpOASES::SUCCESSFUL_RETURN ? rval : 0);
#else
  (void)Q; (void)x; (void)nwsr_out;
  return -1;
#endif
}

template <class T>
static int solve_one(const update_data_t* u, const problem_setup* s, double* q_soln, int* info, bool trig_as_compiled = false,
                     bool gemv_by4 = false, int gemm_kc = 0)
{
  Formulation<T> F;
  formulate<T>(u, s, F, trig_as_compiled, gemv_by4, gemm_kc);
  ReducedQP Q;
  eliminate(F, Q);
  std::vector<double> x;
  int nwsr = 0;
  int rc = solve_reduced(Q, x, &nwsr);
  const int nu = 12 * s->horizon;
  // SolverMPC.cpp:720-732 scatter, eliminated variables = 0
  for (int i = 0; i < nu; i++) q_soln[i] = 0.0;
  for (int i = 0; i < Q.nv; i++) q_soln[Q.var_ind[i]] = x[i];
  if (info) { info[0] = rc; info[1] = nwsr; info[2] = Q.nv; info[3] = Q.nc; }
  return rc;
}

// =================================================================================================
// C entry points (ctypes-friendly)
// =================================================================================================
extern "C" {

int oracle_has_qpoases(void)
{
#ifdef ORACLE_WITH_QPOASES
  return 1;
#else
  return 0;
#endif
}

size_t oracle_sizeof_update_data(void) { return sizeof(update_data_t); }

/* Solve `n` records.  `mode` is a bit mask:
 *   bit 0 (ORACLE_MODE_FP64)  the formulation carried in double (sensitivity probe, not the reference);
 *                             0 = the reference's arithmetic (fp32 formulation, fp64 solve);
 *   bit 1 (ORACLE_MODE_TRIG_AS_COMPILED)  trig calls resolve as in the reference's translation unit
 *                             (float overloads, see "trig resolution" above); 0 = canonical double trig.
 *   bit 2 (ORACLE_MODE_GEMV_BY4)  sensitivity probe: the two matrix-vector products of :570 summed the way Eigen 3.3's
 *                             column-major gemv kernel groups them (four columns at a time) instead of sequentially.
 * q_soln [n][12N]; info [n][4] = {return code, nWSR, reduced vars, reduced cons}. */
int oracle_solve_batch(const update_data_t* u, int n, const problem_setup* s, int mode,
                       double* q_soln, int* info)
{
  int bad = 0;
  const int nu = 12 * s->horizon;
  const bool tac = (mode & 2) != 0, by4 = (mode & 4) != 0;
  const int kc = (mode >> 8) & 0xff;  // bits 8-15: depth block of the Hessian product's k-sum (0: one sequential sum)
  for (int i = 0; i < n; i++) {
    int rc = (mode & 1) ? solve_one<double>(&u[i], s, q_soln + (size_t)i * nu, info ? info + 4 * i : NULL, tac, by4, kc)
                        : solve_one<float>(&u[i], s, q_soln + (size_t)i * nu, info ? info + 4 * i : NULL, tac, by4, kc);
    if (rc) bad++;
  }
  return bad;
}

/* Timed variant for the CPU baseline: solves records [0,n) `reps` times round-robin on the calling
 * thread and returns per-solve wall times (seconds) measured with clock_gettime(CLOCK_MONOTONIC)
 * like the reference's Timer (include/common/Utilities/Timer.h:15-49).  The timed region is the
 * equivalent of SolverMPC.cpp:374-732 plus the per-tick buffer (re)allocation the reference does in
 * resize_qp_mats (all formulation buffers here are allocated per call, as there). */
int oracle_time_solves(const update_data_t* u, int n, const problem_setup* s, int total,
                       double* seconds_out)
{
  std::vector<double> q((size_t)12 * s->horizon);
  int info[4];
  int bad = 0;
  for (int t = 0; t < total; t++) {
    struct timespec a, b;
    clock_gettime(CLOCK_MONOTONIC, &a);
    int rc = solve_one<float>(&u[t % n], s, q.data(), info);
    clock_gettime(CLOCK_MONOTONIC, &b);
    if (rc) bad++;
    seconds_out[t] = (double)(b.tv_sec - a.tv_sec) + 1e-9 * (double)(b.tv_nsec - a.tv_nsec);
  }
  return bad;
}

/* Formulation only (fp32, the reference's arithmetic): full un-reduced QP data of one record.
 * H [n*n] row-major, g [n], Fblk [192], lb/ub [16N]; optional intermediates (may be NULL):
 * x0 [13], Acd [169], Bcd [156], Rfoot [18], Rmat [9], A_qp [13N*13]. */
int oracle_formulate_f32_mode(const update_data_t* u, const problem_setup* s, int mode, float* H, float* g,
                              float* Fblk, float* lb, float* ub, float* x0, float* Acd, float* Bcd, float* Rfoot,
                              float* Rmat, float* A_qp)
{
  Formulation<float> F;
  formulate<float>(u, s, F, (mode & 2) != 0);
  const int N = s->horizon, nu = 12 * N;
  if (H) memcpy(H, F.H.data(), sizeof(float) * (size_t)nu * nu);
  if (g) memcpy(g, F.g.data(), sizeof(float) * nu);
  if (Fblk) memcpy(Fblk, F.Fblk, sizeof(F.Fblk));
  if (lb) memcpy(lb, F.lb.data(), sizeof(float) * 16 * N);
  if (ub) memcpy(ub, F.ub.data(), sizeof(float) * 16 * N);
  if (x0) memcpy(x0, F.x0, sizeof(F.x0));
  if (Acd) memcpy(Acd, F.Acd, sizeof(F.Acd));
  if (Bcd) memcpy(Bcd, F.Bcd, sizeof(F.Bcd));
  if (Rfoot) memcpy(Rfoot, F.Rfoot, sizeof(F.Rfoot));
  if (Rmat) memcpy(Rmat, F.R, sizeof(F.R));
  if (A_qp) memcpy(A_qp, F.A_qp.data(), sizeof(float) * (size_t)13 * N * 13);
  return 0;
}

int oracle_formulate_f32(const update_data_t* u, const problem_setup* s, float* H, float* g, float* Fblk,
                         float* lb, float* ub, float* x0, float* Acd, float* Bcd, float* Rfoot,
                         float* Rmat, float* A_qp)
{
  return oracle_formulate_f32_mode(u, s, 0, H, g, Fblk, lb, ub, x0, Acd, Bcd, Rfoot, Rmat, A_qp);
}

/* The reduced QP exactly as handed to qpOASES (doubles), for independent QP cross-checks.
 * Buffers sized for the un-reduced problem.  Returns nv, writes nc. */
int oracle_reduced_qp(const update_data_t* u, const problem_setup* s, int mode, double* H,
                      double* g, double* A, double* lb, double* ub, int* var_ind, int* con_ind, int* nc_out)
{
  ReducedQP Q;
  const bool tac = (mode & 2) != 0;
  if (mode & 1) { Formulation<double> F; formulate<double>(u, s, F, tac); eliminate(F, Q); }
  else { Formulation<float> F; formulate<float>(u, s, F, tac); eliminate(F, Q); }
  memcpy(H, Q.H.data(), sizeof(double) * Q.H.size());
  memcpy(g, Q.g.data(), sizeof(double) * Q.g.size());
  memcpy(A, Q.A.data(), sizeof(double) * Q.A.size());
  memcpy(lb, Q.lb.data(), sizeof(double) * Q.lb.size());
  memcpy(ub, Q.ub.data(), sizeof(double) * Q.ub.size());
  memcpy(var_ind, Q.var_ind.data(), sizeof(int) * Q.nv);
  memcpy(con_ind, Q.con_ind.data(), sizeof(int) * Q.nc);
  *nc_out = Q.nc;
  return Q.nv;
}

/* ------------------------------------------------------------------------------------------------
 * Row f-2 of SURVEY.md §8: joint torques from the first-step wrench, restating
 *   LegController.cpp:108-166 (computeLegJacobianAndPosition, J_force_moment 6x5, double)
 *   LegController.cpp:57-63   (legtau = J_force_moment^T * feedforwardForce)
 *   ConvexMPCLocomotion.cpp:419-440 (f_ff[leg] = -rBody * [GRF; GRM])
 * q5 = the leg's joint angles as LegController holds them (raw + first 0.3/-0.6/0.3*3.14159 offset).
 * J is written with named sub-sums instead of the reference's expanded literals.
 * ------------------------------------------------------------------------------------------------ */
void oracle_leg_jacobian_fm(const double* q5, int leg, double* J /* [6][5] row-major */)
{
  const double side = (leg == 0) ? 1.0 : -1.0;
  const double s0 = sin(q5[0]), c0 = cos(q5[0]), s1 = sin(q5[1]), c1 = cos(q5[1]);
  const double q23 = q5[2] + q5[3], q234 = q5[2] + q5[3] + q5[4];
  // lever sums of the thigh (0.22), calf (0.22) and foot (0.04) links seen from joints 2, 3 and 4
  const double S[3] = {0.04 * sin(q234) + 0.22 * sin(q23) + 0.22 * sin(q5[2]), 0.04 * sin(q234) + 0.22 * sin(q23), 0.04 * sin(q234)};
  const double C[3] = {0.04 * cos(q234) + 0.22 * cos(q23) + 0.22 * cos(q5[2]), 0.04 * cos(q234) + 0.22 * cos(q23), 0.04 * cos(q234)};
  const double h = 0.018 * side + 0.0025, e = 0.015 * side;
  for (int i = 0; i < 30; i++) J[i] = 0.0;
  const double a = e + c1 * h - 1.0 * s1 * C[0];
  J[0 * 5 + 0] = s0 * (S[0] + 0.0135) + c0 * a;
  J[1 * 5 + 0] = s0 * a - 1.0 * c0 * (S[0] + 0.0135);
  J[5 * 5 + 0] = 1.0;
  const double b = s1 * h + c1 * C[0];
  J[0 * 5 + 1] = -1.0 * s0 * b;
  J[1 * 5 + 1] = c0 * b;
  J[2 * 5 + 1] = s1 * C[0] - 1.0 * c1 * h;
  J[3 * 5 + 1] = c0;
  J[4 * 5 + 1] = s0;
  for (int k = 0; k < 3; k++) {
    J[0 * 5 + 2 + k] = s0 * s1 * S[k] - 1.0 * c0 * C[k];
    J[1 * 5 + 2 + k] = -1.0 * s0 * C[k] - 1.0 * c0 * s1 * S[k];
    J[2 * 5 + 2 + k] = c1 * S[k];
    J[3 * 5 + 2 + k] = -c1 * s0;
    J[4 * 5 + 2 + k] = c0 * c1;
    J[5 * 5 + 2 + k] = s1;
  }
}

/* tau[leg][j] for n robots: wrench12 = first-step wrench (get_solution(0..11)), rBody row-major world->body,
 * q_leg[2][5] as above, contact[2] = first-step contact flags (swing legs get no feed-forward force). */
void oracle_joint_torques(const double* wrench12, const double* rBody, const double* q_leg, const int* contact, int n,
                          double* tau /* [n][10] */)
{
  for (int i = 0; i < n; i++) {
    const double* w = wrench12 + 12 * i;
    const double* Rb = rBody + 9 * i;
    for (int leg = 0; leg < 2; leg++) {
      double f[6] = {0, 0, 0, 0, 0, 0};
      if (contact[2 * i + leg]) {
        for (int r = 0; r < 3; r++) {
          f[r] = -(Rb[r * 3] * w[3 * leg] + Rb[r * 3 + 1] * w[3 * leg + 1] + Rb[r * 3 + 2] * w[3 * leg + 2]);
          f[3 + r] = -(Rb[r * 3] * w[6 + 3 * leg] + Rb[r * 3 + 1] * w[6 + 3 * leg + 1] + Rb[r * 3 + 2] * w[6 + 3 * leg + 2]);
        }
      }
      double J[30];
      oracle_leg_jacobian_fm(q_leg + 10 * i + 5 * leg, leg, J);
      for (int j = 0; j < 5; j++) {
        double t = 0;
        for (int r = 0; r < 6; r++) t += J[r * 5 + j] * f[r];
        tau[10 * i + 5 * leg + j] = t;
      }
    }
  }
}

}  // extern "C"
