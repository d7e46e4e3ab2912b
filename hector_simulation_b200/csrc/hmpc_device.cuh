// hmpc_device.cuh — sm_100a device code of the batched force-and-moment MPC solver.
//
// One CTA solves one robot's per-tick QP end to end:
//   stage 0  cp.async.bulk (TMA 1-D) of the packed record into shared memory
//   stage 1  SRBD linearisation + foot rotations + constraint rows      (SolverMPC.cpp:374-433, 463-548)
//   stage 2  forward-Euler discretisation, powers, Toeplitz blocks       (SolverMPC.cpp:133-193)
//   stage 3  Hessian / gradient of the condensed QP, swing-leg removal   (SolverMPC.cpp:450-461, 557-570, 589-697)
//   stage 4  in-register symmetric sweep inversion of H (fp64, 6x6 block per thread)
//   stage 5  dual active-set iterations on the explicit inverse           (replaces qpOASES, SolverMPC.cpp:702-712)
//   stage 6  scatter of the optimal wrenches, eliminated entries = 0     (SolverMPC.cpp:720-732)
//
// Stages 1-3 reproduce the reference's float32 arithmetic operation by operation (separately rounded
// multiply and add, same summation order) so that the QP data equals the oracle's bit for bit; they use
// the *_rn intrinsics, which the compiler never contracts into FMAs.  Stages 4-5 work in float64 on the
// float32-rounded data, like the reference hands float data to a double solver (SolverMPC.cpp:573-577).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hmpc {

#define FM(a, b) __fmul_rn((a), (b))
#define FA(a, b) __fadd_rn((a), (b))
#define FS(a, b) __fsub_rn((a), (b))
#define FD(a, b) __fdiv_rn((a), (b))
#define DM(a, b) __dmul_rn((a), (b))
#define DA(a, b) __dadd_rn((a), (b))
#define DS(a, b) __dsub_rn((a), (b))

// termination codes (low byte of the status word, include/hector_mpc_b200.h)
enum : int { ST_OK = 0, ST_ITER_CAP = 1, ST_WS_CAP = 2, ST_INFEASIBLE = 3, ST_NOT_SPD = 4 };

struct KernelArgs {
  const unsigned char* records;  // packed device records
  int rec_stride;                // bytes, multiple of 16
  int batch;
  int horizon;                   // N
  float dt;
  float f_max;
  int nb_lo, nb_hi;              // this launch handles instances with nb_lo < NB <= nb_hi
  int nb_cap;                    // capacity (blocks of 6 variables) the shared-memory carve is sized for
  int qmax;                      // working-set capacity
  int max_iter;
  float* wrench;                 // [batch][12N]
  int* status;                   // [batch]
  // assembly dump (parity hook); all null in production launches
  float* dbg_H;                  // [batch][12N*12N]
  float* dbg_g;                  // [batch][12N]
  float* dbg_F;                  // [batch][192]
  float* dbg_lb;                 // [batch][16N]
  float* dbg_ub;                 // [batch][16N]
};

// ------------------------------------------------------------------------------------------------
// shared-memory carve-up (byte offsets), identical on host and device
// ------------------------------------------------------------------------------------------------
struct Layout {
  int H, gq, x0, x, w, HA, nrm, rhs, blk, misc, uni;
  // solver view of the union
  int Li, lam, dv, yv, rv, Wc, act;
  // assembly view of the union
  int rec, x0f, Acd, Bcd, P, M, T, dd, fbl, wts;
  int total;
};

__host__ __device__ inline int align16(int x) { return (x + 15) & ~15; }

__host__ __device__ inline Layout make_layout(int N, int nb_cap, int qmax, int rec_stride)
{
  Layout L;
  const int nbt = nb_cap * (nb_cap + 1) / 2;
  const int n = 6 * nb_cap, m = 10 * nb_cap;
  int o = 0;
  L.H = o;    o += nbt * 36 * 8;
  L.gq = o;   o += n * 8;
  L.x0 = o;   o += n * 8;
  L.x = o;    o += n * 8;   // doubles as sweep pivot-column buffer 0
  L.w = o;    o += n * 8;   // doubles as sweep pivot-column buffer 1
  L.HA = o;   o += n * 8;
  L.nrm = o;  o += 2 * 10 * 6 * 8;
  L.rhs = o;  o += m * 8;
  L.blk = o;  o += align16(nb_cap * 4 + 2 * N * 4 * 2);  // block -> (step,leg) and (step,leg) -> block
  L.misc = o; o += 512;
  L.uni = o;
  // solver view
  int s = L.uni;
  L.Li = s;   s += (qmax + 1) * (qmax + 2) / 2 * 8;
  L.lam = s;  s += (qmax + 2) * 8;
  L.dv = s;   s += (qmax + 2) * 8;
  L.yv = s;   s += (qmax + 2) * 8;
  L.rv = s;   s += (qmax + 2) * 8;
  L.Wc = s;   s += align16((qmax + 2) * 4);
  L.act = s;  s += align16(m);
  // assembly view
  int a = L.uni;
  L.rec = a;  a += align16(rec_stride);
  L.x0f = a;  a += 16 * 4;
  L.Acd = a;  a += align16(169 * 4);
  L.Bcd = a;  a += align16(156 * 4);
  L.P = a;    a += align16((N + 1) * 169 * 4);
  L.M = a;    a += align16(N * 156 * 4);
  L.T = a;    a += align16(N * 144 * 4);
  L.dd = a;   a += align16(13 * N * 4);
  L.fbl = a;  a += 192 * 4;
  L.wts = a;  a += 16 * 4;
  L.total = align16(s > a ? s : a);
  return L;
}

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase)
{
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ int leg_of(int c12) { return (c12 / 3) & 1; }            // column of a 12-wide step -> leg
__device__ __forceinline__ int loc_of(int c12) { return (c12 % 3) + (c12 >= 6 ? 3 : 0); }  // -> slot in [F(3) M(3)]
__device__ __forceinline__ int col12_of(int leg, int loc) { return (loc < 3) ? 3 * leg + loc : 6 + 3 * leg + (loc - 3); }

// element (i,j) of the symmetric matrix stored as lower 6x6 blocks; i = 6*ib+r, j = 6*jb+c
__device__ __forceinline__ int blk_off(int ib, int jb) { return (ib * (ib + 1) / 2 + jb) * 36; }
__device__ __forceinline__ double hsym(const double* H, int ib, int r, int jb, int c)
{
  return (ib >= jb) ? H[blk_off(ib, jb) + r * 6 + c] : H[blk_off(jb, ib) + c * 6 + r];
}

// Eigen 3x3 inverse restated (oracle: inverse3)
__device__ __forceinline__ float cof3(const float* m, int i, int j)
{
  int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
  return FS(FM(m[i1 * 3 + j1], m[i2 * 3 + j2]), FM(m[i1 * 3 + j2], m[i2 * 3 + j1]));
}
__device__ inline void inverse3(const float* m, float* inv)
{
  float c00 = cof3(m, 0, 0), c10 = cof3(m, 1, 0), c20 = cof3(m, 2, 0);
  float det = FA(FA(FM(c00, m[0]), FM(c10, m[3])), FM(c20, m[6]));
  float id = FD(1.0f, det);
  inv[0] = FM(c00, id);
  inv[1] = FM(c10, id);
  inv[2] = FM(c20, id);
  inv[3] = FM(cof3(m, 0, 1), id);
  inv[4] = FM(cof3(m, 1, 1), id);
  inv[5] = FM(cof3(m, 2, 1), id);
  inv[6] = FM(cof3(m, 0, 2), id);
  inv[7] = FM(cof3(m, 1, 2), id);
  inv[8] = FM(cof3(m, 2, 2), id);
}
// row-major 3x3 product, sequential k (oracle: matmul)
__device__ inline void mul3(const float* A, const float* B, float* C)
{
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      float acc = FM(A[i * 3], B[j]);
      acc = FA(acc, FM(A[i * 3 + 1], B[3 + j]));
      acc = FA(acc, FM(A[i * 3 + 2], B[6 + j]));
      C[i * 3 + j] = acc;
    }
}

// foot rotation from five offset-corrected joint angles (SolverMPC.cpp:428-433; oracle: foot_rotation)
__device__ inline void foot_rotation(const float* q, float* Rf)
{
  double s0, c0, s1, c1, s2, c2, s3, c3, s4, c4;
  sincos((double)q[0], &s0, &c0);
  sincos((double)q[1], &s1, &c1);
  sincos((double)q[2], &s2, &c2);
  sincos((double)q[3], &s3, &c3);
  sincos((double)q[4], &s4, &c4);
  double a = DA(DM(c0, s2), DM(DM(c2, s0), s1));
  double b = DS(DM(c0, c2), DM(DM(s0, s1), s2));
  double c = DA(DM(c2, s0), DM(DM(c0, s1), s2));
  double d = DS(DM(s0, s2), DM(DM(c0, c2), s1));
  float q234 = FA(FA(q[2], q[3]), q[4]);
  double s234, c234;
  sincos((double)q234, &s234, &c234);
  double c3a_s3b = DA(DM(c3, a), DM(s3, b));
  double s3a_c3b = DS(DM(s3, a), DM(c3, b));
  double c3c_s3d = DS(DM(c3, c), DM(s3, d));
  double s3c_c3d = DA(DM(s3, c), DM(c3, d));
  Rf[0] = (float)DS(DM(-s4, c3a_s3b), DM(c4, s3a_c3b));
  Rf[1] = (float)DM(-c1, s0);
  Rf[2] = (float)DS(DM(c4, c3a_s3b), DM(s4, s3a_c3b));
  Rf[3] = (float)DS(DM(c4, c3c_s3d), DM(s4, s3c_c3d));
  Rf[4] = (float)DM(c0, c1);
  Rf[5] = (float)DA(DM(c4, s3c_c3d), DM(s4, c3c_s3d));
  Rf[6] = (float)DM(-s234, c1);
  Rf[7] = (float)s1;
  Rf[8] = (float)DM(c234, c1);
}

// ------------------------------------------------------------------------------------------------
// stage 1: scalar prologue (one thread): fills x0f, Acd, Bcd, Fblk rows
// record floats: p[0..2] v[3..5] q[6..9] w[10..12] r[13..18] joint[19..28] yaw[29] weights[30..41]
//                alpha[42..53] traj[54..54+12N)  then gait bytes
// ------------------------------------------------------------------------------------------------
__device__ inline void prologue(const float* rf, float dt, float* x0f, float* Acd, float* Bcd, float* Fblk)
{
  // joint angles: SolverMPC.cpp:374-393
  const double PI = 3.14159265359;
  float q[10];
  for (int i = 0; i < 10; i++) q[i] = rf[19 + i];
  q[2] = (float)DA((double)q[2], DM(0.3, PI));
  q[3] = (float)DS((double)q[3], DM(0.6, PI));
  q[4] = (float)DA((double)q[4], DM(0.3, PI));
  q[7] = (float)DA((double)q[7], DM(0.3, PI));
  q[8] = (float)DS((double)q[8], DM(0.6, PI));
  q[9] = (float)DA((double)q[9], DM(0.3, PI));
  const double PI2 = DM(2.0, PI);
  for (int i = 0; i < 10; i++) q[i] = (float)fmod((double)q[i], PI2);

  // RobotState::set — Quaternionf::toRotationMatrix
  float R[9];
  {
    float w = rf[6], x = rf[7], y = rf[8], z = rf[9];
    float tx = FM(2.f, x), ty = FM(2.f, y), tz = FM(2.f, z);
    float twx = FM(tx, w), twy = FM(ty, w), twz = FM(tz, w);
    float txx = FM(tx, x), txy = FM(ty, x), txz = FM(tz, x);
    float tyy = FM(ty, y), tyz = FM(tz, y), tzz = FM(tz, z);
    R[0] = FS(1.f, FA(tyy, tzz));
    R[1] = FS(txy, twz);
    R[2] = FA(txz, twy);
    R[3] = FA(txy, twz);
    R[4] = FS(1.f, FA(txx, tzz));
    R[5] = FS(tyz, twx);
    R[6] = FS(txz, twy);
    R[7] = FA(tyz, twx);
    R[8] = FS(1.f, FA(txx, tyy));
  }
  // quat_to_rpy: SolverMPC.cpp:333-342
  float rpy[3];
  {
    float qw = rf[6], qx = rf[7], qy = rf[8], qz = rf[9];
    double as_d = DM(2.0, (double)FS(FM(qw, qy), FM(qx, qz)));
    if (!(as_d < .99999)) as_d = .99999;
    float as = (float)as_d;
    rpy[0] = (float)atan2((double)FM(2.f, FA(FM(qw, qx), FM(qy, qz))),
                          DS(1.0, (double)FM(2.f, FA(FM(qx, qx), FM(qy, qy)))));
    rpy[1] = (float)asin((double)as);
    rpy[2] = (float)atan2((double)FM(2.f, FA(FM(qw, qz), FM(qx, qy))),
                          DS(1.0, (double)FM(2.f, FA(FM(qy, qy), FM(qz, qz)))));
  }
  // euler_to_rotation: SolverMPC.cpp:65-89
  float Rb[9];
  {
    double sp, cp, sy, cy;
    sincos((double)rpy[1], &sp, &cp);
    sincos((double)rpy[2], &sy, &cy);
    float Rbm[9] = {(float)DM(cy, cp), (float)(-sy), 0.f, (float)DM(sy, cp), (float)cy, 0.f, (float)(-sp), 0.f, 1.f};
    inverse3(Rbm, Rb);
  }
  for (int i = 0; i < 3; i++) {
    x0f[i] = rpy[i];
    x0f[3 + i] = rf[i];
    x0f[6 + i] = rf[10 + i];
    x0f[9 + i] = rf[3 + i];
  }
  x0f[12] = 9.81f;
  // I_world, I_inv: SolverMPC.cpp:421, 320
  float Iinv[9];
  {
    const float Ib[3] = {0.5413f, 0.5200f, 0.0691f};
    float RI[9], Rt[9], Iw[9];
    for (int i = 0; i < 3; i++)
      for (int k = 0; k < 3; k++) {
        RI[i * 3 + k] = FM(R[i * 3 + k], Ib[k]);
        Rt[i * 3 + k] = R[k * 3 + i];
      }
    mul3(RI, Rt, Iw);
    inverse3(Iw, Iinv);
  }
  // ct_ss_mats + c2qp's Acd/Bcd: SolverMPC.cpp:312-331, 145-146
  for (int i = 0; i < 169; i++) Acd[i] = 0.f;
  for (int i = 0; i < 156; i++) Bcd[i] = 0.f;
  for (int i = 0; i < 13; i++) Acd[i * 13 + i] = 1.f;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) Acd[i * 13 + 6 + j] = FA(0.f, FM(dt, Rb[i * 3 + j]));
  for (int i = 0; i < 3; i++) Acd[(3 + i) * 13 + 9 + i] = FA(0.f, FM(dt, 1.f));
  Acd[11 * 13 + 12] = FA(0.f, FM(dt, -1.f));
  for (int b = 0; b < 2; b++) {
    float rx = rf[13 + 0 + b], ry = rf[13 + 2 + b], rz = rf[13 + 4 + b];
    float cm[9] = {0.f, -rz, ry, rz, 0.f, -rx, -ry, rx, 0.f};
    float blk[9];
    mul3(Iinv, cm, blk);
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) Bcd[(6 + i) * 12 + b * 3 + j] = FM(dt, blk[i * 3 + j]);
  }
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      float v = FM(dt, Iinv[i * 3 + j]);
      Bcd[(6 + i) * 12 + 6 + j] = v;
      Bcd[(6 + i) * 12 + 9 + j] = v;
    }
  {
    float v = FM(dt, FD(1.f, 9.0f));  // mass literal 9.0, SolverMPC.cpp:423
    for (int i = 0; i < 3; i++) {
      Bcd[(9 + i) * 12 + i] = v;
      Bcd[(9 + i) * 12 + 3 + i] = v;
    }
  }
  // F_control: SolverMPC.cpp:488-548
  for (int i = 0; i < 192; i++) Fblk[i] = 0.f;
  const float mu = 2.0f, lt = 0.09f, lh = 0.06f;
  for (int leg = 0; leg < 2; leg++) {
    float Rf[9];
    foot_rotation(&q[5 * leg], Rf);
    const int r0 = 8 * leg, cF = 3 * leg, cM = 6 + 3 * leg;
    Fblk[(r0 + 0) * 12 + cF + 0] = -mu; Fblk[(r0 + 0) * 12 + cF + 2] = 1.f;
    Fblk[(r0 + 1) * 12 + cF + 0] = mu;  Fblk[(r0 + 1) * 12 + cF + 2] = 1.f;
    Fblk[(r0 + 2) * 12 + cF + 1] = -mu; Fblk[(r0 + 2) * 12 + cF + 2] = 1.f;
    Fblk[(r0 + 3) * 12 + cF + 1] = mu;  Fblk[(r0 + 3) * 12 + cF + 2] = 1.f;
    float v1t[3] = {FM(-lt, Rf[2]), FM(-lt, Rf[5]), FM(-lt, Rf[8])};
    float v1h[3] = {FM(-lh, Rf[2]), FM(-lh, Rf[5]), FM(-lh, Rf[8])};
    for (int j = 0; j < 3; j++) {
      float xw = FA(FA(FM(Rf[0], R[j * 3]), FM(Rf[3], R[j * 3 + 1])), FM(Rf[6], R[j * 3 + 2]));
      float yw = FA(FA(FM(Rf[1], R[j * 3]), FM(Rf[4], R[j * 3 + 1])), FM(Rf[7], R[j * 3 + 2]));
      float zt = FA(FA(FM(v1t[0], R[j * 3]), FM(v1t[1], R[j * 3 + 1])), FM(v1t[2], R[j * 3 + 2]));
      float zh = FA(FA(FM(v1h[0], R[j * 3]), FM(v1h[1], R[j * 3 + 1])), FM(v1h[2], R[j * 3 + 2]));
      Fblk[(r0 + 4) * 12 + cM + j] = xw;
      Fblk[(r0 + 5) * 12 + cF + j] = zt;
      Fblk[(r0 + 5) * 12 + cM + j] = yw;
      Fblk[(r0 + 6) * 12 + cF + j] = zh;
      Fblk[(r0 + 6) * 12 + cM + j] = (leg == 0) ? -yw : yw;  // quirk Q5
    }
    Fblk[(r0 + 7) * 12 + cF + 2] = 2.f;
  }
}

// block argmin over (value, index); result broadcast through `red` (32 doubles followed by 32 ints)
__device__ inline void block_argmin(double v, int idx, double* red, double& vout, int& iout)
{
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  for (int o = 16; o > 0; o >>= 1) {
    double ov = __shfl_down_sync(0xffffffffu, v, o);
    int oi = __shfl_down_sync(0xffffffffu, idx, o);
    if (ov < v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
  int* redi = reinterpret_cast<int*>(red + 32);
  if (lane == 0) { red[wid] = v; redi[wid] = idx; }
  __syncthreads();
  if (wid == 0) {
    v = (lane < nw) ? red[lane] : 1e300;
    idx = (lane < nw) ? redi[lane] : 0x7fffffff;
    for (int o = 16; o > 0; o >>= 1) {
      double ov = __shfl_down_sync(0xffffffffu, v, o);
      int oi = __shfl_down_sync(0xffffffffu, idx, o);
      if (ov < v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    if (lane == 0) { red[31] = v; redi[31] = idx; }
  }
  __syncthreads();
  vout = red[31];
  iout = redi[31];
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) hmpc_solve_kernel(const KernelArgs ka)
{
  extern __shared__ __align__(16) unsigned char smem[];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int N = ka.horizon;
  const Layout L = make_layout(N, ka.nb_cap, ka.qmax, ka.rec_stride);
  const bool dump = (ka.dbg_H != nullptr);

  double* H = reinterpret_cast<double*>(smem + L.H);
  double* gq = reinterpret_cast<double*>(smem + L.gq);
  double* x0 = reinterpret_cast<double*>(smem + L.x0);
  double* xv = reinterpret_cast<double*>(smem + L.x);
  double* wv = reinterpret_cast<double*>(smem + L.w);
  double* HA = reinterpret_cast<double*>(smem + L.HA);
  double* nrm = reinterpret_cast<double*>(smem + L.nrm);  // [leg][type][6]
  double* rhs = reinterpret_cast<double*>(smem + L.rhs);  // [block*10 + type]
  int* blk_sl = reinterpret_cast<int*>(smem + L.blk);     // block -> step*2+leg
  int* sl_blk = blk_sl + ka.nb_cap;                       // step*2+leg -> block or -1
  double* red = reinterpret_cast<double*>(smem + L.misc);       // 32 doubles + 32 ints of reduction scratch
  int* flags = reinterpret_cast<int*>(smem + L.misc + 384);     // [0]=NB [1]=stance0 [2]=stance1 [3]=code [4]=decision [5]=drop
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L.misc + 448);

  double* Li = reinterpret_cast<double*>(smem + L.Li);
  double* lam = reinterpret_cast<double*>(smem + L.lam);
  double* dv = reinterpret_cast<double*>(smem + L.dv);
  double* yv = reinterpret_cast<double*>(smem + L.yv);
  double* rv = reinterpret_cast<double*>(smem + L.rv);
  int* Wc = reinterpret_cast<int*>(smem + L.Wc);
  unsigned char* act = smem + L.act;

  unsigned char* rec = smem + L.rec;
  const float* rf = reinterpret_cast<const float*>(rec);
  float* x0f = reinterpret_cast<float*>(smem + L.x0f);
  float* Acd = reinterpret_cast<float*>(smem + L.Acd);
  float* Bcd = reinterpret_cast<float*>(smem + L.Bcd);
  float* P = reinterpret_cast<float*>(smem + L.P);
  float* Mb = reinterpret_cast<float*>(smem + L.M);
  float* Tb = reinterpret_cast<float*>(smem + L.T);
  float* dd = reinterpret_cast<float*>(smem + L.dd);
  float* Fblk = reinterpret_cast<float*>(smem + L.fbl);

  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t phase = 0;

  for (int inst = blockIdx.x; inst < ka.batch; inst += gridDim.x) {
    // ---------------- stage 0: record -> shared memory (TMA bulk copy) ----------------
    if (tid == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic-proxy use of the union
      mbar_expect_tx(bar, (uint32_t)ka.rec_stride);
      bulk_g2s(rec, ka.records + (size_t)inst * ka.rec_stride, (uint32_t)ka.rec_stride, bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;

    // ---------------- contact table -> reduced block list (SolverMPC.cpp:589-637) ----------------
    const unsigned char* gait = rec + (54 + 12 * N) * 4;
    if (tid == 0) {
      int nb = 0;
      unsigned st0 = 0, st1 = 0;
      for (int s = 0; s < N; s++)
        for (int l = 0; l < 2; l++) {
          float ub = FM(ka.f_max, (float)gait[2 * s + l]);
          bool swing = (ub < 0.0001f && ub > -0.0001f) && !dump;  // near_zero(lb) && near_zero(ub); lb == 0
          if (swing) sl_blk[2 * s + l] = -1;
          else {
            sl_blk[2 * s + l] = (nb < ka.nb_cap) ? nb : -1;
            if (nb < ka.nb_cap) blk_sl[nb] = 2 * s + l;
            nb++;
            if (l == 0) st0 |= 1u << s; else st1 |= 1u << s;
          }
        }
      flags[0] = nb;
      flags[1] = (int)st0;
      flags[2] = (int)st1;
      flags[3] = ST_OK;
    }
    __syncthreads();
    const int NB = flags[0];
    if (!(NB > ka.nb_lo && NB <= ka.nb_hi)) {  // another launch's instance
      __syncthreads();
      continue;
    }
    const int n = 6 * NB, m = 10 * NB;
    const unsigned stmask[2] = {(unsigned)flags[1], (unsigned)flags[2]};

    // ---------------- stage 1: prologue ----------------
    if (tid == 0) prologue(rf, ka.dt, x0f, Acd, Bcd, Fblk);
    if (tid == 32 || (nt <= 32 && tid == 0)) {
      for (int i = 0; i < 169; i++) P[i] = (i / 13 == i % 13) ? 1.f : 0.f;
    }
    __syncthreads();

    // constraint normals (fp64 copies of the fp32 rows) and right-hand sides, "c'x >= d" form
    for (int e = tid; e < 2 * 10 * 6; e += nt) {
      int leg = e / 60, t = (e / 6) % 10, c = e % 6;
      int col = col12_of(leg, c);
      // one-sided rows in "c'x >= d" form: t0-3 friction (lower), t4/t5 Mx lower/upper, t6/t7 line
      // contact (upper), t8/t9 Fz lower/upper
      const int row = (t < 5) ? t : (t == 5 ? 4 : (t < 8 ? t - 1 : 7));
      const bool neg = (t == 5 || t == 6 || t == 7 || t == 9);
      float v = Fblk[(8 * leg + row) * 12 + col];
      nrm[e] = (double)(neg ? -v : v);
    }
    for (int e = tid; e < m; e += nt) {
      int k = e / 10, t = e % 10;
      int sl = blk_sl[k];
      double d = 0.0;
      if (t == 5) d = -(double)0.01f;
      if (t == 9) d = -(double)FM(ka.f_max, (float)gait[sl]);
      rhs[e] = d;
    }

    // ---------------- stage 2: powers of Acd, Toeplitz blocks ----------------
    for (int k = 1; k <= N; k++) {
      const float* Pp = P + (k - 1) * 169;
      float* Pn = P + k * 169;
      for (int e = tid; e < 169; e += nt) {
        int i = e / 13, j = e % 13;
        float acc = FM(Pp[i * 13], Acd[j]);
#pragma unroll
        for (int t = 1; t < 13; t++) acc = FA(acc, FM(Pp[i * 13 + t], Acd[t * 13 + j]));
        Pn[e] = acc;
      }
      __syncthreads();
    }
    // M_d = P_d * Bcd (rows 0..11 used), T_d = M_d .* w, dd = A_qp x0 - X_d
    for (int e = tid; e < N * 144; e += nt) {
      int d = e / 144, r = (e % 144) / 12, c = e % 12;
      const float* Pd = P + d * 169 + r * 13;
      float acc = FM(Pd[0], Bcd[c]);
#pragma unroll
      for (int t = 1; t < 13; t++) acc = FA(acc, FM(Pd[t], Bcd[t * 12 + c]));
      Mb[d * 156 + r * 12 + c] = acc;
      Tb[d * 144 + r * 12 + c] = FM(acc, rf[30 + r]);
    }
    for (int e = tid; e < N * 12; e += nt) {
      int s = e / 12, r = e % 12;
      const float* Ps = P + (s + 1) * 169 + r * 13;
      float acc = FM(Ps[0], x0f[0]);
#pragma unroll
      for (int t = 1; t < 13; t++) acc = FA(acc, FM(Ps[t], x0f[t]));
      dd[13 * s + r] = FS(acc, rf[54 + 12 * s + r]);
    }
    __syncthreads();

    // ---------------- stage 3: Hessian prefix chains + gradient ----------------
    if (dump) {
      float* oF = ka.dbg_F + (size_t)inst * 192;
      for (int e = tid; e < 192; e += nt) oF[e] = Fblk[e];
      float* olb = ka.dbg_lb + (size_t)inst * 16 * N;
      float* oub = ka.dbg_ub + (size_t)inst * 16 * N;
      for (int e = tid; e < 16 * N; e += nt) {
        int s = e / 16, r = e % 16, leg = r / 8, rr = r % 8;
        float lo = 0.f, hi = 0.f;
        if (rr < 4) hi = (float)5e10;
        else if (rr == 4) hi = 0.01f;
        else if (rr < 7) lo = (float)(-5e10);
        else hi = FM(ka.f_max, (float)gait[2 * s + leg]);
        olb[e] = lo;
        oub[e] = hi;
      }
    }
    for (int id = tid; id < N * 144; id += nt) {
      const int delta = id / 144, ii = (id % 144) / 12, jj = id % 12;
      if (delta == 0 && ii > jj) continue;
      const int li = leg_of(ii), lj = leg_of(jj);
      const unsigned need = stmask[li] & (stmask[lj] >> delta);  // bit a: entry (a,ii)-(a+delta,jj) wanted
      if (!need) continue;
      const int amin = __ffs(need) - 1;
      const int Kmax = N - 1 - delta - amin;
      const int ci = loc_of(ii), cj = loc_of(jj);
      const float alpha = (delta == 0 && ii == jj) ? rf[42 + ii] : 0.f;
      float acc = 0.f;
      for (int K = 0; K <= Kmax; K++) {
        const float* Tk = Tb + (K + delta) * 144 + ii;
        const float* Mk = Mb + K * 156 + jj;
#pragma unroll
        for (int r = 0; r < 12; r++) acc = FA(acc, FM(Tk[r * 12], Mk[r * 12]));
        const int a = N - 1 - K - delta, b = a + delta;
        if ((need >> a) & 1u) {
          const float hv = FM(2.f, FA(acc, alpha));  // qH = 2*(B'SB + Alpha_rep)
          if (dump) {
            float* oH = ka.dbg_H + (size_t)inst * (144 * N * N);
            oH[(size_t)(12 * a + ii) * (12 * N) + 12 * b + jj] = hv;
            oH[(size_t)(12 * b + jj) * (12 * N) + 12 * a + ii] = hv;
          } else {
            const int ka_ = sl_blk[2 * a + li], kb_ = sl_blk[2 * b + lj];
            const double hd = (double)hv;
            if (ka_ == kb_) {
              H[blk_off(ka_, ka_) + ci * 6 + cj] = hd;
              H[blk_off(ka_, ka_) + cj * 6 + ci] = hd;
            } else if (kb_ > ka_) H[blk_off(kb_, ka_) + cj * 6 + ci] = hd;
            else H[blk_off(ka_, kb_) + ci * 6 + cj] = hd;
          }
        }
      }
    }
    for (int e = tid; e < N * 12; e += nt) {
      const int a = e / 12, ii = e % 12, li = leg_of(ii);
      if (!((stmask[li] >> a) & 1u)) continue;
      float acc = 0.f;
      for (int s = a; s < N; s++) {
        const float* Tk = Tb + (s - a) * 144 + ii;
        const float* dk = dd + 13 * s;
#pragma unroll
        for (int r = 0; r < 12; r++) acc = FA(acc, FM(FM(Tk[r * 12], 2.f), dk[r]));
      }
      if (dump) ka.dbg_g[(size_t)inst * 12 * N + e] = acc;
      else gq[6 * sl_blk[2 * a + li] + loc_of(ii)] = (double)acc;
    }
    __syncthreads();
    if (dump) continue;

    float* out = ka.wrench + (size_t)inst * 12 * N;
    if (NB == 0) {
      for (int e = tid; e < 12 * N; e += nt) out[e] = 0.f;
      if (tid == 0) ka.status[inst] = ST_OK;
      __syncthreads();
      continue;
    }

    // ---------------- stage 4: sweep inversion, one 6x6 block per thread in registers ----------------
    // After sweeping every pivot the matrix holds -H^-1 (Goodnight's sweep operator on an SPD matrix).
    {
      const int nbt = NB * (NB + 1) / 2;
      const bool own = tid < nbt;
      int ib = 0, jb = 0;
      if (own) {
        ib = (int)((sqrtf(8.f * (float)tid + 1.f) - 1.f) * 0.5f);
        while ((ib + 1) * (ib + 2) / 2 <= tid) ib++;
        while (ib * (ib + 1) / 2 > tid) ib--;
        jb = tid - ib * (ib + 1) / 2;
      }
      double a[36];
      if (own) {
        const double* src = H + blk_off(ib, jb);
#pragma unroll
        for (int e = 0; e < 36; e++) a[e] = src[e];
      }
      double* colbuf[2] = {xv, wv};
      bool bad = false;
      for (int kb = 0; kb < NB; kb++) {
#pragma unroll
        for (int kk = 0; kk < 6; kk++) {
          double* col = colbuf[kk & 1];
          if (own) {
            if (jb == kb) {
#pragma unroll
              for (int r = 0; r < 6; r++) col[6 * ib + r] = a[r * 6 + kk];
            } else if (ib == kb) {
#pragma unroll
              for (int c = 0; c < 6; c++) col[6 * jb + c] = a[kk * 6 + c];
            }
          }
          __syncthreads();
          if (own) {
            const double d = col[6 * kb + kk];
            if (!(d > 0.0)) bad = true;
            const double inv = 1.0 / d;
            double ci[6], cj[6];
#pragma unroll
            for (int r = 0; r < 6; r++) ci[r] = col[6 * ib + r];
#pragma unroll
            for (int c = 0; c < 6; c++) cj[c] = col[6 * jb + c] * inv;
#pragma unroll
            for (int r = 0; r < 6; r++)
#pragma unroll
              for (int c = 0; c < 6; c++) a[r * 6 + c] = fma(-ci[r], cj[c], a[r * 6 + c]);
            if (jb == kb) {
#pragma unroll
              for (int r = 0; r < 6; r++) a[r * 6 + kk] = ci[r] * inv;
            }
            if (ib == kb) {
#pragma unroll
              for (int c = 0; c < 6; c++) a[kk * 6 + c] = cj[c];
              if (jb == kb) a[kk * 6 + kk] = -inv;
            }
          }
        }
      }
      if (own) {
        double* dst = H + blk_off(ib, jb);
#pragma unroll
        for (int e = 0; e < 36; e++) dst[e] = -a[e];
        if (bad) atomicExch(&flags[3], ST_NOT_SPD);
      }
      __syncthreads();
    }

    // ---------------- stage 5: dual active-set iterations ----------------
    // x0 = -H^-1 g
    for (int i = tid; i < n; i += nt) {
      const int ibk = i / 6, r = i % 6;
      double acc = 0.0;
      for (int jbk = 0; jbk < NB; jbk++) {
#pragma unroll
        for (int c = 0; c < 6; c++) acc = fma(hsym(H, ibk, r, jbk, c), gq[6 * jbk + c], acc);
      }
      x0[i] = -acc;
    }
    for (int e = tid; e < m; e += nt) act[e] = 0;
    __syncthreads();

    int q = 0, iters = 0;
    int code = flags[3];
    double xscale = 1.0;
    {
      double mx = 0.0;
      for (int i = tid; i < n; i += nt) mx = fmax(mx, fabs(x0[i]));
      int dummy;
      block_argmin(-mx, tid, red, mx, dummy);
      xscale = fmax(1.0, -mx);
    }
    const double tol = 1e-9 * xscale;

    while (code == ST_OK) {
      // w = A_W' lam ; x = x0 + H^-1 w
      for (int i = tid; i < n; i += nt) {
        const int k = i / 6, c = i % 6, leg = blk_sl[k] & 1;
        double acc = 0.0;
        for (int j = 0; j < q; j++) {
          const int cj = Wc[j];
          if (cj / 10 == k) acc = fma(lam[j], nrm[(leg * 10 + cj % 10) * 6 + c], acc);
        }
        wv[i] = acc;
      }
      __syncthreads();
      for (int i = tid; i < n; i += nt) {
        const int ibk = i / 6, r = i % 6;
        double acc = x0[i];
        for (int jbk = 0; jbk < NB; jbk++) {
          const double* wj = wv + 6 * jbk;
          if (wj[0] != 0.0 || wj[1] != 0.0 || wj[2] != 0.0 || wj[3] != 0.0 || wj[4] != 0.0 || wj[5] != 0.0) {
#pragma unroll
            for (int c = 0; c < 6; c++) acc = fma(hsym(H, ibk, r, jbk, c), wj[c], acc);
          }
        }
        xv[i] = acc;
      }
      __syncthreads();
      // most violated inactive constraint
      double sbest = 1e300;
      int pbest = 0x7fffffff;
      for (int e = tid; e < m; e += nt) {
        if (act[e]) continue;
        const int k = e / 10, t = e % 10, leg = blk_sl[k] & 1;
        const double* nn = nrm + (leg * 10 + t) * 6;
        const double* xb = xv + 6 * k;
        double s = -rhs[e];
#pragma unroll
        for (int c = 0; c < 6; c++) s = fma(nn[c], xb[c], s);
        if (s < sbest) { sbest = s; pbest = e; }
      }
      double sp;
      int p;
      block_argmin(sbest, pbest, red, sp, p);
      if (!(sp < -tol)) break;  // KKT point reached

      // ---- add constraint p (possibly after dropping blocking ones) ----
      const int kp = p / 10, tp = p % 10, legp = blk_sl[kp] & 1;
      const double* np_ = nrm + (legp * 10 + tp) * 6;
      double lam_p = 0.0;
      while (true) {
        iters++;
        if (iters > ka.max_iter) { code = ST_ITER_CAP; break; }
        // HA = H^-1 a_p
        for (int i = tid; i < n; i += nt) {
          const int ibk = i / 6, r = i % 6;
          double acc = 0.0;
#pragma unroll
          for (int c = 0; c < 6; c++) acc = fma(hsym(H, ibk, r, kp, c), np_[c], acc);
          HA[i] = acc;
        }
        __syncthreads();
        // warp 0: step direction in the dual space through the inverse Cholesky factor of the Schur complement
        if (tid < 32) {
          const int lane = tid;
          double cHc = 0.0;
#pragma unroll
          for (int c = 0; c < 6; c++) cHc = fma(np_[c], HA[6 * kp + c], cHc);
          for (int j = lane; j < q; j += 32) {
            const int cj = Wc[j], kj = cj / 10, lj = blk_sl[kj] & 1;
            const double* nj = nrm + (lj * 10 + cj % 10) * 6;
            double acc = 0.0;
#pragma unroll
            for (int c = 0; c < 6; c++) acc = fma(nj[c], HA[6 * kj + c], acc);
            dv[j] = acc;
          }
          __syncwarp();
          double yy = 0.0;
          for (int j = lane; j < q; j += 32) {
            const double* row = Li + j * (j + 1) / 2;
            double acc = 0.0;
            for (int i = 0; i <= j; i++) acc = fma(row[i], dv[i], acc);
            yv[j] = acc;
            yy = fma(acc, acc, yy);
          }
          for (int o = 16; o > 0; o >>= 1) yy += __shfl_xor_sync(0xffffffffu, yy, o);
          __syncwarp();
          const double zn = cHc - yy;
          const bool dependent = !(zn > 1e-11 * cHc);
          double t1 = 1e300;
          int l1 = 0x7fffffff;
          for (int i = lane; i < q; i += 32) {
            double acc = 0.0;
            for (int j = i; j < q; j++) acc = fma(Li[j * (j + 1) / 2 + i], yv[j], acc);
            rv[i] = acc;
            if (acc > 0.0) {
              const double ratio = lam[i] / acc;
              if (ratio < t1 || (ratio == t1 && i < l1)) { t1 = ratio; l1 = i; }
            }
          }
          for (int o = 16; o > 0; o >>= 1) {
            double ot = __shfl_xor_sync(0xffffffffu, t1, o);
            int ol = __shfl_xor_sync(0xffffffffu, l1, o);
            if (ot < t1 || (ot == t1 && ol < l1)) { t1 = ot; l1 = ol; }
          }
          __syncwarp();
          const double t2 = dependent ? 1e300 : fmax(0.0, -sp / zn);
          const double t = fmin(t1, t2);
          int decision;  // 0 = full step (p joins W), 1 = partial step (drop l1, retry), 2 = infeasible, 3 = W full
          if (!(t < 1e299)) decision = 2;
          else if (t2 <= t1) decision = (q >= ka.qmax) ? 3 : 0;
          else decision = 1;
          if (decision < 2) {
            for (int i = lane; i < q; i += 32) lam[i] = fmax(0.0, lam[i] - t * rv[i]);
          }
          __syncwarp();
          if (decision == 0) {
            // new row of the inverse factor: [-r'/rho, 1/rho], rho = sqrt(zn)
            const double rho = sqrt(zn), irho = 1.0 / rho;
            double* row = Li + q * (q + 1) / 2;
            for (int i = lane; i < q; i += 32) row[i] = -rv[i] * irho;
            if (lane == 0) {
              row[q] = irho;
              Wc[q] = p;
              lam[q] = lam_p + t;
              act[p] = 1;
            }
          } else if (decision == 1 && lane == 0) {
            lam[q] = lam_p + t;  // pending multiplier of p, parked behind the working set
            Wc[q] = p;
            flags[5] = l1;
          }
          if (lane == 0) flags[4] = decision;
        }
        __syncthreads();
        const int decision = flags[4];
        if (decision == 0) { q++; break; }
        if (decision == 2) { code = ST_INFEASIBLE; break; }
        if (decision == 3) { code = ST_WS_CAP; break; }
        // ---- partial step: drop working-set entry l, rebuild the inverse factor, refresh s_p ----
        {
          const int l = flags[5];
          lam_p = lam[q];
          __syncthreads();
          if (tid == 0) {
            act[Wc[l]] = 0;
            for (int j = l; j < q; j++) { Wc[j] = Wc[j + 1]; lam[j] = lam[j + 1]; }  // includes the parked p at q
          }
          q--;
          __syncthreads();
          // rebuild Li by appending the q remaining constraints one at a time (warp 0)
          if (tid < 32) {
            const int lane = tid;
            for (int jn = 0; jn < q; jn++) {
              const int cn = Wc[jn], kn = cn / 10, ln = blk_sl[kn] & 1;
              const double* nn = nrm + (ln * 10 + cn % 10) * 6;
              // S[jn][i] = a_n' H^-1 a_i, i <= jn
              for (int i = lane; i <= jn; i += 32) {
                const int ci_ = Wc[i], ki = ci_ / 10, li_ = blk_sl[ki] & 1;
                const double* ni = nrm + (li_ * 10 + ci_ % 10) * 6;
                double acc = 0.0;
#pragma unroll
                for (int r = 0; r < 6; r++) {
                  double hr = 0.0;
#pragma unroll
                  for (int c = 0; c < 6; c++) hr = fma(hsym(H, kn, r, ki, c), ni[c], hr);
                  acc = fma(nn[r], hr, acc);
                }
                dv[i] = acc;
              }
              __syncwarp();
              double yy = 0.0;
              for (int j = lane; j < jn; j += 32) {
                const double* row = Li + j * (j + 1) / 2;
                double acc = 0.0;
                for (int i = 0; i <= j; i++) acc = fma(row[i], dv[i], acc);
                yv[j] = acc;
                yy = fma(acc, acc, yy);
              }
              for (int o = 16; o > 0; o >>= 1) yy += __shfl_xor_sync(0xffffffffu, yy, o);
              __syncwarp();
              const double znn = dv[jn] - yy;
              const double irho = 1.0 / sqrt(fmax(znn, 1e-300));
              double* row = Li + jn * (jn + 1) / 2;
              for (int i = lane; i < jn; i += 32) {
                double acc = 0.0;
                for (int j = i; j < jn; j++) acc = fma(Li[j * (j + 1) / 2 + i], yv[j], acc);
                row[i] = -acc * irho;
              }
              if (lane == 0) row[jn] = irho;
              __syncwarp();
            }
          }
          __syncthreads();
          // x with the parked multiplier of p included -> refreshed slack of p
          for (int i = tid; i < n; i += nt) {
            const int k = i / 6, c = i % 6, leg = blk_sl[k] & 1;
            double acc = 0.0;
            for (int j = 0; j <= q; j++) {
              const int cj = Wc[j];
              if (cj / 10 == k) acc = fma(lam[j], nrm[(leg * 10 + cj % 10) * 6 + c], acc);
            }
            wv[i] = acc;
          }
          __syncthreads();
          for (int i = tid; i < n; i += nt) {
            const int ibk = i / 6, r = i % 6;
            double acc = x0[i];
            for (int jbk = 0; jbk < NB; jbk++) {
              const double* wj = wv + 6 * jbk;
#pragma unroll
              for (int c = 0; c < 6; c++) acc = fma(hsym(H, ibk, r, jbk, c), wj[c], acc);
            }
            xv[i] = acc;
          }
          __syncthreads();
          sp = -rhs[p];
#pragma unroll
          for (int c = 0; c < 6; c++) sp = fma(np_[c], xv[6 * kp + c], sp);
        }
      }
    }

    // ---------------- stage 6: scatter (eliminated variables are exactly 0) ----------------
    for (int e = tid; e < 12 * N; e += nt) {
      const int s = e / 12, c12 = e % 12, leg = leg_of(c12);
      const int k = sl_blk[2 * s + leg];
      out[e] = (k >= 0) ? (float)xv[6 * k + loc_of(c12)] : 0.f;
    }
    if (tid == 0) ka.status[inst] = (code & 0xff) | ((iters & 0xfff) << 8) | ((q & 0xff) << 20);
    __syncthreads();
  }
}

}  // namespace hmpc
