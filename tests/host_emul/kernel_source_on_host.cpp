// tests/host_emul/kernel_source_on_host.cpp — TEST INFRASTRUCTURE (CPU suite only), see fake_cuda/cuda_runtime.h.
//
// Includes the product's device header through the fake CUDA prelude and runs, on the host,
//   * the one-thread-per-robot kernels  hmpc_prepare_kernel (row f-1), hmpc_advance_kernel (f-3), hmpc_swing_kernel (f-4)
//     — one call per robot with threadIdx/blockIdx set the way a launch would set them;
//   * the stage-1 device functions of the solve kernel (role_leg / role_state / role_inertia: SRBD linearisation, foot
//     rotations, constraint rows) and the joint-torque epilogue function (leg_torque, f-2).
//   * the classification kernel and the solve kernel itself (all stages), one OS thread per CUDA thread of one CTA at a
//     time, launched the way hmpc_capi.cu's enqueue_solve launches them (classification -> class 0 -> class 1 -> class 2
//     with working-set-overflow escalation through the lists)
// so that tests/test_kernel_source_on_host.py can hold the kernels' SOURCE to the oracle / to the reference's compiled
// controller without a GPU.
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

// HMPC_DEVICE_HEADER: hector_simulation_b200/csrc/hmpc_device.cuh after the test's build step (test_kernel_source_on_host.py,
// `_host_buildable`) has made the handful of substitutions a host compiler needs, each asserted to hit exactly once:
//   - the four PTX helper bodies (mbar_init / mbar_expect_tx / mbar_wait / bulk_g2s) -> calls to hmpc_emul_* (cuda_runtime.h)
//   - the `rcp.approx.ftz.f64` seed of fast_rcp -> `1.0 / x` (the Newton steps that follow are kept)
//   - the dynamic shared-memory declaration -> a pointer to the emulated CTA's buffer
//   - the body of dmma884 (mma.sync.m8n8k4.f64) -> hmpc_emul_dmma884, a warp-collective exchange + fused multiply-adds
//   - every other `asm volatile(...)` (fences, griddepcontrol: no arithmetic) -> nothing
// Everything else is the product's source, byte for byte.
#include HMPC_DEVICE_HEADER

thread_local hmpc_emul_dim3 threadIdx, blockIdx, blockDim, gridDim;
thread_local hmpc_emul::Cta* hmpc_emul_cta = nullptr;

namespace hmpc_emul {
unsigned char* cta_smem()
{
  alignas(16) static unsigned char buf[256 * 1024];
  return buf;
}
}  // namespace hmpc_emul

namespace {
// run `fn()` on NT OS threads as the NT CUDA threads of the single CTA of a 1-CTA grid
template <class F>
void run_cta(int NT, F fn)
{
  hmpc_emul::Cta* cta = new hmpc_emul::Cta;
  cta->bar.count = NT;
  for (int w = 0; w < 32; w++) cta->warps[w].bar.count = 32;
  std::vector<std::thread> th;
  th.reserve(NT);
  for (int t = 0; t < NT; t++)
    th.emplace_back([=] {
      threadIdx = {(unsigned)t, 0, 0};
      blockIdx = {0, 0, 0};
      blockDim = {(unsigned)NT, 1, 1};
      gridDim = {1, 1, 1};
      hmpc_emul_cta = cta;
      fn();
      hmpc_emul_cta = nullptr;
    });
  for (auto& x : th) x.join();
  delete cta;
}

// the kernel variants of hmpc_capi.cu (HMPC_FOR_VARIANT): <threads, min CTAs/SM, fixed horizon, class>
const int kBucketThreads[5] = {64, 128, 192, 256, 384};
int variant_threads(int variant) { return variant == 0 ? 128 : (variant == 1 ? 256 : kBucketThreads[(variant - 10) % 5]); }
void launch_variant(int variant, const hmpc::KernelArgs& ka)
{
  switch (variant) {
    case 0: run_cta(128, [=] { hmpc::hmpc_solve_kernel<128, 7, 10, 0>(ka); }); break;
    case 1: run_cta(256, [=] { hmpc::hmpc_solve_kernel<256, 2, 10, 1>(ka); }); break;
    case 10: run_cta(64, [=] { hmpc::hmpc_solve_kernel<64, 8, 0, 0>(ka); }); break;
    case 11: run_cta(128, [=] { hmpc::hmpc_solve_kernel<128, 6, 0, 0>(ka); }); break;
    case 12: run_cta(192, [=] { hmpc::hmpc_solve_kernel<192, 3, 0, 0>(ka); }); break;
    case 13: run_cta(256, [=] { hmpc::hmpc_solve_kernel<256, 2, 0, 0>(ka); }); break;
    case 14: run_cta(384, [=] { hmpc::hmpc_solve_kernel<384, 1, 0, 0>(ka); }); break;
    case 15: run_cta(64, [=] { hmpc::hmpc_solve_kernel<64, 8, 0, 1>(ka); }); break;
    case 16: run_cta(128, [=] { hmpc::hmpc_solve_kernel<128, 6, 0, 1>(ka); }); break;
    case 17: run_cta(192, [=] { hmpc::hmpc_solve_kernel<192, 3, 0, 1>(ka); }); break;
    case 18: run_cta(256, [=] { hmpc::hmpc_solve_kernel<256, 2, 0, 1>(ka); }); break;
    default: run_cta(384, [=] { hmpc::hmpc_solve_kernel<384, 1, 0, 1>(ka); }); break;
  }
}
struct ClassCfg {
  int variant, nb_cap, qmax, tcap;
  hmpc::Layout L;
};
// hmpc_capi.cu build_classes, minus the CUDA occupancy calls
int build_classes(int N, ClassCfg* cls)
{
  const int rs = hmpc::record_stride(N);
  int ncls = 3;
  for (int i = 0; i < 2; i++) {
    ClassCfg& k = cls[i];
    k.nb_cap = hmpc::class_nb_cap(N, i);
    const int warps = hmpc::class_warps(N, i);
    int bucket = 0;
    while (bucket < 4 && kBucketThreads[bucket] < 32 * warps) bucket++;
    k.variant = (N == 10) ? i : 10 + 5 * i + bucket;
    k.qmax = hmpc::class_qmax(N, i);
    k.tcap = hmpc::class_tcap(N, i);
    k.L = hmpc::make_layout(N, k.nb_cap, k.qmax, rs, variant_threads(k.variant) / 32, k.tcap);
  }
  {
    ClassCfg& k = cls[2];
    k = cls[1];
    if (N == 10) k.variant = 18;
    const int n = 6 * k.nb_cap, nw = variant_threads(k.variant) / 32;
    k.qmax = n;
    k.tcap = 0;
    k.L = hmpc::make_layout(N, k.nb_cap, k.qmax, rs, nw, 0);
    while (k.L.total > 226 * 1024 && k.qmax > cls[1].qmax) {
      k.qmax -= 4;
      k.L = hmpc::make_layout(N, k.nb_cap, k.qmax, rs, nw, 0);
    }
    if (k.qmax <= cls[1].qmax || cls[1].qmax >= n) ncls = 2;
  }
  return ncls;
}
const unsigned kBlock = 128;
inline void set_thread(int i)
{
  blockDim = {kBlock, 1, 1};
  blockIdx = {(unsigned)i / kBlock, 0, 0};
  threadIdx = {(unsigned)i % kBlock, 0, 0};
  gridDim = {1u << 20, 1, 1};
}
}  // namespace

namespace {
int* g_ws = nullptr;  // [B][WS_STATE_INTS] working sets kept between emul_solve* calls (closed-loop warm start), or null
int g_ws_shift = 0;
}  // namespace

extern "C" {

/* working-set memory for the following emul_solve* calls: written back by every solve, proposed to the next one when that
 * call passes warm_start = 1 (shift = MPC steps the horizon moved in between) */
void emul_set_ws(int* ws, int shift)
{
  g_ws = ws;
  g_ws_shift = shift;
}
int emul_ws_ints() { return hmpc::WS_STATE_INTS; }

int emul_record_stride(int N) { return hmpc::record_stride(N); }

void emul_prepare(const unsigned char* states, int batch, int N, double dtMPC, unsigned char* records)
{
  for (int i = 0; i < batch; i++) {
    set_thread(i);
    hmpc::hmpc_prepare_kernel(states, batch, N, dtMPC, records, hmpc::record_stride(N));
  }
}

void emul_advance(unsigned char* states, unsigned char* loop, int batch, int N, double dtMPC, const float* wrench,
                  const int* status)
{
  for (int i = 0; i < batch; i++) {
    set_thread(i);
    hmpc::hmpc_advance_kernel(states, loop, batch, N, dtMPC, wrench, status, nullptr);
  }
}

void emul_swing(const unsigned char* states, const unsigned char* loop, const double* phase, unsigned char* swing, int batch,
                int n_iterations, double dt, double dtSwing, unsigned char* cmd)
{
  for (int i = 0; i < batch; i++) {
    set_thread(i);
    hmpc::hmpc_swing_kernel(states, loop, phase, swing, batch, n_iterations, dt, dtSwing, cmd);
  }
}

/* stage 1 of the solve kernel on one packed record's floats `rf` (layout: hmpc_device.cuh, "record floats"):
 * Fblk [192] (both legs' constraint rows), x0 [13], Acd [169], Bcd [156] — pre-zeroed as the kernel does. */
void emul_stage1(const float* rf, float dt, float* Fblk, float* x0, float* Acd, float* Bcd)
{
  memset(Fblk, 0, 192 * sizeof(float));
  memset(Acd, 0, 169 * sizeof(float));
  memset(Bcd, 0, 156 * sizeof(float));
  static float x0f[16];
  alignas(16) static unsigned char scr[512];
  memset(x0f, 0, sizeof(x0f));
  // the roles are warp-cooperative (libm calls spread over lanes): one emulated warp runs them the way stage 1 does
  run_cta(32, [=] {
    const int lane = (int)threadIdx.x;
    hmpc::role_leg(rf, lane, Fblk, scr);
    hmpc::role_state(rf, dt, x0f, Acd, lane, scr + 256);
    if (lane == 31) hmpc::role_inertia(rf, dt, Bcd);
  });
  memcpy(x0, x0f, 13 * sizeof(float));
}

/* The solve paths of hmpc_capi.cu on B <= 1024 robots.
 *   records != NULL: the device-resident path (enqueue_solve): classification kernel, then one launch per size class with
 *                    escalation lists, reading packed records.
 *   raw != NULL    : the in-place host-buffer path (enqueue_solve_hostlists): class lists built on the host from the contact
 *                    tables, kernels gather the live bytes of the caller's update_data_t records (3016-byte stride), no
 *                    escalation (the caller retries overflow through the first path).
 * Outputs: wrench [B][12N] floats and/or wrench64 [B][12N] doubles, status [B], tau [B][10] or NULL, launched[3] =
 * instances each class processed (NULL to skip).  warm_start: KernelArgs::warm_start.  Optional assembly dump (all NULL,
 * or all given): H [B][n*n], g [B][n], Fblk [B][192], lb/ub [B][16N]. */
int emul_solve_ex(const unsigned char* records, const unsigned char* raw, int B, int N, float dt, float f_max, int max_iter,
                  int warm_start, float* wrench, double* wrench64, int* status, float* tau, int* launched, float* dH, float* dg,
                  float* dF, float* dlb, float* dub)
{
  if (B < 1 || B > 1024 || (!records && !raw)) return 1;
  ClassCfg cls[3];
  int ncls = build_classes(N, cls);
  std::vector<int> block(8 + 3 * (size_t)B, 0);
  int* counts = block.data();      // [4] list lengths (+ [4] the next call's, which the class-0 launch clears)
  int* lists = counts + 8;
  const int rs = hmpc::record_stride(N);
  const int nb_hi0 = cls[0].nb_cap;
  if (raw) {
    // hmpc_capi.cu classify_host: gait bytes at offsetof(update_data_t, gait) = 1944
    for (int i = 0; i < B; i++) {
      int k = 0;
      for (int e = 0; e < 2 * N; e++) {
        const float ub = f_max * (float)raw[(size_t)i * 3016 + 1944 + e];
        k += !(ub < 0.0001f && ub > -0.0001f);
      }
      const int cl = (k <= nb_hi0) ? 0 : 1;
      lists[(size_t)cl * B + counts[cl]++] = i;
    }
    ncls = 2;
  } else {
    counts[0] = B;       // device-resident path: class 0 runs over every instance and classifies on the way
    counts[4] = counts[5] = counts[6] = counts[7] = 0x55;  // must come back cleared
  }
  for (int i = 0; i < ncls; i++) {
    if (launched) launched[i] = counts[i];
    if (counts[i] == 0) continue;
    const bool fused = !raw && i == 0;
    hmpc::KernelArgs ka{};
    ka.records = records;
    ka.raw_records = raw;
    ka.rec_stride = rs;
    ka.batch = B;
    ka.horizon = N;
    ka.dt = dt;
    ka.f_max = f_max;
    ka.max_iter = max_iter;
    ka.tol_kkt = getenv("HMPC_TOL_KKT") ? atof(getenv("HMPC_TOL_KKT")) : 1e-9;   // hmpc_capi.cu's constants; the
    ka.tol_dep = getenv("HMPC_TOL_DEP") ? atof(getenv("HMPC_TOL_DEP")) : 1e-11;  // tolerance-sweep test overrides them
    ka.kappa_max = getenv("HMPC_KAPPA_MAX") ? atof(getenv("HMPC_KAPPA_MAX")) : 1.5e5;  // hmpc_capi.cu's default
    ka.block_min = getenv("HMPC_BLOCK_MIN") ? atoi(getenv("HMPC_BLOCK_MIN")) : 2;  // hmpc_capi.cu's default
    ka.block_rounds = getenv("HMPC_BLOCK_ROUNDS") ? atoi(getenv("HMPC_BLOCK_ROUNDS")) : 4;  // hmpc_capi.cu's default
    ka.wrench = wrench;
    ka.wrench64 = wrench64;
    ka.status = status;
    ka.tau = tau;
    ka.warm_start = (g_ws && warm_start) ? 1 : 0;
    ka.ws_state = g_ws;
    ka.ws_shift = g_ws_shift;
    ka.list = fused ? nullptr : lists + (size_t)i * B;
    ka.split_nb = fused ? nb_hi0 : -1;
    ka.counts_next = fused ? counts + 4 : nullptr;
    ka.wave_sync = fused ? reinterpret_cast<unsigned*>(counts + 3) : nullptr;  // one emulated CTA walks all instances: B waves
    ka.counts = counts;
    ka.cls = i;
    ka.esc_list = (!raw && i + 1 < ncls) ? lists + (size_t)(i + 1) * B : nullptr;
    ka.nb_cap = cls[i].nb_cap;
    ka.qmax = cls[i].qmax;
    ka.tcap = cls[i].tcap;
    ka.L = cls[i].L;
    ka.dbg_H = dH; ka.dbg_g = dg; ka.dbg_F = dF; ka.dbg_lb = dlb; ka.dbg_ub = dub;
    launch_variant(cls[i].variant, ka);
    if (fused) {
      if (counts[4] | counts[5] | counts[6] | counts[7]) return 2;  // the next call's lengths were not cleared
      if (launched) launched[0] = B - counts[1];                     // instances class 0 kept
    }
  }
  return 0;
}

int emul_solve(const unsigned char* records, int B, int N, float dt, float f_max, int max_iter, float* wrench, int* status,
               float* tau, int* launched, float* dH, float* dg, float* dF, float* dlb, float* dub)
{
  return emul_solve_ex(records, nullptr, B, N, dt, f_max, max_iter, 0, wrench, nullptr, status, tau, launched, dH, dg, dF, dlb, dub);
}

/* the torque epilogue's per-joint function: tau_j = column j of J_force_moment(q5, leg) . f6 */
double emul_leg_torque(const double* q5, int leg, int j, const double* f6) { return hmpc::leg_torque(q5, leg, j, f6); }

}  // extern "C"
