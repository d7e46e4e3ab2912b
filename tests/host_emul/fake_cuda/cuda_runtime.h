// tests/host_emul/fake_cuda/cuda_runtime.h — TEST INFRASTRUCTURE (CPU suite only).
//
// Lets g++ parse hector_simulation_b200/csrc/hmpc_device.cuh on a machine without a GPU so that the *source text*
// of the one-thread-per-robot kernels (data preparation, closed-loop advance, swing-leg controller) and of the stage-1
// device functions can be executed on the host and compared with the oracle in the CPU test suite.  This is NOT a CPU
// path of the product: it is compiled only by tests/test_kernel_source_on_host.py into a throw-away library under
// tests/host_emul/_build/, nothing in the package links or loads it.  One-thread-per-robot kernels run as plain loops;
// the cooperative kernels (solve, classification) run with ONE OS THREAD PER CUDA THREAD of a CTA: __syncthreads and the
// full-mask warp primitives (__syncwarp, __shfl*_sync, __ballot_sync, __reduce_*_sync) are blocking barriers plus an
// exchange array, shared memory is one process-wide buffer (one CTA at a time), the TMA bulk copy is a memcpy that
// completes an emulated mbarrier.  It is slow (tens of ms per QP) and exists to check results, not speed.
//
// Arithmetic: the CUDA `__f*_rn / __d*_rn` intrinsics are single IEEE operations that the compiler may not contract;
// here they are plain operators in a translation unit built with -ffp-contract=off and without -march (no FMA).
// Math functions resolve to glibc's double versions (the kernels call them on doubles only).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __shared__ static
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

struct hmpc_emul_dim3 { unsigned x, y, z; };
extern thread_local hmpc_emul_dim3 threadIdx, blockIdx, blockDim, gridDim;

inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline double __dsqrt_rn(double a) { return std::sqrt(a); }
inline double __drcp_rn(double a) { return 1.0 / a; }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
inline unsigned __float_as_uint(float f) { unsigned i; std::memcpy(&i, &f, 4); return i; }
inline float __uint_as_float(unsigned i) { float f; std::memcpy(&f, &i, 4); return f; }
inline long long __double_as_longlong(double d) { long long i; std::memcpy(&i, &d, 8); return i; }
inline double __longlong_as_double(long long i) { double d; std::memcpy(&d, &i, 8); return d; }
inline int __double2hiint(double d) { return (int)(__double_as_longlong(d) >> 32); }
inline int __double2loint(double d) { return (int)(__double_as_longlong(d) & 0xffffffffll); }
inline double __hiloint2double(int hi, int lo) { return __longlong_as_double(((long long)hi << 32) | (unsigned)lo); }

// ---- cooperative execution: one OS thread per CUDA thread of the CTA that is "resident" ---------------------------------
#include <sched.h>

#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <mutex>

namespace hmpc_emul {
struct Barrier {  // reusable blocking barrier
  std::mutex m;
  std::condition_variable cv;
  int count = 0, waiting = 0;
  unsigned gen = 0;
  void wait()
  {
    std::unique_lock<std::mutex> l(m);
    const unsigned g = gen;
    if (++waiting == count) {
      waiting = 0;
      gen++;
      cv.notify_all();
    } else {
      cv.wait(l, [&] { return gen != g; });
    }
  }
};
struct Warp {
  Barrier bar;
  unsigned long long slot[32];
};
struct Cta {
  Barrier bar;
  Warp warps[32];
  int orflag[2] = {0, 0};
  unsigned orphase = 0;
};
[[noreturn]] inline void die(const char* what)
{
  std::fprintf(stderr, "host emulation: %s\n", what);
  std::abort();
}
unsigned char* cta_smem();  // 256 KB, 16-byte aligned (kernel_source_on_host.cpp)
}  // namespace hmpc_emul
extern thread_local hmpc_emul::Cta* hmpc_emul_cta;  // null while a one-thread-per-robot kernel runs as a plain loop

inline hmpc_emul::Warp& hmpc_emul_warp()
{
  if (!hmpc_emul_cta) hmpc_emul::die("cooperative primitive outside an emulated CTA");
  return hmpc_emul_cta->warps[threadIdx.x >> 5];
}
inline void __syncthreads()
{
  if (!hmpc_emul_cta) hmpc_emul::die("__syncthreads outside an emulated CTA");
  hmpc_emul_cta->bar.wait();
}
inline int __syncthreads_or(int pred)
{
  if (!hmpc_emul_cta) hmpc_emul::die("__syncthreads_or outside an emulated CTA");
  __atomic_fetch_or(&hmpc_emul_cta->orflag[hmpc_emul_cta->orphase & 1], pred, __ATOMIC_SEQ_CST);
  hmpc_emul_cta->bar.wait();
  const int r = __atomic_load_n(&hmpc_emul_cta->orflag[hmpc_emul_cta->orphase & 1], __ATOMIC_SEQ_CST);
  hmpc_emul_cta->bar.wait();
  if (threadIdx.x == 0) { hmpc_emul_cta->orflag[hmpc_emul_cta->orphase & 1] = 0; hmpc_emul_cta->orphase++; }
  hmpc_emul_cta->bar.wait();
  return r;
}
inline void __syncwarp(unsigned mask = 0xffffffffu)
{
  if (mask != 0xffffffffu) hmpc_emul::die("partial-mask __syncwarp is not emulated");
  hmpc_emul_warp().bar.wait();
}
// all-lanes exchange: every lane deposits 8 bytes, reads what it needs after a barrier, and a second barrier keeps the
// slots intact until everybody has read
template <class T, class F>
inline T hmpc_emul_exchange(unsigned mask, T v, F&& pick)
{
  static_assert(sizeof(T) <= 8, "exchange of up to 8 bytes");
  if (mask != 0xffffffffu) hmpc_emul::die("partial-mask warp primitive is not emulated");
  hmpc_emul::Warp& w = hmpc_emul_warp();
  unsigned long long bits = 0;
  std::memcpy(&bits, &v, sizeof(T));
  w.slot[threadIdx.x & 31] = bits;
  w.bar.wait();
  const T r = pick(w.slot);
  w.bar.wait();
  return r;
}
template <class T>
inline T hmpc_emul_from_slot(unsigned long long bits)
{
  T r;
  std::memcpy(&r, &bits, sizeof(T));
  return r;
}
template <class T>
inline T __shfl_sync(unsigned mask, T v, int src, int = 32)
{
  return hmpc_emul_exchange(mask, v, [&](const unsigned long long* s) { return hmpc_emul_from_slot<T>(s[src & 31]); });
}
template <class T>
inline T __shfl_xor_sync(unsigned mask, T v, int lane_mask, int = 32)
{
  const int src = (threadIdx.x & 31) ^ lane_mask;
  return hmpc_emul_exchange(mask, v, [&](const unsigned long long* s) { return hmpc_emul_from_slot<T>(s[src & 31]); });
}
inline unsigned __ballot_sync(unsigned mask, int pred)
{
  return hmpc_emul_exchange(mask, (unsigned)(pred != 0), [&](const unsigned long long* s) {
    unsigned m = 0;
    for (int l = 0; l < 32; l++) m |= (unsigned)(s[l] & 1u) << l;
    return m;
  });
}
// group masks are supported for the unsigned minimum: every lane of the warp executes the call (convergent code), each
// with the mask of its own group, and reduces over the lanes its mask names
inline unsigned __reduce_min_sync(unsigned mask, unsigned v)
{
  if (!((mask >> (threadIdx.x & 31)) & 1u)) hmpc_emul::die("__reduce_min_sync: the caller is not in its mask");
  return hmpc_emul_exchange(0xffffffffu, v, [&](const unsigned long long* s) {
    unsigned m = 0xffffffffu;
    for (int l = 0; l < 32; l++)
      if ((mask >> l) & 1u) m = (unsigned)s[l] < m ? (unsigned)s[l] : m;
    return m;
  });
}
inline int __reduce_min_sync(unsigned mask, int v)
{
  return hmpc_emul_exchange(mask, v, [&](const unsigned long long* s) {
    int m = hmpc_emul_from_slot<int>(s[0]);
    for (int l = 1; l < 32; l++) {
      const int o = hmpc_emul_from_slot<int>(s[l]);
      m = o < m ? o : m;
    }
    return m;
  });
}
inline unsigned __reduce_max_sync(unsigned mask, unsigned v)
{
  return hmpc_emul_exchange(mask, v, [&](const unsigned long long* s) {
    unsigned m = (unsigned)s[0];
    for (int l = 1; l < 32; l++) m = (unsigned)s[l] > m ? (unsigned)s[l] : m;
    return m;
  });
}
inline unsigned __reduce_or_sync(unsigned mask, unsigned v)
{
  return hmpc_emul_exchange(mask, v, [&](const unsigned long long* s) {
    unsigned m = 0u;
    for (int l = 0; l < 32; l++) m |= (unsigned)s[l];
    return m;
  });
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0u; }
// mma.sync.aligned.m8n8k4.row.col.f64 (the kernel's dmma884 body is replaced by a call to this, test build step):
// lane 4g+t holds A[g][t], B[t][g] and C[g][2t], C[g][2t+1]; the k-sum runs in index order with fused multiply-adds
inline void hmpc_emul_dmma884(double& c0, double& c1, double a, double b)
{
  hmpc_emul::Warp& w = hmpc_emul_warp();
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  double A[4], B0[4], B1[4];
  unsigned long long bits;
  std::memcpy(&bits, &a, 8);
  w.slot[lane] = bits;
  w.bar.wait();
  for (int k = 0; k < 4; k++) std::memcpy(&A[k], &w.slot[4 * g + k], 8);
  w.bar.wait();
  std::memcpy(&bits, &b, 8);
  w.slot[lane] = bits;
  w.bar.wait();
  for (int k = 0; k < 4; k++) {
    std::memcpy(&B0[k], &w.slot[4 * (2 * t) + k], 8);
    std::memcpy(&B1[k], &w.slot[4 * (2 * t + 1) + k], 8);
  }
  w.bar.wait();
  for (int k = 0; k < 4; k++) {
    c0 = std::fma(A[k], B0[k], c0);
    c1 = std::fma(A[k], B1[k], c1);
  }
}
inline unsigned atomicOr(unsigned* p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicAnd(unsigned* p, unsigned v) { return __atomic_fetch_and(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline long long clock64() { return 0; }
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
inline unsigned __brev(unsigned x)
{
  unsigned r = 0;
  for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i);
  return r;
}
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __trap() { hmpc_emul::die("__trap"); }
template <class T> inline T __ldg(const T* p) { return *p; }
// mbarrier + 1-D bulk copy (TMA): the kernel's PTX helpers are replaced by calls to these (test build step)
inline void hmpc_emul_mbar_init(uint64_t* bar) { __atomic_store_n(bar, (uint64_t)0, __ATOMIC_SEQ_CST); }
inline void hmpc_emul_mbar_wait(uint64_t* bar, uint32_t phase)
{
  while ((__atomic_load_n(bar, __ATOMIC_ACQUIRE) & 1u) == phase) sched_yield();  // phase `phase` not yet complete
}
inline void hmpc_emul_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
  std::memcpy(dst, src, bytes);
  __atomic_fetch_add(bar, (uint64_t)1, __ATOMIC_RELEASE);  // the expected bytes have arrived: the phase completes
}
inline unsigned long long __cvta_generic_to_shared(const void*) { hmpc_emul::die("__cvta_generic_to_shared"); }
struct double2 { double x, y; };
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
inline double2 make_double2(double x, double y) { return double2{x, y}; }
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
inline int2 make_int2(int x, int y) { return int2{x, y}; }
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
using std::fmax;
using std::fmin;
using std::fabs;
