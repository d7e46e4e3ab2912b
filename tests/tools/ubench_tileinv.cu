// tests/tools/ubench_tileinv.cu — developer microbenchmark (B200): latency of the in-register 8x8 SPD tile inversion that
// is the critical path of stage 4's block step, alone and with the seven CTAs of an SM doing it at once.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ubench_tileinv tests/tools/ubench_tileinv.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int NEWTON>
__device__ __forceinline__ double rcp_n(double x)
{
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
#pragma unroll
  for (int i = 0; i < NEWTON; i++) {
    const double e = fma(-x, r, 1.0);
    r = fma(e, r, r);
  }
  return r;
}

// the kernel's routine (hmpc_device.cuh tile_inverse_spd), NEWTON = Newton steps per reciprocal, PREDICT = next pivot's
// reciprocal computed while the current one is applied
template <int NEWTON, bool PREDICT>
__device__ __forceinline__ bool tile_inverse(double& a0, double& a1, int lane)
{
  const int g = lane >> 2, t4 = lane & 3;
  bool bad = false;
  double inv = rcp_n<NEWTON>(__shfl_sync(0xffffffffu, a0, 0));
#pragma unroll
  for (int p = 0; p < 8; p++) {
    const double selp = (p & 1) ? a1 : a0;
    const double colp = __shfl_sync(0xffffffffu, selp, 4 * g + (p >> 1));
    const double r0 = __shfl_sync(0xffffffffu, a0, 4 * p + t4);
    const double r1 = __shfl_sync(0xffffffffu, a1, 4 * p + t4);
    const double d = __shfl_sync(0xffffffffu, selp, 4 * p + (p >> 1));
    bad |= !(d > 0.0);
    if (!PREDICT) inv = rcp_n<NEWTON>(d);
    double invn = 0.0;
    if (PREDICT && p < 7) {
      const double e = __shfl_sync(0xffffffffu, selp, 4 * (p + 1) + (p >> 1));
      const double dn0 = __shfl_sync(0xffffffffu, ((p + 1) & 1) ? a1 : a0, 4 * (p + 1) + ((p + 1) >> 1));
      invn = rcp_n<NEWTON>(fma(-e * inv, e, dn0));
    }
    const double f = colp * inv;
    double n0 = fma(-f, r0, a0), n1 = fma(-f, r1, a1);
    if (g == p) { n0 = r0 * inv; n1 = r1 * inv; }
    if (t4 == (p >> 1)) {
      const double pc = (g == p) ? -inv : f;
      if (p & 1) n1 = pc; else n0 = pc;
    }
    a0 = n0;
    a1 = n1;
    if (PREDICT) inv = invn;
  }
  a0 = -a0;
  a1 = -a1;
  return bad;
}

// `which` = the warp of the CTA that works (-1: all warps; -2: warp (blockIdx.x / nsm) % 4, a different one in every CTA
// of an SM if CTAs are dealt round-robin); the others wait at the barrier like the kernel's other warps
template <int NEWTON, bool PREDICT>
__global__ void k_inv(double* out, long long* clk, int iters, int which, int nsm)
{
  extern __shared__ unsigned char pad[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  // SPD tile: 4 I + small symmetric part
  double a0 = (g == 2 * t4 ? 4.0 : 0.0) + 0.01 * (g + 2 * t4 + 1), a1 = (g == 2 * t4 + 1 ? 4.0 : 0.0) + 0.01 * (g + 2 * t4 + 2);
  const int w = which == -2 ? (blockIdx.x / nsm) % 4 : which;
  bool bad = false;
  __syncthreads();
  const long long t0 = clock64();
  if (which == -1 || wid == w)
    for (int i = 0; i < iters; i++) bad |= tile_inverse<NEWTON, PREDICT>(a0, a1, lane);  // inverse of the inverse: stays SPD
  const long long t1 = clock64();
  __syncthreads();
  if (which == -1 || wid == w) {
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + (bad ? 1.0 : 0.0);
    if (lane == 0) clk[blockIdx.x] = t1 - t0;
  }
  if (pad[0] == 77) out[0] = 1.0;
}

int main()
{
  double* d;
  long long* clk;
  const int NB = 148 * 7;
  CK(cudaMalloc(&d, sizeof(double) * NB * 128));
  CK(cudaMalloc(&clk, sizeof(long long) * NB));
  long long* h = (long long*)malloc(sizeof(long long) * NB);
  const int IT = 64;
  auto report = [&](const char* name, int nb) {
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, clk, sizeof(long long) * nb, cudaMemcpyDeviceToHost));
    double s = 0, mx = 0;
    for (int i = 0; i < nb; i++) { s += (double)h[i]; if ((double)h[i] > mx) mx = (double)h[i]; }
    printf("%-78s: mean %.0f cyc, max %.0f cyc per inversion\n", name, s / nb / IT, mx / IT);
  };
#define RUN(NEWTON, PREDICT, label)                                                                                   \
  CK(cudaFuncSetAttribute(k_inv<NEWTON, PREDICT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32000));               \
  CK(cudaMemset(clk, 0, sizeof(long long) * NB));                                                                     \
  k_inv<NEWTON, PREDICT><<<1, 32, 32000>>>(d, clk, IT, 0, 148); report(label ", one warp alone", 1);                    \
  k_inv<NEWTON, PREDICT><<<1, 128, 32000>>>(d, clk, IT, -1, 148); report(label ", 4 warps of one CTA (one per scheduler)", 1); \
  k_inv<NEWTON, PREDICT><<<NB, 128, 32000>>>(d, clk, IT, 2, 148); report(label ", 7 CTAs/SM, warp 2 of every CTA", NB);  \
  k_inv<NEWTON, PREDICT><<<NB, 128, 32000>>>(d, clk, IT, -2, 148); report(label ", 7 CTAs/SM, warp (block/148)%4", NB);
  RUN(2, true, "2 Newton steps, predicted reciprocal (the kernel's)")
  RUN(1, true, "1 Newton step, predicted reciprocal")
  RUN(2, false, "2 Newton steps, reciprocal of the shuffled pivot")
  RUN(0, true, "seed only (accuracy 1e-6: timing reference, not usable)")
  return 0;
}
