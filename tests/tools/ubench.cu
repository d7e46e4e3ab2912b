// tests/tools/ubench.cu — developer microbenchmarks (B200): latencies that shape the solve kernel's critical path.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ubench tests/tools/ubench.cu && gpurun_out/ubench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void k_dfma_lat(double* out, long long* clk, int iters)
{
  double a = out[0], b = out[1], c = out[2];
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) c = fma(a, c, b);
  }
  long long t1 = clock64();
  out[3] = c;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void k_ffma_lat(float* out, long long* clk, int iters)
{
  float a = out[0], b = out[1], c = out[2];
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) c = __fadd_rn(__fmul_rn(a, c), b);
  }
  long long t1 = clock64();
  out[3] = c;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}
// DMMA: NACC independent accumulator pairs per warp; NACC = 1 -> dependent-chain latency
template <int NACC>
__global__ void k_dmma(double* out, long long* clk, int iters)
{
  double c0[NACC], c1[NACC];
  for (int i = 0; i < NACC; i++) { c0[i] = out[i]; c1[i] = out[i + 1]; }
  double a = out[threadIdx.x & 7], b = out[(threadIdx.x >> 2) & 7];
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < NACC; u++) dmma884(c0[u], c1[u], a, b);
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < NACC; i++) s += c0[i] + c1[i];
  out[64 + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
// DFMA throughput: NACC independent chains per thread
template <int NACC>
__global__ void k_dfma_tp(double* out, long long* clk, int iters)
{
  double c[NACC];
  for (int i = 0; i < NACC; i++) c[i] = out[i];
  double a = out[8], b = out[9];
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < NACC; u++) c[u] = fma(a, c[u], b);
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < NACC; i++) s += c[i];
  out[64 + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
__global__ void k_sync(long long* clk, int iters)
{
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) __syncthreads();
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void k_shfl(double* out, long long* clk, int iters)
{
  double v = out[threadIdx.x & 31];
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) v = __shfl_sync(0xffffffffu, v, (threadIdx.x + 1) & 31) + 1.0;
  }
  long long t1 = clock64();
  out[64 + threadIdx.x] = v;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void k_shfl32(float* out, long long* clk, int iters)
{
  float v = out[threadIdx.x & 31];
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) v = __shfl_sync(0xffffffffu, v, (threadIdx.x + 1) & 31);
  }
  long long t1 = clock64();
  out[64 + threadIdx.x] = v;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void k_lds(int* out, long long* clk, int iters)
{
  __shared__ int chain[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) chain[i] = (i * 33 + 17) & 1023;
  __syncthreads();
  int p = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) p = chain[p];
  }
  long long t1 = clock64();
  out[threadIdx.x] = p;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__device__ __forceinline__ double fast_rcp(double x)
{
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(e, r, r);
  e = fma(-x, r, 1.0);
  r = fma(e, r, r);
  e = fma(-x, r, 1.0);
  r = fma(e, r, r);
  return r;
}
__global__ void k_rcp(double* out, long long* clk, int iters, int mode)
{
  double v = out[0] + 1.5;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (mode == 0) v = fast_rcp(v) + 1.25;
      else if (mode == 1) v = 1.0 / v + 1.25;
      else if (mode == 2) v = rsqrt(v) + 1.25;
      else if (mode == 3) v = sqrt(v) + 1.25;
      else v = (double)(1.0f / (float)v) + 1.25;
    }
  }
  long long t1 = clock64();
  out[64 + threadIdx.x] = v;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void k_trig(double* out, long long* clk, int iters, int mode)
{
  double v = out[0] + 0.3 + 0.001 * threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    if (mode == 0) { double s, c; sincos(v, &s, &c); v = s * 0.5 + c * 0.25 + 0.3; }
    else if (mode == 1) v = fmod(v + 7.0, 6.28318530718) * 0.2 + 0.1;
    else if (mode == 2) v = atan2(v, 1.0 - v * 0.1) + 0.2;
    else if (mode == 3) v = asin(v * 0.3) + 0.4;
  }
  long long t1 = clock64();
  out[64 + threadIdx.x] = v;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void k_satom(int* out, long long* clk, int iters)
{
  __shared__ int ctr;
  if (threadIdx.x == 0) ctr = 0;
  __syncthreads();
  int v = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    int c = 0;
    if ((threadIdx.x & 31) == 0) c = atomicAdd(&ctr, 1);
    v += __shfl_sync(0xffffffffu, c, 0);
  }
  long long t1 = clock64();
  out[threadIdx.x] = v;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void k_redux(int* out, long long* clk, int iters)
{
  unsigned v = out[threadIdx.x & 31];
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) v = __reduce_min_sync(0xffffffffu, v + threadIdx.x) + 3;
  }
  long long t1 = clock64();
  out[64 + threadIdx.x] = v;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}

__device__ __forceinline__ double fast_rcp2(double x)
{
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(e, r, r);
  e = fma(-x, r, 1.0);
  r = fma(e, r, r);
  return r;
}
// accuracy of the reciprocal with two / three Newton steps against IEEE division, in ulps
__global__ void k_rcp_acc(double* out, int n)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double worst2 = 0.0, worst3 = 0.0, seed = 0.0;
  for (int k = i; k < n; k += gridDim.x * blockDim.x) {
    // values over many binades, deterministic
    unsigned long long h = (unsigned long long)k * 0x9E3779B97F4A7C15ull;
    h ^= h >> 29;
    const double m = 1.0 + (double)(h & 0xFFFFFFFFFFFFFull) * (1.0 / 4503599627370496.0);
    const int ex = (int)((h >> 52) % 120) - 60;
    const double x = ldexp(m, ex);
    const double ref = 1.0 / x;
    const double ulp = ldexp(1.0, -52) * fabs(ref);
    worst2 = fmax(worst2, fabs(fast_rcp2(x) - ref) / ulp);
    worst3 = fmax(worst3, fabs(fast_rcp(x) - ref) / ulp);
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    seed = fmax(seed, fabs(r - ref) / fabs(ref));
  }
  out[3 * i] = worst2;
  out[3 * i + 1] = worst3;
  out[3 * i + 2] = seed;
}

int main()
{
  double* d; float* f; int* ii; long long* clk;
  CK(cudaMalloc(&d, 4096 * 8)); CK(cudaMalloc(&f, 4096 * 4)); CK(cudaMalloc(&ii, 4096 * 4)); CK(cudaMalloc(&clk, 4096 * 8));
  double hd[128]; for (int i = 0; i < 128; i++) hd[i] = 0.5 + 0.001 * i;
  float hf[128]; for (int i = 0; i < 128; i++) hf[i] = 0.5f + 0.001f * i;
  int hi[128]; for (int i = 0; i < 128; i++) hi[i] = i * 7;
  CK(cudaMemcpy(d, hd, sizeof(hd), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(f, hf, sizeof(hf), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(ii, hi, sizeof(hi), cudaMemcpyHostToDevice));
  long long h[1024];
  const int IT = 2000;
  auto get = [&](int n = 1) { CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h, clk, 8 * n, cudaMemcpyDeviceToHost)); long long m = 0; for (int i = 0; i < n; i++) m = h[i] > m ? h[i] : m; return (double)m; };
  for (int rep = 0; rep < 2; rep++) {
    k_dfma_lat<<<1, 32>>>(d, clk, IT); printf("DFMA dependent latency          : %.2f cyc\n", get() / (IT * 16.0));
    k_ffma_lat<<<1, 32>>>(f, clk, IT); printf("FMUL+FADD dependent pair        : %.2f cyc\n", get() / (IT * 16.0));
    k_dmma<1><<<1, 32>>>(d, clk, IT); printf("DMMA m8n8k4 dependent latency   : %.2f cyc\n", get() / (IT * 1.0));
    k_dmma<2><<<1, 32>>>(d, clk, IT); printf("DMMA 1 warp, 2 indep: per DMMA   : %.2f cyc\n", get() / (IT * 2.0));
    k_dmma<4><<<1, 32>>>(d, clk, IT); printf("DMMA 1 warp, 4 indep: per DMMA   : %.2f cyc\n", get() / (IT * 4.0));
    k_dmma<8><<<1, 32>>>(d, clk, IT); printf("DMMA 1 warp, 8 indep: per DMMA   : %.2f cyc\n", get() / (IT * 8.0));
    k_dmma<8><<<1, 128>>>(d, clk, IT); printf("DMMA 4 warps(1/sched), 8 indep   : %.2f cyc per DMMA per warp\n", get() / (IT * 8.0));
    k_dmma<8><<<1, 256>>>(d, clk, IT); printf("DMMA 8 warps, 8 indep            : %.2f cyc per DMMA per warp\n", get() / (IT * 8.0));
    k_dmma<8><<<1, 512>>>(d, clk, IT); printf("DMMA 16 warps, 8 indep           : %.2f cyc per DMMA per warp\n", get() / (IT * 8.0));
    k_dfma_tp<8><<<1, 32>>>(d, clk, IT); printf("DFMA 1 warp, 8 indep: per DFMA   : %.2f cyc\n", get() / (IT * 8.0));
    k_dfma_tp<8><<<1, 128>>>(d, clk, IT); printf("DFMA 4 warps, 8 indep            : %.2f cyc per DFMA per warp\n", get() / (IT * 8.0));
    k_dfma_tp<8><<<1, 256>>>(d, clk, IT); printf("DFMA 8 warps, 8 indep            : %.2f cyc per DFMA per warp\n", get() / (IT * 8.0));
    k_dfma_tp<8><<<1, 512>>>(d, clk, IT); printf("DFMA 16 warps, 8 indep           : %.2f cyc per DFMA per warp\n", get() / (IT * 8.0));
    for (int nt = 64; nt <= 512; nt *= 2) { k_sync<<<1, nt>>>(clk, IT); printf("__syncthreads %3d threads        : %.2f cyc\n", nt, get() / (IT * 8.0)); }
    k_shfl<<<1, 32>>>(d, clk, IT); printf("shfl f64 (+DADD) dependent       : %.2f cyc\n", get() / (IT * 8.0));
    k_shfl32<<<1, 32>>>(f, clk, IT); printf("shfl f32 dependent               : %.2f cyc\n", get() / (IT * 8.0));
    k_lds<<<1, 32>>>(ii, clk, IT); printf("LDS dependent (pointer chase)    : %.2f cyc\n", get() / (IT * 8.0));
    k_redux<<<1, 32>>>(ii, clk, IT); printf("REDUX.MIN dependent (+2 IADD)    : %.2f cyc\n", get() / (IT * 8.0));
    k_satom<<<1, 32>>>(ii, clk, IT); printf("smem atomicAdd + shfl bcast      : %.2f cyc\n", get() / (IT * 1.0));
    const char* nm[5] = {"fast_rcp (MUFU+3 Newton)", "IEEE 1.0/x", "rsqrt(double)", "sqrt(double)", "fp32 rcp round trip"};
    for (int m = 0; m < 5; m++) { k_rcp<<<1, 32>>>(d, clk, IT, m); printf("%-28s    : %.2f cyc (incl. 1 DADD)\n", nm[m], get() / (IT * 4.0)); }
    const char* tn[4] = {"sincos(double)", "fmod(double)", "atan2(double)", "asin(double)"};
    for (int m = 0; m < 4; m++) { k_trig<<<1, 32>>>(d, clk, 500, m); printf("%-28s    : %.2f cyc (incl. ~2 DFMA)\n", tn[m], get() / 500.0); }
    printf("----\n");
  }
  {
    double* acc; CK(cudaMalloc(&acc, 3 * 148 * 256 * 8));
    k_rcp_acc<<<148, 256>>>(acc, 50000000);
    CK(cudaDeviceSynchronize());
    static double ha[3 * 148 * 256];
    CK(cudaMemcpy(ha, acc, sizeof(ha), cudaMemcpyDeviceToHost));
    double w2 = 0, w3 = 0, sd = 0;
    for (int i = 0; i < 148 * 256; i++) { w2 = ha[3 * i] > w2 ? ha[3 * i] : w2; w3 = ha[3 * i + 1] > w3 ? ha[3 * i + 1] : w3; sd = ha[3 * i + 2] > sd ? ha[3 * i + 2] : sd; }
    printf("reciprocal vs IEEE 1/x over 5e7 values: seed rel err %.3e; 2 Newton steps worst %.2f ulp; 3 steps worst %.2f ulp\n", sd, w2, w3);
  }
  // whole-SM DMMA vs DFMA throughput: 148*? CTAs of 256 threads
  {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int mode = 0; mode < 2; mode++) {
      const int ctas = 148 * 4, IT2 = 20000;
      CK(cudaEventRecord(e0));
      if (mode == 0) k_dmma<8><<<ctas, 256>>>(d, clk, IT2); else k_dfma_tp<8><<<ctas, 256>>>(d, clk, IT2);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      const double flop = mode == 0 ? (double)ctas * 8 * IT2 * 8 * 512.0 : (double)ctas * 256 * IT2 * 8 * 2.0;
      printf("%s whole GPU: %.2f TFLOP/s (%.3f ms)\n", mode == 0 ? "DMMA" : "DFMA", flop / ms * 1e-9, ms);
    }
  }
  return 0;
}
