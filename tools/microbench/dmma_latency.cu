// DMMA (mma.sync m8n8k4 f64) and DFMA issue interval and dependent-issue latency on sm_100a: `warps_per_sm` warps per SM, NACC
// independent accumulator chains issued round-robin.  cycles per DMMA = max(issue interval, latency / NACC).  Not product code.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_latency dmma_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NACC>
__global__ void k(double *out, long long *cyc, int iters) {
  double c0[NACC], c1[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) c0[i] = c1[i] = 0.0;
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) dmma884(c0[i], c1[i], a, b);
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int NACC>
__global__ void kf(double *out, long long *cyc, int iters) {
  double c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) c[i] = i * 1e-3;
  const double a = 1.0 + threadIdx.x * 1e-12, b = 1e-9 * threadIdx.x;
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = fma(c[i], a, b);
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int NACC>
void runf(double *out, long long *cyc, int warps) {
  const int iters = 20000;
  kf<NACC><<<148, 32 * warps>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("{\"probe\": \"dfma_chains\", \"warps_per_sm\": %d, \"chains_per_warp\": %d, \"cycles_per_dfma_per_warp\": %.2f}\n", warps, NACC,
         (double)h / ((double)iters * NACC));
}
template <int NACC>
void run(double *out, long long *cyc, int warps) {
  const int iters = 20000;
  k<NACC><<<148, 32 * warps>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("{\"probe\": \"dmma_chains\", \"warps_per_sm\": %d, \"chains_per_warp\": %d, \"cycles_per_dmma_per_warp\": %.2f}\n", warps, NACC,
         (double)h / ((double)iters * NACC));
}
int main() {
  double *out;
  long long *cyc;
  cudaMalloc(&out, 148 * 1024 * 8);
  cudaMalloc(&cyc, 8);
  for (int warps : {1, 4, 8, 12}) {
    run<1>(out, cyc, warps);
    run<2>(out, cyc, warps);
    run<3>(out, cyc, warps);
    run<4>(out, cyc, warps);
    run<6>(out, cyc, warps);
    run<8>(out, cyc, warps);
    run<12>(out, cyc, warps);
    run<16>(out, cyc, warps);
  }
  for (int warps : {1, 4, 8}) {
    runf<1>(out, cyc, warps);
    runf<2>(out, cyc, warps);
    runf<4>(out, cyc, warps);
    runf<8>(out, cyc, warps);
    runf<16>(out, cyc, warps);
  }
  return 0;
}
