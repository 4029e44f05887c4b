// FP64 peak probes for B200 (sm_100a): DMMA (mma.sync m8n8k4 f64), DFMA, exp(), HBM copy.
// Not product code: it measures the roofline denominators that MEASURED_PEAKS.json lacks (FP64).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o fp64_peaks fp64_peaks.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void k_dmma(double *out, int iters) {
  double c0[NACC], c1[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c0[i] = 0.0; c1[i] = 0.0; }
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) dmma884(c0[i], c1[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dfma(double *out, int iters) {
  double c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) c[i] = i * 1e-3;
  double a = 1.0 + threadIdx.x * 1e-12, b = 1e-9 * threadIdx.x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mixed: DMMA and DFMA interleaved, to see whether they share a pipe
template <int NACC>
__global__ void k_mixed(double *out, int iters) {
  double c0[NACC], c1[NACC], f[NACC * 2];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c0[i] = 0.0; c1[i] = 0.0; f[2*i] = i; f[2*i+1] = -i; }
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) {
      dmma884(c0[i], c1[i], a, b);
      // 8 DFMA per DMMA = same FMA-lane count as one DMMA.8x8x4 (256 FMA = 8 warp-DFMA)
#pragma unroll
      for (int j = 0; j < 4; j++) { f[2*i] = fma(f[2*i], a, b); f[2*i+1] = fma(f[2*i+1], a, b); }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c0[i] + c1[i] + f[2*i] + f[2*i+1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_exp(const double *in, double *out, int iters) {
  double x = in[threadIdx.x & 31];
  double s = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int j = 0; j < 8; j++) { s += exp(x); x -= 1e-3; }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_log(const double *in, double *out, int iters) {
  double x = in[threadIdx.x & 31] + 3.0;
  double s = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int j = 0; j < 8; j++) { s += log(x); x += 1e-3; }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_copy(const double2 *__restrict__ in, double2 *__restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i + 3 * stride < n; i += 4 * stride) {
    double2 a = in[i], b = in[i + stride], c = in[i + 2 * stride], d = in[i + 3 * stride];
    out[i] = a; out[i + stride] = b; out[i + 2 * stride] = c; out[i + 3 * stride] = d;
  }
  for (; i < n; i += stride) out[i] = in[i];
}
__global__ void k_read(const double2 *__restrict__ in, double *out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  double s = 0;
  for (; i + 3 * stride < n; i += 4 * stride) {
    double2 a = in[i], b = in[i + stride], c = in[i + 2 * stride], d = in[i + 3 * stride];
    s += a.x + a.y + b.x + b.y + c.x + c.y + d.x + d.y;
  }
  for (; i < n; i += stride) s += in[i].x + in[i].y;
  if (s == 1.2345e-300) out[0] = s;
}

template <typename F> float timeit(F f, int reps = 5) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, sms, p.clockRate);
  double *out; CK(cudaMalloc(&out, sizeof(double) * sms * 32 * 1024));
  double *in; CK(cudaMalloc(&in, 256)); CK(cudaMemset(in, 0, 256));
  const int iters = 20000;
  for (int warps : {1, 2, 4, 8, 16, 32}) {
    int threads = warps * 32 > 1024 ? 1024 : warps * 32;
    int blocks = sms * (warps * 32 / threads);
    float ms = timeit([&] { k_dmma<8><<<blocks, threads>>>(out, iters); });
    double fl = (double)blocks * (threads / 32) * iters * 8.0 * 512.0;
    printf("{\"probe\": \"dmma_m8n8k4\", \"warps_per_sm\": %d, \"tflops\": %.2f, \"ms\": %.3f}\n", warps, fl / ms * 1e-9, ms);
  }
  for (int warps : {4, 8, 16, 32}) {
    int threads = warps * 32 > 1024 ? 1024 : warps * 32;
    int blocks = sms;
    float ms = timeit([&] { k_dfma<16><<<blocks, threads>>>(out, iters); });
    double fl = (double)blocks * threads * iters * 16.0 * 2.0;
    printf("{\"probe\": \"dfma\", \"warps_per_sm\": %d, \"tflops\": %.2f, \"ms\": %.3f}\n", warps, fl / ms * 1e-9, ms);
  }
  for (int warps : {8, 16, 32}) {
    int threads = warps * 32, blocks = sms;
    float ms = timeit([&] { k_mixed<8><<<blocks, threads>>>(out, iters / 4); });
    double fl = (double)blocks * warps * (iters / 4) * 8.0 * (512.0 + 8 * 64.0);
    printf("{\"probe\": \"dmma+dfma_mixed\", \"warps_per_sm\": %d, \"tflops_total\": %.2f, \"ms\": %.3f}\n", warps, fl / ms * 1e-9, ms);
  }
  for (int warps : {8, 16, 32}) {
    int threads = warps * 32, blocks = sms;
    float ms = timeit([&] { k_exp<<<blocks, threads>>>(in, out, 2000); });
    double n = (double)blocks * threads * 2000 * 8.0;
    printf("{\"probe\": \"exp_f64\", \"warps_per_sm\": %d, \"gexp_per_s\": %.2f, \"ms\": %.3f}\n", warps, n / ms * 1e-6, ms);
    ms = timeit([&] { k_log<<<blocks, threads>>>(in, out, 2000); });
    printf("{\"probe\": \"log_f64\", \"warps_per_sm\": %d, \"glog_per_s\": %.2f, \"ms\": %.3f}\n", warps, n / ms * 1e-6, ms);
  }
  {
    size_t bytes = (size_t)2 << 30; // 2 GiB each
    double2 *a, *b; CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes)); CK(cudaMemset(a, 1, bytes)); CK(cudaMemset(b, 0, bytes));
    size_t n = bytes / sizeof(double2);
    for (int mult : {4, 8, 16}) {
      float ms = timeit([&] { k_copy<<<sms * mult, 512>>>(a, b, n); });
      printf("{\"probe\": \"hbm_copy_ldg128\", \"ctas_per_sm\": %d, \"gbs\": %.1f, \"ms\": %.3f}\n", mult, 2.0 * bytes / ms * 1e-6, ms);
      ms = timeit([&] { k_read<<<sms * mult, 512>>>(a, out, n); });
      printf("{\"probe\": \"hbm_read_ldg128\", \"ctas_per_sm\": %d, \"gbs\": %.1f, \"ms\": %.3f}\n", mult, 1.0 * bytes / ms * 1e-6, ms);
    }
    float ms = timeit([&] { CK(cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice)); });
    printf("{\"probe\": \"hbm_copy_cudaMemcpy\", \"gbs\": %.1f, \"ms\": %.3f}\n", 2.0 * bytes / ms * 1e-6, ms);
  }
  return 0;
}
