// Cost of a grid-wide barrier among 148 persistent CTAs (one per SM, 384 threads) on B200: what one phase boundary of the fused
// VB-loop kernel costs.  Variants: cooperative_groups grid.sync(); a monotone atomic counter polled with ld.acquire.gpu by one
// thread per CTA; the same with red.release + a separate "generation" word written by the last arriver (one poller line).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -rdc=false -o grid_barrier grid_barrier.cu ; run: ./grid_barrier
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ unsigned ld_acquire(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release(unsigned *p, unsigned v) { asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned atom_acq_rel(unsigned *p, unsigned v) {
  unsigned o;
  asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(o) : "l"(p), "r"(v) : "memory");
  return o;
}
__device__ __forceinline__ void st_release(unsigned *p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

__global__ void __launch_bounds__(384, 1) k_cg(int rounds) {
  cg::grid_group g = cg::this_grid();
  for (int r = 0; r < rounds; ++r) g.sync();
}
// counter barrier: every CTA adds 1, polls until the counter reaches (r+1)*grid
__global__ void __launch_bounds__(384, 1) k_counter(unsigned *ctr, int rounds) {
  for (int r = 0; r < rounds; ++r) {
    __syncthreads();
    if (threadIdx.x == 0) {
      red_release(ctr, 1u);
      const unsigned target = (unsigned)(r + 1) * gridDim.x;
      while (ld_acquire(ctr) < target) {
      }
    }
    __syncthreads();
  }
}
// last arriver publishes a generation word; the others poll the generation (read-only line shared by all pollers)
__global__ void __launch_bounds__(384, 1) k_gen(unsigned *ctr, unsigned *gen, int rounds) {
  for (int r = 0; r < rounds; ++r) {
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned old = atom_acq_rel(ctr, 1u);
      if (old == (unsigned)(r + 1) * gridDim.x - 1) st_release(gen, (unsigned)(r + 1));
      else
        while (ld_acquire(gen) < (unsigned)(r + 1)) {
        }
    }
    __syncthreads();
  }
}

int main() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  unsigned *ctr, *gen;
  cudaMalloc(&ctr, 256);
  gen = ctr + 32;
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  const int rounds = 2000;
  for (int variant = 0; variant < 3; ++variant) {
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
      cudaMemset(ctr, 0, 256);
      int rr = rounds;
      void *args_cg[] = {&rr};
      void *args_c[] = {&ctr, &rr};
      void *args_g[] = {&ctr, &gen, &rr};
      cudaEventRecord(a);
      if (variant == 0) cudaLaunchCooperativeKernel((void *)k_cg, dim3(sms), dim3(384), args_cg, 0, 0);
      if (variant == 1) cudaLaunchCooperativeKernel((void *)k_counter, dim3(sms), dim3(384), args_c, 0, 0);
      if (variant == 2) cudaLaunchCooperativeKernel((void *)k_gen, dim3(sms), dim3(384), args_g, 0, 0);
      cudaEventRecord(b);
      cudaError_t e = cudaEventSynchronize(b);
      if (e != cudaSuccess) {
        printf("{\"variant\": %d, \"error\": \"%s\"}\n", variant, cudaGetErrorString(e));
        return 1;
      }
      float ms;
      cudaEventElapsedTime(&ms, a, b);
      if (ms < best) best = ms;
    }
    const char *names[] = {"cg_grid_sync", "counter_poll", "generation_word"};
    printf("{\"barrier\": \"%s\", \"ctas\": %d, \"threads\": 384, \"rounds\": %d, \"us_per_barrier\": %.3f}\n", names[variant], sms, rounds, best * 1e3f / rounds);
  }
  return 0;
}
