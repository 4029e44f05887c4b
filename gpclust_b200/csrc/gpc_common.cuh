// gpclust_b200 -- shared device helpers for the sm_100a kernels (FP64 throughout).
//
// FP64 tensor work on sm_100a is warp-level `mma.sync.m8n8k4.f64` (SASS DMMA.8x8x4); tcgen05 has no
// f64 kind, so there is no TMEM path for this data type.  Operand tiles are staged in shared memory
// by 1-D bulk async copies (cp.async.bulk + mbarrier complete_tx; SASS UBLKCP) into a row-padded
// layout that makes the 64-bit fragment reads bank-conflict free.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gpclust_b200.h"  // GPC_ROW_TILE (rows per CTA tile; every N-sized device array is
                                         // allocated with roundup(N, GPC_ROW_TILE) rows, pad rows zero) and enums
#define GPC_MAX_KP 64            // padded cluster count handled by the tensor kernels (8 n-tiles)
#define GPC_MAX_DP 64            // padded feature count (time points / MOG features)
#define GPC_THREADS 256          // 8 warps; warp 0 doubles as the F-tile producer
#define GPC_CONSUMER_WARPS 8

namespace gpc {

// ---------------------------------------------------------------- DMMA m8n8k4
// A (8x4, row): lane (g = lane/4, t = lane%4) holds A[g][t]
// B (4x8, col): lane holds B[t][g]
// C (8x8):      lane holds C[g][2t], C[g][2t+1]
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// ---------------------------------------------------------------- mbarrier + bulk async copy
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy (bytes multiple of 16, both addresses 16-B aligned), completion on mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------- vector global access (16 B)
__device__ __forceinline__ double2 ldg2(const double *p) { return *reinterpret_cast<const double2 *>(p); }
__device__ __forceinline__ double2 ldg2_stream(const double *p) {
  double2 v;
  asm volatile("ld.global.L1::no_allocate.v2.f64 {%0,%1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg2(double *p, double a, double b) { *reinterpret_cast<double2 *>(p) = make_double2(a, b); }

// ---------------------------------------------------------------- reductions
__device__ __forceinline__ double quad_sum(double v) {  // over the 4 lanes sharing g = lane/4
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}
__device__ __forceinline__ double quad_max(double v) {
  v = fmax(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = fmax(v, __shfl_xor_sync(0xffffffffu, v, 2));
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// sum over lanes with equal t = lane%4 (i.e. over the 8 row groups g)
__device__ __forceinline__ double colgroup_sum(double v) {
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  return v;
}

// ---------------------------------------------------------------- special functions (FP64)
// digamma for x > 0: upward recurrence to x >= 10, then the asymptotic series through B_14.
// |error| ~ 1e-16 relative for the arguments on this path (alpha + phi_hat > 0).
__device__ __forceinline__ double digamma_pos(double x) {
  double r = 0.0;
  while (x < 10.0) {
    r -= 1.0 / x;
    x += 1.0;
  }
  const double xi = 1.0 / x, x2 = xi * xi;
  double s = x2 * (1.0 / 12.0 - x2 * (1.0 / 120.0 - x2 * (1.0 / 252.0 - x2 * (1.0 / 240.0 - x2 * (1.0 / 132.0 - x2 * (691.0 / 32760.0 - x2 * (1.0 / 12.0)))))));
  return r + log(x) - 0.5 * xi - s;
}

}  // namespace gpc
