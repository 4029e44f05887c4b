// gpclust_b200 -- shared device helpers for the sm_100a kernels (FP64 throughout).
//
// FP64 tensor work on sm_100a is warp-level `mma.sync.m8n8k4.f64` (SASS DMMA.8x8x4); tcgen05 has no
// f64 kind, so there is no TMEM path for this data type.  Operands are staged in shared memory by 1-D bulk
// async copies (cp.async.bulk + mbarrier complete_tx; SASS UBLKCP); the HBM layout itself is fragment-major
// (gpc_vb_kernels.cuh) so that the copied bytes are already in the order the lanes read them.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gpclust_b200.h"  // GPC_ROW_TILE (rows per CTA tile; every N-sized device array is
                                         // allocated with roundup(N, GPC_ROW_TILE) rows, pad rows zero) and enums
#define GPC_MAX_KP 64            // padded cluster count handled by one launch of the tensor kernels (8 n-tiles)
#define GPC_MAX_KTOT 512         // more clusters are processed in chunks of GPC_MAX_KP
#define GPC_MAX_DP 64            // padded feature count (time points / MOG features)

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
  // try_wait with a suspend-time hint: the warp sleeps in hardware instead of burning issue slots in a spin
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(0x989680u)
      : "memory");
}
// global -> shared bulk copy (bytes multiple of 16, both addresses 16-B aligned), completion on mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------- thread-block clusters (distributed shared memory)
// K > 64: the CTAs of a cluster share the same row groups, CTA `rank` owning the clusters [64 rank, 64 rank + 64); the
// row-coupled scalars (softmax max / sum, sum_k phi*a) are exchanged through each other's shared memory.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_count_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// shared::cta address of this CTA -> shared::cluster address of the same variable in CTA `rank`
__device__ __forceinline__ uint32_t cluster_map(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
// asynchronous remote store whose completion is counted (complete_tx, 8 bytes) on the RECEIVER's mbarrier: the
// receiver arms its barrier with arrive.expect_tx(total bytes) and waits on it like on a bulk copy -- no MEMBAR on the
// sender, no L1 invalidation (CCTL.IVALL) on the receiver, which is what a release/acquire pair at cluster scope costs
__device__ __forceinline__ void st_async_f64(uint32_t remote_addr, double v, uint32_t remote_bar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];\n" ::"r"(remote_addr), "d"(v), "r"(remote_bar) : "memory");
}

__device__ __forceinline__ void mbar_inval(uint64_t *bar) { asm volatile("mbarrier.inval.shared::cta.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory"); }
// generic-proxy writes (st.global / st.shared by this thread, or observed by it) ordered before later async-proxy accesses
// (cp.async.bulk reads of the same bytes)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;\n" ::: "memory"); }
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int *p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add_u32(unsigned int *p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}

// bulk prefetch of a contiguous global range into L2 (no destination, no completion tracking)
__device__ __forceinline__ void bulk_prefetch_l2(const void *src_gmem, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(src_gmem), "r"(bytes) : "memory");
}

// ---------------------------------------------------------------- device-resident loop state (gpc_loop_t)
// true when this kernel of the enqueued sequence has nothing to do: the loop has terminated, or it is a RETRY-phase
// kernel and the conjugate step was accepted.  Uniform over the grid (the state was written by an earlier kernel).
__device__ __forceinline__ bool loop_skip(const gpc_loop_t *loop, int phase) {
  if (loop == nullptr) return false;
  if (loop->done != 0) return true;
  if (loop->slot_mode != 0) return false;  // one evaluation per slot: only the G pass ever skips (it checks need_retry itself)
  return phase == GPC_PHASE_RETRY && loop->need_retry == 0;
}
// phase of this kernel: the launch argument, or (slot mode) RETRY exactly when the previous slot's conjugate step was rejected.
// Read at kernel entry by every CTA; need_retry is only written by the last CTA of a posterior kernel after all CTAs are done.
__device__ __forceinline__ int loop_phase(const gpc_loop_t *loop, int phase) {
  if (loop != nullptr && loop->slot_mode != 0) return loop->need_retry ? GPC_PHASE_RETRY : GPC_PHASE_CONJ;
  return phase;
}
template <typename T>
__device__ __forceinline__ T sel3(T p0, T p1, T p2, int i) { return i == 0 ? p0 : (i == 1 ? p1 : p2); }

// ---------------------------------------------------------------- peer-memory exchange (gpc_peers_t)
__host__ __device__ inline long long xchg_l2(int len) { return (len + 1) & ~1; }
__host__ __device__ inline long long xchg_off_stats(int W, int len, int slot, int sender) { return ((long long)slot * W + sender) * xchg_l2(len); }
__host__ __device__ inline long long xchg_off_dots(int W, int len, int slot, int sender) { return 2LL * W * xchg_l2(len) + ((long long)slot * W + sender) * 4; }
__host__ __device__ inline long long xchg_off_flags(int W, int len, int which, int slot, int sender) {  // in doubles (= u64 words)
  return 2LL * W * xchg_l2(len) + 2LL * W * 4 + ((long long)which * 2 + slot) * W + sender;
}
__host__ __device__ inline long long xchg_doubles(int W, int len) { return 2LL * W * xchg_l2(len) + 2LL * W * 4 + 4LL * W; }

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
// Raise this rank's flag of exchange (`which`, `slot`) in every peer's inbox.  Called by ONE WARP after the block-wide barrier
// that follows the data stores: a single system-scope fence (executed once for the warp), then lane r stores the flag of peer r
// -- the eight NVLink round trips overlap instead of eight st.release.sys (= fence + store each) issued one after the other.
__device__ __forceinline__ void xchg_raise_flags(const gpc_peers_t *P, int which, int slot, unsigned long long epoch) {
  const int lane = threadIdx.x & 31;
  __threadfence_system();
  if (lane < P->world)
    st_relaxed_sys(reinterpret_cast<unsigned long long *>(P->inbox[lane]) + xchg_off_flags(P->world, P->stats_len, which, slot, P->rank), epoch);
}
// one thread: wait until every rank's flag of exchange `which` (0 statistics, 1 dots) in `slot` of MY inbox has reached
// `epoch`.  Bounded (a few seconds): returns false if a peer never shows up.
__device__ inline bool xchg_wait(const gpc_peers_t *P, int which, int slot, unsigned long long epoch) {
  const unsigned long long *flags = reinterpret_cast<const unsigned long long *>(P->inbox[P->rank]) + xchg_off_flags(P->world, P->stats_len, which, slot, 0);
  for (int r = 0; r < P->world; ++r) {
    long long spins = 0;
    while (ld_acquire_sys(flags + r) < epoch) {
      if (++spins > (1LL << 23)) return false;
      __nanosleep(64);
    }
  }
  return true;
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

// ---------------------------------------------------------------- exp for the softmax passes (FP64)
// exp(x) for a BATCH of arguments (x <= ~0 on this path: x - logsumexp or x - rowmax), ~1.5 ulp.
// Cody-Waite reduction x = n ln2 + r with the 2^52 rounding trick, degree-11 near-minimax polynomial on
// |r| <= ln2/2 (Chebyshev-node interpolant of exp, coefficients generated with mpmath at 60 digits),
// scaling by 2^n through the exponent field.  Arguments are clamped at -708 (exp = 3.3e-308, the smallest
// normal range) instead of being flushed to zero / denormals: no branch, no special cases, and every
// Horner step is issued for all N chains back to back -- FP64 FMA latency, not throughput, is what a
// one-at-a-time evaluation pays for.  NaN maps to exp(-708).
__constant__ double kExpC[17] = {
    1.4426950408889634074,        // 0  log2(e)
    6755399441055744.0,           // 1  1.5 * 2^52
    -6.93147180369123816490e-01,  // 2  -ln2 hi
    -1.90821492927058770002e-10,  // 3  -ln2 lo
    -708.0,                       // 4  clamp
    0x1.af632e7d3bc8bp-26,        // 5  c11
    0x1.28b413b5ef026p-22,        // 6  c10
    0x1.71ddf56b28fcep-19,        // 7  c9
    0x1.a019919d2943ep-16,        // 8  c8
    0x1.a01a01b147016p-13,        // 9  c7
    0x1.6c16c188019e5p-10,        // 10 c6
    0x1.111111110f21cp-7,         // 11 c5
    0x1.555555554f0b3p-5,         // 12 c4
    0x1.555555555555ap-3,         // 13 c3
    0x1.0000000000011p-1,         // 14 c2
    1.0,                          // 15 c1
    1.0};                         // 16 c0

template <int N>
__device__ __forceinline__ void exp_batch(double (&x)[N]) {  // in place: x[i] <- exp(x[i])
  int k[N];
  const double l2e = kExpC[0], shift = kExpC[1], ln2hi = kExpC[2], ln2lo = kExpC[3], lo = kExpC[4];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const double xi = (x[i] > lo) ? x[i] : lo;
    const double t = fma(xi, l2e, shift);
    const double n = t - shift;
    k[i] = __double2loint(t);
    x[i] = fma(n, ln2lo, fma(n, ln2hi, xi));  // r
  }
  double p[N];
  {
    const double c11 = kExpC[5], c10 = kExpC[6];
#pragma unroll
    for (int i = 0; i < N; ++i) p[i] = fma(c11, x[i], c10);
  }
#pragma unroll
  for (int c = 7; c < 17; ++c) {
    const double cc = kExpC[c];
#pragma unroll
    for (int i = 0; i < N; ++i) p[i] = fma(p[i], x[i], cc);
  }
#pragma unroll
  for (int i = 0; i < N; ++i) x[i] = __hiloint2double(__double2hiint(p[i]) + (k[i] << 20), __double2loint(p[i]));
}

// exp for the G pass (FP64-issue-bound): table-driven variant of exp_batch.  x = (32 k + j) ln2/32 + r, |r| <= ln2/64, so a
// degree-6 Taylor polynomial suffices (truncation 3.5e-18) and exp(x) = 2^k * 2^(j/32) * p(r): 11 FP64 instructions per value instead
// of 15.  2^(j/32) comes from a 32-entry table held one entry per LANE (`tab_lane` = kExpTab[lane]) and fetched with a warp shuffle:
// no shared memory, no bank conflicts.  Every lane of the warp must be active at the call.  ~1.4 ulp (checked against a
// 64-bit-mantissa reference over [-708, 3]); arguments clamped at -708 as in exp_batch.
__constant__ double kExpTab[32] = {
    0x1.0000000000000p+0, 0x1.059b0d3158574p+0, 0x1.0b5586cf9890fp+0, 0x1.11301d0125b51p+0,
    0x1.172b83c7d517bp+0, 0x1.1d4873168b9aap+0, 0x1.2387a6e756238p+0, 0x1.29e9df51fdee1p+0,
    0x1.306fe0a31b715p+0, 0x1.371a7373aa9cbp+0, 0x1.3dea64c123422p+0, 0x1.44e086061892dp+0,
    0x1.4bfdad5362a27p+0, 0x1.5342b569d4f82p+0, 0x1.5ab07dd485429p+0, 0x1.6247eb03a5585p+0,
    0x1.6a09e667f3bcdp+0, 0x1.71f75e8ec5f74p+0, 0x1.7a11473eb0187p+0, 0x1.82589994cce13p+0,
    0x1.8ace5422aa0dbp+0, 0x1.93737b0cdc5e5p+0, 0x1.9c49182a3f090p+0, 0x1.a5503b23e255dp+0,
    0x1.ae89f995ad3adp+0, 0x1.b7f76f2fb5e47p+0, 0x1.c199bdd85529cp+0, 0x1.cb720dcef9069p+0,
    0x1.d5818dcfba487p+0, 0x1.dfc97337b9b5fp+0, 0x1.ea4afa2a490dap+0, 0x1.f50765b6e4540p+0};
template <int N>
__device__ __forceinline__ void exp_batch_tab(double (&x)[N], double tab_lane) {  // in place: x[i] <- exp(x[i])
  int k[N];
  const double l2e32 = kExpC[0] * 32.0, shift = kExpC[1], ln2hi = kExpC[2] * 0.03125, ln2lo = kExpC[3] * 0.03125, lo = kExpC[4];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const double xi = (x[i] > lo) ? x[i] : lo;
    const double t = fma(xi, l2e32, shift);
    const double n = t - shift;
    k[i] = __double2loint(t);
    x[i] = fma(n, ln2lo, fma(n, ln2hi, xi));  // r
  }
  double p[N];
#pragma unroll
  for (int i = 0; i < N; ++i) p[i] = fma(1.0 / 720.0, x[i], 1.0 / 120.0);
#pragma unroll
  for (int c = 0; c < 5; ++c) {
    const double cc = (c == 0) ? 1.0 / 24.0 : (c == 1) ? 1.0 / 6.0 : (c == 2) ? 0.5 : 1.0;
#pragma unroll
    for (int i = 0; i < N; ++i) p[i] = fma(p[i], x[i], cc);
  }
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const double v = __shfl_sync(0xffffffffu, tab_lane, k[i] & 31) * p[i];
    x[i] = __hiloint2double(__double2hiint(v) + ((k[i] >> 5) << 20), __double2loint(v));
  }
}

__device__ __forceinline__ void exp_pair(double x0, double x1, double &e0, double &e1) {
  double v[2] = {x0, x1};
  exp_batch<2>(v);
  e0 = v[0];
  e1 = v[1];
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
