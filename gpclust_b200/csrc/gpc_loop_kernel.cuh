// gpclust_b200 -- the conjugate natural-gradient loop of MOHGP as ONE persistent cooperative kernel (FP64, sm_100a).
//
// One launch runs up to `max_evals` bound evaluations of collapsed_vb.py:147-303.  Per evaluation the phases of the former launch
// sequence  G pass -> S pass -> reduce / push -> posterior  run back to back on the same 148 CTAs (one per SM, 12 warps):
//
//   G rows   (g_rows, gpc_vb_kernels.cuh)       natural gradient + per-CTA partial dots            MOHGP.py:137-153
//   B1       grid barrier; fixed-order sum of the partial dots; with rows sharded over GPUs CTA 0 pushes the three dots into every
//            peer's inbox over NVLink and every CTA sums the W contributions in rank order           collapsed_vb.py:154,160-164
//   S rows   (s_producer / s_softmax / s_stats)  CG step, softmax, entropy, per-CTA statistics      collapsed_vb.py:167,172, MOHGP.py:76
//   B2       grid barrier
//   post     CTA k < KP: fixed-order reduction of cluster k's statistics row over the CTAs; sharded rows: the row goes straight into
//            every peer's inbox with ONE flag per (sender, cluster), the receiver CTA waits for its own cluster's W flags only and
//            sums in rank order; posterior in the eigenbasis (post_cluster_math); the last CTA assembles the bound and takes the
//            accept / reject / termination decision (post_tail -> cg_decide)                        MOHGP.py:79-90,124-135
//   B3       grid barrier; every CTA re-reads the loop state
//
// What this removes compared with the launch sequence: four kernel boundaries per evaluation (drain + launch + ramp-up), the
// separate reduce kernel and its full-vector push, one all-to-all rendezvous at CTA-0 granularity.  What it adds: three grid
// barriers of ~1.2 us (tools/microbench/grid_barrier.cu).  The S phase's producer warp does not take part in B1: it fills the ring
// with [F | x | ng | sd] of this CTA's first groups while the other warps wait for the dots -- the group -> CTA assignment is the
// same in both row phases (group = cta + m * grid), so everything a CTA stages in the S phase was written by ITSELF in the G phase.
//
// Coherence inside one launch: N-sized operands are CTA-private (see above; generic-proxy writes are ordered before the next
// phase's bulk copies by fence.proxy.async + the block barrier).  Everything K- or D-sized that crosses CTAs (loop state, Wt, bias,
// partials, per-cluster scalars, dots, beta) is read through L2 (__ldcg / ld.acquire) after a grid barrier -- never through L1.
#pragma once
#include "gpc_cluster_kernels.cuh"
#include "gpc_vb_kernels.cuh"

namespace gpc {

constexpr int kLoopHeaderBytes = 1024;  // loop-state snapshot + scalars at the start of dynamic shared memory
constexpr long long kSpinLimitNs = 4000000000LL;  // bounded waits (grid barrier, peer flags): 4 s

// ---- inbox layout of the fused exchange: FLAG-IN-DATA messages, no fences and no separate flags.
// Every double travels as ONE 16-byte store {lo32, epoch32, hi32, epoch32} (the scheme of NCCL's LL protocol): the receiver polls the
// 16-byte word itself until BOTH epoch fields carry the epoch it expects -- a torn 16-byte write (8-byte atomicity is all that is
// guaranteed) can then never be taken for a complete one.  The sender fires its stores and moves on: no __threadfence_system, no
// flag round trip.  Two slots (epoch parity): a rank is at most one exchange ahead of the slowest (every exchange is all-to-all).
//   in 16-byte units:  [2][W][64 clusters][kLLRow] statistics rows (row k: T[k][0..DP) | phi_hat_k | H (with cluster 0))
//                      [2][W][4] dots
constexpr int kLLRow = GPC_MAX_DP + 8;
__host__ __device__ inline long long ll_off_stats(int W, int slot, int sender, int k) { return (((long long)slot * W + sender) * GPC_MAX_KP + k) * kLLRow; }
__host__ __device__ inline long long ll_off_dots(int W, int slot, int sender) { return 2LL * W * GPC_MAX_KP * kLLRow + ((long long)slot * W + sender) * 4; }
__host__ __device__ inline long long ll_units(int W) { return 2LL * W * GPC_MAX_KP * kLLRow + 2LL * W * 4; }
__host__ __device__ inline long long fx_doubles(int W, int len) { (void)len; return 2 * ll_units(W); }

__device__ __forceinline__ void ll_send(double *inbox_peer, long long unit, double v, unsigned int ep) {
  const unsigned int lo = (unsigned int)__double2loint(v), hi = (unsigned int)__double2hiint(v);
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};\n" ::"l"(reinterpret_cast<uint4 *>(inbox_peer) + unit), "r"(lo), "r"(ep), "r"(hi), "r"(ep)
               : "memory");
}
// poll one message of this rank's own inbox until it carries epoch `ep`; bounded.  *ok is cleared on a timeout.
__device__ __forceinline__ double ll_recv(const double *inbox_self, long long unit, unsigned int ep, bool *ok) {
  const uint4 *p = reinterpret_cast<const uint4 *>(inbox_self) + unit;
  unsigned int lo, f0, hi, f1;
  unsigned long long t0 = 0ull;
  int spins = 0;
  for (;;) {
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];\n" : "=r"(lo), "=r"(f0), "=r"(hi), "=r"(f1) : "l"(p) : "memory");
    if (f0 == ep && f1 == ep) break;
    if (((++spins) & 255) == 0) {
      const unsigned long long t = globaltimer_ns();
      if (t0 == 0ull) t0 = t;
      else if ((long long)(t - t0) > kSpinLimitNs) {
        *ok = false;
        break;
      }
    }
  }
  return __hiloint2double((int)hi, (int)lo);
}

struct LoopKernelArgs {
  NatgradArgs g;    // F, Wt, bias, partials (grid x 4), dots_out, N, K, DP, num_groups, stages, bufs      (loop / peers unused)
  StepStatsArgs s;  // F, beta_out, partials (grid x stats_len), N, K, DP, num_groups, stages, bufs         (loop / peers / dots unused)
  MohgpPostArgs p;  // stats, ws, Sy_inv, Sf, const_term, Wt, muk_t, ..., info, alpha, prior, K, D, KP, DP, dots, beta, loop (GLOBAL)
  const gpc_peers_t *peers;     // nullable; inbox layout ll_* above
  unsigned int *sync;           // [0] grid-barrier counter, [1] exit counter, [2] posterior ticket -- all zero between launches
  unsigned long long *timing;   // nullable: [0..3] ns spent in G rows / B1 + dots / S rows / B2 + posterior + B3, [4] evaluations, [5] G phases
  int max_evals;
};

struct LoopShared {  // header of dynamic shared memory: everything that lives across phases is here, not in registers of the row loops
  gpc_loop_t L;      // snapshot of the device loop state, taken once per evaluation through L2
  MohgpPostArgs p;   // posterior arguments (copied from the kernel parameters once: the out-of-line phases read them from here)
  const gpc_peers_t *peers;
  double *inbox_self;
  unsigned int *sync;
  unsigned long long *timing;
  const double *gpart, *spart;
  double *dots_out;
  double dots[4];
  unsigned long long e_stats, e_dots;  // exchange epochs (uniform over the grid)
  unsigned long long t_acc[4], n_eval, n_g, t[4];
  unsigned long long p_acc[7], pt[7], dbg[8], dbg_acc[8];  // posterior phase in detail (CTA 0): B2 wait / local row reduce / exchange / math + ticket / tail + B3
  unsigned int bar_target;
  int abort, W, my_rank, len;
  int ev, max_evals, retry, g_stages, s_stages, KP, DP;
  const gpc_loop_t *loop_g;  // the loop state in global memory
  double *spart_mine;        // this CTA's block of partial statistics
};
static_assert(sizeof(LoopShared) <= kLoopHeaderBytes, "loop header");

// Grid-wide barrier among all CTAs of the (cooperative) launch.  NTHREADS threads of this CTA take part (named barrier BAR_ID;
// 0 = the whole CTA).  Monotone counter; hs.bar_target (thread 0) is the value the counter reaches when everybody arrived.
// Returns false when the wait ran into the spin limit (a CTA of this grid is stuck in a peer wait that timed out elsewhere).
template <int BAR_ID, int NTHREADS>
__device__ __forceinline__ bool grid_barrier(LoopShared &hs) {
  asm volatile("bar.sync %0, %1;\n" ::"n"(BAR_ID), "n"(NTHREADS) : "memory");
  if (threadIdx.x == 0 && hs.abort == 0) {  // once a wait of this CTA has timed out no further barrier spins: the loop is being left
    const unsigned int target = hs.bar_target + gridDim.x;
    hs.bar_target = target;
    __threadfence();
    red_release_gpu_add_u32(hs.sync, 1u);
    const unsigned long long t0 = globaltimer_ns();
    int spins = 0;
    while ((int)(ld_acquire_gpu_u32(hs.sync) - target) < 0) {
      if (((++spins) & 1023) == 0 && (long long)(globaltimer_ns() - t0) > kSpinLimitNs) {
        hs.abort = 1;
        break;
      }
    }
  }
  asm volatile("bar.sync %0, %1;\n" ::"n"(BAR_ID), "n"(NTHREADS) : "memory");
  return *reinterpret_cast<volatile int *>(&hs.abort) == 0;
}

// one thread: wait until the u64 word at `flag` (this rank's own inbox) has reached `epoch`; bounded
__device__ __forceinline__ bool flag_wait(const unsigned long long *flag, unsigned long long epoch) {
  const unsigned long long t0 = globaltimer_ns();
  int spins = 0;
  while (ld_acquire_sys(flag) < epoch) {
    if (((++spins) & 255) == 0 && (long long)(globaltimer_ns() - t0) > kSpinLimitNs) return false;
  }
  return true;
}

constexpr int kRowThreads = (kSSoftmaxWarps + kSStatWarps) * 32;  // the S phase's consumers (everything but the producer warp)

// The phases between the row loops are OUT OF LINE (their own register allocation, arguments through the shared-memory header):
// the row loops run at the register limit of the CTA shape (168) and must not share it with anything that lives across phases.

// B1 + the three dots (collapsed_vb.py:154,160-164), run by the S phase's consumer warps.  On return hs.dots holds the dots summed
// over CTAs (and ranks, in rank order); CTA 0 has stored them for the host.
__device__ __noinline__ void loop_dots_phase(LoopShared &hs) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cid = (int)blockIdx.x, ncl = (int)gridDim.x;
  grid_barrier<3, kRowThreads>(hs);  // every CTA's partial dots are in L2 (a timed-out wait sets hs.abort; the loop ends at B2)
  const gpc_peers_t *P = hs.peers;
  if (P == nullptr) {
    if (warp < 3) {
      const double v = reduce_dot_warp(hs.gpart, ncl, warp);
      if (lane == 0) {
        hs.dots[warp] = v;
        if (cid == 0) hs.dots_out[warp] = v;
      }
    }
  } else {
    const int W = hs.W, my_rank = hs.my_rank;
    const unsigned long long e_dots = hs.e_dots + 1;  // uniform; thread 0 stores it back after the last barrier below
    const int slot = (int)(e_dots & 1);
    const unsigned int ep = (unsigned int)e_dots;
    if (cid == 0 && warp < 3) {
      const double v = reduce_dot_warp(hs.gpart, ncl, warp);
      if (lane < W) ll_send(P->inbox[lane], ll_off_dots(W, slot, my_rank) + warp, v, ep);  // lane r -> peer r, over NVLink
    }
    if (warp == 0) {
      // lane (r, q) = (lane / 4, lane % 4) receives dot q of rank r; the W contributions are summed in rank order (identical on all ranks)
      bool ok = true;
      const int r = lane >> 2, q = lane & 3;
      double d = 0.0;
      if (r < W && q < 3) d = ll_recv(hs.inbox_self, ll_off_dots(W, slot, r) + q, ep, &ok);
      ok = __all_sync(0xffffffffu, ok);
      double tot = 0.0;
      for (int rr = 0; rr < W; ++rr) tot += __shfl_sync(0xffffffffu, d, rr * 4 + q);
      if (lane < 3) {
        hs.dots[lane] = ok ? tot : nan("");
        if (cid == 0) hs.dots_out[lane] = hs.dots[lane];
      }
      if (!ok && lane == 0) atomicOr(&hs.p.loop->done, GPC_DONE_PEER_TIMEOUT);
    }
  }
  asm volatile("bar.sync 3, %0;\n" ::"n"(kRowThreads) : "memory");
  if (P != nullptr && tid == 0) hs.e_dots += 1;  // every consumer has read the old value before the barrier above
  if (hs.timing != nullptr && cid == 0 && tid == 0) hs.t[2] = globaltimer_ns();
}

// B2, the per-cluster posterior (CTA k < KP, first kClusterThreads threads), the decision, B3.  Returns false when a barrier wait
// timed out.  smem_raw: the (now idle) ring memory.
__device__ __noinline__ bool loop_post_phase(LoopShared &hs, unsigned char *smem_raw, int phase) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cid = (int)blockIdx.x, ncl = (int)gridDim.x;
  if (!grid_barrier<0, kGThreads>(hs)) return false;
  const bool stamp = hs.timing != nullptr && cid == 0 && tid == 0;
  if (stamp) hs.pt[0] = hs.pt[1] = hs.pt[2] = hs.pt[3] = globaltimer_ns();
  const MohgpPostArgs &pa = hs.p;
  const int KP = pa.KP, DP = pa.DP, D = pa.D, len = hs.len;
  const gpc_peers_t *P = hs.peers;
  const unsigned long long e_stats = hs.e_stats + 1;  // read by every thread here; thread 0 stores it back after B3
  if (cid < KP && tid < kClusterThreads) {
    const int k = cid;
    double *ws_s = reinterpret_cast<double *>(smem_raw);
    PostSmem &ps = *reinterpret_cast<PostSmem *>(smem_raw + (((size_t)eig_ws_len(D) * 8 + 127) & ~(size_t)127));
    if (tid == 0) post_stage_ws(pa.ws, D, ps, ws_s);
    double phi_hat = post_reduce_row<true>(hs.spart, ncl, len, k, KP, DP, ps);
    if (stamp) hs.pt[1] = hs.pt[2] = globaltimer_ns();
    if (P != nullptr) {
      const int W = hs.W, my_rank = hs.my_rank;
      const int slot = (int)(e_stats & 1);
      const unsigned int ep = (unsigned int)e_stats;
      // this rank's row of cluster k (+ its entropy, with cluster 0) -> every peer's inbox as flag-in-data messages
      double Hloc = 0.0;
      if (k == 0) {
        for (int c = tid; c < ncl; c += kClusterThreads) Hloc += __ldcg(hs.spart + (size_t)c * len + (size_t)KP * DP + KP);
        Hloc = block_sum<true>(Hloc, ps.red_s);
      }
      const long long base = ll_off_stats(W, slot, my_rank, k);
      if (tid < DP + 2) {
        const double v = tid < DP ? ps.yb_s[tid] : (tid == DP ? phi_hat : Hloc);
        for (int r = 0; r < W; ++r) ll_send(P->inbox[r], base + tid, v, ep);
      }
      // the W rows of cluster k, one message per thread and round; summed in rank order (identical on all ranks)
      bool ok = true;
      const int rowlen = DP + 2;
      for (int i = tid; i < W * rowlen; i += kClusterThreads) {
        const int r = i / rowlen, j = i - r * rowlen;
        ps.xrow[r][j] = ll_recv(hs.inbox_self, ll_off_stats(W, slot, r, k) + j, ep, &ok);
      }
      if (!ok) {
        atomicOr(pa.info, GPC_INFO_PEER_TIMEOUT);
        atomicOr(&pa.loop->done, GPC_DONE_PEER_TIMEOUT);
      }
      psync<true>();
      double tot = 0.0;
      if (tid < DP) {
        for (int r = 0; r < W; ++r) tot += ps.xrow[r][tid];
        ps.yb_s[tid] = tot;
      }
      phi_hat = 0.0;
      for (int r = 0; r < W; ++r) phi_hat += ps.xrow[r][DP];
      psync<true>();
    }
    if (tid < DP) pa.stats[(size_t)k * DP + tid] = ps.yb_s[tid];
    if (tid == 0) pa.stats[(size_t)KP * DP + k] = phi_hat;
    if (stamp) hs.pt[2] = globaltimer_ns();
    mbar_wait(&ps.stage_bar, 0);
    if (stamp) hs.pt[4] = globaltimer_ns();
    post_cluster_math<true>(pa, k, phi_hat, ps, ws_s, stamp ? hs.dbg : nullptr);
    if (stamp) {
      hs.pt[5] = globaltimer_ns();
      hs.dbg_acc[0] += hs.dbg[0] - hs.pt[4];
      for (int q = 1; q < 5; ++q) hs.dbg_acc[q] += hs.dbg[q] - hs.dbg[q - 1];
      hs.dbg_acc[5] += hs.pt[5] - hs.dbg[4];
    }
    __threadfence();
    psync<true>();
    if (tid == 0) {
      atomicAdd(pa.counter, 1u);  // ticket: the scalar CTA finishes the evaluation once all KP clusters are in
      mbar_inval(&ps.stage_bar);
    }
    if (stamp) hs.pt[3] = globaltimer_ns();
  } else if (cid == KP && tid < kClusterThreads) {
    // ---- the "scalar" CTA: everything of the evaluation's tail that needs the statistics only -- entropy, phi_hat of every cluster,
    // the mixing-proportion terms (lgamma / digamma per cluster) -- runs HERE, concurrently with the cluster CTAs' algebra; then it
    // waits for their ticket and assembles the bound, the bias and the accept / reject / termination decision.
    PostSmem &ps = *reinterpret_cast<PostSmem *>(smem_raw);
    if (P != nullptr) {
      // phi_hat of every cluster and the entropy from the same messages the cluster CTAs consume, summed in the same (rank) order
      const int W = hs.W, slot = (int)(e_stats & 1);
      const unsigned int ep = (unsigned int)e_stats;
      bool ok = true;
      for (int k = tid; k < pa.K; k += kClusterThreads) {
        double v = 0.0;
        for (int r = 0; r < W; ++r) v += ll_recv(hs.inbox_self, ll_off_stats(W, slot, r, k) + DP, ep, &ok);
        ps.phi_hat_sh[k] = v;
      }
      if (tid == kClusterThreads - 1) {
        double v = 0.0;
        for (int r = 0; r < W; ++r) v += ll_recv(hs.inbox_self, ll_off_stats(W, slot, r, 0) + DP + 1, ep, &ok);
        ps.H_s = v;
      }
      if (!ok) atomicOr(&pa.loop->done, GPC_DONE_PEER_TIMEOUT);
      psync<true>();
      post_tail_mixing<true>(pa, ps);
    } else {
      for (int k = warp; k < pa.K; k += kClusterThreads / 32) {
        const double v = reduce_phi_hat_warp(hs.spart, ncl, len, KP, DP, k);  // the same routine as cluster k's own CTA: the same bits
        if (lane == 0) ps.phi_hat_sh[k] = v;
      }
      psync<true>();
      post_tail_prepare<true>(pa, ps, hs.spart, ncl, len);
    }
    if (tid == 0) {
      const unsigned long long t0 = globaltimer_ns();
      int spins = 0;
      while (ld_acquire_gpu_u32(pa.counter) < (unsigned)KP) {
        if (((++spins) & 1023) == 0 && (long long)(globaltimer_ns() - t0) > kSpinLimitNs) {
          atomicOr(&pa.loop->done, GPC_DONE_PEER_TIMEOUT);
          break;
        }
      }
    }
    psync<true>();
    post_tail_finish<true>(pa, ps, phase);  // resets the ticket (pa.counter)
    __threadfence();
  }
  // the posterior's scratch and the staged workspace live in the ring memory: generic-proxy accesses ordered before the next phase's
  // bulk copies into the same bytes
  fence_proxy_async();
  // B3: the decision and the new weights are visible to every CTA
  if (!grid_barrier<0, kGThreads>(hs)) return false;
  if (tid == 0) {
    hs.e_stats = e_stats;
    if (hs.timing != nullptr && cid == 0) {
      const unsigned long long t4 = globaltimer_ns();
      hs.t_acc[0] += hs.t[1] - hs.t[0];
      hs.t_acc[1] += hs.t[2] - hs.t[1];
      hs.t_acc[2] += hs.t[3] - hs.t[2];
      hs.t_acc[3] += t4 - hs.t[3];
      hs.n_eval += 1;
      hs.n_g += (phase == GPC_PHASE_RETRY) ? 0 : 1;
      hs.p_acc[0] += hs.pt[0] - hs.t[3];
      hs.p_acc[1] += hs.pt[1] - hs.pt[0];
      hs.p_acc[2] += hs.pt[2] - hs.pt[1];
      hs.p_acc[3] += hs.pt[3] - hs.pt[2];
      hs.p_acc[4] += t4 - hs.pt[3];
      hs.p_acc[5] += hs.pt[4] - hs.pt[2];  // of the algebra interval: waiting for the staged workspace
      hs.p_acc[6] += hs.pt[5] - hs.pt[4];  // ... and the algebra itself
    }
  }
  return true;
}

#ifndef GPC_LOOP_G_NOINLINE
#define GPC_LOOP_G_NOINLINE 0
#endif
#ifndef GPC_LOOP_S_NOINLINE
#define GPC_LOOP_S_NOINLINE 0
#endif
#if GPC_LOOP_G_NOINLINE
#define GPC_G_ATTR __noinline__
#else
#define GPC_G_ATTR __forceinline__
#endif
#if GPC_LOOP_S_NOINLINE
#define GPC_S_ATTR __noinline__
#else
#define GPC_S_ATTR __forceinline__
#endif
// the row phases as callable units (kernel parameters are __grid_constant__: their address can be passed on without a local copy)
template <int KP, bool FULLK>
__device__ GPC_G_ATTR void loop_g_phase(const NatgradArgs &ga, LoopShared &hs, unsigned char *smem_raw) {
  const G2Smem gm = g2_carve(smem_raw, ga.stages, KP, ga.DP);
  const GPtrs gp = g_resolve(ga, KP, 0, &hs.L);
  g2_setup(gm, ga.stages);
  __syncthreads();
  g2_rows<KP, FULLK>(ga, gm, gp, (int)blockIdx.x, (int)gridDim.x);  // the producer warp starts the HBM traffic while the product warps load W
  fence_proxy_async();  // this CTA's natural gradient is staged by its own bulk copies in the S phase
  __syncthreads();
}
template <int KP, bool FULLK>
__device__ GPC_S_ATTR void loop_s_softmax(const StepStatsArgs &sa, LoopShared &hs, unsigned char *smem_raw, double beta) {
  const SSmem sm = s_carve(smem_raw, sa.stages, KP, sa.DP);
  const SPtrs sp = s_resolve(sa, KP, 0, &hs.L, hs.retry ? GPC_PHASE_RETRY : GPC_PHASE_CONJ);
  s_softmax<KP, FULLK, false>(sa, sm, sp, (int)blockIdx.x, (int)gridDim.x, 0, 1, beta, true, beta != 0.0);
}
template <int KP>
__device__ GPC_S_ATTR void loop_s_stats(const StepStatsArgs &sa, unsigned char *smem_raw) {
  const SSmem sm = s_carve(smem_raw, sa.stages, KP, sa.DP);
  s_stats<KP, false>(sa, sm, s_part<KP, false>(sa, (int)blockIdx.x, 0, 1), (int)blockIdx.x, (int)gridDim.x);
}

// start of an evaluation: snapshot of the loop state through L2; false when the loop is over (terminated, or the launch's budget used)
__device__ __noinline__ bool loop_begin_eval(LoopShared &hs) {
  __syncthreads();
  if (threadIdx.x < (int)(sizeof(gpc_loop_t) / 4))
    reinterpret_cast<int *>(&hs.L)[threadIdx.x] = __ldcg(reinterpret_cast<const int *>(hs.loop_g) + threadIdx.x);
  __syncthreads();
  const bool go = hs.L.done == 0 && hs.ev < hs.max_evals;
  __syncthreads();
  if (threadIdx.x == 0) {
    hs.ev += 1;
    hs.retry = hs.L.need_retry != 0;
    if (hs.timing != nullptr && blockIdx.x == 0) hs.t[0] = hs.t[1] = hs.t[2] = hs.t[3] = globaltimer_ns();
  }
  __syncthreads();
  return go;
}
// end of the G rows: per-CTA partial dots, ring barriers invalidated (the memory is re-carved by the S phase)
__device__ __noinline__ void loop_after_g(LoopShared &hs, unsigned char *smem_raw) {
  const G2Smem gm = g2_carve(smem_raw, hs.g_stages, hs.KP, hs.DP);
  g_cta_partial(gm.red_s, const_cast<double *>(hs.gpart));
  if (threadIdx.x == 0) {
    for (int s = 0; s < hs.g_stages; ++s) {
      mbar_inval(&gm.full_bar[s]);
      mbar_inval(&gm.a_bar[s]);
      mbar_inval(&gm.empty_bar[s]);
    }
    if (hs.timing != nullptr && blockIdx.x == 0) hs.t[1] = hs.t[2] = globaltimer_ns();
  }
}
// end of the S rows: entropy partial of this CTA, ring barriers invalidated
__device__ __noinline__ void loop_after_s(LoopShared &hs, unsigned char *smem_raw) {
  const SSmem sm = s_carve(smem_raw, hs.s_stages, hs.KP, hs.DP);
  if (threadIdx.x == 0) {
    double v = 0.0;
    for (int w = 0; w < kSSoftmaxWarps; ++w) v += sm.red_s[w];
    double *h = hs.spart_mine + (size_t)hs.KP * hs.DP + hs.KP;
    h[0] = v;
    h[1] = 0.0;
    for (int s = 0; s < hs.s_stages; ++s) {
      mbar_inval(&sm.full_bar[s]);
      mbar_inval(&sm.phi_bar[s]);
      mbar_inval(&sm.empty_bar[s]);
    }
    if (hs.timing != nullptr && blockIdx.x == 0) hs.t[3] = globaltimer_ns();
  }
}

template <int KP, bool FULLK>
__global__ void __launch_bounds__(kGThreads, 1) k_mohgp_loop(const __grid_constant__ LoopKernelArgs a) {
  static_assert(kGThreads == kSThreads, "both row phases use the same CTA shape");
  extern __shared__ __align__(128) unsigned char smem_all[];
  LoopShared &hs = *reinterpret_cast<LoopShared *>(smem_all);
  unsigned char *smem_raw = smem_all + kLoopHeaderBytes;
  const NatgradArgs &ga = a.g;
  const StepStatsArgs &sa = a.s;
  if (threadIdx.x == 0) {
    const gpc_peers_t *P = a.peers;
    hs.p = a.p;
    hs.peers = P;
    hs.W = P ? P->world : 1;
    hs.my_rank = P ? P->rank : 0;
    hs.inbox_self = P ? P->inbox[P->rank] : nullptr;
    hs.sync = a.sync;
    hs.timing = a.timing;
    hs.gpart = ga.partials;
    hs.spart = sa.partials;
    hs.dots_out = ga.dots_out;
    hs.len = KP * ga.DP + KP + 2;
    hs.ev = 0;
    hs.max_evals = a.max_evals;
    hs.retry = 0;
    hs.g_stages = ga.stages;
    hs.s_stages = sa.stages;
    hs.KP = KP;
    hs.DP = ga.DP;
    hs.loop_g = a.p.loop;
    hs.spart_mine = sa.partials + (size_t)blockIdx.x * (size_t)(KP * ga.DP + KP + 2);
    hs.abort = 0;
    hs.bar_target = 0u;
    hs.e_stats = P ? __ldcg(P->epochs + 0) : 0ull;
    hs.e_dots = P ? __ldcg(P->epochs + 1) : 0ull;
    for (int q = 0; q < 4; ++q) hs.t_acc[q] = 0ull;
    for (int q = 0; q < 7; ++q) hs.p_acc[q] = 0ull;
    for (int q = 0; q < 8; ++q) hs.dbg_acc[q] = 0ull;
    hs.n_eval = hs.n_g = 0ull;
  }

  while (loop_begin_eval(hs)) {
    // ------------------------------------------------------------------ G rows (skipped in a retry slot: same point, same gradient)
    if (!hs.retry) {
      loop_g_phase<KP, FULLK>(ga, hs, smem_raw);
      loop_after_g(hs, smem_raw);
    }

    // ------------------------------------------------------------------ S rows (the producer runs ahead of B1)
    {
      const SSmem sm = s_carve(smem_raw, sa.stages, KP, sa.DP);
      const SPtrs sp = s_resolve(sa, KP, 0, &hs.L, hs.retry ? GPC_PHASE_RETRY : GPC_PHASE_CONJ);  // slot mode: the phase follows need_retry
      __syncthreads();  // every warp is done with the G phase's shared memory (and thread 0 with the invalidation)
      s_init_barriers<false>(sm, sa.stages);
      __syncthreads();
      const int warp = threadIdx.x >> 5;
      if (warp == kSSoftmaxWarps + kSStatWarps) {
        // beta is not known yet: the old search direction is staged whenever the step may use it
        s_producer<KP>(sa, sm, sp, (int)blockIdx.x, (int)gridDim.x, true, sp.beta_mode != GPC_BETA_ZERO);
      } else {
        double beta = 0.0;
        if (!hs.retry) {
          loop_dots_phase(hs);
          beta = cg_beta(sp.beta_mode, hs.dots[0], hs.dots[1], hs.dots[2], sp.sq_old);
          if (blockIdx.x == 0 && threadIdx.x == 0 && sp.beta_out) *sp.beta_out = beta;
        }
        if (warp < kSSoftmaxWarps) loop_s_softmax<KP, FULLK>(sa, hs, smem_raw, beta);
        else loop_s_stats<KP>(sa, smem_raw);
      }
      fence_proxy_async();  // x_new / sd written here are staged by this CTA's bulk copies in the next G / S phase
      __syncthreads();
    }
    loop_after_s(hs, smem_raw);

    // ------------------------------------------------------------------ B2, posterior, decision, B3
    if (!loop_post_phase(hs, smem_raw, hs.retry ? GPC_PHASE_RETRY : GPC_PHASE_CONJ)) break;
  }

  // ---- exit: the last CTA to leave re-arms the counters for the next launch; CTA 0 stores the exchange epochs
  __syncthreads();
  if (threadIdx.x == 0) {
    if (hs.abort) atomicOr(&a.p.loop->done, GPC_DONE_PEER_TIMEOUT);
    if (a.peers != nullptr && blockIdx.x == 0) {
      a.peers->epochs[0] = hs.e_stats;
      a.peers->epochs[1] = hs.e_dots;
    }
    if (a.timing != nullptr && blockIdx.x == 0) {
      for (int q = 0; q < 4; ++q) a.timing[q] += hs.t_acc[q];
      a.timing[4] += hs.n_eval;
      a.timing[5] += hs.n_g;
      for (int q = 0; q < 7; ++q) a.timing[6 + q] += hs.p_acc[q];
      for (int q = 0; q < 6; ++q) a.timing[16 + q] += hs.dbg_acc[q];
    }
    __threadfence();
    const unsigned int left = atomicAdd(a.sync + 1, 1u);
    if (left == gridDim.x - 1) {
      a.sync[0] = 0u;
      a.sync[1] = 0u;
      a.sync[2] = 0u;
      __threadfence();
    }
  }
}

}  // namespace gpc
