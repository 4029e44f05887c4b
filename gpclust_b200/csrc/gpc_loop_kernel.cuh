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

constexpr int kLoopHeaderBytes = 512;  // loop-state snapshot + scalars at the start of dynamic shared memory
constexpr long long kSpinLimitNs = 4000000000LL;  // bounded waits (grid barrier, peer flags): 4 s

// ---- inbox layout of the fused exchange (doubles / u64 words), W ranks, statistics length L (even), up to 64 clusters
//   [2][W][L] statistics | [2][W][4] dots | [2][W][64] statistics flags (per sender and cluster) | [2][W] dots flags
__host__ __device__ inline long long fx_off_stats(int W, int len, int slot, int sender) { return ((long long)slot * W + sender) * xchg_l2(len); }
__host__ __device__ inline long long fx_off_dots(int W, int len, int slot, int sender) { return 2LL * W * xchg_l2(len) + ((long long)slot * W + sender) * 4; }
__host__ __device__ inline long long fx_off_sflag(int W, int len, int slot, int sender, int k) {
  return 2LL * W * xchg_l2(len) + 2LL * W * 4 + ((long long)slot * W + sender) * GPC_MAX_KP + k;
}
__host__ __device__ inline long long fx_off_dflag(int W, int len, int slot, int sender) {
  return 2LL * W * xchg_l2(len) + 2LL * W * 4 + 2LL * W * GPC_MAX_KP + (long long)slot * W + sender;
}
__host__ __device__ inline long long fx_doubles(int W, int len) { return 2LL * W * xchg_l2(len) + 2LL * W * 4 + 2LL * W * GPC_MAX_KP + 2LL * W; }

struct LoopKernelArgs {
  NatgradArgs g;    // F, Wt, bias, partials (grid x 4), dots_out, N, K, DP, num_groups, stages, bufs      (loop / peers unused)
  StepStatsArgs s;  // F, beta_out, partials (grid x stats_len), N, K, DP, num_groups, stages, bufs         (loop / peers / dots unused)
  MohgpPostArgs p;  // stats, ws, Sy_inv, Sf, const_term, Wt, muk_t, ..., info, alpha, prior, K, D, KP, DP, dots, beta, loop (GLOBAL)
  const gpc_peers_t *peers;     // nullable; inbox layout fx_* above
  unsigned int *sync;           // [0] grid-barrier counter, [1] exit counter, [2] posterior ticket -- all zero between launches
  unsigned long long *timing;   // nullable: [0..3] ns spent in G rows / B1 + dots / S rows / B2 + posterior + B3, [4] evaluations, [5] G phases
  int max_evals;
};

struct LoopShared {  // header of dynamic shared memory (long-lived scalars live here, not in registers of the row loops)
  gpc_loop_t L;      // snapshot of the device loop state, taken once per evaluation through L2
  double dots[4];
  unsigned long long e_stats, e_dots;  // exchange epochs (uniform over the grid)
  unsigned long long t_acc[4], n_eval, n_g, t[4];
  unsigned int bar_target;
  int abort;
};
static_assert(sizeof(LoopShared) <= kLoopHeaderBytes, "loop header");

// Grid-wide barrier among all CTAs of the (cooperative) launch.  `nthreads` threads of this CTA take part (named barrier `bar_id`;
// 0 = the whole CTA).  Monotone counter; `*target_s` (shared memory, thread 0) is the value the counter reaches when everybody arrived.
// Returns false when the wait ran into the spin limit (a CTA of this grid is stuck in a peer wait that timed out elsewhere).
template <int BAR_ID, int NTHREADS>
__device__ __forceinline__ bool grid_barrier(unsigned int *ctr, unsigned int *target_s, int *abort_s) {
  asm volatile("bar.sync %0, %1;\n" ::"n"(BAR_ID), "n"(NTHREADS) : "memory");
  if (threadIdx.x == 0 && *abort_s == 0) {  // once a wait of this CTA has timed out no further barrier spins: the loop is being left
    const unsigned int target = *target_s + gridDim.x;
    *target_s = target;
    __threadfence();
    red_release_gpu_add_u32(ctr, 1u);
    const unsigned long long t0 = globaltimer_ns();
    int spins = 0;
    while ((int)(ld_acquire_gpu_u32(ctr) - target) < 0) {
      if (((++spins) & 1023) == 0 && (long long)(globaltimer_ns() - t0) > kSpinLimitNs) {
        *abort_s = 1;
        break;
      }
    }
  }
  asm volatile("bar.sync %0, %1;\n" ::"n"(BAR_ID), "n"(NTHREADS) : "memory");
  return *reinterpret_cast<volatile int *>(abort_s) == 0;
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

template <int KP, bool FULLK>
__global__ void __launch_bounds__(kGThreads, 1) k_mohgp_loop(const LoopKernelArgs a) {
  static_assert(kGThreads == kSThreads, "both row phases use the same CTA shape");
  extern __shared__ __align__(128) unsigned char smem_all[];
  LoopShared &hs = *reinterpret_cast<LoopShared *>(smem_all);
  unsigned char *smem_raw = smem_all + kLoopHeaderBytes;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cid = (int)blockIdx.x, ncl = (int)gridDim.x;
  const int DP = a.g.DP, D = a.p.D;
  const int len = KP * DP + KP + 2;
  constexpr int kRowThreads = (kSSoftmaxWarps + kSStatWarps) * 32;  // the S phase's consumers (everything but the producer warp)
  const bool is_producer = (warp == kSSoftmaxWarps + kSStatWarps);
  const gpc_peers_t *P = a.peers;
  const int W = P ? P->world : 1, my_rank = P ? P->rank : 0;
  double *inbox_self = P ? P->inbox[my_rank] : nullptr;
  const bool timekeeper = (a.timing != nullptr) && cid == 0 && tid == 0;
  if (tid == 0) {
    hs.abort = 0;
    hs.bar_target = 0u;
    hs.e_stats = P ? __ldcg(P->epochs + 0) : 0ull;
    hs.e_dots = P ? __ldcg(P->epochs + 1) : 0ull;
    for (int q = 0; q < 4; ++q) hs.t_acc[q] = 0ull;
    hs.n_eval = hs.n_g = 0ull;
  }

  const NatgradArgs &ga = a.g;
  const StepStatsArgs &sa = a.s;

  for (int ev = 0; ev < a.max_evals; ++ev) {
    // ---- loop state of this evaluation (written by the previous evaluation's decision, or by the host), through L2
    __syncthreads();
    if (tid < (int)(sizeof(gpc_loop_t) / 4)) reinterpret_cast<int *>(&hs.L)[tid] = __ldcg(reinterpret_cast<const int *>(a.p.loop) + tid);
    __syncthreads();
    if (hs.L.done != 0) break;
    const bool retry = (hs.L.need_retry != 0);
    const int phase = retry ? GPC_PHASE_RETRY : GPC_PHASE_CONJ;
    if (timekeeper) hs.t[0] = hs.t[1] = hs.t[2] = hs.t[3] = globaltimer_ns();

    // ------------------------------------------------------------------ G rows
    if (!retry) {
      const GSmem gm = g_carve(smem_raw, ga.stages, KP, DP);
      const GPtrs gp = g_resolve(ga, KP, 0, &hs.L);
      g_setup<KP, false>(ga, gm, gp);
      __syncthreads();
      g_rows<KP, FULLK, false>(ga, gm, gp, cid, ncl, 0, 1);
      fence_proxy_async();  // this CTA's natural gradient is staged by its own bulk copies in the S phase
      __syncthreads();
      g_cta_partial(gm, ga.partials);
      if (tid == 0)
        for (int s = 0; s < ga.stages; ++s) mbar_inval(&gm.full_bar[s]);  // the ring memory is re-carved below
      if (timekeeper) hs.t[1] = hs.t[2] = globaltimer_ns();
    }

    // ------------------------------------------------------------------ S rows (the producer runs ahead of B1)
    const SSmem sm = s_carve(smem_raw, sa.stages, KP, DP);
    const SPtrs sp = s_resolve(sa, KP, 0, &hs.L, phase);  // slot mode: the phase follows need_retry of the snapshot
    const SPart part = s_part<KP, false>(sa, cid, 0, 1);
    __syncthreads();  // every warp is done with the G phase's shared memory (and thread 0 with the invalidation)
    s_init_barriers<false>(sm, sa.stages);
    __syncthreads();
    const bool has_ng = true;
    const bool load_old = (sp.beta_mode != GPC_BETA_ZERO);  // beta is not known yet: stage the old search direction whenever it may be used
    if (is_producer) {
      s_producer<KP>(sa, sm, sp, cid, ncl, has_ng, load_old);
    } else {
      double beta = 0.0;
      if (!retry) {
        // B1 among the consumer warps: every CTA's partial dots are in L2
        if (!grid_barrier<3, kRowThreads>(a.sync, &hs.bar_target, &hs.abort)) {
          // fall through to the block barrier below with abort set; the loop ends there
        }
        if (P == nullptr) {
          if (warp < 3) {
            double v = 0.0;
            for (int c = lane; c < ncl; c += 32) v += __ldcg(ga.partials + (size_t)c * 4 + warp);
            v = warp_sum(v);  // fixed lane -> CTA assignment and fixed shuffle tree: identical on every CTA, run to run
            if (lane == 0) {
              hs.dots[warp] = v;
              if (cid == 0) ga.dots_out[warp] = v;
            }
          }
        } else {
          const unsigned long long e_dots = hs.e_dots + 1;  // uniform: thread 0 stores it back after the barrier below
          const int slot = (int)(e_dots & 1);
          if (cid == 0) {
            if (warp < 3) {
              double v = 0.0;
              for (int c = lane; c < ncl; c += 32) v += __ldcg(ga.partials + (size_t)c * 4 + warp);
              v = warp_sum(v);
              if (lane < W) P->inbox[lane][fx_off_dots(W, len, slot, my_rank) + warp] = v;  // lane r -> peer r, over NVLink
            }
            asm volatile("bar.sync 3, %0;\n" ::"n"(kRowThreads) : "memory");
            if (warp == 0) {
              __threadfence_system();
              if (lane < W) st_relaxed_sys(reinterpret_cast<unsigned long long *>(P->inbox[lane]) + fx_off_dflag(W, len, slot, my_rank), e_dots);
            }
          }
          if (warp == 0) {
            bool ok = true;
            if (lane < W) ok = flag_wait(reinterpret_cast<const unsigned long long *>(inbox_self) + fx_off_dflag(W, len, slot, lane), e_dots);
            ok = __all_sync(0xffffffffu, ok);
            if (lane < 3) {
              double d = 0.0;
              for (int r = 0; r < W; ++r) d += __ldcg(inbox_self + fx_off_dots(W, len, slot, r) + lane);  // rank order: identical on all ranks
              hs.dots[lane] = ok ? d : nan("");
              if (cid == 0) ga.dots_out[lane] = hs.dots[lane];
            }
            if (!ok && lane == 0) atomicOr(&a.p.loop->done, GPC_DONE_PEER_TIMEOUT);
          }
        }
        asm volatile("bar.sync 3, %0;\n" ::"n"(kRowThreads) : "memory");
        if (P != nullptr && tid == 0) hs.e_dots += 1;  // every consumer has read the old value before the barrier above
        beta = cg_beta(sp.beta_mode, hs.dots[0], hs.dots[1], hs.dots[2], sp.sq_old);
        if (cid == 0 && tid == 0 && sp.beta_out) *sp.beta_out = beta;
        if (timekeeper) hs.t[2] = globaltimer_ns();
      }
      const bool use_old = (beta != 0.0);
      if (warp < kSSoftmaxWarps) s_softmax<KP, FULLK, false>(sa, sm, sp, cid, ncl, 0, 1, beta, has_ng, use_old);
      else s_stats<KP, false>(sa, sm, part, cid, ncl);
    }
    fence_proxy_async();  // x_new / sd written here are staged by this CTA's bulk copies in the next G / S phase
    __syncthreads();
    s_cta_entropy<false>(sa, sm, part, 0);
    if (tid == 0)
      for (int s = 0; s < sa.stages; ++s) {
        mbar_inval(&sm.full_bar[s]);
        mbar_inval(&sm.phi_bar[s]);
        mbar_inval(&sm.empty_bar[s]);
      }
    if (timekeeper) hs.t[3] = globaltimer_ns();

    // ------------------------------------------------------------------ B2, posterior
    if (!grid_barrier<0, kGThreads>(a.sync, &hs.bar_target, &hs.abort)) break;
    const unsigned long long e_stats = hs.e_stats + 1;  // read by every thread before thread 0 bumps it (after the next barrier)
    if (cid < KP && tid < kClusterThreads) {
      const int k = cid;
      double *ws_s = reinterpret_cast<double *>(smem_raw);
      PostSmem &ps = *reinterpret_cast<PostSmem *>(smem_raw + (((size_t)eig_ws_len(D) * 8 + 127) & ~(size_t)127));
      if (tid == 0) post_stage_ws(a.p.ws, D, ps, ws_s);
      double phi_hat = post_reduce_row<true>(sa.partials, ncl, len, k, KP, DP, ps);
      const double *hparts = sa.partials;
      int hnum = ncl;
      if (P != nullptr) {
        const int slot = (int)(e_stats & 1);
        // this rank's row of cluster k (+ its entropy, with cluster 0) -> every peer's inbox; one flag per (sender, cluster)
        double Hloc = 0.0;
        if (k == 0) {
          for (int c = tid; c < ncl; c += kClusterThreads) Hloc += __ldcg(sa.partials + (size_t)c * len + (size_t)KP * DP + KP);
          Hloc = block_sum<true>(Hloc, ps.red_s);
        }
        const long long base = fx_off_stats(W, len, slot, my_rank);
        for (int r = 0; r < W; ++r) {
          double *dst = P->inbox[r] + base;
          if (tid < DP) dst[(size_t)k * DP + tid] = ps.yb_s[tid];
          if (tid == 64) dst[(size_t)KP * DP + k] = phi_hat;
          if (tid == 65 && k == 0) dst[(size_t)KP * DP + KP] = Hloc;
        }
        psync<true>();
        if (warp == 0) {
          __threadfence_system();
          if (lane < W) st_relaxed_sys(reinterpret_cast<unsigned long long *>(P->inbox[lane]) + fx_off_sflag(W, len, slot, my_rank, k), e_stats);
          bool ok = true;
          if (lane < W) ok = flag_wait(reinterpret_cast<const unsigned long long *>(inbox_self) + fx_off_sflag(W, len, slot, lane, k), e_stats);
          ok = __all_sync(0xffffffffu, ok);
          if (!ok && lane == 0) {
            atomicOr(a.p.info, GPC_INFO_PEER_TIMEOUT);
            atomicOr(&a.p.loop->done, GPC_DONE_PEER_TIMEOUT);
          }
        }
        psync<true>();
        hparts = inbox_self + fx_off_stats(W, len, slot, 0);
        hnum = W;
        phi_hat = post_reduce_row<true>(hparts, W, len, k, KP, DP, ps);  // sum over the ranks, in rank order
      }
      if (tid < DP) a.p.stats[(size_t)k * DP + tid] = ps.yb_s[tid];
      if (tid == 0) a.p.stats[(size_t)KP * DP + k] = phi_hat;
      mbar_wait(&ps.stage_bar, 0);
      post_cluster_math<true>(a.p, k, phi_hat, ps, ws_s);
      __threadfence();
      psync<true>();
      if (tid == 0) {
        const unsigned int done = atomicAdd(a.sync + 2, 1u);
        ps.is_last = (done == (unsigned)KP - 1) ? 1 : 0;
        mbar_inval(&ps.stage_bar);
      }
      psync<true>();
      if (ps.is_last) {
        __threadfence();
        if (P != nullptr && warp == 0) {
          // the entropies travel with cluster 0's rows: make sure this CTA has observed those flags itself
          const int slot = (int)(e_stats & 1);
          if (lane < W) flag_wait(reinterpret_cast<const unsigned long long *>(inbox_self) + fx_off_sflag(W, len, slot, lane, 0), e_stats);
        }
        psync<true>();
        MohgpPostArgs pa = a.p;
        pa.counter = a.sync + 2;  // the tail resets the ticket
        post_tail<true>(pa, ps, hparts, hnum, len, phase);
        __threadfence();
      }
    }
    // ------------------------------------------------------------------ B3: the decision and the new weights are visible
    if (!grid_barrier<0, kGThreads>(a.sync, &hs.bar_target, &hs.abort)) break;
    if (tid == 0) hs.e_stats = e_stats;
    if (timekeeper) {
      const unsigned long long t4 = globaltimer_ns();
      hs.t_acc[0] += hs.t[1] - hs.t[0];
      hs.t_acc[1] += hs.t[2] - hs.t[1];
      hs.t_acc[2] += hs.t[3] - hs.t[2];
      hs.t_acc[3] += t4 - hs.t[3];
      hs.n_eval += 1;
      hs.n_g += retry ? 0 : 1;
    }
  }

  // ---- exit: the last CTA to leave re-arms the counters for the next launch; CTA 0 stores the exchange epochs
  __syncthreads();
  if (tid == 0) {
    if (hs.abort) atomicOr(&a.p.loop->done, GPC_DONE_PEER_TIMEOUT);
    if (P != nullptr && cid == 0) {
      P->epochs[0] = hs.e_stats;
      P->epochs[1] = hs.e_dots;
    }
    if (timekeeper) {
      for (int q = 0; q < 4; ++q) a.timing[q] += hs.t_acc[q];
      a.timing[4] += hs.n_eval;
      a.timing[5] += hs.n_g;
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
