// gpclust_b200 -- the two N-sized passes of one collapsed-VB iteration (FP64, sm_100a).
//
//   k_natgrad    : G pass.  a[n,k] = F[n,:] . W[:,k] + bias[k] - x[n,k]   (DMMA)
//                  natgrad = a - sum_k phi*a ; grad = natgrad*phi ; the three CG dot products.
//                  Restates MOHGP.py:137-153 (multiple_mahalanobis utilities.py:108-172 in GEMM form,
//                  row-constant terms dropped because the row-centering at MOHGP.py:150 removes them),
//                  MOG.py:72-84 and the dots of collapsed_vb.py:154,160-164.
//   k_step_stats : S pass.  searchDir = natgrad + beta*searchDir_old ; x_new = x + step*searchDir
//                  (collapsed_vb.py:167,172) fused with the softmax / entropy / phi_hat /
//                  phi^T F sufficient statistics of collapsed_mixture.py:53-56, MOHGP.py:76, MOG.py:56-57.
//
// HBM layout ("fragment-major", private to this library -- gpc_pack_rows / gpc_unpack_rows translate):
// rows are handled in GROUPS of 8 (the M of mma.m8n8k4).  Within a group
//   N x K arrays (logits, natural gradient, search direction):  element (r, c) of the group sits at
//       ((c/8)*32 + r*4 + (c%8)/2)*2 + c%2          -- the DMMA C-fragment order: lane (g=r, t) owns columns 8nt+2t, +1
//   features F:                                                  element (r, d) sits at (d/4)*32 + r*4 + d%4
//                                                                -- the DMMA A-fragment order: lane (g=r, t) owns column 4ks+t
// so one group of every operand is ONE contiguous 16-B-aligned block (4 KB at K = D = 64) that a single
// cp.async.bulk lands in shared memory exactly in the order the lanes read it: conflict-free LDS, no padding
// bytes in HBM, and the global stores of a warp are 512 contiguous bytes.
//
// Both kernels are persistent, one CTA per SM (12 warps, 168 registers each), built around rings of group-sized
// shared-memory stages filled by cp.async.bulk + mbarrier (complete_tx):
//   G pass: six units of two warps; a unit owns one 8-row group at a time (warp h: clusters [32h, 32h+32)) and feeds
//           its own 3-4 ring slots [F | x | lse]; the previous-point operands go straight from HBM to registers; W stays
//           resident in shared memory in B-fragment order.
//   S pass: 1 producer warp (13-stage ring of [F | x | ng | sd]) + 7 softmax warps (CG step, softmax, entropy; leave phi
//           in the stage in A-fragment order) + 4 statistics warps (every group: T += phi^T F on the tensor pipe, 16 of the
//           64 8x8 tiles each).
// Consumers never wait on HBM, only on the ring.  The S pass leaves lse[n] = logsumexp_k x[n,:] behind so that the G pass
// gets phi = exp(x - lse) without row reductions.
// K > 64 ("wide" mode, template parameter CL): the same kernels launched as thread-block clusters of KP/64 CTAs.  The CTAs
// of a cluster walk the same row groups, CTA r owning clusters [64r, 64r+64); the row-coupled scalars (softmax max / sum /
// sum e*(x-max); sum_k phi*a) are exchanged through distributed shared memory: st.async into every peer's mailbox, counted
// by complete_tx on the peer's mbarrier, two mailbox parities per unit / softmax warp.
// Device-resident loop (gpc_loop_t), peer exchange over NVLink (gpc_peers_t): see include/gpclust_b200.h.
#pragma once
#include "gpc_common.cuh"

namespace gpc {

constexpr int kGroupRows = 8;
constexpr int kCtasPerSm = 1;
constexpr int kGConsumers = 12;                    // G pass: every warp computes (6 units of 2; 12 warps -> 168 registers each)
constexpr int kGThreads = kGConsumers * 32;
constexpr int kSSoftmaxWarps = 7;                  // S pass (7 + 4 + 1 = 12 warps -> 168 registers each)
constexpr int kSStatWarps = 4;
constexpr int kSThreads = (kSSoftmaxWarps + kSStatWarps + 1) * 32;
constexpr int kSmemBudget = 227 * 1024;
constexpr int kMaxCl = GPC_MAX_KTOT / GPC_MAX_KP;  // CTAs per cluster in the wide (K > 64) mode: one 64-cluster chunk each

// ------------------------------------------------------------------------------------------------
// G pass
// ------------------------------------------------------------------------------------------------
struct NatgradArgs {
  const double *F;        // groups x (8*DP) features, A-fragment order
  const double *x;        // groups x (8*KP) logits at the current point, C-fragment order
  const double *lse;      // Npad : logsumexp of the rows of x (written by the S pass that produced x)
  const double *Wt;       // KP x DP row-major, Wt[k][d] = W[d][k]
  const double *bias;     // KP
  const double *x_old;    // logits at the previous accepted point (nullable -> no HS/PR dots)
  const double *lse_old;  // Npad : logsumexp of the rows of x_old
  const double *ng_old;   // natural gradient at the previous accepted point (nullable)
  double *ng_out;         // natural gradient (ascent direction, un-negated), C-fragment order
  double *grad_out;       // nullable: gradient = natgrad * phi
  double *partials;       // gridDim.x * 4 doubles
  unsigned int *counter;  // zero on entry; reset to zero on exit
  double *dots_out;       // [0]=natgrad.grad  [1]=(ng-ng_old).grad  [2]=(ng-ng_old).grad_old
  int N, K, DP, num_groups, stages;
  const gpc_loop_t *loop;  // nullable: device-resident loop state; the N x K operands then come from `bufs`
  gpc_bufs_t bufs;
  const gpc_peers_t *peers;  // nullable: push the three dots into every peer's inbox (rows sharded over GPUs)
  // K > 64: the clusters are processed in chunks of 64, one launch per chunk (pointers pre-offset to the chunk).
  // xg_stride: doubles between consecutive 8-row groups of the N x K operands (0 = 8*KP: single chunk).
  // sa_vec != NULL selects sweep 1 of the chunked pass: ng_out receives a = F.W + bias - x (not yet centred) and
  // sa_vec[row] accumulates sum_k phi*a over the chunks (sa_first: this launch initialises it); no dots.
  long long xg_stride;
  double *sa_vec;
  int sa_first;
  // wide mode (template CL): launched with clusters of `csize` CTAs; K is the total cluster count, Wt / bias / the N x K
  // operands are the full KPtot = 64*csize wide arrays (xg_stride = 8*KPtot) and CTA `rank` of a cluster works on the
  // columns [64 rank, 64 rank + 64) of the row groups of its cluster.
  int csize;
};

// ring stage: [F group | x group | lse (8) | lse_old (8)]; the previous-point operands x_old / ng_old go straight
// from HBM to registers (the fragment-major layout makes those loads 512 contiguous bytes per warp instruction)
__host__ __device__ inline int g_stage_doubles(int KP, int DP) { return kGroupRows * DP + kGroupRows * KP + 2 * kGroupRows; }
__host__ __device__ inline int g_fixed_doubles(int KP, int DP) { return KP * DP + KP + kGConsumers * 4 + 2 * kGConsumers * kGroupRows; }
// wide mode: mailbox [2 parities][units][kMaxCl ranks][8 rows] + 2*units mbarriers
__host__ __device__ inline int g_cl_doubles() { return 2 * kGConsumers * kMaxCl * kGroupRows + 2 * kGConsumers; }

// shared-memory carve of the G pass (dynamic shared memory; `stages` ring slots)
struct GSmem {
  double *stages, *w_s, *bias_s, *red_s, *sax_s, *sa_mb;
  uint64_t *full_bar, *mb_bar;
};
__device__ __forceinline__ GSmem g_carve(unsigned char *smem_raw, int stages, int KP, int DP) {
  GSmem m;
  m.stages = reinterpret_cast<double *>(smem_raw);
  m.w_s = m.stages + (size_t)stages * g_stage_doubles(KP, DP);  // B fragments: ((ks*NT + nt)*32 + lane)
  m.bias_s = m.w_s + KP * DP;                                    // KP
  m.red_s = m.bias_s + KP;                                       // kGConsumers * 4
  m.sax_s = m.red_s + kGConsumers * 4;                           // [2 buffers][UNITS][NSPLIT][8] row-sum exchange
  m.full_bar = reinterpret_cast<uint64_t *>(m.sax_s + 2 * kGConsumers * kGroupRows);
  m.mb_bar = m.full_bar + stages;                                           // wide mode: [2][UNITS]
  m.sa_mb = reinterpret_cast<double *>(m.mb_bar + 2 * kGConsumers);         // wide mode: [2][UNITS][kMaxCl][8]
  return m;
}
// operands of one G pass, resolved from the launch arguments or from the device-resident loop state
struct GPtrs {
  const double *x, *lse, *x_old, *lse_old, *ng_old, *Wt, *bias;
  double *ng_out, *grad_out;
};
// `loop`: the loop state to resolve against (a.loop for the stand-alone kernels; the fused loop kernel passes its per-evaluation
// snapshot in shared memory)
__device__ __forceinline__ GPtrs g_resolve(const NatgradArgs &a, int KP, int crank, const gpc_loop_t *loop) {
  const size_t coff = (size_t)crank * (kGroupRows * KP);  // this CTA's chunk inside a group of the N x K operands (wide mode)
  GPtrs p;
  p.x = a.x + coff;
  p.lse = a.lse;
  p.x_old = a.x_old ? a.x_old + coff : nullptr;
  p.lse_old = a.lse_old;
  p.ng_old = a.ng_old ? a.ng_old + coff : nullptr;
  p.ng_out = a.ng_out + coff;
  p.grad_out = a.grad_out ? a.grad_out + coff : nullptr;
  p.Wt = a.Wt + (size_t)crank * KP * a.DP;
  p.bias = a.bias + crank * KP;
  if (loop != nullptr) {
    const int xi = loop->xi, xp = loop->xprev, ngi = loop->ngi, method = loop->method;
    p.x = sel3(a.bufs.x[0], a.bufs.x[1], a.bufs.x[2], xi) + coff;
    p.lse = sel3(a.bufs.lse[0], a.bufs.lse[1], a.bufs.lse[2], xi);
    p.ng_out = (ngi ? a.bufs.ng[1] : a.bufs.ng[0]) + coff;
    // HS / PR need the previous accepted point (collapsed_vb.py:159-164); none on the first pass
    const bool ho = (loop->first == 0) && xp >= 0 && (method == GPC_BETA_HS || method == GPC_BETA_PR);
    p.x_old = ho ? sel3(a.bufs.x[0], a.bufs.x[1], a.bufs.x[2], xp) + coff : nullptr;
    p.lse_old = ho ? sel3(a.bufs.lse[0], a.bufs.lse[1], a.bufs.lse[2], xp) : nullptr;
    p.ng_old = ho ? (ngi ? a.bufs.ng[0] : a.bufs.ng[1]) + coff : nullptr;
  }
  return p;
}
// barriers, resident weights (B-fragment order) and bias; the caller synchronises the block (cluster) afterwards
template <int KP, bool CL>
__device__ __forceinline__ void g_setup_bars(const NatgradArgs &a, const GSmem &m) {
  constexpr int NT = KP / 8;
  constexpr int UNITS = kGConsumers / ((NT >= 2) ? 2 : 1);
  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) mbar_init(&m.full_bar[s], 1);
    if (CL)
      for (int s = 0; s < 2 * UNITS; ++s) mbar_init(&m.mb_bar[s], 1);
    mbar_fence_init();
  }
}
template <int KP>
__device__ __forceinline__ void g_setup_w(const NatgradArgs &a, const GSmem &m, const GPtrs &p);
template <int KP, bool CL>
__device__ __forceinline__ void g_setup(const NatgradArgs &a, const GSmem &m, const GPtrs &p) {
  g_setup_bars<KP, CL>(a, m);
  g_setup_w<KP>(a, m, p);
}
template <int KP>
__device__ __forceinline__ void g_setup_w(const NatgradArgs &a, const GSmem &m, const GPtrs &p) {
  constexpr int NT = KP / 8;
  const int DP = a.DP;
  for (int i = threadIdx.x; i < KP * DP; i += blockDim.x) {
    const int ln = i & 31, f = i >> 5;  // f = ks*NT + nt
    const int nt = f % NT, ks = f / NT;
    m.w_s[i] = __ldcg(p.Wt + (size_t)(8 * nt + (ln >> 2)) * DP + 4 * ks + (ln & 3));  // L2: rewritten by other CTAs inside the fused loop kernel
  }
  for (int i = threadIdx.x; i < KP; i += blockDim.x) m.bias_s[i] = __ldcg(p.bias + i);
  for (int i = threadIdx.x; i < kGConsumers * 4; i += blockDim.x) m.red_s[i] = 0.0;
}

// FULLK: K == KP, no column masking.
// All 12 warps compute.  They work in units of NSPLIT warps that share one group: warp h of a unit owns the n-tiles
// [h*NTW, (h+1)*NTW) (clusters), so the unit's row reduction sum_k phi*a is exchanged through shared memory once per
// group.  Every unit feeds ITSELF: it owns R = stages / UNITS ring slots, and lane 0 of its first warp re-arms the
// slot it has just finished with the unit's group R visits ahead (cp.async.bulk + mbarrier complete_tx) -- no
// producer warp, no empty barriers (the unit's own exchange barrier orders "both warps are done reading" before
// the refill), and the four schedulers carry three compute warps each.
// The row loop of one CTA: on return red_s[warp*4 + {0,1,2}] hold the warp's partial dots (the caller synchronises).
// `cid` / `ncl`: index of this CTA (cluster of CTAs in wide mode) among those that share the row groups.
// the first R ring slots of every unit, armed by the unit's leader lane (what g_rows does on entry unless told `prefed`): the fused
// loop kernel issues them BEFORE it loads the weights, so that the first HBM round trip overlaps the L2 reads of W
template <int KP>
__device__ __forceinline__ void g_prefeed(const NatgradArgs &a, const GSmem &m, const GPtrs &p, int cid, int ncl) {
  constexpr int NT = KP / 8;
  constexpr int NSPLIT = (NT >= 2) ? 2 : 1;
  constexpr int UNITS = kGConsumers / NSPLIT;
  constexpr int XG = kGroupRows * KP;
  const int DP = a.DP, FG = kGroupRows * DP, R = a.stages / UNITS, stage_doubles = g_stage_doubles(KP, DP);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, unit = warp / NSPLIT, h = warp % NSPLIT;
  if (h != 0 || lane != 0) return;
  const bool has_old = (p.ng_old != nullptr);
  const size_t GS = a.xg_stride ? (size_t)a.xg_stride : (size_t)XG;
  const uint32_t bytes = (uint32_t)(FG + XG + kGroupRows + (has_old ? kGroupRows : 0)) * 8u;
  double *ustages = m.stages + (size_t)unit * R * stage_doubles;
  uint64_t *ubar = m.full_bar + unit * R;
  for (int i = 0; i < R; ++i) {
    const long long grp = (long long)cid + (long long)unit * ncl + (long long)i * UNITS * ncl;
    if (grp >= a.num_groups) break;
    double *st = ustages + (size_t)i * stage_doubles;
    mbar_arrive_expect_tx(&ubar[i], bytes);
    bulk_g2s(st, a.F + (size_t)grp * FG, FG * 8u, &ubar[i]);
    bulk_g2s(st + FG, p.x + (size_t)grp * GS, XG * 8u, &ubar[i]);
    bulk_g2s(st + FG + XG, p.lse + (size_t)grp * kGroupRows, kGroupRows * 8u, &ubar[i]);
    if (has_old) bulk_g2s(st + FG + XG + kGroupRows, p.lse_old + (size_t)grp * kGroupRows, kGroupRows * 8u, &ubar[i]);
  }
}

template <int KP, bool FULLK, bool CL>
__device__ __forceinline__ void g_rows(const NatgradArgs &a, const GSmem &m, const GPtrs &p, int cid, int ncl, int crank, int csize,
                                       bool prefed = false) {
  constexpr int NT = KP / 8;
  constexpr int NSPLIT = (NT >= 2) ? 2 : 1;
  constexpr int NTW = NT / NSPLIT;           // n-tiles per warp
  constexpr int UNITS = kGConsumers / NSPLIT;
  constexpr int XG = kGroupRows * KP;        // doubles of one N x K operand per group
  const int DP = a.DP, FG = kGroupRows * DP;
  const int R = a.stages / UNITS;            // ring slots per unit
  const int stage_doubles = g_stage_doubles(KP, DP);
  double *stages = m.stages, *w_s = m.w_s, *bias_s = m.bias_s, *red_s = m.red_s, *sax_s = m.sax_s, *sa_mb = m.sa_mb;
  uint64_t *full_bar = m.full_bar, *mb_bar = m.mb_bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const double exp_tab = kExpTab[lane];  // this lane's entry of the 2^(j/32) table (exp_batch_tab)
  const int Kc = FULLK ? KP : (CL ? min(KP, a.K - KP * crank) : a.K);  // real clusters among this CTA's KP columns
  const double *p_x = p.x, *p_lse = p.lse, *p_x_old = p.x_old, *p_lse_old = p.lse_old, *p_ng_old = p.ng_old;
  double *p_ng_out = p.ng_out, *p_grad_out = p.grad_out;
  const bool has_old = (p_ng_old != nullptr);
  const size_t GS = a.xg_stride ? (size_t)a.xg_stride : (size_t)XG;  // global group stride of the N x K operands
  (void)csize;
  {
    const int unit = warp / NSPLIT, h = warp % NSPLIT, nt0 = h * NTW;
    const int ksteps = DP / 4;
    const uint32_t bytes = (uint32_t)(FG + XG + kGroupRows + (has_old ? kGroupRows : 0)) * 8u;
    double *ustages = stages + (size_t)unit * R * stage_doubles;
    uint64_t *ubar = full_bar + unit * R;
    const int gstride = UNITS * ncl;
    const int grp0 = cid + unit * ncl;
    // arm ring slot `slot` with group `grp` (one elected lane of the unit)
    auto feed = [&](int slot, int grp) {
      double *st = ustages + (size_t)slot * stage_doubles;
      mbar_arrive_expect_tx(&ubar[slot], bytes);
      bulk_g2s(st, a.F + (size_t)grp * FG, FG * 8u, &ubar[slot]);
      bulk_g2s(st + FG, p_x + (size_t)grp * GS, XG * 8u, &ubar[slot]);
      bulk_g2s(st + FG + XG, p_lse + (size_t)grp * kGroupRows, kGroupRows * 8u, &ubar[slot]);
      if (has_old) bulk_g2s(st + FG + XG + kGroupRows, p_lse_old + (size_t)grp * kGroupRows, kGroupRows * 8u, &ubar[slot]);
    };
    if (!prefed && h == 0 && lane == 0)
      for (int i = 0; i < R; ++i) {
        const long long grp = (long long)grp0 + (long long)i * gstride;
        if (grp < a.num_groups) feed(i, (int)grp);
      }
    double d0 = 0.0, d1 = 0.0, d2 = 0.0;
    int i = 0;
    for (int grp = grp0; grp < a.num_groups; grp += gstride, ++i) {
      const int slot = i % R;
      const double *st = ustages + (size_t)slot * stage_doubles;
      const int row = grp * kGroupRows + g;
      const bool row_ok = row < a.N;
      const size_t goff = (size_t)grp * GS + (size_t)nt0 * 64 + 2 * lane;

      // previous-point operands: HBM -> registers, in flight behind the DMMA loop and pass 1
      double2 xo[NTW], ngo[NTW];
      if (has_old) {
#pragma unroll
        for (int q = 0; q < NTW; ++q) {
          xo[q] = ldg2_stream(p_x_old + goff + q * 64);
          ngo[q] = ldg2_stream(p_ng_old + goff + q * 64);
        }
      }
      mbar_wait(&ubar[slot], (uint32_t)((i / R) & 1));

      double acc[NTW][2];
#pragma unroll
      for (int q = 0; q < NTW; ++q) acc[q][0] = acc[q][1] = 0.0;
      const double *fa = st + lane;
      const double *wb = w_s + nt0 * 32 + lane;
#pragma unroll 4
      for (int ks = 0; ks < ksteps; ++ks) {
        const double af = fa[ks * 32];
#pragma unroll
        for (int q = 0; q < NTW; ++q) dmma884(acc[q][0], acc[q][1], af, wb[(ks * NT + q) * 32]);
      }
      const double *xs = st + FG + nt0 * 64 + 2 * lane;
      const double lse = st[FG + XG + g];
      const double lse_o = has_old ? st[FG + XG + kGroupRows + g] : 0.0;

      // ---- pass 1: phi = exp(x - lse), a = F.W + bias - x (in place of acc), partial sa = sum phi*a
      double phi[2 * NTW];
#pragma unroll
      for (int q = 0; q < NTW; ++q) {
        const int c = 8 * (nt0 + q) + 2 * t;
        const bool ok0 = FULLK || c < Kc, ok1 = FULLK || c + 1 < Kc;
        const double2 xv = *reinterpret_cast<const double2 *>(xs + q * 64);
        const double2 bb = *reinterpret_cast<const double2 *>(bias_s + c);
        acc[q][0] = ok0 ? (acc[q][0] + bb.x) - xv.x : 0.0;
        acc[q][1] = ok1 ? (acc[q][1] + bb.y) - xv.y : 0.0;
        phi[2 * q] = xv.x - lse;
        phi[2 * q + 1] = xv.y - lse;
      }
      exp_batch_tab<2 * NTW>(phi, exp_tab);
      double sa = 0.0;
#pragma unroll
      for (int q = 0; q < NTW; ++q) {
        const int c = 8 * (nt0 + q) + 2 * t;
        if (!(FULLK || c < Kc)) phi[2 * q] = 0.0;
        if (!(FULLK || c + 1 < Kc)) phi[2 * q + 1] = 0.0;
        sa += phi[2 * q] * acc[q][0] + phi[2 * q + 1] * acc[q][1];
      }
      sa = quad_sum(sa);
      if (NSPLIT > 1) {
        // the unit's other warp holds the rest of the row: exchange through shared memory (fixed summation order)
        double *sx = sax_s + ((i & 1) * UNITS + unit) * (NSPLIT * kGroupRows);
        if (t == 0) sx[h * kGroupRows + g] = sa;
        asm volatile("bar.sync %0, %1;\n" ::"r"(1 + unit), "n"(NSPLIT * 32) : "memory");
        sa = sx[g] + sx[kGroupRows + g];
      } else {
        __syncwarp();
      }
      // every warp of the unit has finished reading the slot: re-arm it with the group R visits ahead
      if (h == 0 && lane == 0) {
        const long long nxt = (long long)grp + (long long)R * gstride;
        if (nxt < a.num_groups) feed(slot, (int)nxt);
      }

      // previous-point responsibilities (independent of the row sums: in wide mode they hide the cluster exchange)
      double po[2 * NTW];
      auto old_point = [&]() {
        if (has_old) {
#pragma unroll
          for (int q = 0; q < NTW; ++q) {
            po[2 * q] = xo[q].x - lse_o;
            po[2 * q + 1] = xo[q].y - lse_o;
          }
          exp_batch_tab<2 * NTW>(po, exp_tab);
        }
      };
      if (CL) {
        // the row sums of this CTA's 64 clusters -> every CTA of the cluster (st.async into the peers' mailboxes, counted on
        // the peers' mailbox barriers); total = sum over ranks in rank order (identical on every CTA).  Two parities: a peer can
        // post visit i+2 only after it has consumed MY visit i+1, which this unit posts after both its warps finished visit i.
        const int par = i & 1;
        double *mb = sa_mb + (size_t)((par * UNITS + unit) * kMaxCl) * kGroupRows;
        uint64_t *mbb = mb_bar + par * UNITS + unit;
        {
          // arm this CTA's mailbox barrier for the csize x 8 doubles of this visit, then every lane of the unit (all hold the
          // row sum) posts row g to ONE peer, rank h*4 + t: asynchronous remote stores that complete on the peer's barrier
          if (h == 0 && lane == 0) mbar_arrive_expect_tx(mbb, (uint32_t)(csize * kGroupRows * 8));
          const uint32_t la = smem_u32(mb + crank * kGroupRows + g), lb = smem_u32(mbb);
          for (int rc = h * 4 + t; rc < csize; rc += 4 * NSPLIT) st_async_f64(cluster_map(la, (uint32_t)rc), sa, cluster_map(lb, (uint32_t)rc));
        }
        old_point();
        mbar_wait(mbb, (uint32_t)((i >> 1) & 1));
        double tot = 0.0;
        for (int rc = 0; rc < csize; ++rc) tot += mb[rc * kGroupRows + g];
        sa = tot;
      }

      if (a.sa_vec != nullptr) {
        // sweep 1 of the chunked pass (K > 64): leave a = F.W + bias - x behind, accumulate the row sums over the chunks
        if (h == 0 && t == 0 && row_ok) a.sa_vec[row] = a.sa_first ? sa : a.sa_vec[row] + sa;
        double *a_g = p_ng_out + goff;
#pragma unroll
        for (int q = 0; q < NTW; ++q) stg2(a_g + q * 64, acc[q][0], acc[q][1]);
        continue;
      }

      // ---- pass 2: natural gradient, gradient, dots.  Outside the wide mode the previous-point responsibilities are evaluated here,
      // in batches of HB n-tiles, each consumed at once: the peak register demand of the loop is what decides whether the row loop
      // fits the 168 registers of the CTA shape (it must, also inside the fused loop kernel).
#ifndef GPC_G_OLD_BATCH
#define GPC_G_OLD_BATCH 2
#endif
      constexpr int HB = (!CL && NTW > GPC_G_OLD_BATCH) ? GPC_G_OLD_BATCH : NTW;
      double *ngo_g = p_ng_out + goff;
      double *gro_g = p_grad_out ? p_grad_out + goff : nullptr;
#pragma unroll
      for (int q0 = 0; q0 < NTW; q0 += HB) {
        double pb[2 * HB];
        if (has_old) {
          if (CL) {
#pragma unroll
            for (int q = 0; q < HB; ++q) {
              pb[2 * q] = po[2 * (q0 + q)];
              pb[2 * q + 1] = po[2 * (q0 + q) + 1];
            }
          } else {
#pragma unroll
            for (int q = 0; q < HB; ++q) {
              pb[2 * q] = xo[q0 + q].x - lse_o;
              pb[2 * q + 1] = xo[q0 + q].y - lse_o;
            }
            exp_batch_tab<2 * HB>(pb, exp_tab);
          }
        }
#pragma unroll
        for (int qq = 0; qq < HB; ++qq) {
          const int q = q0 + qq;
          const int c = 8 * (nt0 + q) + 2 * t;
          const bool ok0 = FULLK || c < Kc, ok1 = FULLK || c + 1 < Kc;
          const double n0 = (ok0 && row_ok) ? acc[q][0] - sa : 0.0;
          const double n1 = (ok1 && row_ok) ? acc[q][1] - sa : 0.0;
          const double g0 = n0 * phi[2 * q], g1 = n1 * phi[2 * q + 1];
          stg2(ngo_g + q * 64, n0, n1);
          if (gro_g) stg2(gro_g + q * 64, g0, g1);
          d0 += n0 * g0 + n1 * g1;
          if (has_old) {
            const double po0 = (ok0 && row_ok) ? pb[2 * qq] : 0.0, po1 = (ok1 && row_ok) ? pb[2 * qq + 1] : 0.0;
            const double e0 = n0 - ngo[q].x, e1 = n1 - ngo[q].y;
            d1 += e0 * g0 + e1 * g1;
            d2 += e0 * (ngo[q].x * po0) + e1 * (ngo[q].y * po1);
          }
        }
      }
    }
    d0 = warp_sum(d0);
    d1 = warp_sum(d1);
    d2 = warp_sum(d2);
    if (lane == 0) {
      red_s[warp * 4 + 0] = d0;
      red_s[warp * 4 + 1] = d1;
      red_s[warp * 4 + 2] = d2;
    }
  }
}

// per-CTA partial of the three dots (fixed order over the warps); call after the block-wide barrier that follows g_rows
__device__ __forceinline__ void g_cta_partial(const double *red_s, double *partials) {
  if (threadIdx.x < 3) {
    double v = 0.0;
    for (int w = 0; w < kGConsumers; ++w) v += red_s[w * 4 + threadIdx.x];
    partials[(size_t)blockIdx.x * 4 + threadIdx.x] = v;
  }
}

// dot `q` summed over the per-CTA partials by ONE WARP: all loads in flight at once (up to 256 CTAs), fixed order
__device__ __forceinline__ double reduce_dot_warp(const double *gpart, int ncl, int q) {
  const int lane = threadIdx.x & 31;
  double w[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) w[u] = __ldcg(gpart + (size_t)min(lane + 32 * u, ncl - 1) * 4 + q);
#pragma unroll
  for (int u = 0; u < 8; ++u) w[u] = (lane + 32 * u < ncl) ? w[u] : 0.0;
  return warp_sum(((w[0] + w[1]) + (w[2] + w[3])) + ((w[4] + w[5]) + (w[6] + w[7])));  // identical on every CTA, run to run
}

// ------------------------------------------------------------------------------------------------
// G pass, K <= 64: warp-specialised row loop (the shape of the S pass)
// ------------------------------------------------------------------------------------------------
// 1 producer warp + 4 PRODUCT warps (tensor pipe) + 7 ELEMENT warps (exp / row sums / dots), around ONE ring of group-sized stages
//   [ F group, later a = F.W + bias - x | x | x_old | ng_old | lse (8) | lse_old (8) ]
// filled by the producer with cp.async.bulk -- the previous-point operands included, so that nobody waits on an HBM load with
// registers as the landing zone.  Every stage is visited by all four product warps (warp p: n-tiles p, p + 4; its B fragments of W
// live in REGISTERS for the whole pass, the accumulators start at bias - x) -- they take the A fragments of F into registers, meet on
// a named barrier and overwrite the head of the stage with a -- and then by ONE element warp (stage j -> warp j mod 7: the whole row
// of 8*KP/32 elements per lane, so the row sum sum_k phi*a needs the quad shuffle only).  exp(x - lse) is evaluated by the element
// warp while the product warps still work on the stage.
// Why: in the unit design (g_rows) every warp alternates between a DMMA chain and the elementwise work and waits ~3 000 cycles per
// visit for the previous-point loads; three such warps per scheduler keep the shared FP64 / DMMA pipe 55-62 % busy.  Specialised
// roles overlap the two kinds of pipe work and the memory latency the way the S pass does (0.94-0.96 of HBM).
__host__ __device__ inline int g2_head_doubles(int KP, int DP) { return kGroupRows * (DP > KP ? DP : KP); }
__host__ __device__ inline int g2_stage_doubles(int KP, int DP) { return g2_head_doubles(KP, DP) + 3 * kGroupRows * KP + 2 * kGroupRows; }
constexpr int kG2ProductBar = 7;  // named barrier of the four product warps

struct G2Smem {
  double *stages, *red_s;
  uint64_t *full_bar, *a_bar, *empty_bar;
};
constexpr int kG2FixedDoubles = kGConsumers * 4;  // dot partials
__device__ __forceinline__ G2Smem g2_carve(unsigned char *smem_raw, int S, int KP, int DP) {
  G2Smem m;
  m.stages = reinterpret_cast<double *>(smem_raw);
  m.red_s = m.stages + (size_t)S * g2_stage_doubles(KP, DP);  // kGConsumers * 4
  m.full_bar = reinterpret_cast<uint64_t *>(m.red_s + kGConsumers * 4);
  m.a_bar = m.full_bar + S;
  m.empty_bar = m.a_bar + S;
  return m;
}
// barriers + zeroed dot partials; the caller synchronises the block afterwards
__device__ __forceinline__ void g2_setup(const G2Smem &m, int S) {
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&m.full_bar[s], 1);
      mbar_init(&m.a_bar[s], kSStatWarps);
      mbar_init(&m.empty_bar[s], 1);
    }
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < kGConsumers * 4; i += blockDim.x) m.red_s[i] = 0.0;
}

// ---------------- producer (one lane): bulk copies of the stage operands
template <int KP>
__device__ __forceinline__ void g2_producer(const NatgradArgs &a, const G2Smem &m, const GPtrs &p, int cid, int ncl) {
  constexpr int XG = kGroupRows * KP;
  const int DP = a.DP, FG = kGroupRows * DP, S = a.stages, HD = g2_head_doubles(KP, DP), SD = g2_stage_doubles(KP, DP);
  if ((threadIdx.x & 31) != 0) return;
  const bool has_old = (p.ng_old != nullptr);
  const uint32_t bytes = (uint32_t)(FG + XG + kGroupRows + (has_old ? 2 * XG + kGroupRows : 0)) * 8u;
  const int Mc = cid < a.num_groups ? (a.num_groups - cid + ncl - 1) / ncl : 0;  // row groups of this CTA: cid + j * ncl
  for (int j = 0; j < Mc; ++j) {
    const size_t grp = (size_t)cid + (size_t)j * ncl;
    const int s = j % S;
    if (j >= S) mbar_wait(&m.empty_bar[s], (uint32_t)(((j / S) & 1) ^ 1));
    double *st = m.stages + (size_t)s * SD;
    mbar_arrive_expect_tx(&m.full_bar[s], bytes);
    bulk_g2s(st, a.F + grp * FG, FG * 8u, &m.full_bar[s]);
    bulk_g2s(st + HD, p.x + grp * XG, XG * 8u, &m.full_bar[s]);
    bulk_g2s(st + HD + 3 * XG, p.lse + grp * kGroupRows, kGroupRows * 8u, &m.full_bar[s]);
    if (has_old) {
      bulk_g2s(st + HD + XG, p.x_old + grp * XG, XG * 8u, &m.full_bar[s]);
      bulk_g2s(st + HD + 2 * XG, p.ng_old + grp * XG, XG * 8u, &m.full_bar[s]);
      bulk_g2s(st + HD + 3 * XG + kGroupRows, p.lse_old + grp * kGroupRows, kGroupRows * 8u, &m.full_bar[s]);
    }
  }
}

// ---------------- product warps: a = F.W + bias - x for the n-tiles pw, pw + 4 of every stage, left in the head of the stage.
// A DMMA m8n8k4 occupies the scheduler's FP64 pipe for 16 cycles and has 26 cycles of dependent-issue latency
// (tools/microbench/dmma_latency.cu: 26.0 cycles per DMMA with one accumulator chain, 16.2 with two): two chains keep one warp at the
// peak rate, provided it issues DMMAs and next to nothing else.  Hence: all A fragments of the stage into registers first, then the
// 2 x 16 DMMAs back to back as kG2Split chains per n-tile (k-step ks -> chain ks mod kG2Split; chain 0 starts at -x, chain 1 at the
// bias, so that a = F.W + bias - x costs one addition per element).
#ifndef GPC_G2_SPLIT
#define GPC_G2_SPLIT 2
#endif
constexpr int kG2Split = GPC_G2_SPLIT;
template <int KP, bool FULLK, bool FULLD>
__device__ __forceinline__ void g2_product_loop(const NatgradArgs &a, const G2Smem &m, const GPtrs &p, int cid, int ncl) {
  constexpr int NT = KP / 8;
  constexpr int NTP = (NT + kSStatWarps - 1) / kSStatWarps;  // n-tiles per product warp
  constexpr int KS = GPC_MAX_DP / 4;                         // k-steps at the widest feature block
  const int DP = a.DP, S = a.stages, HD = g2_head_doubles(KP, DP), SD = g2_stage_doubles(KP, DP), ksteps = FULLD ? KS : DP / 4;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int pw = warp - kSSoftmaxWarps;
  const int Kc = FULLK ? KP : a.K;
  // B fragments of this warp's n-tiles (lane (g, t): W[4 ks + t][8 nt + g]) and its bias entries in registers for the whole pass
  double wr[KS][NTP], b0[NTP], b1[NTP];
#pragma unroll
  for (int q = 0; q < NTP; ++q) {
    const int nt = pw + kSStatWarps * q;
    const bool has = nt < NT;
    const double *wrow = p.Wt + (size_t)(8 * (has ? nt : 0) + g) * DP + t;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) wr[ks][q] = (has && (FULLD || ks < ksteps)) ? __ldcg(wrow + 4 * ks) : 0.0;  // L2: rewritten by other CTAs inside the fused loop kernel
    const int c = 8 * (has ? nt : 0) + 2 * t;
    b0[q] = (has && (FULLK || c < Kc)) ? __ldcg(p.bias + c) : 0.0;
    b1[q] = (has && (FULLK || c + 1 < Kc)) ? __ldcg(p.bias + c + 1) : 0.0;
  }
  const int Mc = cid < a.num_groups ? (a.num_groups - cid + ncl - 1) / ncl : 0;
  for (int j = 0; j < Mc; ++j) {
    const int s = j % S;
    double *st = m.stages + (size_t)s * SD;
    mbar_wait(&m.full_bar[s], (uint32_t)((j / S) & 1));
    double af[KS];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) af[ks] = (FULLD || ks < ksteps) ? st[ks * 32 + lane] : 0.0;
    // chain 0 starts at -x, chain 1 at the bias, the others at 0
    double acc[NTP][kG2Split][2];
#pragma unroll
    for (int q = 0; q < NTP; ++q) {
      const int nt = pw + kSStatWarps * q;
      double2 xv = make_double2(0.0, 0.0);
      if (nt < NT) xv = *reinterpret_cast<const double2 *>(st + HD + nt * 64 + 2 * lane);
      acc[q][0][0] = -xv.x;
      acc[q][0][1] = -xv.y;
      acc[q][1][0] = b0[q];
      acc[q][1][1] = b1[q];
#pragma unroll
      for (int c = 2; c < kG2Split; ++c) acc[q][c][0] = acc[q][c][1] = 0.0;
    }
    if (FULLD) {
#pragma unroll
      for (int ks = 0; ks < KS; ++ks)
#pragma unroll
        for (int q = 0; q < NTP; ++q) dmma884(acc[q][ks % kG2Split][0], acc[q][ks % kG2Split][1], af[ks], wr[ks][q]);
    } else {
#pragma unroll
      for (int ks = 0; ks < KS; ++ks)
        if (ks < ksteps) {
#pragma unroll
          for (int q = 0; q < NTP; ++q) dmma884(acc[q][ks % kG2Split][0], acc[q][ks % kG2Split][1], af[ks], wr[ks][q]);
        }
    }
    // every product warp holds its A fragments of the stage in registers: the head may be overwritten
    asm volatile("bar.sync %0, %1;\n" ::"n"(kG2ProductBar), "n"(kSStatWarps * 32) : "memory");
#pragma unroll
    for (int q = 0; q < NTP; ++q) {
      const int nt = pw + kSStatWarps * q;
      if (nt < NT) {
        const int c = 8 * nt + 2 * t;
        double a0 = acc[q][0][0] + acc[q][1][0], a1 = acc[q][0][1] + acc[q][1][1];
#pragma unroll
        for (int cc = 2; cc < kG2Split; ++cc) {
          a0 += acc[q][cc][0];
          a1 += acc[q][cc][1];
        }
        if (!(FULLK || c < Kc)) a0 = 0.0;
        if (!(FULLK || c + 1 < Kc)) a1 = 0.0;
        *reinterpret_cast<double2 *>(st + nt * 64 + 2 * lane) = make_double2(a0, a1);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&m.a_bar[s]);
  }
}
template <int KP, bool FULLK>
__device__ __forceinline__ void g2_product(const NatgradArgs &a, const G2Smem &m, const GPtrs &p, int cid, int ncl) {
  static_assert(kG2Split >= 2, "chain 0 carries -x, chain 1 the bias");
  if (a.DP == GPC_MAX_DP) g2_product_loop<KP, FULLK, true>(a, m, p, cid, ncl);  // no predicates around the DMMAs
  else g2_product_loop<KP, FULLK, false>(a, m, p, cid, ncl);
}

// ---------------- element warps: natural gradient, gradient and the three dots of one group each; red_s[warp*4 + {0,1,2}] <- dots
template <int KP, bool FULLK>
__device__ __forceinline__ void g2_element(const NatgradArgs &a, const G2Smem &m, const GPtrs &p, int cid, int ncl) {
  constexpr int NT = KP / 8;
  constexpr int XG = kGroupRows * KP;
  constexpr int EB = NT > 4 ? 4 : NT;  // n-tiles per exp batch at the current point
  constexpr int HB = NT > 4 ? 4 : NT;  // ... at the previous point (evaluated and consumed batch by batch: register peak)
  const int DP = a.DP, S = a.stages, HD = g2_head_doubles(KP, DP), SD = g2_stage_doubles(KP, DP);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const double exp_tab = kExpTab[lane];  // this lane's entry of the 2^(j/32) table (exp_batch_tab)
  const int Kc = FULLK ? KP : a.K;
  const bool has_old = (p.ng_old != nullptr);
  double *p_ng_out = p.ng_out, *p_grad_out = p.grad_out;
  double d0 = 0.0, d1 = 0.0, d2 = 0.0, d0b = 0.0, d1b = 0.0, d2b = 0.0;  // two chains per dot (even / odd column of the lane)
  const int Mc = cid < a.num_groups ? (a.num_groups - cid + ncl - 1) / ncl : 0;
  for (int j = warp; j < Mc; j += kSSoftmaxWarps) {
    const size_t grp = (size_t)cid + (size_t)j * ncl;
    const int s = j % S;
    const uint32_t par = (uint32_t)((j / S) & 1);
    const double *st = m.stages + (size_t)s * SD;
    const bool row_ok = (long long)(grp * kGroupRows + g) < (long long)a.N;
    mbar_wait(&m.full_bar[s], par);
    const double lse = st[HD + 3 * XG + g];
    const double lse_o = has_old ? st[HD + 3 * XG + kGroupRows + g] : 0.0;
    // ---- phi = exp(x - lse): needs the stage only, runs while the product warps still work on it
    double phi[2 * NT];
    const double *xs = st + HD + 2 * lane;
#pragma unroll
    for (int q0 = 0; q0 < NT; q0 += EB) {
      double eb[2 * EB];
#pragma unroll
      for (int q = 0; q < EB; ++q) {
        const double2 xv = *reinterpret_cast<const double2 *>(xs + (q0 + q) * 64);
        eb[2 * q] = xv.x - lse;
        eb[2 * q + 1] = xv.y - lse;
      }
      exp_batch_tab<2 * EB>(eb, exp_tab);
#pragma unroll
      for (int q = 0; q < EB; ++q) {
        const int c = 8 * (q0 + q) + 2 * t;
        phi[2 * (q0 + q)] = (FULLK || c < Kc) ? eb[2 * q] : 0.0;
        phi[2 * (q0 + q) + 1] = (FULLK || c + 1 < Kc) ? eb[2 * q + 1] : 0.0;
      }
    }
    // ---- a from the product warps; row sum sa = sum_k phi*a (the quad holds the whole row)
    mbar_wait(&m.a_bar[s], par);
    double sa = 0.0, sa2 = 0.0;
    const double *as = st + 2 * lane;
#pragma unroll
    for (int q = 0; q < NT; ++q) {
      const double2 v = *reinterpret_cast<const double2 *>(as + q * 64);
      sa = fma(phi[2 * q], v.x, sa);
      sa2 = fma(phi[2 * q + 1], v.y, sa2);
    }
    sa = quad_sum(sa + sa2);
    // ---- natural gradient, gradient, dots (MOHGP.py:150-153, collapsed_vb.py:154,160-164)
    const size_t goff = grp * XG + 2 * lane;
    double *ngo_g = p_ng_out + goff;
    double *gro_g = p_grad_out ? p_grad_out + goff : nullptr;
    const double *xos = st + HD + XG + 2 * lane, *ngs = st + HD + 2 * XG + 2 * lane;
#pragma unroll
    for (int q0 = 0; q0 < NT; q0 += HB) {
      double pb[2 * HB];
      if (has_old) {
#pragma unroll
        for (int q = 0; q < HB; ++q) {
          const double2 xo = *reinterpret_cast<const double2 *>(xos + (q0 + q) * 64);
          pb[2 * q] = xo.x - lse_o;
          pb[2 * q + 1] = xo.y - lse_o;
        }
        exp_batch_tab<2 * HB>(pb, exp_tab);
      }
#pragma unroll
      for (int qq = 0; qq < HB; ++qq) {
        const int q = q0 + qq;
        const int c = 8 * q + 2 * t;
        const bool ok0 = (FULLK || c < Kc) && row_ok, ok1 = (FULLK || c + 1 < Kc) && row_ok;
        // a is re-read from the stage rather than kept across the row sum: 32 registers that the exp batches of 8 need more
        const double2 av = *reinterpret_cast<const double2 *>(as + q * 64);
        const double n0 = ok0 ? av.x - sa : 0.0;
        const double n1 = ok1 ? av.y - sa : 0.0;
        const double g0 = n0 * phi[2 * q], g1 = n1 * phi[2 * q + 1];
        stg2(ngo_g + q * 64, n0, n1);
        if (gro_g) stg2(gro_g + q * 64, g0, g1);
        d0 = fma(n0, g0, d0);
        d0b = fma(n1, g1, d0b);
        if (has_old) {
          const double2 ngo = *reinterpret_cast<const double2 *>(ngs + q * 64);
          const double po0 = ok0 ? pb[2 * qq] : 0.0, po1 = ok1 ? pb[2 * qq + 1] : 0.0;
          const double e0 = n0 - ngo.x, e1 = n1 - ngo.y;
          d1 = fma(e0, g0, d1);
          d1b = fma(e1, g1, d1b);
          d2 = fma(e0, ngo.x * po0, d2);
          d2b = fma(e1, ngo.y * po1, d2b);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&m.empty_bar[s]);  // the stage goes back to the producer
  }
  d0 = warp_sum(d0 + d0b);
  d1 = warp_sum(d1 + d1b);
  d2 = warp_sum(d2 + d2b);
  if (lane == 0) {
    m.red_s[warp * 4 + 0] = d0;
    m.red_s[warp * 4 + 1] = d1;
    m.red_s[warp * 4 + 2] = d2;
  }
}
// the three roles; the caller has run g2_setup + a block barrier and synchronises the block afterwards
template <int KP, bool FULLK>
__device__ __forceinline__ void g2_rows(const NatgradArgs &a, const G2Smem &m, const GPtrs &p, int cid, int ncl) {
  static_assert(kGThreads == kSThreads, "the specialised G pass uses the warp roles of the S pass");
  const int warp = threadIdx.x >> 5;
  if (warp == kSSoftmaxWarps + kSStatWarps) g2_producer<KP>(a, m, p, cid, ncl);
  else if (warp < kSSoftmaxWarps) g2_element<KP, FULLK>(a, m, p, cid, ncl);
  else g2_product<KP, FULLK>(a, m, p, cid, ncl);
}

// end of a stand-alone G pass: deterministic two-level reduction of the three dots (per-CTA partial, the last CTA sums all partials
// in index order) and, with rows sharded over GPUs, the push of this rank's dots into every peer's inbox
__device__ __forceinline__ void g_finish(const NatgradArgs &a, const double *red_s) {
  __shared__ bool is_last;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // deterministic two-level reduction: per-CTA partial, last CTA sums all partials in index order
  g_cta_partial(red_s, a.partials);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int done = atomicAdd(a.counter, 1u);
    is_last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    if (warp < 3) {
      // fixed lane->CTA assignment and fixed shuffle tree => run-to-run deterministic (the same routine as the fused loop kernel)
      double v = 0.0;
      if ((int)gridDim.x <= 256) v = reduce_dot_warp(a.partials, (int)gridDim.x, warp);
      else {
        for (int c = lane; c < (int)gridDim.x; c += 32) v += __ldcg(a.partials + (size_t)c * 4 + warp);
        v = warp_sum(v);
      }
      if (lane == 0) {
        a.dots_out[warp] = v;
        if (a.peers != nullptr) {  // this rank's contribution -> slot of the NEXT dots exchange in every inbox
          const gpc_peers_t *P = a.peers;
          const unsigned long long e = P->epochs[1] + 1;
          for (int r = 0; r < P->world; ++r) P->inbox[r][xchg_off_dots(P->world, P->stats_len, (int)(e & 1), P->rank) + warp] = v;
        }
      }
    }
    __syncthreads();
    if (warp == 0) {
      if (a.peers != nullptr) {
        const gpc_peers_t *P = a.peers;
        const unsigned long long e = P->epochs[1] + 1;
        xchg_raise_flags(P, 1, (int)(e & 1), e);
        __syncwarp();
        if (lane == 0) P->epochs[1] = e;
      }
      if (lane == 0) *a.counter = 0u;
    }
  }
}

template <int KP, bool FULLK, bool CL = false>
__global__ void __launch_bounds__(kGThreads, 1) k_natgrad(const NatgradArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int crank = CL ? (int)cluster_ctarank() : 0, csize = CL ? a.csize : 1;
  const int cid = CL ? (int)cluster_id_x() : (int)blockIdx.x, ncl = CL ? (int)cluster_count_x() : (int)gridDim.x;
  if (a.loop != nullptr) {
    if (loop_skip(a.loop, GPC_PHASE_CONJ)) return;
    if (a.loop->slot_mode != 0 && a.loop->need_retry != 0) return;  // retry slot: same point, the gradient and its dots stand
  }
  const GSmem m = g_carve(smem_raw, a.stages, KP, a.DP);
  const GPtrs p = g_resolve(a, KP, crank, a.loop);
  g_setup<KP, CL>(a, m, p);
  if (CL) cluster_sync_all();  // every peer's mailbox barriers exist before the first remote arrive
  else __syncthreads();
  g_rows<KP, FULLK, CL>(a, m, p, cid, ncl, crank, csize);
  if (CL) cluster_sync_all();  // no CTA leaves while a peer may still store into its mailbox
  else __syncthreads();
  g_finish(a, m.red_s);
}

// K <= 64, one chunk: the warp-specialised row loop
template <int KP, bool FULLK>
__global__ void __launch_bounds__(kGThreads, 1) k_natgrad2(const NatgradArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  if (a.loop != nullptr) {
    if (loop_skip(a.loop, GPC_PHASE_CONJ)) return;
    if (a.loop->slot_mode != 0 && a.loop->need_retry != 0) return;  // retry slot: same point, the gradient and its dots stand
  }
  const G2Smem m = g2_carve(smem_raw, a.stages, KP, a.DP);
  const GPtrs p = g_resolve(a, KP, 0, a.loop);
  g2_setup(m, a.stages);
  __syncthreads();
  g2_rows<KP, FULLK>(a, m, p, (int)blockIdx.x, (int)gridDim.x);
  __syncthreads();
  g_finish(a, m.red_s);
}

// ------------------------------------------------------------------------------------------------
// S pass
// ------------------------------------------------------------------------------------------------
// beta modes: GPC_BETA_* of include/gpclust_b200.h

struct StepStatsArgs {
  const double *F;       // groups x (8*DP), A-fragment order
  const double *x_base;  // logits at the accepted point, C-fragment order
  const double *ng;      // natural gradient at x_base (nullable: "set" mode, x_new := x_base, nothing written)
  const double *sd_old;  // previous search direction (nullable or ignored when beta == 0)
  double *sd_new;        // nullable
  double *x_new;         // trial logits (may be nullptr in "set" mode)
  double *lse_new;       // Npad : logsumexp of the rows of x_new (of x_base in "set" mode); nullable
  const double *dots;    // [0]=sq [1]=num [2]=den of the G pass at x_base (device scalars, read when beta_mode != 0)
  double *beta_out;      // device scalar, written by CTA 0
  double *partials;      // gridDim.x * stat_len, stat_len = KP*DP + KP + 2  ([T KPxDP | phi_hat KP | H | pad])
  double step;
  double sq_old;         // natgrad.grad of the previous iteration (PR / FR only; the host read it back already)
  int beta_mode;
  int N, K, DP, num_groups, stages;
  const gpc_loop_t *loop;  // nullable: device-resident loop state (operands from `bufs`, beta mode / step / sq_old from it)
  gpc_bufs_t bufs;
  int phase;               // GPC_PHASE_CONJ / GPC_PHASE_RETRY
  const gpc_peers_t *peers;  // nullable: the dots are the rank-ordered sum of the W inbox contributions
  double *dots_sum;          // peers != NULL: CTA 0 stores the summed dots here (the loop tail and the host read them)
  // K > 64: statistics of one 64-cluster chunk per launch (x_base pre-offset to the chunk, ng == NULL).
  // lse_in != NULL: phi = exp(x - lse_in[row]) with the row logsumexp over ALL clusters (from k_step_lse) instead of a
  // softmax over the chunk.  part_stride / t_off / ph_off / h_off place the chunk's [T | phi_hat | H] inside the
  // per-CTA block of the full statistics vector (0 / 0 / 0 / 0 = single chunk).  h_in: per-CTA entropy partials of
  // k_step_lse, added into the H slot when write_h.
  long long xg_stride;
  const double *lse_in;
  const double *h_in;
  long long part_stride;
  int t_off, ph_off, h_off, write_h;
  // wide mode (template CL): clusters of `csize` CTAs, see NatgradArgs.  K is the total cluster count; partials holds one
  // full [T KPtot x DP | phi_hat KPtot | H | pad] block PER CLUSTER (rank r fills rows / entries [64 r, 64 r + 64), rank 0 H).
  int csize;
  // reverse != 0: this CTA walks its row groups from the last to the first.  The G pass walks them first to last, so the S pass that
  // follows starts on the operands (F, x, natural gradient) the G pass touched LAST -- still in L2 -- and ends on the groups the next
  // G pass starts with (its x_new is that pass's x).
  int reverse;
};

__host__ __device__ inline int s_stage_doubles(int KP, int DP) { return kGroupRows * DP + 3 * kGroupRows * KP; }
// wide mode: mailbox [softmax warps][2 parities][kMaxCl ranks][8 rows][max, sum, sum e*(x-max)] + 2 mbarriers per softmax warp
__host__ __device__ inline int s_cl_doubles() { return kSSoftmaxWarps * 2 * kMaxCl * kGroupRows * 3 + 2 * kSSoftmaxWarps; }

// shared-memory carve of the S pass (dynamic shared memory; S ring stages)
struct SSmem {
  double *stages, *red_s, *sm_mb;
  uint64_t *full_bar, *phi_bar, *empty_bar, *mb_bar;
};
__device__ __forceinline__ SSmem s_carve(unsigned char *smem_raw, int S, int KP, int DP) {
  SSmem m;
  m.stages = reinterpret_cast<double *>(smem_raw);
  m.red_s = m.stages + (size_t)S * s_stage_doubles(KP, DP);  // kSSoftmaxWarps
  m.full_bar = reinterpret_cast<uint64_t *>(m.red_s + kSSoftmaxWarps);
  m.phi_bar = m.full_bar + S;
  m.empty_bar = m.phi_bar + S;
  m.mb_bar = m.empty_bar + S;                                             // wide mode: [softmax warp][2]
  m.sm_mb = reinterpret_cast<double *>(m.mb_bar + 2 * kSSoftmaxWarps);    // wide mode: [softmax warp][2][kMaxCl][8][3]
  return m;
}
template <bool CL>
__device__ __forceinline__ void s_init_barriers(const SSmem &m, int S) {
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&m.full_bar[s], 1);
      mbar_init(&m.phi_bar[s], 1);
      mbar_init(&m.empty_bar[s], kSStatWarps);
    }
    if (CL)
      for (int s = 0; s < 2 * kSSoftmaxWarps; ++s) mbar_init(&m.mb_bar[s], 1);
    mbar_fence_init();
  }
}
// operands and scalars of one S pass, resolved from the launch arguments or from the device-resident loop state
struct SPtrs {
  const double *x_base, *ng, *sd_old;
  double *sd_new, *x_new, *lse_new, *beta_out;
  double sq_old, step;
  int beta_mode;
};
__device__ __forceinline__ SPtrs s_resolve(const StepStatsArgs &a, int KP, int crank, const gpc_loop_t *loop, int phase_arg) {
  const size_t coff = (size_t)crank * (kGroupRows * KP);
  SPtrs p;
  p.x_base = a.x_base + coff;
  p.ng = a.ng ? a.ng + coff : nullptr;
  p.sd_old = a.sd_old ? a.sd_old + coff : nullptr;
  p.sd_new = a.sd_new ? a.sd_new + coff : nullptr;
  p.x_new = a.x_new ? a.x_new + coff : nullptr;
  p.lse_new = a.lse_new;
  p.beta_out = a.beta_out;
  p.beta_mode = a.beta_mode;
  p.sq_old = a.sq_old;
  p.step = a.step;
  if (loop != nullptr) {
    const int phase = loop_phase(loop, phase_arg);
    const int xi = loop->xi, xt = loop->xtrial, ngi = loop->ngi;
    p.x_base = sel3(a.bufs.x[0], a.bufs.x[1], a.bufs.x[2], xi) + coff;
    p.x_new = sel3(a.bufs.x[0], a.bufs.x[1], a.bufs.x[2], xt) + coff;
    p.lse_new = sel3(a.bufs.lse[0], a.bufs.lse[1], a.bufs.lse[2], xt);
    p.ng = (ngi ? a.bufs.ng[1] : a.bufs.ng[0]) + coff;
    p.sd_old = p.sd_new = a.bufs.sd + coff;
    // collapsed_vb.py:157-158 (first pass / steepest) and :182-184 (retry with searchDir = -natgrad)
    p.beta_mode = (phase == GPC_PHASE_RETRY || loop->first != 0) ? GPC_BETA_ZERO : loop->method;
    if (phase == GPC_PHASE_RETRY) p.beta_out = nullptr;  // the trace reports the conjugate beta
    p.sq_old = loop->sq_old;
    p.step = loop->step;
  }
  return p;
}
// beta of collapsed_vb.py:157-166 from the three dots
__device__ __forceinline__ double cg_beta(int beta_mode, double sq, double num, double den, double sq_old) {
  double beta = 0.0;
  if (beta_mode != GPC_BETA_ZERO) {
    if (beta_mode == GPC_BETA_HS) beta = num / den;
    else if (beta_mode == GPC_BETA_PR) beta = num / sq_old;
    else beta = sq / sq_old;
    if (isnan(beta) || beta < 0.0) beta = 0.0;  // collapsed_vb.py:165-166
  }
  return beta;
}
// where this CTA's partial statistics go
struct SPart {
  double *t, *ph, *h;
};
template <int KP, bool CL>
__device__ __forceinline__ SPart s_part(const StepStatsArgs &a, int cid, int crank, int csize) {
  const int DP = a.DP;
  const bool chunked = (a.part_stride != 0);
  const int KPT = KP * csize;  // total padded cluster count
  double *part = CL ? a.partials + (size_t)cid * (size_t)(KPT * DP + KPT + 2)
                    : a.partials + (size_t)blockIdx.x * (chunked ? (size_t)a.part_stride : (size_t)(KP * DP + KP + 2));
  SPart q;
  q.t = part + (CL ? crank * KP * DP : (chunked ? a.t_off : 0));                  // T rows of this chunk
  q.ph = part + (CL ? KPT * DP + crank * KP : (chunked ? a.ph_off : KP * DP));    // phi_hat of this chunk
  q.h = part + (CL ? KPT * DP + KPT : (chunked ? a.h_off : KP * DP + KP));        // H (+ pad)
  return q;
}

// ---------------- producer (one lane): bulk copies of [F | x | ng | sd] groups into the ring
template <int KP>
__device__ __forceinline__ void s_producer(const StepStatsArgs &a, const SSmem &m, const SPtrs &p, int cid, int ncl, bool has_ng, bool load_old) {
  constexpr int XG = kGroupRows * KP;
  const int DP = a.DP, FG = kGroupRows * DP, S = a.stages;
  const int stage_doubles = s_stage_doubles(KP, DP);
  double *stages = m.stages;
  uint64_t *full_bar = m.full_bar, *empty_bar = m.empty_bar;
  const double *p_x_base = p.x_base, *p_ng = p.ng, *p_sd_old = p.sd_old;
  const size_t GS = a.xg_stride ? (size_t)a.xg_stride : (size_t)XG;
  const int lane = threadIdx.x & 31;
  {
    if (lane == 0) {
      const uint32_t bytes = (uint32_t)(FG + XG + (has_ng ? XG : 0) + (load_old ? XG : 0)) * 8u;
      const int Mc = cid < a.num_groups ? (a.num_groups - cid + ncl - 1) / ncl : 0;  // row groups of this CTA: cid + m * ncl
      for (int j = 0; j < Mc; ++j) {
        const int grp = cid + (a.reverse ? Mc - 1 - j : j) * ncl;
        const int s = j % S;
        if (j >= S) mbar_wait(&empty_bar[s], (uint32_t)(((j / S) & 1) ^ 1));
        double *st = stages + (size_t)s * stage_doubles;
        mbar_arrive_expect_tx(&full_bar[s], bytes);
        bulk_g2s(st, a.F + (size_t)grp * FG, FG * 8u, &full_bar[s]);
        bulk_g2s(st + FG, p_x_base + (size_t)grp * GS, XG * 8u, &full_bar[s]);
        if (has_ng) bulk_g2s(st + FG + XG, p_ng + (size_t)grp * GS, XG * 8u, &full_bar[s]);
        if (load_old) bulk_g2s(st + FG + 2 * XG, p_sd_old + (size_t)grp * GS, XG * 8u, &full_bar[s]);
      }
    }
  }
}

// ---------------- softmax warps: CG step + softmax statistics of one group each; red_s[warp] <- entropy partial
template <int KP, bool FULLK, bool CL>
__device__ __forceinline__ void s_softmax(const StepStatsArgs &a, const SSmem &m, const SPtrs &p, int cid, int ncl, int crank, int csize, double beta,
                                          bool has_ng, bool use_old) {
  constexpr int NT = KP / 8;
  constexpr int XG = kGroupRows * KP;
  const int DP = a.DP, FG = kGroupRows * DP, S = a.stages;
  const int stage_doubles = s_stage_doubles(KP, DP);
  double *stages = m.stages, *red_s = m.red_s, *sm_mb = m.sm_mb;
  uint64_t *full_bar = m.full_bar, *phi_bar = m.phi_bar, *mb_bar = m.mb_bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int Kc = FULLK ? KP : (CL ? min(KP, a.K - KP * crank) : a.K);
  double *p_sd_new = p.sd_new, *p_x_new = p.x_new, *p_lse_new = p.lse_new;
  const double step = p.step;
  const size_t GS = a.xg_stride ? (size_t)a.xg_stride : (size_t)XG;
  (void)csize;
  {
    // ---------------- softmax warps: CG step + softmax statistics of one group each
    double Hpart = 0.0;
    const int Mc = cid < a.num_groups ? (a.num_groups - cid + ncl - 1) / ncl : 0;
    int iw = 0;
    for (int j = warp; j < Mc; j += kSSoftmaxWarps, ++iw) {
      const int grp = cid + (a.reverse ? Mc - 1 - j : j) * ncl;
      const int s = j % S;
      double *st = stages + (size_t)s * stage_doubles;
      const int row = grp * kGroupRows + g;
      const bool row_ok = row < a.N;
      mbar_wait(&full_bar[s], (uint32_t)((j / S) & 1));
      double *xs = st + FG + 2 * lane;
      double x[NT][2];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const double2 v = *reinterpret_cast<const double2 *>(xs + nt * 64);
        x[nt][0] = v.x;
        x[nt][1] = v.y;
      }
      if (has_ng) {
        double *sdo_g = p_sd_new ? p_sd_new + (size_t)grp * GS + 2 * lane : nullptr;
        double *xn_g = p_x_new + (size_t)grp * GS + 2 * lane;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const double2 ngv = *reinterpret_cast<const double2 *>(xs + XG + nt * 64);
          double s0 = ngv.x, s1 = ngv.y;
          if (use_old) {
            const double2 sdv = *reinterpret_cast<const double2 *>(xs + 2 * XG + nt * 64);
            s0 += beta * sdv.x;  // searchDir = natgrad + beta*searchDir_old  (collapsed_vb.py:167, ascent sign)
            s1 += beta * sdv.y;
          }
          x[nt][0] += step * s0;  // collapsed_vb.py:172
          x[nt][1] += step * s1;
          if (sdo_g) stg2(sdo_g + nt * 64, s0, s1);
          stg2(xn_g + nt * 64, x[nt][0], x[nt][1]);
        }
      }
      if (a.lse_in != nullptr) {
        // chunk of a wider row: phi = exp(x - logsumexp over all clusters); entropy and lse come from k_step_lse
        const double l = a.lse_in[row];
        double ex[2 * NT];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          ex[2 * nt] = x[nt][0] - l;
          ex[2 * nt + 1] = x[nt][1] - l;
        }
        exp_batch<2 * NT>(ex);
        __syncwarp();
        double *psl = st + FG + (g >> 2) * 32 + (g & 3);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const int c = 8 * nt + 2 * t;
          psl[nt * 64 + (2 * t) * 4] = ((FULLK || c < Kc) && row_ok) ? ex[2 * nt] : 0.0;
          psl[nt * 64 + (2 * t + 1) * 4] = ((FULLK || c + 1 < Kc) && row_ok) ? ex[2 * nt + 1] : 0.0;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&phi_bar[s]);
        continue;
      }
      if (CL) {
        // softmax over a row spread across the CTAs of the cluster: local (max, sum e, sum e*(x - max)) of this CTA's 64 columns,
        // exchanged through every peer's mailbox; all ranks then combine the csize triples in the same order.
        double m = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const int c = 8 * nt + 2 * t;
          if (FULLK || c < Kc) m = fmax(m, x[nt][0]);
          if (FULLK || c + 1 < Kc) m = fmax(m, x[nt][1]);
        }
        m = quad_max(m);
        double ex[2 * NT];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          x[nt][0] -= m;
          x[nt][1] -= m;
          ex[2 * nt] = x[nt][0];
          ex[2 * nt + 1] = x[nt][1];
        }
        exp_batch<2 * NT>(ex);
        double ssum = 0.0, px = 0.0;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const int c = 8 * nt + 2 * t;
          const bool ok0 = FULLK || c < Kc, ok1 = FULLK || c + 1 < Kc;
          ex[2 * nt] = ok0 ? ex[2 * nt] : 0.0;
          ex[2 * nt + 1] = ok1 ? ex[2 * nt + 1] : 0.0;
          ssum += ex[2 * nt] + ex[2 * nt + 1];
          if (ok0) px += ex[2 * nt] * x[nt][0];
          if (ok1) px += ex[2 * nt + 1] * x[nt][1];
        }
        ssum = quad_sum(ssum);
        px = quad_sum(px);
        const int par = iw & 1;
        double *mb = sm_mb + (size_t)((warp * 2 + par) * kMaxCl) * (kGroupRows * 3);
        uint64_t *mbb = mb_bar + warp * 2 + par;
        {
          // arm this CTA's mailbox barrier for the csize x 8 triples of this visit; all four lanes of a quad hold the row's
          // triple: lane t posts it to the ranks t and t + 4 (asynchronous remote stores counted on the peer's barrier)
          if (lane == 0) mbar_arrive_expect_tx(mbb, (uint32_t)(csize * kGroupRows * 24));
          const uint32_t la = smem_u32(mb + (crank * kGroupRows + g) * 3), lb = smem_u32(mbb);
          for (int rc = t; rc < csize; rc += 4) {
            const uint32_t ra = cluster_map(la, (uint32_t)rc), rb = cluster_map(lb, (uint32_t)rc);
            st_async_f64(ra, m, rb);
            st_async_f64(ra + 8, ssum, rb);
            st_async_f64(ra + 16, px, rb);
          }
        }
        mbar_wait(mbb, (uint32_t)((iw >> 1) & 1));
        // lane t of the quad combines the ranks t and t + 4
        double mm[2], ss[2], pp[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int rc = t + 4 * q;
          const bool v = rc < csize;
          const double *src = mb + ((v ? rc : 0) * kGroupRows + g) * 3;
          mm[q] = v ? src[0] : -INFINITY;
          ss[q] = v ? src[1] : 0.0;
          pp[q] = v ? src[2] : 0.0;
        }
        const double M = quad_max(fmax(mm[0], mm[1]));
        double ee[3] = {mm[0] - M, mm[1] - M, m - M};
        const double dm0 = (t < csize) ? ee[0] : 0.0, dm1 = (t + 4 < csize) ? ee[1] : 0.0;
        exp_batch<3>(ee);
        const double Ssum = quad_sum(ss[0] * ee[0] + ss[1] * ee[1]);
        const double Psum = quad_sum(ee[0] * (pp[0] + dm0 * ss[0]) + ee[1] * (pp[1] + dm1 * ss[1]));
        const double inv = row_ok ? ee[2] / Ssum : 0.0;
        const double lognorm = log(Ssum);
        if (crank == 0 && row_ok && t == 0) Hpart += lognorm - Psum / Ssum;
        if (crank == 0 && p_lse_new && t == 0) p_lse_new[row] = M + lognorm;
        __syncwarp();
        double *ps = st + FG + (g >> 2) * 32 + (g & 3);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          ps[nt * 64 + (2 * t) * 4] = ex[2 * nt] * inv;
          ps[nt * 64 + (2 * t + 1) * 4] = ex[2 * nt + 1] * inv;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&phi_bar[s]);
        continue;
      }
      // row softmax: max, exp, sum   (np_utilities.py:25-40)
      double m = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int c = 8 * nt + 2 * t;
        if (FULLK || c < Kc) m = fmax(m, x[nt][0]);
        if (FULLK || c + 1 < Kc) m = fmax(m, x[nt][1]);
      }
      m = quad_max(m);
      double ssum = 0.0, px = 0.0;
      double ex[2 * NT];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        x[nt][0] -= m;
        x[nt][1] -= m;
        ex[2 * nt] = x[nt][0];
        ex[2 * nt + 1] = x[nt][1];
      }
      exp_batch<2 * NT>(ex);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int c = 8 * nt + 2 * t;
        const bool ok0 = FULLK || c < Kc, ok1 = FULLK || c + 1 < Kc;
        const double e0 = ok0 ? ex[2 * nt] : 0.0, e1 = ok1 ? ex[2 * nt + 1] : 0.0;
        ssum += e0 + e1;
        if (ok0) px += e0 * x[nt][0];
        if (ok1) px += e1 * x[nt][1];
        x[nt][0] = e0;
        x[nt][1] = e1;
      }
      ssum = quad_sum(ssum);
      px = quad_sum(px);
      const double inv = row_ok ? 1.0 / ssum : 0.0;
      const double lognorm = log(ssum);
      // row entropy -sum phi*log(phi) with log(phi) = (x - max) - log(sum)  ==  log(sum) - sum(e*(x - max))/sum
      if (row_ok && t == 0) Hpart += lognorm - px * inv;
      if (p_lse_new && t == 0) p_lse_new[row] = m + lognorm;

      // phi -> the stage's x region, in the A-fragment order of phi^T: element (row r, cluster k) at
      // ((k/8)*2 + r/4)*32 + (k%8)*4 + r%4.  Every lane has read its x values above.
      __syncwarp();
      double *ps = st + FG + (g >> 2) * 32 + (g & 3);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        ps[nt * 64 + (2 * t) * 4] = x[nt][0] * inv;
        ps[nt * 64 + (2 * t + 1) * 4] = x[nt][1] * inv;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&phi_bar[s]);
    }
    Hpart = warp_sum(Hpart);
    if (lane == 0) red_s[warp] = Hpart;
  }
}

// ---------------- statistics warps: T[k][d] += sum_rows phi[row][k] F[row][d] for every group of this CTA
template <int KP, bool CL>
__device__ __forceinline__ void s_stats(const StepStatsArgs &a, const SSmem &m, const SPart &part, int cid, int ncl) {
  constexpr int NT = KP / 8;
  constexpr int TPW = NT * 8 / kSStatWarps;  // 8x8 tiles of T per statistics warp q: tile i -> (mt = i/2, dt = 4*(i%2) + q)
  const int DP = a.DP, FG = kGroupRows * DP, S = a.stages, DT = DP / 8;
  const int stage_doubles = s_stage_doubles(KP, DP);
  double *stages = m.stages;
  uint64_t *phi_bar = m.phi_bar, *empty_bar = m.empty_bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  double *part_t = part.t, *part_ph = part.ph;
  {
    // ---------------- statistics warps: T[k][d] += sum_rows phi[row][k] F[row][d] for every group of this CTA
    const int q = warp - kSSoftmaxWarps;
    double acc[TPW][2];
#pragma unroll
    for (int i = 0; i < TPW; ++i) acc[i][0] = acc[i][1] = 0.0;
    double colsum[NT];  // sum over rows == t (mod 4) of phi[row][8*mt + g]
#pragma unroll
    for (int mt = 0; mt < NT; ++mt) colsum[mt] = 0.0;
    const int Mc = cid < a.num_groups ? (a.num_groups - cid + ncl - 1) / ncl : 0;
    for (int j = 0; j < Mc; ++j) {
      const int s = j % S;
      const double *st = stages + (size_t)s * stage_doubles;
      mbar_wait(&phi_bar[s], (uint32_t)((j / S) & 1));
      const double *pa = st + FG + lane;
      // B fragment of F for (ks', dt): lane (g, t) reads F[row 4ks'+t][8dt+g] = st[(2dt + g/4)*32 + (4ks'+t)*4 + g%4]
      const double *fb = st + (g >> 2) * 32 + t * 4 + (g & 3);
      // all fragments of the group first (one shared-memory round trip per group), then the 2*TPW DMMAs back to back
      double af[2][NT], bf[2][2];
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
        for (int mt = 0; mt < NT; ++mt) af[ks][mt] = pa[(mt * 2 + ks) * 32];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int dt = (e << 2) + q;
          bf[ks][e] = (dt < DT) ? fb[dt * 64 + ks * 16] : 0.0;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);  // the stage is in registers now
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
        for (int mt = 0; mt < NT; ++mt) colsum[mt] += af[ks][mt];
#pragma unroll
        for (int i = 0; i < TPW; ++i) dmma884(acc[i][0], acc[i][1], af[ks][i >> 1], bf[ks][i & 1]);  // tile (mt = i/2, dt = 4*(i%2) + q)
      }
    }
    // ---- per-CTA partial statistics
#pragma unroll
    for (int i = 0; i < TPW; ++i) {
      const int mt = i >> 1, dt = ((i & 1) << 2) + q;
      if (dt < DT) stg2(part_t + (size_t)(8 * mt + g) * DP + 8 * dt + 2 * t, acc[i][0], acc[i][1]);
    }
    if (q == 0) {
#pragma unroll
      for (int mt = 0; mt < NT; ++mt) {
        const double cs = quad_sum(colsum[mt]);
        if (t == 0) part_ph[8 * mt + g] = cs;
      }
    }
  }
}
// entropy partial of this CTA (+ the chunked-mode carry-in); call after the block-wide barrier that follows the role functions
template <bool CL>
__device__ __forceinline__ void s_cta_entropy(const StepStatsArgs &a, const SSmem &m, const SPart &part, int crank) {
  const bool chunked = (a.part_stride != 0);
  if (threadIdx.x == 0 && (CL ? crank == 0 : (!chunked || a.write_h))) {
    double v = 0.0;
    if (a.lse_in == nullptr)
      for (int w = 0; w < kSSoftmaxWarps; ++w) v += m.red_s[w];
    if (a.h_in != nullptr) v += a.h_in[blockIdx.x];
    part.h[0] = v;
    part.h[1] = 0.0;
  }
}

template <int KP, bool FULLK, bool CL = false>
__global__ void __launch_bounds__(kSThreads, 1) k_step_stats(const StepStatsArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5;
  const int crank = CL ? (int)cluster_ctarank() : 0, csize = CL ? a.csize : 1;
  const int cid = CL ? (int)cluster_id_x() : (int)blockIdx.x, ncl = CL ? (int)cluster_count_x() : (int)gridDim.x;
  const SSmem m = s_carve(smem_raw, a.stages, KP, a.DP);
  s_init_barriers<CL>(m, a.stages);
  if (CL) cluster_sync_all();
  else __syncthreads();

  if (a.loop != nullptr && loop_skip(a.loop, a.phase)) return;
  const SPtrs p = s_resolve(a, KP, crank, a.loop, a.phase);
  __shared__ double dots_sh[4];
  // rows sharded over GPUs: the dots of the G pass arrive through the peer inbox.  Only the consumers need them (beta); the
  // producer warp goes straight to filling the ring, so the exchange latency / rank skew hides behind the first HBM loads
  const bool is_producer = (warp == kSSoftmaxWarps + kSStatWarps);
  if (a.peers != nullptr && !is_producer) {
    if (threadIdx.x == 0) {
      const gpc_peers_t *P = a.peers;
      const unsigned long long e = P->epochs[1];  // the G pass of this rank has pushed exchange e
      const bool ok = xchg_wait(P, 1, (int)(e & 1), e);
      double d[3] = {0.0, 0.0, 0.0};
      for (int r = 0; r < P->world; ++r) {
        const double *src = P->inbox[P->rank] + xchg_off_dots(P->world, P->stats_len, (int)(e & 1), r);
        for (int q = 0; q < 3; ++q) d[q] += __ldcg(src + q);
      }
      for (int q = 0; q < 3; ++q) dots_sh[q] = ok ? d[q] : nan("");
      // a peer that never shows up is a hard fault of the exchange, not a numerical outcome: sticky bit, the host raises
      if (!ok && a.loop != nullptr) atomicOr(const_cast<int *>(&a.loop->done), GPC_DONE_PEER_TIMEOUT);
      if (blockIdx.x == 0 && a.dots_sum)
        for (int q = 0; q < 3; ++q) a.dots_sum[q] = dots_sh[q];
    }
    asm volatile("bar.sync 1, %0;\n" ::"n"((kSSoftmaxWarps + kSStatWarps) * 32) : "memory");
  }
  double beta = 0.0;
  if (p.beta_mode != GPC_BETA_ZERO && !(a.peers != nullptr && is_producer))
    beta = a.peers ? cg_beta(p.beta_mode, dots_sh[0], dots_sh[1], dots_sh[2], p.sq_old) : cg_beta(p.beta_mode, a.dots[0], a.dots[1], a.dots[2], p.sq_old);
  if (blockIdx.x == 0 && threadIdx.x == 0 && p.beta_out) *p.beta_out = beta;
  const bool has_ng = (p.ng != nullptr);
  const bool use_old = has_ng && (beta != 0.0) && (p.sd_old != nullptr);
  // what the producer stages: with the peer exchange it does not know beta, so the old search direction is staged whenever the
  // step may use it (beta == 0 then simply leaves it unread)
  const bool load_old = (a.peers != nullptr) ? (has_ng && p.beta_mode != GPC_BETA_ZERO && p.sd_old != nullptr) : use_old;
  const SPart part = s_part<KP, CL>(a, cid, crank, csize);

  if (is_producer) s_producer<KP>(a, m, p, cid, ncl, has_ng, load_old);
  else if (warp < kSSoftmaxWarps) s_softmax<KP, FULLK, CL>(a, m, p, cid, ncl, crank, csize, beta, has_ng, use_old);
  else s_stats<KP, CL>(a, m, part, cid, ncl);
  if (CL) cluster_sync_all();  // no CTA leaves while a peer may still store into its mailbox
  else __syncthreads();
  s_cta_entropy<CL>(a, m, part, crank);
}

// out[j] = sum over CTAs of partials[c][j], fixed order (deterministic).  Block = 32 columns x 8 part-groups:
// group q sums parts q, q+8, ... (two independent chains), the 8 group sums are combined in a fixed tree.
__global__ void __launch_bounds__(256) k_reduce_partials(const double *__restrict__ partials, int num_parts, int len, double *__restrict__ out) {
  __shared__ double red[8][33];
  const int c = threadIdx.x & 31, q = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + c;
  double v0 = 0.0, v1 = 0.0;
  if (j < len) {
    int p = q;
    for (; p + 8 < num_parts; p += 16) {
      v0 += partials[(size_t)p * len + j];
      v1 += partials[(size_t)(p + 8) * len + j];
    }
    if (p < num_parts) v0 += partials[(size_t)p * len + j];
  }
  red[q][c] = v0 + v1;
  __syncthreads();
  if (q == 0 && j < len)
    out[j] = ((red[0][c] + red[1][c]) + (red[2][c] + red[3][c])) + ((red[4][c] + red[5][c]) + (red[6][c] + red[7][c]));
}

// ------------------------------------------------------------------------------------------------
// K > 64 (clusters processed in chunks of 64): the two row-coupled steps that span all chunks.
// ------------------------------------------------------------------------------------------------
// S1: CG step on whole rows + row logsumexp + entropy.  One warp per 8-row group, lane (g, t) walks the n-tiles of
// row g; two sweeps (max, then sum of exponentials) exactly as np_utilities.py:25-40, the second one re-reading the
// x_new this lane has just written (L1/L2 hits).
struct StepLseArgs {
  const double *x_base, *ng, *sd_old;
  double *sd_new, *x_new, *lse_new;
  const double *dots;
  double *beta_out;
  double *hpart;  // gridDim.x entropy partials
  double step, sq_old;
  int beta_mode, N, K, KP, num_groups;
};
__global__ void __launch_bounds__(256) k_step_lse(const StepLseArgs a) {
  __shared__ double red_s[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int NTT = a.KP / 8;
  const size_t GS = (size_t)kGroupRows * a.KP;
  double beta = 0.0;
  if (a.beta_mode != GPC_BETA_ZERO) {
    const double sq = a.dots[0], num = a.dots[1], den = a.dots[2];
    if (a.beta_mode == GPC_BETA_HS) beta = num / den;
    else if (a.beta_mode == GPC_BETA_PR) beta = num / a.sq_old;
    else beta = sq / a.sq_old;
    if (isnan(beta) || beta < 0.0) beta = 0.0;  // collapsed_vb.py:165-166
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && a.beta_out) *a.beta_out = beta;
  const bool has_ng = (a.ng != nullptr);
  const bool use_old = has_ng && beta != 0.0 && a.sd_old != nullptr;
  double Hpart = 0.0;
  for (int grp = blockIdx.x * 8 + warp; grp < a.num_groups; grp += gridDim.x * 8) {
    const int row = grp * kGroupRows + g;
    const bool row_ok = row < a.N;
    const size_t off = (size_t)grp * GS + 2 * lane;
    const double *xsrc = has_ng ? a.x_new : a.x_base;
    double m = -INFINITY;
    for (int nt = 0; nt < NTT; ++nt) {
      const int c = 8 * nt + 2 * t;
      double2 xv = ldg2(a.x_base + off + nt * 64);
      if (has_ng) {
        const double2 nv = ldg2_stream(a.ng + off + nt * 64);
        double s0 = nv.x, s1 = nv.y;
        if (use_old) {
          const double2 sv = ldg2(a.sd_old + off + nt * 64);
          s0 += beta * sv.x;  // collapsed_vb.py:167
          s1 += beta * sv.y;
        }
        xv.x += a.step * s0;  // collapsed_vb.py:172
        xv.y += a.step * s1;
        if (a.sd_new) stg2(a.sd_new + off + nt * 64, s0, s1);
        stg2(a.x_new + off + nt * 64, xv.x, xv.y);
      }
      if (c < a.K) m = fmax(m, xv.x);
      if (c + 1 < a.K) m = fmax(m, xv.y);
    }
    m = quad_max(m);
    double ssum = 0.0, px = 0.0;
    for (int nt0 = 0; nt0 < NTT; nt0 += 4) {
      double z[8];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const double2 xv = (nt0 + q < NTT) ? ldg2(xsrc + off + (nt0 + q) * 64) : make_double2(m, m);
        z[2 * q] = xv.x - m;
        z[2 * q + 1] = xv.y - m;
      }
      double e[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) e[q] = z[q];
      exp_batch<8>(e);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int c = 8 * (nt0 + (q >> 1)) + 2 * t + (q & 1);
        if (c < a.K) {
          ssum += e[q];
          px += e[q] * z[q];
        }
      }
    }
    ssum = quad_sum(ssum);
    px = quad_sum(px);
    const double lognorm = log(ssum);
    if (row_ok && t == 0) Hpart += lognorm - px / ssum;
    if (a.lse_new && t == 0) a.lse_new[row] = m + lognorm;
  }
  Hpart = warp_sum(Hpart);
  if (lane == 0) red_s[warp] = Hpart;
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += red_s[w];
    a.hpart[blockIdx.x] = v;
  }
}

// G2: a (left in the natural-gradient buffer by the per-chunk sweep) -> natural gradient, gradient and the CG dots over
// whole rows: natgrad = a - sa[row], grad = natgrad * exp(x - lse)   (MOHGP.py:150-151, collapsed_vb.py:154,160-164)
struct NatgradFinishArgs {
  double *ng;             // in: a ; out: natural gradient
  const double *x, *lse, *sa_vec;
  const double *x_old, *lse_old, *ng_old;  // nullable together
  double *grad_out;       // nullable
  double *partials;       // gridDim.x * 4
  unsigned int *counter;
  double *dots_out;
  int N, K, KP, num_groups;
};
__global__ void __launch_bounds__(256) k_natgrad_finish(const NatgradFinishArgs a) {
  __shared__ double red_s[8 * 4];
  __shared__ bool is_last;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int NTT = a.KP / 8;
  const size_t GS = (size_t)kGroupRows * a.KP;
  const bool has_old = (a.ng_old != nullptr);
  double d0 = 0.0, d1 = 0.0, d2 = 0.0;
  for (int grp = blockIdx.x * 8 + warp; grp < a.num_groups; grp += gridDim.x * 8) {
    const int row = grp * kGroupRows + g;
    const bool row_ok = row < a.N;
    const size_t off = (size_t)grp * GS + 2 * lane;
    const double sa = a.sa_vec[row], lse = a.lse[row], lse_o = has_old ? a.lse_old[row] : 0.0;
    for (int nt0 = 0; nt0 < NTT; nt0 += 4) {
      double phi[8], po[8];
      double2 av[4], ngo[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const bool in = nt0 + q < NTT;
        const double2 xv = in ? ldg2_stream(a.x + off + (nt0 + q) * 64) : make_double2(0.0, 0.0);
        av[q] = in ? ldg2(a.ng + off + (nt0 + q) * 64) : make_double2(0.0, 0.0);
        phi[2 * q] = xv.x - lse;
        phi[2 * q + 1] = xv.y - lse;
        if (has_old) {
          const double2 xo = in ? ldg2_stream(a.x_old + off + (nt0 + q) * 64) : make_double2(0.0, 0.0);
          ngo[q] = in ? ldg2_stream(a.ng_old + off + (nt0 + q) * 64) : make_double2(0.0, 0.0);
          po[2 * q] = xo.x - lse_o;
          po[2 * q + 1] = xo.y - lse_o;
        }
      }
      exp_batch<8>(phi);
      if (has_old) exp_batch<8>(po);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (nt0 + q >= NTT) continue;
        const int c = 8 * (nt0 + q) + 2 * t;
        const bool ok0 = c < a.K && row_ok, ok1 = c + 1 < a.K && row_ok;
        const double n0 = ok0 ? av[q].x - sa : 0.0, n1 = ok1 ? av[q].y - sa : 0.0;
        const double p0 = ok0 ? phi[2 * q] : 0.0, p1 = ok1 ? phi[2 * q + 1] : 0.0;
        const double g0 = n0 * p0, g1 = n1 * p1;
        stg2(a.ng + off + (nt0 + q) * 64, n0, n1);
        if (a.grad_out) stg2(a.grad_out + off + (nt0 + q) * 64, g0, g1);
        d0 += n0 * g0 + n1 * g1;
        if (has_old) {
          const double e0 = n0 - ngo[q].x, e1 = n1 - ngo[q].y;
          d1 += e0 * g0 + e1 * g1;
          d2 += e0 * (ngo[q].x * (ok0 ? po[2 * q] : 0.0)) + e1 * (ngo[q].y * (ok1 ? po[2 * q + 1] : 0.0));
        }
      }
    }
  }
  d0 = warp_sum(d0);
  d1 = warp_sum(d1);
  d2 = warp_sum(d2);
  if (lane == 0) {
    red_s[warp * 4 + 0] = d0;
    red_s[warp * 4 + 1] = d1;
    red_s[warp * 4 + 2] = d2;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += red_s[w * 4 + threadIdx.x];
    a.partials[(size_t)blockIdx.x * 4 + threadIdx.x] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(a.counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (is_last) {
    __threadfence();
    if (warp < 3) {
      double v = 0.0;
      for (int c = lane; c < (int)gridDim.x; c += 32) v += __ldcg(a.partials + (size_t)c * 4 + warp);
      v = warp_sum(v);
      if (lane == 0) a.dots_out[warp] = v;
    }
    if (threadIdx.x == 0) *a.counter = 0u;
  }
}

// The same reduction, but the result goes straight into every peer's inbox (this rank's slot of the next statistics
// exchange) over NVLink; the last CTA raises the flags.  Replaces reduce -> NCCL all-reduce between the S pass and the
// posterior kernel: the consumer sums the W inbox vectors in rank order.
__global__ void __launch_bounds__(256) k_reduce_push(const double *__restrict__ partials, int num_parts, int len, const gpc_peers_t *P,
                                                     unsigned int *counter, const gpc_loop_t *loop, int phase) {
  if (loop_skip(loop, phase)) return;
  __shared__ double red[8][33];
  __shared__ bool is_last;
  const int c = threadIdx.x & 31, q = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + c;
  const unsigned long long e = P->epochs[0] + 1;
  double v0 = 0.0, v1 = 0.0;
  if (j < len) {
    int p = q;
    for (; p + 8 < num_parts; p += 16) {
      v0 += partials[(size_t)p * len + j];
      v1 += partials[(size_t)(p + 8) * len + j];
    }
    if (p < num_parts) v0 += partials[(size_t)p * len + j];
  }
  red[q][c] = v0 + v1;
  __syncthreads();
  if (q < P->world && j < len) {  // warp q stores to peer q, q + 8, ...: coalesced 256-byte rows over NVLink
    const double v = ((red[0][c] + red[1][c]) + (red[2][c] + red[3][c])) + ((red[4][c] + red[5][c]) + (red[6][c] + red[7][c]));
    for (int r = q; r < P->world; r += 8) P->inbox[r][xchg_off_stats(P->world, len, (int)(e & 1), P->rank) + j] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (is_last && threadIdx.x < 32) {
    xchg_raise_flags(P, 0, (int)(e & 1), e);
    __syncwarp();
    if (threadIdx.x == 0) {
      P->epochs[0] = e;
      *counter = 0u;
    }
  }
}

}  // namespace gpc
