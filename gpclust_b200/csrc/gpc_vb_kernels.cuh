// gpclust_b200 -- the two N-sized passes of one collapsed-VB iteration (FP64, sm_100a).
//
//   k_natgrad    : G pass.  a[n,k] = F[n,:] . W[:,k] + bias[k] - x[n,k]   (DMMA, F tile staged by bulk copy)
//                  natgrad = a - sum_k phi*a ; grad = natgrad*phi ; the three CG dot products.
//                  Restates MOHGP.py:137-153 (multiple_mahalanobis utilities.py:108-172 in GEMM form,
//                  row-constant terms dropped because the row-centering at MOHGP.py:150 removes them),
//                  MOG.py:72-84 and the dots of collapsed_vb.py:154,160-164.
//   k_step_stats : S pass.  searchDir = natgrad + beta*searchDir_old ; x_new = x + step*searchDir
//                  (collapsed_vb.py:167,172) fused with the softmax / entropy / phi_hat /
//                  phi^T F sufficient statistics of collapsed_mixture.py:53-56, MOHGP.py:76, MOG.py:56-57.
//
// Layout: every N x K array is row-major with leading dimension KP (K padded to 8/16/32/64); the
// feature matrix F (= Y for MOHGP, [x_i x_j, x_i] for MOG) is row-major with DP = D padded to a
// multiple of 8 columns (pad zero) and leading dimension DP + 4: the 4 pad doubles per row make the
// 64-bit DMMA fragment reads bank-conflict free once a 64-row tile has been copied VERBATIM into
// shared memory by a single cp.async.bulk.  All N-sized arrays are allocated with roundup(N, 64) rows.
//
// Both kernels are persistent (grid = 2 x SM count, two CTAs resident per SM so that one CTA's
// memory phase overlaps the other's FP64 phase) and 8 warps each own 8 rows of a 64-row tile; warp 0 also
// streams the F row tile of the NEXT iteration into a 2-stage shared-memory ring with cp.async.bulk +
// mbarrier (16 warps x 112 registers fill the four 16K-register files of an SM exactly).
// The S pass leaves lse[n] = logsumexp_k x[n,:] behind, so the G pass gets phi = exp(x - lse) without
// row reductions and can stream the "previous point" operands (x_old, ng_old) in small column chunks.
#pragma once
#include "gpc_common.cuh"

namespace gpc {

constexpr int kStages = 2;
constexpr int kCtasPerSm = 2;
#ifndef GPC_DMMA_UNROLL
#define GPC_DMMA_UNROLL 2
#endif
#ifndef GPC_PF_G
#define GPC_PF_G 2
#endif
#ifndef GPC_PF_S
#define GPC_PF_S 4
#endif
constexpr int kDmmaUnroll = GPC_DMMA_UNROLL;

struct VbSmemLayout {
  int f_stride;       // doubles per staged F row  (DP + 4)
  int stage_doubles;  // GPC_ROW_TILE * f_stride
  __host__ __device__ VbSmemLayout(int DP) : f_stride(DP + 4), stage_doubles(GPC_ROW_TILE * (DP + 4)) {}
};

// Producer duty (done by warp 0, one tile ahead of its use): copy F rows [tile*64, tile*64+64) of this
// CTA's j-th tile into stage j % kStages with one bulk copy (64 * (DP+4) * 8 bytes, contiguous in HBM).
__device__ __forceinline__ void produce_f_tile(const double *__restrict__ F, int stage_doubles, double *stages, uint64_t *full_bar,
                                               uint64_t *empty_bar, int j, int tile) {
  const int s = j % kStages;
  const uint32_t ph = (uint32_t)((j / kStages) & 1);
  mbar_wait(&empty_bar[s], ph ^ 1u);  // every warp has finished the previous tile staged here
  if ((threadIdx.x & 31) == 0) {
    const uint32_t bytes = (uint32_t)stage_doubles * 8u;
    mbar_arrive_expect_tx(&full_bar[s], bytes);
    bulk_g2s(stages + (size_t)s * stage_doubles, F + (size_t)tile * stage_doubles, bytes, &full_bar[s]);
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// G pass
// ------------------------------------------------------------------------------------------------
struct NatgradArgs {
  const double *F;        // Npad x (DP+4) features
  const double *x;        // Npad x KP logits at the current point
  const double *lse;      // Npad : logsumexp of the rows of x (written by the S pass that produced x)
  const double *Wt;       // KP x DP, Wt[k][d] = W[d][k]
  const double *bias;     // KP
  const double *x_old;    // logits at the previous accepted point (nullable -> no HS/PR dots)
  const double *lse_old;  // Npad : logsumexp of the rows of x_old
  const double *ng_old;   // natural gradient at the previous accepted point (nullable)
  double *ng_out;         // Npad x KP natural gradient (ascent direction, un-negated)
  double *grad_out;       // nullable: Npad x KP gradient = natgrad * phi
  double *partials;       // gridDim.x * 4 doubles
  unsigned int *counter;  // zero on entry; reset to zero on exit
  double *dots_out;       // [0]=natgrad.grad  [1]=(ng-ng_old).grad  [2]=(ng-ng_old).grad_old
  int N, K, DP, num_tiles;
};

// FULLK: K == KP, no column masking
template <int KP, bool FULLK>
__global__ void __launch_bounds__(GPC_THREADS, kCtasPerSm) k_natgrad(const NatgradArgs a) {
  constexpr int NT = KP / 8;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const VbSmemLayout L(a.DP);
  double *stages = reinterpret_cast<double *>(smem_raw);
  double *wt_s = stages + (size_t)kStages * L.stage_doubles;  // KP x f_stride
  double *bias_s = wt_s + KP * L.f_stride;                    // KP
  double *red_s = bias_s + KP;                                // GPC_CONSUMER_WARPS * 4
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(red_s + GPC_CONSUMER_WARPS * 4);
  uint64_t *empty_bar = full_bar + kStages;
  __shared__ bool is_last;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], GPC_CONSUMER_WARPS);
    }
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < KP * a.DP; i += blockDim.x) {
    const int k = i / a.DP, d = i - k * a.DP;
    wt_s[k * L.f_stride + d] = a.Wt[i];
  }
  for (int i = threadIdx.x; i < KP; i += blockDim.x) bias_s[i] = a.bias[i];
  __syncthreads();

  {
    const bool has_old = (a.ng_old != nullptr);
    const int ksteps = a.DP / 4;
    double d0 = 0.0, d1 = 0.0, d2 = 0.0;
    int it = 0;
    if (warp == 0 && (int)blockIdx.x < a.num_tiles)
      produce_f_tile(a.F, L.stage_doubles, stages, full_bar, empty_bar, 0, blockIdx.x);
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++it) {
      if (warp == 0 && tile + (int)gridDim.x < a.num_tiles) {
        produce_f_tile(a.F, L.stage_doubles, stages, full_bar, empty_bar, it + 1, tile + gridDim.x);
        // pull the next tile's N x K operands into L2 so that the demand loads below are L2 hits
        const size_t nxt = (size_t)(tile + gridDim.x) * GPC_ROW_TILE * KP;
        const uint32_t bytes = GPC_ROW_TILE * KP * 8u;
        if (lane == 0) bulk_prefetch_l2(a.x + nxt, bytes);
        if (has_old && lane == 1) bulk_prefetch_l2(a.x_old + nxt, bytes);
        if (has_old && lane == 2) bulk_prefetch_l2(a.ng_old + nxt, bytes);
      }
      const int s = it % kStages;
      const uint32_t ph = (uint32_t)((it / kStages) & 1);
      const int row = tile * GPC_ROW_TILE + warp * 8 + g;
      const bool row_ok = row < a.N;
      const size_t off = (size_t)row * KP + 2 * t;

      // current-point logits: issued before the DMMA loop so their HBM latency overlaps it
      double x[NT][2];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const double2 v = ldg2_stream(a.x + off + 8 * nt);
        x[nt][0] = v.x;
        x[nt][1] = v.y;
      }
      const double lse = a.lse[row];
      const double lse_o = has_old ? a.lse_old[row] : 0.0;

      double acc[NT][2];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = 0.0;

      mbar_wait(&full_bar[s], ph);
      const double *fa = stages + (size_t)s * L.stage_doubles + (warp * 8 + g) * L.f_stride + t;
      const double *wb = wt_s + g * L.f_stride + t;
#pragma unroll(kDmmaUnroll)
      for (int ks = 0; ks < ksteps; ++ks) {
        const double af = fa[4 * ks];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) dmma884(acc[nt][0], acc[nt][1], af, wb[nt * 8 * L.f_stride + 4 * ks]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);

      // previous-point operands are streamed in column chunks, PF chunks ahead of their use
      constexpr int PF = (NT < GPC_PF_G) ? NT : GPC_PF_G;
      double2 xo_q[PF], ngo_q[PF];
      if (has_old) {
#pragma unroll
        for (int q = 0; q < PF; ++q) {
          xo_q[q] = ldg2_stream(a.x_old + off + 8 * q);
          ngo_q[q] = ldg2_stream(a.ng_old + off + 8 * q);
        }
      }

      // ---- pass 1: phi = exp(x - lse) (in place of x), a = F.W + bias - x (in place of acc), sa = sum phi*a
      double sa = 0.0;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int c = 8 * nt + 2 * t;
        const bool ok0 = FULLK || c < a.K, ok1 = FULLK || c + 1 < a.K;
        const double x0 = x[nt][0], x1 = x[nt][1];
        const double2 bb = *reinterpret_cast<const double2 *>(bias_s + c);
        acc[nt][0] = ok0 ? (acc[nt][0] + bb.x) - x0 : 0.0;
        acc[nt][1] = ok1 ? (acc[nt][1] + bb.y) - x1 : 0.0;
        double p0, p1;
        exp_pair(x0 - lse, x1 - lse, p0, p1);
        x[nt][0] = ok0 ? p0 : 0.0;
        x[nt][1] = ok1 ? p1 : 0.0;
        sa += x[nt][0] * acc[nt][0] + x[nt][1] * acc[nt][1];
      }
      sa = quad_sum(sa);

      // ---- pass 2: natural gradient, gradient, dots
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int c = 8 * nt + 2 * t;
        const bool ok0 = FULLK || c < a.K, ok1 = FULLK || c + 1 < a.K;
        const double n0 = (ok0 && row_ok) ? acc[nt][0] - sa : 0.0;
        const double n1 = (ok1 && row_ok) ? acc[nt][1] - sa : 0.0;
        const double g0 = n0 * x[nt][0], g1 = n1 * x[nt][1];
        stg2(a.ng_out + off + 8 * nt, n0, n1);
        if (a.grad_out) stg2(a.grad_out + off + 8 * nt, g0, g1);
        d0 += n0 * g0 + n1 * g1;
        if (has_old) {
          const double2 xo = xo_q[nt % PF], ngo = ngo_q[nt % PF];
          if (nt + PF < NT) {
            xo_q[nt % PF] = ldg2_stream(a.x_old + off + 8 * (nt + PF));
            ngo_q[nt % PF] = ldg2_stream(a.ng_old + off + 8 * (nt + PF));
          }
          if (row_ok) {
            double po0, po1;
            exp_pair(xo.x - lse_o, xo.y - lse_o, po0, po1);
            if (!ok0) po0 = 0.0;
            if (!ok1) po1 = 0.0;
            const double e0 = n0 - ngo.x, e1 = n1 - ngo.y;
            d1 += e0 * g0 + e1 * g1;
            d2 += e0 * (ngo.x * po0) + e1 * (ngo.y * po1);
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
  __syncthreads();
  // deterministic two-level reduction: per-CTA partial, last CTA sums all partials in index order
  if (threadIdx.x < 3) {
    double v = 0.0;
    for (int w = 0; w < GPC_CONSUMER_WARPS; ++w) v += red_s[w * 4 + threadIdx.x];
    a.partials[(size_t)blockIdx.x * 4 + threadIdx.x] = v;
  }
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
      double v = 0.0;
      for (int c = lane; c < (int)gridDim.x; c += 32) v += a.partials[(size_t)c * 4 + warp];
      // fixed lane->CTA assignment and fixed shuffle tree => run-to-run deterministic
      v = warp_sum(v);
      if (lane == 0) a.dots_out[warp] = v;
    }
    if (threadIdx.x == 0) *a.counter = 0u;
  }
}

// ------------------------------------------------------------------------------------------------
// S pass
// ------------------------------------------------------------------------------------------------
// beta modes: GPC_BETA_* of include/gpclust_b200.h

struct StepStatsArgs {
  const double *F;       // Npad x (DP+4)
  const double *x_base;  // Npad x KP logits at the accepted point
  const double *ng;      // natural gradient at x_base (nullable: "set" mode, x_new := x_base, nothing written)
  const double *sd_old;  // previous search direction (nullable or ignored when beta == 0)
  double *sd_new;        // nullable
  double *x_new;         // trial logits (may be nullptr in "set" mode)
  double *lse_new;       // Npad : logsumexp of the rows of x_new (of x_base in "set" mode); nullable
  const double *dots;    // [0]=sq [1]=num [2]=den of the G pass at x_base (device scalars, read when beta_mode != 0)
  double *beta_out;      // device scalar, written by CTA 0
  double *partials;      // gridDim.x * stat_len, stat_len = KP*DP + KP + 2  ([T KPxDP | phi_hat KP | H | pad]; even => 16-B aligned blocks)
  double step;
  double sq_old;         // natgrad.grad of the previous iteration (PR / FR only; the host read it back already)
  int beta_mode;
  int N, K, DP, num_tiles;
};

template <int KP, bool FULLK>
__global__ void __launch_bounds__(GPC_THREADS, kCtasPerSm) k_step_stats(const StepStatsArgs a) {
  constexpr int NT = KP / 8;
  constexpr int NGRP = GPC_CONSUMER_WARPS / NT;  // consumer warps sharing one 8-cluster m-tile in the stats GEMM
  constexpr int MAXDT = GPC_MAX_DP / 8;
  constexpr int DT_PER = (MAXDT + NGRP - 1) / NGRP;
  constexpr int P_STRIDE = KP + 4;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const VbSmemLayout L(a.DP);
  double *stages = reinterpret_cast<double *>(smem_raw);
  double *p_s = stages + (size_t)kStages * L.stage_doubles;  // GPC_ROW_TILE x P_STRIDE (single buffer)
  double *red_s = p_s + GPC_ROW_TILE * P_STRIDE;             // GPC_CONSUMER_WARPS
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(red_s + GPC_CONSUMER_WARPS);
  uint64_t *empty_bar = full_bar + kStages;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], GPC_CONSUMER_WARPS);
    }
    mbar_fence_init();
  }
  __syncthreads();

  double beta = 0.0;
  if (a.beta_mode != GPC_BETA_ZERO) {
    const double sq = a.dots[0], num = a.dots[1], den = a.dots[2];
    const double sq_old = a.sq_old;
    if (a.beta_mode == GPC_BETA_HS) beta = num / den;
    else if (a.beta_mode == GPC_BETA_PR) beta = num / sq_old;
    else beta = sq / sq_old;
    if (isnan(beta) || beta < 0.0) beta = 0.0;  // collapsed_vb.py:165-166
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && a.beta_out) *a.beta_out = beta;
  const bool use_old = (beta != 0.0) && (a.sd_old != nullptr);

  const int DT = a.DP / 8;
  {
    const int mt = warp % NT, dgrp = warp / NT;
    double acc[DT_PER][2];
#pragma unroll
    for (int i = 0; i < DT_PER; ++i) acc[i][0] = acc[i][1] = 0.0;
    double colsum = 0.0;  // sum over rows == t (mod 4) of phi[row][8*mt + g]
    double Hpart = 0.0;

    int it = 0;
    if (warp == 0 && (int)blockIdx.x < a.num_tiles)
      produce_f_tile(a.F, L.stage_doubles, stages, full_bar, empty_bar, 0, blockIdx.x);
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++it) {
      if (warp == 0 && tile + (int)gridDim.x < a.num_tiles) {
        produce_f_tile(a.F, L.stage_doubles, stages, full_bar, empty_bar, it + 1, tile + gridDim.x);
        const size_t nxt = (size_t)(tile + gridDim.x) * GPC_ROW_TILE * KP;
        const uint32_t bytes = GPC_ROW_TILE * KP * 8u;
        if (lane == 0) bulk_prefetch_l2(a.x_base + nxt, bytes);
        if (a.ng != nullptr && lane == 1) bulk_prefetch_l2(a.ng + nxt, bytes);
        if (use_old && lane == 2) bulk_prefetch_l2(a.sd_old + nxt, bytes);
      }
      const int s = it % kStages;
      const uint32_t ph = (uint32_t)((it / kStages) & 1);
      const int row = tile * GPC_ROW_TILE + warp * 8 + g;
      const bool row_ok = row < a.N;
      const size_t off = (size_t)row * KP + 2 * t;

      // ---- phase 1: CG step + softmax statistics for this warp's 8 rows
      double x[NT][2];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const double2 v = ldg2_stream(a.x_base + off + 8 * nt);
        x[nt][0] = v.x;
        x[nt][1] = v.y;
      }
      if (a.ng != nullptr) {
        // operands streamed in column chunks, PF chunks ahead of their use
        constexpr int PF = (NT < GPC_PF_S) ? NT : GPC_PF_S;
        double2 ng_q[PF], sd_q[PF];
#pragma unroll
        for (int q = 0; q < PF; ++q) {
          ng_q[q] = ldg2_stream(a.ng + off + 8 * q);
          if (use_old) sd_q[q] = ldg2_stream(a.sd_old + off + 8 * q);
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          double s0 = ng_q[nt % PF].x, s1 = ng_q[nt % PF].y;
          if (use_old) {
            s0 += beta * sd_q[nt % PF].x;  // searchDir = natgrad + beta*searchDir_old  (collapsed_vb.py:167, ascent sign)
            s1 += beta * sd_q[nt % PF].y;
          }
          if (nt + PF < NT) {
            ng_q[nt % PF] = ldg2_stream(a.ng + off + 8 * (nt + PF));
            if (use_old) sd_q[nt % PF] = ldg2_stream(a.sd_old + off + 8 * (nt + PF));
          }
          x[nt][0] += a.step * s0;  // collapsed_vb.py:172
          x[nt][1] += a.step * s1;
          if (a.sd_new) stg2(a.sd_new + off + 8 * nt, s0, s1);
          stg2(a.x_new + off + 8 * nt, x[nt][0], x[nt][1]);
        }
      }
      // row softmax: max, exp, sum   (np_utilities.py:25-40)
      double m = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int c = 8 * nt + 2 * t;
        if (FULLK || c < a.K) m = fmax(m, x[nt][0]);
        if (FULLK || c + 1 < a.K) m = fmax(m, x[nt][1]);
      }
      m = quad_max(m);
      double ssum = 0.0, px = 0.0;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int c = 8 * nt + 2 * t;
        const bool ok0 = FULLK || c < a.K, ok1 = FULLK || c + 1 < a.K;
        const double z0 = x[nt][0] - m, z1 = x[nt][1] - m;
        double e0, e1;
        exp_pair(z0, z1, e0, e1);
        if (!ok0) e0 = 0.0;
        if (!ok1) e1 = 0.0;
        ssum += e0 + e1;
        if (ok0) px += e0 * z0;
        if (ok1) px += e1 * z1;
        x[nt][0] = e0;
        x[nt][1] = e1;
      }
      ssum = quad_sum(ssum);
      px = quad_sum(px);
      const double inv = row_ok ? 1.0 / ssum : 0.0;
      const double lognorm = log(ssum);
      // row entropy -sum phi*log(phi) with log(phi) = (x - max) - log(sum)  ==  log(sum) - sum(e*(x - max))/sum
      if (row_ok && t == 0) Hpart += lognorm - px * inv;
      if (a.lse_new && t == 0) a.lse_new[row] = m + lognorm;

      // the previous tile's GEMM has finished reading p_s
      asm volatile("bar.sync 1, %0;\n" ::"n"(GPC_CONSUMER_WARPS * 32) : "memory");
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int c = 8 * nt + 2 * t;
        *reinterpret_cast<double2 *>(p_s + (warp * 8 + g) * P_STRIDE + c) = make_double2(x[nt][0] * inv, x[nt][1] * inv);
      }
      // all consumer warps have written their phi rows
      asm volatile("bar.sync 2, %0;\n" ::"n"(GPC_CONSUMER_WARPS * 32) : "memory");

      // ---- phase 2: T[k][d] += sum_rows phi[row][k] * F[row][d]   (this warp: m-tile mt, d-tiles dgrp + i*NGRP)
      mbar_wait(&full_bar[s], ph);
      const double *fb = stages + (size_t)s * L.stage_doubles + t * L.f_stride + g;
      const double *pa = p_s + t * P_STRIDE + 8 * mt + g;
#pragma unroll(kDmmaUnroll)
      for (int ks = 0; ks < GPC_ROW_TILE / 4; ++ks) {
        const double af = pa[4 * ks * P_STRIDE];
        colsum += af;
#pragma unroll
        for (int i = 0; i < DT_PER; ++i) {
          const int dt = dgrp + i * NGRP;
          if (dt < DT) dmma884(acc[i][0], acc[i][1], af, fb[4 * ks * L.f_stride + 8 * dt]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);
    }

    // ---- per-CTA partial statistics
    double *part = a.partials + (size_t)blockIdx.x * (KP * a.DP + KP + 2);
#pragma unroll
    for (int i = 0; i < DT_PER; ++i) {
      const int dt = dgrp + i * NGRP;
      if (dt < DT) stg2(part + (size_t)(8 * mt + g) * a.DP + 8 * dt + 2 * t, acc[i][0], acc[i][1]);
    }
    colsum = quad_sum(colsum);
    Hpart = warp_sum(Hpart);
    if (dgrp == 0 && t == 0) part[(size_t)KP * a.DP + 8 * mt + g] = colsum;
    if (lane == 0) red_s[warp] = Hpart;
    asm volatile("bar.sync 1, %0;\n" ::"n"(GPC_CONSUMER_WARPS * 32) : "memory");
    if (threadIdx.x == 0) {
      double v = 0.0;
      for (int w = 0; w < GPC_CONSUMER_WARPS; ++w) v += red_s[w];
      part[(size_t)KP * a.DP + KP] = v;
      part[(size_t)KP * a.DP + KP + 1] = 0.0;
    }
  }
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

}  // namespace gpc
