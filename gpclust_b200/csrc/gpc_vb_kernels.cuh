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
// Layout: every N x K array is row-major with leading dimension KP (K padded to 16/32/64); the
// feature matrix F (= Y for MOHGP, [x_i x_j, x_i] for MOG) is row-major N x DP, DP = D padded to a
// multiple of 8, pad columns zero.  All of them are allocated with roundup(N, 64) rows.
#pragma once
#include "gpc_common.cuh"

namespace gpc {

constexpr int kStages = 4;

struct VbSmemLayout {
  int f_stride;       // doubles per staged F row  (DP + 4)
  int stage_doubles;  // GPC_ROW_TILE * f_stride
  __host__ __device__ VbSmemLayout(int DP) : f_stride(DP + 4), stage_doubles(GPC_ROW_TILE * (DP + 4)) {}
};

// Producer warp: stream F row tiles [tile*64, tile*64+64) into the stage ring.
__device__ __forceinline__ void produce_f_tiles(const double *__restrict__ F, int DP, int f_stride, double *stages, int stage_doubles,
                                                uint64_t *full_bar, uint64_t *empty_bar, int first_tile, int tile_step, int num_tiles) {
  const int lane = threadIdx.x & 31;
  const uint32_t row_bytes = (uint32_t)DP * 8u;
  int it = 0;
  for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++it) {
    const int s = it % kStages;
    const uint32_t ph = (uint32_t)((it / kStages) & 1);
    mbar_wait(&empty_bar[s], ph ^ 1u);
    if (lane == 0) mbar_arrive_expect_tx(&full_bar[s], row_bytes * GPC_ROW_TILE);
    __syncwarp();
    double *dst = stages + (size_t)s * stage_doubles;
    const double *src = F + (size_t)tile * GPC_ROW_TILE * DP;
#pragma unroll
    for (int r = lane; r < GPC_ROW_TILE; r += 32) bulk_g2s(dst + r * f_stride, src + (size_t)r * DP, row_bytes, &full_bar[s]);
  }
}

// exp(x - m) row softmax pieces for the C-fragment layout: each lane holds 2*NT entries of one row.
template <int NT>
__device__ __forceinline__ void row_softmax(const double (&x)[NT][2], int t, int K, double (&p)[NT][2], double &rowmax, double &rowsum) {
  double m = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const int c = 8 * nt + 2 * t;
    if (c < K) m = fmax(m, x[nt][0]);
    if (c + 1 < K) m = fmax(m, x[nt][1]);
  }
  m = quad_max(m);
  double s = 0.0;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const int c = 8 * nt + 2 * t;
    p[nt][0] = (c < K) ? exp(x[nt][0] - m) : 0.0;
    p[nt][1] = (c + 1 < K) ? exp(x[nt][1] - m) : 0.0;
    s += p[nt][0] + p[nt][1];
  }
  s = quad_sum(s);
  const double inv = 1.0 / s;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    p[nt][0] *= inv;
    p[nt][1] *= inv;
  }
  rowmax = m;
  rowsum = s;
}

// ------------------------------------------------------------------------------------------------
// G pass
// ------------------------------------------------------------------------------------------------
struct NatgradArgs {
  const double *F;       // Npad x DP features
  const double *x;       // Npad x KP logits at the current point
  const double *Wt;      // KP x DP, Wt[k][d] = W[d][k]
  const double *bias;    // KP
  const double *x_old;   // logits at the previous accepted point (nullable -> no HS/PR dots)
  const double *ng_old;  // natural gradient at the previous accepted point (nullable)
  double *ng_out;        // Npad x KP natural gradient (ascent direction, un-negated)
  double *grad_out;      // nullable: Npad x KP gradient = natgrad * phi
  double *partials;      // gridDim.x * 4 doubles
  unsigned int *counter; // zero on entry; reset to zero on exit
  double *dots_out;      // [0]=natgrad.grad  [1]=(ng-ng_old).grad  [2]=(ng-ng_old).grad_old
  int N, K, DP, num_tiles;
};

template <int KP>
__global__ void __launch_bounds__(GPC_THREADS, 1) k_natgrad(const NatgradArgs a) {
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

  if (warp == GPC_CONSUMER_WARPS) {
    produce_f_tiles(a.F, a.DP, L.f_stride, stages, L.stage_doubles, full_bar, empty_bar, blockIdx.x, gridDim.x, a.num_tiles);
  } else {
    const bool has_old = (a.ng_old != nullptr);
    const int ksteps = a.DP / 4;
    double d0 = 0.0, d1 = 0.0, d2 = 0.0;
    int it = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++it) {
      const int s = it % kStages;
      const uint32_t ph = (uint32_t)((it / kStages) & 1);
      const int row = tile * GPC_ROW_TILE + warp * 8 + g;
      const bool row_ok = row < a.N;
      const size_t off = (size_t)row * KP + 2 * t;

      // epilogue operands: issued before the DMMA loop so their HBM latency overlaps it
      double x[NT][2], xo[NT][2], ngo[NT][2];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const double2 v = ldg2_stream(a.x + off + 8 * nt);
        x[nt][0] = v.x;
        x[nt][1] = v.y;
      }
      if (has_old) {
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const double2 v = ldg2_stream(a.x_old + off + 8 * nt);
          xo[nt][0] = v.x;
          xo[nt][1] = v.y;
          const double2 u = ldg2_stream(a.ng_old + off + 8 * nt);
          ngo[nt][0] = u.x;
          ngo[nt][1] = u.y;
        }
      }

      double acc[NT][2];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = 0.0;

      mbar_wait(&full_bar[s], ph);
      const double *fa = stages + (size_t)s * L.stage_doubles + (warp * 8 + g) * L.f_stride + t;
      const double *wb = wt_s + g * L.f_stride + t;
#pragma unroll 4
      for (int ks = 0; ks < ksteps; ++ks) {
        const double af = fa[4 * ks];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) dmma884(acc[nt][0], acc[nt][1], af, wb[nt * 8 * L.f_stride + 4 * ks]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);

      // ---- epilogue: softmax, natural gradient, dots
      double p[NT][2], m, ssum;
      row_softmax<NT>(x, t, a.K, p, m, ssum);
      double sa = 0.0;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int c = 8 * nt + 2 * t;
        acc[nt][0] = (c < a.K) ? (acc[nt][0] + bias_s[c]) - x[nt][0] : 0.0;
        acc[nt][1] = (c + 1 < a.K) ? (acc[nt][1] + bias_s[c + 1]) - x[nt][1] : 0.0;
        sa += p[nt][0] * acc[nt][0] + p[nt][1] * acc[nt][1];
      }
      sa = quad_sum(sa);
      double po[NT][2];
      if (has_old) {
        double mo, so;
        row_softmax<NT>(xo, t, a.K, po, mo, so);
      }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int c = 8 * nt + 2 * t;
        const double n0 = (c < a.K && row_ok) ? acc[nt][0] - sa : 0.0;
        const double n1 = (c + 1 < a.K && row_ok) ? acc[nt][1] - sa : 0.0;
        const double g0 = n0 * p[nt][0], g1 = n1 * p[nt][1];
        stg2(a.ng_out + off + 8 * nt, n0, n1);
        if (a.grad_out) stg2(a.grad_out + off + 8 * nt, g0, g1);
        d0 += n0 * g0 + n1 * g1;
        if (has_old && row_ok) {
          const double e0 = n0 - ngo[nt][0], e1 = n1 - ngo[nt][1];
          d1 += e0 * g0 + e1 * g1;
          d2 += e0 * (ngo[nt][0] * po[nt][0]) + e1 * (ngo[nt][1] * po[nt][1]);
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
  const double *F;       // Npad x DP
  const double *x_base;  // Npad x KP logits at the accepted point
  const double *ng;      // natural gradient at x_base (nullable: "set" mode, x_new := x_base, nothing written)
  const double *sd_old;  // previous search direction (nullable or ignored when beta == 0)
  double *sd_new;        // nullable
  double *x_new;         // trial logits (may be nullptr in "set" mode)
  const double *dots;    // [0]=sq [1]=num [2]=den of the G pass at x_base (device scalars, read when beta_mode != 0)
  double *beta_out;      // device scalar, written by CTA 0
  double *partials;      // gridDim.x * stat_len, stat_len = KP*DP + KP + 2  ([T KPxDP | phi_hat KP | H | pad]; even => 16-B aligned blocks)
  double step;
  double sq_old;         // natgrad.grad of the previous iteration (PR / FR only; the host read it back already)
  int beta_mode;
  int N, K, DP, num_tiles;
};

template <int KP>
__global__ void __launch_bounds__(GPC_THREADS, 1) k_step_stats(const StepStatsArgs a) {
  constexpr int NT = KP / 8;
  constexpr int NGRP = GPC_CONSUMER_WARPS / NT;  // consumer warps sharing one 8-cluster m-tile in the stats GEMM
  constexpr int MAXDT = GPC_MAX_DP / 8;
  constexpr int DT_PER = (MAXDT + NGRP - 1) / NGRP;
  constexpr int P_STRIDE = KP + 4;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const VbSmemLayout L(a.DP);
  double *stages = reinterpret_cast<double *>(smem_raw);
  double *p_s = stages + (size_t)kStages * L.stage_doubles;  // 2 x GPC_ROW_TILE x P_STRIDE
  double *red_s = p_s + 2 * GPC_ROW_TILE * P_STRIDE;         // GPC_CONSUMER_WARPS * (KP + 1)
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(red_s + GPC_CONSUMER_WARPS * (KP + 1));
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
  if (warp == GPC_CONSUMER_WARPS) {
    produce_f_tiles(a.F, a.DP, L.f_stride, stages, L.stage_doubles, full_bar, empty_bar, blockIdx.x, gridDim.x, a.num_tiles);
  } else {
    const int mt = warp % NT, dgrp = warp / NT;
    double acc[DT_PER][2];
#pragma unroll
    for (int i = 0; i < DT_PER; ++i) acc[i][0] = acc[i][1] = 0.0;
    double colsum[NT][2];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) colsum[nt][0] = colsum[nt][1] = 0.0;
    double Hpart = 0.0;

    int it = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++it) {
      const int s = it % kStages;
      const uint32_t ph = (uint32_t)((it / kStages) & 1);
      const int row = tile * GPC_ROW_TILE + warp * 8 + g;
      const bool row_ok = row < a.N;
      const size_t off = (size_t)row * KP + 2 * t;
      double *pt = p_s + (it & 1) * GPC_ROW_TILE * P_STRIDE;

      // ---- phase 1: CG step + softmax statistics for this warp's 8 rows
      double x[NT][2];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const double2 v = ldg2_stream(a.x_base + off + 8 * nt);
        x[nt][0] = v.x;
        x[nt][1] = v.y;
      }
      if (a.ng != nullptr) {
        double sd[NT][2];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const double2 v = ldg2_stream(a.ng + off + 8 * nt);
          sd[nt][0] = v.x;
          sd[nt][1] = v.y;
        }
        if (use_old) {
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            const double2 v = ldg2_stream(a.sd_old + off + 8 * nt);
            sd[nt][0] += beta * v.x;  // searchDir = natgrad + beta*searchDir_old  (collapsed_vb.py:167, ascent sign)
            sd[nt][1] += beta * v.y;
          }
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          x[nt][0] += a.step * sd[nt][0];  // collapsed_vb.py:172
          x[nt][1] += a.step * sd[nt][1];
          if (a.sd_new) stg2(a.sd_new + off + 8 * nt, sd[nt][0], sd[nt][1]);
          stg2(a.x_new + off + 8 * nt, x[nt][0], x[nt][1]);
        }
      }
      double p[NT][2], m, ssum;
      row_softmax<NT>(x, t, a.K, p, m, ssum);
      const double lognorm = log(ssum);
      double h = 0.0;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int c = 8 * nt + 2 * t;
        if (!row_ok) p[nt][0] = p[nt][1] = 0.0;
        // entropy term -phi*log(phi), log(phi) = (x - max) - log(sum)   (np_utilities.py:30-36)
        if (c < a.K) h -= p[nt][0] * ((x[nt][0] - m) - lognorm);
        if (c + 1 < a.K) h -= p[nt][1] * ((x[nt][1] - m) - lognorm);
        colsum[nt][0] += p[nt][0];
        colsum[nt][1] += p[nt][1];
        *reinterpret_cast<double2 *>(pt + (warp * 8 + g) * P_STRIDE + c) = make_double2(p[nt][0], p[nt][1]);
      }
      Hpart += h;  // h is zero for pad rows because p is zero there

      // all consumer warps have written their phi rows (and finished the GEMM of the previous tile)
      asm volatile("bar.sync 1, %0;\n" ::"n"(GPC_CONSUMER_WARPS * 32) : "memory");

      // ---- phase 2: ybark[k][d] += sum_rows phi[row][k] * F[row][d]   (this warp: m-tile mt, d-tiles dgrp + i*NGRP)
      mbar_wait(&full_bar[s], ph);
      const double *fb = stages + (size_t)s * L.stage_doubles + t * L.f_stride + g;
      const double *pa = pt + t * P_STRIDE + 8 * mt + g;
#pragma unroll 4
      for (int ks = 0; ks < GPC_ROW_TILE / 4; ++ks) {
        const double af = pa[4 * ks * P_STRIDE];
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
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const double c0 = colgroup_sum(colsum[nt][0]), c1 = colgroup_sum(colsum[nt][1]);
      if (g == 0) {
        red_s[warp * (KP + 1) + 8 * nt + 2 * t] = c0;
        red_s[warp * (KP + 1) + 8 * nt + 2 * t + 1] = c1;
      }
    }
    Hpart = warp_sum(Hpart);
    if (lane == 0) red_s[warp * (KP + 1) + KP] = Hpart;
    asm volatile("bar.sync 1, %0;\n" ::"n"(GPC_CONSUMER_WARPS * 32) : "memory");
    for (int j = threadIdx.x; j < KP + 1; j += GPC_CONSUMER_WARPS * 32) {
      double v = 0.0;
      for (int w = 0; w < GPC_CONSUMER_WARPS; ++w) v += red_s[w * (KP + 1) + j];
      part[(size_t)KP * a.DP + j] = v;
    }
    if (threadIdx.x == 0) part[(size_t)KP * a.DP + KP + 1] = 0.0;
  }
}

// out[j] = sum over CTAs of partials[c][j], fixed order (deterministic)
__global__ void k_reduce_partials(const double *__restrict__ partials, int num_parts, int len, double *__restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= len) return;
  double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
  int c = 0;
  for (; c + 3 < num_parts; c += 4) {
    v0 += partials[(size_t)c * len + j];
    v1 += partials[(size_t)(c + 1) * len + j];
    v2 += partials[(size_t)(c + 2) * len + j];
    v3 += partials[(size_t)(c + 3) * len + j];
  }
  for (; c < num_parts; ++c) v0 += partials[(size_t)c * len + j];
  out[j] = (v0 + v1) + (v2 + v3);
}

}  // namespace gpc
