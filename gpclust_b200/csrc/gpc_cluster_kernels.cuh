// gpclust_b200 -- per-cluster posterior kernels (small D x D FP64 linear algebra in shared memory)
// and the scalar "finalize" kernels that assemble the collapsed bound and the per-cluster gradient
// constants.  These sit on the serial critical path between the two N-sized passes.
//
//   k_mohgp_cluster : MOHGP.py:79-90   (jitchol, dtrtrs, Lambda_inv, dpotrs, muk) + the per-cluster
//                     pieces of MOHGP.py:129-135 and :146-147
//   k_mog_cluster   : MOG.py:52-61 (+ np_utilities.py:6-22 multiple_pdinv) and the per-cluster pieces of
//                     MOG.py:63-70, :79
//   k_*_finalize    : collapsed_mixture.py:66-94 (mixing-proportion bound and gradient), bound assembly
#pragma once
#include "gpc_common.cuh"

namespace gpc {

constexpr int kClusterThreads = 256;
// GPC_PRIOR_* and GPC_INFO_NOT_PD come from include/gpclust_b200.h

// In-place lower Cholesky of the n x n matrix M (row stride ld) in shared memory, all threads of the
// CTA.  Returns false when a pivot is not positive / not finite.  Strict upper triangle is left as is.
__device__ inline bool chol_lower_smem(double *M, int n, int ld, int *flag_s) {
  const int tid = threadIdx.x, nth = blockDim.x;
  if (tid == 0) *flag_s = 0;
  __syncthreads();
  for (int j = 0; j < n; ++j) {
    if (tid == 0) {
      const double d = M[j * ld + j];
      if (!(d > 0.0) || isinf(d)) *flag_s = 1;
      else M[j * ld + j] = sqrt(d);
    }
    __syncthreads();
    if (*flag_s) return false;
    const double ljj = M[j * ld + j];
    for (int i = j + 1 + tid; i < n; i += nth) M[i * ld + j] /= ljj;
    __syncthreads();
    const int m = n - j - 1;  // trailing size
    for (int e = tid; e < m * m; e += nth) {
      const int i = j + 1 + e / m, k = j + 1 + e % m;
      if (k <= i) M[i * ld + k] -= M[i * ld + j] * M[k * ld + j];
    }
    __syncthreads();
  }
  return true;
}

// x := (L L^T)^-1 x for one vector in shared memory, executed by warp 0 (other warps idle).
// Column-oriented substitution; L lower, row stride ld.
__device__ inline void potrs_vec_warp0(const double *L, int n, int ld, double *x) {
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    for (int i = 0; i < n; ++i) {  // forward: L y = x
      __syncwarp();
      const double xi = x[i] / L[i * ld + i];
      __syncwarp();
      if (lane == 0) x[i] = xi;
      for (int m = i + 1 + lane; m < n; m += 32) x[m] -= L[m * ld + i] * xi;
    }
    for (int i = n - 1; i >= 0; --i) {  // backward: L^T z = y
      __syncwarp();
      const double xi = x[i] / L[i * ld + i];
      __syncwarp();
      if (lane == 0) x[i] = xi;
      for (int m = lane; m < i; m += 32) x[m] -= L[i * ld + m] * xi;
    }
    __syncwarp();
  }
}

// Block-wide barrier of the per-cluster routines.  SUB = false: the whole CTA (__syncthreads).  SUB = true: the routine is run by
// the first kClusterThreads threads of a LARGER CTA (the fused loop kernel, 384 threads): named barrier 2 over those threads.
template <bool SUB>
__device__ __forceinline__ void psync() {
  if (SUB) asm volatile("bar.sync 2, %0;\n" ::"n"(kClusterThreads) : "memory");
  else __syncthreads();
}
template <bool SUB>
__device__ __forceinline__ int pthreads() {
  return SUB ? kClusterThreads : (int)blockDim.x;
}
template <bool SUB = false>
__device__ inline double block_sum(double v, double *red_s) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = pthreads<SUB>() >> 5;
  psync<SUB>();
  if (lane == 0) red_s[warp] = v;
  psync<SUB>();
  double r = 0.0;
  for (int w = 0; w < nw; ++w) r += red_s[w];
  return r;
}

// ------------------------------------------------------------------------------------------------
struct MohgpClusterArgs {
  const double *stats;   // [ybark KP x DP | phi_hat KP | H], already summed over CTAs (and ranks)
  const double *A;       // D x D : Ly^-1 Sf Ly^-T             (MOHGP.py:79)
  const double *Ly;      // D x D : chol(Sy + 1e-6 I), lower   (MOHGP.py:49)
  const double *Ly_inv;  // D x D : Ly^-1 (Sy_chol_inv of MOHGP.py:49) -- the dpotrs of MOHGP.py:88 as two triangular products
  const double *Sy;      // D x D : un-jittered                (MOHGP.py:85)
  const double *Sy_inv;  // D x D
  const double *Sf;      // D x D
  double *Wt;            // out KP x DP : Wt[k][:] = Sy^-1 mu_k (pad = 0)
  double *muk_t;         // out K x D   : row k = mu_k
  double *per_k;         // out KP x 4  : log_det_diff, s^T Lambda_inv s, tr(Lambda_inv Sy_inv), mu^T Sy^-1 mu
  double *Syi_ybark_t;   // out K x D   : row k = Sy^-1 ybark_k  (nullable)
  double *Lambda_inv;    // out K x D x D (nullable)
  double *C_chols;       // out K x D x D (nullable)
  int *info;
  int K, D, KP, DP;
};

constexpr int kClMaxQ = GPC_MAX_DP / 4;  // register slots per thread: entries m = 4*mm + q of one row / column

// Left-looking lower Cholesky of the D x D matrix in Cs (row stride ld), in place, by 4*D threads:
// quad r = tid/4 owns row r; its factor entries L[r][4mm+q] stay in registers, the pivot row is read from
// shared memory.  Two block barriers per column.  *flag_s must be 0 on entry (set on a bad pivot).
__device__ __forceinline__ bool chol_lower_quadrows(double *Cs, int D, int ld, int *flag_s) {
  const int r = threadIdx.x >> 2, q = threadIdx.x & 3;
  double Lr[kClMaxQ];
#pragma unroll
  for (int mm = 0; mm < kClMaxQ; ++mm) Lr[mm] = 0.0;
  bool ok = true;
  for (int j = 0; j < D && ok; ++j) {
    double pp[4] = {0.0, 0.0, 0.0, 0.0};  // four independent chains: FP64 FMA latency, not throughput, bounds this loop
#pragma unroll
    for (int mm = 0; mm < kClMaxQ; ++mm) {
      const int m = 4 * mm + q;
      pp[mm & 3] += (m < j) ? Lr[mm] * Cs[j * ld + m] : 0.0;
    }
    const double part = quad_sum((pp[0] + pp[1]) + (pp[2] + pp[3]));
    const double v = (r >= j && r < D) ? Cs[r * ld + j] - part : 0.0;
    if (r == j && q == 0) {
      if (!(v > 0.0) || isinf(v)) *flag_s = 1;
      else Cs[j * ld + j] = sqrt(v);
    }
    __syncthreads();
    ok = (*flag_s == 0);
    const double dj = Cs[j * ld + j];
    const double l = v / dj;
    const bool mine = ok && r > j && r < D && q == (j & 3);
    if (mine) Cs[r * ld + j] = l;
#pragma unroll
    for (int mm = 0; mm < kClMaxQ; ++mm) Lr[mm] = (mine && mm == (j >> 2)) ? l : Lr[mm];
    __syncthreads();
  }
  return ok;
}

// One CTA (256 threads) per cluster.  Thread (r = tid/4, q = tid%4): the quad r owns row r of the
// Cholesky factor and column r of the triangular solve, split over q by m = 4*mm + q and held in
// registers, so the only block-wide synchronisation is two barriers per Cholesky column.
__global__ void __launch_bounds__(kClusterThreads) k_mohgp_cluster(const MohgpClusterArgs a) {
  extern __shared__ __align__(16) double sm[];
  const int D = a.D, ld = GPC_MAX_DP + 1, k = blockIdx.x, tid = threadIdx.x;
  const int r = tid >> 2, q = tid & 3;
  double *Cs = sm;                  // D x ld : C_k -> its Cholesky factor -> Lambda_inv
  double *Ts = Cs + GPC_MAX_DP * ld;  // D x ld : T = L_k^-1 Ly^T
  double *Lis = Ts + GPC_MAX_DP * ld; // D x ld : Ly^-1
  double *v_s = Lis + GPC_MAX_DP * ld;  // D : s_k
  double *u_s = v_s + GPC_MAX_DP;       // D : scratch
  double *mu_s = u_s + GPC_MAX_DP;      // D
  double *red_s = mu_s + GPC_MAX_DP;    // 32
  __shared__ int flag_s;

  if (k >= a.K) {  // pad clusters: zero weights so that the G pass sees finite numbers
    for (int d = tid; d < a.DP; d += blockDim.x) a.Wt[(size_t)k * a.DP + d] = 0.0;
    if (tid < 4) a.per_k[k * 4 + tid] = 0.0;
    return;
  }
  const double phi_hat = a.stats[(size_t)a.KP * a.DP + k];
  for (int e = tid; e < D * D; e += blockDim.x) Lis[(e / D) * ld + (e % D)] = a.Ly_inv[e];
  for (int d = tid; d < D; d += blockDim.x) v_s[d] = a.stats[(size_t)k * a.DP + d];

  // ---- C_k = I + phi_hat*A and its Cholesky factor with GPy's jitchol retry
  // (plain attempt, then jitter = mean(diag)*1e-6 * 10^i, i = 0..4)
  bool ok = false;
  double diag_mean = 0.0;
  for (int attempt = 0; attempt < 6 && !ok; ++attempt) {
    double jitter = 0.0;
    if (attempt > 0) {
      jitter = diag_mean * 1e-6;
      for (int p = 1; p < attempt; ++p) jitter *= 10.0;
      if (!isfinite(jitter)) break;
    }
    __syncthreads();
    for (int e = tid; e < D * D; e += blockDim.x) {
      const int i = e / D, j = e % D;
      Cs[i * ld + j] = ((i == j) ? 1.0 + jitter : 0.0) + a.A[e] * phi_hat;  // MOHGP.py:80
    }
    if (tid == 0) flag_s = 0;
    __syncthreads();
    if (attempt == 0) {
      double dsum = 0.0, dmin = INFINITY;
      for (int i = 0; i < D; ++i) {  // every thread, same order: identical result everywhere
        dsum += Cs[i * ld + i];
        dmin = fmin(dmin, Cs[i * ld + i]);
      }
      diag_mean = dsum / D;
      if (!(dmin > 0.0)) break;  // "not pd: non-positive diagonal elements"
    }
    ok = chol_lower_quadrows(Cs, D, ld, &flag_s);
  }
  if (!ok) {
    if (tid == 0) atomicOr(a.info, GPC_INFO_NOT_PD);
    for (int d = tid; d < a.DP; d += blockDim.x) a.Wt[(size_t)k * a.DP + d] = 0.0;
    if (tid < 4) a.per_k[k * 4 + tid] = 0.0;
    return;
  }
  if (a.C_chols)
    for (int e = tid; e < D * D; e += blockDim.x) {
      const int i = e / D, j = e % D;
      a.C_chols[(size_t)k * D * D + e] = (j <= i) ? Cs[i * ld + j] : 0.0;
    }
  double ldd = 0.0;
  if (tid < 32) {
    for (int i = tid; i < D; i += 32) ldd += log(Cs[i * ld + i]);
    ldd = 2.0 * warp_sum(ldd);  // MOHGP.py:83
  }

  // ---- T = L_k^-1 Ly^T  (MOHGP.py:84): quad r solves right-hand-side column r, entries T[4mm+q][r] in registers
  for (int e = tid; e < D * D; e += blockDim.x) {
    const int i = e / D, j = e % D;
    Ts[i * ld + j] = (j >= i) ? a.Ly[(size_t)j * D + i] : 0.0;  // Ly^T
  }
  for (int d = tid; d < D; d += blockDim.x) u_s[d] = 1.0 / Cs[d * ld + d];
  __syncthreads();
  {
    double Tr[kClMaxQ];
#pragma unroll
    for (int mm = 0; mm < kClMaxQ; ++mm) Tr[mm] = 0.0;
    for (int i = 0; i < D; ++i) {
      double pp[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int mm = 0; mm < kClMaxQ; ++mm) {
        const int m = 4 * mm + q;
        pp[mm & 3] += (m < i) ? Cs[i * ld + m] * Tr[mm] : 0.0;
      }
      const double part = quad_sum((pp[0] + pp[1]) + (pp[2] + pp[3]));
      const double rhs = (r >= i && r < D) ? Ts[i * ld + r] : 0.0;  // (Ly^T)[i][r], staged below
      const double val = (rhs - part) * u_s[i];                     // u_s[i] = 1 / L[i][i]
      const bool mine = (q == (i & 3));
#pragma unroll
      for (int mm = 0; mm < kClMaxQ; ++mm) Tr[mm] = (mine && mm == (i >> 2)) ? val : Tr[mm];
    }
    __syncthreads();  // every quad has read its right-hand side out of Ts
    if (r < D) {
#pragma unroll
      for (int mm = 0; mm < kClMaxQ; ++mm) {
        const int m = 4 * mm + q;
        if (m < D) Ts[m * ld + r] = Tr[mm];
      }
    }
  }
  __syncthreads();

  // ---- Lambda_inv = (Sy - T^T T) / phi_hat   or Sf when phi_hat <= 1e-6   (MOHGP.py:85); 4 x 4 block per thread
  {
    const bool tiny = !(phi_hat > 1e-6);
    const int a0 = 4 * (tid >> 4), b0 = 4 * (tid & 15);
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    if (!tiny && a0 < D && b0 < D) {
      for (int m = 0; m < D; ++m) {
        double ta[4], tb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          ta[i] = Ts[m * ld + a0 + i];
          tb[i] = Ts[m * ld + b0 + i];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] += ta[i] * tb[j];
      }
    }
    // every thread has finished reading the factor (the solve above) -- Cs can take Lambda_inv
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ai = a0 + i, bj = b0 + j;
        if (ai < D && bj < D) {
          const double v = tiny ? a.Sf[ai * D + bj] : (a.Sy[ai * D + bj] - acc[i][j]) / phi_hat;
          Cs[ai * ld + bj] = v;
          if (a.Lambda_inv) a.Lambda_inv[(size_t)k * D * D + ai * D + bj] = v;
        }
      }
  }
  // ---- s_k = (Ly Ly^T)^-1 ybark_k = Ly^-T (Ly^-1 ybark_k)   (MOHGP.py:88)
  auto solve_sy = [&](double *vec) {  // vec <- Sy_jittered^-1 vec, result also left in u_s? no: in vec
    double part = 0.0;
    if (r < D) {
#pragma unroll
      for (int mm = 0; mm < kClMaxQ; ++mm) {
        const int m = 4 * mm + q;
        if (m <= r) part += Lis[r * ld + m] * vec[m];
      }
    }
    part = quad_sum(part);
    __syncthreads();
    if (r < D && q == 0) u_s[r] = part;
    __syncthreads();
    part = 0.0;
    if (r < D) {
#pragma unroll
      for (int mm = 0; mm < kClMaxQ; ++mm) {
        const int m = 4 * mm + q;
        if (m >= r && m < D) part += Lis[m * ld + r] * u_s[m];
      }
    }
    part = quad_sum(part);
    __syncthreads();
    if (r < D && q == 0) vec[r] = part;
    __syncthreads();
  };
  __syncthreads();
  solve_sy(v_s);
  if (a.Syi_ybark_t)
    for (int d = tid; d < D; d += blockDim.x) a.Syi_ybark_t[(size_t)k * D + d] = v_s[d];
  // mu_k[j] = sum_i Lambda_inv[i][j] s[i]   (MOHGP.py:90)
  {
    double part = 0.0;
    if (r < D) {
#pragma unroll
      for (int mm = 0; mm < kClMaxQ; ++mm) {
        const int i = 4 * mm + q;
        if (i < D) part += Cs[i * ld + r] * v_s[i];
      }
    }
    part = quad_sum(part);
    if (r < D && q == 0) {
      mu_s[r] = part;
      a.muk_t[(size_t)k * D + r] = part;
    }
  }
  // quad = sum_ij s_i s_j Lambda_inv_ij (MOHGP.py:133) ; trace = sum_ij Lambda_inv_ij Sy_inv_ij (MOHGP.py:147)
  double quad = 0.0, tr = 0.0;
  for (int e = tid; e < D * D; e += blockDim.x) {
    const int i = e / D, j = e % D;
    const double l = Cs[i * ld + j];
    quad += v_s[i] * v_s[j] * l;
    tr += l * a.Sy_inv[e];
  }
  quad = block_sum(quad, red_s);
  tr = block_sum(tr, red_s);
  // w_k = (Ly Ly^T)^-1 mu_k : the GEMM form of multiple_mahalanobis(Y, muk.T, Sy_chol) (MOHGP.py:144)
  __syncthreads();
  for (int d = tid; d < D; d += blockDim.x) v_s[d] = mu_s[d];
  __syncthreads();
  solve_sy(v_s);
  double m2 = 0.0;
  for (int d = tid; d < D; d += blockDim.x) m2 += mu_s[d] * v_s[d];
  m2 = block_sum(m2, red_s);
  for (int d = tid; d < a.DP; d += blockDim.x) a.Wt[(size_t)k * a.DP + d] = (d < D) ? v_s[d] : 0.0;
  if (tid == 0) {
    a.per_k[k * 4 + 0] = ldd;
    a.per_k[k * 4 + 1] = quad;
    a.per_k[k * 4 + 2] = tr;
    a.per_k[k * 4 + 3] = m2;
  }
}

// ------------------------------------------------------------------------------------------------
// Mixing-proportion terms (collapsed_mixture.py:66-94).  Block-cooperative: the special functions are
// evaluated one cluster per thread, the K-long sums / cumulative sums serially by thread 0 (same
// left-to-right order as numpy's cumsum).  Writes mixgrad[0..K) (shared) and *mix_out (shared).
constexpr int kMaxMixK = 512;
struct MixScratch {
  double a[kMaxMixK], b[kMaxMixK], c[kMaxMixK], da[kMaxMixK], db[kMaxMixK], dc[kMaxMixK];
};

template <bool SUB = false>
__device__ inline void mixing_terms(const double *phi_hat, int K, double alpha, int prior, MixScratch &w, double *mixgrad, double *mix_out) {
  const int tid = threadIdx.x;
  if (prior == GPC_PRIOR_SYMMETRIC) {
    for (int k = tid; k < K; k += pthreads<SUB>()) {
      const double v = alpha + phi_hat[k];
      w.a[k] = lgamma(v);
      mixgrad[k] = digamma_pos(v);  // collapsed_mixture.py:86
    }
    psync<SUB>();
    if (tid == 0) {
      // ln_dirichlet_C(alpha*1) - ln_dirichlet_C(alpha + phi_hat)   (np_utilities.py:67-69)
      double sum_a = 0.0, sum_lg = 0.0;
      for (int k = 0; k < K; ++k) {
        sum_a += alpha + phi_hat[k];
        sum_lg += w.a[k];
      }
      *mix_out = (lgamma(alpha * K) - K * lgamma(alpha)) - (lgamma(sum_a) - sum_lg);
    }
    psync<SUB>();
    return;
  }
  // truncated DP, collapsed_mixture.py:72-77 and :88-92
  if (tid == 0) {
    double tail = 0.0;  // phi_tilde_plus_hat[k] = sum_{j>=k} phi_hat[j]: the reference's reversed cumsum
    for (int k = K - 1; k >= 0; --k) {
      tail += phi_hat[k];
      w.c[k] = tail;
    }
  }
  psync<SUB>();
  for (int k = tid; k < K; k += pthreads<SUB>()) {
    const double ph = phi_hat[k], tph = w.c[k], til = tph - ph;
    w.a[k] = lgamma(1.0 + ph);
    w.b[k] = lgamma(alpha + til);
    w.da[k] = digamma_pos(ph + 1.0);
    w.db[k] = digamma_pos(til + alpha);
    w.dc[k] = digamma_pos(tph + alpha + 1.0);
  }
  psync<SUB>();
  for (int k = tid; k < K; k += pthreads<SUB>()) w.c[k] = lgamma(alpha + 1.0 + w.c[k]);
  psync<SUB>();
  if (tid == 0) {
    double A = 0.0, B = 0.0, C = 0.0, cumB = 0.0, cumC = 0.0;
    for (int k = 0; k < K; ++k) {
      A += w.a[k];
      B += w.b[k];
      C += w.c[k];
      cumC += w.dc[k];
      mixgrad[k] = w.da[k] + cumB - cumC;  // the B term is shifted by one cluster (collapsed_mixture.py:90)
      cumB += w.db[k];
    }
    *mix_out = A + B - C + K * (lgamma(1.0 + alpha) - lgamma(alpha));
  }
  psync<SUB>();
}

// ------------------------------------------------------------------------------------------------
// Device-side tail of one bound evaluation inside the conjugate natural-gradient loop (one thread):
// accept / reject (collapsed_vb.py:170-193), buffer rotation, iteration count, trace row, termination tests
// (collapsed_vb.py:229-291 with optimize_parameters() == 0, i.e. no free kernel parameters), carry-over (:294-303).
__device__ inline void cg_decide(gpc_loop_t *L, int phase, double bound_trial, bool failed, double sq, double beta) {
  if (L->done & GPC_DONE_PEER_TIMEOUT) return;  // the statistics of this evaluation are incomplete: no decision is taken on them
  const double bound_old = L->bound_old;
  int evals = 1;
  if (phase == GPC_PHASE_CONJ) {
    const double bound = failed ? bound_old - 1.0 : bound_trial;  // collapsed_vb.py:174-176
    if (bound < bound_old) {                                       // :182 -- fall back to the steepest step
      L->need_retry = 1;
      return;
    }
    bound_trial = bound;
  } else {
    L->need_retry = 0;
    if (failed) {  // :187-190 -- the host restores the accepted point and carries on from there
      L->done |= GPC_DONE_FAULT;
      return;
    }
    evals = 2;
  }
  // the trial point becomes the accepted point: rotate the logits buffers, swap the natural-gradient buffers
  const int xi = L->xi, xp = L->xprev, xt = L->xtrial;
  const int free_buf = (xp >= 0) ? xp : 3 - xi - xt;
  L->xprev = xi;
  L->xi = xt;
  L->xtrial = free_buf;
  L->ngi = 1 - L->ngi;
  L->first = 0;
  const int iteration = L->iteration + evals;
  L->iteration = iteration;
  if (L->n_trace < L->trace_cap) {
    double *row = reinterpret_cast<double *>(L + 1) + 4 * L->n_trace;
    row[0] = (double)iteration;
    row[1] = bound_trial;
    row[2] = sq;
    row[3] = beta;
    L->n_trace += 1;
  }
  int done = 0;
  if (fabs(bound_trial - bound_old) <= L->ftol) done |= GPC_DONE_FTOL;
  if (sq <= L->gtol) done |= GPC_DONE_GTOL;
  if ((double)iteration >= L->maxiter) done |= GPC_DONE_MAXITER;
  if (L->pass_budget > 0) {  // run exactly this many accepted steps (bench / profiling)
    L->pass_budget -= 1;
    if (L->pass_budget == 0) done |= GPC_DONE_BUDGET;
  }
  L->done = done | (L->done & GPC_DONE_PEER_TIMEOUT);
  L->sq_old = sq;
  L->bound_old = bound_trial;
}

struct MohgpFinalizeArgs {
  const double *stats;  // [ybark | phi_hat | H]
  const double *per_k;  // KP x 4
  double *bias;         // out KP : mixgrad_k - 0.5*trace_k - 0.5*mu^T Sy^-1 mu   (MOHGP.py:146-148 + GEMM-form constant)
  double *scalars;      // out: [0]=bound [1]=mixing_prop_bound [2]=H
  double *mixgrad_out;  // out K (nullable): mixing_prop_bound_grad, collapsed_mixture.py:80-94
  double n_total;       // N over all ranks
  const double *const_term;  // device scalar: -0.5*(N*D*log(2pi) + N*Sy_logdet + sum(YTY*Sy_inv))   (MOHGP.py:129-132)
  double alpha;
  int prior, K, KP, DP;
};

__global__ void __launch_bounds__(kClusterThreads) k_mohgp_finalize(const MohgpFinalizeArgs a) {
  __shared__ double mixgrad[kMaxMixK];
  __shared__ MixScratch scratch;
  __shared__ double mix_s;
  const double *phi_hat = a.stats + (size_t)a.KP * a.DP;
  mixing_terms(phi_hat, a.K, a.alpha, a.prior, scratch, mixgrad, &mix_s);
  if (threadIdx.x == 0) {
    const double H = phi_hat[a.KP];
    const double mix = mix_s;
    double ldd = 0.0, quad = 0.0;
    for (int k = 0; k < a.K; ++k) {
      ldd += a.per_k[k * 4 + 0];
      quad += a.per_k[k * 4 + 1];
    }
    a.scalars[0] = (*a.const_term - 0.5 * ldd) + 0.5 * quad + mix + H;  // MOHGP.py:129-135
    a.scalars[1] = mix;
    a.scalars[2] = H;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < a.KP; k += blockDim.x)
    a.bias[k] = (k < a.K) ? (mixgrad[k] - 0.5 * a.per_k[k * 4 + 2]) - 0.5 * a.per_k[k * 4 + 3] : 0.0;
  if (a.mixgrad_out)
    for (int k = threadIdx.x; k < a.K; k += blockDim.x) a.mixgrad_out[k] = mixgrad[k];
}

// ------------------------------------------------------------------------------------------------
// MOHGP posterior in the eigenbasis of A = Ly^-1 Sf Ly^-T  (DESIGN.md section 5).
//
// C_k = I + phi_hat_k A has the same eigenvectors for every cluster and every iteration, so with
// A = Q diag(lam) Q^T (k_mohgp_eig, once per kernel-parameter change), V = Ly Q, U = Ly^-T Q,
// den_i = 1 + phi_hat_k lam_i, c_i = lam_i / den_i and z = U^T ybark_k:
//     log_det_diff_k          = sum_i log(den_i)                                   MOHGP.py:83
//     s_k = Sy^-1 ybark_k     = U z                                                MOHGP.py:88
//     Lambda_inv_k            = V diag(c) V^T - (1e-6/phi_hat_k) I                 MOHGP.py:85 (Sy vs Sy + 1e-6 I)
//     mu_k = Lambda_inv_k s_k = V (c o z) - (1e-6/phi_hat_k) s_k                   MOHGP.py:90
//     s^T Lambda_inv s        = sum c z^2 - (1e-6/phi_hat_k) |s|^2                 MOHGP.py:133
//     tr(Lambda_inv Sy_inv)   = sum c - (1e-6/phi_hat_k) tr(Sy_inv)                MOHGP.py:147
//     w_k = Sy^-1 mu_k        = U (c o z) - (1e-6/phi_hat_k) Sy_inv s_k            (GEMM form of MOHGP.py:144)
// i.e. five D x D matrix-vector products per cluster instead of a Cholesky, a D x D triangular solve and a SYRK.
// phi_hat_k <= 1e-6 takes the reference's Lambda_inv := Sf branch.

// workspace layout (doubles): [lam 64 | U D*D | Ut D*D | Vt D*D | Sy_inv D*D | tr(Sy_inv), sum(Sf o Sy_inv), sweeps, off-norm, min lam, sum lam, 2 spare],
// length rounded up to an even number of doubles so that the whole block is one 16-byte-granular bulk copy
__host__ __device__ inline int eig_ws_len(int D) { return (GPC_MAX_DP + 4 * D * D + 8 + 1) & ~1; }

struct MohgpEigArgs {
  const double *A, *Ly, *Ly_inv, *Sy_inv, *Sf;
  double *ws;
  int D;
};

// One CTA of 1024 threads; round-robin parallel ordering: n/2 disjoint rotations per round.  Stage 1: one-sided (Hestenes) Jacobi, one
// warp per column pair and one barrier per round (4.07 -> 2.7 ms at D = 64); stage 2 (fall-back): two-sided cyclic Jacobi, three
// barrier-separated phases per round.  Latency, not throughput, bounds both.
constexpr int kEigThreads = 1024;
__global__ void __launch_bounds__(kEigThreads) k_mohgp_eig(const MohgpEigArgs a) {
  extern __shared__ __align__(16) double sm[];
  const int D = a.D, n = (D + 1) & ~1, ld = GPC_MAX_DP + 1, tid = threadIdx.x, nth = blockDim.x;
  double *As = sm;                    // n x ld
  double *Qs = As + GPC_MAX_DP * ld;  // n x ld
  double *cs = Qs + GPC_MAX_DP * ld;  // n/2 cosines
  double *sn = cs + GPC_MAX_DP / 2;   // n/2 sines
  double *red_s = sn + GPC_MAX_DP / 2;  // 32
  __shared__ int pp[GPC_MAX_DP / 2], qq[GPC_MAX_DP / 2];
  for (int e = tid; e < n * n; e += nth) {
    const int i = e / n, j = e % n;
    // symmetrise: A is symmetric up to rounding, Jacobi wants it exactly so
    As[i * ld + j] = (i < D && j < D) ? 0.5 * (a.A[i * D + j] + a.A[j * D + i]) : 0.0;
    Qs[i * ld + j] = (i == j) ? 1.0 : 0.0;
  }
  __syncthreads();
  double fro = 0.0;
  for (int e = tid; e < n * n; e += nth) fro += As[(e / n) * ld + (e % n)] * As[(e / n) * ld + (e % n)];
  fro = block_sum(fro, red_s);
  int sweeps = 0;
  double off = 0.0;
  // ---- stage 1: ONE-SIDED Jacobi (Hestenes) on G = A V, V = I: a rotation of the column pair (p, q) makes G_p and G_q orthogonal and
  // touches nothing else, so the n/2 pairs of a round are fully independent -- one warp per pair (three warp-wide dot products, one
  // rotation of two columns of G and of V), ONE block barrier per round instead of three.  At convergence the columns of G are
  // lam_j v_j.  Valid whenever no two eigenvalues are +lam / -lam (always for positive semi-definite A = Ly^-1 Sf Ly^-T); the result is
  // verified by its residual and the two-sided iteration below is the fall-back.
  bool have = false;
  {
    double *Gs = red_s + 32;  // n x ld, behind the reduction scratch
    __shared__ int rotated;
    const int warp = tid >> 5, lane = tid & 31;
    for (int e = tid; e < n * n; e += nth) Gs[(e / n) * ld + (e % n)] = As[(e / n) * ld + (e % n)];
    __syncthreads();
    int sw = 0;
    for (; sw < 30; ++sw) {
      if (tid == 0) rotated = 0;
      __syncthreads();
      for (int round = 0; round < n - 1; ++round) {
        if (warp < n / 2) {
          int p, q;
          if (warp == 0) {
            p = n - 1;
            q = round;
          } else {
            p = (round + warp) % (n - 1);
            q = (round - warp + (n - 1)) % (n - 1);
          }
          const int i0 = lane, i1 = lane + 32;
          const double gp0 = Gs[i0 * ld + p], gq0 = Gs[i0 * ld + q];
          const double gp1 = i1 < n ? Gs[i1 * ld + p] : 0.0, gq1 = i1 < n ? Gs[i1 * ld + q] : 0.0;
          const double al = warp_sum(gp0 * gp0 + gp1 * gp1), be = warp_sum(gq0 * gq0 + gq1 * gq1), ga = warp_sum(gp0 * gq0 + gp1 * gq1);
          if (al > 0.0 && be > 0.0 && fabs(ga) > 1e-15 * sqrt(al * be)) {
            const double zeta = (be - al) / (2.0 * ga);
            const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
            const double c = 1.0 / sqrt(1.0 + t * t), sn_ = c * t;
            Gs[i0 * ld + p] = c * gp0 - sn_ * gq0;
            Gs[i0 * ld + q] = sn_ * gp0 + c * gq0;
            const double vp0 = Qs[i0 * ld + p], vq0 = Qs[i0 * ld + q];
            Qs[i0 * ld + p] = c * vp0 - sn_ * vq0;
            Qs[i0 * ld + q] = sn_ * vp0 + c * vq0;
            if (i1 < n) {
              Gs[i1 * ld + p] = c * gp1 - sn_ * gq1;
              Gs[i1 * ld + q] = sn_ * gp1 + c * gq1;
              const double vp1 = Qs[i1 * ld + p], vq1 = Qs[i1 * ld + q];
              Qs[i1 * ld + p] = c * vp1 - sn_ * vq1;
              Qs[i1 * ld + q] = sn_ * vp1 + c * vq1;
            }
            if (lane == 0) rotated = 1;
          }
        }
        __syncthreads();
      }
      if (rotated == 0) break;
      __syncthreads();
    }
    // eigenvalues as Rayleigh quotients v_j . (A v_j) = v_j . G_j, and the residual |G_j - lam_j v_j|^2 summed over j
    double res = 0.0;
    for (int j = warp; j < n; j += nth / 32) {
      const int i0 = lane, i1 = lane + 32;
      const double g0 = Gs[i0 * ld + j], v0 = Qs[i0 * ld + j], g1 = i1 < n ? Gs[i1 * ld + j] : 0.0, v1 = i1 < n ? Qs[i1 * ld + j] : 0.0;
      const double lj = warp_sum(g0 * v0 + g1 * v1);
      const double d0 = g0 - lj * v0, d1 = g1 - lj * v1;
      res += d0 * d0 + d1 * d1;  // summed over lanes below
      if (lane == 0) Gs[j * ld + j] = lj;               // the diagonal of Gs now carries lam_j (G itself is no longer needed on it)
    }
    res = block_sum(res, red_s);
    have = (sw < 30) && (res <= 1e-24 * fro);
    if (have) {
      for (int j = tid; j < n; j += nth) As[j * ld + j] = Gs[j * ld + j];
      sweeps = sw;
      off = res;
    } else {
      for (int e = tid; e < n * n; e += nth) Qs[(e / n) * ld + (e % n)] = ((e / n) == (e % n)) ? 1.0 : 0.0;
    }
    __syncthreads();
  }
  // ---- stage 2 (fall-back): two-sided cyclic Jacobi
  for (; !have && sweeps < 30; ++sweeps) {
    off = 0.0;
    for (int e = tid; e < n * n; e += nth) {
      const int i = e / n, j = e % n;
      if (i != j) off += As[i * ld + j] * As[i * ld + j];
    }
    off = block_sum(off, red_s);
    if (!(off > 1e-30 * fro)) break;  // off-diagonal norm below 1e-15 of the matrix norm
    for (int round = 0; round < n - 1; ++round) {
      if (tid < n / 2) {
        int p, q;
        if (tid == 0) {
          p = n - 1;
          q = round;
        } else {
          p = (round + tid) % (n - 1);
          q = (round - tid + (n - 1)) % (n - 1);
        }
        if (p > q) {
          const int t = p;
          p = q;
          q = t;
        }
        const double apq = As[p * ld + q], app = As[p * ld + p], aqq = As[q * ld + q];
        double c = 1.0, s = 0.0;
        if (fabs(apq) > 1e-300) {
          const double tau = (aqq - app) / (2.0 * apq);
          const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
          c = 1.0 / sqrt(1.0 + t * t);
          s = t * c;
        }
        pp[tid] = p;
        qq[tid] = q;
        cs[tid] = c;
        sn[tid] = s;
      }
      __syncthreads();
      // columns: [A_ip, A_iq] <- [c A_ip - s A_iq, s A_ip + c A_iq]; same for the eigenvector accumulator
      for (int e = tid; e < (n / 2) * n; e += nth) {
        const int r = e / n, i = e % n;
        const int p = pp[r], q = qq[r];
        const double c = cs[r], s = sn[r];
        const double x = As[i * ld + p], y = As[i * ld + q];
        As[i * ld + p] = c * x - s * y;
        As[i * ld + q] = s * x + c * y;
        const double u = Qs[i * ld + p], v = Qs[i * ld + q];
        Qs[i * ld + p] = c * u - s * v;
        Qs[i * ld + q] = s * u + c * v;
      }
      __syncthreads();
      // rows
      for (int e = tid; e < (n / 2) * n; e += nth) {
        const int r = e / n, j = e % n;
        const int p = pp[r], q = qq[r];
        const double c = cs[r], s = sn[r];
        const double x = As[p * ld + j], y = As[q * ld + j];
        As[p * ld + j] = c * x - s * y;
        As[q * ld + j] = s * x + c * y;
      }
      __syncthreads();
    }
  }
  double *lam = a.ws, *U = a.ws + GPC_MAX_DP, *Ut = U + D * D, *Vt = Ut + D * D, *SyI = Vt + D * D, *scal = SyI + D * D;
  for (int i = tid; i < GPC_MAX_DP; i += nth) lam[i] = (i < D) ? As[i * ld + i] : 0.0;
  for (int e = tid; e < D * D; e += nth) SyI[e] = a.Sy_inv[e];
  // U = Ly^-T Q ; V = Ly Q   (Ly, Ly_inv lower triangular with zero strict upper)
  for (int e = tid; e < D * D; e += nth) {
    const int i = e / D, j = e % D;
    double u = 0.0, v = 0.0;
    for (int m = i; m < D; ++m) u += a.Ly_inv[m * D + i] * Qs[m * ld + j];
    for (int m = 0; m <= i; ++m) v += a.Ly[i * D + m] * Qs[m * ld + j];
    U[i * D + j] = u;
    Ut[j * D + i] = u;
    Vt[j * D + i] = v;
  }
  double t1 = 0.0, t2 = 0.0;
  for (int e = tid; e < D * D; e += nth) {
    if (e / D == e % D) t1 += a.Sy_inv[e];
    t2 += a.Sf[e] * a.Sy_inv[e];
  }
  t1 = block_sum(t1, red_s);
  t2 = block_sum(t2, red_s);
  if (tid == 0) {
    scal[0] = t1;
    scal[1] = t2;
    scal[2] = (double)sweeps;
    scal[3] = sqrt(off / (fro > 0.0 ? fro : 1.0));
    // smallest eigenvalue and eigenvalue sum (= tr A): the jitchol emulation of the posterior needs them for every cluster and every
    // evaluation, so they are found once here (same order of operations for every consumer)
    double lmin = INFINITY, lsum = 0.0;
    for (int i = 0; i < D; ++i) {
      lmin = fmin(lmin, As[i * ld + i]);
      lsum += As[i * ld + i];
    }
    scal[4] = lmin;
    scal[5] = lsum;
  }
}

struct MohgpPostArgs {
  const double *partials;  // num_parts x stats_len per-CTA statistics of the S pass (or the all-reduced vector, num_parts = 1)
  int num_parts;
  double *stats;           // out: [ybark KP x DP | phi_hat KP | H | pad]  (may alias partials when num_parts == 1)
  const double *ws;        // k_mohgp_eig workspace
  const double *Sy_inv;    // D x D
  const double *Sf;        // D x D
  const double *const_term;
  double *Wt;              // out KP x DP
  double *muk_t;           // out K x D
  double *Syi_ybark_t;     // out K x D
  double *per_k;           // out KP x 4 : log_det_diff, s^T Lambda_inv s, tr(Lambda_inv Sy_inv), mu^T Sy^-1 mu
  double *bias;            // out KP
  double *scalars;         // out [0]=bound [1]=mix [2]=H
  double *mixgrad_out;     // out K (nullable)
  unsigned int *counter;   // zero on entry, reset on exit
  int *info;               // NOT cleared here: the caller's flag word (cleared by the launcher)
  double alpha;
  int prior, K, D, KP, DP;
  const double *dots;  // loop mode: [0] = natgrad.grad of the G pass
  const double *beta;  // loop mode: beta of the conjugate S pass
  gpc_loop_t *loop;    // nullable
  int phase;
  const gpc_peers_t *peers;  // nullable: statistics = the W inbox vectors of the peer exchange (partials / num_parts ignored)
};

// partial of y[r] = sum_j Mt[j*D + r] x[j] over j == pg (mod 4), for r < D; x in shared memory, Mt in shared memory
// (staged by bulk copy) or global memory.
// D == GPC_MAX_DP (= 64, config 3): no clamping, compile-time strides; thread group pg takes the 16 CONSECUTIVE entries
// j in [16 pg, 16 pg + 16) of x (read as 8 double2), four independent FMA chains.
__device__ __forceinline__ double matvec_t_part64(const double *Mt, const double *x_s, int r, int pg) {
  const double *m = Mt + (size_t)(16 * pg) * GPC_MAX_DP + r;
  const double2 *x2 = reinterpret_cast<const double2 *>(x_s + 16 * pg);
  double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
#pragma unroll
  for (int u = 0; u < 8; u += 2) {
    const double2 xa = x2[u], xb = x2[u + 1];
    v0 = fma(m[(2 * u + 0) * GPC_MAX_DP], xa.x, v0);
    v1 = fma(m[(2 * u + 1) * GPC_MAX_DP], xa.y, v1);
    v2 = fma(m[(2 * u + 2) * GPC_MAX_DP], xb.x, v2);
    v3 = fma(m[(2 * u + 3) * GPC_MAX_DP], xb.y, v3);
  }
  return (v0 + v1) + (v2 + v3);
}
__device__ __forceinline__ double matvec_t_part_any(const double *Mt, const double *x_s, int D, int r, int pg);
__device__ __forceinline__ double matvec_t_part(const double *Mt, const double *x_s, int D, int r, int pg) {
  return D == GPC_MAX_DP ? matvec_t_part64(Mt, x_s, r, pg) : matvec_t_part_any(Mt, x_s, D, r, pg);
}
__device__ __forceinline__ double matvec_t_part_any(const double *Mt, const double *x_s, int D, int r, int pg) {
  // unconditional loads from clamped addresses (a predicated load is not hoisted above its use by the compiler)
  const int rc = r < D ? r : D - 1;
  double v[4] = {0.0, 0.0, 0.0, 0.0};
  for (int j = pg; j < D; j += 64) {
    double m[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int jj = j + 4 * u;
      m[u] = Mt[(jj < D ? jj : D - 1) * D + rc];
    }
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int jj = j + 4 * u;
      v[u & 3] += m[u] * (jj < D ? x_s[jj] : 0.0);
    }
  }
  return r < D ? (v[0] + v[1]) + (v[2] + v[3]) : 0.0;
}

// shared memory of the posterior routine: static in k_mohgp_post, carved from the (then idle) dynamic ring memory in the fused loop kernel
struct PostSmem {
  double red[4][GPC_MAX_DP], red1[4][GPC_MAX_DP], red2[4][GPC_MAX_DP];
  double yb_s[GPC_MAX_DP], z_s[GPC_MAX_DP], cz_s[GPC_MAX_DP], s_s[GPC_MAX_DP], mu_s[GPC_MAX_DP], w_s[GPC_MAX_DP], c_s[GPC_MAX_DP];
  double red_s[32];
  double mixgrad[kMaxMixK];
  double phi_hat_sh[kMaxMixK];
  MixScratch scratch;
  double mix_s, H_s;
  double xrow[GPC_MAX_PEERS][GPC_MAX_DP + 8];  // fused exchange: the W ranks' rows of this cluster
  uint64_t stage_bar;
  int xchg_ok;
  int is_last;
};

// the replicated D x D operands (eigen workspace incl. a copy of Sy_inv: <= 132 KB) -> shared memory by one bulk copy (one thread)
__device__ __forceinline__ void post_stage_ws(const double *ws, int D, PostSmem &sm, double *ws_s) {
  mbar_init(&sm.stage_bar, 1);
  mbar_fence_init();
  const uint32_t bytes = (uint32_t)eig_ws_len(D) * 8u;  // even number of doubles: a multiple of 16 bytes
  mbar_arrive_expect_tx(&sm.stage_bar, bytes);
  bulk_g2s(ws_s, ws, bytes, &sm.stage_bar);
}

// phi_hat_k = sum over `num_parts` vectors of their phi_hat entry k, by ONE WARP (all 32 lanes), fixed order.  Every consumer of
// phi_hat_k (the cluster's own CTA, the CTA that evaluates the mixing-proportion terms) uses this routine: bit-identical values.
__device__ __forceinline__ double reduce_phi_hat_warp(const double *partials, int num_parts, int len, int KP, int DP, int k) {
  const int lane = threadIdx.x & 31;
  const double *src = partials + (size_t)KP * DP + k;
  const int last = num_parts - 1;
  double tot = 0.0;
  // lane takes the vectors lane, lane + 32, ...: eight loads in flight per round (one L2 round trip for up to 256 vectors), clamped
  // addresses as in post_reduce_row
  for (int p0 = lane; p0 < num_parts; p0 += 256) {
    double w[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) w[u] = __ldcg(src + min(p0 + 32 * u, last) * len);
#pragma unroll
    for (int u = 0; u < 8; ++u) w[u] = (p0 + 32 * u <= last) ? w[u] : 0.0;
    tot += ((w[0] + w[1]) + (w[2] + w[3])) + ((w[4] + w[5]) + (w[6] + w[7]));
  }
  return warp_sum(tot);
}

// this cluster's statistics: fixed-order sum over `num_parts` vectors of stride `len` -> sm.yb_s[0..DP); returns phi_hat_k on every
// thread.  Ends with a barrier.  (8 independent loads in flight per thread.)
template <bool SUB>
__device__ __forceinline__ double post_reduce_row(const double *partials, int num_parts, int len, int k, int KP, int DP, PostSmem &sm) {
  const int tid = threadIdx.x, r = tid & 63, pg = tid >> 6;
  {
    double v[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    {
      const double *src = partials + (size_t)k * DP + (r < DP ? r : DP - 1);
      const int last = num_parts - 1;
      // up to 160 partial vectors (one per SM and then some): all 40 loads of this thread are issued before the first
      // use -- clamped, unconditional addresses, because a predicated load is not hoisted above its use.  32-bit element offsets
      // (num_parts * len < 2^31): one multiply-add per address.
      for (int p0 = pg; p0 < num_parts; p0 += 160) {
        double w[40];
        const int nmine = (last - p0) >> 2;  // this thread's partials are p0 + 4u, u = 0 .. nmine
        const int step4 = 4 * len, off0 = p0 * len;
#pragma unroll
        for (int u = 0; u < 40; ++u) w[u] = __ldcg(src + (off0 + min(u, nmine) * step4));
#pragma unroll
        for (int u = 0; u < 40; ++u) v[u & 7] += (u <= nmine) ? w[u] : 0.0;
      }
    }
    sm.red[pg][r] = ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
  }
  double phi_hat = 0.0;
  if (tid < 32) {
    phi_hat = reduce_phi_hat_warp(partials, num_parts, len, KP, DP, k);
    if (tid == 0) sm.red_s[0] = phi_hat;
  }
  psync<SUB>();
  phi_hat = sm.red_s[0];
  if (tid < DP) sm.yb_s[tid] = (sm.red[0][tid] + sm.red[1][tid]) + (sm.red[2][tid] + sm.red[3][tid]);
  psync<SUB>();
  return phi_hat;
}

// per-cluster posterior of cluster k from sm.yb_s / phi_hat (eigenbasis form, see above); ws_s: the staged workspace.
// Three dependent levels of matrix-vector products (z; then s, V(c o z), U(c o z) together; then Sy_inv s), each level = partial
// products by 4 thread groups + one barrier + the 4-way sums; one combined block reduction for the five per-cluster scalars.
template <bool SUB>
__device__ __forceinline__ void post_cluster_math(const MohgpPostArgs &a, int k, double phi_hat, PostSmem &sm, const double *ws_s,
                                                  unsigned long long *dbg = nullptr) {
  const int D = a.D, DP = a.DP, tid = threadIdx.x;
  const int r = tid & 63, pg = tid >> 6;
  double (*red)[GPC_MAX_DP] = sm.red;
  double (*red1)[GPC_MAX_DP] = sm.red1;
  double (*red2)[GPC_MAX_DP] = sm.red2;
  double *yb_s = sm.yb_s, *z_s = sm.z_s, *cz_s = sm.cz_s, *s_s = sm.s_s, *mu_s = sm.mu_s, *w_s = sm.w_s, *c_s = sm.c_s;
  const double *lam = ws_s, *U = ws_s + GPC_MAX_DP, *Ut = U + D * D, *Vt = Ut + D * D, *syi_s = Vt + D * D, *scal = syi_s + D * D;
  auto sum4 = [](double (*q)[GPC_MAX_DP], int i) { return (q[0][i] + q[1][i]) + (q[2][i] + q[3][i]); };
  if (k < a.K) {
    // ---- den_i = 1 + jitter + phi_hat lam_i with GPy's jitchol retry: a failed factorisation <=> some den_i <= 0
    double jitter = 0.0;
    bool ok = false;
    {
      const double lmin = scal[4];
      ok = (1.0 + phi_hat * lmin > 0.0) && isfinite(phi_hat);  // first attempt, no jitter: the only one taken in a healthy run
      if (!ok && isfinite(phi_hat)) {
        const double diag_mean = 1.0 + phi_hat * scal[5] / D;
        for (int attempt = 1; attempt < 6 && !ok; ++attempt) {
          if (attempt == 1) jitter = diag_mean * 1e-6;
          else jitter *= 10.0;
          if (!isfinite(jitter)) break;
          ok = (1.0 + jitter + phi_hat * lmin > 0.0);
        }
      }
    }
    if (!ok) {
      if (tid == 0) atomicOr(a.info, GPC_INFO_NOT_PD);
      for (int d = tid; d < DP; d += pthreads<SUB>()) a.Wt[(size_t)k * DP + d] = 0.0;
      if (tid < 4) a.per_k[k * 4 + tid] = 0.0;
    } else {
      const bool tiny = !(phi_hat > 1e-6);
      const double shift = tiny ? 0.0 : 1e-6 / phi_hat;
      if (dbg) dbg[0] = globaltimer_ns();
      // level 1: z = U^T ybark
      red[pg][r] = matvec_t_part(U, yb_s, D, r, pg);
      psync<SUB>();
      double ldd = 0.0;
      if (tid < D) {
        const double z = sum4(red, tid);
        const double pl = phi_hat * lam[tid];
        const double den = (1.0 + jitter) + pl;
        const double c = (jitter == 0.0) ? lam[tid] / den : (jitter + pl) / (den * phi_hat);
        z_s[tid] = z;
        c_s[tid] = c;
        cz_s[tid] = c * z;
        ldd = (jitter == 0.0) ? log1p(pl) : log(den);
      }
      psync<SUB>();
      if (dbg) dbg[1] = globaltimer_ns();
      // level 2: s = U z ; V (c o z) ; U (c o z)
      red[pg][r] = matvec_t_part(Ut, z_s, D, r, pg);
      red1[pg][r] = tiny ? 0.0 : matvec_t_part(Vt, cz_s, D, r, pg);
      red2[pg][r] = tiny ? 0.0 : matvec_t_part(Ut, cz_s, D, r, pg);
      psync<SUB>();
      if (tid < D) {
        s_s[tid] = sum4(red, tid);
        mu_s[tid] = sum4(red1, tid);
        w_s[tid] = sum4(red2, tid);
      }
      psync<SUB>();
      if (dbg) dbg[2] = globaltimer_ns();
      // level 3
      if (tiny) {
        // Lambda_inv := Sf (MOHGP.py:85): mu = Sf^T s, then w = Sy_inv mu (Sy_inv symmetric)
        red[pg][r] = matvec_t_part(a.Sf, s_s, D, r, pg);
        psync<SUB>();
        if (tid < D) mu_s[tid] = sum4(red, tid);
        psync<SUB>();
        red[pg][r] = matvec_t_part(syi_s, mu_s, D, r, pg);
        psync<SUB>();
        if (tid < D) w_s[tid] = sum4(red, tid);
      } else {
        red[pg][r] = matvec_t_part(syi_s, s_s, D, r, pg);  // h = Sy_inv s
        psync<SUB>();
        if (tid < D) {
          const double h = sum4(red, tid);
          mu_s[tid] -= shift * s_s[tid];
          w_s[tid] -= shift * h;
        }
      }
      if (dbg) dbg[3] = globaltimer_ns();
      // the five per-cluster scalars in ONE block reduction (each thread < D touches only its own entries of mu_s / w_s above)
      double quad = 0.0, tr = 0.0, m2 = 0.0, ss = 0.0;
      if (tid < D) {
        if (tiny) quad = s_s[tid] * mu_s[tid];
        else {
          quad = c_s[tid] * z_s[tid] * z_s[tid];
          ss = s_s[tid] * s_s[tid];
          tr = c_s[tid];
        }
        m2 = mu_s[tid] * w_s[tid];
        a.muk_t[(size_t)k * D + tid] = mu_s[tid];
        if (a.Syi_ybark_t) a.Syi_ybark_t[(size_t)k * D + tid] = s_s[tid];
      }
      for (int d = tid; d < DP; d += pthreads<SUB>()) a.Wt[(size_t)k * DP + d] = (d < D) ? w_s[d] : 0.0;
      ldd = warp_sum(ldd);
      quad = warp_sum(quad);
      ss = warp_sum(ss);
      tr = warp_sum(tr);
      m2 = warp_sum(m2);
      if ((tid & 31) == 0 && tid < 64) {  // D <= 64: only the first two warps hold contributions
        double *q = sm.red_s + (tid >> 5) * 8;
        q[0] = ldd;
        q[1] = quad;
        q[2] = ss;
        q[3] = tr;
        q[4] = m2;
      }
      psync<SUB>();
      if (dbg) dbg[4] = globaltimer_ns();
      if (tid == 0) {
        const double *q = sm.red_s;
        ldd = q[0] + q[8];
        quad = q[1] + q[9];
        ss = q[2] + q[10];
        tr = q[3] + q[11];
        m2 = q[4] + q[12];
        a.per_k[k * 4 + 0] = ldd;
        a.per_k[k * 4 + 1] = tiny ? quad : quad - shift * ss;
        a.per_k[k * 4 + 2] = tiny ? scal[1] : tr - shift * scal[0];
        a.per_k[k * 4 + 3] = m2;
      }
    }
  } else {  // pad clusters: zero weights so that the G pass sees finite numbers
    for (int d = tid; d < DP; d += pthreads<SUB>()) a.Wt[(size_t)k * DP + d] = 0.0;
    if (tid < 4) a.per_k[k * 4 + tid] = 0.0;
  }
}

// Tail of a bound evaluation, in two parts.
// post_tail_prepare: entropy H, mixing-proportion bound and gradient (collapsed_mixture.py:66-94) from sm.phi_hat_sh[0..K) -- needs the
//   statistics only, not the per-cluster posteriors: the fused loop kernel runs it on a CTA of its own WHILE the cluster CTAs work.
// post_tail_finish: the per-cluster scalars -> bound (MOHGP.py:124-135), G-pass bias (:146-148), and in loop mode the accept /
//   reject / termination decision.  `partials` / `num_parts` / `len`: where the entropy partials are.
template <bool SUB>
__device__ __forceinline__ void post_tail_prepare(const MohgpPostArgs &a, PostSmem &sm, const double *partials, int num_parts, int len) {
  const int KP = a.KP, DP = a.DP, tid = threadIdx.x;
  // every load of this serial tail is issued by a different thread (one L2 round trip, not num_parts of them)
  double H = 0.0;
  for (int p = tid; p < num_parts; p += pthreads<SUB>()) H += __ldcg(partials + (size_t)p * len + (size_t)KP * DP + KP);
  H = block_sum<SUB>(H, sm.red_s);
  if (tid == 0) sm.H_s = H;
  psync<SUB>();
  mixing_terms<SUB>(sm.phi_hat_sh, a.K, a.alpha, a.prior, sm.scratch, sm.mixgrad, &sm.mix_s);
}
// the same with sm.H_s already in place
template <bool SUB>
__device__ __forceinline__ void post_tail_mixing(const MohgpPostArgs &a, PostSmem &sm) {
  mixing_terms<SUB>(sm.phi_hat_sh, a.K, a.alpha, a.prior, sm.scratch, sm.mixgrad, &sm.mix_s);
}
template <bool SUB>
__device__ __forceinline__ void post_tail_finish(const MohgpPostArgs &a, PostSmem &sm, int phase) {
  const int KP = a.KP, DP = a.DP, tid = threadIdx.x;
  double *mixgrad = sm.mixgrad;
  double ldd = 0.0, quad = 0.0;
  for (int kk = tid; kk < a.K; kk += pthreads<SUB>()) {
    ldd += __ldcg(a.per_k + kk * 4 + 0);
    quad += __ldcg(a.per_k + kk * 4 + 1);
  }
  ldd = block_sum<SUB>(ldd, sm.red_s);
  quad = block_sum<SUB>(quad, sm.red_s);
  if (tid == 0) {
    const double H = sm.H_s, mix = sm.mix_s;
    const double bound = (*a.const_term - 0.5 * ldd) + 0.5 * quad + mix + H;
    a.stats[(size_t)KP * DP + KP] = H;
    a.scalars[0] = bound;
    a.scalars[1] = mix;
    a.scalars[2] = H;
    *a.counter = 0u;
    if (a.loop) {
      // loop mode: the flag word is consumed here and re-armed for the next evaluation (no memset between kernels)
      cg_decide(a.loop, phase, bound, (atomicExch(a.info, 0) & GPC_INFO_NOT_PD) != 0, __ldcg(a.dots), __ldcg(a.beta));
    }
  }
  for (int kk = tid; kk < KP; kk += pthreads<SUB>())
    a.bias[kk] = (kk < a.K) ? (mixgrad[kk] - 0.5 * __ldcg(a.per_k + kk * 4 + 2)) - 0.5 * __ldcg(a.per_k + kk * 4 + 3) : 0.0;
  if (a.mixgrad_out)
    for (int kk = tid; kk < a.K; kk += pthreads<SUB>()) a.mixgrad_out[kk] = mixgrad[kk];
}

__global__ void __launch_bounds__(kClusterThreads) k_mohgp_post(const MohgpPostArgs a) {
  __shared__ PostSmem sm;
  if (loop_skip(a.loop, a.phase)) return;
  const int phase = loop_phase(a.loop, a.phase);  // every CTA reads it here, before the last CTA's decision can change it
  const int D = a.D, DP = a.DP, KP = a.KP, k = blockIdx.x, tid = threadIdx.x;
  const int len = KP * DP + KP + 2;
  const double *partials = a.partials;
  int num_parts = a.num_parts;
  // the two N-sized passes have swept the L2 since the last evaluation: the replicated D x D operands (eigen
  // workspace incl. a copy of Sy_inv: <= 132 KB) are pulled into shared memory by one bulk copy while the statistics are being
  // reduced (and, with rows sharded over GPUs, while the peers' statistics are still on their way), so that the five dependent
  // matrix-vector products below never wait on HBM
  extern __shared__ __align__(128) double post_dyn[];
  double *ws_s = post_dyn;
  if (tid == 0) post_stage_ws(a.ws, D, sm, ws_s);
  if (a.peers != nullptr) {
    const unsigned long long e = a.peers->epochs[0];  // gpc_reduce_push of this rank has pushed exchange e
    if (tid == 0) sm.xchg_ok = xchg_wait(a.peers, 0, (int)(e & 1), e) ? 1 : 0;
    __syncthreads();
    if (!sm.xchg_ok && tid == 0) {  // a peer never arrived: a hard fault of the exchange, reported as such (never as not-PD)
      atomicOr(a.info, GPC_INFO_PEER_TIMEOUT);
      if (a.loop != nullptr) atomicOr(&a.loop->done, GPC_DONE_PEER_TIMEOUT);
    }
    partials = a.peers->inbox[a.peers->rank] + xchg_off_stats(a.peers->world, len, (int)(e & 1), 0);
    num_parts = a.peers->world;
  }
  const double phi_hat = post_reduce_row<false>(partials, num_parts, len, k, KP, DP, sm);
  if (tid < DP) a.stats[(size_t)k * DP + tid] = sm.yb_s[tid];
  if (tid == 0) a.stats[(size_t)KP * DP + k] = phi_hat;

  mbar_wait(&sm.stage_bar, 0);
  post_cluster_math<false>(a, k, phi_hat, sm, ws_s);

  // ---- the last CTA to finish assembles the bound and the G-pass bias (MOHGP.py:124-135, :146-148)
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned int done = atomicAdd(a.counter, 1u);
    sm.is_last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (!sm.is_last) return;
  __threadfence();
  for (int kk = tid; kk < a.K; kk += blockDim.x) sm.phi_hat_sh[kk] = __ldcg(a.stats + (size_t)KP * DP + kk);
  __syncthreads();
  post_tail_prepare<false>(a, sm, partials, num_parts, len);
  post_tail_finish<false>(a, sm, phase);
}

// ------------------------------------------------------------------------------------------------
// dL/dSf and dL/dSy of the MOHGP bound (MOHGP.update_kern_grads, MOHGP.py:92-122) in the eigenbasis of A.
//
// The reference forms, per cluster, tmp = L_k^-1 Ly^-1, B_inv_k = phi_hat_k tmp^T tmp (a triangular solve and a D^3 product each) and
// sums K dense D x D matrices.  With C_k = Q diag(den_k) Q^T, den_ki = 1 + jitter_k + phi_hat_k lam_i and U = Ly^-T Q:
//     B_inv_k              = phi_hat_k U diag(1/den_k) U^T
//     sum_k B_inv_k        = U diag(sum_k phi_hat_k/den_ki) U^T                 ONE D x D product instead of K
//     sum_k B_inv_k/ph_k   = U diag(sum_{k: ph_k > 1e-6} 1/den_ki) U^T
//     tmp1_k = (I - Sf B_inv_k)^T s_k = s_k - B_inv_k (Sf s_k)       (Sf symmetric)   MOHGP.py:100-101
//     Byk    = B_inv_k ybark_k                                                        MOHGP.py:111
//   dL_dSf = 1/2 sum_k tmp1_k tmp1_k^T - 1/2 sum_k B_inv_k                            MOHGP.py:104-106
//   dL_dSy = 1/2 [ sum_{k: ph_k > 1e-6} (Byk Byk^T / ph_k^3 - s_k s_k^T / ph_k - B_inv_k / ph_k) + (K - N) Sy_inv + Sy_inv YTY Sy_inv ]   :112-118
// Kernel 1 (one CTA per cluster): the per-cluster vectors and weights, four matrix-vector products each.
// Kernel 2 (one CTA): the two D x D matrices.
struct MohgpKgradArgs {
  const double *stats;    // [ybark KP x DP | phi_hat KP | ...] at the current point
  const double *ws;       // k_mohgp_eig workspace
  const double *Sf;       // D x D
  const double *s_t;      // K x D : rows s_k = Sy^-1 ybark_k (Syi_ybark_t of the posterior kernel)
  double *t1, *by;        // out K x D : tmp1_k, Byk
  double *w1, *w2;        // out K x D : phi_hat_k / den_ki ; 1 / den_ki (zero row when phi_hat_k <= 1e-6)
  double *cw;             // out K x 2 : 1/phi_hat_k^3, 1/phi_hat_k (zeros when phi_hat_k <= 1e-6)
  int *info;
  int K, D, KP, DP;
};
__global__ void __launch_bounds__(kClusterThreads) k_mohgp_kgrad_cluster(const MohgpKgradArgs a) {
  __shared__ double red[4][GPC_MAX_DP];
  __shared__ double yb_s[GPC_MAX_DP], s_s[GPC_MAX_DP], v_s[GPC_MAX_DP], u_s[GPC_MAX_DP], q_s[GPC_MAX_DP];
  const int D = a.D, DP = a.DP, KP = a.KP, k = blockIdx.x, tid = threadIdx.x, r = tid & 63, pg = tid >> 6;
  const double *lam = a.ws, *U = a.ws + GPC_MAX_DP, *Ut = U + D * D, *scal = Ut + 3 * D * D;
  const double phi_hat = a.stats[(size_t)KP * DP + k];
  auto sum4 = [&](int i) { return (red[0][i] + red[1][i]) + (red[2][i] + red[3][i]); };
  if (tid < D) {
    yb_s[tid] = a.stats[(size_t)k * DP + tid];
    s_s[tid] = a.s_t[(size_t)k * D + tid];
  }
  // jitter exactly as the posterior (GPy jitchol emulation)
  double jitter = 0.0;
  bool ok = (1.0 + phi_hat * scal[4] > 0.0) && isfinite(phi_hat);
  if (!ok && isfinite(phi_hat)) {
    const double diag_mean = 1.0 + phi_hat * scal[5] / D;
    for (int attempt = 1; attempt < 6 && !ok; ++attempt) {
      if (attempt == 1) jitter = diag_mean * 1e-6;
      else jitter *= 10.0;
      if (!isfinite(jitter)) break;
      ok = (1.0 + jitter + phi_hat * scal[4] > 0.0);
    }
  }
  if (!ok && tid == 0) atomicOr(a.info, GPC_INFO_NOT_PD);
  const bool big = phi_hat > 1e-6;
  __syncthreads();
  // v = Sf s  (global Sf, symmetric)
  red[pg][r] = matvec_t_part(a.Sf, s_s, D, r, pg);
  __syncthreads();
  if (tid < D) v_s[tid] = sum4(tid);
  __syncthreads();
  // u = U^T v ; q = U^T ybark   -> scaled by phi_hat / den
  red[pg][r] = matvec_t_part(U, v_s, D, r, pg);
  __syncthreads();
  double den = 1.0;
  if (tid < D) {
    den = (1.0 + jitter) + phi_hat * lam[tid];
    u_s[tid] = sum4(tid) * (phi_hat / den);
    a.w1[(size_t)k * D + tid] = ok ? phi_hat / den : 0.0;
    a.w2[(size_t)k * D + tid] = (ok && big) ? 1.0 / den : 0.0;
  }
  __syncthreads();
  red[pg][r] = matvec_t_part(U, yb_s, D, r, pg);
  __syncthreads();
  if (tid < D) q_s[tid] = sum4(tid) * (phi_hat / den);
  __syncthreads();
  // tmp1 = s - U u ; Byk = U q
  red[pg][r] = matvec_t_part(Ut, u_s, D, r, pg);
  __syncthreads();
  if (tid < D) a.t1[(size_t)k * D + tid] = ok ? s_s[tid] - sum4(tid) : 0.0;
  __syncthreads();
  red[pg][r] = matvec_t_part(Ut, q_s, D, r, pg);
  __syncthreads();
  if (tid < D) a.by[(size_t)k * D + tid] = ok ? sum4(tid) : 0.0;
  if (tid == 0) {
    a.cw[2 * k + 0] = (ok && big) ? 1.0 / (phi_hat * phi_hat * phi_hat) : 0.0;
    a.cw[2 * k + 1] = (ok && big) ? 1.0 / phi_hat : 0.0;
  }
}

struct MohgpKgradFinalArgs {
  const double *t1, *by, *s_t, *w1, *w2, *cw;  // K x D (cw: K x 2)
  const double *ws, *Sy_inv, *YTY;
  double *dL_dSf, *dL_dSy;                     // out D x D
  double *scratch;                             // D x D : Sy_inv YTY
  double n_total;
  int K, D;
};
constexpr int kKgradThreads = 1024;
__global__ void __launch_bounds__(kKgradThreads) k_mohgp_kgrad_final(const MohgpKgradFinalArgs a) {
  __shared__ double a1[GPC_MAX_DP], a2[GPC_MAX_DP];
  const int D = a.D, K = a.K, tid = threadIdx.x, nth = blockDim.x;
  const double *U = a.ws + GPC_MAX_DP;
  for (int i = tid; i < D; i += nth) {
    double x = 0.0, y = 0.0;
    for (int k = 0; k < K; ++k) {
      x += a.w1[(size_t)k * D + i];
      y += a.w2[(size_t)k * D + i];
    }
    a1[i] = x;
    a2[i] = y;
  }
  for (int e = tid; e < D * D; e += nth) {  // M = Sy_inv YTY
    const int i = e / D, j = e % D;
    double v = 0.0;
    for (int m = 0; m < D; ++m) v += a.Sy_inv[i * D + m] * a.YTY[m * D + j];
    a.scratch[e] = v;
  }
  __syncthreads();
  for (int e = tid; e < D * D; e += nth) {
    const int i = e / D, j = e % D;
    double tt = 0.0, bb = 0.0, ss = 0.0;
    for (int k = 0; k < K; ++k) {
      tt += a.t1[(size_t)k * D + i] * a.t1[(size_t)k * D + j];
      bb += a.by[(size_t)k * D + i] * a.by[(size_t)k * D + j] * a.cw[2 * k];
      ss += a.s_t[(size_t)k * D + i] * a.s_t[(size_t)k * D + j] * a.cw[2 * k + 1];
    }
    double u1 = 0.0, u2 = 0.0, sys = 0.0;
    for (int m = 0; m < D; ++m) {
      const double uu = U[i * D + m] * U[j * D + m];
      u1 += uu * a1[m];
      u2 += uu * a2[m];
      sys += a.scratch[i * D + m] * a.Sy_inv[m * D + j];
    }
    a.dL_dSf[e] = 0.5 * tt - 0.5 * u1;
    a.dL_dSy[e] = 0.5 * (((bb - ss) - u2) + ((double)K - a.n_total) * a.Sy_inv[e] + sys);
  }
}

// ------------------------------------------------------------------------------------------------
// MOG: features f = [x_i x_j (i <= j, row-major upper) | x_i], NF = D(D+1)/2 + D, computed from data
// centred at `centre` (the posterior is translation invariant; centring limits cancellation).
struct MogClusterArgs {
  const double *stats;  // [T KP x DP | phi_hat KP | H] ; T[k][f] = sum_n phi_nk f_n
  const double *m0c;    // D : prior mean, centred
  const double *S0;     // D x D
  double *Wt;           // out KP x DP
  double *per_k;        // out KP x 4 : kN, vN, Sns_halflogdet, grad const (without mixgrad)
  double *mun_c;        // out K x D (centred posterior means, row k)
  double *Sns;          // out K x D x D
  double *Sns_inv;      // out K x D x D
  int *info;
  double k0, v0;
  int K, D, KP, DP;
  const gpc_loop_t *loop;  // nullable
  int phase;
};

constexpr int kMogMaxD = 10;

__global__ void k_mog_cluster(const MogClusterArgs a) {
  if (loop_skip(a.loop, a.phase)) return;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= a.KP) return;
  const int D = a.D;
  if (k >= a.K) {
    for (int d = 0; d < a.DP; ++d) a.Wt[(size_t)k * a.DP + d] = 0.0;
    for (int q = 0; q < 4; ++q) a.per_k[k * 4 + q] = 0.0;
    return;
  }
  const double *T = a.stats + (size_t)k * a.DP;
  const double phi_hat = a.stats[(size_t)a.KP * a.DP + k];
  const double kN = phi_hat + a.k0, vN = phi_hat + a.v0;  // MOG.py:54-55
  const int NQ = D * (D + 1) / 2;
  double mun[kMogMaxD], S[kMogMaxD][kMogMaxD], L[kMogMaxD][kMogMaxD], Si[kMogMaxD][kMogMaxD];
  for (int i = 0; i < D; ++i) mun[i] = (a.k0 * a.m0c[i] + T[NQ + i]) / kN;  // MOG.py:58
  {
    int f = 0;
    for (int i = 0; i < D; ++i)
      for (int j = i; j < D; ++j, ++f) {
        const double v = a.S0[i * D + j] + T[f] + a.k0 * a.m0c[i] * a.m0c[j] - kN * mun[i] * mun[j];  // MOG.py:60
        S[i][j] = v;
        S[j][i] = v;
      }
  }
  // jitchol
  bool ok = false;
  double dsum = 0.0, dmin = INFINITY;
  for (int i = 0; i < D; ++i) {
    dsum += S[i][i];
    dmin = fmin(dmin, S[i][i]);
  }
  double jitter = 0.0;
  for (int attempt = 0; attempt < 6 && !ok; ++attempt) {
    if (attempt == 1) {
      if (!(dmin > 0.0)) break;
      jitter = dsum / D * 1e-6;
    } else if (attempt > 1) jitter *= 10.0;
    ok = true;
    for (int j = 0; j < D && ok; ++j) {
      double d = S[j][j] + jitter;
      for (int m = 0; m < j; ++m) d -= L[j][m] * L[j][m];
      if (!(d > 0.0) || isinf(d)) {
        ok = false;
        break;
      }
      L[j][j] = sqrt(d);
      for (int i = j + 1; i < D; ++i) {
        double v = S[i][j];
        for (int m = 0; m < j; ++m) v -= L[i][m] * L[j][m];
        L[i][j] = v / L[j][j];
      }
    }
  }
  if (!ok) {
    atomicOr(a.info, GPC_INFO_NOT_PD);
    for (int d = 0; d < a.DP; ++d) a.Wt[(size_t)k * a.DP + d] = 0.0;
    for (int q = 0; q < 4; ++q) a.per_k[k * 4 + q] = 0.0;
    return;
  }
  double hld = 0.0;
  for (int i = 0; i < D; ++i) hld += log(L[i][i]);  // np_utilities.py:19
  // inverse from the factor (dpotri): Si = L^-T L^-1
  double Li[kMogMaxD][kMogMaxD];
  for (int j = 0; j < D; ++j) {
    for (int i = 0; i < D; ++i) Li[i][j] = 0.0;
    Li[j][j] = 1.0 / L[j][j];
    for (int i = j + 1; i < D; ++i) {
      double v = 0.0;
      for (int m = j; m < i; ++m) v -= L[i][m] * Li[m][j];
      Li[i][j] = v / L[i][i];
    }
  }
  for (int i = 0; i < D; ++i)
    for (int j = i; j < D; ++j) {
      double v = 0.0;
      for (int m = j; m < D; ++m) v += Li[m][i] * Li[m][j];
      Si[i][j] = v;
      Si[j][i] = v;
    }
  // gradient weights: -0.5*vN*(x-mu)^T Si (x-mu) expanded on the features   (MOG.py:74-79)
  double Sm[kMogMaxD], mSm = 0.0;
  for (int i = 0; i < D; ++i) {
    double v = 0.0;
    for (int j = 0; j < D; ++j) v += Si[i][j] * mun[j];
    Sm[i] = v;
    mSm += v * mun[i];
  }
  {
    double *w = a.Wt + (size_t)k * a.DP;
    int f = 0;
    for (int i = 0; i < D; ++i)
      for (int j = i; j < D; ++j, ++f) w[f] = (i == j) ? -0.5 * vN * Si[i][i] : -vN * Si[i][j];
    for (int i = 0; i < D; ++i) w[NQ + i] = vN * Sm[i];
    for (int d = NQ + D; d < a.DP; ++d) w[d] = 0.0;
  }
  double dg = 0.0;
  for (int d = 0; d < D; ++d) dg += digamma_pos((vN - d) / 2.0);
  const double gconst = -0.5 * D / kN + 0.5 * dg - hld - 1.0;  // MOG.py:79 without mixing_prop_bound_grad
  a.per_k[k * 4 + 0] = kN;
  a.per_k[k * 4 + 1] = vN;
  a.per_k[k * 4 + 2] = hld;
  a.per_k[k * 4 + 3] = gconst - 0.5 * vN * mSm;
  for (int i = 0; i < D; ++i) {
    a.mun_c[(size_t)k * D + i] = mun[i];
    for (int j = 0; j < D; ++j) {
      a.Sns[(size_t)k * D * D + i * D + j] = S[i][j];
      a.Sns_inv[(size_t)k * D * D + i * D + j] = Si[i][j];
    }
  }
}

struct MogFinalizeArgs {
  const double *stats;
  const double *per_k;
  double *bias;     // out KP
  double *scalars;  // [0]=bound [1]=mix [2]=H
  double *mixgrad_out;  // nullable
  double n_total, k0, v0, S0_halflogdet, alpha;
  int prior, K, KP, DP, D;
  const double *dots;  // loop mode
  const double *beta;
  const int *info;
  gpc_loop_t *loop;    // nullable
  int phase;
};

__device__ inline double lngammad_dev(double v, int D) {  // np_utilities.py:62-64
  double s = 0.0;
  for (int d = 1; d <= D; ++d) s += lgamma((v + 1.0 - d) / 2.0);
  return s;
}

__global__ void __launch_bounds__(kClusterThreads) k_mog_finalize(const MogFinalizeArgs a) {
  if (loop_skip(a.loop, a.phase)) return;
  __shared__ double mixgrad[kMaxMixK];
  __shared__ MixScratch scratch;
  __shared__ double mix_s;
  const double *phi_hat = a.stats + (size_t)a.KP * a.DP;
  mixing_terms(phi_hat, a.K, a.alpha, a.prior, scratch, mixgrad, &mix_s);
  if (threadIdx.x == 0) {
    const double H = phi_hat[a.KP];
    const double mix = mix_s;
    double t1 = 0.0, t2 = 0.0, t3 = 0.0;
    for (int k = 0; k < a.K; ++k) {
      const double kN = a.per_k[k * 4 + 0], vN = a.per_k[k * 4 + 1], hld = a.per_k[k * 4 + 2];
      t1 += log(kN / a.k0);
      t2 += vN * hld;
      t3 += lngammad_dev(vN, a.D);
    }
    // MOG.py:63-70
    const double bound = -0.5 * a.D * t1 + a.K * a.v0 * a.S0_halflogdet - t2 + t3 - a.K * lngammad_dev(a.v0, a.D) + mix + H -
                         0.5 * a.n_total * a.D * log(M_PI);
    a.scalars[0] = bound;
    a.scalars[1] = mix;
    a.scalars[2] = H;
  }
  __syncthreads();  // bias / mixgrad below are written before the loop state may flip (same thread order is not needed: next kernel)
  if (threadIdx.x == 0 && a.loop) cg_decide(a.loop, loop_phase(a.loop, a.phase), a.scalars[0], *a.info != 0, a.dots[0], a.beta[0]);
  __syncthreads();
  for (int k = threadIdx.x; k < a.KP; k += blockDim.x) a.bias[k] = (k < a.K) ? a.per_k[k * 4 + 3] + mixgrad[k] : 0.0;
  if (a.mixgrad_out)
    for (int k = threadIdx.x; k < a.K; k += blockDim.x) a.mixgrad_out[k] = mixgrad[k];
}

}  // namespace gpc
