// gpclust_b200 -- setup / read-out kernels around the two N-sized passes (FP64, sm_100a).
//
//   k_kern_fill        : GPy covariance functions .K(X, X2)  (call sites MOHGP.py:47-48,61-62; OMGP.py:104,132)
//   k_mog_features     : MOG feature rows [x_i x_j | x_i] that turn MOG.py:56-57,74-77 into the same
//                        two GEMMs as MOHGP
//   k_softmax_rows     : np_utilities.py:25-40 / utilities.py:61-102 -- materialises phi, log phi, H
//   k_mahalanobis_pairs: utilities.py:147-170, per-pair forward substitution, same operation order
//   k_gram_partials    : Y^T Y of MOHGP.py:53 (per-CTA partials, fixed-order reduction afterwards)
//   k_mohgp_setup      : pdinv(Sy + 1e-6 I) and backsub_both_sides(Ly, Sf) of MOHGP.py:49,63,79
//   k_batched_pdinv    : np_utilities.py:6-22 (multiple_pdinv)
#pragma once
#include "gpc_cluster_kernels.cuh"
#include "gpc_common.cuh"

namespace gpc {

// ------------------------------------------------------------------------------------------------
// GPC_KERN_* come from include/gpclust_b200.h
constexpr int kMaxKernParts = 8;
struct KernSpec {
  int nparts;
  int kind[kMaxKernParts];
  double variance[kMaxKernParts];
  double lengthscale[kMaxKernParts];
};

// sum_parts k_p(x_i, x_j).  r^2 is formed the way GPy's Stationary._unscaled_dist does:
// -2 x.x2 + |x|^2 + |x2|^2, diagonal forced to zero when X2 is absent (`self`), clipped at zero.
// `same` = (self && i == j): White contributes only there.
__device__ __forceinline__ double kern_value(const KernSpec &spec, const double *xi, const double *xj, int q, bool self, bool same) {
  double dot = 0.0, si = 0.0, sj = 0.0;
  for (int c = 0; c < q; ++c) {
    dot += xi[c] * xj[c];
    si += xi[c] * xi[c];
    sj += xj[c] * xj[c];
  }
  double r2 = -2.0 * dot + (si + sj);
  if (same) r2 = 0.0;
  r2 = fmax(r2, 0.0);
  const double r_un = sqrt(r2);
  double v = 0.0;
  for (int p = 0; p < spec.nparts; ++p) {
    const double var = spec.variance[p];
    const double r = r_un / spec.lengthscale[p];
    double kp;
    switch (spec.kind[p]) {
      case GPC_KERN_RBF: kp = var * exp(-0.5 * (r * r)); break;
      case GPC_KERN_MATERN32: kp = var * (1.0 + sqrt(3.0) * r) * exp(-sqrt(3.0) * r); break;
      case GPC_KERN_MATERN52: kp = var * (1.0 + sqrt(5.0) * r + 5.0 / 3.0 * (r * r)) * exp(-sqrt(5.0) * r); break;
      case GPC_KERN_BIAS: kp = var; break;
      default: kp = same ? var : 0.0; break;  // white
    }
    v = (p == 0) ? kp : v + kp;
  }
  return v;
}

// out[i][j] = sum_parts k_p(x_i, x2_j)  (+ diag_add on the diagonal when X2 is absent)
__global__ void k_kern_fill(const double *__restrict__ X, const double *__restrict__ X2, int n, int m, int q, const KernSpec spec,
                            double diag_add, double *__restrict__ out) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)n * m) return;
  const int i = (int)(e / m), j = (int)(e % m);
  const bool self = (X2 == nullptr);
  double v = kern_value(spec, X + (size_t)i * q, (self ? X : X2) + (size_t)j * q, q, self, self && i == j);
  if (self && i == j) v += diag_add;
  out[e] = v;
}

// ------------------------------------------------------------------------------------------------
// Offsets inside one 8-row group of the fragment-major HBM layouts (gpc_vb_kernels.cuh header)
__host__ __device__ __forceinline__ int frag_off_features(int r, int d) { return (d >> 2) * 32 + r * 4 + (d & 3); }
__host__ __device__ __forceinline__ int frag_off_logits(int r, int c) { return (((c >> 3) * 32 + r * 4 + ((c & 7) >> 1)) << 1) + (c & 1); }

// MOG features: row n -> [ xc_i*xc_j (i<=j, row-major upper) | xc_i | 0 pad ], xc = x - centre; written in the
// fragment-major feature layout (DP columns per row).
__global__ void k_mog_features(const double *__restrict__ X, const double *__restrict__ centre, int N, int Npad, int D, int DP,
                               double *__restrict__ F) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)Npad * DP) return;
  const int n = (int)(e / DP), f = (int)(e % DP);
  const int NQ = D * (D + 1) / 2;
  double v = 0.0;
  if (n < N) {
    if (f < NQ) {
      int i = 0, rem = f;
      while (rem >= D - i) {
        rem -= D - i;
        ++i;
      }
      const int j = i + rem;
      v = (X[(size_t)n * D + i] - centre[i]) * (X[(size_t)n * D + j] - centre[j]);
    } else if (f < NQ + D) {
      const int i = f - NQ;
      v = X[(size_t)n * D + i] - centre[i];
    }
  }
  F[(size_t)(n >> 3) * 8 * DP + frag_off_features(n & 7, f)] = v;
}

// contiguous N x C row-major  ->  fragment-major (Npad rows, CP columns, pad = 0); LOGITS selects the C-fragment order
template <bool LOGITS>
__global__ void k_pack_rows(const double *__restrict__ src, int N, int C, int Npad, int CP, double *__restrict__ dst) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // index into dst
  if (e >= (long long)Npad * CP) return;
  const int grp = (int)(e / (8 * CP)), o = (int)(e % (8 * CP));
  int r, c;
  if (LOGITS) {  // o = ((c/8)*32 + r*4 + (c%8)/2)*2 + c%2
    const int h = o >> 1;
    c = ((h >> 5) << 3) + ((h & 3) << 1) + (o & 1);
    r = (h >> 2) & 7;
  } else {  // o = (d/4)*32 + r*4 + d%4
    c = ((o >> 5) << 2) + (o & 3);
    r = (o >> 2) & 7;
  }
  const int n = grp * 8 + r;
  dst[e] = (n < N && c < C) ? src[(size_t)n * C + c] : 0.0;
}
// and back: fragment-major -> contiguous N x C
template <bool LOGITS>
__global__ void k_unpack_rows(const double *__restrict__ src, int N, int C, int CP, double *__restrict__ dst) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)N * C) return;
  const int n = (int)(e / C), c = (int)(e % C);
  const int o = LOGITS ? frag_off_logits(n & 7, c) : frag_off_features(n & 7, c);
  dst[e] = src[(size_t)(n >> 3) * 8 * CP + o];
}

// Copy an N x D matrix into the padded Npad x DP feature layout (pad = 0).
__global__ void k_pad_rows(const double *__restrict__ src, int N, int D, int Npad, int DP, double *__restrict__ dst) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)Npad * DP) return;
  const int n = (int)(e / DP), d = (int)(e % DP);
  dst[e] = (n < N && d < D) ? src[(size_t)n * D + d] : 0.0;
}
// and back: padded Npad x ld -> contiguous N x K
__global__ void k_unpad_rows(const double *__restrict__ src, int N, int K, int ld, double *__restrict__ dst) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)N * K) return;
  const int n = (int)(e / K), k = (int)(e % K);
  dst[e] = src[(size_t)n * ld + k];
}

// ------------------------------------------------------------------------------------------------
// Ingest (the step before the path; drosophila.ipynb cells [2] and [4]).
// Row-wise normalisation  y <- (y - mean(y)) / std(y)   (population std, as numpy's .std(1)); one warp per row.
__global__ void __launch_bounds__(256) k_normalise_rows(const double *__restrict__ src, int N, int D, double *__restrict__ dst) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (long long row = (long long)blockIdx.x * 8 + warp; row < N; row += (long long)gridDim.x * 8) {
    const double *y = src + (size_t)row * D;
    double s = 0.0;
    for (int d = lane; d < D; d += 32) s += y[d];
    const double mean = warp_sum(s) / D;
    double v = 0.0;
    for (int d = lane; d < D; d += 32) v += (y[d] - mean) * (y[d] - mean);
    const double sd = sqrt(warp_sum(v) / D);
    for (int d = lane; d < D; d += 32) dst[(size_t)row * D + d] = (y[d] - mean) / sd;
  }
}
// Replicate score of a series: std over time points of the replicate means / mean over time points of the replicate stds
// (drosophila.ipynb cell [4]).  group[d] in [0, G) is the time-point index of column d; one thread per row, G <= 64.
__global__ void k_replicate_scores(const double *__restrict__ Y, int N, int D, const int *__restrict__ group, int G, double *__restrict__ score) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= N) return;
  const double *y = Y + (size_t)row * D;
  double sum[64], sq[64];
  int cnt[64];
  for (int q = 0; q < G; ++q) {
    sum[q] = sq[q] = 0.0;
    cnt[q] = 0;
  }
  for (int d = 0; d < D; ++d) {
    sum[group[d]] += y[d];
    cnt[group[d]] += 1;
  }
  for (int q = 0; q < G; ++q) sum[q] /= cnt[q];  // replicate means
  for (int d = 0; d < D; ++d) {
    const double e = y[d] - sum[group[d]];
    sq[group[d]] += e * e;
  }
  double mm = 0.0, ms = 0.0;
  for (int q = 0; q < G; ++q) {
    mm += sum[q];
    ms += sqrt(sq[q] / cnt[q]);
  }
  mm /= G;
  ms /= G;
  double vm = 0.0;
  for (int q = 0; q < G; ++q) vm += (sum[q] - mm) * (sum[q] - mm);
  score[row] = sqrt(vm / G) / ms;
}

// ------------------------------------------------------------------------------------------------
// Row softmax read-out: one warp per row.  x has leading dimension ldx; phi / logphi (nullable) are
// written with leading dimension ldo; Hpart[blockIdx.x] = sum over this CTA's rows of -sum_k phi log phi.
constexpr int kSoftmaxWarps = 8;
__global__ void __launch_bounds__(kSoftmaxWarps * 32) k_softmax_rows(const double *__restrict__ x, int N, int K, int ldx, double *__restrict__ phi,
                                                                     double *__restrict__ logphi, int ldo, double *__restrict__ Hpart) {
  __shared__ double red_s[kSoftmaxWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double hsum = 0.0;
  for (long long row = (long long)blockIdx.x * kSoftmaxWarps + warp; row < N; row += (long long)gridDim.x * kSoftmaxWarps) {
    const double *xr = x + (size_t)row * ldx;
    double m = -INFINITY;
    for (int k = lane; k < K; k += 32) m = fmax(m, xr[k]);
    m = warp_max(m);
    double s = 0.0;
    for (int k = lane; k < K; k += 32) s += exp(xr[k] - m);
    s = warp_sum(s);
    const double lognorm = log(s);
    double h = 0.0;
    for (int k = lane; k < K; k += 32) {
      const double lp0 = xr[k] - m;
      const double p = exp(lp0) / s;
      const double lp = lp0 - lognorm;
      h -= p * lp;
      if (phi) phi[(size_t)row * ldo + k] = p;
      if (logphi) logphi[(size_t)row * ldo + k] = lp;
    }
    hsum += warp_sum(h);
  }
  if (lane == 0) red_s[warp] = hsum;
  __syncthreads();
  if (threadIdx.x == 0 && Hpart) {
    double v = 0.0;
    for (int w = 0; w < kSoftmaxWarps; ++w) v += red_s[w];
    Hpart[blockIdx.x] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// Pairwise Mahalanobis by forward substitution (the reference's own native kernel, utilities.py:147-170):
// out[n][m] = | L^-1 (x1_n - x2_m) |^2.  One thread per pair, L staged in shared memory.
constexpr int kMahaMaxD = 128;
__global__ void __launch_bounds__(128) k_mahalanobis_pairs(const double *__restrict__ X1, const double *__restrict__ X2, const double *__restrict__ L,
                                                           int N1, int N2, int D, double *__restrict__ out) {
  extern __shared__ __align__(16) double Ls[];  // D x D
  for (int e = threadIdx.x; e < D * D; e += blockDim.x) Ls[e] = L[e];
  __syncthreads();
  const long long pair = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pair >= (long long)N1 * N2) return;
  const int n = (int)(pair / N2), m = (int)(pair % N2);
  double t[kMahaMaxD];
  double acc2 = 0.0;
  for (int i = 0; i < D; ++i) {
    double v = X1[(size_t)n * D + i] - X2[(size_t)m * D + i];
    for (int j = 0; j < i; ++j) v -= Ls[i * D + j] * t[j];
    v /= Ls[i * D + i];
    t[i] = v;
  }
  for (int i = 0; i < D; ++i) acc2 += t[i] * t[i];
  out[pair] = acc2;
}

// ------------------------------------------------------------------------------------------------
// Gram matrix partials: part[c][i][j] = sum over CTA c's rows of F[n][i] F[n][j]   (i, j < DP); F fragment-major.
// On the FP64 tensor pipe: for one 8-row group the A fragment of F^T for (m-tile mt, k-step ks) and the B fragment of F for
// (n-tile dt, ks) are the SAME shared-memory word pattern  base[tile*64 + ks*16]  (gpc_vb_kernels.cuh, statistics warps),
// so a 64-row slab (8 groups, copied verbatim) feeds 2*8 k-steps; warp w owns the output tiles (mt = w, dt = 0..7).
constexpr int kGramRows = 64;
__global__ void __launch_bounds__(256, 2) k_gram_partials(const double *__restrict__ F, int Npad, int DP, double *__restrict__ partials) {
  extern __shared__ __align__(16) double slab[];  // kGramRows x DP, fragment-major (8 groups)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int DT = DP / 8;
  double acc[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = 0.0;
  const int num_slabs = Npad / kGramRows, slab_doubles = kGramRows * DP;
  const double *fb = slab + (g >> 2) * 32 + t * 4 + (g & 3);
  for (int sl = blockIdx.x; sl < num_slabs; sl += gridDim.x) {
    __syncthreads();
    const double2 *src = reinterpret_cast<const double2 *>(F + (size_t)sl * slab_doubles);
    double2 *dst = reinterpret_cast<double2 *>(slab);
    for (int e = tid; e < slab_doubles / 2; e += 256) dst[e] = src[e];
    __syncthreads();
    if (warp < DT) {
#pragma unroll 2
      for (int grp = 0; grp < kGramRows / 8; ++grp) {
        const double *gb = fb + grp * 8 * DP;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const double af = gb[warp * 64 + ks * 16];
#pragma unroll
          for (int dt = 0; dt < 8; ++dt)
            if (dt < DT) dmma884(acc[dt][0], acc[dt][1], af, gb[dt * 64 + ks * 16]);
        }
      }
    }
  }
  double *part = partials + (size_t)blockIdx.x * DP * DP;
  if (warp < DT) {
#pragma unroll
    for (int dt = 0; dt < 8; ++dt)
      if (dt < DT) stg2(part + (size_t)(8 * warp + g) * DP + 8 * dt + 2 * t, acc[dt][0], acc[dt][1]);
  }
}

// ------------------------------------------------------------------------------------------------
// Lower-triangular inverse in shared memory: Li = L^-1, one column per thread (column-oriented).
__device__ inline void tri_inverse_lower_smem(const double *L, int n, int ld, double *Li) {
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    for (int i = 0; i < j; ++i) Li[i * ld + j] = 0.0;
    Li[j * ld + j] = 1.0 / L[j * ld + j];
    for (int i = j + 1; i < n; ++i) {
      double v = 0.0;
      for (int m = j; m < i; ++m) v -= L[i * ld + m] * Li[m * ld + j];
      Li[i * ld + j] = v / L[i * ld + i];
    }
  }
  __syncthreads();
}

// jitchol (GPy.util.linalg.jitchol) of the n x n matrix `src` (global, row-major, + diag_add on the
// diagonal) into shared memory M.  Returns false when not positive definite even with jitter.
__device__ inline bool jitchol_smem(const double *src, double diag_add, int n, int ld, double *M, int *flag_s) {
  bool ok = false;
  double diag_mean = 0.0;
  for (int attempt = 0; attempt < 6 && !ok; ++attempt) {
    double jitter = 0.0;
    if (attempt > 0) {
      jitter = diag_mean * 1e-6;
      for (int q = 1; q < attempt; ++q) jitter *= 10.0;
      if (!isfinite(jitter)) break;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
      const int i = e / n, j = e % n;
      M[i * ld + j] = src[e] + ((i == j) ? diag_add + jitter : 0.0);
    }
    __syncthreads();
    if (attempt == 0) {
      double dsum = 0.0, dmin = INFINITY;
      for (int i = 0; i < n; ++i) {
        dsum += M[i * ld + i];
        dmin = fmin(dmin, M[i * ld + i]);
      }
      diag_mean = dsum / n;
      ok = chol_lower_smem(M, n, ld, flag_s);
      if (!ok && !(dmin > 0.0)) break;
    } else {
      ok = chol_lower_smem(M, n, ld, flag_s);
    }
  }
  return ok;
}

struct MohgpSetupArgs {
  const double *Sf;    // D x D
  const double *Sy;    // D x D (un-jittered)
  double *Ly;          // out D x D : chol(Sy + 1e-6 I), lower, strict upper zero   (Sy_chol)
  double *Ly_inv;      // out D x D : Ly^-1                                          (Sy_chol_inv)
  double *Sy_inv;      // out D x D : (Sy + 1e-6 I)^-1
  double *A;           // out D x D : Ly^-1 Sf Ly^-T                                 (MOHGP.py:79)
  double *Sy_logdet;   // out scalar
  int *info;
  int D;
};

__global__ void __launch_bounds__(kClusterThreads) k_mohgp_setup(const MohgpSetupArgs a) {
  extern __shared__ __align__(16) double sm[];
  const int D = a.D, ld = D + 1, tid = threadIdx.x;
  double *Lm = sm;            // factor
  double *Li = Lm + D * ld;   // its inverse
  double *W = Li + D * ld;    // work
  __shared__ int flag_s;
  if (!jitchol_smem(a.Sy, 1e-6, D, ld, Lm, &flag_s)) {
    if (tid == 0) atomicOr(a.info, GPC_INFO_NOT_PD);
    return;
  }
  if (tid < 32) {
    double ldd = 0.0;
    for (int i = tid; i < D; i += 32) ldd += log(Lm[i * ld + i]);
    ldd = 2.0 * warp_sum(ldd);
    if (tid == 0) *a.Sy_logdet = ldd;
  }
  __syncthreads();
  tri_inverse_lower_smem(Lm, D, ld, Li);
  for (int e = tid; e < D * D; e += blockDim.x) {
    const int i = e / D, j = e % D;
    a.Ly[e] = (j <= i) ? Lm[i * ld + j] : 0.0;
    a.Ly_inv[e] = (j <= i) ? Li[i * ld + j] : 0.0;
    // Sy_inv = Li^T Li
    double v = 0.0;
    for (int m = (i > j ? i : j); m < D; ++m) v += Li[m * ld + i] * Li[m * ld + j];
    a.Sy_inv[e] = v;
  }
  // W = Li Sf ; A = W Li^T
  for (int e = tid; e < D * D; e += blockDim.x) {
    const int i = e / D, j = e % D;
    double v = 0.0;
    for (int m = 0; m <= i; ++m) v += Li[i * ld + m] * a.Sf[m * D + j];
    W[i * ld + j] = v;
  }
  __syncthreads();
  for (int e = tid; e < D * D; e += blockDim.x) {
    const int i = e / D, j = e % D;
    double v = 0.0;
    for (int m = 0; m <= j; ++m) v += W[i * ld + m] * Li[j * ld + m];
    a.A[e] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// multiple_pdinv (np_utilities.py:6-22): A is D x D x K with K LAST (stride K between matrix entries);
// inv has the same layout; hld[k] = sum log diag chol.  One CTA per k.
__global__ void __launch_bounds__(kClusterThreads) k_batched_pdinv(const double *__restrict__ A, int D, int K, double *__restrict__ inv,
                                                                   double *__restrict__ hld, int *info) {
  extern __shared__ __align__(16) double sm[];
  const int ld = D + 1, k = blockIdx.x, tid = threadIdx.x;
  double *G = sm;            // gathered matrix (contiguous D x D)
  double *Lm = G + D * D;    // factor
  double *Li = Lm + D * ld;  // inverse factor
  __shared__ int flag_s;
  for (int e = tid; e < D * D; e += blockDim.x) G[e] = A[(size_t)e * K + k];
  __syncthreads();
  if (!jitchol_smem(G, 0.0, D, ld, Lm, &flag_s)) {
    if (tid == 0) atomicOr(info, GPC_INFO_NOT_PD);
    return;
  }
  if (tid < 32) {
    double s = 0.0;
    for (int i = tid; i < D; i += 32) s += log(Lm[i * ld + i]);
    s = warp_sum(s);
    if (tid == 0) hld[k] = s;
  }
  __syncthreads();
  tri_inverse_lower_smem(Lm, D, ld, Li);
  for (int e = tid; e < D * D; e += blockDim.x) {
    const int i = e / D, j = e % D;
    double v = 0.0;
    for (int m = (i > j ? i : j); m < D; ++m) v += Li[m * ld + i] * Li[m * ld + j];
    inv[(size_t)e * K + k] = v;
  }
}

}  // namespace gpc
