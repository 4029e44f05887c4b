// gpclust_b200 -- OMGP (overlapping mixtures of GPs) device path: dense N x N work per component.
//
// OMGP.py:96-147 needs, for every component i,
//     B = K_i(X) + diag(variance / (phi_i + 1e-6)),  L = jitchol(B),  logdet B,  alpha = B^-1 Y,
//     tr(B^-1 Y Y^T) = sum(Y o alpha),  diag(B^-1)
// The matrix lives in HBM padded to a multiple of 64 rows/columns (pad = identity, which leaves every
// quantity above unchanged).  Factorisation and triangular inverse are right-looking blocked algorithms
// with 64 x 64 tiles: the diagonal tile is handled by one CTA in shared memory, every other tile update is a
// 64 x 64 x 64 FP64 DMMA product (mma.sync.m8n8k4.f64).  diag(B^-1) are the squared column norms of
// W = L^-1 and alpha = W^T (W Y), so nothing is ever solved against an N x N right-hand side
// (the reference's dpotrs(LB, YYT), OMGP.py:113, is an N^3 solve whose trace is all that is used).
#pragma once
#include "gpc_cluster_kernels.cuh"
#include "gpc_common.cuh"
#include "gpc_util_kernels.cuh"

namespace gpc {

constexpr int kOB = 64;        // tile edge
constexpr int kOStrideT = 68;  // shared-memory row stride of row-major fragments read along k (bank skew 4)
constexpr int kOStrideN = 72;  // shared-memory row stride of the non-transposed B operand (bank skew 8)
constexpr int kOmgpMaxDy = 8;
constexpr int kOmgpSlabRows = 512;  // rows per partial-sum slab of the column statistics

// ------------------------------------------------------------------------------------------------
// B (Npad x Npad) = K(X, X) + diag(1 / ((w + eps) / variance) + jitter); pad rows/columns = identity.
// OMGP.py:104-105 (bound), :132-135 (gradient).  w has stride w_stride (a column of the N x K phi).
__global__ void k_omgp_build(const double *__restrict__ X, int N, int Npad, int q, const KernSpec spec, const double *__restrict__ w,
                             int w_stride, double variance, double eps, double jitter, double *__restrict__ B) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)Npad * Npad) return;
  const int i = (int)(e / Npad), j = (int)(e % Npad);
  double v;
  if (i < N && j < N) {
    v = kern_value(spec, X + (size_t)i * q, X + (size_t)j * q, q, true, i == j);
    if (i == j) v += 1.0 / ((w[(size_t)i * w_stride] + eps) / variance) + jitter;
  } else {
    v = (i == j) ? 1.0 : 0.0;
  }
  B[e] = v;
}

// ------------------------------------------------------------------------------------------------
// 64 x 64 x 64 tile product on the FP64 tensor pipe.  256 threads; warp w owns rows 8w..8w+7.
//   C = (SUB ? C - P*op(Q) : P*op(Q)),  op(Q) = Q^T when QT (Q given as n x k rows), Q otherwise (k x n rows)
// P may alias C (both k-halves of P are consumed before anything is written).
// The operands are staged in two k-halves of 32 (36.9 KB of shared memory per CTA instead of 71.7 KB): five CTAs per SM instead of
// three, which is what overlaps one CTA's global loads / read-modify-write epilogue with the DMMAs of the others.
constexpr int kOKH = 32;          // k columns staged at a time
constexpr int kOStrideH = 36;     // row stride of a 32-wide half read along k (bank skew 4)
// nk > 1: C -= sum over nk consecutive 64-wide k tiles (P advances by 64 columns per tile; Q by 64 columns when QT, by 64 rows otherwise):
// ONE read-modify-write of C for nk * 64 of k -- the arithmetic intensity of the blocked updates grows with nk (section 6 of DESIGN.md).
template <bool QT, bool SUB>
__device__ __forceinline__ void tile_product_64(const double *__restrict__ P, int ldp, const double *__restrict__ Q, int ldq, double *C, int ldc,
                                                double *Ps, double *Qs, int nk = 1) {
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  // warp w owns the 16 x 32 block (rows 16 (w/2) .., columns 32 (w%2) ..): per k-step 2 A fragments + 4 B fragments feed 8 DMMAs
  // (an 8 x 64 strip per warp needs 1 + 8)
  const int r0 = (warp >> 1) * 16, c0 = (warp & 1) * 32;
  double acc[2][4][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
  const double *pa = Ps + (r0 + g) * kOStrideH + t;
#pragma unroll 1
  for (int kh = 0; kh < nk * (kOB / kOKH); ++kh) {
    __syncthreads();  // previous user of the staging buffers is done
    for (int e = tid; e < kOB * kOKH / 2; e += blockDim.x) {  // P rows x 32 columns of this half
      const int r = e >> 4, c2 = (e & 15) * 2;
      *reinterpret_cast<double2 *>(Ps + r * kOStrideH + c2) = ldg2(P + (size_t)r * ldp + kh * kOKH + c2);
    }
    if (QT) {
      for (int e = tid; e < kOB * kOKH / 2; e += blockDim.x) {  // Q given as n x k rows: 64 rows x 32 columns
        const int r = e >> 4, c2 = (e & 15) * 2;
        *reinterpret_cast<double2 *>(Qs + r * kOStrideH + c2) = ldg2(Q + (size_t)r * ldq + kh * kOKH + c2);
      }
    } else {
      for (int e = tid; e < kOKH * kOB / 2; e += blockDim.x) {  // Q given as k x n rows: 32 rows x 64 columns
        const int r = e >> 5, c2 = (e & 31) * 2;
        *reinterpret_cast<double2 *>(Qs + r * kOStrideN + c2) = ldg2(Q + (size_t)(kh * kOKH + r) * ldq + c2);
      }
    }
    __syncthreads();
#pragma unroll 4
    for (int ks = 0; ks < kOKH / 4; ++ks) {
      const double af0 = pa[4 * ks], af1 = pa[8 * kOStrideH + 4 * ks];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const double bf = QT ? Qs[(c0 + 8 * nt + g) * kOStrideH + 4 * ks + t] : Qs[(4 * ks + t) * kOStrideN + c0 + 8 * nt + g];
        dmma884(acc[0][nt][0], acc[0][nt][1], af0, bf);
        dmma884(acc[1][nt][0], acc[1][nt][1], af1, bf);
      }
    }
  }
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    double *crow = C + (size_t)(r0 + 8 * mt + g) * ldc + c0 + 2 * t;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      if (SUB) {
        const double2 c = *reinterpret_cast<const double2 *>(crow + 8 * nt);
        stg2(crow + 8 * nt, c.x - acc[mt][nt][0], c.y - acc[mt][nt][1]);
      } else {
        stg2(crow + 8 * nt, acc[mt][nt][0], acc[mt][nt][1]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Diagonal tile jb: in-place lower Cholesky (strict upper of the tile zeroed), inverse of the factor to
// dinv[jb] (64 x 64 row-major), sum of log diag to ldsum[jb].  One CTA of 256 threads.
__global__ void __launch_bounds__(kClusterThreads) k_omgp_chol_diag(double *A, int ld, int jb, double *__restrict__ dinv, double *__restrict__ ldsum,
                                                                    int *info) {
  extern __shared__ __align__(16) double sm[];
  double *Ls = sm, *Li = sm + kOB * (kOB + 1);
  __shared__ int flag_s;
  const int tid = threadIdx.x, lds = kOB + 1;
  double *T = A + (size_t)jb * kOB * ld + (size_t)jb * kOB;
  for (int e = tid; e < kOB * kOB; e += blockDim.x) Ls[(e >> 6) * lds + (e & 63)] = T[(size_t)(e >> 6) * ld + (e & 63)];
  if (tid == 0) flag_s = 0;
  __syncthreads();
  const bool ok = chol_lower_quadrows(Ls, kOB, lds, &flag_s);
  if (!ok) {
    if (tid == 0) atomicOr(info, GPC_INFO_NOT_PD);
    // leave an identity behind so that the remaining steps stay finite; the host re-runs with jitter
    for (int e = tid; e < kOB * kOB; e += blockDim.x) {
      const int i = e >> 6, j = e & 63;
      T[(size_t)i * ld + j] = (i == j) ? 1.0 : 0.0;
      dinv[(size_t)jb * kOB * kOB + e] = (i == j) ? 1.0 : 0.0;
    }
    if (tid == 0) ldsum[jb] = 0.0;
    return;
  }
  __syncthreads();
  if (tid < 32) {
    double s = 0.0;
    for (int i = tid; i < kOB; i += 32) s += log(Ls[i * lds + i]);
    s = warp_sum(s);
    if (tid == 0) ldsum[jb] = s;
  }
  // inverse of the factor: quad r = tid/4 solves column r of L X = I by forward substitution, the column held
  // in registers split over q = tid%4 (same scheme as the triangular solve of k_mohgp_cluster)
  {
    const int r = tid >> 2, q = tid & 3;
    double Xr[kOB / 4];
#pragma unroll
    for (int mm = 0; mm < kOB / 4; ++mm) Xr[mm] = 0.0;
    for (int i = 0; i < kOB; ++i) {
      double pp[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int mm = 0; mm < kOB / 4; ++mm) {
        const int m = 4 * mm + q;
        pp[mm & 3] += (m < i) ? Ls[i * lds + m] * Xr[mm] : 0.0;
      }
      const double part = quad_sum((pp[0] + pp[1]) + (pp[2] + pp[3]));
      const double rhs = (i == r) ? 1.0 : 0.0;
      const double val = (i >= r) ? (rhs - part) / Ls[i * lds + i] : 0.0;
      const bool mine = (q == (i & 3));
#pragma unroll
      for (int mm = 0; mm < kOB / 4; ++mm) Xr[mm] = (mine && mm == (i >> 2)) ? val : Xr[mm];
    }
#pragma unroll
    for (int mm = 0; mm < kOB / 4; ++mm) Li[(4 * mm + q) * lds + r] = Xr[mm];
  }
  __syncthreads();
  for (int e = tid; e < kOB * kOB; e += blockDim.x) {
    const int i = e >> 6, j = e & 63;
    T[(size_t)i * ld + j] = (j <= i) ? Ls[i * lds + j] : 0.0;
    dinv[(size_t)jb * kOB * kOB + e] = (j <= i) ? Li[i * lds + j] : 0.0;
  }
}

// Panel below diagonal tile jb: A[i][jb] <- A[i][jb] * Linv_jb^T   (i = jb+1+blockIdx.x)
__global__ void __launch_bounds__(256) k_omgp_chol_panel(double *A, int ld, int jb, const double *__restrict__ dinv) {
  extern __shared__ __align__(16) double sm[];
  const int ib = jb + 1 + blockIdx.x;
  double *T = A + (size_t)ib * kOB * ld + (size_t)jb * kOB;
  tile_product_64<true, false>(T, ld, dinv + (size_t)jb * kOB * kOB, kOB, T, ld, sm, sm + kOB * kOStrideH);
}

// ---- two-level blocking: block columns are factorised in GROUPS of kOGroup tiles.  Inside a group the right-looking update touches
// only the group's own remaining columns (k = 64, thin); everything to the right of the group is updated ONCE per group with
// k = 64 * (group width).  The 64-wide update reads two 32 KB panels and read-modify-writes a 32 KB tile per 0.5 MFLOP (4 flop/B: bound
// by HBM at ~26 TFLOP/s, measured 20-22); with four tiles of k per pass over C it is 6.5 flop/B -- above the DMMA ridge.
constexpr int kOGroup = 4;
// within the group: A[ib][kb] -= A[ib][jb] A[kb][jb]^T for jb < kb < p1, ib >= kb   (grid: x = ib - jb - 1, y = kb - jb - 1)
__global__ void __launch_bounds__(256, 4) k_omgp_chol_within(double *A, int ld, int jb) {
  extern __shared__ __align__(16) double sm[];
  const int ib = jb + 1 + blockIdx.x, kb = jb + 1 + blockIdx.y;
  if (ib < kb) return;
  tile_product_64<true, true>(A + (size_t)ib * kOB * ld + (size_t)jb * kOB, ld, A + (size_t)kb * kOB * ld + (size_t)jb * kOB, ld,
                              A + (size_t)ib * kOB * ld + (size_t)kb * kOB, ld, sm, sm + kOB * kOStrideH);
}
// right of the group [p0, p1): A[ib][kb] -= sum_{t in group} A[ib][t] A[kb][t]^T for p1 <= kb <= ib   (triangular tile index)
__global__ void __launch_bounds__(256, 4) k_omgp_chol_big(double *A, int ld, int p0, int p1) {
  extern __shared__ __align__(16) double sm[];
  int a = (int)((sqrt(8.0 * (double)blockIdx.x + 1.0) - 1.0) * 0.5);
  while ((a + 1) * (a + 2) / 2 <= (int)blockIdx.x) ++a;
  while (a * (a + 1) / 2 > (int)blockIdx.x) --a;
  const int b = blockIdx.x - a * (a + 1) / 2;
  const int ib = p1 + a, kb = p1 + b;
  tile_product_64<true, true>(A + (size_t)ib * kOB * ld + (size_t)p0 * kOB, ld, A + (size_t)kb * kOB * ld + (size_t)p0 * kOB, ld,
                              A + (size_t)ib * kOB * ld + (size_t)kb * kOB, ld, sm, sm + kOB * kOStrideH, p1 - p0);
}

// ------------------------------------------------------------------------------------------------
// W = L^-1 by right-looking blocked forward substitution on the identity.
// init: W = 0, diagonal tiles W[i][i] = Linv_i.
__global__ void k_omgp_inv_init(double *__restrict__ W, int Npad, const double *__restrict__ dinv) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)Npad * Npad) return;
  const int i = (int)(e / Npad), j = (int)(e % Npad);
  W[e] = ((i >> 6) == (j >> 6)) ? dinv[(size_t)(i >> 6) * kOB * kOB + (i & 63) * kOB + (j & 63)] : 0.0;
}
// the same initialisation split for the pipelined variant (gpc_chol_inverse): all zeros first, diagonal tile ib at step ib
__global__ void k_omgp_inv_zero(double *__restrict__ W, long long total) {
  const long long e = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (e < total) stg2(W + e, 0.0, 0.0);
}
__global__ void __launch_bounds__(256) k_omgp_inv_diag(double *__restrict__ W, int ld, int ib, const double *__restrict__ dinv) {
  for (int e = threadIdx.x; e < kOB * kOB; e += blockDim.x)
    W[(size_t)(ib * kOB + (e >> 6)) * ld + ib * kOB + (e & 63)] = dinv[(size_t)ib * kOB * kOB + e];
}
// step ib, part (a): W[ib][j] <- Linv_ib * W[ib][j]  for j < ib   (blockIdx.x = j)
__global__ void __launch_bounds__(256) k_omgp_inv_row(double *W, int ld, int ib, const double *__restrict__ dinv) {
  extern __shared__ __align__(16) double sm[];
  double *T = W + (size_t)ib * kOB * ld + (size_t)blockIdx.x * kOB;
  tile_product_64<false, false>(dinv + (size_t)ib * kOB * kOB, kOB, T, ld, T, ld, sm, sm + kOB * kOStrideH);
}
// two-level blocking of the inverse (as for the factorisation): inside the group [p0, p1) step ib pushes row ib only to the rows of
// the group below it; the rows below the group receive the whole group at once with k = 64 * (group width).
//   within: W[m][j] -= L[m][ib] W[ib][j]  for ib < m < p1, j <= ib        (grid: x = j, y = m - ib - 1)
__global__ void __launch_bounds__(256, 4) k_omgp_inv_within(double *W, int ld, const double *__restrict__ L, int ldl, int ib) {
  extern __shared__ __align__(16) double sm[];
  const int jb = blockIdx.x, mb = ib + 1 + blockIdx.y;
  tile_product_64<false, true>(L + (size_t)mb * kOB * ldl + (size_t)ib * kOB, ldl, W + (size_t)ib * kOB * ld + (size_t)jb * kOB, ld,
                               W + (size_t)mb * kOB * ld + (size_t)jb * kOB, ld, sm, sm + kOB * kOStrideH);
}
//   big:    W[m][j] -= sum_{t in group} L[m][t] W[t][j]  for m >= p1, j < p1   (grid: x = j, y = m - p1); W[t][j] = 0 for t < j
__global__ void __launch_bounds__(256, 4) k_omgp_inv_big(double *W, int ld, const double *__restrict__ L, int ldl, int p0, int p1) {
  extern __shared__ __align__(16) double sm[];
  const int jb = blockIdx.x, mb = p1 + blockIdx.y;
  // columns right of the group's first tile: the rows t < j of the group hold zeros -- start the k loop at max(p0, j)
  const int t0 = jb > p0 ? jb : p0;
  tile_product_64<false, true>(L + (size_t)mb * kOB * ldl + (size_t)t0 * kOB, ldl, W + (size_t)t0 * kOB * ld + (size_t)jb * kOB, ld,
                               W + (size_t)mb * kOB * ld + (size_t)jb * kOB, ld, sm, sm + kOB * kOStrideH, p1 - t0);
}

// ------------------------------------------------------------------------------------------------
// t = W Y  (W lower triangular Npad x Npad, Y Npad x Dy with zero pad rows): one warp per row.
__global__ void __launch_bounds__(256) k_omgp_wy(const double *__restrict__ W, int Npad, const double *__restrict__ Y, int Dy, double *__restrict__ t) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x * 8 + warp;
  if (m >= Npad) return;
  double acc[kOmgpMaxDy];
#pragma unroll
  for (int d = 0; d < kOmgpMaxDy; ++d) acc[d] = 0.0;
  const double *wr = W + (size_t)m * Npad;
  for (int j = lane; j <= m; j += 32) {
    const double w = wr[j];
#pragma unroll
    for (int d = 0; d < kOmgpMaxDy; ++d)
      if (d < Dy) acc[d] += w * Y[(size_t)j * Dy + d];
  }
#pragma unroll
  for (int d = 0; d < kOmgpMaxDy; ++d)
    if (d < Dy) {
      const double v = warp_sum(acc[d]);
      if (lane == 0) t[(size_t)m * Dy + d] = v;
    }
}

// Column statistics of W over a slab of rows: part[s][j][0] = sum_m W[m][j]^2, part[s][j][1+d] = sum_m W[m][j] t[m][d]
// grid: x = column block of 64, y = slab; block 256 = 64 columns x 4 row lanes.
__global__ void __launch_bounds__(256) k_omgp_colstats(const double *__restrict__ W, int Npad, const double *__restrict__ t, int Dy, int slab_rows,
                                                       double *__restrict__ part) {
  __shared__ double red[4][64][kOmgpMaxDy + 1];
  const int c = threadIdx.x & 63, rl = threadIdx.x >> 6;
  const int j = blockIdx.x * 64 + c;
  const int r0 = blockIdx.y * slab_rows, r1 = min(Npad, r0 + slab_rows);
  double acc[kOmgpMaxDy + 1];
#pragma unroll
  for (int d = 0; d <= kOmgpMaxDy; ++d) acc[d] = 0.0;
  const int start = max(r0, blockIdx.x * 64);  // rows above the column block hold zeros (W is lower triangular)
  for (int m = start + rl; m < r1; m += 4) {
    const double w = W[(size_t)m * Npad + j];
    acc[0] += w * w;
#pragma unroll
    for (int d = 0; d < kOmgpMaxDy; ++d)
      if (d < Dy) acc[1 + d] += w * t[(size_t)m * Dy + d];
  }
#pragma unroll
  for (int d = 0; d <= kOmgpMaxDy; ++d) red[rl][c][d] = acc[d];
  __syncthreads();
  if (rl == 0) {
    double *o = part + ((size_t)blockIdx.y * Npad + j) * (kOmgpMaxDy + 1);
#pragma unroll
    for (int d = 0; d <= kOmgpMaxDy; ++d) o[d] = (red[0][c][d] + red[1][c][d]) + (red[2][c][d] + red[3][c][d]);
  }
}

// Per-component wrap-up (one CTA): diag(B^-1), alpha, grad_Lm column, and the scalars
//   comp[0] = tr(B^-1 Y Y^T) = sum(Y o alpha)   comp[1] = logdet B   comp[2] = sum_j phi_ji   (OMGP.py:113,117,121)
//   grad_Lm[j] = -0.5 variance (sum_d alpha_jd^2 - diag(B^-1)_j) / (phi_ji^2 + 1e-6)             (OMGP.py:138-140)
struct OmgpCompArgs {
  const double *part;   // nslab x Npad x (kOmgpMaxDy+1)
  const double *Y;      // Npad x Dy
  const double *phi;    // column i of the N x K responsibilities (stride phi_stride)
  const double *ldsum;  // Npad/64
  double *grad_col;     // column i of grad_Lm (stride grad_stride)
  double *comp;         // 4 doubles
  double variance;
  int N, Npad, Dy, nslab, phi_stride, grad_stride;
};
__global__ void __launch_bounds__(1024) k_omgp_component(const OmgpCompArgs a) {
  __shared__ double red_s[32];
  double tr = 0.0, ps = 0.0;
  for (int j = threadIdx.x; j < a.N; j += blockDim.x) {
    double v[kOmgpMaxDy + 1];
#pragma unroll
    for (int d = 0; d <= kOmgpMaxDy; ++d) v[d] = 0.0;
    for (int s = 0; s < a.nslab; ++s) {
      const double *p = a.part + ((size_t)s * a.Npad + j) * (kOmgpMaxDy + 1);
#pragma unroll
      for (int d = 0; d <= kOmgpMaxDy; ++d) v[d] += p[d];
    }
    double a2 = 0.0;
#pragma unroll
    for (int d = 0; d < kOmgpMaxDy; ++d)
      if (d < a.Dy) {
        a2 += v[1 + d] * v[1 + d];
        tr += a.Y[(size_t)j * a.Dy + d] * v[1 + d];
      }
    const double ph = a.phi[(size_t)j * a.phi_stride];
    ps += ph;
    a.grad_col[(size_t)j * a.grad_stride] = -0.5 * a.variance * (a2 - v[0]) / (ph * ph + 1e-6);
  }
  tr = block_sum(tr, red_s);
  ps = block_sum(ps, red_s);
  double ld = 0.0;
  for (int b = threadIdx.x; b < a.Npad / kOB; b += blockDim.x) ld += a.ldsum[b];
  ld = block_sum(ld, red_s);
  if (threadIdx.x == 0) {
    a.comp[0] = tr;
    a.comp[1] = 2.0 * ld;
    a.comp[2] = ps;
    a.comp[3] = 0.0;
  }
}

// ------------------------------------------------------------------------------------------------
// Kernel hyper-parameter gradients (OMGP.update_kern_grads, OMGP.py:55-94).  For one component, with
//   B = K(X) + diag(variance / phi)   (no 1e-6 here -- quirk of OMGP.py:63),  W = L^-1,  alpha = B^-1 Y:
//   dL_dB = alpha alpha^T - B^-1                                              (OMGP.py:71-75)
// k_omgp_alpha turns the column statistics of W (k_omgp_colstats) into alpha and diag(dL_dB).
__global__ void __launch_bounds__(256) k_omgp_alpha(const double *__restrict__ part, int nslab, int N, int Npad, int Dy, double *__restrict__ alpha,
                                                    double *__restrict__ dldb_diag) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= Npad) return;
  double v[kOmgpMaxDy + 1];
#pragma unroll
  for (int d = 0; d <= kOmgpMaxDy; ++d) v[d] = 0.0;
  for (int s = 0; s < nslab; ++s) {
    const double *p = part + ((size_t)s * Npad + j) * (kOmgpMaxDy + 1);
#pragma unroll
    for (int d = 0; d <= kOmgpMaxDy; ++d) v[d] += p[d];
  }
  double a2 = 0.0;
#pragma unroll
  for (int d = 0; d < kOmgpMaxDy; ++d)
    if (d < Dy) {
      const double a = (j < N) ? v[1 + d] : 0.0;
      alpha[(size_t)j * Dy + d] = a;
      a2 += a * a;
    }
  dldb_diag[j] = (j < N) ? a2 - v[0] : 0.0;
}

// Tile (ib >= jb) of B^-1 = W^T W on the FP64 tensor pipe (k-loop over the row tiles kb >= ib of the lower-triangular W),
// contracted in the epilogue against the parameter derivatives of every covariance part (GPy Stationary /
// Bias / White update_gradients_full with dL_dK = 0.5 dL_dB -- the 0.5 and the 1/lengthscale are applied by the host):
//   part[cta][2p]   = sum_ij w_ij dL_dB_ij k_p(r_ij) / variance_p
//   part[cta][2p+1] = sum_ij w_ij dL_dB_ij dk_p/dr (r_ij) r_ij            (stationary parts; 0 otherwise)
// w_ij = 2 for off-diagonal tiles (the symmetric partner is not visited), 1 on diagonal tiles.  Nothing N x N is stored.
struct OmgpKgradArgs {
  const double *W;      // Npad x Npad, lower triangular
  const double *X;      // N x q
  const double *alpha;  // Npad x Dy (zero pad rows)
  double *part;         // gridDim.x x 16
  KernSpec spec;
  int N, Npad, q, Dy;
};
__global__ void __launch_bounds__(256) k_omgp_kgrad_tiles(const OmgpKgradArgs a) {
  extern __shared__ __align__(16) double sm[];
  double *Ps = sm, *Qs = sm + kOB * kOStrideN;
  __shared__ double red_s[32];
  int ta = (int)((sqrt(8.0 * (double)blockIdx.x + 1.0) - 1.0) * 0.5);
  while ((ta + 1) * (ta + 2) / 2 <= (int)blockIdx.x) ++ta;
  while (ta * (ta + 1) / 2 > (int)blockIdx.x) --ta;
  const int ib = ta, jb = blockIdx.x - ta * (ta + 1) / 2;
  const int nb = a.Npad / kOB, ld = a.Npad;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  double acc[8][2];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) acc[nt][0] = acc[nt][1] = 0.0;
  for (int kb = ib; kb < nb; ++kb) {
    const double *P = a.W + (size_t)kb * kOB * ld + (size_t)ib * kOB;
    const double *Q = a.W + (size_t)kb * kOB * ld + (size_t)jb * kOB;
    __syncthreads();
    for (int e = tid; e < kOB * kOB / 2; e += blockDim.x) {
      const int r = e >> 5, c2 = (e & 31) * 2;
      *reinterpret_cast<double2 *>(Ps + r * kOStrideN + c2) = ldg2(P + (size_t)r * ld + c2);
      *reinterpret_cast<double2 *>(Qs + r * kOStrideN + c2) = ldg2(Q + (size_t)r * ld + c2);
    }
    __syncthreads();
#pragma unroll 4
    for (int ks = 0; ks < kOB / 4; ++ks) {
      const double af = Ps[(4 * ks + t) * kOStrideN + warp * 8 + g];  // A[m][k] = P[k][m]
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) dmma884(acc[nt][0], acc[nt][1], af, Qs[(4 * ks + t) * kOStrideN + 8 * nt + g]);
    }
  }
  // epilogue: lane (g, t) of warp w holds (row 8w+g, cols 8nt+2t, +1) of the tile
  double sums[2 * kMaxKernParts];
#pragma unroll
  for (int p = 0; p < 2 * kMaxKernParts; ++p) sums[p] = 0.0;
  const int i = ib * kOB + warp * 8 + g;
  const double wgt = (ib == jb) ? 1.0 : 2.0;
  if (i < a.N) {
    double ai[kOmgpMaxDy];
#pragma unroll
    for (int d = 0; d < kOmgpMaxDy; ++d) ai[d] = (d < a.Dy) ? a.alpha[(size_t)i * a.Dy + d] : 0.0;
    const double *xi = a.X + (size_t)i * a.q;
    double si = 0.0;
    for (int c = 0; c < a.q; ++c) si += xi[c] * xi[c];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = jb * kOB + 8 * nt + 2 * t + e;
        if (j >= a.N) continue;
        double aa = 0.0;
#pragma unroll
        for (int d = 0; d < kOmgpMaxDy; ++d)
          if (d < a.Dy) aa += ai[d] * a.alpha[(size_t)j * a.Dy + d];
        const double dl = wgt * (aa - acc[nt][e]);
        const double *xj = a.X + (size_t)j * a.q;
        double dot = 0.0, sj = 0.0;
        for (int c = 0; c < a.q; ++c) {
          dot += xi[c] * xj[c];
          sj += xj[c] * xj[c];
        }
        double r2 = -2.0 * dot + (si + sj);  // GPy Stationary._unscaled_dist (see kern_value)
        if (i == j) r2 = 0.0;
        const double r_un = sqrt(fmax(r2, 0.0));
#pragma unroll
        for (int p = 0; p < kMaxKernParts; ++p) {
          if (p >= a.spec.nparts) continue;
          const double r = r_un / a.spec.lengthscale[p], var = a.spec.variance[p];
          double f, dr;  // k_p / variance, dk_p/dr * r
          switch (a.spec.kind[p]) {
            case GPC_KERN_RBF: f = exp(-0.5 * (r * r)); dr = -r * (var * f) * r; break;
            case GPC_KERN_MATERN32: { const double ex = exp(-sqrt(3.0) * r); f = (1.0 + sqrt(3.0) * r) * ex; dr = (-3.0 * var * r * ex) * r; } break;
            case GPC_KERN_MATERN52: {
              const double ex = exp(-sqrt(5.0) * r);
              f = (1.0 + sqrt(5.0) * r + 5.0 / 3.0 * (r * r)) * ex;
              dr = (var * (10.0 / 3.0 * r - 5.0 * r - 5.0 * sqrt(5.0) / 3.0 * (r * r)) * ex) * r;
            } break;
            case GPC_KERN_BIAS: f = 1.0; dr = 0.0; break;
            default: f = (i == j) ? 1.0 : 0.0; dr = 0.0; break;  // white: trace(dL_dK)
          }
          sums[2 * p] += dl * f;
          sums[2 * p + 1] += dl * dr;
        }
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int p = 0; p < 2 * kMaxKernParts; ++p) {
    const double v = block_sum(sums[p], red_s);
    if (tid == 0) a.part[(size_t)blockIdx.x * 2 * kMaxKernParts + p] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// Bound assembly (OMGP.py:96-123) + mixing-proportion gradient.  comp: K x 4 from k_omgp_component.
struct OmgpFinalizeArgs {
  const double *comp;
  const double *Hpart;  // softmax entropy partials
  int n_hpart;
  double *scalars;      // [0]=bound [1]=mix [2]=H
  double *phi_hat;      // out K
  double *mixgrad;      // out K
  double variance, alpha;
  int prior, K, Dy;
};
__global__ void __launch_bounds__(kClusterThreads) k_omgp_finalize(const OmgpFinalizeArgs a) {
  __shared__ double mixgrad[kMaxMixK];
  __shared__ double phi_hat[kMaxMixK];
  __shared__ MixScratch scratch;
  __shared__ double mix_s;
  for (int k = threadIdx.x; k < a.K; k += blockDim.x) phi_hat[k] = a.comp[k * 4 + 2];
  __syncthreads();
  mixing_terms(phi_hat, a.K, a.alpha, a.prior, scratch, mixgrad, &mix_s);
  if (threadIdx.x == 0) {
    double H = 0.0;
    for (int c = 0; c < a.n_hpart; ++c) H += a.Hpart[c];
    double gp = 0.0;
    for (int k = 0; k < a.K; ++k) {
      gp -= 0.5 * a.comp[k * 4 + 0];
      gp -= 0.5 * a.comp[k * 4 + 1];
      gp -= 0.5 * a.Dy * (a.comp[k * 4 + 2] * log(2.0 * M_PI * a.variance));
    }
    a.scalars[0] = gp + mix_s + H;
    a.scalars[1] = mix_s;
    a.scalars[2] = H;
  }
  for (int k = threadIdx.x; k < a.K; k += blockDim.x) {
    a.phi_hat[k] = phi_hat[k];
    a.mixgrad[k] = mixgrad[k];
  }
}

// ------------------------------------------------------------------------------------------------
// Natural gradient for models whose per-point likelihood gradient is materialised (OMGP.py:142-147):
//   grad_phi = grad_Lm + mixgrad + Hgrad ; natgrad = grad_phi - sum_k phi grad_phi ; grad = natgrad * phi
// plus the CG dots of collapsed_vb.py:154,160-164 against the previous (ng_old, grad_old), nullable.
// One thread per row; per-CTA partial dots in partials[blockIdx.x*4 + {0,1,2}].
__global__ void __launch_bounds__(128) k_dense_natgrad(const double *__restrict__ grad_Lm, const double *__restrict__ mixgrad, const double *__restrict__ phi,
                                                       const double *__restrict__ logphi, const double *__restrict__ ng_old,
                                                       const double *__restrict__ grad_old, int N, int K, double *__restrict__ ng_out,
                                                       double *__restrict__ grad_out, double *__restrict__ partials) {
  __shared__ double red_s[32];
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  double d0 = 0.0, d1 = 0.0, d2 = 0.0;
  if (n < N) {
    const size_t o = (size_t)n * K;
    double s = 0.0;
    for (int k = 0; k < K; ++k) s += phi[o + k] * ((grad_Lm[o + k] + mixgrad[k]) - logphi[o + k]);
    for (int k = 0; k < K; ++k) {
      const double gp = (grad_Lm[o + k] + mixgrad[k]) - logphi[o + k];
      const double ng = gp - s, gr = ng * phi[o + k];
      ng_out[o + k] = ng;
      grad_out[o + k] = gr;
      d0 += ng * gr;
      if (ng_old) {
        const double e = ng - ng_old[o + k];
        d1 += e * gr;
        d2 += e * grad_old[o + k];
      }
    }
  }
  d0 = block_sum(d0, red_s);
  d1 = block_sum(d1, red_s);
  d2 = block_sum(d2, red_s);
  if (threadIdx.x == 0) {
    partials[(size_t)blockIdx.x * 4 + 0] = d0;
    partials[(size_t)blockIdx.x * 4 + 1] = d1;
    partials[(size_t)blockIdx.x * 4 + 2] = d2;
    partials[(size_t)blockIdx.x * 4 + 3] = 0.0;
  }
}

// sd_new = ng + beta*sd_old (sd_old nullable / beta == 0) ; x_new = x + step*sd_new    (collapsed_vb.py:167,172)
__global__ void k_cg_axpy(const double *__restrict__ x, const double *__restrict__ ng, const double *sd_old, double beta, double step,
                          long long len, double *sd_new, double *__restrict__ x_new) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= len) return;
  double s = ng[e];
  if (sd_old != nullptr && beta != 0.0) s += beta * sd_old[e];
  sd_new[e] = s;
  x_new[e] = x[e] + step * s;
}

}  // namespace gpc
