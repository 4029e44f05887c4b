// gpclust_b200 -- C-ABI launchers (see include/gpclust_b200.h for the contract and the reference
// lines each entry point replaces).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo
#include "../../include/gpclust_b200.h"

#include <cuda_runtime.h>
#include <math.h>

#include <map>
#include <mutex>
#include <utility>

#include "gpc_cluster_kernels.cuh"
#include "gpc_common.cuh"
#include "gpc_loop_kernel.cuh"
#include "gpc_omgp_kernels.cuh"
#include "gpc_util_kernels.cuh"
#include "gpc_vb_kernels.cuh"

using namespace gpc;

namespace {

inline int cuda_rc(cudaError_t e) { return e == cudaSuccess ? 0 : -(int)e; }
inline int launch_rc() { return cuda_rc(cudaGetLastError()); }
inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

// cudaFuncAttributeMaxDynamicSharedMemorySize is set ONCE per (device, kernel) -- raised only when a launch needs more than
// what is already granted -- instead of on every launch (pure host overhead on the host-driven loops: callbacks, free
// hyper-parameters, MOG multi-GPU, OMGP's thousands of tile launches).  A process-wide cache of opt-in sizes; not device state.
template <typename Kern>
cudaError_t ensure_dyn_smem(Kern kern, size_t smem) {
  static std::mutex mu;
  static std::map<std::pair<int, const void *>, size_t> granted;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const std::pair<int, const void *> key(dev, reinterpret_cast<const void *>(kern));
  std::lock_guard<std::mutex> lock(mu);
  auto it = granted.find(key);
  if (it != granted.end() && it->second >= smem) return cudaSuccess;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) granted[key] = smem;
  return e;
}

// ring depth: as many group stages as fit next to the fixed part in the 227 KB of one SM (capped)
constexpr int kMaxStages = 24;
constexpr int kMaxUnitSlots = 4;
int natgrad_stages(int KP, int DP) {
  // every unit of consumer warps owns the same number of ring slots (at most kMaxUnitSlots)
  const int units = kGConsumers / ((KP / 8 >= 2) ? 2 : 1);
  const size_t fixed = sizeof(double) * (size_t)g_fixed_doubles(KP, DP) + 64;
  const size_t per = sizeof(double) * (size_t)g_stage_doubles(KP, DP) + sizeof(uint64_t);
  int r = (int)((kSmemBudget - fixed) / per) / units;
  if (r > kMaxUnitSlots) r = kMaxUnitSlots;
  return r * units;
}
size_t natgrad_smem(int KP, int DP, int S) {
  return sizeof(double) * ((size_t)S * g_stage_doubles(KP, DP) + g_fixed_doubles(KP, DP)) + sizeof(uint64_t) * S;
}
int step_stats_stages(int KP, int DP) {
  const size_t fixed = sizeof(double) * kSSoftmaxWarps + 64;
  const size_t per = sizeof(double) * (size_t)s_stage_doubles(KP, DP) + 3 * sizeof(uint64_t);
  const int s = (int)((kSmemBudget - fixed) / per);
  return s > kMaxStages ? kMaxStages : s;
}
size_t step_stats_smem(int KP, int DP, int S) {
  return sizeof(double) * ((size_t)S * s_stage_doubles(KP, DP) + kSSoftmaxWarps) + sizeof(uint64_t) * 3 * S;
}

template <typename Kern, typename Args>
int launch_vb(Kern kern, const Args &a, size_t smem, int grid, int threads, cudaStream_t st) {
  cudaError_t e = ensure_dyn_smem(kern, smem);
  if (e != cudaSuccess) return cuda_rc(e);
  kern<<<grid, threads, smem, st>>>(a);
  return launch_rc();
}
// the warp-specialised G pass (one chunk of K <= 64 clusters): ring depth and shared memory
int natgrad2_stages(int KP, int DP, size_t budget) {
  const size_t fixed = sizeof(double) * (size_t)kG2FixedDoubles + 64;
  const size_t per = sizeof(double) * (size_t)g2_stage_doubles(KP, DP) + 3 * sizeof(uint64_t);
  const int s = (int)((budget - fixed) / per);
  return s > kMaxStages ? kMaxStages : s;
}
size_t natgrad2_smem(int KP, int DP, int S) {
  return sizeof(double) * ((size_t)S * g2_stage_doubles(KP, DP) + kG2FixedDoubles) + sizeof(uint64_t) * 3 * S;
}
template <int KP>
int launch_natgrad(NatgradArgs &a, int grid, cudaStream_t st) {
  if (a.sa_vec == nullptr && a.xg_stride == 0) {
    a.stages = natgrad2_stages(KP, a.DP, kSmemBudget);
    const size_t smem = natgrad2_smem(KP, a.DP, a.stages);
    return a.K == KP ? launch_vb(k_natgrad2<KP, true>, a, smem, grid, kGThreads, st) : launch_vb(k_natgrad2<KP, false>, a, smem, grid, kGThreads, st);
  }
  // sweep 1 of the per-chunk launches (K > 64 without thread-block clusters): the unit design
  a.stages = natgrad_stages(KP, a.DP);
  const size_t smem = natgrad_smem(KP, a.DP, a.stages);
  return a.K == KP ? launch_vb(k_natgrad<KP, true>, a, smem, grid, kGThreads, st) : launch_vb(k_natgrad<KP, false>, a, smem, grid, kGThreads, st);
}
template <int KP>
int launch_step_stats(StepStatsArgs &a, int grid, cudaStream_t st) {
  a.stages = step_stats_stages(KP, a.DP);
  const size_t smem = step_stats_smem(KP, a.DP, a.stages);
  return a.K == KP ? launch_vb(k_step_stats<KP, true>, a, smem, grid, kSThreads, st)
                   : launch_vb(k_step_stats<KP, false>, a, smem, grid, kSThreads, st);
}

// ---- wide mode (K > 64): clusters of KP/64 CTAs, one 64-cluster chunk per CTA (gpc_vb_kernels.cuh)
int natgrad_stages_wide(int DP) {
  const int units = kGConsumers / 2;
  const size_t fixed = sizeof(double) * ((size_t)g_fixed_doubles(GPC_MAX_KP, DP) + g_cl_doubles()) + 64;
  const size_t per = sizeof(double) * (size_t)g_stage_doubles(GPC_MAX_KP, DP) + sizeof(uint64_t);
  int r = (int)((kSmemBudget - fixed) / per) / units;
  if (r > kMaxUnitSlots) r = kMaxUnitSlots;
  return r * units;
}
size_t natgrad_smem_wide(int DP, int S) { return natgrad_smem(GPC_MAX_KP, DP, S) + sizeof(double) * g_cl_doubles(); }
int step_stats_stages_wide(int DP) {
  const size_t fixed = sizeof(double) * ((size_t)kSSoftmaxWarps + s_cl_doubles()) + 64;
  const size_t per = sizeof(double) * (size_t)s_stage_doubles(GPC_MAX_KP, DP) + 3 * sizeof(uint64_t);
  const int s = (int)((kSmemBudget - fixed) / per);
  return s > kMaxStages ? kMaxStages : s;
}
size_t step_stats_smem_wide(int DP, int S) { return step_stats_smem(GPC_MAX_KP, DP, S) + sizeof(double) * s_cl_doubles(); }

template <typename Kern, typename Args>
int launch_wide(Kern kern, const Args &a, size_t smem, int csize, int nclusters, int threads, cudaStream_t st, int *max_clusters) {
  cudaError_t e = ensure_dyn_smem(kern, smem);
  if (e != cudaSuccess) return cuda_rc(e);
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)csize;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (max_clusters != nullptr) {  // query only: how many clusters of this shape are co-resident
    cfg.gridDim = dim3((unsigned)csize);
    e = cudaOccupancyMaxActiveClusters(max_clusters, kern, &cfg);
    return cuda_rc(e);
  }
  cfg.gridDim = dim3((unsigned)(csize * nclusters));
  e = cudaLaunchKernelEx(&cfg, kern, a);
  return e == cudaSuccess ? launch_rc() : cuda_rc(e);
}
int dispatch_natgrad_wide(NatgradArgs &a, int csize, int nclusters, cudaStream_t st, int *max_clusters) {
  a.stages = natgrad_stages_wide(a.DP);
  a.csize = csize;
  const size_t smem = natgrad_smem_wide(a.DP, a.stages);
  return a.K == csize * GPC_MAX_KP ? launch_wide(k_natgrad<GPC_MAX_KP, true, true>, a, smem, csize, nclusters, kGThreads, st, max_clusters)
                                   : launch_wide(k_natgrad<GPC_MAX_KP, false, true>, a, smem, csize, nclusters, kGThreads, st, max_clusters);
}
int dispatch_step_stats_wide(StepStatsArgs &a, int csize, int nclusters, cudaStream_t st, int *max_clusters) {
  a.stages = step_stats_stages_wide(a.DP);
  a.csize = csize;
  const size_t smem = step_stats_smem_wide(a.DP, a.stages);
  return a.K == csize * GPC_MAX_KP ? launch_wide(k_step_stats<GPC_MAX_KP, true, true>, a, smem, csize, nclusters, kSThreads, st, max_clusters)
                                   : launch_wide(k_step_stats<GPC_MAX_KP, false, true>, a, smem, csize, nclusters, kSThreads, st, max_clusters);
}
bool wide_geometry_ok(long long N, int K, int KP, int DP, int nclusters) {
  return N >= 1 && N <= 0x7fffffffLL && K > GPC_MAX_KP && K <= KP && KP <= GPC_MAX_KTOT && KP == gpc_padded_k(K) && DP >= 8 && DP <= GPC_MAX_DP &&
         (DP % 8) == 0 && nclusters >= 1;
}

constexpr size_t kTileSmem = sizeof(double) * ((size_t)kOB * kOStrideH + (size_t)kOKH * kOStrideN);  // P half + the larger Q half
template <typename Kern>
int tile_attr(Kern kern) {
  return cuda_rc(ensure_dyn_smem(kern, kTileSmem));
}
int launch_post(const MohgpPostArgs &a, int KP, int D, cudaStream_t st) {
  // dynamic shared memory: eigen workspace + Sy_inv staged by bulk copy (sizes must be multiples of 16 bytes)
  const size_t smem = sizeof(double) * (size_t)eig_ws_len(D);
  cudaError_t e = ensure_dyn_smem(k_mohgp_post, smem);
  if (e != cudaSuccess) return cuda_rc(e);
  k_mohgp_post<<<KP, kClusterThreads, smem, st>>>(a);
  return launch_rc();
}

bool make_kern_spec(int nparts, const int *kinds, const double *variances, const double *lengthscales, KernSpec &spec) {
  if (nparts < 1 || nparts > kMaxKernParts || !kinds || !variances || !lengthscales) return false;
  spec.nparts = nparts;
  for (int p = 0; p < kMaxKernParts; ++p) {
    spec.kind[p] = p < nparts ? kinds[p] : 0;
    spec.variance[p] = p < nparts ? variances[p] : 0.0;
    spec.lengthscale[p] = p < nparts ? lengthscales[p] : 1.0;
    if (p < nparts && (kinds[p] < GPC_KERN_RBF || kinds[p] > GPC_KERN_WHITE)) return false;
  }
  return true;
}

bool geometry_ok(long long N, int K, int KP, int DP, int grid) {
  return N >= 1 && K >= 1 && K <= KP && KP <= GPC_MAX_KP && KP == gpc_padded_k(K) && DP >= 8 && DP <= GPC_MAX_DP && (DP % 8) == 0 && grid >= 1 &&
         gpc_padded_rows(N) / GPC_ROW_TILE < 0x7fffffffLL;
}

}  // namespace

extern "C" {

int gpc_abi_version(void) { return 1; }

int gpc_padded_k(int K) {
  if (K < 1 || K > GPC_MAX_KTOT) return -1;
  if (K > GPC_MAX_KP) return (K + GPC_MAX_KP - 1) / GPC_MAX_KP * GPC_MAX_KP;  // chunks of 64 clusters
  int kp = 8;
  while (kp < K) kp *= 2;
  return kp;
}
int gpc_padded_d(int D) {
  if (D < 1 || D > GPC_MAX_DP) return -1;
  return (D + 7) / 8 * 8;
}
long long gpc_padded_rows(long long N) { return (N + GPC_ROW_TILE - 1) / GPC_ROW_TILE * GPC_ROW_TILE; }
int gpc_stats_len(int KP, int DP) { return KP * DP + KP + 2; }

int gpc_default_grid(int device) {
  int sms = 0;
  cudaError_t e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  if (e != cudaSuccess) return cuda_rc(e);
  return sms * kCtasPerSm;
}

namespace {
int dispatch_natgrad(NatgradArgs &a, int KP, int grid, cudaStream_t st) {
  if (grid > a.num_groups) grid = a.num_groups;
  switch (KP) {
    case 8: return launch_natgrad<8>(a, grid, st);
    case 16: return launch_natgrad<16>(a, grid, st);
    case 32: return launch_natgrad<32>(a, grid, st);
    case 64: return launch_natgrad<64>(a, grid, st);
  }
  return GPC_EINVAL;
}
int dispatch_step_stats(StepStatsArgs &a, int KP, int grid, cudaStream_t st) {
  // NOTE: every CTA of the launch grid writes one partial block, so the grid is NOT clamped here:
  // CTAs without a group write zeros.
  switch (KP) {
    case 8: return launch_step_stats<8>(a, grid, st);
    case 16: return launch_step_stats<16>(a, grid, st);
    case 32: return launch_step_stats<32>(a, grid, st);
    case 64: return launch_step_stats<64>(a, grid, st);
  }
  return GPC_EINVAL;
}
bool bufs_ok(const gpc_bufs_t *b) {
  if (!b || !b->sd) return false;
  for (int i = 0; i < 3; ++i)
    if (!b->x[i] || !b->lse[i]) return false;
  return b->ng[0] && b->ng[1];
}
}  // namespace

int gpc_natgrad(const double *F, const double *x, const double *lse, const double *Wt, const double *bias, const double *x_old,
                const double *lse_old, const double *ng_old, double *ng_out, double *grad_out, double *partials, unsigned int *counter,
                double *dots, long long N, int K, int KP, int DP, int grid, void *stream) {
  if (!F || !x || !lse || !Wt || !bias || !ng_out || !partials || !counter || !dots) return GPC_EINVAL;
  if ((x_old == nullptr) != (ng_old == nullptr) || (x_old == nullptr) != (lse_old == nullptr)) return GPC_EINVAL;
  if (!geometry_ok(N, K, KP, DP, grid) || N > 0x7fffffffLL) return GPC_EINVAL;
  NatgradArgs a{};
  a.F = F;
  a.x = x;
  a.lse = lse;
  a.lse_old = lse_old;
  a.Wt = Wt;
  a.bias = bias;
  a.x_old = x_old;
  a.ng_old = ng_old;
  a.ng_out = ng_out;
  a.grad_out = grad_out;
  a.partials = partials;
  a.counter = counter;
  a.dots_out = dots;
  a.N = (int)N;
  a.K = K;
  a.DP = DP;
  a.num_groups = (int)(gpc_padded_rows(N) / kGroupRows);
  a.loop = nullptr;
  a.peers = nullptr;
  return dispatch_natgrad(a, KP, grid, as_stream(stream));
}

int gpc_loop_natgrad(const double *F, const gpc_bufs_t *bufs, const double *Wt, const double *bias, double *partials, unsigned int *counter,
                     double *dots, const gpc_loop_t *loop, long long N, int K, int KP, int DP, int grid, void *stream) {
  return gpc_loop_natgrad_px(F, bufs, Wt, bias, partials, counter, dots, loop, nullptr, N, K, KP, DP, grid, stream);
}

int gpc_loop_natgrad_px(const double *F, const gpc_bufs_t *bufs, const double *Wt, const double *bias, double *partials, unsigned int *counter,
                        double *dots, const gpc_loop_t *loop, const gpc_peers_t *peers, long long N, int K, int KP, int DP, int grid,
                        void *stream) {
  if (!F || !bufs_ok(bufs) || !Wt || !bias || !partials || !counter || !dots || !loop) return GPC_EINVAL;
  if (!geometry_ok(N, K, KP, DP, grid) || N > 0x7fffffffLL) return GPC_EINVAL;
  NatgradArgs a{};
  a.F = F;
  a.Wt = Wt;
  a.bias = bias;
  a.partials = partials;
  a.counter = counter;
  a.dots_out = dots;
  a.N = (int)N;
  a.K = K;
  a.DP = DP;
  a.num_groups = (int)(gpc_padded_rows(N) / kGroupRows);
  a.loop = loop;
  a.bufs = *bufs;
  a.peers = peers;
  return dispatch_natgrad(a, KP, grid, as_stream(stream));
}

int gpc_step_stats(const double *F, const double *x_base, const double *ng, const double *sd_old, double *sd_new, double *x_new,
                   double *lse_new, const double *dots, double sq_old, double *beta_out, double *partials, double step, int beta_mode, long long N,
                   int K, int KP, int DP, int grid, void *stream) {
  if (!F || !x_base || !partials) return GPC_EINVAL;
  if (ng && !x_new) return GPC_EINVAL;
  if (beta_mode != GPC_BETA_ZERO && (!dots || !ng)) return GPC_EINVAL;
  if (beta_mode < GPC_BETA_ZERO || beta_mode > GPC_BETA_FR) return GPC_EINVAL;
  if (!geometry_ok(N, K, KP, DP, grid) || N > 0x7fffffffLL) return GPC_EINVAL;
  StepStatsArgs a{};
  a.F = F;
  a.x_base = x_base;
  a.ng = ng;
  a.sd_old = sd_old;
  a.sd_new = sd_new;
  a.x_new = x_new;
  a.lse_new = lse_new;
  a.dots = dots;
  a.sq_old = sq_old;
  a.beta_out = beta_out;
  a.partials = partials;
  a.step = step;
  a.beta_mode = beta_mode;
  a.N = (int)N;
  a.K = K;
  a.DP = DP;
  a.num_groups = (int)(gpc_padded_rows(N) / kGroupRows);
  a.loop = nullptr;
  a.phase = GPC_PHASE_CONJ;
  return dispatch_step_stats(a, KP, grid, as_stream(stream));
}

int gpc_loop_step_stats(const double *F, const gpc_bufs_t *bufs, const double *dots, double *beta_out, double *partials,
                        const gpc_loop_t *loop, int phase, long long N, int K, int KP, int DP, int grid, void *stream) {
  return gpc_loop_step_stats_px(F, bufs, const_cast<double *>(dots), beta_out, partials, loop, nullptr, phase, N, K, KP, DP, grid, stream);
}

int gpc_loop_step_stats_px(const double *F, const gpc_bufs_t *bufs, double *dots, double *beta_out, double *partials, const gpc_loop_t *loop,
                           const gpc_peers_t *peers, int phase, long long N, int K, int KP, int DP, int grid, void *stream) {
  if (!F || !bufs_ok(bufs) || !dots || !beta_out || !partials || !loop) return GPC_EINVAL;
  if (phase != GPC_PHASE_CONJ && phase != GPC_PHASE_RETRY) return GPC_EINVAL;
  if (!geometry_ok(N, K, KP, DP, grid) || N > 0x7fffffffLL) return GPC_EINVAL;
  StepStatsArgs a{};
  a.F = F;
  a.dots = dots;
  a.beta_out = beta_out;
  a.partials = partials;
  a.N = (int)N;
  a.K = K;
  a.DP = DP;
  a.num_groups = (int)(gpc_padded_rows(N) / kGroupRows);
  a.loop = loop;
  a.bufs = *bufs;
  a.phase = phase;
  a.peers = peers;
  a.dots_sum = peers ? dots : nullptr;
  return dispatch_step_stats(a, KP, grid, as_stream(stream));
}

// ------------------------------------------------------------------------------------------------ K > 64: chunks of 64 clusters
int gpc_natgrad_chunk(const double *F, const double *x, const double *lse, const double *Wt, const double *bias, double *a_out, double *sa_vec,
                      int first, double *partials, unsigned int *counter, double *dots_scratch, long long N, int Kc, int DP, long long xg_stride,
                      int grid, void *stream) {
  if (!F || !x || !lse || !Wt || !bias || !a_out || !sa_vec || !partials || !counter || !dots_scratch) return GPC_EINVAL;
  if (N < 1 || N > 0x7fffffffLL || Kc < 1 || Kc > GPC_MAX_KP || DP < 8 || DP > GPC_MAX_DP || (DP % 8) || grid < 1 || xg_stride < 8 * GPC_MAX_KP)
    return GPC_EINVAL;
  NatgradArgs a{};
  a.F = F;
  a.x = x;
  a.lse = lse;
  a.Wt = Wt;
  a.bias = bias;
  a.ng_out = a_out;
  a.partials = partials;
  a.counter = counter;
  a.dots_out = dots_scratch;
  a.N = (int)N;
  a.K = Kc;
  a.DP = DP;
  a.num_groups = (int)(gpc_padded_rows(N) / kGroupRows);
  a.xg_stride = xg_stride;
  a.sa_vec = sa_vec;
  a.sa_first = first;
  return dispatch_natgrad(a, GPC_MAX_KP, grid, as_stream(stream));
}

int gpc_natgrad_finish(double *ng, const double *x, const double *lse, const double *sa_vec, const double *x_old, const double *lse_old,
                       const double *ng_old, double *grad_out, double *partials, unsigned int *counter, double *dots, long long N, int K, int KP,
                       int grid, void *stream) {
  if (!ng || !x || !lse || !sa_vec || !partials || !counter || !dots) return GPC_EINVAL;
  if ((x_old == nullptr) != (ng_old == nullptr) || (x_old == nullptr) != (lse_old == nullptr)) return GPC_EINVAL;
  if (N < 1 || N > 0x7fffffffLL || K < 1 || K > KP || KP != gpc_padded_k(K) || grid < 1) return GPC_EINVAL;
  NatgradFinishArgs a{ng, x, lse, sa_vec, x_old, lse_old, ng_old, grad_out, partials, counter, dots, (int)N, K, KP,
                      (int)(gpc_padded_rows(N) / kGroupRows)};
  k_natgrad_finish<<<grid, 256, 0, as_stream(stream)>>>(a);
  return launch_rc();
}

int gpc_step_lse(const double *x_base, const double *ng, const double *sd_old, double *sd_new, double *x_new, double *lse_new, const double *dots,
                 double sq_old, double *beta_out, double *hpart, double step, int beta_mode, long long N, int K, int KP, int grid, void *stream) {
  if (!x_base || !lse_new || !hpart) return GPC_EINVAL;
  if (ng && !x_new) return GPC_EINVAL;
  if (beta_mode != GPC_BETA_ZERO && (!dots || !ng)) return GPC_EINVAL;
  if (beta_mode < GPC_BETA_ZERO || beta_mode > GPC_BETA_FR) return GPC_EINVAL;
  if (N < 1 || N > 0x7fffffffLL || K < 1 || K > KP || KP != gpc_padded_k(K) || grid < 1) return GPC_EINVAL;
  StepLseArgs a{x_base, ng, sd_old, sd_new, x_new, lse_new, dots, beta_out, hpart, step, sq_old, beta_mode, (int)N, K, KP,
                (int)(gpc_padded_rows(N) / kGroupRows)};
  k_step_lse<<<grid, 256, 0, as_stream(stream)>>>(a);
  return launch_rc();
}

int gpc_stats_chunk(const double *F, const double *x, const double *lse, const double *h_in, double *partials, long long part_stride, int t_off,
                    int ph_off, int h_off, int write_h, long long N, int Kc, int DP, long long xg_stride, int grid, void *stream) {
  if (!F || !x || !lse || !partials || part_stride < 1) return GPC_EINVAL;
  if (N < 1 || N > 0x7fffffffLL || Kc < 1 || Kc > GPC_MAX_KP || DP < 8 || DP > GPC_MAX_DP || (DP % 8) || grid < 1 || xg_stride < 8 * GPC_MAX_KP)
    return GPC_EINVAL;
  StepStatsArgs a{};
  a.F = F;
  a.x_base = x;
  a.partials = partials;
  a.beta_mode = GPC_BETA_ZERO;
  a.N = (int)N;
  a.K = Kc;
  a.DP = DP;
  a.num_groups = (int)(gpc_padded_rows(N) / kGroupRows);
  a.phase = GPC_PHASE_CONJ;
  a.xg_stride = xg_stride;
  a.lse_in = lse;
  a.h_in = h_in;
  a.part_stride = part_stride;
  a.t_off = t_off;
  a.ph_off = ph_off;
  a.h_off = h_off;
  a.write_h = write_h;
  return dispatch_step_stats(a, GPC_MAX_KP, grid, as_stream(stream));
}

// ---- wide mode entry points (K > 64)
int gpc_wide_clusters(int KP, int DP) {
  if (KP <= GPC_MAX_KP || KP > GPC_MAX_KTOT || (KP % GPC_MAX_KP) || DP < 8 || DP > GPC_MAX_DP || (DP % 8)) return GPC_EINVAL;
  const int csize = KP / GPC_MAX_KP;
  NatgradArgs g{};
  g.DP = DP;
  g.K = KP;
  StepStatsArgs s{};
  s.DP = DP;
  s.K = KP;
  int ng = 0, ns = 0;
  int rc = dispatch_natgrad_wide(g, csize, 1, nullptr, &ng);
  if (rc) return rc;
  rc = dispatch_step_stats_wide(s, csize, 1, nullptr, &ns);
  if (rc) return rc;
  const int n = ng < ns ? ng : ns;
  return n >= 1 ? n : -(int)cudaErrorLaunchOutOfResources;
}

int gpc_natgrad_wide(const double *F, const double *x, const double *lse, const double *Wt, const double *bias, const double *x_old,
                     const double *lse_old, const double *ng_old, double *ng_out, double *grad_out, double *partials, unsigned int *counter,
                     double *dots, long long N, int K, int KP, int DP, int nclusters, void *stream) {
  if (!F || !x || !lse || !Wt || !bias || !ng_out || !partials || !counter || !dots) return GPC_EINVAL;
  if ((x_old == nullptr) != (ng_old == nullptr) || (x_old == nullptr) != (lse_old == nullptr)) return GPC_EINVAL;
  if (!wide_geometry_ok(N, K, KP, DP, nclusters)) return GPC_EINVAL;
  NatgradArgs a{};
  a.F = F;
  a.x = x;
  a.lse = lse;
  a.Wt = Wt;
  a.bias = bias;
  a.x_old = x_old;
  a.lse_old = lse_old;
  a.ng_old = ng_old;
  a.ng_out = ng_out;
  a.grad_out = grad_out;
  a.partials = partials;
  a.counter = counter;
  a.dots_out = dots;
  a.N = (int)N;
  a.K = K;
  a.DP = DP;
  a.num_groups = (int)(gpc_padded_rows(N) / kGroupRows);
  a.xg_stride = (long long)kGroupRows * KP;
  return dispatch_natgrad_wide(a, KP / GPC_MAX_KP, nclusters, as_stream(stream), nullptr);
}

int gpc_step_stats_wide(const double *F, const double *x_base, const double *ng, const double *sd_old, double *sd_new, double *x_new,
                        double *lse_new, const double *dots, double sq_old, double *beta_out, double *partials, double step, int beta_mode,
                        long long N, int K, int KP, int DP, int nclusters, void *stream) {
  if (!F || !x_base || !partials) return GPC_EINVAL;
  if (ng && !x_new) return GPC_EINVAL;
  if (beta_mode != GPC_BETA_ZERO && (!dots || !ng)) return GPC_EINVAL;
  if (beta_mode < GPC_BETA_ZERO || beta_mode > GPC_BETA_FR) return GPC_EINVAL;
  if (!wide_geometry_ok(N, K, KP, DP, nclusters)) return GPC_EINVAL;
  StepStatsArgs a{};
  a.F = F;
  a.x_base = x_base;
  a.ng = ng;
  a.sd_old = sd_old;
  a.sd_new = sd_new;
  a.x_new = x_new;
  a.lse_new = lse_new;
  a.dots = dots;
  a.sq_old = sq_old;
  a.beta_out = beta_out;
  a.partials = partials;
  a.step = step;
  a.beta_mode = beta_mode;
  a.N = (int)N;
  a.K = K;
  a.DP = DP;
  a.num_groups = (int)(gpc_padded_rows(N) / kGroupRows);
  a.phase = GPC_PHASE_CONJ;
  a.xg_stride = (long long)kGroupRows * KP;
  return dispatch_step_stats_wide(a, KP / GPC_MAX_KP, nclusters, as_stream(stream), nullptr);
}

int gpc_loop_natgrad_wide(const double *F, const gpc_bufs_t *bufs, const double *Wt, const double *bias, double *partials, unsigned int *counter,
                          double *dots, const gpc_loop_t *loop, const gpc_peers_t *peers, long long N, int K, int KP, int DP, int nclusters,
                          void *stream) {
  if (!F || !bufs_ok(bufs) || !Wt || !bias || !partials || !counter || !dots || !loop) return GPC_EINVAL;
  if (!wide_geometry_ok(N, K, KP, DP, nclusters)) return GPC_EINVAL;
  NatgradArgs a{};
  a.F = F;
  a.Wt = Wt;
  a.bias = bias;
  a.partials = partials;
  a.counter = counter;
  a.dots_out = dots;
  a.N = (int)N;
  a.K = K;
  a.DP = DP;
  a.num_groups = (int)(gpc_padded_rows(N) / kGroupRows);
  a.loop = loop;
  a.bufs = *bufs;
  a.peers = peers;
  a.xg_stride = (long long)kGroupRows * KP;
  return dispatch_natgrad_wide(a, KP / GPC_MAX_KP, nclusters, as_stream(stream), nullptr);
}

int gpc_loop_step_stats_wide(const double *F, const gpc_bufs_t *bufs, double *dots, double *beta_out, double *partials, const gpc_loop_t *loop,
                             const gpc_peers_t *peers, int phase, long long N, int K, int KP, int DP, int nclusters, void *stream) {
  if (!F || !bufs_ok(bufs) || !dots || !beta_out || !partials || !loop) return GPC_EINVAL;
  if (phase != GPC_PHASE_CONJ && phase != GPC_PHASE_RETRY) return GPC_EINVAL;
  if (!wide_geometry_ok(N, K, KP, DP, nclusters)) return GPC_EINVAL;
  StepStatsArgs a{};
  a.F = F;
  a.dots = dots;
  a.beta_out = beta_out;
  a.partials = partials;
  a.N = (int)N;
  a.K = K;
  a.DP = DP;
  a.num_groups = (int)(gpc_padded_rows(N) / kGroupRows);
  a.loop = loop;
  a.bufs = *bufs;
  a.phase = phase;
  a.peers = peers;
  a.dots_sum = peers ? dots : nullptr;
  a.xg_stride = (long long)kGroupRows * KP;
  return dispatch_step_stats_wide(a, KP / GPC_MAX_KP, nclusters, as_stream(stream), nullptr);
}

int gpc_reduce_partials(const double *partials, int num_parts, int len, double *out, void *stream) {
  if (!partials || !out || num_parts < 1 || len < 1) return GPC_EINVAL;
  k_reduce_partials<<<(len + 31) / 32, 256, 0, as_stream(stream)>>>(partials, num_parts, len, out);
  return launch_rc();
}

long long gpc_xchg_doubles(int world, int stats_len) {
  return (world < 1 || world > GPC_MAX_PEERS || stats_len < 1) ? -1 : xchg_doubles(world, stats_len);
}

int gpc_reduce_push(const double *partials, int num_parts, int len, const gpc_peers_t *peers, unsigned int *counter, const gpc_loop_t *loop,
                    int phase, void *stream) {
  if (!partials || !peers || !counter || num_parts < 1 || len < 1) return GPC_EINVAL;
  k_reduce_push<<<(len + 31) / 32, 256, 0, as_stream(stream)>>>(partials, num_parts, len, peers, counter, loop, phase);
  return launch_rc();
}

int gpc_mohgp_setup(const double *Sf, const double *Sy, double *Ly, double *Ly_inv, double *Sy_inv, double *A, double *Sy_logdet,
                    int *info, int D, void *stream) {
  if (!Sf || !Sy || !Ly || !Ly_inv || !Sy_inv || !A || !Sy_logdet || !info || D < 1 || D > GPC_MAX_DP) return GPC_EINVAL;
  cudaStream_t st = as_stream(stream);
  cudaError_t e = cudaMemsetAsync(info, 0, sizeof(int), st);
  if (e != cudaSuccess) return cuda_rc(e);
  const size_t smem = sizeof(double) * 3 * (size_t)D * (D + 1);
  e = ensure_dyn_smem(k_mohgp_setup, smem);
  if (e != cudaSuccess) return cuda_rc(e);
  MohgpSetupArgs a{Sf, Sy, Ly, Ly_inv, Sy_inv, A, Sy_logdet, info, D};
  k_mohgp_setup<<<1, kClusterThreads, smem, st>>>(a);
  return launch_rc();
}

__global__ void k_mohgp_const(const double *YTY, const double *Sy_inv, const double *Sy_logdet, double n_total, int D, double *out) {
  __shared__ double red_s[32];
  double v = 0.0;
  for (int e = threadIdx.x; e < D * D; e += blockDim.x) v += YTY[e] * Sy_inv[e];
  v = block_sum(v, red_s);
  if (threadIdx.x == 0) *out = -0.5 * (n_total * D * log(2.0 * M_PI) + n_total * (*Sy_logdet) + v);
}

int gpc_mohgp_const(const double *YTY, const double *Sy_inv, const double *Sy_logdet, double n_total, int D, double *const_term,
                    void *stream) {
  if (!YTY || !Sy_inv || !Sy_logdet || !const_term || D < 1) return GPC_EINVAL;
  k_mohgp_const<<<1, 256, 0, as_stream(stream)>>>(YTY, Sy_inv, Sy_logdet, n_total, D, const_term);
  return launch_rc();
}

int gpc_mohgp_cluster(const double *stats, const double *A, const double *Ly, const double *Ly_inv, const double *Sy,
                      const double *Sy_inv, const double *Sf, double *Wt, double *muk_t, double *per_k, double *Syi_ybark_t, double *Lambda_inv, double *C_chols, int *info,
                      int K, int D, int KP, int DP, void *stream) {
  if (!stats || !A || !Ly || !Ly_inv || !Sy || !Sy_inv || !Sf || !Wt || !muk_t || !per_k || !info) return GPC_EINVAL;
  if (K < 1 || K > KP || D < 1 || D > DP || DP > GPC_MAX_DP) return GPC_EINVAL;
  cudaStream_t st = as_stream(stream);
  cudaError_t e = cudaMemsetAsync(info, 0, sizeof(int), st);
  if (e != cudaSuccess) return cuda_rc(e);
  const size_t smem = sizeof(double) * (3 * (size_t)GPC_MAX_DP * (GPC_MAX_DP + 1) + 3 * GPC_MAX_DP + 32);
  e = ensure_dyn_smem(k_mohgp_cluster, smem);
  if (e != cudaSuccess) return cuda_rc(e);
  MohgpClusterArgs a{stats, A, Ly, Ly_inv, Sy, Sy_inv, Sf, Wt, muk_t, per_k, Syi_ybark_t, Lambda_inv, C_chols, info, K, D, KP, DP};
  k_mohgp_cluster<<<KP, kClusterThreads, smem, st>>>(a);
  return launch_rc();
}

int gpc_mohgp_eig_ws_len(int D) { return (D < 1 || D > GPC_MAX_DP) ? -1 : eig_ws_len(D); }

int gpc_mohgp_eig(const double *A, const double *Ly, const double *Ly_inv, const double *Sy_inv, const double *Sf, int D, double *ws,
                  void *stream) {
  if (!A || !Ly || !Ly_inv || !Sy_inv || !Sf || !ws || D < 1 || D > GPC_MAX_DP) return GPC_EINVAL;
  const size_t smem = sizeof(double) * (3 * (size_t)GPC_MAX_DP * (GPC_MAX_DP + 1) + GPC_MAX_DP + 32);  // A, V, G = A V + rotation scratch
  cudaError_t e = ensure_dyn_smem(k_mohgp_eig, smem);
  if (e != cudaSuccess) return cuda_rc(e);
  MohgpEigArgs a{A, Ly, Ly_inv, Sy_inv, Sf, ws, D};
  k_mohgp_eig<<<1, kEigThreads, smem, as_stream(stream)>>>(a);
  return launch_rc();
}

int gpc_mohgp_kern_grads(const double *stats, const double *eig_ws, const double *Sf, const double *Sy_inv, const double *YTY, const double *Syi_ybark_t,
                         double n_total, int K, int D, int KP, int DP, double *work, double *dL_dSf, double *dL_dSy, int *info, void *stream) {
  if (!stats || !eig_ws || !Sf || !Sy_inv || !YTY || !Syi_ybark_t || !work || !dL_dSf || !dL_dSy || !info) return GPC_EINVAL;
  if (K < 1 || K > KP || KP > GPC_MAX_KTOT || D < 1 || D > DP || DP > GPC_MAX_DP) return GPC_EINVAL;
  cudaStream_t st = as_stream(stream);
  int rc = cuda_rc(cudaMemsetAsync(info, 0, sizeof(int), st));
  if (rc) return rc;
  const size_t KD = (size_t)K * D;
  double *t1 = work, *by = work + KD, *w1 = work + 2 * KD, *w2 = work + 3 * KD, *cw = work + 4 * KD, *scratch = cw + 2 * (size_t)K;
  MohgpKgradArgs a{stats, eig_ws, Sf, Syi_ybark_t, t1, by, w1, w2, cw, info, K, D, KP, DP};
  k_mohgp_kgrad_cluster<<<K, kClusterThreads, 0, st>>>(a);
  if ((rc = launch_rc())) return rc;
  MohgpKgradFinalArgs f{t1, by, Syi_ybark_t, w1, w2, cw, eig_ws, Sy_inv, YTY, dL_dSf, dL_dSy, scratch, n_total, K, D};
  k_mohgp_kgrad_final<<<1, kKgradThreads, 0, st>>>(f);
  return launch_rc();
}
long long gpc_mohgp_kern_grads_work(int K, int D) { return (K < 1 || D < 1) ? -1 : 4LL * K * D + 2LL * K + (long long)D * D; }

int gpc_loop_mohgp_post(const double *partials, int num_parts, double *stats, const double *eig_ws, const double *Sy_inv, const double *Sf,
                        const double *const_term, double *Wt, double *muk_t, double *Syi_ybark_t, double *per_k, double *bias, double *scalars,
                        double *mixgrad_out, unsigned int *counter, int *info, double alpha, int prior, int K, int D, int KP, int DP,
                        const double *dots, const double *beta, gpc_loop_t *loop, int phase, void *stream) {
  if (!partials || num_parts < 1 || !stats || !eig_ws || !Sy_inv || !Sf || !const_term || !Wt || !muk_t || !per_k || !bias || !scalars ||
      !counter || !info)
    return GPC_EINVAL;
  if (K < 1 || K > KP || KP > GPC_MAX_KTOT || D < 1 || D > DP || DP > GPC_MAX_DP) return GPC_EINVAL;
  if (prior != GPC_PRIOR_SYMMETRIC && prior != GPC_PRIOR_DP) return GPC_EINVAL;
  if (loop && (!dots || !beta || (phase != GPC_PHASE_CONJ && phase != GPC_PHASE_RETRY))) return GPC_EINVAL;
  cudaStream_t st = as_stream(stream);
  // the flag word is cleared by a tiny stream-ordered memset; in loop mode the kernel's last CTA consumes and re-arms it
  // (the first loop evaluation follows a synchronous one that left it clear unless it failed -- and then raised)
  if (!loop) {
    cudaError_t e = cudaMemsetAsync(info, 0, sizeof(int), st);
    if (e != cudaSuccess) return cuda_rc(e);
  }
  MohgpPostArgs a{partials, num_parts, stats, eig_ws, Sy_inv, Sf, const_term, Wt, muk_t, Syi_ybark_t, per_k, bias, scalars, mixgrad_out,
                  counter, info, alpha, prior, K, D, KP, DP, dots, beta, loop, phase, nullptr};
  return launch_post(a, KP, D, st);
}

int gpc_loop_mohgp_post_px(const gpc_peers_t *peers, int world, double *stats, const double *eig_ws, const double *Sy_inv, const double *Sf,
                           const double *const_term, double *Wt, double *muk_t, double *Syi_ybark_t, double *per_k, double *bias,
                           double *scalars, double *mixgrad_out, unsigned int *counter, int *info, double alpha, int prior, int K, int D,
                           int KP, int DP, const double *dots, const double *beta, gpc_loop_t *loop, int phase, void *stream) {
  if (!peers || world < 1 || world > GPC_MAX_PEERS || !stats || !eig_ws || !Sy_inv || !Sf || !const_term || !Wt || !muk_t || !per_k || !bias ||
      !scalars || !counter || !info || !dots || !beta || !loop)
    return GPC_EINVAL;
  if (K < 1 || K > KP || KP > GPC_MAX_KTOT || D < 1 || D > DP || DP > GPC_MAX_DP) return GPC_EINVAL;
  if (prior != GPC_PRIOR_SYMMETRIC && prior != GPC_PRIOR_DP) return GPC_EINVAL;
  if (phase != GPC_PHASE_CONJ && phase != GPC_PHASE_RETRY) return GPC_EINVAL;
  cudaStream_t st = as_stream(stream);
  MohgpPostArgs a{stats, world, stats, eig_ws, Sy_inv, Sf, const_term, Wt, muk_t, Syi_ybark_t, per_k, bias, scalars, mixgrad_out,
                  counter, info, alpha, prior, K, D, KP, DP, dots, beta, loop, phase, peers};
  return launch_post(a, KP, D, st);
}

int gpc_mohgp_post(const double *partials, int num_parts, double *stats, const double *eig_ws, const double *Sy_inv, const double *Sf,
                   const double *const_term, double *Wt, double *muk_t, double *Syi_ybark_t, double *per_k, double *bias, double *scalars,
                   double *mixgrad_out, unsigned int *counter, int *info, double alpha, int prior, int K, int D, int KP, int DP, void *stream) {
  return gpc_loop_mohgp_post(partials, num_parts, stats, eig_ws, Sy_inv, Sf, const_term, Wt, muk_t, Syi_ybark_t, per_k, bias, scalars, mixgrad_out,
                             counter, info, alpha, prior, K, D, KP, DP, nullptr, nullptr, nullptr, GPC_PHASE_CONJ, stream);
}

int gpc_mohgp_finalize(const double *stats, const double *per_k, const double *const_term, double *bias, double *scalars,
                       double *mixgrad_out, double alpha, int prior, int K, int KP, int DP, void *stream) {
  if (!stats || !per_k || !const_term || !bias || !scalars || K < 1 || K > KP || KP > kMaxMixK) return GPC_EINVAL;
  if (prior != GPC_PRIOR_SYMMETRIC && prior != GPC_PRIOR_DP) return GPC_EINVAL;
  MohgpFinalizeArgs a{stats, per_k, bias, scalars, mixgrad_out, 0.0, const_term, alpha, prior, K, KP, DP};
  k_mohgp_finalize<<<1, kClusterThreads, 0, as_stream(stream)>>>(a);
  return launch_rc();
}

int gpc_mog_features(const double *X, const double *centre, long long N, int D, int DP, double *F, void *stream) {
  if (!X || !centre || !F || N < 1 || D < 1 || D * (D + 1) / 2 + D > DP) return GPC_EINVAL;
  if (DP < 8 || DP > GPC_MAX_DP || (DP % 8)) return GPC_EINVAL;
  const long long total = gpc_padded_rows(N) * DP;
  k_mog_features<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(X, centre, (int)N, (int)gpc_padded_rows(N), D, DP, F);
  return launch_rc();
}

int gpc_mog_cluster(const double *stats, const double *m0c, const double *S0, double *Wt, double *per_k, double *mun_c, double *Sns,
                    double *Sns_inv, int *info, double k0, double v0, int K, int D, int KP, int DP, void *stream) {
  if (!stats || !m0c || !S0 || !Wt || !per_k || !mun_c || !Sns || !Sns_inv || !info) return GPC_EINVAL;
  if (K < 1 || K > KP || D < 1 || D > kMogMaxD || D * (D + 1) / 2 + D > DP) return GPC_EINVAL;
  cudaStream_t st = as_stream(stream);
  cudaError_t e = cudaMemsetAsync(info, 0, sizeof(int), st);
  if (e != cudaSuccess) return cuda_rc(e);
  MogClusterArgs a{stats, m0c, S0, Wt, per_k, mun_c, Sns, Sns_inv, info, k0, v0, K, D, KP, DP, nullptr, GPC_PHASE_CONJ};
  k_mog_cluster<<<(KP + 31) / 32, 32, 0, st>>>(a);
  return launch_rc();
}

int gpc_mog_finalize(const double *stats, const double *per_k, double *bias, double *scalars, double *mixgrad_out, double n_total,
                     double k0, double v0, double S0_halflogdet, double alpha, int prior, int K, int KP, int DP, int D, void *stream) {
  if (!stats || !per_k || !bias || !scalars || K < 1 || K > KP || KP > kMaxMixK) return GPC_EINVAL;
  if (prior != GPC_PRIOR_SYMMETRIC && prior != GPC_PRIOR_DP) return GPC_EINVAL;
  MogFinalizeArgs a{stats, per_k, bias, scalars, mixgrad_out, n_total, k0, v0, S0_halflogdet, alpha, prior, K, KP, DP, D, nullptr, nullptr, nullptr, nullptr, GPC_PHASE_CONJ};
  k_mog_finalize<<<1, kClusterThreads, 0, as_stream(stream)>>>(a);
  return launch_rc();
}

int gpc_loop_mog_post(const double *stats, const double *m0c, const double *S0, double *Wt, double *per_k, double *mun_c, double *Sns,
                      double *Sns_inv, int *info, double k0, double v0, double *bias, double *scalars, double *mixgrad_out, double n_total,
                      double S0_halflogdet, double alpha, int prior, int K, int D, int KP, int DP, const double *dots, const double *beta,
                      gpc_loop_t *loop, int phase, void *stream) {
  if (!stats || !m0c || !S0 || !Wt || !per_k || !mun_c || !Sns || !Sns_inv || !info || !bias || !scalars || !dots || !beta || !loop) return GPC_EINVAL;
  if (K < 1 || K > KP || KP > kMaxMixK || D < 1 || D > kMogMaxD || D * (D + 1) / 2 + D > DP) return GPC_EINVAL;
  if (prior != GPC_PRIOR_SYMMETRIC && prior != GPC_PRIOR_DP) return GPC_EINVAL;
  if (phase != GPC_PHASE_CONJ && phase != GPC_PHASE_RETRY) return GPC_EINVAL;
  cudaStream_t st = as_stream(stream);
  cudaError_t e = cudaMemsetAsync(info, 0, sizeof(int), st);
  if (e != cudaSuccess) return cuda_rc(e);
  MogClusterArgs c{stats, m0c, S0, Wt, per_k, mun_c, Sns, Sns_inv, info, k0, v0, K, D, KP, DP, loop, phase};
  k_mog_cluster<<<(KP + 31) / 32, 32, 0, st>>>(c);
  MogFinalizeArgs f{stats, per_k, bias, scalars, mixgrad_out, n_total, k0, v0, S0_halflogdet, alpha, prior, K, KP, DP, D, dots, beta, info, loop, phase};
  k_mog_finalize<<<1, kClusterThreads, 0, st>>>(f);
  return launch_rc();
}

int gpc_kern_fill(const double *X, const double *X2, int n, int m, int q, int nparts, const int *kinds, const double *variances,
                  const double *lengthscales, double diag_add, double *out, void *stream) {
  if (!X || !out || n < 1 || m < 1 || q < 1 || nparts < 1 || nparts > kMaxKernParts || !kinds || !variances || !lengthscales)
    return GPC_EINVAL;
  if (!X2 && m != n) return GPC_EINVAL;
  KernSpec spec;
  if (!make_kern_spec(nparts, kinds, variances, lengthscales, spec)) return GPC_EINVAL;
  const long long total = (long long)n * m;
  k_kern_fill<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(X, X2, n, m, q, spec, diag_add, out);
  return launch_rc();
}

int gpc_softmax_rows(const double *x, long long N, int K, int ldx, double *phi, double *logphi, int ldo, double *Hpart, int grid,
                     void *stream) {
  if (!x || N < 1 || K < 1 || ldx < K || ((phi || logphi) && ldo < K) || grid < 1 || N > 0x7fffffffLL) return GPC_EINVAL;
  k_softmax_rows<<<grid, kSoftmaxWarps * 32, 0, as_stream(stream)>>>(x, (int)N, K, ldx, phi, logphi, ldo, Hpart);
  return launch_rc();
}

int gpc_multiple_mahalanobis(const double *X1, const double *X2, const double *L, long long N1, int N2, int D, double *out,
                             void *stream) {
  if (!X1 || !X2 || !L || !out || N1 < 1 || N2 < 1 || D < 1 || D > kMahaMaxD) return GPC_EINVAL;
  const size_t smem = sizeof(double) * (size_t)D * D;
  cudaError_t e = ensure_dyn_smem(k_mahalanobis_pairs, smem);
  if (e != cudaSuccess) return cuda_rc(e);
  const long long pairs = N1 * N2;
  k_mahalanobis_pairs<<<(unsigned)((pairs + 127) / 128), 128, smem, as_stream(stream)>>>(X1, X2, L, (int)N1, N2, D, out);
  return launch_rc();
}

int gpc_multiple_pdinv(const double *A, int D, int K, double *inv, double *hld, int *info, void *stream) {
  if (!A || !inv || !hld || !info || D < 1 || D > GPC_MAX_DP || K < 1) return GPC_EINVAL;
  cudaStream_t st = as_stream(stream);
  cudaError_t e = cudaMemsetAsync(info, 0, sizeof(int), st);
  if (e != cudaSuccess) return cuda_rc(e);
  const size_t smem = sizeof(double) * ((size_t)D * D + 2 * (size_t)D * (D + 1));
  e = ensure_dyn_smem(k_batched_pdinv, smem);
  if (e != cudaSuccess) return cuda_rc(e);
  k_batched_pdinv<<<K, kClusterThreads, smem, st>>>(A, D, K, inv, hld, info);
  return launch_rc();
}

int gpc_gram_partials(const double *F, long long Npad, int DP, double *partials, int grid, void *stream) {
  if (!F || !partials || Npad < kGramRows || (Npad % kGramRows) || DP < 8 || DP > GPC_MAX_DP || grid < 1) return GPC_EINVAL;
  k_gram_partials<<<grid, 256, sizeof(double) * kGramRows * DP, as_stream(stream)>>>(F, (int)Npad, DP, partials);
  return launch_rc();
}

int gpc_pad_rows(const double *src, long long N, int D, int DP, double *dst, void *stream) {
  if (!src || !dst || N < 1 || D < 1 || DP < D) return GPC_EINVAL;
  const long long total = gpc_padded_rows(N) * DP;
  k_pad_rows<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(src, (int)N, D, (int)gpc_padded_rows(N), DP, dst);
  return launch_rc();
}

int gpc_normalise_rows(const double *src, long long N, int D, double *dst, void *stream) {
  if (!src || !dst || N < 1 || D < 1 || N > 0x7fffffffLL) return GPC_EINVAL;
  const long long blocks = (N + 7) / 8;
  k_normalise_rows<<<(unsigned)(blocks > 65536 ? 65536 : blocks), 256, 0, as_stream(stream)>>>(src, (int)N, D, dst);
  return launch_rc();
}

int gpc_replicate_scores(const double *Y, long long N, int D, const int *group, int G, double *score, void *stream) {
  if (!Y || !group || !score || N < 1 || D < 1 || G < 1 || G > 64 || N > 0x7fffffffLL) return GPC_EINVAL;
  k_replicate_scores<<<(unsigned)((N + 127) / 128), 128, 0, as_stream(stream)>>>(Y, (int)N, D, group, G, score);
  return launch_rc();
}

int gpc_pack_rows(const double *src, long long N, int C, int CP, int layout, double *dst, void *stream) {
  if (!src || !dst || N < 1 || C < 1 || CP < C || (CP % 8) || N > 0x7fffffffLL) return GPC_EINVAL;
  const long long total = gpc_padded_rows(N) * CP;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  if (layout == GPC_LAYOUT_LOGITS)
    k_pack_rows<true><<<blocks, 256, 0, as_stream(stream)>>>(src, (int)N, C, (int)gpc_padded_rows(N), CP, dst);
  else if (layout == GPC_LAYOUT_FEATURES)
    k_pack_rows<false><<<blocks, 256, 0, as_stream(stream)>>>(src, (int)N, C, (int)gpc_padded_rows(N), CP, dst);
  else
    return GPC_EINVAL;
  return launch_rc();
}

int gpc_unpack_rows(const double *src, long long N, int C, int CP, int layout, double *dst, void *stream) {
  if (!src || !dst || N < 1 || C < 1 || CP < C || (CP % 8) || N > 0x7fffffffLL) return GPC_EINVAL;
  const long long total = N * C;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  if (layout == GPC_LAYOUT_LOGITS)
    k_unpack_rows<true><<<blocks, 256, 0, as_stream(stream)>>>(src, (int)N, C, CP, dst);
  else if (layout == GPC_LAYOUT_FEATURES)
    k_unpack_rows<false><<<blocks, 256, 0, as_stream(stream)>>>(src, (int)N, C, CP, dst);
  else
    return GPC_EINVAL;
  return launch_rc();
}

int gpc_unpad_rows(const double *src, long long N, int K, int ld, double *dst, void *stream) {
  if (!src || !dst || N < 1 || K < 1 || ld < K) return GPC_EINVAL;
  const long long total = N * K;
  k_unpad_rows<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(src, (int)N, K, ld, dst);
  return launch_rc();
}

// ------------------------------------------------------------------------------------------------ OMGP
long long gpc_omgp_padded_n(long long N) { return (N + kOB - 1) / kOB * kOB; }
int gpc_omgp_slabs(long long Npad) { return (int)((Npad + kOmgpSlabRows - 1) / kOmgpSlabRows); }
int gpc_omgp_part_stride(void) { return kOmgpMaxDy + 1; }

int gpc_omgp_build(const double *X, long long N, int q, int nparts, const int *kinds, const double *variances, const double *lengthscales,
                   const double *w, int w_stride, double variance, double eps, double jitter, double *B, void *stream) {
  if (!X || !w || !B || N < 1 || N > 0x3fffffffLL || q < 1 || w_stride < 1) return GPC_EINVAL;
  KernSpec spec;
  if (!make_kern_spec(nparts, kinds, variances, lengthscales, spec)) return GPC_EINVAL;
  const long long Npad = gpc_omgp_padded_n(N), total = Npad * Npad;
  k_omgp_build<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(X, (int)N, (int)Npad, q, spec, w, w_stride, variance, eps, jitter, B);
  return launch_rc();
}

namespace {
constexpr size_t kDiagSmem = sizeof(double) * 2 * kOB * (kOB + 1);
int omgp_attrs() {
  int rc;
  if ((rc = cuda_rc(ensure_dyn_smem(k_omgp_chol_diag, kDiagSmem)))) return rc;
  if ((rc = tile_attr(k_omgp_chol_panel))) return rc;
  if ((rc = tile_attr(k_omgp_chol_within))) return rc;
  if ((rc = tile_attr(k_omgp_chol_big))) return rc;
  if ((rc = tile_attr(k_omgp_inv_row))) return rc;
  if ((rc = tile_attr(k_omgp_inv_within))) return rc;
  if ((rc = tile_attr(k_omgp_inv_big))) return rc;
  return 0;
}
// factorisation of the block-column group [p0, p1): diagonal tile, panel and the group's own remaining columns, tile column by tile column
void chol_group(double *A, int ld, int nb, int p0, int p1, double *dinv, double *ldsum, int *info, cudaStream_t st) {
  for (int jb = p0; jb < p1; ++jb) {
    k_omgp_chol_diag<<<1, kClusterThreads, kDiagSmem, st>>>(A, ld, jb, dinv, ldsum, info);
    const int m = nb - jb - 1;
    if (m > 0) k_omgp_chol_panel<<<m, 256, kTileSmem, st>>>(A, ld, jb, dinv);
    if (m > 0 && p1 - jb - 1 > 0) k_omgp_chol_within<<<dim3(m, p1 - jb - 1), 256, kTileSmem, st>>>(A, ld, jb);
  }
}
void chol_big(double *A, int ld, int nb, int p0, int p1, cudaStream_t st) {
  const int m = nb - p1;
  if (m > 0) k_omgp_chol_big<<<m * (m + 1) / 2, 256, kTileSmem, st>>>(A, ld, p0, p1);
}
// inverse steps of the group [p0, p1) (W's diagonal tiles of the group must be in place) and the push to the rows below it
void inv_group(double *W, int ld, const double *L, int nb, int p0, int p1, const double *dinv, cudaStream_t st) {
  for (int ib = p0; ib < p1; ++ib) {
    if (ib > 0) k_omgp_inv_row<<<ib, 256, kTileSmem, st>>>(W, ld, ib, dinv);
    if (p1 - ib - 1 > 0) k_omgp_inv_within<<<dim3(ib + 1, p1 - ib - 1), 256, kTileSmem, st>>>(W, ld, L, ld, ib);
  }
  if (nb - p1 > 0) k_omgp_inv_big<<<dim3(p1, nb - p1), 256, kTileSmem, st>>>(W, ld, L, ld, p0, p1);
}
}  // namespace

int gpc_chol_blocked(double *A, long long Npad, double *dinv, double *ldsum, int *info, void *stream) {
  if (!A || !dinv || !ldsum || !info || Npad < kOB || (Npad % kOB) || Npad > 0x3fffffffLL) return GPC_EINVAL;
  cudaStream_t st = as_stream(stream);
  int rc = cuda_rc(cudaMemsetAsync(info, 0, sizeof(int), st));
  if (rc) return rc;
  if ((rc = omgp_attrs())) return rc;
  const int nb = (int)(Npad / kOB), ld = (int)Npad;
  for (int p0 = 0; p0 < nb; p0 += kOGroup) {
    const int p1 = p0 + kOGroup < nb ? p0 + kOGroup : nb;
    chol_group(A, ld, nb, p0, p1, dinv, ldsum, info, st);
    chol_big(A, ld, nb, p0, p1, st);
  }
  return launch_rc();
}

// Factorisation and triangular inverse pipelined over two streams.  The inverse steps of a group need only that group's block columns of
// L and its diagonal-tile inverses, final once the group is factorised: the aux stream runs inverse group g underneath the main
// stream's trailing update of group g and the factorisation of group g+1.  Capturable (the events become graph edges).
int gpc_chol_inverse(double *A, long long Npad, double *dinv, double *ldsum, int *info, double *W, void *stream, void *aux_stream) {
  if (!A || !dinv || !ldsum || !info || !W || !aux_stream || stream == aux_stream) return GPC_EINVAL;
  if (Npad < kOB || (Npad % kOB) || Npad > 0x3fffffffLL) return GPC_EINVAL;
  cudaStream_t st = as_stream(stream), ax = as_stream(aux_stream);
  int rc = cuda_rc(cudaMemsetAsync(info, 0, sizeof(int), st));
  if (rc) return rc;
  if ((rc = omgp_attrs())) return rc;
  const int nb = (int)(Npad / kOB), ld = (int)Npad;
  const long long total = Npad * Npad;
  cudaEvent_t ev;
  if ((rc = cuda_rc(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)))) return rc;
  // every event call is an edge of the two-stream pipeline: the first failure is kept and returned (a lost edge would silently
  // drop a dependency between the factorisation and the inverse)
  auto edge = [&rc](cudaError_t e) {
    if (rc == 0 && e != cudaSuccess) rc = cuda_rc(e);
  };
  edge(cudaEventRecord(ev, st));  // fork: the aux stream starts after everything enqueued on the main stream so far
  edge(cudaStreamWaitEvent(ax, ev, 0));
  k_omgp_inv_zero<<<(unsigned)((total / 2 + 255) / 256), 256, 0, ax>>>(W, total);
  for (int p0 = 0; p0 < nb && rc == 0; p0 += kOGroup) {
    const int p1 = p0 + kOGroup < nb ? p0 + kOGroup : nb;
    chol_group(A, ld, nb, p0, p1, dinv, ldsum, info, st);
    edge(cudaEventRecord(ev, st));  // the group's block columns of L are final (one event object, re-recorded per group:
    edge(cudaStreamWaitEvent(ax, ev, 0));  // a wait captures the record that precedes it)
    chol_big(A, ld, nb, p0, p1, st);
    for (int ib = p0; ib < p1; ++ib) k_omgp_inv_diag<<<1, 256, 0, ax>>>(W, ld, ib, dinv);
    inv_group(W, ld, A, nb, p0, p1, dinv, ax);
  }
  edge(cudaEventRecord(ev, ax));  // join (always attempted: a forked capture must be joined back)
  edge(cudaStreamWaitEvent(st, ev, 0));
  cudaEventDestroy(ev);
  return rc ? rc : launch_rc();
}

int gpc_tri_inverse_blocked(const double *L, long long Npad, const double *dinv, double *W, void *stream) {
  if (!L || !dinv || !W || Npad < kOB || (Npad % kOB) || Npad > 0x3fffffffLL) return GPC_EINVAL;
  cudaStream_t st = as_stream(stream);
  int rc;
  if ((rc = omgp_attrs())) return rc;
  const int nb = (int)(Npad / kOB), ld = (int)Npad;
  const long long total = Npad * Npad;
  k_omgp_inv_init<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(W, ld, dinv);
  for (int p0 = 0; p0 < nb; p0 += kOGroup) {
    const int p1 = p0 + kOGroup < nb ? p0 + kOGroup : nb;
    inv_group(W, ld, L, nb, p0, p1, dinv, st);
  }
  return launch_rc();
}

int gpc_omgp_component(const double *W, const double *Y, const double *ldsum, const double *phi_col, int phi_stride, long long N, int Dy,
                       double variance, double *tbuf, double *part, double *grad_col, int grad_stride, double *comp, void *stream) {
  if (!W || !Y || !ldsum || !phi_col || !tbuf || !part || !grad_col || !comp || N < 1 || Dy < 1 || Dy > kOmgpMaxDy) return GPC_EINVAL;
  cudaStream_t st = as_stream(stream);
  const int Npad = (int)gpc_omgp_padded_n(N), nslab = gpc_omgp_slabs(Npad);
  k_omgp_wy<<<(Npad + 7) / 8, 256, 0, st>>>(W, Npad, Y, Dy, tbuf);
  k_omgp_colstats<<<dim3(Npad / 64, nslab), 256, 0, st>>>(W, Npad, tbuf, Dy, kOmgpSlabRows, part);
  OmgpCompArgs a{part, Y, phi_col, ldsum, grad_col, comp, variance, (int)N, Npad, Dy, nslab, phi_stride, grad_stride};
  k_omgp_component<<<1, 1024, 0, st>>>(a);
  return launch_rc();
}

long long gpc_omgp_kgrad_tiles(long long Npad) {
  const long long nb = Npad / kOB;
  return nb * (nb + 1) / 2;
}

int gpc_omgp_kern_grads(const double *W, const double *X, int q, const double *Y, long long N, int Dy, int nparts, const int *kinds,
                        const double *variances, const double *lengthscales, double *tbuf, double *part, double *alpha, double *dldb_diag,
                        double *tile_part, double *out, void *stream) {
  if (!W || !X || !Y || !tbuf || !part || !alpha || !dldb_diag || !tile_part || !out) return GPC_EINVAL;
  if (N < 1 || N > 0x3fffffffLL || q < 1 || Dy < 1 || Dy > kOmgpMaxDy) return GPC_EINVAL;
  OmgpKgradArgs a{};
  if (!make_kern_spec(nparts, kinds, variances, lengthscales, a.spec)) return GPC_EINVAL;
  cudaStream_t st = as_stream(stream);
  const int Npad = (int)gpc_omgp_padded_n(N), nslab = gpc_omgp_slabs(Npad);
  const long long tiles = gpc_omgp_kgrad_tiles(Npad);
  if (tiles > 0x7fffffffLL) return GPC_EINVAL;
  k_omgp_wy<<<(Npad + 7) / 8, 256, 0, st>>>(W, Npad, Y, Dy, tbuf);
  k_omgp_colstats<<<dim3(Npad / 64, nslab), 256, 0, st>>>(W, Npad, tbuf, Dy, kOmgpSlabRows, part);
  k_omgp_alpha<<<(Npad + 255) / 256, 256, 0, st>>>(part, nslab, (int)N, Npad, Dy, alpha, dldb_diag);
  a.W = W;
  a.X = X;
  a.alpha = alpha;
  a.part = tile_part;
  a.N = (int)N;
  a.Npad = Npad;
  a.q = q;
  a.Dy = Dy;
  constexpr size_t smem = sizeof(double) * 2 * kOB * kOStrideN;
  int rc = cuda_rc(ensure_dyn_smem(k_omgp_kgrad_tiles, smem));
  if (rc) return rc;
  k_omgp_kgrad_tiles<<<(unsigned)tiles, 256, smem, st>>>(a);
  k_reduce_partials<<<1, 256, 0, st>>>(tile_part, (int)tiles, 2 * kMaxKernParts, out);
  return launch_rc();
}

int gpc_omgp_finalize(const double *comp, const double *Hpart, int n_hpart, double *scalars, double *phi_hat, double *mixgrad, double variance,
                      double alpha, int prior, int K, int Dy, void *stream) {
  if (!comp || !Hpart || !scalars || !phi_hat || !mixgrad || K < 1 || K > kMaxMixK || n_hpart < 1) return GPC_EINVAL;
  if (prior != GPC_PRIOR_SYMMETRIC && prior != GPC_PRIOR_DP) return GPC_EINVAL;
  OmgpFinalizeArgs a{comp, Hpart, n_hpart, scalars, phi_hat, mixgrad, variance, alpha, prior, K, Dy};
  k_omgp_finalize<<<1, kClusterThreads, 0, as_stream(stream)>>>(a);
  return launch_rc();
}

int gpc_dense_natgrad(const double *grad_Lm, const double *mixgrad, const double *phi, const double *logphi, const double *ng_old,
                      const double *grad_old, long long N, int K, double *ng_out, double *grad_out, double *partials, void *stream) {
  if (!grad_Lm || !mixgrad || !phi || !logphi || !ng_out || !grad_out || !partials || N < 1 || K < 1 || N > 0x7fffffffLL) return GPC_EINVAL;
  if ((ng_old == nullptr) != (grad_old == nullptr)) return GPC_EINVAL;
  k_dense_natgrad<<<(unsigned)((N + 127) / 128), 128, 0, as_stream(stream)>>>(grad_Lm, mixgrad, phi, logphi, ng_old, grad_old, (int)N, K, ng_out,
                                                                             grad_out, partials);
  return launch_rc();
}

int gpc_cg_axpy(const double *x, const double *ng, const double *sd_old, double beta, double step, long long len, double *sd_new, double *x_new,
                void *stream) {
  if (!x || !ng || !sd_new || !x_new || len < 1) return GPC_EINVAL;
  k_cg_axpy<<<(unsigned)((len + 255) / 256), 256, 0, as_stream(stream)>>>(x, ng, sd_old, beta, step, len, sd_new, x_new);
  return launch_rc();
}

// ---- launch-sequence replay (CUDA graph of a batch of loop-body passes)
}  // extern "C"

// ---- the whole conjugate natural-gradient loop of MOHGP as one persistent cooperative kernel (gpc_loop_kernel.cuh)
namespace {
constexpr size_t kLoopRingBudget = 180000;
struct LoopGeom {
  int g_stages, s_stages;
  size_t smem;
};
LoopGeom loop_geometry(int KP, int DP, int D) {
  LoopGeom q;
  // Shared memory deliberately NOT maxed out: neither ring gains anything beyond nine stages (measured: G rows 520-521 us with
  // 10 .. 13 stages, 525 with 9; S rows flat from 8 up), while the posterior phase between them loses 5 us per evaluation when the
  // launch's carve-out leaves L1 only 28 KB.  180 KB of rings (+ header) selects the 196 KB configuration: ten stages each.
  const size_t budget = kLoopRingBudget;
  q.g_stages = natgrad2_stages(KP, DP, budget);
  {
    const size_t fixed = sizeof(double) * kSSoftmaxWarps + 64;
    const size_t per = sizeof(double) * (size_t)s_stage_doubles(KP, DP) + 3 * sizeof(uint64_t);
    const int st = (int)((budget - fixed) / per);
    q.s_stages = st > kMaxStages ? kMaxStages : st;
  }
  const size_t post = (((size_t)eig_ws_len(D) * 8 + 127) & ~(size_t)127) + sizeof(PostSmem);
  size_t m = natgrad2_smem(KP, DP, q.g_stages);
  if (step_stats_smem(KP, DP, q.s_stages) > m) m = step_stats_smem(KP, DP, q.s_stages);
  if (post > m) m = post;
  q.smem = kLoopHeaderBytes + ((m + 127) & ~(size_t)127);
  return q;
}
template <int KP>
int launch_loop(LoopKernelArgs &a, int D, int grid, cudaStream_t st) {
  const LoopGeom q = loop_geometry(KP, a.g.DP, D);
  if (q.smem > (size_t)kSmemBudget) return GPC_EINVAL;
  a.g.stages = q.g_stages;
  a.s.stages = q.s_stages;
  void *params[] = {&a};
  cudaError_t e;
  if (a.g.K == KP) {
    if ((e = ensure_dyn_smem(k_mohgp_loop<KP, true>, q.smem)) != cudaSuccess) return cuda_rc(e);
    e = cudaLaunchCooperativeKernel(reinterpret_cast<void *>(k_mohgp_loop<KP, true>), dim3(grid), dim3(kGThreads), params, q.smem, st);
  } else {
    if ((e = ensure_dyn_smem(k_mohgp_loop<KP, false>, q.smem)) != cudaSuccess) return cuda_rc(e);
    e = cudaLaunchCooperativeKernel(reinterpret_cast<void *>(k_mohgp_loop<KP, false>), dim3(grid), dim3(kGThreads), params, q.smem, st);
  }
  return e == cudaSuccess ? launch_rc() : cuda_rc(e);
}
}  // namespace

extern "C" {

long long gpc_xchg_doubles_fused(int world, int stats_len) {
  if (world < 1 || world > GPC_MAX_PEERS || stats_len < 1) return -1;
  const long long a = fx_doubles(world, stats_len), b = xchg_doubles(world, stats_len);
  return a > b ? a : b;  // one buffer serves both exchange layouts (they are never live at the same time)
}

int gpc_mohgp_loop(const double *F, const gpc_bufs_t *bufs, double *Wt, double *bias, double *gpart, double *spart, double *dots, double *beta,
                   double *stats, const double *eig_ws, const double *Sy_inv, const double *Sf, const double *const_term, double *muk_t,
                   double *Syi_ybark_t, double *per_k, double *scalars, double *mixgrad_out, int *info, double alpha, int prior, gpc_loop_t *loop,
                   const gpc_peers_t *peers, unsigned int *sync, unsigned long long *timing, int max_evals, long long N, int K, int D, int KP,
                   int DP, int grid, void *stream) {
  if (!F || !bufs_ok(bufs) || !Wt || !bias || !gpart || !spart || !dots || !beta || !stats || !eig_ws || !Sy_inv || !Sf || !const_term || !muk_t ||
      !per_k || !scalars || !info || !loop || !sync)
    return GPC_EINVAL;
  if (!geometry_ok(N, K, KP, DP, grid) || N > 0x7fffffffLL || D < 1 || D > DP || max_evals < 1) return GPC_EINVAL;
  if (prior != GPC_PRIOR_SYMMETRIC && prior != GPC_PRIOR_DP) return GPC_EINVAL;
  if (grid <= KP) return GPC_EINVAL;  // CTA k < KP computes the posterior of cluster k, CTA KP the scalar tail of the evaluation
  LoopKernelArgs a{};
  a.g.F = F;
  a.g.Wt = Wt;
  a.g.bias = bias;
  a.g.partials = gpart;
  a.g.dots_out = dots;
  a.g.N = (int)N;
  a.g.K = K;
  a.g.DP = DP;
  a.g.num_groups = (int)(gpc_padded_rows(N) / kGroupRows);
  a.g.bufs = *bufs;
  a.s.F = F;
  a.s.beta_out = beta;
  a.s.partials = spart;
  a.s.N = (int)N;
  a.s.K = K;
  a.s.DP = DP;
  a.s.num_groups = a.g.num_groups;
  a.s.bufs = *bufs;
  a.s.reverse = 0;  // no measurable gain from the reverse sweep at 125k .. 1M rows per GPU (and it changes the summation order)
  a.p = MohgpPostArgs{spart, grid, stats, eig_ws, Sy_inv, Sf, const_term, Wt, muk_t, Syi_ybark_t, per_k, bias, scalars, mixgrad_out,
                      sync + 2, info, alpha, prior, K, D, KP, DP, dots, beta, loop, GPC_PHASE_CONJ, nullptr};
  a.peers = peers;
  a.sync = sync;
  a.timing = timing;
  a.max_evals = max_evals;
  cudaStream_t st = as_stream(stream);
  switch (KP) {
    case 8: return launch_loop<8>(a, D, grid, st);
    case 16: return launch_loop<16>(a, D, grid, st);
    case 32: return launch_loop<32>(a, D, grid, st);
    case 64: return launch_loop<64>(a, D, grid, st);
  }
  return GPC_EINVAL;
}

int gpc_graph_begin(void *stream) {
  if (!stream) return GPC_EINVAL;  // the legacy default stream cannot be captured
  return cuda_rc(cudaStreamBeginCapture(as_stream(stream), cudaStreamCaptureModeRelaxed));
}

int gpc_graph_end(void *stream, void **exec) {
  if (!stream || !exec) return GPC_EINVAL;
  cudaGraph_t graph = nullptr;
  cudaError_t e = cudaStreamEndCapture(as_stream(stream), &graph);
  if (e != cudaSuccess || graph == nullptr) return e != cudaSuccess ? cuda_rc(e) : GPC_EINVAL;
  cudaGraphExec_t ge = nullptr;
  e = cudaGraphInstantiate(&ge, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) return cuda_rc(e);
  *exec = reinterpret_cast<void *>(ge);
  return 0;
}

int gpc_graph_launch(void *exec, void *stream) {
  if (!exec) return GPC_EINVAL;
  return cuda_rc(cudaGraphLaunch(reinterpret_cast<cudaGraphExec_t>(exec), as_stream(stream)));
}

int gpc_graph_destroy(void *exec) {
  if (!exec) return GPC_EINVAL;
  return cuda_rc(cudaGraphExecDestroy(reinterpret_cast<cudaGraphExec_t>(exec)));
}

}  // extern "C"
