/* gpclust_b200 -- C-ABI of the B200 (sm_100a) collapsed-VB hot path of GPclust.
 *
 * The reference (/root/reference/GPclust) is pure Python; its only backend seam is the L1 utility
 * module pair utilities.py (scipy.weave inline C++) / np_utilities.py (numpy), bound at import time by
 *   collapsed_mixture.py:6-10, MOHGP.py:9-12, MOG.py:6-9.
 * The entry points below are what a ctypes binding in those modules (and in the template methods
 * set_vb_param / do_computations / bound / vb_grad_natgrad, collapsed_vb.py:41-54) would call.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to IEEE double (row-major) unless stated otherwise;
 *   - the caller owns every buffer; the library allocates nothing and keeps no global state;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued asynchronously on it;
 *   - return value: 0 = enqueued, GPC_EINVAL = bad argument, negative = -(cudaError_t);
 *   - numerical failure (matrix not positive definite, the reference's LinAlgError raised by
 *     GPy.util.linalg.jitchol) is reported asynchronously through an `int *info` device flag
 *     (bit GPC_INFO_NOT_PD), which the host mirror turns into numpy.linalg.LinAlgError so that
 *     collapsed_vb.py:174,187 keep their meaning.
 *
 * Layout of the N-sized arrays ("series" rows n, "cluster" columns k): FRAGMENT-MAJOR, private to the library.
 *   Rows are padded to Npad = roundup(N, GPC_ROW_TILE) and handled in groups of 8; within a group
 *   - logits / natural gradient / search direction (KP = gpc_padded_k(K) in {8,16,32,64} columns): element
 *     (r, c) at ((c/8)*32 + r*4 + (c%8)/2)*2 + c%2 -- the mma.m8n8k4 C-fragment order; group = 8*KP doubles;
 *   - features F (DP = gpc_padded_d(D) columns, multiple of 8, <= 64): element (r, d) at (d/4)*32 + r*4 + d%4
 *     -- the A-fragment order; group = 8*DP doubles.  MOHGP: F = Y.  MOG: F = [x_i x_j (i<=j) | x_i] of the
 *     centred data (gpc_mog_features).
 *   One group of any operand is one contiguous block that cp.async.bulk lands in shared memory in lane order.
 *   gpc_pack_rows / gpc_unpack_rows convert from / to contiguous row-major; pad rows and columns hold zeros.
 *   - stats vector: [ T (KP x DP) | phi_hat (KP) | H (1) | pad (1) ],  T[k][f] = sum_n phi[n][k] F[n][f].
 */
#ifndef GPCLUST_B200_H
#define GPCLUST_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define GPC_ROW_TILE 64
#define GPC_EINVAL 2
#define GPC_INFO_NOT_PD 1
#define GPC_INFO_PEER_TIMEOUT 2 /* a peer of the NVLink exchange never raised its flag (distinct from a numerical failure) */

#define GPC_PRIOR_SYMMETRIC 0
#define GPC_PRIOR_DP 1

#define GPC_BETA_ZERO 0 /* steepest / first iteration  collapsed_vb.py:157-158 */
#define GPC_BETA_HS 1   /* collapsed_vb.py:163-164 */
#define GPC_BETA_PR 2   /* collapsed_vb.py:159-160 */
#define GPC_BETA_FR 3   /* collapsed_vb.py:161-162 */

#define GPC_LAYOUT_FEATURES 0 /* fragment-major feature rows F (DMMA A-fragment order)      */
#define GPC_LAYOUT_LOGITS 1   /* fragment-major N x K arrays (DMMA C-fragment order)          */

#define GPC_KERN_RBF 0
#define GPC_KERN_MATERN32 1
#define GPC_KERN_MATERN52 2
#define GPC_KERN_BIAS 3
#define GPC_KERN_WHITE 4

/* ---- device-resident loop state of the conjugate natural-gradient ascent (collapsed_vb.py:143-303) ----
 * With a gpc_loop_t in device memory the accept / reject decision (collapsed_vb.py:182-193), the buffer rotation,
 * the iteration count and the termination tests (collapsed_vb.py:229-291) are taken ON THE DEVICE by the kernel
 * that finishes a bound evaluation, and every kernel of the sequence
 *     gpc_loop_natgrad -> gpc_loop_step_stats(CONJ) -> [reduce, all-reduce] -> gpc_loop_*_post(CONJ)
 *                      -> gpc_loop_step_stats(RETRY) -> [reduce, all-reduce] -> gpc_loop_*_post(RETRY)
 * reads it to pick its buffers or to return at once (`done`, or RETRY without `need_retry`).  The host enqueues a
 * batch of such sequences without synchronising and reads the trace back once per batch.
 * slot_mode = 1 shortens the sequence to ONE bound evaluation per slot,
 *     gpc_loop_natgrad -> gpc_loop_step_stats -> [reduce / push] -> gpc_loop_*_post          (`phase` arguments ignored)
 * where a slot that follows a rejected conjugate step (need_retry) skips the G pass (same point, same gradient) and
 * runs the other two in the RETRY phase: no launches that return at once after every accepted step.            */
#define GPC_PHASE_CONJ 0
#define GPC_PHASE_RETRY 1
#define GPC_DONE_FTOL 1
#define GPC_DONE_GTOL 2
#define GPC_DONE_MAXITER 4
#define GPC_DONE_FAULT 8
#define GPC_DONE_BUDGET 16
#define GPC_DONE_PEER_TIMEOUT 32 /* sticky: a peer exchange timed out on this rank; every later kernel returns at once and the
                                    host raises a hard error (never folded into a numerical outcome) */
typedef struct {
  double bound_old; /* bound at the accepted point                                   */
  double sq_old;    /* natgrad.grad of the previous iteration (PR / FR)              */
  double ftol, gtol, maxiter, step;
  int iteration;    /* bound evaluations so far (collapsed_vb.py:177,191)            */
  int method;       /* GPC_BETA_*                                                    */
  int first;        /* 1 until the first step has been taken (beta = 0, no history)  */
  int xi, xprev, xtrial, ngi; /* buffer selectors; xprev = -1: no previous point     */
  int need_retry;   /* the conjugate step lowered the bound: the RETRY kernels run   */
  int done;         /* GPC_DONE_* bits; once set every later kernel returns at once  */
  int n_trace, trace_cap;
  int slot_mode;    /* 1: every kernel takes its phase from need_retry (see below)   */
  int pass_budget;  /* > 0: accepted steps left before GPC_DONE_BUDGET is raised     */
  int pad;
  /* followed by trace_cap rows [iteration, bound, squareNorm, beta] of doubles      */
} gpc_loop_t;
typedef struct {
  double *x[3];   /* logits: accepted / previous accepted / trial, rotating            */
  double *lse[3]; /* row logsumexp of the matching logits buffer                        */
  double *ng[2];  /* natural gradient: current / previous                               */
  double *sd;     /* search direction (updated in place)                                */
} gpc_bufs_t;

/* ---- peer-memory exchange (rows sharded over the GPUs of one NVSwitch box) -------------------------------
 * Instead of an NCCL all-reduce between kernels, the kernel that PRODUCES a small replicated vector (the reduced
 * statistics, the three CG dots) stores it straight into every peer's inbox over NVLink and raises a flag; the
 * kernel that CONSUMES it waits for the W flags and sums the W contributions in rank order (bit-identical on all
 * ranks).  inbox[r] is rank r's exchange buffer mapped into this process (symmetric memory / CUDA IPC); its layout
 * for W ranks and statistics length L (gpc_xchg_doubles gives the size):
 *   [2][W][L2] statistics | [2][W][4] dots | [2][W] statistics flags (u64) | [2][W] dots flags (u64),  L2 = L rounded up to 2
 * Two slots: a rank can be at most one exchange ahead of the slowest.  `epochs` (2 local u64: statistics, dots)
 * count the exchanges this rank has pushed.                                                                   */
#define GPC_MAX_PEERS 8
typedef struct {
  int world, rank, stats_len, pad;
  double *inbox[GPC_MAX_PEERS];
  unsigned long long *epochs; /* local: [0] statistics exchanges pushed, [1] dots exchanges pushed */
} gpc_peers_t;
long long gpc_xchg_doubles(int world, int stats_len);

/* ---- library / geometry queries (host only, no GPU needed) ---------------------------------- */
int gpc_abi_version(void);
int gpc_padded_k(int K);           /* 8,16,32,64, then multiples of 64 up to 512; -1 when K < 1 or K > 512 */
int gpc_padded_d(int D);           /* roundup(D, 8);  -1 when D < 1 or D > 64                      */
long long gpc_padded_rows(long long N);
int gpc_stats_len(int KP, int DP); /* KP*DP + KP + 2 (one pad double keeps blocks 16-B aligned)    */
int gpc_default_grid(int device);  /* number of persistent CTAs (1 per SM); <0 on CUDA error          */

/* ---- G pass: natural gradient + conjugate-gradient dot products ------------------------------
 * Replaces MOHGP.vb_grad_natgrad (MOHGP.py:137-153) / MOG.vb_grad_natgrad (MOG.py:72-84) including
 * multiple_mahalanobis (utilities.py:108-172) and the dots of collapsed_vb.py:154,160-164.
 *   a[n,k]   = F[n,:] . Wt[k,:] + bias[k] - x[n,k]
 *   natgrad  = a - sum_k phi a ;  grad = natgrad * phi ;  phi = softmax_k(x)
 *   dots[0] = natgrad.grad  dots[1] = (natgrad - ng_old).grad  dots[2] = (natgrad - ng_old).grad_old
 * lse / lse_old: Npad row-wise logsumexp of x / x_old, as left behind by gpc_step_stats (phi = exp(x - lse)).
 * x_old / lse_old / ng_old may all be NULL (first iteration): dots[1], dots[2] are then 0.
 * grad_out may be NULL.  partials: grid*4 doubles.  counter: one zero-initialised unsigned int.     */
int gpc_natgrad(const double *F, const double *x, const double *lse, const double *Wt, const double *bias, const double *x_old,
                const double *lse_old, const double *ng_old, double *ng_out, double *grad_out, double *partials, unsigned int *counter,
                double *dots, long long N, int K, int KP, int DP, int grid, void *stream);

/* ---- S pass: conjugate step + softmax + sufficient statistics --------------------------------
 * Replaces collapsed_vb.py:167,172 (searchDir, phi_old + step*searchDir) fused with
 * CollapsedMixture.set_vb_param (collapsed_mixture.py:49-59: softmax, phi_hat, H) and the N-sized part
 * of do_computations (MOHGP.py:76 ybark; MOG.py:56-57 Xsumk, Ck).
 *   beta from dots (device) according to beta_mode, nan / negative -> 0 (collapsed_vb.py:165-166)
 *   sd_new = ng + beta*sd_old ; x_new = x_base + step*sd_new ; stats of softmax(x_new)
 * lse_new (Npad, nullable) receives the row-wise logsumexp of x_new, the G pass needs it.
 * ng == NULL: "set" mode -- statistics (and lse) of x_base itself, no N x K array is written.
 * sd_new may alias sd_old.  partials: grid * gpc_stats_len doubles -> reduce with gpc_reduce_partials. */
int gpc_step_stats(const double *F, const double *x_base, const double *ng, const double *sd_old, double *sd_new, double *x_new,
                   double *lse_new, const double *dots, double sq_old, double *beta_out, double *partials, double step, int beta_mode, long long N,
                   int K, int KP, int DP, int grid, void *stream);

/* the same two passes driven by a device-resident gpc_loop_t (see above); `bufs` is read on the host.        */
int gpc_loop_natgrad(const double *F, const gpc_bufs_t *bufs, const double *Wt, const double *bias, double *partials, unsigned int *counter,
                     double *dots, const gpc_loop_t *loop, long long N, int K, int KP, int DP, int grid, void *stream);
int gpc_loop_step_stats(const double *F, const gpc_bufs_t *bufs, const double *dots, double *beta_out, double *partials,
                        const gpc_loop_t *loop, int phase, long long N, int K, int KP, int DP, int grid, void *stream);

/* peer-exchange variants: `peers` is a DEVICE pointer to a gpc_peers_t (NULL = single GPU / NCCL outside)    */
int gpc_loop_natgrad_px(const double *F, const gpc_bufs_t *bufs, const double *Wt, const double *bias, double *partials, unsigned int *counter,
                        double *dots, const gpc_loop_t *loop, const gpc_peers_t *peers, long long N, int K, int KP, int DP, int grid,
                        void *stream);
int gpc_loop_step_stats_px(const double *F, const gpc_bufs_t *bufs, double *dots, double *beta_out, double *partials, const gpc_loop_t *loop,
                           const gpc_peers_t *peers, int phase, long long N, int K, int KP, int DP, int grid, void *stream);
/* fixed-order reduction of the per-CTA statistics, pushed into every peer's inbox (+ flag)                  */
int gpc_reduce_push(const double *partials, int num_parts, int len, const gpc_peers_t *peers, unsigned int *counter, const gpc_loop_t *loop,
                    int phase, void *stream);

/* ---- the whole loop as ONE persistent cooperative launch (MOHGP, K <= 64) ------------------------------------------------
 * Replaces the per-evaluation launch sequence gpc_loop_natgrad_px -> gpc_loop_step_stats_px -> gpc_reduce_push -> gpc_loop_mohgp_post_px
 * of collapsed_vb.py:147-303 + MOHGP.py:70-153: one launch runs up to `max_evals` bound evaluations (or until the device-side
 * termination tests of gpc_loop_t fire); the phases are separated by grid-wide barriers instead of kernel boundaries, the statistics
 * are reduced per cluster by the CTA that computes that cluster's posterior, and with rows sharded over GPUs each cluster's row goes
 * straight into every peer's inbox with its own flag.  Arguments as for the entry points it replaces.
 *   grid      = gpc_default_grid(device) (one CTA per SM, all co-resident: cooperative launch); must be > KP (CTA k: cluster k; CTA KP:
 *             entropy, mixing-proportion terms, bound and the accept / reject decision)
 *   gpart     grid*4 doubles, spart grid*gpc_stats_len(KP,DP) doubles
 *   dots/beta device scalars [3] / [1] (read back by the host with the control block)
 *   sync      4 zero-initialised unsigned ints (the kernel leaves them zero)
 *   timing    nullable: 16 unsigned 64-bit accumulators (measured by CTA 0 with %globaltimer) -- [0..3] ns in G rows / dots barrier +
 *             exchange / S rows / posterior incl. barriers, [4] evaluations, [5] G phases, [6..10] the posterior phase in detail:
 *             wait at the barrier / local row reduction / NVLink exchange / posterior algebra + ticket / tail + barrier
 *   peers     NULL on one GPU; otherwise the inbox must hold gpc_xchg_doubles_fused(world, stats_len) doubles:
 *             [2][W][L] statistics | [2][W][4] dots | [2][W][64] statistics flags (sender, cluster) | [2][W] dots flags
 * A peer that never arrives sets GPC_DONE_PEER_TIMEOUT in loop->done (bounded waits, 4 s).                                      */
long long gpc_xchg_doubles_fused(int world, int stats_len);
int gpc_mohgp_loop(const double *F, const gpc_bufs_t *bufs, double *Wt, double *bias, double *gpart, double *spart, double *dots, double *beta,
                   double *stats, const double *eig_ws, const double *Sy_inv, const double *Sf, const double *const_term, double *muk_t,
                   double *Syi_ybark_t, double *per_k, double *scalars, double *mixgrad_out, int *info, double alpha, int prior, gpc_loop_t *loop,
                   const gpc_peers_t *peers, unsigned int *sync, unsigned long long *timing, int max_evals, long long N, int K, int D, int KP,
                   int DP, int grid, void *stream);

/* ---- K > 64: the clusters are processed in chunks of 64 (one launch of the tensor kernels per chunk) ----------
 * N x K arrays keep the fragment-major layout with KP/8 n-tiles per 8-row group (xg_stride = 8*KP doubles between
 * groups); a chunk's slice starts 512*c doubles into the group.  Pointers named x / a_out / Wt / bias are PRE-OFFSET to
 * the chunk by the caller (x + 512*c, Wt + 64*c*DP, bias + 64*c).
 * S pass = gpc_step_lse (CG step on whole rows, row logsumexp, entropy partials; collapsed_vb.py:167,172 +
 *          np_utilities.py:25-40) followed by gpc_stats_chunk per chunk (phi = exp(x - lse), T_c = phi_c^T F, phi_hat_c).
 * G pass = gpc_natgrad_chunk per chunk (a_c = F.W_c + bias_c - x_c into the natural-gradient buffer, sa[row] += sum phi*a)
 *          followed by gpc_natgrad_finish (natgrad = a - sa, gradient, the CG dots).                                    */
int gpc_natgrad_chunk(const double *F, const double *x, const double *lse, const double *Wt, const double *bias, double *a_out, double *sa_vec,
                      int first, double *partials, unsigned int *counter, double *dots_scratch, long long N, int Kc, int DP, long long xg_stride,
                      int grid, void *stream);
int gpc_natgrad_finish(double *ng, const double *x, const double *lse, const double *sa_vec, const double *x_old, const double *lse_old,
                       const double *ng_old, double *grad_out, double *partials, unsigned int *counter, double *dots, long long N, int K, int KP,
                       int grid, void *stream);
int gpc_step_lse(const double *x_base, const double *ng, const double *sd_old, double *sd_new, double *x_new, double *lse_new, const double *dots,
                 double sq_old, double *beta_out, double *hpart, double step, int beta_mode, long long N, int K, int KP, int grid, void *stream);
int gpc_stats_chunk(const double *F, const double *x, const double *lse, const double *h_in, double *partials, long long part_stride, int t_off,
                    int ph_off, int h_off, int write_h, long long N, int Kc, int DP, long long xg_stride, int grid, void *stream);

/* ---- wide mode: K > 64 in ONE launch per pass ---------------------------------------------------
 * The kernels run as thread-block clusters of KP/64 CTAs: the CTAs of a cluster walk the same 8-row groups, CTA r owning the
 * clusters [64r, 64r+64) (its own slice of Wt / bias / the N x K operands, F re-read from L2), and the row-coupled scalars --
 * (max, sum e, sum e*(x-max)) of the softmax (np_utilities.py:25-40) and sum_k phi*a of the natural gradient
 * (MOHGP.py:150) -- are exchanged through each other's shared memory.  Same arguments and meaning as gpc_natgrad /
 * gpc_step_stats with KP = gpc_padded_k(K) in (64, 512]; N x K operands in the fragment-major layout of width KP.
 * nclusters = gpc_wide_clusters(KP, DP) (co-resident clusters on this device); `partials` of the S pass holds nclusters
 * blocks of gpc_stats_len(KP, DP) doubles, of the G pass nclusters*(KP/64)*4 doubles.                            */
int gpc_wide_clusters(int KP, int DP);
int gpc_natgrad_wide(const double *F, const double *x, const double *lse, const double *Wt, const double *bias, const double *x_old,
                     const double *lse_old, const double *ng_old, double *ng_out, double *grad_out, double *partials, unsigned int *counter,
                     double *dots, long long N, int K, int KP, int DP, int nclusters, void *stream);
int gpc_step_stats_wide(const double *F, const double *x_base, const double *ng, const double *sd_old, double *sd_new, double *x_new,
                        double *lse_new, const double *dots, double sq_old, double *beta_out, double *partials, double step, int beta_mode,
                        long long N, int K, int KP, int DP, int nclusters, void *stream);
/* the wide passes driven by a device-resident gpc_loop_t (as gpc_loop_natgrad_px / gpc_loop_step_stats_px)       */
int gpc_loop_natgrad_wide(const double *F, const gpc_bufs_t *bufs, const double *Wt, const double *bias, double *partials, unsigned int *counter,
                          double *dots, const gpc_loop_t *loop, const gpc_peers_t *peers, long long N, int K, int KP, int DP, int nclusters,
                          void *stream);
int gpc_loop_step_stats_wide(const double *F, const gpc_bufs_t *bufs, double *dots, double *beta_out, double *partials, const gpc_loop_t *loop,
                             const gpc_peers_t *peers, int phase, long long N, int K, int KP, int DP, int nclusters, void *stream);

/* out[j] = sum_c partials[c*len + j], c in fixed order (bit-reproducible).                          */
int gpc_reduce_partials(const double *partials, int num_parts, int len, double *out, void *stream);

/* ---- MOHGP per-cluster posterior (MOHGP.py:79-90) + bound (MOHGP.py:124-135) ------------------ */
/* once per kernel-parameter change (MOHGP.py:47-49 / :61-63): Ly = chol(Sy+1e-6 I), Ly_inv, Sy_inv,
 * A = Ly^-1 Sf Ly^-T, Sy_logdet.  D <= 64.                                                         */
int gpc_mohgp_setup(const double *Sf, const double *Sy, double *Ly, double *Ly_inv, double *Sy_inv, double *A, double *Sy_logdet,
                    int *info, int D, void *stream);
/* const_term = -0.5*(n_total*D*log(2 pi) + n_total*Sy_logdet + sum(YTY o Sy_inv))  (MOHGP.py:129-132) */
int gpc_mohgp_const(const double *YTY, const double *Sy_inv, const double *Sy_logdet, double n_total, int D, double *const_term,
                    void *stream);
/* per cluster k (one CTA each): C_k = I + phi_hat_k A, jitchol, log_det_diff, Lambda_inv, s_k, mu_k,
 * and the G-pass weights Wt[k,:] = Sy^-1 mu_k.  per_k: KP x 4 = [log_det_diff, s^T Lambda_inv s,
 * tr(Lambda_inv Sy_inv), mu^T Sy^-1 mu].  Syi_ybark_t, Lambda_inv, C_chols may be NULL.  *info is
 * cleared first.                                                                                     */
int gpc_mohgp_cluster(const double *stats, const double *A, const double *Ly, const double *Ly_inv, const double *Sy,
                      const double *Sy_inv, const double *Sf, double *Wt, double *muk_t, double *per_k, double *Syi_ybark_t, double *Lambda_inv, double *C_chols, int *info,
                      int K, int D, int KP, int DP, void *stream);
/* ---- the same posterior WITHOUT per-cluster factorisations (the hot-loop form) ----
 * once per kernel-parameter change: A = Q diag(lam) Q^T by one-CTA cyclic Jacobi; ws (gpc_mohgp_eig_ws_len(D)
 * doubles) receives lam, U = Ly^-T Q, U^T, (Ly Q)^T, tr(Sy_inv), sum(Sf o Sy_inv).                             */
int gpc_mohgp_eig_ws_len(int D);
int gpc_mohgp_eig(const double *A, const double *Ly, const double *Ly_inv, const double *Sy_inv, const double *Sf, int D, double *ws,
                  void *stream);
/* per bound evaluation, ONE launch (KP CTAs): fixed-order reduction of this cluster's statistics over `num_parts`
 * per-CTA partial vectors of gpc_step_stats (num_parts = 1 and partials == stats after an all-reduce), posterior
 * mean / G-pass weights / per-cluster scalars in the eigenbasis (MOHGP.py:79-90), and -- in the last CTA to
 * finish -- the mixing-proportion terms, the bound (MOHGP.py:124-135) and the G-pass bias.  Outputs as
 * gpc_mohgp_cluster + gpc_mohgp_finalize.  counter: one zero-initialised unsigned int.  *info is cleared first. */
int gpc_mohgp_post(const double *partials, int num_parts, double *stats, const double *eig_ws, const double *Sy_inv, const double *Sf,
                   const double *const_term, double *Wt, double *muk_t, double *Syi_ybark_t, double *per_k, double *bias, double *scalars,
                   double *mixgrad_out, unsigned int *counter, int *info, double alpha, int prior, int K, int D, int KP, int DP, void *stream);
/* ---- kernel hyper-parameter gradients of MOHGP (MOHGP.update_kern_grads, MOHGP.py:92-122) on the device, in the eigenbasis of A:
 * dL_dSf and dL_dSy (D x D each) from the statistics and the posterior of the current point -- O(K D^2) + two D x K x D products
 * instead of the reference's K triangular solves and K dense D^3 products.  stats / Syi_ybark_t as left by gpc_mohgp_post;
 * work: gpc_mohgp_kern_grads_work(K, D) doubles.  *info is cleared first (GPC_INFO_NOT_PD as in gpc_mohgp_post).
 * The host applies GPy's update_gradients_full (D x D contractions with dK/dtheta) to the two matrices.                     */
long long gpc_mohgp_kern_grads_work(int K, int D);
int gpc_mohgp_kern_grads(const double *stats, const double *eig_ws, const double *Sf, const double *Sy_inv, const double *YTY, const double *Syi_ybark_t,
                         double n_total, int K, int D, int KP, int DP, double *work, double *dL_dSf, double *dL_dSy, int *info, void *stream);
/* gpc_mohgp_post + the device-side accept / reject / termination step (dots, beta: as written by the two passes) */
int gpc_loop_mohgp_post(const double *partials, int num_parts, double *stats, const double *eig_ws, const double *Sy_inv, const double *Sf,
                        const double *const_term, double *Wt, double *muk_t, double *Syi_ybark_t, double *per_k, double *bias, double *scalars,
                        double *mixgrad_out, unsigned int *counter, int *info, double alpha, int prior, int K, int D, int KP, int DP,
                        const double *dots, const double *beta, gpc_loop_t *loop, int phase, void *stream);
/* as gpc_loop_mohgp_post, the statistics coming from the W inbox vectors of the peer exchange                  */
int gpc_loop_mohgp_post_px(const gpc_peers_t *peers, int world, double *stats, const double *eig_ws, const double *Sy_inv, const double *Sf,
                           const double *const_term, double *Wt, double *muk_t, double *Syi_ybark_t, double *per_k, double *bias,
                           double *scalars, double *mixgrad_out, unsigned int *counter, int *info, double alpha, int prior, int K, int D,
                           int KP, int DP, const double *dots, const double *beta, gpc_loop_t *loop, int phase, void *stream);
/* mixing-proportion bound and gradient (collapsed_mixture.py:66-94), bound assembly, G-pass bias.
 * scalars: [0] bound, [1] mixing_prop_bound, [2] H.  mixgrad_out (K doubles) may be NULL.             */
int gpc_mohgp_finalize(const double *stats, const double *per_k, const double *const_term, double *bias, double *scalars,
                       double *mixgrad_out, double alpha, int prior, int K, int KP, int DP, void *stream);

/* ---- MOG (MOG.py:34-84) ------------------------------------------------------------------------ */
/* F (fragment-major, Npad x DP) <- [xc_i xc_j (i<=j) | xc_i | 0], xc = x - centre                        */
int gpc_mog_features(const double *X, const double *centre, long long N, int D, int DP, double *F, void *stream);
/* per_k: KP x 4 = [kN, vN, Sns_halflogdet, grad const].  mun_c: K x D (centred), Sns / Sns_inv: K x D x D.
 * D <= 10.  *info is cleared first.                                                                   */
int gpc_mog_cluster(const double *stats, const double *m0c, const double *S0, double *Wt, double *per_k, double *mun_c, double *Sns,
                    double *Sns_inv, int *info, double k0, double v0, int K, int D, int KP, int DP, void *stream);
int gpc_mog_finalize(const double *stats, const double *per_k, double *bias, double *scalars, double *mixgrad_out, double n_total,
                     double k0, double v0, double S0_halflogdet, double alpha, int prior, int K, int KP, int DP, int D, void *stream);

/* gpc_mog_cluster + gpc_mog_finalize driven by a gpc_loop_t (skip logic + device-side decision)            */
int gpc_loop_mog_post(const double *stats, const double *m0c, const double *S0, double *Wt, double *per_k, double *mun_c, double *Sns,
                      double *Sns_inv, int *info, double k0, double v0, double *bias, double *scalars, double *mixgrad_out, double n_total,
                      double S0_halflogdet, double alpha, int prior, int K, int D, int KP, int DP, const double *dots, const double *beta,
                      gpc_loop_t *loop, int phase, void *stream);

/* ---- covariance fills (GPy kern.K(X, X2); call sites MOHGP.py:47-48,61-62, OMGP.py:104,132) ---- */
/* out (n x m) = sum_p k_p(X_i, X2_j) ; X2 == NULL -> X2 = X, m = n (White contributes, diagonal r = 0).
 * kinds/variances/lengthscales: HOST arrays of nparts (<= 8) entries.  diag_add is added to the
 * diagonal when X2 == NULL.                                                                          */
int gpc_kern_fill(const double *X, const double *X2, int n, int m, int q, int nparts, const int *kinds, const double *variances,
                  const double *lengthscales, double diag_add, double *out, void *stream);

/* ---- OMGP (OMGP.py:96-147): dense N x N work per component ------------------------------------
 * Matrices are Npad x Npad row-major, Npad = gpc_omgp_padded_n(N) (multiple of 64; pad = identity).
 * Per component i the host mirror calls build -> chol -> tri_inverse -> component; then once finalize.   */
long long gpc_omgp_padded_n(long long N);
int gpc_omgp_slabs(long long Npad);  /* row slabs of the column statistics: part holds slabs*Npad*gpc_omgp_part_stride() doubles */
int gpc_omgp_part_stride(void);
/* B = K(X,X) + diag(1/((w + eps)/variance) + jitter)  (OMGP.py:104-105,132-135; eps = 0 for :158).  X: N x q,
 * w: phi column with stride w_stride.  kinds/variances/lengthscales: HOST arrays as in gpc_kern_fill.      */
int gpc_omgp_build(const double *X, long long N, int q, int nparts, const int *kinds, const double *variances, const double *lengthscales,
                   const double *w, int w_stride, double variance, double eps, double jitter, double *B, void *stream);
/* in-place blocked lower Cholesky (the dpotrf inside GPy pdinv/jitchol): diagonal-tile inverses to dinv
 * (Npad x 64), per-tile sum(log diag) to ldsum (Npad/64).  *info is cleared first; GPC_INFO_NOT_PD on a bad
 * pivot (the host retries with jitter = mean(diag)*1e-6*10^i, GPy jitchol).                               */
int gpc_chol_blocked(double *A, long long Npad, double *dinv, double *ldsum, int *info, void *stream);
/* W = L^-1 (lower triangular, strict upper zero) -- the dtrtri of GPy pdinv                                */
int gpc_tri_inverse_blocked(const double *L, long long Npad, const double *dinv, double *W, void *stream);
/* gpc_chol_blocked + gpc_tri_inverse_blocked pipelined over two streams (same results): step j of the inverse starts as soon as
 * the panel of step j is factorised and runs underneath the trailing updates.  `aux_stream` is a second stream of the caller; it is
 * forked from / joined back into `stream`, so the call behaves like one launch sequence on `stream` (and can be captured).    */
int gpc_chol_inverse(double *A, long long Npad, double *dinv, double *ldsum, int *info, double *W, void *stream, void *aux_stream);
/* comp[0] = tr(B^-1 Y Y^T), comp[1] = logdet B, comp[2] = sum_n phi_ni; grad_col (stride grad_stride) =
 * -0.5 variance (sum_d alpha^2 - diag(B^-1)) / (phi^2 + 1e-6)  (OMGP.py:113,117,121,138-140).
 * Y: Npad x Dy (pad rows zero), Dy <= 8.  tbuf: Npad x Dy.  part: see gpc_omgp_slabs.                       */
int gpc_omgp_component(const double *W, const double *Y, const double *ldsum, const double *phi_col, int phi_stride, long long N, int Dy,
                       double variance, double *tbuf, double *part, double *grad_col, int grad_stride, double *comp, void *stream);
/* Kernel hyper-parameter gradients of one component (OMGP.update_kern_grads, OMGP.py:55-94).  W = L^-1 of
 * B = K + diag(variance / phi) (built with eps = 0: OMGP.py:63 has no 1e-6).  With dL_dB = alpha alpha^T - B^-1:
 *   out[2p]   = sum_ij dL_dB_ij k_p(x_i,x_j) / variance_p         (-> GPy variance.gradient = 0.5 * out[2p])
 *   out[2p+1] = sum_ij dL_dB_ij dk_p/dr r_ij                      (-> lengthscale.gradient = -0.5 * out[2p+1] / lengthscale_p)
 * for part p of the covariance (White: trace; Bias: plain sum), and dldb_diag[j] = diag(dL_dB) (OMGP.py:88-91).
 * B^-1 = W^T W is formed tile by tile on the FP64 tensor pipe and contracted in the epilogue; nothing N x N is stored.
 * Scratch: tbuf Npad x Dy, part as gpc_omgp_component, alpha Npad x Dy, dldb_diag Npad,
 * tile_part gpc_omgp_kgrad_tiles(Npad) x 16, out 16 doubles.                                                */
long long gpc_omgp_kgrad_tiles(long long Npad);
int gpc_omgp_kern_grads(const double *W, const double *X, int q, const double *Y, long long N, int Dy, int nparts, const int *kinds,
                        const double *variances, const double *lengthscales, double *tbuf, double *part, double *alpha, double *dldb_diag,
                        double *tile_part, double *out, void *stream);
/* bound (OMGP.py:96-123) from comp (K x 4) + entropy partials; phi_hat, mixing-proportion gradient (K each) */
int gpc_omgp_finalize(const double *comp, const double *Hpart, int n_hpart, double *scalars, double *phi_hat, double *mixgrad, double variance,
                      double alpha, int prior, int K, int Dy, void *stream);
/* OMGP.py:142-147 + the CG dots of collapsed_vb.py:154,160-164 on contiguous N x K arrays; ng_old / grad_old
 * nullable together.  partials: ceil(N/128) x 4 -> gpc_reduce_partials(len = 4).                              */
int gpc_dense_natgrad(const double *grad_Lm, const double *mixgrad, const double *phi, const double *logphi, const double *ng_old,
                      const double *grad_old, long long N, int K, double *ng_out, double *grad_out, double *partials, void *stream);
/* sd_new = ng + beta*sd_old ; x_new = x + step*sd_new  (collapsed_vb.py:167,172); sd_old nullable         */
int gpc_cg_axpy(const double *x, const double *ng, const double *sd_old, double beta, double step, long long len, double *sd_new, double *x_new,
                void *stream);

/* ---- the reference's L1 utilities, same meaning as np_utilities.py ---------------------------- */
/* softmax (np_utilities.py:25-40): x is N x K with leading dimension ldx; phi / logphi (either may be
 * NULL) get leading dimension ldo; Hpart must hold `grid` doubles, H = sum(Hpart) (gpc_reduce_partials). */
int gpc_softmax_rows(const double *x, long long N, int K, int ldx, double *phi, double *logphi, int ldo, double *Hpart, int grid,
                     void *stream);
/* multiple_mahalanobis (utilities.py:108-172): out[n][m] = |L^-1 (X1_n - X2_m)|^2, D <= 128.         */
int gpc_multiple_mahalanobis(const double *X1, const double *X2, const double *L, long long N1, int N2, int D, double *out,
                             void *stream);
/* multiple_pdinv (np_utilities.py:6-22): A, inv are D x D x K (K last); hld K doubles.  D <= 64.     */
int gpc_multiple_pdinv(const double *A, int D, int K, double *inv, double *hld, int *info, void *stream);
/* Y^T Y (MOHGP.py:53) of the feature rows F (fragment-major): partials grid x DP x DP, then
 * gpc_reduce_partials(len = DP*DP).                                                                    */
int gpc_gram_partials(const double *F, long long Npad, int DP, double *partials, int grid, void *stream);
/* ---- ingest helpers (drosophila.ipynb cells [2], [4]) ------------------------------------------------- */
/* dst = (src - rowmean) / rowstd  (population std), N x D row-major                                   */
int gpc_normalise_rows(const double *src, long long N, int D, double *dst, void *stream);
/* score[n] = std_t(mean over replicates) / mean_t(std over replicates); group[d] = time-point index of column d (device int array), G <= 64 */
int gpc_replicate_scores(const double *Y, long long N, int D, const int *group, int G, double *score, void *stream);

/* contiguous N x C row-major  <->  the fragment-major device layout (Npad rows, CP columns, pad = 0)   */
int gpc_pack_rows(const double *src, long long N, int C, int CP, int layout, double *dst, void *stream);
int gpc_unpack_rows(const double *src, long long N, int C, int CP, int layout, double *dst, void *stream);
/* layout helpers: contiguous N x D  <->  row-major padded Npad x ld                                  */
int gpc_pad_rows(const double *src, long long N, int D, int DP, double *dst, void *stream);
int gpc_unpad_rows(const double *src, long long N, int K, int ld, double *dst, void *stream);

/* ---- launch-sequence replay -------------------------------------------------------------------
 * The device-resident loop (gpc_loop_t above) makes every launch of a loop-body pass argument-invariant, so a batch of
 * passes is captured once into a CUDA graph and replayed with a single launch (no per-kernel host work).
 * gpc_graph_begin starts a (relaxed-mode) capture on `stream` (must not be the legacy default stream); every gpc_* launch
 * issued on that stream until gpc_graph_end is recorded instead of executed.  *exec is an opaque handle.       */
int gpc_graph_begin(void *stream);
int gpc_graph_end(void *stream, void **exec);
int gpc_graph_launch(void *exec, void *stream);
int gpc_graph_destroy(void *exec);

#ifdef __cplusplus
}
#endif
#endif
