/* gpo_b200.h -- C ABI of the B200-native GP-posterior + acquisition engine.
 *
 * Drop-in boundary for ONE path of SheffieldML/GPyOpt: exact-GP posterior prediction fused with the
 * acquisition value/gradient over very large candidate sets.  The reference is pure Python; each entry
 * point below names the reference method whose arithmetic it replaces (paths relative to /root/reference).
 * A maintainer binds these with ctypes; no torch / C++ types cross the boundary.
 *
 * Conventions
 *   - every array is float64, C-contiguous; "candidate" arrays are row-major [N, d]
 *   - pointer arguments marked (host|device) may be either: the library inspects them with
 *     cudaPointerGetAttributes and stages host memory through its own pinned buffers
 *   - `stream` is a cudaStream_t (NULL = legacy default stream).  With device pointers the call is
 *     asynchronous on `stream`; with host pointers it returns after the results are in host memory
 *   - every function returns 0 (GPO_OK) or a negative gpo_status; never throws; one engine per thread
 *   - S = number of hyper-parameter sets held by the engine: 1 for GPModel, #HMC samples for GPModel_MCMC
 */
#ifndef GPO_B200_H
#define GPO_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gpo_engine gpo_engine;

typedef enum {
  GPO_OK = 0,
  GPO_ERR_CUDA = -1,      /* CUDA runtime/driver failure; text via gpo_last_error */
  GPO_ERR_NOT_PD = -2,    /* Gram matrix not positive definite even with jitter -> numpy.linalg.LinAlgError
                             (caught by GPyOpt/core/bo.py:143) */
  GPO_ERR_ARG = -3,       /* bad argument / unsupported size */
  GPO_ERR_STATE = -4,     /* call order: predict before fit, ... */
  GPO_ERR_NO_DEVICE = -5  /* no usable sm_100 GPU: the engine has no CPU fallback */
} gpo_status;

typedef enum { GPO_KERN_RBF = 0, GPO_KERN_MATERN52 = 1, GPO_KERN_MATERN32 = 2 } gpo_kernel;

typedef enum {
  GPO_ACQ_EI = 0,       /* GPyOpt/acquisitions/EI.py:32-51        par = jitter */
  GPO_ACQ_MPI = 1,      /* GPyOpt/acquisitions/MPI.py:32-51       par = jitter */
  GPO_ACQ_LCB = 2,      /* GPyOpt/acquisitions/LCB.py:35-50       par = exploration_weight */
  GPO_ACQ_EI_MCMC = 3,  /* GPyOpt/acquisitions/EI_mcmc.py:29-59   par = jitter (keeps the +jitter quirk) */
  GPO_ACQ_MPI_MCMC = 4, /* GPyOpt/acquisitions/MPI_mcmc.py:29-59 */
  GPO_ACQ_LCB_MCMC = 5  /* GPyOpt/acquisitions/LCB_mcmc.py:26-52 */
} gpo_acq_kind;

typedef enum { GPO_LP_NONE = 0, GPO_LP_SOFTPLUS = 1 } gpo_lp_transform;

const char* gpo_strerror(int status);
const char* gpo_last_error(const gpo_engine* eng);
const char* gpo_version(void);

/* Engine lifetime.  `device` is the CUDA ordinal.  Fails with GPO_ERR_NO_DEVICE when there is no GPU. */
int gpo_create(int device, gpo_engine** out);
int gpo_destroy(gpo_engine* eng);

/* Tunable: candidates per internal chunk (0 = automatic).  The K* chunk [S, chunk, n_pad] lives in HBM/L2. */
int gpo_set_chunk(gpo_engine* eng, long long chunk);

/* Refit given hyper-parameters: GPModel.updateModel / GPModel_MCMC per-sample `_trigger_params_changed`
 * (GPyOpt/models/gpmodel.py:78-93, 267-271) == GPy ExactGaussianInference.inference:
 *   Ky = K(X,X) + (noise + 1e-8) I;  L = jitchol(Ky);  alpha = Ky^-1 Y;  L^-1;  W^-1 = Ky^-1;  fmin = min(K alpha)
 * X (host) [n, d]; Y (host) [n] (already normalised by the caller, GPyOpt/core/bo.py:253-256);
 * variance [S]; lengthscale [S, d] (non-ARD: the single lengthscale repeated d times); noise [S].
 * Returns GPO_ERR_NOT_PD after GPy's jitchol retry ladder (5 tries, jitter = mean(diag)*1e-6 * 10^i). */
int gpo_fit(gpo_engine* eng, int kernel, int n, int d, const double* X, const double* Y, int S,
            const double* variance, const double* lengthscale, const double* noise);

/* GPModel.get_fmin / GPModel_MCMC.get_fmin (gpmodel.py:125-129, 279-295): fmin (host) [S]. */
int gpo_get_fmin(gpo_engine* eng, double* fmin);

/* log marginal likelihood and log|Ky| per hyper-parameter set (host) [S] each; either may be NULL.
 * (GPy ExactGaussianInference.inference; used by the ML-II fitter above the boundary.) */
int gpo_get_loglik(gpo_engine* eng, double* loglik, double* logdet);

/* d loglik / d [variance, lengthscale_1..d, noise] (host) [S, d + 2]: GPy dL_dK = 0.5 (alpha alpha^T - W^-1)
 * contracted with dK/dtheta (Stationary.update_gradients_full) and its trace (Gaussian likelihood). */
int gpo_get_loglik_grad(gpo_engine* eng, double* grad);

/* Copy posterior state to the host for inspection / broadcast checks; any pointer may be NULL.
 * L, Linv, Winv: [S, n, n]; alpha: [S, n]. */
int gpo_get_state(gpo_engine* eng, double* L, double* Linv, double* Winv, double* alpha);

/* GPModel.predict (gpmodel.py:95-112): m, s (host|device) [S, N]; s = sqrt(clip(var (+noise), 1e-10)). */
int gpo_predict(gpo_engine* eng, long long N, const double* Xs, int with_noise, double* m, double* s, void* stream);

/* GPModel.predict_withGradients (gpmodel.py:131-142): additionally dmdx, dsdx (host|device) [S, N, d]. */
int gpo_predict_grad(gpo_engine* eng, long long N, const double* Xs, double* m, double* s, double* dmdx,
                     double* dsdx, void* stream);

/* Posterior mean and its input gradient only (no variance: the dense contraction is skipped).  Serves the callers
 * that reach through to GPy's GP.predictive_gradients for the mean Jacobian alone -- estimate_L's Lipschitz sweep and
 * its L-BFGS (GPyOpt/core/evaluators/batch_local_penalization.py:52-72).  m [S, N], dmdx [S, N, d] (host|device). */
int gpo_predict_mean_grad(gpo_engine* eng, long long N, const double* Xs, double* m, double* dmdx, void* stream);

/* Acquisition*._compute_acq / _compute_acq_withGradients: f [N], df [N, d] (df may be NULL: value only).
 * For S > 1 the *_MCMC kinds average over the hyper-parameter sets.  If a Local-Penalization state is set
 * (gpo_lp_set) the output is AcquisitionLP.acquisition_function / d_acquisition_function instead
 * (GPyOpt/acquisitions/LP.py:70-132). */
int gpo_acq(gpo_engine* eng, int kind, double par, long long N, const double* Xs, double* f, double* df,
            void* stream);

/* Local-Penalization state: K penalisers centred at Xb (host) [K, d] with radii r [K], scales s [K]
 * (AcquisitionLP.update_batches, LP.py:40-62).  K = 0 with transform >= 0 keeps only the log transform;
 * transform < 0 clears LP. */
int gpo_lp_set(gpo_engine* eng, int K, const double* Xb, const double* r, const double* s, int transform);

/* Fused candidate sweep + selection (AnchorPointsGenerator.get, GPyOpt/optimization/
 * anchor_points_generator.py:25-69): evaluates the acquisition value over Xs [N, d] and returns the k
 * candidates with the smallest score = -f (ties: lowest index).  vals/idx are host arrays [k];
 * vals holds the scores (-f).  f_out (host|device, may be NULL) receives all N acquisition values.
 * k = 1 is served by the per-block argmax that the acquisition epilogue computes with warp shuffles; k > 1 by a
 * multi-level bitonic selection over f. */
int gpo_acq_topk(gpo_engine* eng, int kind, double par, long long N, const double* Xs, int k, double* vals,
                 long long* idx, double* f_out, void* stream);

/* Device-side random design (replaces samples_multidimensional_uniform,
 * GPyOpt/experiment_design/random_design.py:66-76, for continuous box domains): rows [row0, row0+N) of the
 * uniform design  numpy.random.Generator(numpy.random.Philox(key=key)).uniform(lo, hi, size=(rows, d))
 * (Philox4x64-10, row-major, bit-identical to NumPy), written to X (host|device) [N, d].
 * key: two 64-bit words; bounds (host) [d, 2] = (lo, hi) per dimension.  Needs no fitted model. */
int gpo_design_uniform(gpo_engine* eng, const unsigned long long* key, long long row0, long long N, int d,
                       const double* bounds, double* X, void* stream);

/* gpo_acq_topk over rows [row0, row0+N) of that design, generated on the device chunk by chunk inside the
 * sweep: no candidate ever crosses PCIe (initial_design('random') + AnchorPointsGenerator.get,
 * anchor_points_generator.py:25-69, in one call).  idx are LOCAL row numbers (global row = row0 + idx);
 * x_top (host, may be NULL) [k, d] receives the winning rows. */
int gpo_acq_topk_uniform(gpo_engine* eng, int kind, double par, const unsigned long long* key, long long row0,
                         long long N, const double* bounds, int k, double* vals, long long* idx, double* x_top,
                         void* stream);

/* Batched multi-start L-BFGS on the device: the local optimisation of the acquisition from M anchor points
 * (GPyOpt/optimization/acquisition_optimizer.py:73-74 -> optimizer.py:28-61, one SciPy L-BFGS-B run per anchor with one
 * f(x) and one f_df(x) host call per step).  All M <= 64 anchors advance in lock-step: per round one fused value + gradient
 * evaluation at the M trial points and one optimiser-state kernel (projected L-BFGS with 10 curvature pairs, Armijo
 * backtracking along the projected path, SciPy's stopping rules: projected gradient <= pgtol, relative decrease <=
 * factr * eps, maxiter iterations).  Minimises acquisition_function(x) = -acq(x) -- or the Local-Penalization value when
 * gpo_lp_set is active -- inside the box `bounds` (host [d, 2]).  x0 (host [M, d]) start points; x_out (host [M, d]) and
 * f_out (host [M]) the optimised points and the minimised values; iters / nfev (host [M], may be NULL) accepted steps and
 * evaluations per anchor. */
int gpo_lbfgs(gpo_engine* eng, int kind, double par, int M, const double* x0, const double* bounds, int maxiter, double pgtol,
              double factr, double* x_out, double* f_out, int* iters, int* nfev);

/* k smallest (largest = 0) or largest (largest = 1) entries of a device/host array v [N].  vals / idx (host|device,
 * both of the same kind): with device pointers the records stay in HBM and the call is asynchronous on `stream` (the
 * multi-GPU path all-gathers them without a host round trip); entries past min(k, N) are (NaN, -1). */
int gpo_topk(gpo_engine* eng, long long N, const double* v, int k, int largest, double* vals, long long* idx,
             void* stream);

/* Full posterior covariance between two small point sets X1 (host) [N1, d] and X2 (host) [N2, d]:
 * GPModel.get_covariance_between_points (gpmodel.py:173-177) = K(X1,X2) - (L^-1 K(X,X1))^T (L^-1 K(X,X2)).
 * X2 == NULL means X2 = X1 and is GPModel.predict_covariance's matrix (gpmodel.py:114-123, before its 1e-10 clip),
 * with the noise variance added on the diagonal when with_noise != 0.  out (host) [N1, N2].  S = 1 only; not a
 * large-sweep path (Entropy Search is the only caller in the reference). */
int gpo_posterior_cov(gpo_engine* eng, long long N1, const double* X1, long long N2, const double* X2, int with_noise,
                      double* out);

/* Variance form of gradient sweeps.  The reference takes the posterior variance from sigma_f^2 - |L^-1 k|^2
 * (GPy PosteriorExact._raw_predict via GPyOpt/models/gpmodel.py:98,136) and the variance gradient from b = W^-1 k
 * (GP.predictive_gradients, gpmodel.py:138).
 *   mode 1 (default): the reference's form, always.  Large sweeps run b = L^-T (L^-1 k) as two triangular products of
 *                     n^2 flop each; the first one IS the sum-of-squares form, so the variance matches the reference by
 *                     construction at the 2 n^2 flop per candidate of the single W^-1 product.
 *   mode 0:           fast form, variance from sigma_f^2 - k.b with b = W^-1 k (one full product; cancels badly on
 *                     ill-conditioned Gram matrices).  Opt-in.
 *   mode -1:          automatic: after every refit both forms are compared at (up to 512 of) the training inputs and
 *                     the fast form is kept only where max |s_L - s_W| <= 7e-12 sqrt(sf2 + noise).  Opt-in.
 * gpo_get_info: info[0] = reference form active, info[1] = measured discrepancy (automatic mode, scaled),
 * info[2] = chunk (candidates), info[3] = padded n, info[4] = SM count, info[5] = S. */
int gpo_set_strict_variance(gpo_engine* eng, int mode);
int gpo_get_info(gpo_engine* eng, double* info, int count);

/* Tuning knobs for measurements (the defaults are the product configuration; results do not depend on them beyond the
 * summation order of the partial rows).
 *   GPO_OPT_SERPENTINE  1 (default) / 0: deal the GEMM work items to the CTAs in serpentine rounds
 *   GPO_OPT_BN_SWEEP    64 (default) / 128: column-tile width of the sweep GEMMs (64 = two CTAs per SM)
 *   GPO_OPT_BN_FIT      64 (default) / 128: the same for the rectangular K = 128 products of the refit
 *   GPO_OPT_INVERSE     0 automatic (default) / 1 interleaved with the Cholesky / 2 recursive triangular inverse
 *   GPO_OPT_FUSED_SMALL 1 (default) / 0: value sweeps of models with n <= 128 take the fused shared-memory-resident kernel
 *   GPO_OPT_PREFETCH_KT 0 (default) / 1: L2 prefetch of the K* tile that the gradient epilogue of the second triangular product
 *                       re-reads (measured: +0.03 % RBF, -0.4 % Matern at n = 2048 -- the epilogue is not waiting on those loads)
 *   GPO_OPT_FIT_PATH    0 (default): refit with the short critical chain (diagonal-block Cholesky -> panel by block substitution ->
 *                       block-column update; inverse, trailing update and W^-1 on side streams) / 1: the round-1 look-ahead scheme
 *   GPO_OPT_FIT_TRACE   0 (default) / 1: time every launch of the refit with events and print the timeline to stderr (slows the refit)
 *   GPO_OPT_FIT_GRAPH   1 (default) / 0: replay the refit's launch sequence as a CUDA graph captured once per problem shape
 *   GPO_OPT_FIT_CHAIN_MAX  largest number of 128-row blocks (n_pad / 128) the short-chain refit is used for
 *   GPO_OPT_DESYNC      1 (default) / 0: the two GEMM CTAs sharing an SM walk their work items in opposite orders, so that their
 *                       epilogues do not coincide (results are unchanged bit for bit)
 *   GPO_OPT_K1_SLICE    training rows per slice of the covariance kernel K1 (128 / 256 / 512); takes effect with the next gpo_fit */
typedef enum { GPO_OPT_SERPENTINE = 0, GPO_OPT_BN_SWEEP = 1, GPO_OPT_BN_FIT = 2, GPO_OPT_INVERSE = 3, GPO_OPT_FUSED_SMALL = 4,
               GPO_OPT_PREFETCH_KT = 5, GPO_OPT_FIT_PATH = 6, GPO_OPT_FIT_TRACE = 7, GPO_OPT_FIT_GRAPH = 8,
               GPO_OPT_FIT_CHAIN_MAX = 9, GPO_OPT_DESYNC = 10,
               GPO_OPT_K1_SLICE = 11 } gpo_option;
int gpo_set_option(gpo_engine* eng, int key, long long value);

/* Multi-GPU replication of the posterior state (one process per GPU; the collective itself is issued by the
 * host through torch.distributed / NCCL on the buffers listed here).
 *   gpo_alloc          allocate an engine for (kernel, n, d, S) without fitting (receiving ranks)
 *   gpo_state_buffers  device pointers + byte sizes of every buffer a sweep reads (X, alpha, L^-1, W^-1,
 *                      hyper-parameters, fmin ...); returns their count (<= max) or a negative status
 *   gpo_state_commit   after the buffers were overwritten (broadcast from the fitting rank): refresh the
 *                      host mirrors and mark the engine fitted */
int gpo_alloc(gpo_engine* eng, int kernel, int n, int d, int S);
int gpo_state_buffers(gpo_engine* eng, void** ptrs, long long* bytes, int max);
int gpo_state_commit(gpo_engine* eng);

/* Live timing of the dominant kernel (the variance GEMM launched once per sweep chunk), measured with CUDA events on
 * the stream the kernel runs on.  While profiling is enabled the sweep chunks are serialised on one stream slot (so the
 * events bracket that kernel alone; the normal two-slot pipeline overlaps neighbouring chunks).  gpo_get_profile
 * synchronises, returns the accumulated milliseconds, launches and candidates since the last call, and resets. */
int gpo_set_profiling(gpo_engine* eng, int enable);
int gpo_get_profile(gpo_engine* eng, double* gemm_ms, long long* gemm_launches, long long* gemm_candidates);

/* Counters for the benchmark harness: kernels launched and seconds of GPU time since the last reset. */
int gpo_launch_count(gpo_engine* eng, long long* launches, int reset);

#ifdef __cplusplus
}
#endif
#endif /* GPO_B200_H */
