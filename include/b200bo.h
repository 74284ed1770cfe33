/* b200bo.h -- C ABI of libb200bo.so: the B200-native (sm_100a) GP-posterior + acquisition-search path.
 *
 * The reference (jbrea/BayesianOptimization.jl) has no FFI; its plugin boundary is Julia multiple dispatch
 * on the model type (src/models/gp.jl:2-18,42-77) plus the NLopt-backed search (src/acquisition.jl:20-68).
 * Each entry below names the reference interface it replaces.  A Julia host binds these with `ccall`
 * (bayesianoptimization.jl_b200/julia/B200BayesOpt.jl); the Python host mirror binds the
 * same symbols with ctypes.
 *
 * Conventions
 *  - all reals IEEE Float64, all sizes int64_t, matrices COLUMN-MAJOR with points as columns (model.x is
 *    D x N, src/models/gp.jl:9; candidates D x M, test/acquisitionfunctions.jl:6).
 *  - every entry returns an int32 status (B200BO_OK == 0, negative = error) and never throws.
 *  - pointers are HOST pointers unless the entry name ends in `_dev`; `_dev` entries take DEVICE pointers,
 *    only enqueue work on the handle's stream and return without synchronising (use b200bo_sync).
 *  - a handle is not re-entrant; distinct handles are independent.
 *  - there is NO CPU fallback: without a usable CUDA device every compute entry returns B200BO_ERR_CUDA.
 */
#ifndef B200BO_H
#define B200BO_H
#include <stdint.h>

#if defined(__GNUC__)
#define B200BO_API __attribute__((visibility("default")))
#else
#define B200BO_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200bo_handle_s* b200bo_handle_t;

/* status codes */
enum {
  B200BO_OK = 0,
  B200BO_ERR_ARG = -1,       /* invalid argument (the reference throws ArgumentError / DimensionMismatch) */
  B200BO_ERR_CUDA = -2,      /* CUDA runtime / driver error, or no device */
  B200BO_ERR_NOTPD = -3,     /* not positive definite after 10 jitter retries (EXT make_posdef!) */
  B200BO_ERR_STATE = -4,     /* call needs a fitted model */
  B200BO_ERR_ALLOC = -5,
  B200BO_ERR_NCCL = -6       /* NCCL could not be loaded or a collective failed */
};

/* EXT GaussianProcesses.jl kernels reachable from the reference (README.md:22-26, BayesianOptimization.jl:259-264,
 * test/acquisitionfunctions.jl:4).  Kernel parameter vector: Iso = [ll, lsigma], Ard = [ll_1..ll_D, lsigma]. */
enum {
  B200BO_KERNEL_SEISO = 0, B200BO_KERNEL_SEARD = 1,
  B200BO_KERNEL_MAT12ISO = 2, B200BO_KERNEL_MAT12ARD = 3,
  B200BO_KERNEL_MAT32ISO = 4, B200BO_KERNEL_MAT32ARD = 5,
  B200BO_KERNEL_MAT52ISO = 6, B200BO_KERNEL_MAT52ARD = 7
};
enum { B200BO_MEAN_ZERO = 0, B200BO_MEAN_CONST = 1 };

/* acquisition functors, src/acquisitionfunctions.jl:21-28 (PI), :40-50 (EI), :81-96 (UCB), :107-108 (TS),
 * :126-141 (MI), :110-111 (MaxMean).  acq_params: PI/EI = {tau}; UCB = {beta_t}; MI = {sqrt_alpha, gamma_hat};
 * TS/MaxMean = none.  The host computes them in setparams! (:3,44-46,91-95,131-140). */
enum {
  B200BO_ACQ_PI = 0, B200BO_ACQ_EI = 1, B200BO_ACQ_UCB = 2, B200BO_ACQ_TS = 3, B200BO_ACQ_MI = 4,
  B200BO_ACQ_MAXMEAN = 5
};

/* parameter-selection mask for b200bo_mll_sweep == get_params_kwargs(gp; noise, domean, kern) (gp.jl:55-58) */
enum { B200BO_MASK_NOISE = 1, B200BO_MASK_MEAN = 2, B200BO_MASK_KERN = 4 };

/* variants of b200bo_acquire_direct: NLopt :GN_DIRECT_L (locally biased, the reference's default for ThompsonSamplingSimple) and :GN_DIRECT
 * (Jones' original) */
enum { B200BO_DIRECT_L = 0, B200BO_DIRECT_ORIG = 1 };

/* which device timing b200bo_last_timing_ms returns (CUDA events on the handle's stream) */
enum {
  B200BO_T_KMAT = 0,      /* kernel-matrix assembly (K1) of the last fit */
  B200BO_T_CHOL = 1,      /* blocked Cholesky (K2-K4) of the last fit */
  B200BO_T_SYRK = 2,      /* sum of trailing-update launches (K4) of the last fit */
  B200BO_T_ALPHA = 3,     /* alpha / logdet (K5) of the last fit */
  B200BO_T_ACQ = 4,       /* fused acquisition launch (K6) of the last predict/acquire */
  B200BO_T_MLL = 5,       /* last mll sweep, all settings */
  B200BO_T_ACQ_GEMM = 6,  /* sum over the int8 slice-product launches (acq_i8_gemm_kernel) of the last acquire; needs knob "acq_gemm_timing" */
  B200BO_T_COUNT = 7
};

typedef struct { double value; int64_t index; } b200bo_best_t;   /* index = -1: nothing beat -Inf */

/* -- lifecycle: replaces ElasticGPE(D; mean, kernel, logNoise, capacity) (README.md:22-26) ------------------- */
B200BO_API int32_t b200bo_create(b200bo_handle_t* h, int32_t device, int32_t D, int64_t capacity,
                      int32_t kernel_kind, int32_t mean_kind);
B200BO_API int32_t b200bo_destroy(b200bo_handle_t h);

/* -- multi-GPU (the shard axis is the plain loop over starts of acquire_max, src/acquisition.jl:58-66) ----------------------
 * b200bo_create_multi: ONE process, n_gpus replicas of the model (devices[i], or devices 0..n_gpus-1 when NULL).  Model updates (fit /
 *   append / refit / set_params) go to every replica; b200bo_acquire / b200bo_acquire_lhs / b200bo_acquire_ascent / b200bo_predict shard
 *   their candidate columns and b200bo_mll_sweep its settings over the replicas in contiguous blocks.  b200bo_acquire's exchange step is
 *   ONE ncclAllGather of a 272-byte record per rank (value, global index, point) + a deterministic merge on the device (largest value,
 *   lowest index): the result equals the single-GPU result bit for bit.  `_dev` entries of a multi handle address replica 0 only.
 * b200bo_comm_*: one process per GPU (torchrun / MPI-style hosts): rank 0 calls b200bo_comm_unique_id, the host distributes the 128
 *   bytes, every rank calls b200bo_comm_init_rank on its own handle; from then on b200bo_acquire and b200bo_acquire_dev return the
 *   GLOBAL best (and b200bo_acquire's best_x the global winner's point) on every rank -- a collective call: all ranks must make it.
 *   (b200bo_acquire_lhs / b200bo_acquire_ascent stay local: they return the best of the block they were given.) */
B200BO_API int32_t b200bo_create_multi(b200bo_handle_t* h, int32_t n_gpus, const int32_t* devices, int32_t D, int64_t capacity,
                                       int32_t kernel_kind, int32_t mean_kind);
B200BO_API int32_t b200bo_num_gpus(b200bo_handle_t h, int32_t* n_gpus);
B200BO_API int32_t b200bo_comm_unique_id(uint8_t* id128);
B200BO_API int32_t b200bo_comm_init_rank(b200bo_handle_t h, int32_t world, int32_t rank, const uint8_t* id128);
B200BO_API int32_t b200bo_comm_destroy(b200bo_handle_t h);
B200BO_API const char* b200bo_last_error(b200bo_handle_t h);          /* valid until the next call on h; h may be NULL */
B200BO_API int32_t b200bo_set_stream(b200bo_handle_t h, void* cuda_stream);   /* borrow a caller stream (NULL = own) */
B200BO_API int32_t b200bo_sync(b200bo_handle_t h);

/* -- hyper-parameters: EXT get_params / set_params! (gp.jl:55-58,60,74). theta = [logNoise, (beta), kernel...] */
B200BO_API int32_t b200bo_num_params(b200bo_handle_t h, int32_t* P);
B200BO_API int32_t b200bo_set_params(b200bo_handle_t h, const double* theta, int32_t P);   /* invalidates the factor */
B200BO_API int32_t b200bo_get_params(b200bo_handle_t h, double* theta, int32_t P);
/* EXT set_priors! (reference src/models/gp.jl:30-35: "non-flat priors can be specified directly on the GP parameters"): kind[i] = 0 flat,
   1 Normal(a[i], b[i]) per parameter in the order of theta; b200bo_mll_sweep / b200bo_map_fit then work on mll + log prior.  P = 0 clears. */
B200BO_API int32_t b200bo_set_priors(b200bo_handle_t h, int32_t P, const int32_t* kind, const double* a, const double* b);

/* -- model update: update!(model, X, y) (gp.jl:11-18).  fit = full refactor (GP.fit!), append = elastic append! */
B200BO_API int32_t b200bo_fit(b200bo_handle_t h, const double* X, const double* y, int64_t N);
B200BO_API int32_t b200bo_append(b200bo_handle_t h, const double* Xnew, const double* ynew, int64_t m);
B200BO_API int32_t b200bo_refit(b200bo_handle_t h);                   /* refactor with current data + params */

/* -- model queries: dims / maxy / model.x / model.y (gp.jl:9-10, BayesianOptimization.jl:117-119) ------------ */
B200BO_API int32_t b200bo_dims(b200bo_handle_t h, int32_t* D, int64_t* N);
B200BO_API int32_t b200bo_maxy(b200bo_handle_t h, double* maxy);      /* -Inf when empty */
B200BO_API int32_t b200bo_get_data(b200bo_handle_t h, double* X, double* y);
B200BO_API int32_t b200bo_get_mll(b200bo_handle_t h, double* mll);    /* gp.mll / gp.target */
B200BO_API int32_t b200bo_get_alpha(b200bo_handle_t h, double* alpha);
B200BO_API int32_t b200bo_get_factor(b200bo_handle_t h, double* U);   /* N x N column-major upper factor, Sigma = U'U */
B200BO_API int32_t b200bo_jitter_tries(b200bo_handle_t h, int32_t* tries);

/* -- posterior: mean_var(model, X) = GP.predict_f (gp.jl:2-5,8) ---------------------------------------------- */
B200BO_API int32_t b200bo_predict(b200bo_handle_t h, const double* Xs, int64_t M, double* mu, double* var);

/* -- joint posterior sample: myrand(model, X::Matrix) = EXT rand(gp, X) (gp.jl:7): ONE draw
 *    mu + chol(make_posdef!(Sigma_post)) eps over the M columns of Xs with the full M x M posterior covariance (latent, no noise term);
 *    eps_j = the Thompson Philox stream at (seed, idx_offset + j).  The covariance is never formed: the points are appended noise-free to
 *    a worker copy of the model and the augmented matrix is factorised by the fit kernels (csrc/joint.cu).  N + M <= 32768.
 *    mu_out (optional) = the posterior mean; tries (optional) = make_posdef! jitter retries (1e-6 tr/M each, at most 10).
 *    The vector form myrand(model, x) -- an independent draw per column -- is b200bo_acquire with B200BO_ACQ_TS. */
B200BO_API int32_t b200bo_rand_joint(b200bo_handle_t h, const double* Xs, int64_t M, uint64_t seed, int64_t idx_offset, double* sample /*M*/,
                                     double* mu_out /*M or NULL*/, int32_t* tries /*or NULL*/);

/* -- acquisition step: acquisitionfunction(a, model)(X) + the selection rule of acquire_max
 *    (acquisitionfunctions.jl:4-9, acquisition.jl:54-68).  One fused launch over the M candidate columns:
 *    k(x*,X), both triangular solves, mu, sigma^2, a(mu, sigma^2), optional gradient D x M, arg-max with
 *    first-strict-maximum (lowest index) tie-break, NaN never wins.  idx_offset = global index of column 0 (keys
 *    the TS Philox stream and is added to best->index) so shards of one candidate matrix agree with the whole.
 *    Host buffers may be pageable or pinned (cudaHostAlloc / cudaHostRegister); with pinned Xs and outputs the library moves the data chunk
 *    by chunk on its stream lanes, under the kernels of the neighbouring chunk (same results, the PCIe time disappears from the call). */
B200BO_API int32_t b200bo_acquire(b200bo_handle_t h, int32_t acq_kind, const double* acq_params, int32_t n_params,
                       const double* Xs, int64_t M, uint64_t seed, int64_t idx_offset,
                       double* values /*M or NULL*/, double* grad /*D x M or NULL*/,
                       double* mu /*M or NULL*/, double* var /*M or NULL*/,
                       b200bo_best_t* best /*or NULL*/, double* best_x /*D or NULL*/);

/* -- MAP objective: the closure of optimizemodel! (gp.jl:59-64): for each column of Theta (P x S) set the
 *    masked params, refactor, return mll[S] and dmll (P x S, may be NULL).  Leaves the model -- parameters AND
 *    factor -- untouched: the settings are factored on worker buffers, several in flight at once (knob
 *    "sweep_workers"), and on a multi handle they shard over the GPUs. */
B200BO_API int32_t b200bo_mll_sweep(b200bo_handle_t h, const double* Theta, int32_t P, int32_t S, int32_t mask,
                         double* mll, double* dmll);

/* -- device-pointer variants (inputs resident in HBM; enqueue only) ------------------------------------------ */
B200BO_API int32_t b200bo_kmat_dev(b200bo_handle_t h, double* dK, int64_t ld);   /* K1 only: Sigma into dK (ld >= N) */
B200BO_API int32_t b200bo_predict_dev(b200bo_handle_t h, const double* dXs, int64_t M, double* dmu, double* dvar);
B200BO_API int32_t b200bo_acquire_dev(b200bo_handle_t h, int32_t acq_kind, const double* acq_params, int32_t n_params,
                           const double* dXs, int64_t M, uint64_t seed, int64_t idx_offset,
                           double* dvalues, double* dgrad, double* dmu, double* dvar,
                           b200bo_best_t* dbest);

/* -- search helpers that keep the candidates in HBM ------------------------------------------
 * b200bo_lhs: latin_hypercube_sampling(lb, ub, n) (src/utils.jl:101-120) on device: columns [offset, offset + n_local) of ONE
 *   global n_total-point design (each stratum used once per dimension; shards of the same (seed, n_total) agree).
 * b200bo_acquire_lhs: ScaledLHSIterator + acquire_max's sweep (src/acquisition.jl:57-66) without host candidates.
 * b200bo_acquire_ascent: M box-constrained gradient ascents in lock-step from the columns of Xs (what NLopt LD_LBFGS adds per
 *   restart, src/acquisition.jl:59), `steps` value+gradient launches; returns refined points / values / best. */
B200BO_API int32_t b200bo_lhs(b200bo_handle_t h, const double* lb, const double* ub, int64_t n_total, int64_t offset, int64_t n_local,
                              uint64_t seed, double* Xs /* host, D x n_local */);
B200BO_API int32_t b200bo_acquire_lhs(b200bo_handle_t h, int32_t acq_kind, const double* acq_params, int32_t n_params,
                                      const double* lb, const double* ub, int64_t n_total, int64_t offset, int64_t n_local,
                                      uint64_t lhs_seed, uint64_t ts_seed, double* values /*n_local or NULL*/,
                                      b200bo_best_t* best, double* best_x /*D or NULL*/);
B200BO_API int32_t b200bo_acquire_ascent(b200bo_handle_t h, int32_t acq_kind, const double* acq_params, int32_t n_params,
                                         const double* Xs, int64_t M, const double* lb, const double* ub, int32_t steps, double step0,
                                         int64_t idx_offset, double* Xout /*D x M or NULL*/, double* values /*M or NULL*/,
                                         b200bo_best_t* best, double* best_x /*D or NULL*/);

/* b200bo_sobol: ScaledSobolIterator (src/utils.jl:64-87) on device: points index0 .. index0+n-1 of the unscrambled Joe-Kuo Sobol sequence
 *   (index 0 = the origin, which EXT Sobol.jl never emits: its k-th point is index k; the reference's iterator starts at index
 *   1 + 2^floor(log2(N+1)), utils.jl:80), scaled to [lb, ub].  D <= 32, indices below 2^32.
 * b200bo_acquire_lbfgs: what NLopt :LD_LBFGS adds per restart (src/acquisition.jl:59) for ALL M starts at once: box-bounded L-BFGS ascents
 *   in lock-step on the fused value + gradient launch, with the options the reference forwards (src/acquisition.jl:24-27): maxeval
 *   (evaluations per start, <= 0 unlimited), ftol_rel / ftol_abs, xtol_rel / xtol_abs (0 disables), maxtime (seconds, <= 0 unlimited).
 *   evals[M] (optional) = evaluations each start used.
 * b200bo_map_fit: optimizemodel!(::MAPGPOptimizer, model) (src/models/gp.jl:54-77): L-BFGS ascent of mll over the masked parameters within
 *   [lb, ub] (gp.jl:65-68) from R starts in lock-step (Theta0 P x R; R = 1 and Theta0 = current parameters is the reference's run);
 *   every evaluation is a device refactorisation + gradient.  Leaves the model at the best parameters; status = NLopt's numbering
 *   (3 FTOL_REACHED, 4 XTOL_REACHED, 5 MAXEVAL_REACHED, 6 = stalled). */
/* b200bo_acquire_direct: the derivative-free global search the reference selects with method = :GN_DIRECT_L (the default for
 *   ThompsonSamplingSimple, src/acquisition.jl:7-9; wrapper :31-36): locally-biased DIRECT on [lb, ub] (csrc/direct.h), every iteration's
 *   new centres -- all potentially optimal rectangles, all their longest sides -- scored by ONE fused acquisition launch.  maxeval = total
 *   evaluations (exactly, as NLopt counts them), maxtime in seconds (<= 0 unlimited), width = rectangles divided per hull size class
 *   (1 = DIRECT-L; ignored by the original variant), variant = B200BO_DIRECT_L or B200BO_DIRECT_ORIG (diagonal size measure, every tie of a
 *   hull class divides, only cubes are cut along all sides).  Evaluation e uses the Thompson stream (seed, global index e).  Xtrace (D x maxeval) / ftrace (maxeval) optionally
 *   receive every evaluated point and value in evaluation order; evals = evaluations used, batches = device launches.  best->index = the
 *   evaluation (0-based, = column of Xtrace) that produced the best value, -1 if nothing beat -Inf. */
B200BO_API int32_t b200bo_acquire_direct(b200bo_handle_t h, int32_t acq_kind, const double* acq_params, int32_t n_params, const double* lb,
                                         const double* ub, int32_t maxeval, double maxtime, int32_t width, int32_t variant, uint64_t seed,
                                         double* Xtrace /*D x maxeval or NULL*/, double* ftrace /*maxeval or NULL*/, int32_t* evals /*or NULL*/,
                                         int32_t* batches /*or NULL*/, b200bo_best_t* best, double* best_x /*D or NULL*/);
B200BO_API int32_t b200bo_sobol(b200bo_handle_t h, const double* lb, const double* ub, uint64_t index0, int64_t n, double* Xs /* host, D x n */);
B200BO_API int32_t b200bo_acquire_lbfgs(b200bo_handle_t h, int32_t acq_kind, const double* acq_params, int32_t n_params, const double* Xs, int64_t M,
                                        const double* lb, const double* ub, int32_t maxeval, double ftol_rel, double ftol_abs, double xtol_rel,
                                        double xtol_abs, double maxtime, double step0, int64_t idx_offset, double* Xout /*D x M or NULL*/,
                                        double* values /*M or NULL*/, double* evals /*M or NULL*/, b200bo_best_t* best, double* best_x /*D or NULL*/);
B200BO_API int32_t b200bo_map_fit(b200bo_handle_t h, const double* Theta0, int32_t P, int32_t R, int32_t mask, const double* lb, const double* ub,
                                  int32_t maxeval, double ftol_rel, double ftol_abs, double xtol_rel, double xtol_abs, double maxtime,
                                  double* theta_best /*P*/, double* mll_best, int32_t* evals, int32_t* status);

/* -- introspection for benches / tests ----------------------------------------------------------------------- */
B200BO_API int32_t b200bo_kmat(b200bo_handle_t h, double* K);         /* N x N Sigma = K + (e^{2 logNoise}+eps) I to host */
B200BO_API int32_t b200bo_last_timing_ms(b200bo_handle_t h, int32_t which, float* ms);
B200BO_API int32_t b200bo_launch_count(b200bo_handle_t h, int64_t* launches);   /* kernels launched since create */
B200BO_API int32_t b200bo_fp64_peak_tflops(b200bo_handle_t h, double* tflops);   /* self-measured DMMA.8x8x4 rate: the
                                                                        FP64 tensor-pipe roofline denominator */
/* engine of the K = 512 trailing updates of the factorisation: 1 = tcgen05 int8-slice product on 128 x 64 tiles (default),
   2 = the same on 128 x 128 tiles in two passes (experimental, no faster), 0 = DMMA tile GEMM.
   Both are FP64-accurate; the switch exists so that tests and benches can compare them in one process. */
B200BO_API int32_t b200bo_set_syrk_engine(b200bo_handle_t h, int32_t engine);
/* engine of the acquisition step (K6): 1 = error-free int8-slice product against the explicit inverse factor on tcgen05 (default),
   0 = blocked triangular solves on the FP64 tensor pipe (DMMA).  Both are FP64-accurate; the switch exists for in-process A/B tests. */
B200BO_API int32_t b200bo_set_acq_engine(b200bo_handle_t h, int32_t engine);
/* self-measured tcgen05.mma kind::i8 rate in TOP/s (2 ops per multiply-add): the int8 tensor-pipe roofline denominator of K4 / K6 */
B200BO_API int32_t b200bo_i8_peak_tops(b200bo_handle_t h, double* tops);
/* developer / bench knobs: "acq_lanes" (1 or 2 chunk lanes of the tcgen05 acquisition path), "acq_chunk_mb" (k* slice bytes per chunk,
   0 = default), "acq_gemm_timing" (1: CUDA-event pairs around every slice-product launch, one lane; read B200BO_T_ACQ_GEMM),
   "sweep_workers" (settings of a MAP sweep in flight at once on one GPU, default 6; 0 = one after the other on the model's own buffers),
   "chol_sched" (schedule of the blocked Cholesky: 1 = look-ahead, only potrf + the fused cluster head of each panel on the critical chain
   (default); 2 = look-ahead with the tile-GEMM kernels as heads; 0 = the in-order schedule of round 1; for A/B timing),
   "chol_graph" (the look-ahead factorisation replays as one CUDA graph per shape, captured at the shape's 6th consecutive factorisation
   (default; capture + instantiation cost 3-17 ms once, a replay saves 0.12-0.35 ms); k >= 2 = capture at the k-th, 0 = always eager) */
B200BO_API int32_t b200bo_set_knob(b200bo_handle_t h, const char* name, int64_t value);
B200BO_API int32_t b200bo_version(void);

#ifdef __cplusplus
}
#endif
#endif /* B200BO_H */
