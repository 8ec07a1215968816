/* amcmc.h -- C ABI of libamcmc.so, the B200-native many-chain adaptive Metropolis sampler.
 *
 * This is the drop-in boundary for the hot path of savelovme/adaptive-mcmc:
 * the ARWMH sampler kernel `python/kernels/arwmh.py` (class ARWMH, :31-276) plus
 * the model log-densities it calls (`python/scripts/run_*_lr_decay.py`).  Every
 * entry point below cites the reference interface it replaces (paths relative
 * to the reference repo root).  Signatures use plain pointers and sizes only --
 * no torch / C++ types -- so any FFI (ctypes, cffi, pybind, JAX custom_call) can
 * bind them; INTEGRATION.md shows the reference-side binding.
 *
 * Conventions
 *   - All state/draw/output pointers are DEVICE pointers on the current CUDA
 *     device unless the function name ends in `_host`.
 *   - Chain state is struct-of-arrays with the chain index fastest:
 *       z[k*C + c], loc[k*C + c]                      k in [0,d)
 *       scale[(i*(i+1)/2 + j)*C + c]                  packed lower triangle, j <= i
 *     (the reference stores `scale` dense d x d with lower-triangular content,
 *     arwmh.py:124; the Python host expands the packed form on request).
 *   - dtype is AMCMC_F32 (reference precision) or AMCMC_F64 (parity path).
 *   - Every function returns 0 on success, a negative amcmc_status otherwise,
 *     never throws, and enqueues work on the caller's `stream` without a host
 *     synchronisation (except `_host` variants, which synchronise before return).
 */
#ifndef AMCMC_H_
#define AMCMC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AMCMC_VERSION 100

enum amcmc_dtype { AMCMC_F32 = 0, AMCMC_F64 = 1 };

enum amcmc_status {
  AMCMC_OK = 0,
  AMCMC_ERR_ARG = -1,       /* bad argument (the reference raises ValueError, arwmh.py:69-70,118-119) */
  AMCMC_ERR_CUDA = -2,      /* a CUDA runtime call failed; see amcmc_last_error() */
  AMCMC_ERR_UNSUPPORTED = -3 /* model / dtype / dimension combination not compiled */
};

/* Model families = the NumPyro model functions the reference scripts define. */
enum amcmc_model_id {
  AMCMC_MODEL_STD_NORMAL = 0,   /* potential_fn = 0.5*|x|^2, python/jupyter/asumptions_check.ipynb cells 17-28 */
  AMCMC_MODEL_EIGHT_SCHOOLS = 1, /* python/scripts/run_eight_schools_lr_decay.py:26-35 */
  AMCMC_MODEL_KIDIQ = 2,        /* python/scripts/run_kidiq_kidscore_lr_decay.py:29-41 */
  AMCMC_MODEL_DIAMONDS = 3,     /* python/scripts/run_diamonds_lr_decay.py:24-40 */
  AMCMC_MODEL_GAUSSIAN = 4,     /* BASELINE.json config 5: N(0, Sigma), precision Cholesky supplied */
  AMCMC_MODEL_CUSTOM = 5        /* potential compiled from user CUDA source into a plugin (amcmc_model_create_custom) */
};

enum amcmc_rng_mode {
  AMCMC_RNG_PHILOX = 0,   /* in-kernel Philox4x32-10 keyed by (seed, global chain id, iteration) */
  AMCMC_RNG_EXTERNAL = 1  /* caller supplies normals[T][d][C] and uniforms[T][C] (shared-draw parity mode) */
};

/* Sampler variants: AMCMC_KERNEL_ARWMH = python/kernels/arwmh.py; AMCMC_KERNEL_RAM = BASELINE.json configs[4]
 * (not in the reference); AMCMC_KERNEL_ASSS = python/kernels/asss.py:192-269, the adaptive stereographic slice
 * sampler (state: log_step_size unused, mean_accept_prob carries the mean number of shrinkage iterations;
 * external draws: normals[T][d+1][C], uniforms[T][52][C] = u_t, theta_0/2pi, 50 shrinkage draws).  ASSS runs on the
 * thread-per-chain kernels (STD_NORMAL, EIGHT_SCHOOLS, KIDIQ, CUSTOM) and on the CTA-per-chain kernels (DIAMONDS,
 * GAUSSIAN with dim <= 32, CUSTOM row models); adapt = 0 is ASSS.sample_Pnx (asss.py:279-315). */
enum amcmc_kernel_kind { AMCMC_KERNEL_ARWMH = 0, AMCMC_KERNEL_RAM = 1, AMCMC_KERNEL_ASSS = 2 };

typedef struct amcmc_model amcmc_model; /* opaque: device copies of the model data */

/* Replaces numpyro.infer.util.initialize_model + potential_fn_gen(*args, **kwargs)
 * (arwmh.py:111-116): binds model data and returns a handle whose potential the
 * kernels inline.  `arrays` are HOST float64 arrays, copied to the current device:
 *   STD_NORMAL    : none                          (dim = d)
 *   EIGHT_SCHOOLS : y[8], sigma[8]                (dim = 10)
 *   KIDIQ         : kid_score[n], mom_hs[n], mom_iq[n]   (dim = 4)
 *   DIAMONDS      : X[n*K] row-major (col 0 = ones), Y[n]  (dim = K+1; lens[0] = n*K, lens[1] = n)
 *   GAUSSIAN      : P[d*d] row-major lower Cholesky of the precision  (dim = d)
 */
int amcmc_model_create(amcmc_model** out, int model_id, int dtype, int dim, int n_arrays,
                       const double* const* arrays, const int64_t* lens);
/* The reference accepts ANY NumPyro model function (arwmh.py:43-78, 111-116: potential_fn traced by JAX).  The
 * counterpart here: the potential is written as a CUDA device function, compiled together with the fused
 * thread-per-chain kernels into a plugin library (adaptive_mcmc_b200/custom.py generates and builds it with nvcc),
 * and bound to its data by this call.  `plugin_path`: the plugin .so; up to 4 HOST float64 arrays are copied to the
 * device in `dtype` and handed to the potential as a0..a3 / n0..n3.  The model dimension comes from the plugin. */
int amcmc_model_create_custom(amcmc_model** out, const char* plugin_path, int dtype, int n_arrays,
                              const double* const* arrays, const int64_t* lens);
int amcmc_model_destroy(amcmc_model* m);
int amcmc_model_dim(const amcmc_model* m);
int amcmc_model_dtype(const amcmc_model* m);

/* ARWMHState + ARWMHAdaptState (arwmh.py:15-28) for C chains, as device SoA pointers. */
typedef struct amcmc_state {
  int64_t n_chains;        /* C */
  int32_t dim;             /* d */
  int32_t dtype;           /* amcmc_dtype */
  int64_t i;               /* ARWMHState.i  -- shared by all chains; advanced by run */
  void* z;                 /* ARWMHState.z (flat, ravel_pytree order)  [d][C] */
  void* potential_energy;  /* [C] */
  void* mean_accept_prob;  /* [C] */
  void* loc;               /* ARWMHAdaptState.loc            [d][C] */
  void* scale;             /* ARWMHAdaptState.scale, packed  [d(d+1)/2][C] */
  void* log_step_size;     /* ARWMHAdaptState.log_step_size  [C] */
  void* as_change;         /* [C] */
} amcmc_state;

/* ARWMH.__init__ hyper-parameters (arwmh.py:43-45) + ARWMH.init's num_warmup (:107)
 * + the driver's collection plan (numpyro.util.fori_collect as used at
 * python/utils/kernel_utils.py:29-32: sample k = state after
 * collect_start + (k+1)*thinning steps). */
typedef struct amcmc_run_args {
  int64_t n_steps;         /* K fused steps in this call */
  int64_t thinning;        /* >= 1 */
  int64_t collect_start;   /* steps to skip before collection starts */
  int64_t num_warmup;      /* arwmh.py:181: n restarts at 1 after warmup */
  double lr_decay;         /* gamma = n^-lr_decay (arwmh.py:183) */
  double target_accept_prob;
  double eps;              /* proposal regularisation (arwmh.py:166) */
  int32_t adapt;           /* 1: ARWMH.sample (:140-207); 0: frozen kernel of sample_Pnx (:230-249) */
  int32_t rng_mode;        /* amcmc_rng_mode */
  uint64_t seed;
  int64_t chain_offset;    /* global id of this shard's first chain (multi-GPU sharding) */
  const void* normals;     /* EXTERNAL: [n_steps][d][C] */
  const void* uniforms;    /* EXTERNAL: [n_steps][C] */
  void* out_z;             /* [S][d][C] or NULL, S = (n_steps - collect_start) / thinning */
  void* out_potential_energy; /* [S][C] or NULL */
  uint8_t* out_accept;     /* [n_steps][C] accept decisions, or NULL (parity tests) */
  int32_t kernel_kind;     /* amcmc_kernel_kind */
  int32_t impl;            /* 0 = auto; 1 = thread-per-chain registers; 2 = block-per-chain smem; 3 = tcgen05
                            * (DIAMONDS, fp32, ARWMH with adapt = 1: auto picks tcgen05 above two chains per SM);
                            * 4 = thread-per-chain registers behind a per-SM work queue (fp32 ARWMH; same trajectories as 1,
                            * auto picks it when the chains give the schedulers an uneven number of warps) */
} amcmc_run_args;

/* ARWMH.init (arwmh.py:84-138).  If use_given_z == 0 draws q0 ~ U(-init_radius, init_radius)^d
 * per chain (NumPyro init_to_uniform, radius 2) from the Philox stream; otherwise keeps
 * state->z.  Then U0 = potential(q0); loc = q0; scale = I; log_step_size = 0;
 * mean_accept_prob = 0; as_change = 0; i = 0. */
int amcmc_arwmh_init(const amcmc_model* m, amcmc_state* state, uint64_t seed, int64_t chain_offset,
                     double init_radius, int use_given_z, void* stream);

/* ARWMH.sample (arwmh.py:140-207) fused over args->n_steps steps for all chains,
 * with fori_collect-style thinned collection.  Updates *state in place (state->i too). */
int amcmc_arwmh_run(const amcmc_model* m, amcmc_state* state, const amcmc_run_args* args, void* stream);

/* potential_fn(z) (arwmh.py:121,170) for n points, q laid out [d][n]. */
int amcmc_potential(const amcmc_model* m, int64_t n, const void* q, void* out, void* stream);

/* Same as amcmc_arwmh_run but every pointer in *state and the out_* / normals /
 * uniforms pointers in *args are HOST buffers: state is copied to the device,
 * the fused steps run, and state + collected samples are copied back before
 * returning (this is what a non-CUDA caller, e.g. the reference's NumPy/JAX-CPU
 * scripts, would bind).  Device scratch is cached inside the model handle. */
int amcmc_arwmh_run_host(amcmc_model* m, amcmc_state* host_state, const amcmc_run_args* host_args);
/* The chunk plan of amcmc_arwmh_run_host: number of samples (thinning periods) the next launch collects when `remaining`
 * samples are left -- half of them, between ceil(128 / thinning) and ceil(2048 / thinning), never leaving a shorter stub.
 * The chain is bit-reproducible for a given plan (a launch boundary stores the proposal factor in the ABI's Cholesky form:
 * one rounding), so a device-pointer caller that wants the host entry's exact numbers cuts its launches the same way. */
int64_t amcmc_host_chunk_samples(int64_t remaining, int64_t thinning);
/* amcmc_arwmh_init with HOST buffers in *host_state (same contract; synchronises before returning). */
int amcmc_arwmh_init_host(amcmc_model* m, amcmc_state* host_state, uint64_t seed, int64_t chain_offset,
                          double init_radius, int use_given_z);

/* ---- Pooled adaptation (BASELINE.json configs[3]; NOT in the reference -- spec in DESIGN.md) ----------
 * All chains of a launch share ONE adaptation state (loc, scale, log_step_size).  Between windows of
 * K frozen steps the shared state moves by the reference's Robbins-Monro rule (arwmh.py:183-193) with
 * the chain-average of the innovation:  delta_c = x_c - loc,
 *     loc += g * mean_c(delta_c);  cov = (1-g) cov + g * mean_c(delta_c delta_c^T);  scale = chol(cov)
 *     log_step_size += g * (mean accept prob of the window - target);   g = window^-lr_decay
 * (C = 1, K = 1 reduces to the reference's per-step update).  Sufficient statistics are additive over
 * shards, so multi-GPU = one all-reduce (NCCL) of 2 + d + d(d+1)/2 doubles per window. */
typedef struct amcmc_pooled {
  int32_t dim;
  int32_t dtype;          /* same dtype as the chain state */
  void* loc;              /* device [d] */
  void* scale;            /* device [d(d+1)/2] packed lower triangle, row-major */
  void* log_step_size;    /* device [1] */
  double* cov;            /* device [d*d] float64 running covariance */
  int64_t window;         /* completed adaptation windows (n of the Robbins-Monro schedule) */
} amcmc_pooled;

/* args->n_steps frozen Metropolis steps for every chain with the shared proposal of *pool
 * (arwmh.py:161-178 with adapt_state fixed, i.e. the kernel of sample_Pnx :230-249).  Updates state->z,
 * potential_energy and mean_accept_prob (:= mean acceptance probability of each chain over this call).
 * diamonds/fp32 runs on the tcgen05 tensor-core path; other models broadcast *pool into the per-chain
 * adaptation arrays of *state and use the CUDA-core kernels. */
int amcmc_pooled_run(const amcmc_model* m, amcmc_state* state, const amcmc_pooled* pool, const amcmc_run_args* args,
                     void* stream);
/* out_stats (device, float64, ZEROED by this call then accumulated):
 *   [0] = chains, [1..d] = sum_c delta_k, [1+d .. 1+d+d(d+1)/2) = sum_c delta_i delta_j (j <= i, row-major),
 *   [1+d+d(d+1)/2] = sum_c mean_accept_prob_c.   Length 2 + d + d(d+1)/2. */
int amcmc_pooled_stats(const amcmc_state* state, const amcmc_pooled* pool, double* out_stats, void* stream);
/* Robbins-Monro update of *pool from (all-reduced) statistics; Cholesky on device; pool->window += 1.
 * A non-positive-definite covariance keeps the previous scale (the reference's NaN guard, arwmh.py:191). */
int amcmc_pooled_update(amcmc_pooled* pool, const double* stats, double lr_decay, double target_accept_prob,
                        void* stream);

/* Hardware self-test of the tcgen05 building blocks (UMMA descriptors, TMEM, mbarrier, TMA bulk copy)
 * used by the diamonds tensor-core path: d_out[128*256] = A[128 x K] * B[256 x K]^T with bf16 inputs
 * (row-major device arrays of uint16 bf16 bit patterns), fp32 accumulation.  K multiple of 16, <= 96.
 * scratch: 256*K*2 bytes of device memory (used when use_tma != 0).  No reference counterpart. */
int amcmc_selftest_umma(const void* a_bf16, const void* b_bf16, int K, void* scratch, float* d_out, int use_tma,
                        int swap_lbo_sbo, void* stream);

/* ---- sample-quality metrics (python/utils/evaluation.py), SURVEY 8f rank 4 ------------------------------------
 * Samples are row-major [n][d] float32 DEVICE arrays, like the reference's jnp arrays.  Functions with an
 * `out_host` argument write HOST memory and synchronise `stream` before returning. */

/* sum_ij exp(-gamma |x_i - y_j|^2): the three sums of mmd_heuristic (evaluation.py:286-291) and, with
 * skip_diagonal != 0 (x == y), the off-diagonal sums of mmd2_unbiased (:251-261).  Replaces gaussian_kernel(...).sum()
 * (:201-222) without materialising the n x m matrix. */
int amcmc_eval_kernel_sum(const float* x, int64_t n, const float* y, int64_t m, int d, double gamma, int skip_diagonal,
                          double* out_host, void* stream);
/* The three sums of mmd2_unbiased / mmd_heuristic (evaluation.py:246-263, :279-294) in one launch on the tensor cores:
 * out_host[0] = sum_{i != j} k(x_i, x_j), [1] = sum_{i != j} k(y_i, y_j), [2] = sum_ij k(x_i, y_j), k(a, b) = exp(-gamma |a - b|^2)
 * (add n resp. m for the diagonals of the biased estimate: k(a, a) = 1).  The cross terms are a bf16-split tcgen05 GEMM
 * (hi.hi + lo.hi + hi.lo, fp32 accumulation), exp in the epilogue.  d <= 32, else AMCMC_ERR_UNSUPPORTED (use
 * amcmc_eval_kernel_sum, any d, CUDA cores).  Synchronises `stream`. */
int amcmc_eval_mmd_sums(const float* x, int64_t n, const float* y, int64_t m, int d, double gamma, double* out_host, void* stream);
/* Median of all m*m squared pairwise distances of y (diagonal zeros and both orders included): the bandwidth
 * heuristic gamma = 4 / median (evaluation.py:283).  Exact radix select on the float32 keys, no m*m buffer. */
int amcmc_eval_sqdist_median(const float* y, int64_t m, int d, double* out_host, void* stream);
/* out[i*m + j] = |x_i - y_j|_ord (ord >= 1): scipy.spatial.distance_matrix(u, v, p=ord) of wasserstein_dist11_p
 * (evaluation.py:58).  `out` is a DEVICE array; asynchronous on `stream`. */
int amcmc_eval_cost_matrix(const float* x, int64_t n, const float* y, int64_t m, int d, double ord, float* out,
                           void* stream);
/* out_host[k] = mean_i x_ik^p, k < d: the moment estimates of pth_moment_rmse (evaluation.py:33-34). */
int amcmc_eval_moment(const float* x, int64_t n, int d, double p, double* out_host, void* stream);
/* scipy.optimize.linear_sum_assignment(cost_matrix) of wasserstein_dist11_p (evaluation.py:59) for a square n x n float32
 * DEVICE cost matrix: forward auction with epsilon-scaling on the costs quantised to 24-bit integers
 * c_ij = rint(cost_ij * (2^24 - 1) / max cost) (optimal for the quantised matrix; equal optimal cost as SciPy on it).
 * col_of_row [n] int32 DEVICE: the column matched to every row.  quantised: DEVICE [n*n] int32 or NULL, receives the
 * integer matrix.  out_host[3] (HOST, optional): sum of the float costs of the matching, sum of the integer costs,
 * auction rounds.  Synchronises `stream`. */
int amcmc_eval_assignment(const float* cost, int64_t n, int32_t* col_of_row, int32_t* quantised, double* out_host, void* stream);

/* wasserstein_sinkhorn (evaluation.py:69-97: OTT-JAX linear.solve on a PointCloud, `ot.ent_reg_cost`) for two uniform samples:
 * log-domain Sinkhorn on the DEVICE float32 cost matrix [n][m].  epsilon <= 0 selects 0.05 x mean cost.  threshold: L1 error of
 * the row marginal, checked every `inner_iterations`; at most `max_iterations` (OTT defaults as recalled: 1e-3, 10, 2000).
 * out_host[5] (HOST): ent_reg_cost = sum a f + sum b g + eps (1 - sum P), iterations, last marginal error, converged, epsilon.
 * f_out / g_out: optional DEVICE potentials [n] / [m].  Synchronises `stream`. */
int amcmc_eval_sinkhorn(const float* cost, int64_t n, int64_t m, double epsilon, double threshold, int max_iterations,
                        int inner_iterations, float* f_out, float* g_out, double* out_host, void* stream);

/* The reference's random stream on the GPU (python/kernels/arwmh.py:162-165,174: split(rng_key, 3), Normal().sample,
 * Uniform().sample with jax's threefry2x32 PRNG): for every chain, n_steps steps of draws from its JAX key, written in the
 * external-draws layout of amcmc_arwmh_run (normals[n_steps][dim][C], uniforms[n_steps][C], `dtype` elements holding float32
 * values).  keys: DEVICE uint32 [2][C], replaced by the keys after n_steps steps (ARWMHState.rng_key).  Asynchronous. */
int amcmc_jax_draws(uint32_t* keys, int64_t n_chains, int dim, int64_t n_steps, int dtype, void* normals, void* uniforms,
                    void* stream);

const char* amcmc_last_error(void);
int amcmc_version(void);

#ifdef __cplusplus
}
#endif
#endif /* AMCMC_H_ */
