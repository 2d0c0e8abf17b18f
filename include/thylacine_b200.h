/*
 * thylacine_b200.h -- C ABI of the B200-native posterior-evaluation hot path.
 *
 * The reference (gvonness/thylacine, Scala) has no FFI boundary; this header CREATES the seam at the
 * reference's L5/L6 trait surface.  Each entry point cites the reference interface
 * it replaces as path:line relative to /root/reference/src/main/scala/thylacine/.
 *
 * Conventions
 *   - plain pointers and sizes; no C++ or torch types; every call returns a thy_status (0 = ok);
 *   - all numbers are IEEE fp64; parameter batches are row-major "chain-major": X[b*d + j];
 *   - a handle is used by one caller at a time; distinct handles may be used concurrently;
 *   - the library owns all device memory; callers own host buffers; inputs are copied at add / finalize;
 *   - there is NO CPU fallback: every evaluation entry point returns THY_ERR_NO_DEVICE without a GPU;
 *   - *_dev variants take device pointers on the current CUDA device and are asynchronous on the model's
 *     stream: by default a blocking stream the handle creates at finalize (ordered with the legacy default
 *     stream like stream 0 itself, but capturable into CUDA graphs), or the one set by thy_model_set_stream.
 */
#ifndef THYLACINE_B200_H
#define THYLACINE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum thy_status {
  THY_OK = 0,
  THY_ERR_INVALID_ARGUMENT = 1, /* what the reference asserts: dimension mismatch, CI <= 0, hi <= lo, duplicates */
  THY_ERR_UNKNOWN_LABEL = 2,    /* IndexedCollection.scala:25-31 RuntimeException */
  THY_ERR_CUDA = 3,
  THY_ERR_NCCL = 4,
  THY_ERR_STATE = 5,            /* call order (e.g. evaluate before finalize, add after finalize) */
  THY_ERR_UNSUPPORTED = 6,
  THY_ERR_NO_DEVICE = 7
} thy_status;

typedef struct thy_model_s* thy_model_t;
typedef struct thy_hmc_s* thy_hmc_t;
typedef struct thy_lfm_s* thy_lfm_t;
typedef struct thy_slq_s* thy_slq_t;
typedef struct thy_surface_s* thy_surface_t;

/* ---- library ------------------------------------------------------------------------------------- */
const char* thy_version(void);
/* Message of the last failing call on this thread (never NULL). */
const char* thy_last_error(void);
/* Devices (a JVM cannot call the CUDA runtime itself): number of usable GPUs; bind the CALLING THREAD to one.  Every
 * handle lives on the device that was current when it was created / finalized and must be used from a thread bound to
 * that device. */
thy_status thy_device_count(int* out_count);
thy_status thy_set_device(int device);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches claim). */
uint64_t thy_kernel_launch_count(void);

/* ---- model construction --------------------------------------------------------------------------
 * Replaces the Scala case-class constructors; the model is immutable after thy_model_finalize (mirrors
 * the reference's immutable case classes).                                                            */
thy_status thy_model_create(thy_model_t* out);
thy_status thy_model_destroy(thy_model_t m);

/* GaussianPrior.fromConfidenceIntervals(label, values, confidenceIntervals)
 *   components/prior/GaussianPrior.scala:62-75; variance = (ci/2)^2, core/RecordedData.scala:59-66. */
thy_status thy_model_add_gaussian_prior_ci(thy_model_t m, const char* label, int dim,
                                           const double* values, const double* confidence_intervals);
/* GaussianPrior.fromCovarianceMatrix(label, values, covarianceMatrix)  GaussianPrior.scala:78-93
 *   covariance: dim x dim row-major, symmetric positive definite.                                     */
thy_status thy_model_add_gaussian_prior_cov(thy_model_t m, const char* label, int dim,
                                            const double* values, const double* covariance);
/* UniformPrior.fromBounds(label, maxBounds, minBounds)  components/prior/UniformPrior.scala:68-77
 *   (argument order max-then-min as in the reference); box is half-open [min, max),
 *   distributions/UniformDistribution.scala:61-66.                                                    */
thy_status thy_model_add_uniform_prior(thy_model_t m, const char* label, int dim,
                                       const double* max_bounds, const double* min_bounds);
/* CauchyPrior(label, values, confidenceIntervals)  components/prior/CauchyPrior.scala:50-63 */
thy_status thy_model_add_cauchy_prior(thy_model_t m, const char* label, int dim,
                                      const double* values, const double* confidence_intervals);

/* Observation distribution of a likelihood (distributions/{Gaussian,Cauchy,Uniform}Distribution.scala). */
typedef enum thy_distribution { THY_DIST_GAUSSIAN = 0, THY_DIST_CAUCHY = 1, THY_DIST_UNIFORM = 2 } thy_distribution;

/* Forward-model families.  A JVM closure (NonLinearForwardModel.of(evaluation = ...),
 * forwardmodel/NonLinearForwardModel.scala:44-85) cannot run on a GPU; the device-side registry is
 * f(theta) = link(A . theta + offset) applied element-wise.  THY_LINK_IDENTITY with THY_JAC_CONSTANT is the
 * reference's LinearForwardModel (forwardmodel/LinearForwardModel.scala:76-114).                        */
typedef enum thy_link {
  THY_LINK_IDENTITY = 0, THY_LINK_TANH = 1, THY_LINK_EXP = 2, THY_LINK_SQUARE = 3,
  THY_LINK_SOFTPLUS = 4, THY_LINK_SIGMOID = 5
} thy_link;

/* Link ids >= THY_LINK_USER_BASE are user-defined families registered on a model with thy_model_register_link. */
#define THY_LINK_USER_BASE 100

typedef enum thy_jacobian {
  THY_JAC_CONSTANT = 0,           /* LinearForwardModel.jacobianAt: the coefficient blocks (:65-71)            */
  THY_JAC_ANALYTIC = 1,           /* NonLinearForwardModel.of(evaluation, jacobian, ...) (:64-85): link'(z) A   */
  THY_JAC_FINITE_DIFFERENCE = 2   /* core/computation/FiniteDifferenceJacobian.scala:30-55, forward difference, */
                                  /* all d_total+1 perturbed models evaluated in one launch                     */
} thy_jacobian;

typedef struct thy_likelihood_desc {
  int distribution;               /* thy_distribution                                                          */
  int link;                       /* thy_link                                                                  */
  int jacobian;                   /* thy_jacobian (THY_JAC_CONSTANT requires THY_LINK_IDENTITY)                */
  int n;                          /* range dimension (number of observations)                                  */
  int nblocks;                    /* labelled coefficient blocks (IndexedMatrixCollection)                     */
  const char* const* labels;      /* [nblocks] prior labels this forward model reads                           */
  const int* dims;                /* [nblocks] block widths; must equal the prior dimensions                   */
  const double* const* blocks;    /* [nblocks] each n x dims[i], row-major (row = observation)                 */
  const double* offset;           /* [n] or NULL: LinearForwardModel vectorOffset (:103-104)                   */
  const double* measurements;     /* [n]  Gaussian/Cauchy: observed values;  Uniform: upperBounds              */
  const double* uncertainties;    /* [n]  Gaussian/Cauchy: confidence intervals (sigma = ci/2); Uniform: lowerBounds */
  double differential;            /* finite-difference step (NonLinearForwardModel.of(differential = ...))     */
} thy_likelihood_desc;

/* NonLinearForwardModel.of(evaluation = <closure>, ...) (forwardmodel/NonLinearForwardModel.scala:44-85) for a family the
 * registry above does not hold: the link -- f(theta)_i = link((A theta + offset)_i) -- as CUDA C source, an expression in
 * the variable `z` (double), e.g. "z / (1.0 + fabs(z))"; deriv_expr is its derivative (z, and the value f = link(z), are in
 * scope), e.g. "1.0 / ((1.0 + fabs(z)) * (1.0 + fabs(z)))", or NULL.  Compiled with NVRTC for sm_100a (THY_ERR_INVALID_ARGUMENT
 * carries the compiler log; THY_ERR_UNSUPPORTED when libnvrtc.so.12 cannot be loaded) into the element-wise step between
 * the two tensor-core GEMMs of the wide path.  *out_link is the id to put in thy_likelihood_desc.link of likelihoods added
 * to THIS model afterwards.  With THY_JAC_ANALYTIC the derivative expression is required; with THY_JAC_FINITE_DIFFERENCE
 * the link's own forward difference (link(z + h) - link(z)) / h, h = differential, stands in for it
 * (core/computation/FiniteDifferenceJacobian.scala:30-55 has the same limit for this family).  Gaussian / Cauchy only. */
thy_status thy_model_register_link(thy_model_t m, const char* eval_expr, const char* deriv_expr, int* out_link);

/* GaussianLinearLikelihood.of / GaussianLikelihood.apply / CauchyLikelihood.apply,.of / UniformLikelihood.apply
 *   components/likelihood/{GaussianLinearLikelihood.scala:62-83, GaussianLikelihood.scala:55-67,
 *   CauchyLikelihood.scala:59-92, UniformLikelihood.scala:63-73}.  Must follow the priors it refers to. */
thy_status thy_model_add_likelihood(thy_model_t m, const thy_likelihood_desc* desc);

/* Options (set before finalize). */
typedef enum thy_option {
  /* 1 (default) = reproduce CauchyDistribution.scala:70-79 literally (gradient has the wrong sign);
   * 0 = mathematically correct sign.  Samplers built on a model always use the correct sign.             */
  THY_OPT_CAUCHY_REFERENCE_SIGN = 0,
  /* Finite-difference evaluation: 0 (default) = every perturbed point theta + h e_j goes through the full
   * forward model (the reference's arithmetic; the perturbed dot products reuse the unperturbed prefix sums, which
   * is the same rounding sequence); 1 = incremental z + h A[:,j] (same limit, fewer flops); 2 = as 0 with every
   * perturbed dot product recomputed from its first term (diagnostics: bit-identical to 0, ~10x slower).         */
  THY_OPT_FD_INCREMENTAL = 1,
  /* Kernel selection for linear-Gaussian terms: 0 = auto, 1 = force generic tiled kernels,
   * 2 = force fused DMMA kernel (error if the shape is unsupported), 3 = force streaming (small-batch) kernel,
   * 4 = force the wide DMMA GEMM pair (any width / link; not for uniform observations or FD Jacobians),
   * 5 = as 2 but always the variant that streams A through a TMA ring (2 and auto switch to the variant that keeps A
   * resident in shared memory when it fits and the batch covers the SMs; reported as "resident_dmma").        */
  THY_OPT_KERNEL_PATH = 2,
  /* When a posterior is one linear term over the whole parameter vector with per-coordinate priors, the fused DMMA
   * kernel can evaluate the priors in its output step (one launch per evaluation, no read-modify-write of the
   * outputs): 0 = never, 1 (default) = when the cost model says the saved prior pass outweighs the slower output
   * step (small terms, e.g. a collapsed posterior), 2 = always.                                                  */
  THY_OPT_FUSE_PRIOR = 3
} thy_option;
thy_status thy_model_set_option(thy_model_t m, int option, int value);

/* Sorts labels ascending (posterior/Posterior.scala:40-48, core/GenericIdentifier.scala:29-30), builds the flat
 * layout, validates every likelihood label against the priors, uploads constants to the current device.   */
thy_status thy_model_finalize(thy_model_t m);
/* Gaussian collapse (SURVEY.md 8(d) reduced-flop form).  Builds a NEW finalized model with the same priors and label
 * layout in which all GaussianLinearLikelihood terms (likelihood/GaussianLinearLikelihood.scala:62-83) are replaced by
 * ONE d'-observation linear-Gaussian term (A' = L^T, L L^T = sum A^T Sigma^-1 A computed on the device) that has the
 * same log-density and gradient at every point: 4 d'^2 instead of 4 n d flops per evaluation.  Every engine runs on
 * the collapsed model unchanged.  THY_ERR_UNSUPPORTED for non-Gaussian / non-linear terms or a singular Gram matrix.
 * The caller owns *out (thy_model_destroy). */
thy_status thy_model_collapse(thy_model_t m, thy_model_t* out);
/* GaussianAnalyticPosterior (components/posterior/GaussianAnalyticPosterior.scala:141-169: rawDistribution; :72-79
 * mean / covariance / maxLogPdf): Lambda = Sigma_p^-1 + A^T Sigma_d^-1 A on the device, Cholesky on the host.
 * Gaussian priors and GaussianLinearLikelihoods only.  Any output may be NULL.  out_log_evidence = ln of the integral
 * of the unnormalised posterior (the analytic evidence the SLQ tests compare against). */
thy_status thy_model_gaussian_posterior(thy_model_t m, double* out_mean, double* out_covariance, double* out_precision,
                                        double* out_log_evidence, double* out_max_logpdf);
/* Handle-local error text: message of the last failing call made on this model or on a sampler built on it, from any
 * thread (thy_last_error is per thread).  Copies at most capacity - 1 characters. */
thy_status thy_model_last_error(thy_model_t m, char* out, size_t capacity);
/* domainDimension (Posterior.scala:37-38) */
thy_status thy_model_dimension(thy_model_t m, int* out_dim);
/* offset and width of a label inside the flat vector (ModelParameterContext.scala:42-80) */
thy_status thy_model_label_layout(thy_model_t m, const char* label, int* out_offset, int* out_dim);
/* Stream used by the *_dev entry points and by samplers built on this model (a cudaStream_t). */
thy_status thy_model_set_stream(thy_model_t m, void* cuda_stream);
thy_status thy_model_synchronize(thy_model_t m);

/* ---- evaluation ----------------------------------------------------------------------------------
 * ModelParameterPdf.logPdfAt / logPdfGradientAt  (core/values/modelparameters/ModelParameterPdf.scala:33-46,84-95),
 * Posterior.logPdfAt / logPdfGradientAt (posterior/Posterior.scala:50-73), batched over B points.       */
thy_status thy_logpdf_batch(thy_model_t m, int64_t B, const double* X, double* out_logp);
thy_status thy_logpdf_grad_batch(thy_model_t m, int64_t B, const double* X, double* out_logp, double* out_grad);
/* Pipelined host-buffer variant: returns as soon as the copies and kernels are queued (pinned host buffers make the
 * copies asynchronous).  Two staging sets alternate, so call k+1's input copy and call k-1's output copy run under call
 * k's kernels.  *out_ticket identifies the call; the outputs (and X) belong to the library until thy_model_wait returns
 * for that ticket (ticket <= 0: wait for every outstanding call).  A ticket two or more calls old has always completed.
 * out_grad may be NULL (value only). */
thy_status thy_logpdf_grad_batch_async(thy_model_t m, int64_t B, const double* X, double* out_logp, double* out_grad,
                                       int64_t* out_ticket);
thy_status thy_model_wait(thy_model_t m, int64_t ticket);
/* Device-pointer variants (zero-copy; out_grad may be NULL for value only). */
thy_status thy_logpdf_grad_batch_dev(thy_model_t m, int64_t B, const double* X_dev, double* logp_dev, double* grad_dev);
/* Posterior.samplePriors (Posterior.scala:64-65): B independent prior draws with the counter-based RNG. */
thy_status thy_sample_priors(thy_model_t m, int64_t B, uint64_t rng_seed, double* out_X);
thy_status thy_sample_priors_dev(thy_model_t m, int64_t B, uint64_t rng_seed, double* X_dev);
/* Name of the kernel path the last evaluation on this model used ("fused_dmma", "stream", "generic", ...). */
const char* thy_model_last_kernel_path(thy_model_t m);

/* ---- Hamiltonian MCMC ------------------------------------------------------------------------------
 * HmcmcSampledPosterior.apply(hmcmcConfig, posterior, telemetryUpdateCallback, seed)
 *   components/posterior/HmcmcSampledPosterior.scala:63-75; engine sampling/hmcmc/HmcmcEngine.scala:55-214;
 *   config fields config/HmcmcConfig.scala:20-25 (same names).  n_chains independent chains run in lock-step. */
typedef struct thy_hmcmc_config {
  int stepsBetweenSamples;
  int stepsInDynamicsSimulation;
  int warmupStepCount;
  double dynamicsSimulationStepSize;
} thy_hmcmc_config;

/* seed_points: host [n_chains * d] starting points (the reference's `seed`), or NULL => prior draws
 * (HmcmcEngine.scala:183-188).  Samplers always use the mathematically correct Cauchy gradient sign.    */
thy_status thy_hmc_create(thy_model_t m, const thy_hmcmc_config* cfg, int64_t n_chains, uint64_t rng_seed,
                          const double* seed_points, thy_hmc_t* out);
thy_status thy_hmc_destroy(thy_hmc_t h);
/* Replay one trajectory from a CUDA graph (default on; off when the model was given the legacy stream 0 itself). */
thy_status thy_hmc_set_use_graph(thy_hmc_t h, int enable);
/* burnIn (HmcmcEngine.scala:180-197): (re)initialise the chains and run warmupStepCount + 1 trajectories. */
thy_status thy_hmc_burn_in(thy_hmc_t h);
/* Run n_trajectories more trajectories on every chain (each = stepsInDynamicsSimulation leapfrog steps + one
 * Metropolis test); asynchronous on the model's stream.  The benchmark's "leapfrog steps/s" loop.          */
thy_status thy_hmc_run_trajectories(thy_hmc_t h, int n_trajectories);
/* ModelParameterSampler.sample(n) (sampling/ModelParameterSampler.scala:36-37) for every chain:
 * out_samples[(chain * n_per_chain + k) * d + j].  stepsBetweenSamples + 1 trajectories per block; a block in which
 * a chain accepted nothing emits no sample for that chain (the reference's Set semantics, HmcmcEngine.scala:149-152).
 * rerun_burn_in != 0 reproduces the reference (burn-in re-executes on every sample call, :205-207).          */
thy_status thy_hmc_sample(thy_hmc_t h, int n_per_chain, double* out_samples, int rerun_burn_in);
thy_status thy_hmc_sample_dev(thy_hmc_t h, int n_per_chain, double* out_samples_dev, int rerun_burn_in);
/* HmcmcTelemetryUpdate fields (core/telemetry/HmcmcTelemetryUpdate.scala:20-25) per chain; any pointer may be NULL. */
thy_status thy_hmc_stats(thy_hmc_t h, int64_t* out_jump_attempts, int64_t* out_jump_acceptances,
                         double* out_hamiltonian_differential);
/* Current positions [n_chains * d] and log-posteriors [n_chains] (checkpoint / warm start). */
thy_status thy_hmc_get_state(thy_hmc_t h, double* out_x, double* out_logp);
/* Checkpoint / resume (the reference carries its chain position in a TxnVar, HmcmcSampledPosterior.scala:36,55-56, and
 * has no seedable RNG; here the state of a run is the positions plus the trajectory counter that keys every draw):
 * set_state uploads positions [n_chains * d], re-evaluates log-posterior and gradient there and sets the counter, after
 * which the chains continue exactly as the run that was checkpointed (same handle configuration and rng_seed). */
thy_status thy_hmc_set_state(thy_hmc_t h, const double* x, uint64_t trajectory_counter);
thy_status thy_hmc_get_trajectory_counter(thy_hmc_t h, uint64_t* out_counter);
/* When the posterior is one fused-DMMA launch (single linear term + per-coordinate priors) a leapfrog step can be ONE
 * kernel: the gradient kernel's output step applies the momentum kicks and the drift, and on the last step the
 * Hamiltonian, the Metropolis test and the state update (HmcmcEngine.scala:55-172).  mode 2: always do that; 1 (default):
 * when the cost model says it is faster (short gradient kernels: config 1, collapsed models); 0: separate elementwise
 * kernels around the gradient evaluation (always used for models outside that shape). */
thy_status thy_hmc_set_fused_steps(thy_hmc_t h, int mode);

/* ---- Leapfrog MCMC (gradient-free ensemble sampler) -----------------------------------------------
 * LeapfrogMcmcSampledPosterior.of(leapfrogMcmcConfig, distanceCalculation, posterior, setCb, updateCb, seed)
 *   components/posterior/LeapfrogMcmcSampledPosterior.scala:72-106; engine sampling/leapfrog/LeapfrogMcmcEngine.scala:88-294;
 *   config fields config/LeapfrogMcmcConfig.scala:20-24.  The reference's distanceCalculation is a JVM closure
 *   (LeapfrogMcmcSampledPosterior.scala:39); the device registry offers the named distances below.            */
typedef struct thy_leapfrog_mcmc_config {
  int stepsBetweenSamples;
  int warmupStepCount;
  int samplePoolSize;
} thy_leapfrog_mcmc_config;
typedef enum thy_distance {
  THY_DISTANCE_EUCLIDEAN = 0, THY_DISTANCE_SQUARED_EUCLIDEAN = 1, THY_DISTANCE_MANHATTAN = 2
} thy_distance;

/* Draws the pool from the priors (slot 0 <- seed_point if given, LeapfrogMcmcEngine.scala:237-253), evaluates the
 * pool's logPdfs and runs the burn-in (warmupStepCount proposals per pool member) before returning, as `.of` does. */
thy_status thy_lfm_create(thy_model_t m, const thy_leapfrog_mcmc_config* cfg, int distance, uint64_t rng_seed,
                          const double* seed_point, thy_lfm_t* out);
thy_status thy_lfm_destroy(thy_lfm_t h);
/* One sweep = one proposal for every pool member (two half-pool passes); asynchronous on the model's stream. */
thy_status thy_lfm_run_sweeps(thy_lfm_t h, int n_sweeps);
/* sample(n): each returned point has been proposed stepsBetweenSamples times since it was last handed out
 * (LeapfrogMcmcEngine.scala:164-170,196-208).  out_samples: host [n * d].                                  */
thy_status thy_lfm_sample(thy_lfm_t h, int n, double* out_samples);
thy_status thy_lfm_stats(thy_lfm_t h, uint64_t* out_proposals, uint64_t* out_acceptances);
thy_status thy_lfm_get_pool(thy_lfm_t h, double* out_x, double* out_logp);
/* Replace the pool (host [samplePoolSize * d]): warm start near the mode, or resume from thy_lfm_get_pool's output. */
thy_status thy_lfm_set_pool(thy_lfm_t h, const double* x);

/* ---- multi-GPU plumbing ---------------------------------------------------------------------------
 * The reference has no distributed layer (single JVM process).  Here: one process per GPU; chains, pool members and
 * live points shard across ranks; NCCL over NVLink is used only for the SLQ global lowest-logL selection and for
 * sample-moment sums.  The 128-byte unique id is created on rank 0 and handed to the other ranks out of band
 * (torch.distributed object broadcast, a file, the JVM's own RPC ...).  world == 1 needs no NCCL at all.      */
typedef struct thy_comm_s* thy_comm_t;
thy_status thy_comm_unique_id(unsigned char out_id[128]);
thy_status thy_comm_create(int rank, int world, const unsigned char unique_id[128], thy_comm_t* out);
/* Single-process mode (one JVM driving the whole box): communicators for n_devices GPUs of this process at once
 * (ncclCommInitAll), out[i] = rank i on devices[i] (NULL: devices 0 .. n-1).  Use one host thread per GPU, each bound
 * with thy_set_device(devices[i]) before it builds its model / sampler and for every later call on them: collectives of
 * different ranks must be in flight concurrently, which one thread issuing them in turn cannot guarantee. */
thy_status thy_comm_create_all(int n_devices, const int* devices, thy_comm_t* out);
thy_status thy_comm_destroy(thy_comm_t c);
thy_status thy_comm_rank(thy_comm_t c, int* out_rank, int* out_world);
/* In-place sum of a device fp64 buffer over all ranks on the given stream (no-op for world == 1 / NULL comm). */
thy_status thy_comm_all_reduce_sum(thy_comm_t c, double* buf_dev, int64_t count, void* cuda_stream);

/* ---- sample moments ---------------------------------------------------------------------------------
 * Not in the reference (its tests compare sample means by hand, GaussianAnalyticPosteriorSpec.scala:34-50); required
 * for the statistical parity checks.  acc_dev holds [count, sum x (d), sum x x^T (d*d)] = 1 + d + d*d doubles.   */
thy_status thy_moments_reset_dev(thy_model_t m, double* acc_dev);
/* acc += moments of X_dev[n][d]; rows whose first coordinate is NaN are skipped (unfilled HMC sample slots). */
thy_status thy_moments_accumulate_dev(thy_model_t m, int64_t n, const double* X_dev, double* acc_dev);
/* All-reduce over comm (may be NULL), then count, mean[d] and covariance[d*d] (unbiased) to host buffers. */
thy_status thy_moments_finish(thy_model_t m, thy_comm_t comm, double* acc_dev, double* out_count, double* out_mean,
                              double* out_cov);

/* ---- SLQ (stochastic Lebesgue quadrature; nested sampling) ------------------------------------------
 * SlqIntegratedPosterior.of(slqConfig, posterior, callbacks, seedsSpec)
 *   components/posterior/SlqIntegratedPosterior.scala:87-125; engine integration/slq/SlqEngine.scala:107-469;
 *   config fields config/SlqConfig.scala:20-28 (same names).  Live points shard across the ranks of `comm`.   */
typedef struct thy_slq_config {
  int poolSize;                       /* live points (global)                                                  */
  int abscissaNumber;                 /* independent prior-mass simulations (QuadratureAbscissa.scala:33-73)   */
  double domainScalingIncrement;      /* multiplicative step of the proposal-scale controller                  */
  double targetAcceptanceProbability;
  int sampleParallelism;              /* live points retired and replaced per round (reference: fibres in flight) */
  int maxIterationCount;
  int minIterationCount;
  /* extensions (0 => default) */
  int constrainedWalkSteps;           /* constrained-walk steps per replacement (default 20)                   */
  int keepRetiredPoints;              /* keep retired positions for thy_slq_sample (single GPU)                */
  int disableFusedWalk;               /* 1 = always use the per-step kernels (diagnostics; same random numbers) */
  int referenceProposals;             /* 1 = the reference's own proposal mechanism and one-at-a-time pool update: the
                                         cubical cover (PointInCubeCollection.scala:60-111, PointInCube.scala:84-117,
                                         PointInInterval.scala:40-109, QuadratureDomainTelemetry.scala:27-110,
                                         SlqEngine.scala:120-197,216-247,309-345).  Small pools, single GPU: geometry and
                                         pool bookkeeping on the host, every logPdf on the device in batches of
                                         sampleParallelism, quadrature on the device.  Ranks by the full posterior logPdf
                                         as the reference does.  thy_slq_run only.                                  */
} thy_slq_config;

typedef struct thy_slq_stats {
  int64_t iterations;                 /* points retired during the live phase (SlqTelemetryUpdate.iterationCount) */
  int64_t total_retired;              /* + the drained pool (SlqEngine.scala:171-180)                          */
  int64_t rounds;
  int64_t likelihood_evaluations;     /* walk steps of all ranks (global)                                          */
  double log_evidence_mean;           /* mean over abscissas of ln Z                                           */
  double log_evidence_std;            /* spread over abscissas                                                 */
  double neg_entropy_mean;            /* SlqTelemetryUpdate.negEntropyAvg                                      */
  double min_logpdf;                  /* SlqTelemetryUpdate.samplePoolMinimumLogPdf at the end of the live phase */
  double walk_scale;                  /* SlqTelemetryUpdate.domainVolumeScaling analogue                       */
  double walk_acceptance;
} thy_slq_stats;

/* seeds: host [n_seeds * d] points placed in the first pool slots of rank 0 (SlqEngine.scala:384-389,409), or NULL. */
thy_status thy_slq_create(thy_model_t m, const thy_slq_config* cfg, uint64_t rng_seed, const double* seeds, int n_seeds,
                          thy_comm_t comm, thy_slq_t* out);
/* Collective when the handle was built on a communicator with more than one rank (every rank destroys its handle). */
thy_status thy_slq_destroy(thy_slq_t h);
/* rebuildSampleSimulation (SlqIntegratedPosterior.scala:76-80): run to convergence / maxIterationCount, drain. */
thy_status thy_slq_run(thy_slq_t h, thy_slq_stats* out_stats);
/* Run at most n_rounds more rounds of the live phase without draining (the benchmark's loop); *out_converged set. */
thy_status thy_slq_run_rounds(thy_slq_t h, int n_rounds, int* out_converged);
/* The calling rank's live points (host out_x[count * d], out_loglik[count]; either may be NULL) -- the reference's
 * `samples` map, SlqEngine.scala:120-143.  out_count = live points held by this rank. */
thy_status thy_slq_live_points(thy_slq_t h, int64_t* out_count, double* out_x, double* out_loglik);
/* Retired log-likelihoods in retirement order; *out_count receives the total available. */
thy_status thy_slq_retired(thy_slq_t h, int64_t max_count, double* out_loglik, int64_t* out_count);
/* ln X_i of ONE simulated abscissa (0 <= abscissa < abscissaNumber) for the retired points, in retirement order: the
 * `abscissas` the reference keeps beside its logPdf results (SlqEngine.scala:283-297, QuadratureAbscissa.scala:33-73).
 * Replayed on the device from the counter RNG; nothing is stored during the run. */
thy_status thy_slq_retired_log_x(thy_slq_t h, int abscissa, int64_t max_count, double* out_log_x, int64_t* out_count);
/* integrate(integrand) (SlqEngine.scala:460-466 -> QuadratureIntegrator.getIntegrationStats, QuadratureIntegrator.scala:64-83).
 * The integrand is a host callback (the reference's is a JVM closure BigDecimal => BigDecimal); it receives
 * exp(l_i - shift).  Per abscissa a: I_a = sum_i w_ai f(exp(l_i - shift)), Z_a = sum_i w_ai exp(l_i - shift);
 * *out_reference_statistic = mean_a (I_a / Z_a - ln Z_a) -- literally what the reference returns -- and
 * *out_integral = mean_a I_a; out_per_abscissa_integral[abscissaNumber] = I_a.  Any output may be NULL.
 * log_shift: NaN selects the reference's (max l + min l) / 2 (BigDecimal cannot overflow; fp64 can, so pass max l for
 * wide-ranging log-likelihoods). */
typedef double (*thy_slq_integrand)(double scaled_likelihood, void* user);
thy_status thy_slq_integrate(thy_slq_t h, thy_slq_integrand integrand, void* user, double log_shift,
                             double* out_reference_statistic, double* out_integral, double* out_per_abscissa_integral);
/* Phase trace: while enabled, rounds run as plain launches with CUDA events between the phases.  out_us[0..4] are the
 * consecutive phases of a round on the main stream (all-gather of the ranks' keys; radix select + flags + retired
 * values; walk + adapt; wait for the quadrature stream; round control), mean microseconds per traced round; out_us[5] the
 * pool-spread kernels, which run on a side stream beside phases 0-1; out_us[6]: bit 0 = untraced full rounds replay from
 * a CUDA graph, bit 1 = new keys travel over NVLink peer memory instead of an NCCL all-gather (phase 0 is then the wait
 * for the peers' flags + the local apply). */
/* Load balance of the sharded live pool: mean and largest number of points THIS rank had to replace in one round (the
 * K globally lowest points are wherever they are; a round lasts as long as its busiest rank's walk). */
thy_status thy_slq_load_balance(thy_slq_t h, double* out_mean_local, double* out_max_local);
thy_status thy_slq_set_trace(thy_slq_t h, int enable);
thy_status thy_slq_phase_times(thy_slq_t h, double out_us[7], int64_t* out_rounds);
/* Per-abscissa ln Z_a and H_a of the quadrature so far (QuadratureIntegrator.scala:29-62); [abscissaNumber] each. */
thy_status thy_slq_abscissa_results(thy_slq_t h, double* out_log_evidence, double* out_neg_entropy);
/* sample(n) (SamplingSimulation.scala:41-101): every draw picks one simulated abscissa uniformly, then a retired point with
 * probability proportional to that abscissa's trapezoid area; needs keepRetiredPoints (single GPU). */
thy_status thy_slq_sample(thy_slq_t h, int n, uint64_t rng_seed, double* out_samples);

/* ---- CartesianSurface (visualisation/CartesianSurface.scala:30-200) --------------------------------------
 * CartesianSurface.of(xAbscissa, yAbscissa, callbacks): a grid of values[ix * ny + iy], zero-initialised (:85-88).
 * add_samples = addSamples(abscissa, samples, ds, kernelVariance) (:129-150): every sample curve (abscissa_i, value_i)
 * is resampled at arc-length spacing ds (SimplexChain.scala:21-33, Simplex.scala:25-47; the remainder carries across
 * segments; the point the reference's getPoints drops is dropped here too) and every grid point gains
 * sum_p exp(-0.5 * ((px - gx)^2 / xScale^2 + (py - gy)^2 / yScale^2) / kernelVariance)   (:110-127, GraphPoint.scala:26-31).
 * results = getResults (:152-157): normalise_columns != 0 divides every x column by its trapezoid integral over y
 * when that is positive (:53-83, MathOps.scala:26-55; repeated y values or ny < 2 -> THY_ERR_INVALID_ARGUMENT where
 * the reference throws).  samples: host [n_samples * m] row-major; *_dev takes a device pointer (e.g. a sampler's
 * output buffer) and synchronises the device first.  Progress callbacks are the shim's business. */
thy_status thy_surface_create(int nx, const double* x_abscissa, int ny, const double* y_abscissa, thy_surface_t* out);
thy_status thy_surface_destroy(thy_surface_t s);
thy_status thy_surface_add_samples(thy_surface_t s, int m, const double* abscissa, int64_t n_samples, const double* samples,
                                   double ds, double kernel_variance);
thy_status thy_surface_add_samples_dev(thy_surface_t s, int m, const double* abscissa, int64_t n_samples,
                                       const double* samples_dev, double ds, double kernel_variance);
thy_status thy_surface_results(thy_surface_t s, int normalise_columns, double* out_values);

/* Measurement hook: when enabled, a CUDA-event pair is recorded on the model's stream around the dominant kernel of
 * every evaluation; thy_model_kernel_time synchronises, returns the summed duration and launch count, and resets. */
thy_status thy_model_enable_timing(thy_model_t m, int enable);
thy_status thy_model_kernel_time(thy_model_t m, double* out_total_ms, int64_t* out_launches);

#ifdef __cplusplus
}
#endif
#endif /* THYLACINE_B200_H */
