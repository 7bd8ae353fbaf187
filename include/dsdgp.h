/*
 * dsdgp.h -- C-ABI of libdsdgp.so: the B200-native doubly-stochastic DGP hot path.
 *
 * The reference (UCL-SML/Doubly-Stochastic-DGP) has no FFI: its boundary is the Python object API
 * called by GPflow's Model/optimiser machinery.  This header is the boundary BASELINE.json's
 * north_star defines ("Python host over a thin C-ABI (ctypes)"); each entry point names the
 * reference interface it replaces (file:line relative to /root/reference).  The ctypes binding a
 * maintainer would add is shown in INTEGRATION.md and implemented in
 * doubly-stochastic-dgp_b200/doubly_stochastic_dgp/_lib.py.
 *
 * Conventions: every function returns 0 on success or a negative DSDGP_ERR_* code; no C++
 * exception crosses the boundary; dsdgp_last_error() returns a thread-local message.  All arrays
 * are row-major, layouts exactly the reference's: X (N,D_in), Y (N,D_y), Z (M,D_in), q_mu (M,D_out),
 * q_sqrt (D_out,M,M) lower-triangular dense, z / F / Fmean / Fvar (S,N,D_out).
 * Data (X, Y, z, F*) is float32 (the arithmetic type of the per-row kernels); parameters cross the
 * boundary as float64 (the reference's float_type) and are held in fp32 on the device, with the
 * per-step M x M factorisations done in fp64.
 * A ctx is bound to one device and is not thread-safe.  Work is enqueued on the ctx's stream;
 * calls that return a host scalar synchronise.
 */
#ifndef DSDGP_H
#define DSDGP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define DSDGP_API __attribute__((visibility("default")))
#else
#define DSDGP_API
#endif

#define DSDGP_MAX_LAYERS 16

/* error codes */
#define DSDGP_OK 0
#define DSDGP_ERR_INVALID (-1)      /* bad argument / shape */
#define DSDGP_ERR_CUDA (-2)         /* CUDA runtime error (message has the detail) */
#define DSDGP_ERR_NOT_PD (-3)       /* Kuu + jitter*I not positive definite (tf.cholesky would raise) */
#define DSDGP_ERR_NCCL (-4)
#define DSDGP_ERR_UNSUPPORTED (-5)

/* kernels: gpflow.kernels.RBF / Matern52 (call sites layers.py:161,171,184,213) */
#define DSDGP_KERN_RBF 0
#define DSDGP_KERN_MATERN52 1
/* mean functions: gpflow.mean_functions.Zero / Identity / Linear (layers.py:219) */
#define DSDGP_MEAN_ZERO 0
#define DSDGP_MEAN_IDENTITY 1
#define DSDGP_MEAN_LINEAR 2
/* likelihoods: gpflow.likelihoods.Gaussian / MultiClass(RobustMax) / Bernoulli(probit) (utils.py:88-93;
 * tests/test_dgp.py:40-60) */
#define DSDGP_LIK_GAUSSIAN 0
#define DSDGP_LIK_MULTICLASS 1
#define DSDGP_LIK_BERNOULLI 2

/* parameter fields for set/get_param, get_grad */
#define DSDGP_F_Z 0
#define DSDGP_F_Q_MU 1
#define DSDGP_F_Q_SQRT 2
#define DSDGP_F_LENGTHSCALES 3      /* 1 value, or D_in values when ard */
#define DSDGP_F_VARIANCE 4
#define DSDGP_F_MEAN_W 5            /* Linear mean: (D_in,D_out), fixed (layer_initializations.py:41-42) */
#define DSDGP_F_MEAN_B 6
#define DSDGP_F_LIK_VARIANCE 7      /* layer = -1 */
#define DSDGP_F_WHITE_VARIANCE 8    /* variance of the White term of a Sum(kernel, White) layer kernel (kernel_white = 1) */

/* flags */
#define DSDGP_FLAG_DEVICE_PTRS 1u   /* X, Y, zs, outputs are device pointers (default: host) */
#define DSDGP_FLAG_NO_SYNC 2u       /* train_step: do not wait for / return the ELBO */

typedef struct dsdgp_ctx dsdgp_ctx;

/* One SVGP_Layer (layers.py:122-165). */
typedef struct {
    int M;        /* num_inducing */
    int D_in;     /* kern.input_dim */
    int D_out;    /* num_outputs */
    int kernel;   /* DSDGP_KERN_* */
    int ard;      /* 0: one lengthscale, 1: D_in lengthscales */
    int white;    /* layers.py:124 `white` */
    int mean;     /* DSDGP_MEAN_* */
    int kernel_white;    /* 1: the layer kernel is Sum(kernel, White): + variance_w I on Kuu and + variance_w on Kdiag
                          * (demos/demo_step_function.ipynb:111, demos/run_regression.py:65-66) */
    int input_prop_dim;  /* layers.py:105-117: the first input_prop_dim input columns are concatenated in front of the layer's
                          * samples; the next layer's D_in = input_prop_dim + D_out.  0: none */
} dsdgp_layer_desc;

/* The model (dgp.py:42-59 DGP_Base.__init__). */
typedef struct {
    int L;
    dsdgp_layer_desc layers[DSDGP_MAX_LAYERS];
    int likelihood;      /* DSDGP_LIK_* */
    int num_classes;     /* MultiClass only */
    int D_y;             /* columns of Y */
    double jitter;       /* gpflow settings.jitter (layers.py:162,171; utils.py:41) */
    int N_max;           /* largest minibatch / Xnew rows per call (per rank) */
    int S_max;           /* largest num_samples per call */
    int device;          /* CUDA ordinal */
} dsdgp_desc;

DSDGP_API const char* dsdgp_last_error(void);
DSDGP_API const char* dsdgp_version(void);

/* DGP_Base.__init__ (dgp.py:42-59): allocates every workspace once. */
DSDGP_API int dsdgp_create(dsdgp_ctx** out, const dsdgp_desc* desc);
DSDGP_API int dsdgp_destroy(dsdgp_ctx* ctx);

/* `layer.q_mu = ndarray` etc. (tests/test_dgp.py:91-92, demos/run_regression.py:72-74).
 * host float64 in/out; n = element count (checked). */
DSDGP_API int dsdgp_set_param(dsdgp_ctx* ctx, int layer, int field, const double* host, size_t n);
DSDGP_API int dsdgp_get_param(dsdgp_ctx* ctx, int layer, int field, double* host, size_t n);
/* dELBO/dparam of the last dsdgp_elbo_grad call (what tf.gradients(likelihood_tensor) gives). */
DSDGP_API int dsdgp_get_grad(dsdgp_ctx* ctx, int layer, int field, double* host, size_t n);

/* DGP_Base.propagate (dgp.py:61-76) with full_cov=False.  zs: NULL => in-kernel Philox keyed by
 * (seed, layer, s, n, d); else L pointers (entries may be NULL) to (S,N,D_out_l) float32.
 * Fs/Fmeans/Fvars: NULL or arrays of L pointers (entries may be NULL) receiving (S,N,D_out_l). */
DSDGP_API int dsdgp_propagate(dsdgp_ctx* ctx, const float* X, int N, int S, const float* const* zs,
                    uint64_t seed, float* const* Fs, float* const* Fmeans, float* const* Fvars,
                    unsigned flags);

/* DGP_Base.propagate with full_cov=True (dgp.py:61-76; layers.py:66-69,206-217; utils.py:43-51) -- behind
 * predict_f_full_cov / predict_all_layers_full_cov (dgp.py:104-114).  Per sample and output dimension the N x N
 * conditional covariance and a joint draw f = mean + chol(var + jitter I) z; float64 on the device (layers.py:68), float32
 * at the boundary.  Fs, Fmeans: (S,N,D_l); Fvars: (S,N,N,D_l).  zs as in dsdgp_propagate. */
DSDGP_API int dsdgp_propagate_full_cov(dsdgp_ctx* ctx, const float* X, int N, int S, const float* const* zs, uint64_t seed,
                             float* const* Fs, float* const* Fmeans, float* const* Fvars, unsigned flags);

/* DGP_Base.predict_y (dgp.py:116-119): likelihood.predict_mean_and_var (utils.py:110-114) of the last layer's
 * marginals, per sample.  mean, var: (S,N,D_last) float32.  Gaussian: (Fmean, Fvar + variance); MultiClass: class
 * probabilities p_k and p_k - p_k^2. */
DSDGP_API int dsdgp_predict_y(dsdgp_ctx* ctx, const float* X, int N, int S, const float* const* zs, uint64_t seed,
                    float* mean, float* var, unsigned flags);
/* DGP_Base.predict_density (dgp.py:121-126): logsumexp over the S samples of likelihood.predict_density - log S.
 * Y: (N,D_y) (class ids (N,1) for MultiClass); out: (N,D_y) Gaussian, (N,1) MultiClass. */
DSDGP_API int dsdgp_predict_density(dsdgp_ctx* ctx, const float* X, const float* Y, int N, int S, const float* const* zs,
                          uint64_t seed, float* out, unsigned flags);

/* BroadcastingLikelihood.{variational_expectations, predict_mean_and_var, predict_density} (utils.py:88-121) on
 * caller-supplied marginals of the last layer: Fmu, Fvar (S,N,D_last) float32, Y (N,D_y) (ignored by what = 1).
 * what = 0: variational_expectations -> out0 (S,N,Do);  1: predict_mean_and_var -> out0 mean, out1 var (S,N,D_last);
 * 2: predict_density -> out0 (S,N,Do).  Do = D_last (Gaussian, Bernoulli) or 1 (MultiClass).  S*N <= N_max*S_max. */
#define DSDGP_LIK_VE 0
#define DSDGP_LIK_PREDICT_MEAN_AND_VAR 1
#define DSDGP_LIK_PREDICT_DENSITY 2
DSDGP_API int dsdgp_likelihood_apply(dsdgp_ctx* ctx, int what, const float* Fmu, const float* Fvar, const float* Y, int S, int N,
                           float* out0, float* out1, unsigned flags);

/* DGP_Base._build_likelihood (dgp.py:92-98) == compute_log_likelihood(): ELBO scalar. */
DSDGP_API int dsdgp_elbo(dsdgp_ctx* ctx, const float* X, const float* Y, int N, int S, double num_data,
               const float* const* zs, uint64_t seed, unsigned flags, double* elbo);

/* ELBO and its gradient wrt every trainable (TF autodiff of dgp.py:92-98). With a
 * communicator (dsdgp_comm_init) X,Y are this rank's rows and ELBO/gradient are all-reduced. */
DSDGP_API int dsdgp_elbo_grad(dsdgp_ctx* ctx, const float* X, const float* Y, int N, int S, double num_data,
                    const float* const* zs, uint64_t seed, unsigned flags, double* elbo);

/* gpflow.training.AdamOptimizer(lr).minimize step (demos/run_regression.py:83): Adam on GPflow's
 * unconstrained variables (softplus for positive parameters, lower triangle for q_sqrt). */
DSDGP_API int dsdgp_adam_init(dsdgp_ctx* ctx, double lr, double beta1, double beta2, double eps);
/* One session.run(minimize_op): minibatch in, ELBO (before the update) out.  zs as in dsdgp_propagate
 * (NULL: Philox draws keyed by seed, the normal training mode). */
DSDGP_API int dsdgp_train_step(dsdgp_ctx* ctx, const float* X, const float* Y, int N, int S, double num_data,
                     const float* const* zs, uint64_t seed, unsigned flags, double* elbo);

/* DGP_Quad (dgp.py:129-166): the S "samples" are Gauss-Hermite nodes (passed as zs) and the likelihood term is
 * sum_s w_s VE_s instead of the Monte-Carlo mean.  w: S weights summing to 1 (host float64), applied by the likelihood
 * kernels to sample s of every following elbo / elbo_grad / train_step / natgrad_step call with that S; NULL: back to 1/S. */
DSDGP_API int dsdgp_set_sample_weights(dsdgp_ctx* ctx, const double* w, int S);

/* param.set_trainable(flag) (demos/using_natural_gradients.ipynb: the NatGrad-managed q_mu, q_sqrt are taken away
 * from Adam): an untrainable field is skipped by dsdgp_train_step's Adam update; its gradient is still computed. */
DSDGP_API int dsdgp_set_trainable(dsdgp_ctx* ctx, int layer, int field, int trainable);

/* gpflow.training.NatGradOptimizer(gamma).minimize(model, var_list=[[q_mu, q_sqrt] of each listed layer], maxiter=1)
 * (tests/test_collapsed.py:99-100, demos/demo_regression_UCI.ipynb:357-366): one ELBO+gradient pass on the minibatch,
 * then theta <- theta - gamma * d(-ELBO)/d eta in natural parameters (default XiNat), fp64, for every output dimension
 * of the n_layers layers listed in layers[].  elbo = the value before the update.  0 < gamma <= 1.
 * DSDGP_ERR_NOT_PD: the updated precision was not positive definite (that layer's q is left unchanged). */
DSDGP_API int dsdgp_natgrad_step(dsdgp_ctx* ctx, const float* X, const float* Y, int N, int S, double num_data,
                       const float* const* zs, uint64_t seed, unsigned flags, const int* layers, int n_layers,
                       double gamma, double* elbo);

/* NCCL communicator for row-sharded data parallelism (no counterpart in the reference).
 * id = 128-byte ncclUniqueId produced by dsdgp_comm_unique_id on rank 0 and broadcast by the host. */
DSDGP_API int dsdgp_comm_unique_id(void* id128);
DSDGP_API int dsdgp_comm_init(dsdgp_ctx* ctx, const void* id128, int rank, int world);

/* SVGP_Layer.KL() for every layer (layers.py:221-246): kl[L], evaluated from the current parameters. */
DSDGP_API int dsdgp_kl(dsdgp_ctx* ctx, double* kl);

/* Enqueue all following work of this ctx on the caller's CUDA stream (a cudaStream_t passed as void*; NULL: the ctx's own
 * stream).  The caller keeps ownership of the stream.  Replaces nothing upstream (the reference runs inside a TF session);
 * SURVEY 8(b) asks for it so that a PyTorch caller can order the step against its own work. */
DSDGP_API int dsdgp_set_stream(dsdgp_ctx* ctx, void* stream);
/* Device pointers of the flat fp32 parameter and gradient buffers (n elements each; the gradient buffer carries two more
 * floats, the ELBO as hi/lo) and the element offset of a field inside them (-1: no such field), for callers that keep
 * parameters on the device (SURVEY 8(b) "caller owns ... device pointers").  The layout is documented in DESIGN.md section 3. */
DSDGP_API int dsdgp_device_buffers(dsdgp_ctx* ctx, float** params, float** grads, size_t* n);
DSDGP_API long long dsdgp_param_offset(dsdgp_ctx* ctx, int layer, int field);

DSDGP_API int dsdgp_sync(dsdgp_ctx* ctx);
/* Number of kernels of this library launched (or replayed inside CUDA graphs) by this ctx so far. */
DSDGP_API long long dsdgp_launch_count(dsdgp_ctx* ctx);
/* Time the last graph-replayed step spent on the device, from CUDA events on the ctx stream (ms). */
DSDGP_API int dsdgp_last_step_ms(dsdgp_ctx* ctx, float* ms);
/* CUDA-event stopwatch on the ctx stream (the stream every kernel of this ctx is launched on). */
DSDGP_API int dsdgp_timer_start(dsdgp_ctx* ctx);
DSDGP_API int dsdgp_timer_stop(dsdgp_ctx* ctx, float* ms);
/* Per-stage device time of the last step run with option "profile"=1 (eager launches bracketed by events).
 * Fills up to n entries: [0] prep, [1] lik, [2] fin, [3] comm, [4] adam, then per layer l:
 * [5+3l] forward, [6+3l] backward rows, [7+3l] row reductions.  Returns the number of entries. */
DSDGP_API int dsdgp_profile(dsdgp_ctx* ctx, float* ms, int n);
/* Knobs: "graph" (0/1), "profile" (0/1), "path" (0 fp32 SIMT / 1 tcgen05), "overlap" (0/1);
 * data-parallel layout: "n_global", "n_offset" (row sharding: this rank holds rows [n_offset, n_offset+N) of n_global),
 * "s_world", "s_offset" (sample sharding: this rank draws samples [s_offset, s_offset+S) of s_world*S). */
DSDGP_API int dsdgp_set_option(dsdgp_ctx* ctx, const char* name, double value);

#ifdef __cplusplus
}
#endif
#endif /* DSDGP_H */
