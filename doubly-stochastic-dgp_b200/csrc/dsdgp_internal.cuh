// Internal declarations shared by the kernels of libdsdgp.so (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include "../../include/dsdgp.h"

#define DSDGP_NT 256          // threads per CTA of the row-tile kernels
#define DSDGP_KB 16           // k-step of the SIMT tile GEMM
#define DSDGP_NCH 64          // column chunk of the SIMT tile GEMM

// Per-layer view handed to kernels by value.
struct LayerDev {
    int M, Din, Dout, kern, ard, white, mean, n_ls, idx;
    int kwhite;     // 1: the layer kernel is Sum(kernel, White) -- wvar is a trainable parameter
    int ipd;        // input_prop_dim (layers.py:105-117): this layer's F rows are [X[:ipd] | samples], row stride ipd + Dout
    // parameters (fp32, device).  wvar: variance of the White term of Sum(kernel, White) (0 and fixed when there is none):
    // Kuu += wvar I, Kdiag += wvar, K(Z, X) unchanged (gpflow White.K(X, X2) = 0)
    const float *Z, *q_mu, *q_sqrt, *ls, *var, *wvar, *meanW, *meanB;
    // gradient destinations (fp32, same layout as the parameters)
    float *gZ, *gq_mu, *gq_sqrt, *gls, *gvar, *gwvar;
    // per-step small-matrix results
    double *K64, *Lu64, *Linv64, *Kinv64, *Ssum64, *T1, *KbarKL, *Gsym;   // M x M each
    float *Linv32, *LinvT32, *q_sqrtT;                                    // M x M, M x M, D x M x M
    float *wpack_fwd;                                                     // tcgen05 path: packed tf32 weight tiles (or NULL)
    double *scal;   // [0] sum log diag Lu  [1] sum log diag(q_sqrt)^2  [2] tr(Kinv Ssum)  [3] KL  [4] sum q_sqrt^2+q_mu^2 (white)
    // row-reduced accumulators (zeroed every step)
    float *Pd;      // D x M x M : sum_r vbar_rd u_r u_r^T
    float *G;       // M x M     : sum_r w_r u_r^T
    float *qmubar;  // M x D     : sum_r u_r mubar_r^T
};

struct LayerSet {
    LayerDev l[DSDGP_MAX_LAYERS];
    int L;
    // host-side kernel selection of the fp64 prep / gradient-assembly stages (dsdgp_set_option "prep_algo", "prep_threads",
    // "fin_algo"; per context): prep_algo 2 = k_prepA_c4, 1 = k_prepA_ldl, 0 = k_prepA; fin_algo 1 = tiled kernels, 0 = per element
    int prep_algo, prep_threads, fin_algo;
};

// Scalars that change every step live in device memory so that a captured CUDA graph can be replayed.
struct StepArgs {
    unsigned long long seed;
    double lik_scale;      // num_data / (N_global * S_eff)
    double kl_weight;      // 1/world
    int N_global, n_offset;
    int s_offset;          // first global sample index of this rank (S-sharded data parallelism), else 0
    unsigned epoch;        // step counter (> 0): value of the per-tile "published" flags of the persistent chain kernels
    // Adam
    double lr_t, beta1, beta2, eps;
};

struct Accum {             // fp64 scalar accumulators (zeroed every step)
    double lik;            // sum of scaled variational expectations (this rank)
    double kl;             // sum_l KL_l
    double glikvar;        // d/d lik_var
    double elbo;           // lik - kl_weight*kl   (after all-reduce: the ELBO)
    int status;            // nonzero: Cholesky failed (layer index + 1)
    int pad;
    // per layer: 1 when some |q_sqrt| entry exceeds 1e-3 sqrt(kernel variance) -- then |L_d^T u|^2 is not negligible in the
    // conditional variance and the forward kernel runs that product as 3xTF32 instead of 1xTF32 (set by k_pack_fwd)
    int g2flag[DSDGP_MAX_LAYERS];
};

// ----------------------------------------------------------------------------------------------
// Counter-based RNG: Philox4x32-10 keyed by seed; counter = (n_global, s, layer, dchunk) -> 4 normals.
// z(s, n, d) is a pure function of (seed, layer, s, n_global, d): results do not depend on how rows
// are sharded over CTAs or GPUs.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// (kept out of line: it is called from several places in the big row kernels and I-cache footprint matters there)
static __device__ __noinline__ float dsdgp_normal(unsigned long long seed, int layer, int s, int n_global, int d) {
    uint32_t r[4];
    philox4x32_10((uint32_t)n_global, (uint32_t)s, (uint32_t)layer, (uint32_t)(d >> 2),
                  (uint32_t)seed, (uint32_t)(seed >> 32), r);
    // Box-Muller on the pair holding lane d&3
    int p = (d & 2);
    float u1 = ((float)r[p] + 0.5f) * 2.3283064365386963e-10f;       // (0,1)
    float u2 = ((float)r[p + 1] + 0.5f) * 2.3283064365386963e-10f;
    float rad = sqrtf(-2.0f * logf(u1));
    float sn, cs;
    sincospif(2.0f * u2, &sn, &cs);
    return (d & 1) ? rad * sn : rad * cs;
}

// both draws of the Box-Muller pair holding d_even (even) and d_even + 1: bit-identical to dsdgp_normal(.., d_even) and
// dsdgp_normal(.., d_even + 1) at a quarter of the cost per draw (one Philox block, one log/sincos)
__device__ __forceinline__ void dsdgp_normal2_body(unsigned long long seed, int layer, int s, int n_global, int d_even,
                                                   float& z0, float& z1) {
    uint32_t r[4];
    philox4x32_10((uint32_t)n_global, (uint32_t)s, (uint32_t)layer, (uint32_t)(d_even >> 2),
                  (uint32_t)seed, (uint32_t)(seed >> 32), r);
    const bool hi = (d_even & 2) != 0;
    float u1 = ((float)(hi ? r[2] : r[0]) + 0.5f) * 2.3283064365386963e-10f;
    float u2 = ((float)(hi ? r[3] : r[1]) + 0.5f) * 2.3283064365386963e-10f;
    float rad = sqrtf(-2.0f * logf(u1));
    float sn, cs;
    sincospif(2.0f * u2, &sn, &cs);
    z0 = rad * cs;
    z1 = rad * sn;
}
static __device__ __noinline__ void dsdgp_normal2(unsigned long long seed, int layer, int s, int n_global, int d_even,
                                                  float& z0, float& z1) {
    dsdgp_normal2_body(seed, layer, s, n_global, d_even, z0, z1);
}
// Layer-1 fold of the forward pass: the S draws of one (row, output pair) -- F[(s N + row) D + d] = mean + sd z(s, n, d).
// Four samples per iteration with the generator inlined.  Same counters, same arithmetic: bit-identical to dsdgp_normal2 per
// sample.  Measured (clock64 stamps, north-star layer 1, S = 20): 54k -> 46k of the tile's 92k -> 85k cycles; the loop is
// issue-bound (~200 instructions per draw pair: 10 Philox rounds, full-precision log and sincospi), not latency-bound.
static __device__ __noinline__ void dsdgp_draw_fold(unsigned long long seed, int layer, int s_off, int n_global, int d0, int nd,
                                                    int nrep, size_t stride, float* __restrict__ F, float* __restrict__ z_out,
                                                    float m0, float m1, float sd0, float sd1) {
    // F / z_out point at element (sample 0, row, d0); consecutive samples are `stride` elements apart
    for (int ss = 0; ss < nrep; ss += 4) {
        float z[4][2];
#pragma unroll
        for (int e = 0; e < 4; ++e) dsdgp_normal2_body(seed, layer, min(ss + e, nrep - 1) + s_off, n_global, d0, z[e][0], z[e][1]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (ss + e < nrep) {
                const size_t o = (size_t)(ss + e) * stride;
                if (z_out) { z_out[o] = z[e][0]; if (nd == 2) z_out[o + 1] = z[e][1]; }
                F[o] = fmaf(z[e][0], sd0, m0);
                if (nd == 2) F[o + 1] = fmaf(z[e][1], sd1, m1);
            }
        }
    }
}
// 2^x, single MUFU (ex2.approx): relative error 2^-22.5
__device__ __forceinline__ float fast_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// kernel value and d k / d r2   (r2 already scaled by lengthscales)
__device__ __forceinline__ void kern_eval_f(int kern, float r2, float var, float& k, float& kp) {
    if (kern == DSDGP_KERN_RBF) {
        k = var * expf(-0.5f * r2);
        kp = -0.5f * k;
    } else {
        float r = sqrtf(r2 + 1e-12f);
        const float s5 = 2.2360679774997896f;
        float e = expf(-s5 * r);
        k = var * (1.0f + s5 * r + (5.0f / 3.0f) * r * r) * e;
        kp = var * e * (-5.0f / 6.0f) * (1.0f + s5 * r);
    }
}
// fast-path variant for the tensor-core kernels: ex2.approx-based exponential (rel. error ~|x| 2^-23)
__device__ __forceinline__ void kern_eval_fast(int kern, float r2, float var, float& k, float& kp) {
    if (kern == DSDGP_KERN_RBF) {
        k = var * __expf(-0.5f * r2);
        kp = -0.5f * k;
    } else {
        float r = sqrtf(r2 + 1e-12f);
        const float s5 = 2.2360679774997896f;
        float e = __expf(-s5 * r);
        k = var * (1.0f + s5 * r + (5.0f / 3.0f) * r * r) * e;
        kp = var * e * (-5.0f / 6.0f) * (1.0f + s5 * r);
    }
}
__device__ __forceinline__ void kern_eval_d(int kern, double r2, double var, double& k, double& kp) {
    if (kern == DSDGP_KERN_RBF) {
        k = var * exp(-0.5 * r2);
        kp = -0.5 * k;
    } else {
        double r = sqrt(r2 + 1e-12);
        const double s5 = 2.2360679774997896;
        double e = exp(-s5 * r);
        k = var * (1.0 + s5 * r + (5.0 / 3.0) * r * r) * e;
        kp = var * e * (-5.0 / 6.0) * (1.0 + s5 * r);
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ----------------------------------------------------------------------------------------------
// launch interfaces (implemented in the .cu files)
// ----------------------------------------------------------------------------------------------
struct FwdArgs {
    const float* Xin;     // (R, Din)
    int R;                // rows of this layer on this rank
    int N;                // minibatch rows on this rank (row r = s*N + n)
    int S_rep;            // >1: layer evaluated on N rows, sample S_rep times (layer-1 dedup)
    float* U;             // (R, M) out
    float* Fmean;         // (R, Dout) out
    float* Fvar;          // (R, Dout) out
    float* F;             // (S_rep*R, Dout) out or NULL
    const float* z;       // (S_rep*R, Dout) or NULL -> Philox
    float* z_out;         // when z == NULL: the Philox draws are stored here for the backward pass (or NULL)
    float jitter;
    const StepArgs* sa;
    long long* dbg;       // optional clock64() stamps of CTA 0 (diagnostics), else NULL
    const Accum* acc;     // g2flag (see Accum)
    int g2_passes;        // 0: per layer from acc->g2flag; 1 / 3: forced (option "g2_passes")
    // Deferred layer-1 fold (chained forward only): when fold_mean != NULL this layer's input rows are not read from Xin but
    // DRAWN here from the previous (de-duplicated, N-row) layer's marginals -- x(s, n, :) = fold_mean[n] + sqrt(fold_var[n] +
    // jitter) z(s, n, :) -- and written to fold_F / fold_zout for the backward pass.  The previous layer's tiles then publish
    // right after their mean / variance instead of drawing S samples per row first (46k of the 85k cycles of a north-star
    // layer-1 tile, with every other CTA of the chain waiting for it).  Same counters and arithmetic as the in-tile fold.
    const float *fold_mean, *fold_var;   // (N, Din)
    const float* fold_z;                 // injected draws of the previous layer (S*N, Din) or NULL -> Philox
    float *fold_F, *fold_zout;           // (S*N, Din) out; fold_zout may be NULL
    int fold_layer;                      // Philox layer index of the previous layer
};

// all layers' forward tiles as one persistent launch (layer_tc.cu k_chain_fwd_tc)
struct FwdChain {
    FwdArgs a[DSDGP_MAX_LAYERS];
    int L, max_tiles;
    int tiles[DSDGP_MAX_LAYERS];
    int base[DSDGP_MAX_LAYERS + 1];      // first task number of each layer
    unsigned* flags;                     // [L][max_tiles]
    const StepArgs* sa;
};

struct BwdArgs {
    const float* Xin;     // (R, Din)
    int R, N, S_rep;
    const float* U;       // (R, M)
    const float* Fvar;    // (R, Dout)
    const float* fbar;    // (S_rep*R, Dout) upstream dE/dF, or NULL when mubar/vbar are given
    const float* z;       // as in forward
    float* mubar;         // (R, Dout)  in (last layer) / out
    float* vbar;          // (R, Dout)
    float* W;             // (R, M) out
    float* xbar;          // (R, Din) out or NULL (first layer)
    float jitter;
    const StepArgs* sa;
    long long* dbg;       // optional clock64() stamps of CTA 0 (diagnostics), else NULL
    long long* dbg_rr;    // the same for the row-reduction kernel
    // tile-level hand-over between the row kernels of consecutive layers (tcgen05 path): tile t of layer l reads only what
    // tile t of layer l+1 wrote (xbar rows of the same 128-row tile), so the kernel of layer l is launched as a programmatic
    // dependent of the kernel of layer l+1 and each tile waits for ITS producer's flag (== StepArgs::epoch) instead of the grid
    unsigned* tile_done;          // [tiles] written by this launch when a tile's outputs are published, or NULL
    const unsigned* tile_wait;    // [tiles] flags of the upstream launch to wait for before reading fbar, or NULL
    int wait_before_loads;        // 1: nothing of the tile may be read before the flag (last layer behind the forward tail)
    int wait_count;               // > 0: wait for tile_wait[0 .. wait_count) -- the de-duplicated first layer folds the S samples
                                  // of its rows, which are spread over ALL tiles of the layer above
};

void launch_prep(const LayerSet& ls, double jitter, Accum* acc, const StepArgs* sa, cudaStream_t st, cudaStream_t st_kl,
                 cudaEvent_t ev_fork, long long* nlaunch);
void launch_fwd(const LayerDev& P, const FwdArgs& a, int num_sms, cudaStream_t st, long long* nlaunch);
void launch_bwd_rows(const LayerDev& P, const BwdArgs& a, int num_sms, cudaStream_t st, long long* nlaunch);
void launch_bwd_rowred(const LayerDev& P, const BwdArgs& a, int num_sms, cudaStream_t st, long long* nlaunch);
void launch_fin(const LayerSet& ls, int l0, int l1, Accum* acc, const StepArgs* sa, cudaStream_t st, long long* nlaunch, int part = 0);
void launch_lik_gaussian_tiled(const float* Fmean, const float* Fvar, const float* Y, int R, int N, int Dy, const float* lik_var,
                               float* mubar, float* vbar, Accum* acc, const StepArgs* sa, int want_grad, const float* sw,
                               const unsigned* tile_wait, unsigned* tile_done, bool programmatic, cudaStream_t st, long long* nl);
void launch_lik_gaussian(const float* Fmean, const float* Fvar, const float* Y, int R, int N, int Dy,
                         const float* lik_var, float* mubar, float* vbar, Accum* acc, const StepArgs* sa,
                         int want_grad, const float* sample_w, cudaStream_t st, long long* nlaunch);
void launch_lik_multiclass(const float* Fmean, const float* Fvar, const float* Y, int R, int N, int K,
                           float* mubar, float* vbar, Accum* acc, const StepArgs* sa, int want_grad,
                           const float* sample_w, cudaStream_t st, long long* nlaunch);
void launch_lik_bernoulli(const float* Fmean, const float* Fvar, const float* Y, int R, int N, int Dy, float* mubar,
                          float* vbar, Accum* acc, const StepArgs* sa, int want_grad, const float* sample_w,
                          cudaStream_t st, long long* nlaunch);
void launch_ve_elem(int lik, const float* Fmean, const float* Fvar, const float* Y, int R, int N, int D, const float* lik_var,
                    float* out, cudaStream_t st, long long* nlaunch);
void launch_elbo_finish(Accum* acc, const StepArgs* sa, float* glikvar, float* elbo_hi_lo, cudaStream_t st, long long* nlaunch);
void launch_result(const Accum* acc, const float* elbo_hi_lo, int use_hi_lo, double* result, cudaStream_t st, long long* nlaunch);
cudaError_t layer_kernels_init();
cudaError_t layer_tc_init();
bool tc_fwd_supported(const LayerDev& P);
size_t tc_fwd_pack_bytes(int M, int D, int white);
void launch_pack_fwd(const LayerSet& ls, int part, Accum* acc, cudaStream_t st, long long* nlaunch);
void launch_fwd_tc(const LayerDev& P, const FwdArgs& a, cudaStream_t st, long long* nlaunch);
bool tc_chain_fwd_supported(const LayerSet& ls);
bool launch_chain_fwd_tc(const LayerSet& ls, const FwdChain& fc, int num_sms, cudaStream_t st, long long* nlaunch);
cudaError_t layer_tc_bwd_init();
bool tc_bwd_supported(const LayerDev& P);
void launch_bwd_rows_tc(const LayerDev& P, const BwdArgs& a, cudaStream_t st, long long* nlaunch, bool programmatic = false);
cudaError_t rowred_tc_init();
bool tc_rowred_supported(const LayerDev& P);
void launch_bwd_rowred_tc(const LayerDev& P, const BwdArgs& a, int num_sms, cudaStream_t st, long long* nlaunch);
cudaError_t small_matrix_init();
void launch_tail(float* params, float* free_, float* m, float* v, float* grads, const unsigned char* kinds, size_t n,
                 const StepArgs* sa, Accum* acc, size_t off_likvar, int use_hi_lo, int do_adam, double* result,
                 cudaStream_t st, long long* nlaunch);
void launch_constrain_init(const float* params, float* free_, const unsigned char* kinds, size_t n, cudaStream_t st,
                           long long* nlaunch);
void launch_predict_y(int lik, const float* Fmean, const float* Fvar, int R, int D, const float* lik_var, float* mean,
                      float* var, cudaStream_t st, long long* nlaunch);
void launch_predict_density(int lik, const float* Fmean, const float* Fvar, const float* Y, int S, int N, int D, int dedup,
                            const float* lik_var, float* out, cudaStream_t st, long long* nlaunch);
cudaError_t full_cov_init();
size_t full_cov_ws_doubles(int M, int D, int Dmax_io, int N, int S);
void launch_full_cov_x64(const float* X, size_t n, double* X64, cudaStream_t st, long long* nlaunch);
void launch_full_cov_layer(const LayerDev& P, const double* Xin, size_t xs, int N, int S, double jitter, const float* z,
                           const StepArgs* sa, double* ws, double* Fnext, float* F_out, float* mean_out, float* var_out,
                           int* status, cudaStream_t st, long long* nlaunch);
cudaError_t natgrad_init();
size_t natgrad_ws_doubles(int M, int D);
void launch_natgrad_layer(const LayerDev& P, double gamma, double* ws, int* status, float* q_mu, float* q_sqrt,
                          cudaStream_t st, long long* nlaunch);
void launch_natgrad_commit(float* dst, const float* src, size_t n, const int* status, cudaStream_t st, long long* nlaunch);
size_t fwd_smem_bytes(int M, int Din, int TR);
size_t bwd_smem_bytes(int M, int Din, int TR);
