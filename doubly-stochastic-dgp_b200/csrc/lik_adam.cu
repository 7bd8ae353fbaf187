// Likelihood expectations (+ their adjoints), the prediction-side likelihood epilogues, and the optimiser step.
// Reference: utils.py:88-93 (BroadcastingLikelihood.variational_expectations -> gpflow Gaussian / MultiClass
// RobustMax), dgp.py:88-98 (mean over S, sum, scale); Adam = tf.train.AdamOptimizer on GPflow's
// unconstrained variables.
#include "dsdgp_internal.cuh"

__constant__ double c_gh_x[20] = {
    -5.3874808900112328, -4.6036824495507442, -3.9447640401156252, -3.3478545673832163, -2.7888060584281305,
    -2.2549740020892757, -1.7385377121165861, -1.2340762153953231, -0.73747372854539439, -0.24534070830090124,
    0.24534070830090124, 0.73747372854539439, 1.2340762153953231, 1.7385377121165861, 2.2549740020892757,
    2.7888060584281305, 3.3478545673832163, 3.9447640401156252, 4.6036824495507442, 5.3874808900112328};
__constant__ double c_gh_w[20] = {
    2.2293936455341447e-13, 4.3993409922731747e-10, 1.0860693707692782e-07, 7.8025564785320599e-06,
    0.00022833863601635365, 0.0032437733422378567, 0.024810520887463643, 0.10901720602002329, 0.28667550536283415,
    0.46224366960061009, 0.46224366960061009, 0.28667550536283415, 0.10901720602002329, 0.024810520887463643,
    0.0032437733422378567, 0.00022833863601635365, 7.8025564785320599e-06, 1.0860693707692782e-07,
    4.3993409922731747e-10, 2.2293936455341447e-13};

__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Gaussian: VE = -1/2 log 2pi - 1/2 log s2 - 1/2 ((y-mu)^2 + v)/s2 ; rows r = s*N + n share y_n (utils.py:72-73)
__global__ void k_lik_gaussian(const float* __restrict__ Fmean, const float* __restrict__ Fvar, const float* __restrict__ Y,
                               int R, int N, int Dy, const float* __restrict__ lik_var, float* __restrict__ mubar,
                               float* __restrict__ vbar, Accum* acc, const StepArgs* sa, int want_grad,
                               const float* __restrict__ sw) {
    // sw (optional): per-sample weights relative to the uniform 1/S (DGP_Quad's Gauss-Hermite weights, dgp.py:159-166)
    const double s2 = (double)lik_var[0], c0 = sa->lik_scale;
    const double base = -0.5 * 1.8378770664093453 - 0.5 * log(s2);
    double ve = 0.0, gl = 0.0;
    size_t total = (size_t)R * Dy;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        int r = idx / Dy, d = idx % Dy, n = r % N;
        const double c = sw ? c0 * (double)sw[r / N] : c0;
        double y = Y[(size_t)n * Dy + d], mu = Fmean[idx], v = Fvar[idx];
        double e2 = (y - mu) * (y - mu) + v;
        ve += c * (base - 0.5 * e2 / s2);
        if (want_grad) {
            mubar[idx] = (float)(c * (y - mu) / s2);
            vbar[idx] = (float)(-0.5 * c / s2);
            gl += c * (-0.5 / s2 + 0.5 * e2 / (s2 * s2));
        }
    }
    ve = warp_sum_d(ve); gl = warp_sum_d(gl);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&acc->lik, ve);
        if (want_grad) atomicAdd(&acc->glikvar, gl);
    }
}

// The same per 128-row tile of the last layer (the tiling of the tcgen05 kernels): block b waits for forward tile b's flag and
// publishes its own, so the likelihood and the last layer's backward rows start under the forward chain's tail wave.
__global__ void __launch_bounds__(128) k_lik_gaussian_tiled(const float* __restrict__ Fmean, const float* __restrict__ Fvar,
                                                            const float* __restrict__ Y, int R, int N, int Dy,
                                                            const float* __restrict__ lik_var, float* __restrict__ mubar,
                                                            float* __restrict__ vbar, Accum* acc, const StepArgs* sa, int want_grad,
                                                            const float* __restrict__ sw, const unsigned* tile_wait,
                                                            unsigned* tile_done) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const unsigned epoch = sa->epoch;
    if (tile_wait) {
        if (threadIdx.x == 0)
            while (ld_acquire_gpu_u32(tile_wait + blockIdx.x) != epoch) __nanosleep(1000);
        __syncthreads();
    }
    const double s2 = (double)lik_var[0], c0 = sa->lik_scale;
    const double base = -0.5 * 1.8378770664093453 - 0.5 * log(s2);
    double ve = 0.0, gl = 0.0;
    const size_t i0 = (size_t)blockIdx.x * 128 * Dy, i1 = min((size_t)R * Dy, i0 + (size_t)128 * Dy);
    for (size_t idx = i0 + threadIdx.x; idx < i1; idx += 128) {
        int r = idx / Dy, d = idx % Dy, n = r % N;
        const double c = sw ? c0 * (double)sw[r / N] : c0;
        double y = Y[(size_t)n * Dy + d], mu = Fmean[idx], v = Fvar[idx];
        double e2 = (y - mu) * (y - mu) + v;
        ve += c * (base - 0.5 * e2 / s2);
        if (want_grad) {
            mubar[idx] = (float)(c * (y - mu) / s2);
            vbar[idx] = (float)(-0.5 * c / s2);
            gl += c * (-0.5 / s2 + 0.5 * e2 / (s2 * s2));
        }
    }
    ve = warp_sum_d(ve); gl = warp_sum_d(gl);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&acc->lik, ve);
        if (want_grad) atomicAdd(&acc->glikvar, gl);
    }
    __syncthreads();
    if (tile_done && threadIdx.x == 0) {
        __threadfence();
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(tile_done + blockIdx.x), "r"(epoch) : "memory");
    }
}

void launch_lik_gaussian_tiled(const float* Fmean, const float* Fvar, const float* Y, int R, int N, int Dy, const float* lik_var,
                               float* mubar, float* vbar, Accum* acc, const StepArgs* sa, int want_grad, const float* sw,
                               const unsigned* tile_wait, unsigned* tile_done, bool programmatic, cudaStream_t st, long long* nl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((R + 127) / 128); cfg.blockDim = dim3(128); cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = programmatic ? 1 : 0;
    cudaLaunchKernelEx(&cfg, k_lik_gaussian_tiled, Fmean, Fvar, Y, R, N, Dy, lik_var, mubar, vbar, acc, sa, want_grad, sw, tile_wait,
                       tile_done);
    *nl += 1;
}

void launch_lik_gaussian(const float* Fmean, const float* Fvar, const float* Y, int R, int N, int Dy, const float* lik_var,
                         float* mubar, float* vbar, Accum* acc, const StepArgs* sa, int want_grad, const float* sw,
                         cudaStream_t st, long long* nl) {
    size_t total = (size_t)R * Dy;
    int nb = (int)min((size_t)1024, (total + 255) / 256);
    k_lik_gaussian<<<nb, 256, 0, st>>>(Fmean, Fvar, Y, R, N, Dy, lik_var, mubar, vbar, acc, sa, want_grad, sw);
    *nl += 1;
}

// Bernoulli, probit link (gpflow.likelihoods.Bernoulli default; tests/test_dgp.py:48-54): VE = 20-point Gauss-Hermite of
// log p(y | f), p(y=1 | f) = Phi(f) (1 - 2e-3) + 1e-3 (gpflow probit), f = mu + sqrt(2 v) x_h.  One thread per element.
__device__ __forceinline__ double bern_ve(double mu, double v, bool y1, double* gm, double* gv) {
    const bool clipped = v < 1e-10;
    if (clipped) v = 1e-10;
    const double sd2 = sqrt(2.0 * v);
    double ve = 0.0, dm = 0.0, dv = 0.0;
    for (int h = 0; h < 20; ++h) {
        const double x = mu + sd2 * c_gh_x[h];
        const double w = c_gh_w[h] * 0.56418958354775628;       // / sqrt(pi)
        const double p = 0.5 * (1.0 + erf(x * 0.70710678118654752)) * (1.0 - 2e-3) + 1e-3;
        const double dp = 0.3989422804014327 * exp(-0.5 * x * x) * (1.0 - 2e-3);
        ve += w * log(y1 ? p : 1.0 - p);
        const double dl = y1 ? dp / p : -dp / (1.0 - p);         // d log p(y|f) / df
        dm += w * dl;
        dv += w * dl * c_gh_x[h] / sd2;                          // df/dv = x_h / sqrt(2 v)
    }
    if (gm) { *gm = dm; *gv = clipped ? 0.0 : dv; }
    return ve;
}
__device__ __forceinline__ double bern_predict_p(double mu, double v) {
    return 0.5 * (1.0 + erf(mu / sqrt(1.0 + v) * 0.70710678118654752)) * (1.0 - 2e-3) + 1e-3;
}
__global__ void k_lik_bernoulli(const float* __restrict__ Fmean, const float* __restrict__ Fvar, const float* __restrict__ Y,
                                int R, int N, int Dy, float* __restrict__ mubar, float* __restrict__ vbar, Accum* acc,
                                const StepArgs* sa, int want_grad, const float* __restrict__ sw) {
    const double c0 = sa->lik_scale;
    double ve = 0.0;
    const size_t total = (size_t)R * Dy;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int r = idx / Dy, d = idx % Dy, n = r % N;
        const double c = sw ? c0 * (double)sw[r / N] : c0;
        double gm, gv;
        ve += c * bern_ve(Fmean[idx], Fvar[idx], Y[(size_t)n * Dy + d] > 0.5f, want_grad ? &gm : nullptr, &gv);
        if (want_grad) { mubar[idx] = (float)(c * gm); vbar[idx] = (float)(c * gv); }
    }
    ve = warp_sum_d(ve);
    if ((threadIdx.x & 31) == 0) atomicAdd(&acc->lik, ve);
}
void launch_lik_bernoulli(const float* Fmean, const float* Fvar, const float* Y, int R, int N, int Dy, float* mubar,
                          float* vbar, Accum* acc, const StepArgs* sa, int want_grad, const float* sw, cudaStream_t st,
                          long long* nl) {
    const size_t total = (size_t)R * Dy;
    k_lik_bernoulli<<<(int)min((size_t)1024, (total + 127) / 128), 128, 0, st>>>(Fmean, Fvar, Y, R, N, Dy, mubar, vbar, acc, sa,
                                                                                  want_grad, sw);
    *nl += 1;
}

// MultiClass / RobustMax(eps = 1e-3): one thread per row; 20-point Gauss-Hermite over the labelled latent.
#define MC_MAXK 32
__global__ void k_lik_multiclass(const float* __restrict__ Fmean, const float* __restrict__ Fvar, const float* __restrict__ Y,
                                 int R, int N, int K, float* __restrict__ mubar, float* __restrict__ vbar, Accum* acc,
                                 const StepArgs* sa, int want_grad, const float* __restrict__ sw) {
    const double eps = 1e-3;
    double c = sa->lik_scale;
    const double l1 = log(1.0 - eps), l0 = log(eps / (K - 1.0));
    double ve = 0.0;
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < R) {
        int n = r % N;
        if (sw) c *= (double)sw[r / N];
        int y = (int)(Y[n] + 0.5f);
        double mu[MC_MAXK], sd[MC_MAXK], gm[MC_MAXK], gv[MC_MAXK];
        bool clipped[MC_MAXK];
        for (int k = 0; k < K; ++k) {
            double v = Fvar[(size_t)r * K + k];
            clipped[k] = v < 1e-10;
            if (clipped[k]) v = 1e-10;
            mu[k] = Fmean[(size_t)r * K + k]; sd[k] = sqrt(v); gm[k] = 0.0; gv[k] = 0.0;
        }
        double p = 0.0;
        const double s2y = sqrt(2.0) * sd[y];
        for (int h = 0; h < 20; ++h) {
            double x = mu[y] + s2y * c_gh_x[h];
            double w = c_gh_w[h] * 0.56418958354775628;       // / sqrt(pi)
            double prod = 1.0;
            double cdf[MC_MAXK], pdf[MC_MAXK];
            for (int k = 0; k < K; ++k) {
                if (k == y) continue;
                double t = (x - mu[k]) / sd[k];
                cdf[k] = 0.5 * (1.0 + erf(t * 0.70710678118654752)) * (1.0 - 2e-4) + 1e-4;
                pdf[k] = 0.3989422804014327 * exp(-0.5 * t * t) * (1.0 - 2e-4) / sd[k];      // d cdf / dx
                prod *= cdf[k];
            }
            p += w * prod;
            if (want_grad) {
                double dPdx = 0.0;
                for (int k = 0; k < K; ++k) {
                    if (k == y) continue;
                    double others = prod / cdf[k];
                    double dk = w * others * pdf[k];
                    dPdx += dk;
                    gm[k] -= dk;                                           // d/dmu_k = -d/dx
                    if (!clipped[k]) gv[k] -= dk * (x - mu[k]) / (2.0 * sd[k] * sd[k]);
                }
                gm[y] += dPdx;
                if (!clipped[y]) gv[y] += dPdx * c_gh_x[h] / (s2y);       // dx/dv_y = gh/(sqrt(2) sd_y)
            }
        }
        ve = c * (p * l1 + (1.0 - p) * l0);
        if (want_grad) {
            double f = c * (l1 - l0);
            for (int k = 0; k < K; ++k) {
                mubar[(size_t)r * K + k] = (float)(f * gm[k]);
                vbar[(size_t)r * K + k] = (float)(f * gv[k]);
            }
        }
    }
    ve = warp_sum_d(ve);
    if ((threadIdx.x & 31) == 0) atomicAdd(&acc->lik, ve);
}

void launch_lik_multiclass(const float* Fmean, const float* Fvar, const float* Y, int R, int N, int K, float* mubar,
                           float* vbar, Accum* acc, const StepArgs* sa, int want_grad, const float* sw, cudaStream_t st,
                           long long* nl) {
    k_lik_multiclass<<<(R + 127) / 128, 128, 0, st>>>(Fmean, Fvar, Y, R, N, K, mubar, vbar, acc, sa, want_grad, sw);
    *nl += 1;
}

// ----------------------------------------------------------------------------------------------
// Prediction epilogues (dgp.py:116-126): likelihood.predict_mean_and_var / predict_density broadcast over the S samples
// (utils.py:110-121), and the log-mean-exp over S of predict_density (dgp.py:124-126).  Rows r = s*N + n; when `dedup`
// (single-layer model: the conditional was evaluated on the N distinct rows only) every s reads row n.
// GPflow: Gaussian.predict_mean_and_var = (Fmu, Fvar + s2), predict_density = log N(Y; Fmu, Fvar + s2);
// MultiClass(RobustMax): mean_k = P(k is largest), var = mean - mean^2, density = log(p (1-eps) + (1-p) eps/(K-1)).
// ----------------------------------------------------------------------------------------------
__device__ double mc_prob_is_largest(const float* __restrict__ mu_r, const float* __restrict__ var_r, int K, int y) {
    double mu[MC_MAXK], sd[MC_MAXK];
    for (int k = 0; k < K; ++k) {
        double v = var_r[k];
        if (v < 1e-10) v = 1e-10;
        mu[k] = mu_r[k]; sd[k] = sqrt(v);
    }
    const double s2y = sqrt(2.0) * sd[y];
    double p = 0.0;
    for (int h = 0; h < 20; ++h) {
        const double x = mu[y] + s2y * c_gh_x[h];
        double prod = 1.0;
        for (int k = 0; k < K; ++k) {
            if (k == y) continue;
            const double t = (x - mu[k]) / sd[k];
            prod *= 0.5 * (1.0 + erf(t * 0.70710678118654752)) * (1.0 - 2e-4) + 1e-4;
        }
        p += c_gh_w[h] * 0.56418958354775628 * prod;
    }
    return p;
}

__global__ void k_predict_y_gaussian(const float* __restrict__ Fmean, const float* __restrict__ Fvar, size_t total,
                                     const float* __restrict__ lik_var, float* __restrict__ mean, float* __restrict__ var) {
    const float s2 = lik_var[0];
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        mean[i] = Fmean[i];
        var[i] = Fvar[i] + s2;
    }
}
// one thread per (row, class)
__global__ void k_predict_y_multiclass(const float* __restrict__ Fmean, const float* __restrict__ Fvar, int R, int K,
                                       float* __restrict__ mean, float* __restrict__ var) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)R * K) return;
    const int r = (int)(i / K), k = (int)(i % K);
    const double p = mc_prob_is_largest(Fmean + (size_t)r * K, Fvar + (size_t)r * K, K, k);
    mean[i] = (float)p;
    var[i] = (float)(p - p * p);
}
// out (N, Dy): logsumexp_s(log N(y; mu_s, v_s + s2) - log S), streaming over s
__global__ void k_density_gaussian(const float* __restrict__ Fmean, const float* __restrict__ Fvar, const float* __restrict__ Y,
                                   int S, int N, int Dy, int dedup, const float* __restrict__ lik_var, float* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)N * Dy) return;
    const double s2 = (double)lik_var[0], y = Y[i], lS = log((double)S);
    double mx = -INFINITY, acc = 0.0;
    for (int s = 0; s < S; ++s) {
        const size_t j = dedup ? i : (size_t)s * N * Dy + i;
        const double v = (double)Fvar[j] + s2, e = y - (double)Fmean[j];
        const double l = -0.5 * 1.8378770664093453 - 0.5 * log(v) - 0.5 * e * e / v - lS;
        if (l > mx) { acc = acc * exp(mx - l) + 1.0; mx = l; }
        else acc += exp(l - mx);
    }
    out[i] = (float)(mx + log(acc));
}
// out (N, 1)
__global__ void k_density_multiclass(const float* __restrict__ Fmean, const float* __restrict__ Fvar, const float* __restrict__ Y,
                                     int S, int N, int K, int dedup, float* __restrict__ out) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int y = (int)(Y[n] + 0.5f);
    const double eps = 1e-3, lS = log((double)S);
    double mx = -INFINITY, acc = 0.0;
    for (int s = 0; s < S; ++s) {
        const size_t r = dedup ? (size_t)n : (size_t)s * N + n;
        const double p = mc_prob_is_largest(Fmean + r * K, Fvar + r * K, K, y);
        const double l = log(p * (1.0 - eps) + (1.0 - p) * (eps / (K - 1.0))) - lS;
        if (l > mx) { acc = acc * exp(mx - l) + 1.0; mx = l; }
        else acc += exp(l - mx);
    }
    out[n] = (float)(mx + log(acc));
}

// Bernoulli: p = Phi(mu / sqrt(1 + v)) (squashed), var = p - p^2 ; density = log(y ? p : 1 - p), log-mean-exp over S
__global__ void k_predict_y_bernoulli(const float* __restrict__ Fmean, const float* __restrict__ Fvar, size_t total,
                                      float* __restrict__ mean, float* __restrict__ var) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const double p = bern_predict_p(Fmean[i], Fvar[i]);
        mean[i] = (float)p;
        var[i] = (float)(p - p * p);
    }
}
__global__ void k_density_bernoulli(const float* __restrict__ Fmean, const float* __restrict__ Fvar, const float* __restrict__ Y,
                                    int S, int N, int Dy, int dedup, float* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)N * Dy) return;
    const bool y1 = Y[i] > 0.5f;
    const double lS = log((double)S);
    double mx = -INFINITY, acc = 0.0;
    for (int s = 0; s < S; ++s) {
        const size_t j = dedup ? i : (size_t)s * N * Dy + i;
        const double p = bern_predict_p(Fmean[j], Fvar[j]);
        const double l = log(y1 ? p : 1.0 - p) - lS;
        if (l > mx) { acc = acc * exp(mx - l) + 1.0; mx = l; }
        else acc += exp(l - mx);
    }
    out[i] = (float)(mx + log(acc));
}

void launch_predict_y(int lik, const float* Fmean, const float* Fvar, int R, int D, const float* lik_var, float* mean,
                      float* var, cudaStream_t st, long long* nl) {
    const size_t total = (size_t)R * D;
    if (lik == DSDGP_LIK_GAUSSIAN) {
        int nb = (int)min((size_t)1184, (total + 255) / 256);
        k_predict_y_gaussian<<<nb, 256, 0, st>>>(Fmean, Fvar, total, lik_var, mean, var);
    } else if (lik == DSDGP_LIK_BERNOULLI) {
        k_predict_y_bernoulli<<<(int)min((size_t)1184, (total + 255) / 256), 256, 0, st>>>(Fmean, Fvar, total, mean, var);
    } else {
        k_predict_y_multiclass<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(Fmean, Fvar, R, D, mean, var);
    }
    *nl += 1;
}
void launch_predict_density(int lik, const float* Fmean, const float* Fvar, const float* Y, int S, int N, int D, int dedup,
                            const float* lik_var, float* out, cudaStream_t st, long long* nl) {
    if (lik == DSDGP_LIK_GAUSSIAN)
        k_density_gaussian<<<(unsigned)(((size_t)N * D + 127) / 128), 128, 0, st>>>(Fmean, Fvar, Y, S, N, D, dedup, lik_var, out);
    else if (lik == DSDGP_LIK_BERNOULLI)
        k_density_bernoulli<<<(unsigned)(((size_t)N * D + 127) / 128), 128, 0, st>>>(Fmean, Fvar, Y, S, N, D, dedup, out);
    else
        k_density_multiclass<<<(N + 127) / 128, 128, 0, st>>>(Fmean, Fvar, Y, S, N, D, dedup, out);
    *nl += 1;
}

// BroadcastingLikelihood.variational_expectations (utils.py:88-93) on caller-supplied marginals: out (S*N, Do), one value per
// element (Gaussian, Bernoulli) or per row (MultiClass), unscaled -- the host-callable form of what k_lik_* sum up.
__global__ void k_ve_elem(int lik, const float* __restrict__ Fmean, const float* __restrict__ Fvar, const float* __restrict__ Y,
                          int R, int N, int D, const float* __restrict__ lik_var, float* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (lik == DSDGP_LIK_MULTICLASS) {
        if (i >= (size_t)R) return;
        const int y = (int)(Y[i % N] + 0.5f);
        const double p = mc_prob_is_largest(Fmean + i * D, Fvar + i * D, D, y);
        out[i] = (float)(p * log(1.0 - 1e-3) + (1.0 - p) * log(1e-3 / (D - 1.0)));
        return;
    }
    if (i >= (size_t)R * D) return;
    const int r = (int)(i / D), d = (int)(i % D), n = r % N;
    const double y = Y[(size_t)n * D + d], mu = Fmean[i], v = Fvar[i];
    if (lik == DSDGP_LIK_GAUSSIAN) {
        const double s2 = (double)lik_var[0];
        out[i] = (float)(-0.5 * 1.8378770664093453 - 0.5 * log(s2) - 0.5 * ((y - mu) * (y - mu) + v) / s2);
    } else out[i] = (float)bern_ve(mu, v, y > 0.5, nullptr, nullptr);
}
void launch_ve_elem(int lik, const float* Fmean, const float* Fvar, const float* Y, int R, int N, int D, const float* lik_var,
                    float* out, cudaStream_t st, long long* nl) {
    const size_t total = lik == DSDGP_LIK_MULTICLASS ? (size_t)R : (size_t)R * D;
    k_ve_elem<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(lik, Fmean, Fvar, Y, R, N, D, lik_var, out);
    *nl += 1;
}

// ----------------------------------------------------------------------------------------------
// Adam on the unconstrained variables.  kinds: 0 identity, 1 positive (softplus + 1e-6), 2 lower-tri entry
// (trainable, identity transform), 3 structurally-zero upper-tri entry, 4 fixed.
// tf.train.AdamOptimizer: lr_t = lr sqrt(1-b2^t)/(1-b1^t) (host, StepArgs) ; theta -= lr_t m / (sqrt(v) + eps).
// ----------------------------------------------------------------------------------------------
#define POS_LOWER 1e-6f
__device__ __forceinline__ float softplus_f(float x) { return fmaxf(x, 0.f) + log1pf(__expf(-fabsf(x))); }

// One kernel for the tail of a training step: ELBO scalar -> result, likelihood-variance gradient, Adam.
//   use_hi_lo == 0 (single GPU): thread 0 finishes the ELBO from the fp64 accumulators (what k_elbo_finish + k_result did);
//   use_hi_lo == 1 (communicator): k_elbo_finish ran before the all-reduce, the summed (hi, lo) pair sits behind the gradient.
// A failed Cholesky (acc->status != 0: prepA substituted a unit pivot, the gradients are garbage) leaves parameters, the
// unconstrained copy and both moments untouched -- TF would have raised before any update (dgp.py:92-98 under minimize).
__global__ void k_adam(float* __restrict__ params, float* __restrict__ free_, float* __restrict__ m, float* __restrict__ v,
                       float* __restrict__ grads, const unsigned char* __restrict__ kinds, size_t n,
                       const StepArgs* sa, Accum* acc, size_t off_likvar, int use_hi_lo, int do_adam, double* result) {
    const int status = acc->status;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        double e;
        if (use_hi_lo) e = (double)grads[n] + (double)grads[n + 1];
        else {
            e = acc->lik - sa->kl_weight * acc->kl;
            acc->elbo = e;
            const float hi = (float)e;
            grads[n] = hi; grads[n + 1] = (float)(e - (double)hi);
        }
        if (!use_hi_lo && off_likvar != (size_t)-1) grads[off_likvar] = (float)acc->glikvar;      // (k_elbo_finish's job)
        result[0] = e;
        result[1] = (double)status;
    }
    if (!do_adam || status != 0) return;
    const float lr_t = (float)sa->lr_t, b1 = (float)sa->beta1, b2 = (float)sa->beta2, eps = (float)sa->eps;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        unsigned char k = kinds[i];
        if (k >= 3) continue;
        const float gi = (!use_hi_lo && i == off_likvar) ? (float)acc->glikvar : grads[i];   // (thread 0 may not have stored it yet)
        float f = free_[i];
        float g = -gi;                              // objective = -ELBO
        if (k == 1) g *= 1.f / (1.f + __expf(-f));  // d softplus
        float mi = b1 * m[i] + (1.f - b1) * g;
        float vi = b2 * v[i] + (1.f - b2) * g * g;
        f -= lr_t * mi / (sqrtf(vi) + eps);
        m[i] = mi; v[i] = vi; free_[i] = f;
        params[i] = (k == 1) ? softplus_f(f) + POS_LOWER : f;
    }
}

// do_adam == 0: only the ELBO / result / likelihood-variance-gradient part (ELBO and gradient calls)
void launch_tail(float* params, float* free_, float* m, float* v, float* grads, const unsigned char* kinds, size_t n,
                 const StepArgs* sa, Accum* acc, size_t off_likvar, int use_hi_lo, int do_adam, double* result,
                 cudaStream_t st, long long* nl) {
    int nb = do_adam ? (int)min((size_t)592, (n + 255) / 256) : 1;
    k_adam<<<nb, 256, 0, st>>>(params, free_, m, v, grads, kinds, n, sa, acc, off_likvar, use_hi_lo, do_adam, result);
    *nl += 1;
}

__global__ void k_constrain_init(const float* __restrict__ params, float* __restrict__ free_,
                                 const unsigned char* __restrict__ kinds, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float p = params[i];
        if (kinds[i] == 1) {
            double y = (double)p - 1e-6;
            if (y < 1e-12) y = 1e-12;
            free_[i] = (float)(y + log(-expm1(-y)));      // softplus^-1
        } else free_[i] = p;
    }
}

void launch_constrain_init(const float* params, float* free_, const unsigned char* kinds, size_t n, cudaStream_t st,
                           long long* nl) {
    int nb = (int)min((size_t)592, (n + 255) / 256);
    k_constrain_init<<<nb, 256, 0, st>>>(params, free_, kinds, n);
    *nl += 1;
}
