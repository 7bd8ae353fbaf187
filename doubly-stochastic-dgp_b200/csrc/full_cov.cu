// full_cov=True propagation (the prediction / plotting path of the reference: predict_f_full_cov,
// predict_all_layers_full_cov, dgp.py:104-114), float64 throughout -- the reference hard-codes float64 here
// (layers.py:68) and so do we: this is not the training hot path, it is the N* x N* posterior covariance per (sample,
// output) and a Cholesky-based joint draw.
//   per layer, per sample s (independent, layers.py:66-69):  Kuf = k(Z, X_s) ; A = Kuu^-1 Kuf (white: Lu^-1 Kuf)   layers.py:184-188
//     mean = A^T q_mu + mean_function(X_s)                                                                  :190,219
//     SK_d = q_sqrt_d q_sqrt_d^T - Ku (white: - I) ; B_d = SK_d A ; var_d = k(X_s, X_s) + A^T B_d            :194-217
//     f[:, d] = mean[:, d] + chol(var_d + jitter I) z[:, d]                                                  utils.py:43-51
// Every matrix product is one thread per output element with the contraction index walking coalesced rows.
#include "dsdgp_internal.cuh"

// Kuf[s] (M x N) and Kff[s] (N x N).  Xin: (S, N, Din) fp64 with sample stride xs (0: the same X for every s, layer 1).
__global__ void k_fc_gram(LayerDev P, const double* __restrict__ Xin, size_t xs, int N, double* __restrict__ Kuf,
                          double* __restrict__ Kff) {
    const int M = P.M, Din = P.Din, s = blockIdx.y;
    const size_t nuf = (size_t)M * N, nff = (size_t)N * N;
    const double* X = Xin + (size_t)s * xs;
    const double var = (double)P.var[0];
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < nuf + nff; idx += (size_t)gridDim.x * blockDim.x) {
        double r2 = 0.0;
        if (idx < nuf) {
            const int i = (int)(idx / N), n = (int)(idx % N);
            for (int q = 0; q < Din; ++q) {
                const double d = ((double)P.Z[(size_t)i * Din + q] - X[(size_t)n * Din + q]) / (double)P.ls[P.ard ? q : 0];
                r2 += d * d;
            }
        } else {
            const size_t e = idx - nuf;
            const int a = (int)(e / N), b = (int)(e % N);
            for (int q = 0; q < Din; ++q) {
                const double d = (X[(size_t)a * Din + q] - X[(size_t)b * Din + q]) / (double)P.ls[P.ard ? q : 0];
                r2 += d * d;
            }
        }
        double k, kp;
        kern_eval_d(P.kern, r2, var, k, kp);
        if (idx < nuf) Kuf[(size_t)s * nuf + idx] = k;
        else {
            const size_t e = idx - nuf;
            if (e / N == e % N) k += (double)P.wvar[0];        // White.K(X) = variance_w I
            Kff[(size_t)s * nff + e] = k;
        }
    }
}

// A[s] = W Kuf[s],  W = Kuu^-1 (non-white) or Lu^-1 (white), M x M fp64 from the step's prep kernels
__global__ void k_fc_A(LayerDev P, int N, const double* __restrict__ Kuf, double* __restrict__ A) {
    const int M = P.M, s = blockIdx.y;
    const size_t nuf = (size_t)M * N;
    const double* W = P.white ? P.Linv64 : P.Kinv64;
    const double* K = Kuf + (size_t)s * nuf;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < nuf; idx += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx / N), n = (int)(idx % N);
        double acc = 0.0;
        const int kend = P.white ? i + 1 : M;          // Lu^-1 is lower triangular
        for (int k = 0; k < kend; ++k) acc += W[(size_t)i * M + k] * K[(size_t)k * N + n];
        A[(size_t)s * nuf + idx] = acc;
    }
}

// SK[d] = q_sqrt_d q_sqrt_d^T - Ku  (white: - I);  Ku includes the jitter (layers.py:171)
__global__ void k_fc_SK(LayerDev P, double* __restrict__ SK) {
    const int M = P.M, d = blockIdx.y;
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * M) return;
    const int i = idx / M, j = idx % M;
    const float* L = P.q_sqrt + (size_t)d * M * M;
    double s = 0.0;
    const int kmax = min(i, j);
    for (int k = 0; k <= kmax; ++k) s += (double)L[(size_t)i * M + k] * (double)L[(size_t)j * M + k];
    s -= P.white ? (i == j ? 1.0 : 0.0) : P.K64[idx];
    SK[(size_t)d * M * M + idx] = s;
}

// B[s][d] = SK[d] A[s]      (blockIdx.y = s * D + d)
__global__ void k_fc_B(LayerDev P, int N, const double* __restrict__ SK, const double* __restrict__ A, double* __restrict__ B) {
    const int M = P.M, D = P.Dout, s = blockIdx.y / D, d = blockIdx.y % D;
    const size_t nuf = (size_t)M * N;
    const double* Sd = SK + (size_t)d * M * M;
    const double* As = A + (size_t)s * nuf;
    double* Bo = B + (size_t)blockIdx.y * nuf;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < nuf; idx += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx / N), n = (int)(idx % N);
        double acc = 0.0;
        for (int k = 0; k < M; ++k) acc += Sd[(size_t)i * M + k] * As[(size_t)k * N + n];
        Bo[idx] = acc;
    }
}

// cov[s][d] = Kff[s] + A[s]^T B[s][d]  (N x N, fp64, kept for the factorisation) and the fp32 copy in the reference's
// output layout (S, N, N, D)
__global__ void k_fc_cov(LayerDev P, int N, const double* __restrict__ Kff, const double* __restrict__ A,
                         const double* __restrict__ B, double* __restrict__ cov, float* __restrict__ var_out) {
    const int M = P.M, D = P.Dout, s = blockIdx.y / D, d = blockIdx.y % D;
    const size_t nuf = (size_t)M * N, nff = (size_t)N * N;
    const double* As = A + (size_t)s * nuf;
    const double* Bs = B + (size_t)blockIdx.y * nuf;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < nff; idx += (size_t)gridDim.x * blockDim.x) {
        const int a = (int)(idx / N), b = (int)(idx % N);
        double acc = Kff[(size_t)s * nff + idx];
        for (int i = 0; i < M; ++i) acc += As[(size_t)i * N + a] * Bs[(size_t)i * N + b];
        cov[(size_t)blockIdx.y * nff + idx] = acc;
        if (var_out) var_out[((size_t)s * nff + idx) * D + d] = (float)acc;
    }
}

// mean[s][n][d] = sum_i A[s][i][n] q_mu[i][d] + mean_function(X_s[n])_d
__global__ void k_fc_mean(LayerDev P, const double* __restrict__ Xin, size_t xs, int N, const double* __restrict__ A,
                          double* __restrict__ mean, float* __restrict__ mean_out) {
    const int M = P.M, D = P.Dout, Din = P.Din, s = blockIdx.y;
    const double* X = Xin + (size_t)s * xs;
    const double* As = A + (size_t)s * M * N;
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * D) return;
    const int n = idx / D, d = idx % D;
    double acc = 0.0;
    for (int i = 0; i < M; ++i) acc += As[(size_t)i * N + n] * (double)P.q_mu[(size_t)i * D + d];
    if (P.mean == DSDGP_MEAN_IDENTITY) acc += X[(size_t)n * Din + d];
    else if (P.mean == DSDGP_MEAN_LINEAR) {
        double m = (double)P.meanB[d];
        for (int q = 0; q < Din; ++q) m += X[(size_t)n * Din + q] * (double)P.meanW[(size_t)q * D + d];
        acc += m;
    }
    mean[(size_t)s * N * D + idx] = acc;
    if (mean_out) mean_out[(size_t)s * N * D + idx] = (float)acc;
}

// One CTA per (s, d): C = chol(cov + jitter I) in place (lower), then f[s][n][d] = mean + sum_{n' <= n} C[n][n'] z[s][n'][d].
// z == NULL: Philox normals keyed by (seed, layer, s + s_offset, n + n_offset, d) like the diagonal path.
__global__ void __launch_bounds__(1024) k_fc_chol_draw(int layer, int N, int D, double jitter, double* __restrict__ cov,
                                                       const double* __restrict__ mean, const float* __restrict__ z,
                                                       const StepArgs* sa, double* __restrict__ F, float* __restrict__ F_out,
                                                       int* status) {
    extern __shared__ double col[];          // N: the scaled pivot column; then reused for z
    __shared__ double s_piv;
    __shared__ int s_fail;
    const int s = blockIdx.x / D, d = blockIdx.x % D;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
    double* C = cov + (size_t)blockIdx.x * N * N;
    if (tid == 0) s_fail = 0;
    for (int i = tid; i < N; i += nt) C[(size_t)i * N + i] += jitter;
    __syncthreads();
    for (int j = 0; j < N; ++j) {
        if (tid == 0) {
            double piv = C[(size_t)j * N + j];
            if (!(piv > 0.0)) { s_fail = 1; piv = 1.0; }
            const double dg = sqrt(piv);
            C[(size_t)j * N + j] = dg;
            s_piv = 1.0 / dg;
        }
        __syncthreads();
        const double id = s_piv;
        for (int i = j + 1 + tid; i < N; i += nt) { const double v = C[(size_t)i * N + j] * id; C[(size_t)i * N + j] = v; col[i] = v; }
        __syncthreads();
        for (int i = j + 1 + warp; i < N; i += nwarp) {
            const double lij = col[i];
            for (int k = j + 1 + lane; k <= i; k += 32) C[(size_t)i * N + k] -= lij * col[k];
        }
        __syncthreads();
    }
    if (tid == 0 && s_fail) atomicExch(status, 1);
    for (int n = tid; n < N; n += nt) {
        double zz;
        if (z) zz = (double)z[((size_t)s * N + n) * D + d];
        else zz = (double)dsdgp_normal(sa->seed, layer, s + sa->s_offset, n + sa->n_offset, d);
        col[n] = zz;
    }
    __syncthreads();
    for (int n = warp; n < N; n += nwarp) {
        double acc = 0.0;
        for (int k = lane; k <= n; k += 32) acc += C[(size_t)n * N + k] * col[k];
        acc = warp_sum_d(acc);
        if (lane == 0) {
            const size_t o = ((size_t)s * N + n) * D + d;
            const double f = mean[o] + acc;
            F[o] = f;
            if (F_out) F_out[o] = (float)f;
        }
    }
}

__global__ void k_fc_x64(const float* __restrict__ X, size_t n, double* __restrict__ X64) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) X64[i] = (double)X[i];
}

// doubles of workspace for one layer at (N, S): Kuf, A (S M N each), Kff (S N N), SK (D M M), B (S D M N), cov (S D N N),
// mean (S N D), and two activation buffers (S N Dmax)
size_t full_cov_ws_doubles(int M, int D, int Dmax_io, int N, int S) {
    const size_t s = S, n = N, m = M, d = D;
    return 2 * s * m * n + s * n * n + d * m * m + s * d * m * n + s * d * n * n + s * n * d + 2 * s * n * (size_t)Dmax_io;
}

void launch_full_cov_x64(const float* X, size_t n, double* X64, cudaStream_t st, long long* nl) {
    k_fc_x64<<<(unsigned)min((size_t)1184, (n + 255) / 256), 256, 0, st>>>(X, n, X64);
    *nl += 1;
}

// One layer.  Xin: fp64 (S,N,Din) with stride xs (0 for layer 1).  Fnext: fp64 (S,N,Dout).  The *_out pointers are fp32
// device buffers in the reference's layouts, or NULL.
void launch_full_cov_layer(const LayerDev& P, const double* Xin, size_t xs, int N, int S, double jitter, const float* z,
                           const StepArgs* sa, double* ws, double* Fnext, float* F_out, float* mean_out, float* var_out,
                           int* status, cudaStream_t st, long long* nl) {
    const int M = P.M, D = P.Dout;
    const size_t nuf = (size_t)M * N, nff = (size_t)N * N;
    double* Kuf = ws;
    double* A = Kuf + S * nuf;
    double* Kff = A + S * nuf;
    double* SK = Kff + S * nff;
    double* B = SK + (size_t)D * M * M;
    double* cov = B + (size_t)S * D * nuf;
    double* mean = cov + (size_t)S * D * nff;
    auto blocks = [](size_t n) { return (unsigned)min((size_t)4096, (n + 255) / 256); };
    k_fc_gram<<<dim3(blocks(nuf + nff), S), 256, 0, st>>>(P, Xin, xs, N, Kuf, Kff);
    k_fc_A<<<dim3(blocks(nuf), S), 256, 0, st>>>(P, N, Kuf, A);
    k_fc_SK<<<dim3((M * M + 255) / 256, D), 256, 0, st>>>(P, SK);
    k_fc_B<<<dim3(blocks(nuf), S * D), 256, 0, st>>>(P, N, SK, A, B);
    k_fc_cov<<<dim3(blocks(nff), S * D), 256, 0, st>>>(P, N, Kff, A, B, cov, var_out);
    k_fc_mean<<<dim3((N * D + 255) / 256, S), 256, 0, st>>>(P, Xin, xs, N, A, mean, mean_out);
    k_fc_chol_draw<<<S * D, 1024, (size_t)N * sizeof(double), st>>>(P.idx, N, D, jitter, cov, mean, z, sa, Fnext, F_out, status);
    *nl += 7;
}

cudaError_t full_cov_init() {
    return cudaFuncSetAttribute(k_fc_chol_draw, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
}
