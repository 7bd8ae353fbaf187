// Natural-gradient step on (q_mu, q_sqrt) of one layer, fp64, from the row-reduced accumulators of the backward pass.
// Replaces gpflow.training.NatGradOptimizer(gamma).minimize(model, var_list=[[layer.q_mu, layer.q_sqrt]], maxiter=1) as
// the reference calls it (tests/test_collapsed.py:99-100, demos/using_natural_gradients.ipynb, demos/
// demo_regression_UCI.ipynb:357-366).  Math (mirrored in tests/algo_mirror.py::natgrad_update and checked on
// CPU against the theta-space restatement of GPflow's optimiser, tests/test_natgrad_cpu.py):
//     dELBO/dS_d = P_d - 1/2 Prior^-1 + 1/2 S_d^-1 ,   dELBO/dm_d = qmubar_d - Prior^-1 m_d      (Prior = Kuu, or I if white)
//     -2 theta2' = (1-gamma) S_d^-1 + gamma (Prior^-1 - 2 P_d)
//        theta1' = (1-gamma) S_d^-1 m_d + gamma (qmubar_d - 2 P_d m_d)
//     C = chol(-2 theta2') ; V = C^-1 ; S' = V^T V ; m' = S' theta1' ; q_sqrt' = chol(S')      (natural_to_meanvarsqrt)
// One (layer, d) matrix per blockIdx.y; every matrix product is one thread per output element with coalesced operand rows.
#include "dsdgp_internal.cuh"

// A (M x M row-major, symmetric positive definite, lower triangle read) -> A = chol(A) (lower; strict upper left as is),
// X = chol(A)^-1 (lower, strict upper zero).  Fused right-looking elimination, the algorithm of k_prepA (small_matrix.cu)
// with generic strided loops.  One CTA; A, X may live in shared or global memory.
__device__ void cta_chol_inv(double* A, double* X, int M, int* s_fail, double* s_piv) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
    for (int idx = tid; idx < M * M; idx += nt) X[idx] = (idx / M == idx % M) ? 1.0 : 0.0;
    __syncthreads();
    for (int j = 0; j < M; ++j) {
        if (tid == 0) {
            double piv = A[j * M + j];
            if (!(piv > 0.0)) { *s_fail = 1; piv = 1.0; }
            const double d = sqrt(piv);
            s_piv[0] = d; s_piv[1] = 1.0 / d;
            A[j * M + j] = d;
        }
        __syncthreads();
        const double id = s_piv[1];
        for (int i = j + 1 + tid; i < M; i += nt) A[i * M + j] *= id;
        for (int c = tid; c <= j; c += nt) X[j * M + c] *= id;
        __syncthreads();
        for (int i = j + 1 + warp; i < M; i += nwarp) {
            const double lij = A[i * M + j];
            for (int k = j + 1 + lane; k <= i; k += 32) A[i * M + k] -= lij * A[k * M + j];
            for (int c = lane; c <= j; c += 32) X[i * M + c] -= lij * X[j * M + c];
        }
        __syncthreads();
    }
}

// batched over blockIdx.x = d: Ag/Xg are D x M x M.  status != 0 afterwards: some matrix was not positive definite.
__global__ void __launch_bounds__(1024) k_ng_cholinv(double* Ag, double* Xg, int M, int use_smem, int* status) {
    extern __shared__ double smd[];
    __shared__ int s_fail;
    __shared__ double s_piv[2];
    double* Ad = Ag + (size_t)blockIdx.x * M * M;
    double* Xd = Xg + (size_t)blockIdx.x * M * M;
    double *A = Ad, *X = Xd;
    if (threadIdx.x == 0) s_fail = 0;
    if (use_smem) {
        A = smd; X = smd + (size_t)M * M;
        for (int idx = threadIdx.x; idx < M * M; idx += blockDim.x) A[idx] = Ad[idx];
    }
    __syncthreads();
    cta_chol_inv(A, X, M, &s_fail, s_piv);
    if (use_smem)
        for (int idx = threadIdx.x; idx < M * M; idx += blockDim.x) { Ad[idx] = A[idx]; Xd[idx] = X[idx]; }
    if (threadIdx.x == 0 && s_fail) atomicExch(status, 1);
}

// S_d = L_d L_d^T from the fp32 parameters; q_sqrtT (k-major copy written by k_qsqrtT in the step's prep) keeps the loads coalesced
__global__ void k_ng_S(LayerDev P, double* W) {
    const int M = P.M, d = blockIdx.y;
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * M) return;
    int i = idx / M, j = idx % M;
    const float* L = P.q_sqrt + (size_t)d * M * M;
    const float* LT = P.q_sqrtT + (size_t)d * M * M;
    double s = 0.0;
    const int kmax = min(i, j);
    for (int k = 0; k <= kmax; ++k) s += (double)L[i * M + k] * (double)LT[k * M + j];
    W[(size_t)d * M * M + idx] = s;
}

// out = X^T X for lower-triangular X
__global__ void k_ng_ata(const double* __restrict__ Xg, double* __restrict__ Og, int M) {
    const int d = blockIdx.y;
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * M) return;
    int i = idx / M, j = idx % M;
    const double* X = Xg + (size_t)d * M * M;
    double s = 0.0;
    for (int k = max(i, j); k < M; ++k) s += X[k * M + i] * X[k * M + j];
    Og[(size_t)d * M * M + idx] = s;
}

// Pi = -2 theta2' ; t1 = theta1'.   Sinv may be NULL when gamma == 1 (the old S drops out).
__global__ void k_ng_theta(LayerDev P, const double* __restrict__ Sinv, double gamma, double* __restrict__ Pi) {
    const int M = P.M, d = blockIdx.y;
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * M) return;
    int i = idx / M, j = idx % M;
    const float* Pd = P.Pd + (size_t)d * M * M;
    const double prior = P.white ? (i == j ? 1.0 : 0.0) : P.Kinv64[idx];
    double v = gamma * (prior - ((double)Pd[i * M + j] + (double)Pd[j * M + i]));
    if (Sinv) v += (1.0 - gamma) * Sinv[(size_t)d * M * M + idx];
    Pi[(size_t)d * M * M + idx] = v;
}
// one warp per row i
__global__ void k_ng_t1(LayerDev P, const double* __restrict__ Sinv, double gamma, double* __restrict__ t1) {
    const int M = P.M, D = P.Dout, d = blockIdx.y;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= M) return;
    const float* Pd = P.Pd + (size_t)d * M * M;
    double s = 0.0;
    for (int j = lane; j < M; j += 32) {
        double a = -gamma * ((double)Pd[i * M + j] + (double)Pd[j * M + i]);
        if (Sinv) a += (1.0 - gamma) * Sinv[(size_t)d * M * M + i * M + j];
        s += a * (double)P.q_mu[j * D + d];
    }
    s = warp_sum_d(s);
    if (lane == 0) t1[(size_t)d * M + i] = s + gamma * (double)P.qmubar[i * D + d];
}

// q_mu[:, d] = S' t1   (skipped when any factorisation failed: parameters stay as they were)
__global__ void k_ng_newmu(LayerDev P, const double* __restrict__ Snew, const double* __restrict__ t1, float* q_mu,
                           const int* status) {
    if (*status) return;
    const int M = P.M, D = P.Dout, d = blockIdx.y;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= M) return;
    const double* S = Snew + (size_t)d * M * M;
    double s = 0.0;
    for (int j = lane; j < M; j += 32) s += S[i * M + j] * t1[(size_t)d * M + j];
    s = warp_sum_d(s);
    if (lane == 0) q_mu[i * D + d] = (float)s;
}

// q_sqrt[d] = tril(chol S')
__global__ void k_ng_store(LayerDev P, const double* __restrict__ Lnew, float* q_sqrt, const int* status) {
    if (*status) return;
    const int M = P.M, d = blockIdx.y;
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * M) return;
    int i = idx / M, j = idx % M;
    q_sqrt[(size_t)d * M * M + idx] = (j <= i) ? (float)Lnew[(size_t)d * M * M + idx] : 0.f;
}

size_t natgrad_ws_doubles(int M, int D) { return (size_t)3 * D * M * M + (size_t)D * M; }

// ws: natgrad_ws_doubles(M, D) doubles.  q_mu / q_sqrt: staging buffers (M x D, D x M x M) committed by launch_natgrad_commit.
void launch_natgrad_layer(const LayerDev& P, double gamma, double* ws, int* status, float* q_mu, float* q_sqrt,
                          cudaStream_t st, long long* nl) {
    const int M = P.M, D = P.Dout;
    const size_t mm = (size_t)M * M;
    double *W0 = ws, *W1 = ws + D * mm, *W2 = ws + 2 * D * mm, *t1 = ws + 3 * D * mm;
    const int nb = (int)((mm + 255) / 256);
    const size_t sm = 2 * mm * sizeof(double);
    const int use_smem = sm <= 200 * 1024;
    const size_t smb = use_smem ? sm : 0;
    const dim3 gmm(nb, D), grow((M + 7) / 8, D);
    const double* Sinv = nullptr;
    if (gamma != 1.0) {
        k_ng_S<<<gmm, 256, 0, st>>>(P, W0);
        k_ng_cholinv<<<D, 1024, smb, st>>>(W0, W1, M, use_smem, status);
        k_ng_ata<<<gmm, 256, 0, st>>>(W1, W2, M);
        Sinv = W2;
        *nl += 3;
    }
    k_ng_t1<<<grow, 256, 0, st>>>(P, Sinv, gamma, t1);
    k_ng_theta<<<gmm, 256, 0, st>>>(P, Sinv, gamma, W0);
    k_ng_cholinv<<<D, 1024, smb, st>>>(W0, W1, M, use_smem, status);      // W1 = V = chol(-2 theta2')^-1
    k_ng_ata<<<gmm, 256, 0, st>>>(W1, W2, M);                             // W2 = S'
    k_ng_newmu<<<grow, 256, 0, st>>>(P, W2, t1, q_mu, status);
    k_ng_cholinv<<<D, 1024, smb, st>>>(W2, W1, M, use_smem, status);      // W2 = chol(S')
    k_ng_store<<<gmm, 256, 0, st>>>(P, W2, q_sqrt, status);
    *nl += 7;
}

// dst[0..n) = src[0..n) unless *status != 0: commits the staged (q_mu, q_sqrt) of a natural-gradient step only if EVERY
// factorisation of EVERY updated layer succeeded (api.cu: dsdgp_natgrad_step), so a failing step changes nothing.
__global__ void k_ng_commit(float* __restrict__ dst, const float* __restrict__ src, size_t n, const int* status) {
    if (*status) return;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
void launch_natgrad_commit(float* dst, const float* src, size_t n, const int* status, cudaStream_t st, long long* nl) {
    k_ng_commit<<<(unsigned)min((size_t)592, (n + 255) / 256), 256, 0, st>>>(dst, src, n, status);
    *nl += 1;
}

cudaError_t natgrad_init() {
    return cudaFuncSetAttribute(k_ng_cholinv, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
}
