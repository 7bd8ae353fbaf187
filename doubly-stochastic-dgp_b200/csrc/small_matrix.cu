// Per-step M x M work in fp64: Kuu Gram, Cholesky, Lu^-1, the KL term and its gradient, and the
// assembly of the parameter gradients from the row-reduced accumulators.
// Reference: layers.py:167-175 (build_cholesky_if_needed), layers.py:221-246 (KL); the gradients are what
// TF autodiff derives from those; the math is mirrored in tests/algo_mirror.py::layer_fin.
#include "dsdgp_internal.cuh"

// ----------------------------------------------------------------------------------------------
// prepA: one CTA per layer.  K = k(Z,Z) + jitter I ; Lu = chol(K) ; Linv = Lu^-1 (fused elimination).
// Working matrices live in shared memory when they fit, else in the global output buffers.
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_prepA(LayerSet ls, double jitter, Accum* acc, int use_smem) {
    const LayerDev& P = ls.l[blockIdx.x];
    const int M = P.M, Din = P.Din, nt = blockDim.x, tid = threadIdx.x;
    extern __shared__ double smd[];
    double* A = use_smem ? smd : P.Lu64;
    double* X = use_smem ? smd + (size_t)M * M : P.Linv64;
    const double var = (double)P.var[0];

    long long t0 = clock64();
    __shared__ double s_il[64];
    for (int q = tid; q < min(Din, 64); q += nt) s_il[q] = 1.0 / (double)P.ls[P.ard ? q : 0];
    __syncthreads();
    for (int idx = tid; idx < M * M; idx += nt) {
        int i = idx / M, j = idx % M;
        double r2 = 0.0;
        for (int q = 0; q < Din; ++q) {
            double il = q < 64 ? s_il[q] : 1.0 / (double)P.ls[P.ard ? q : 0];
            double d = ((double)P.Z[i * Din + q] - (double)P.Z[j * Din + q]) * il;
            r2 += d * d;
        }
        double k, kp;
        kern_eval_d(P.kern, r2, var, k, kp);
        if (i == j) k += jitter;
        P.K64[idx] = k;
        A[idx] = k;
        X[idx] = (i == j) ? 1.0 : 0.0;
    }
    __syncthreads();

    long long t1 = clock64();
    __shared__ int s_fail;
    __shared__ double s_piv[2];
    if (tid == 0) s_fail = 0;
    const int tx = tid & 31, ty = tid >> 5, nwarp = nt >> 5;
    for (int j = 0; j < M; ++j) {
        // phase A: scale column j of A (below the diagonal) and row j of X by 1/sqrt(A[j][j])
        // (the fp64 sqrt / reciprocal is done by one thread: replicated over 1024 threads it costs ~600 cycles of the
        //  SM's fp64 throughput per column)
        if (tid == 0) {
            double piv = A[j * M + j];
            if (!(piv > 0.0)) { s_fail = 1; piv = 1.0; }
            const double d = sqrt(piv);
            s_piv[0] = d; s_piv[1] = 1.0 / d;
            A[j * M + j] = d;
        }
        __syncthreads();
        const double id = s_piv[1];
        for (int i = j + 1 + tid; i < M; i += nt) A[i * M + j] *= id;
        for (int c = tid; c <= j; c += nt) X[j * M + c] *= id;
        __syncthreads();
        // phase B: trailing update of A (lower triangle) and elimination step on X.  Thread (ty, tx) owns the elements
        // (j+1+ty+32a, j+1+tx+32b) of A and (j+1+ty+32a, tx+32b) of X; the a/b loops are unrolled so that the loads of all
        // of a thread's elements are in flight together (the loop is latency-, not throughput-bound).
        {
            const int n = M - j - 1;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int ii = ty + 32 * a;
                if (ii < n) {
                    const int i = j + 1 + ii;
                    const double lij = A[i * M + j];
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const int kk = tx + 32 * b;
                        if (kk <= ii) { const int k = j + 1 + kk; A[i * M + k] -= lij * A[k * M + j]; }
                    }
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const int c = tx + 32 * b;
                        if (c <= j) X[i * M + c] -= lij * X[j * M + c];
                    }
                }
            }
            // matrices larger than 128: remaining rows / columns (rare; generic strided loops)
            if (M > 128) {
                for (int i = j + 1 + ty; i < M; i += nwarp) {
                    const double lij = A[i * M + j];
                    for (int k = j + 1 + tx; k <= i; k += 32)
                        if (i - (j + 1) >= 128 || k - (j + 1) >= 128) A[i * M + k] -= lij * A[k * M + j];
                    for (int c = tx; c <= j; c += 32)
                        if (i - (j + 1) >= 128 || c >= 128) X[i * M + c] -= lij * X[j * M + c];
                }
            }
        }
        __syncthreads();
    }
    long long t2 = clock64();
    if (tid == 0 && s_fail) atomicExch(&acc->status, blockIdx.x + 1);

    // outputs
    for (int idx = tid; idx < M * M; idx += nt) {
        int i = idx / M, j = idx % M;
        double l = (j <= i) ? A[idx] : 0.0, x = (j <= i) ? X[idx] : 0.0;
        if (use_smem) { P.Lu64[idx] = l; P.Linv64[idx] = x; }
        else { if (j > i) { P.Lu64[idx] = 0.0; P.Linv64[idx] = 0.0; } }
        P.Linv32[idx] = (float)x;
        P.LinvT32[j * M + i] = (float)x;
    }
    // sum log diag Lu
    double s = 0.0;
    for (int i = tid; i < M; i += nt) s += log(A[i * M + i]);
    s = warp_sum_d(s);
    __shared__ double red[32];
    if ((tid & 31) == 0) red[tid >> 5] = s;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < (nt + 31) / 32; ++w) t += red[w];
        P.scal[0] = t; P.scal[2] = 0.0; P.scal[3] = 0.0;
        P.scal[5] = (double)(t1 - t0); P.scal[6] = (double)(t2 - t1); P.scal[7] = (double)(clock64() - t2);
    }
}

// q_sqrtT[d][j][i] = q_sqrt[d][i][j]; scal[1] = sum log diag^2 ; scal[4] = sum q_sqrt^2 + sum q_mu^2
__global__ void k_qsqrtT(LayerSet ls) {
    const LayerDev& P = ls.l[blockIdx.y];
    const int M = P.M, D = P.Dout;
    size_t total = (size_t)D * M * M;
    double lg = 0.0, sq = 0.0;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        int d = idx / ((size_t)M * M), rem = idx % ((size_t)M * M), i = rem / M, j = rem % M;
        float v = (j <= i) ? P.q_sqrt[idx] : 0.0f;
        P.q_sqrtT[(size_t)d * M * M + (size_t)j * M + i] = v;
        if (i == j) lg += log((double)v * (double)v);
        sq += (double)v * (double)v;
    }
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < (size_t)M * D; idx += (size_t)gridDim.x * blockDim.x) {
        double m = P.q_mu[idx]; sq += m * m;
    }
    lg = warp_sum_d(lg); sq = warp_sum_d(sq);
    if ((threadIdx.x & 31) == 0) { atomicAdd(&P.scal[1], lg); atomicAdd(&P.scal[4], sq); }
}

__global__ void k_zero_scal(LayerSet ls) {
    int l = blockIdx.x;
    if (threadIdx.x == 0) { ls.l[l].scal[1] = 0.0; ls.l[l].scal[4] = 0.0; }
}

// Kinv = Linv^T Linv (blockIdx.z == 0) ; Ssum += L_d L_d^T + q_mu_d q_mu_d^T for d = blockIdx.z - 1   (non-white layers)
__global__ void k_kl0(LayerSet ls) {
    const LayerDev& P = ls.l[blockIdx.y];
    if (P.white) return;
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < P.M * P.M) P.Ssum64[idx] = 0.0;
}
__global__ void k_kl1(LayerSet ls) {
    const LayerDev& P = ls.l[blockIdx.y];
    if (P.white) return;
    const int M = P.M, D = P.Dout;
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * M) return;
    int i = idx / M, j = idx % M;
    int lo = max(i, j), hi = min(i, j);
    if (blockIdx.z == 0) {
        double s = 0.0;
        for (int k = lo; k < M; ++k) s += P.Linv64[k * M + i] * P.Linv64[k * M + j];
        P.Kinv64[idx] = s;
    } else {
        const int d = blockIdx.z - 1;
        if (d >= D) return;
        const float* Ld = P.q_sqrt + (size_t)d * M * M;
        double t = (double)P.q_mu[i * D + d] * (double)P.q_mu[j * D + d];
        for (int k = 0; k <= hi; ++k) t += (double)Ld[i * M + k] * (double)Ld[j * M + k];
        atomicAdd(&P.Ssum64[idx], t);
    }
}

// T1 = Kinv Ssum ; scal[2] += sum Kinv o Ssum
__global__ void k_kl2(LayerSet ls) {
    const LayerDev& P = ls.l[blockIdx.y];
    if (P.white) return;
    const int M = P.M;
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    double tr = 0.0;
    if (idx < M * M) {
        int i = idx / M, j = idx % M;
        double s = 0.0;
        for (int k = 0; k < M; ++k) s += P.Kinv64[i * M + k] * P.Ssum64[k * M + j];
        P.T1[idx] = s;
        tr = P.Kinv64[idx] * P.Ssum64[idx];
    }
    tr = warp_sum_d(tr);
    if ((threadIdx.x & 31) == 0) atomicAdd(&P.scal[2], tr);
}

// KbarKL = 1/2 D Kinv - 1/2 T1 Kinv
__global__ void k_kl3(LayerSet ls) {
    const LayerDev& P = ls.l[blockIdx.y];
    if (P.white) return;
    const int M = P.M;
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * M) return;
    int i = idx / M, j = idx % M;
    double s = 0.0;
    for (int k = 0; k < M; ++k) s += P.T1[i * M + k] * P.Kinv64[k * M + j];
    P.KbarKL[idx] = 0.5 * P.Dout * P.Kinv64[idx] - 0.5 * s;
}

// KL value per layer -> scal[3], acc->kl
__global__ void k_klval(LayerSet ls, Accum* acc) {
    int l = threadIdx.x;
    if (l >= ls.L) return;
    const LayerDev& P = ls.l[l];
    double KL = -0.5 * P.Dout * P.M - 0.5 * P.scal[1];
    if (P.white) KL += 0.5 * P.scal[4];
    else KL += P.Dout * P.scal[0] + 0.5 * P.scal[2];
    P.scal[3] = KL;
    atomicAdd(&acc->kl, KL);
}

// prepA on `st` (critical path: everything needs Lu / Linv); the KL-term preparation on `st_kl` (only the final gradient
// assembly and the ELBO scalar need it), which may be a side branch of the step DAG.
void launch_prep(const LayerSet& ls, double jitter, Accum* acc, const StepArgs* sa, cudaStream_t st, cudaStream_t st_kl,
                 cudaEvent_t ev_fork, long long* nl) {
    int Mmax = 0, Dmax = 0;
    for (int l = 0; l < ls.L; ++l) { Mmax = max(Mmax, ls.l[l].M); Dmax = max(Dmax, ls.l[l].Dout); }
    size_t sm = 2 * (size_t)Mmax * Mmax * sizeof(double);
    int use_smem = sm <= 200 * 1024;
    k_prepA<<<ls.L, 1024, use_smem ? sm : 0, st>>>(ls, jitter, acc, use_smem);
    if (st_kl != st) { cudaEventRecord(ev_fork, st); cudaStreamWaitEvent(st_kl, ev_fork, 0); }
    st = st_kl;
    k_zero_scal<<<ls.L, 32, 0, st>>>(ls);
    int nb = (Mmax * Mmax + 255) / 256;
    k_qsqrtT<<<dim3(min(4 * nb, 1024), ls.L), 256, 0, st>>>(ls);
    bool any_nonwhite = false;
    for (int l = 0; l < ls.L; ++l) any_nonwhite |= !ls.l[l].white;
    *nl += 3;
    if (any_nonwhite) {
        k_kl0<<<dim3(nb, ls.L), 256, 0, st>>>(ls);
        k_kl1<<<dim3(nb, ls.L, Dmax + 1), 256, 0, st>>>(ls);
        k_kl2<<<dim3(nb, ls.L), 256, 0, st>>>(ls);
        k_kl3<<<dim3(nb, ls.L), 256, 0, st>>>(ls);
        *nl += 4;
    }
    k_klval<<<1, 32, 0, st>>>(ls, acc);
    *nl += 1;
}

// ----------------------------------------------------------------------------------------------
// fin: parameter gradients from the row-reduced accumulators (tests/algo_mirror.py::layer_fin)
// ----------------------------------------------------------------------------------------------
// gq_sqrt[d] = tril((2 P_d - klw*Kinv) L_d) + klw*diag(1/L_d,ii)     (white: Kinv -> I)
__global__ void k_fin_qsqrt(LayerSet ls, const StepArgs* sa) {
    const LayerDev& P = ls.l[blockIdx.z];
    const int M = P.M, d = blockIdx.y;
    if (d >= P.Dout) return;
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * M) return;
    int i = idx / M, j = idx % M;
    const double klw = sa->kl_weight;
    const float* Ld = P.q_sqrt + (size_t)d * M * M;
    const float* Pd = P.Pd + (size_t)d * M * M;
    double g = 0.0;
    if (j <= i) {
        for (int k = j; k < M; ++k) {
            // symmetrise P_d (atomics make it symmetric only to rounding)
            double a = (double)Pd[i * M + k] + (double)Pd[k * M + i];
            if (!P.white) a -= klw * P.Kinv64[i * M + k];
            g += a * (double)Ld[k * M + j];
        }
        if (P.white) g -= klw * (double)Ld[i * M + j];
        if (i == j) g += klw / (double)Ld[i * M + i];
    }
    P.gq_sqrt[(size_t)d * M * M + idx] = (float)g;
}

// gq_mu = qmubar - klw * Kinv q_mu   (white: - klw q_mu)
__global__ void k_fin_qmu(LayerSet ls, const StepArgs* sa) {
    const LayerDev& P = ls.l[blockIdx.y];
    const int M = P.M, D = P.Dout;
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * D) return;
    int i = idx / D, d = idx % D;
    const double klw = sa->kl_weight;
    double g = (double)P.qmubar[idx];
    if (P.white) g -= klw * (double)P.q_mu[idx];
    else {
        double s = 0.0;
        for (int k = 0; k < M; ++k) s += P.Kinv64[i * M + k] * (double)P.q_mu[k * D + d];
        g -= klw * s;
    }
    P.gq_mu[idx] = (float)g;
}

// white: Phi = tril(Lu^T tril(-G)) with halved diagonal -> T1 ; then T1 <- Phi Linv (into Ssum64) ;
// Kbar = sym(Linv^T (Phi Linv))
__global__ void k_fin_w1(LayerSet ls) {
    const LayerDev& P = ls.l[blockIdx.y];
    if (!P.white) return;
    const int M = P.M;
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * M) return;
    int i = idx / M, j = idx % M;
    double s = 0.0;
    if (j <= i) {
        for (int k = i; k < M; ++k) s += P.Lu64[k * M + i] * (-(double)P.G[k * M + j]);   // Lbar = -tril(G): k >= j holds since k>=i>=j
        if (i == j) s *= 0.5;
    }
    P.T1[idx] = s;
}
__global__ void k_fin_w2(LayerSet ls) {
    const LayerDev& P = ls.l[blockIdx.y];
    if (!P.white) return;
    const int M = P.M;
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * M) return;
    int i = idx / M, j = idx % M;
    double s = 0.0;
    for (int k = j; k <= i; ++k) s += P.T1[i * M + k] * P.Linv64[k * M + j];
    P.Ssum64[idx] = s;
}
__global__ void k_fin_w3(LayerSet ls) {
    const LayerDev& P = ls.l[blockIdx.y];
    if (!P.white) return;
    const int M = P.M;
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * M) return;
    int i = idx / M, j = idx % M;
    double s = 0.0;
    for (int k = i; k < M; ++k) s += P.Linv64[k * M + i] * P.Ssum64[k * M + j];
    P.KbarKL[idx] = s;      // unsymmetrised Kbar
}

// g_ij = Kbar_ij * dk/dr2_ij  (Kbar symmetric)  -> Gsym ; gvar += sum Kbar o k / var
__global__ void k_fin_kbar(LayerSet ls, const StepArgs* sa) {
    const LayerDev& P = ls.l[blockIdx.y];
    const int M = P.M, Din = P.Din;
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    double s2 = 0.0;
    if (idx < M * M) {
        int i = idx / M, j = idx % M;
        double kb;
        if (P.white) kb = 0.5 * (P.KbarKL[i * M + j] + P.KbarKL[j * M + i]);
        else kb = -0.5 * ((double)P.G[i * M + j] + (double)P.G[j * M + i]) - sa->kl_weight * P.KbarKL[idx];
        double r2 = 0.0;
        for (int q = 0; q < Din; ++q) {
            double il = 1.0 / (double)P.ls[P.ard ? q : 0];
            double d = ((double)P.Z[i * Din + q] - (double)P.Z[j * Din + q]) * il;
            r2 += d * d;
        }
        double k, kp;
        kern_eval_d(P.kern, r2, (double)P.var[0], k, kp);
        P.Gsym[idx] = kb * kp;
        s2 = kb * k / (double)P.var[0];
    }
    s2 = warp_sum_d(s2);
    if ((threadIdx.x & 31) == 0) atomicAdd(P.gvar, (float)s2);
}

// Zbar_iq += 4/l_q^2 sum_j g_ij (z_iq - z_jq) ; lsbar_q += -2/l_q^3 sum_ij g_ij (z_iq - z_jq)^2
__global__ void k_fin_kuu(LayerSet ls) {
    const LayerDev& P = ls.l[blockIdx.y];
    const int M = P.M, Din = P.Din;
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * Din) return;
    int i = idx / Din, q = idx % Din;
    double zi = (double)P.Z[idx], a = 0.0, b = 0.0;
    for (int j = 0; j < M; ++j) {
        double d = zi - (double)P.Z[j * Din + q], g = P.Gsym[i * M + j];
        a += g * d; b += g * d * d;
    }
    double l = (double)P.ls[P.ard ? q : 0];
    atomicAdd(&P.gZ[idx], (float)(4.0 * a / (l * l)));
    atomicAdd(&P.gls[P.ard ? q : 0], (float)(-2.0 * b / (l * l * l)));
}

__global__ void k_elbo_finish(Accum* acc, const StepArgs* sa, float* glikvar, float* elbo_hi_lo) {
    double e = acc->lik - sa->kl_weight * acc->kl;
    acc->elbo = e;
    if (glikvar) *glikvar = (float)acc->glikvar;
    if (elbo_hi_lo) {
        float hi = (float)e;
        elbo_hi_lo[0] = hi;
        elbo_hi_lo[1] = (float)(e - (double)hi);
    }
}

void launch_fin(const LayerSet& ls, Accum* acc, const StepArgs* sa, cudaStream_t st, long long* nl) {
    int Mmax = 0, Dmax = 0, MDmax = 0, MDin = 0;
    bool any_white = false;
    for (int l = 0; l < ls.L; ++l) {
        Mmax = max(Mmax, ls.l[l].M); Dmax = max(Dmax, ls.l[l].Dout);
        MDmax = max(MDmax, ls.l[l].M * ls.l[l].Dout); MDin = max(MDin, ls.l[l].M * ls.l[l].Din);
        any_white |= ls.l[l].white != 0;
    }
    int nb = (Mmax * Mmax + 255) / 256;
    k_fin_qsqrt<<<dim3(nb, Dmax, ls.L), 256, 0, st>>>(ls, sa);
    k_fin_qmu<<<dim3((MDmax + 255) / 256, ls.L), 256, 0, st>>>(ls, sa);
    *nl += 2;
    if (any_white) {
        k_fin_w1<<<dim3(nb, ls.L), 256, 0, st>>>(ls);
        k_fin_w2<<<dim3(nb, ls.L), 256, 0, st>>>(ls);
        k_fin_w3<<<dim3(nb, ls.L), 256, 0, st>>>(ls);
        *nl += 3;
    }
    k_fin_kbar<<<dim3(nb, ls.L), 256, 0, st>>>(ls, sa);
    k_fin_kuu<<<dim3((MDin + 255) / 256, ls.L), 256, 0, st>>>(ls);
    *nl += 2;
}

void launch_elbo_finish(Accum* acc, const StepArgs* sa, float* glikvar, float* elbo_hi_lo, cudaStream_t st, long long* nl) {
    // glikvar points at the lik-variance slot of the flat gradient buffer; the two floats after the last
    // gradient hold the ELBO as (hi, lo) so that one all-reduce covers ELBO and gradient.
    k_elbo_finish<<<1, 1, 0, st>>>(acc, sa, glikvar, elbo_hi_lo);
    *nl += 1;
}

// result[0] = ELBO (after the all-reduce when a communicator is attached), result[1] = status
__global__ void k_result(const Accum* acc, const float* elbo_hi_lo, int use_hi_lo, double* result) {
    result[0] = use_hi_lo ? (double)elbo_hi_lo[0] + (double)elbo_hi_lo[1] : acc->elbo;
    result[1] = (double)acc->status;
}
void launch_result(const Accum* acc, const float* elbo_hi_lo, int use_hi_lo, double* result, cudaStream_t st, long long* nl) {
    k_result<<<1, 1, 0, st>>>(acc, elbo_hi_lo, use_hi_lo, result);
    *nl += 1;
}

cudaError_t small_matrix_init() {
    return cudaFuncSetAttribute(k_prepA, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
}
