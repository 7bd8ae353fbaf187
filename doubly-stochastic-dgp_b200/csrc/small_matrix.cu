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
        if (i == j) k += jitter + (double)P.wvar[0];
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


// fp64 reciprocal for the pivot chains below: hardware approximation (>= 20 bits) + two Newton steps (relative error ~2^-80
// before rounding) -- five dependent instructions instead of the IEEE division's ~40; pivots are normal, positive numbers.
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}

// ----------------------------------------------------------------------------------------------
// prepA, fast path (M <= 111: both working matrices in shared memory).  Same results as k_prepA, restructured around the
// latency of the pivot loop, which is what bounds it (measured: 2.9k cycles per pivot in k_prepA, 147 us of a 1.2 ms step):
//   * square-root-free elimination: K = Lt D Lt^T is factored with the pivot d_j read by every thread (one fp64 reciprocal
//     each, no designated thread, no broadcast) -> ONE __syncthreads per pivot instead of three; Lu = Lt D^1/2 and
//     Lu^-1 = D^-1/2 Lt^-1 are formed once at the end (M square roots in parallel);
//   * the factor is held transposed, B[j][i] = A_j[i][j] (i >= j): the pivot column is a contiguous row, so with lanes
//     walking i every shared-memory access is stride-1 (the row-major form reads the pivot column with stride M: 8-way
//     bank conflicts on fp64); each lane keeps its slice of the pivot row (and of row j of the inverse) in registers for all
//     the rows its warp updates;
//   * the Gram matrix is evaluated on the lower triangle only (fp64 exp is the cost) and mirrored.
// ----------------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(NT) k_prepA_ldl(LayerSet ls, double jitter, Accum* acc) {
    const LayerDev& P = ls.l[blockIdx.x];
    const int M = P.M, Din = P.Din, tid = threadIdx.x, MS = M | 1;
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int NW = NT / 32;
    extern __shared__ double smd[];
    double* B = smd;                          // [M][MS]  upper: B[j][i], i >= j
    double* X = smd + (size_t)M * MS;         // [M][MS]  lower: X[i][c], c <= i  (unit-lower inverse of Lt)
    __shared__ double s_il[64];
    __shared__ double s_rsq[128];
    __shared__ int s_fail;
    const double var = (double)P.var[0];

    long long t0 = clock64();
    if (tid == 0) s_fail = 0;
    for (int q = tid; q < min(Din, 64); q += NT) s_il[q] = 1.0 / (double)P.ls[P.ard ? q : 0];
    for (int idx = tid; idx < M * MS; idx += NT) { X[idx] = (idx / MS == idx % MS) ? 1.0 : 0.0; B[idx] = 0.0; }
    __syncthreads();
    // Gram on i >= j: rows p and M-1-p folded into one line of M+1 entries so that all threads carry the same load
    {
        const int H = (M + 1) / 2, W = M + 1;
        for (int idx = tid; idx < H * W; idx += NT) {
            const int p = idx / W, q = idx % W;
            int i, j;
            if (q <= p) { i = p; j = q; }
            else { i = M - 1 - p; j = q - p - 1; if (i == p) continue; }        // odd M: the middle row is not folded twice
            double r2 = 0.0;
            for (int qq = 0; qq < Din; ++qq) {
                const double il = qq < 64 ? s_il[qq] : 1.0 / (double)P.ls[P.ard ? qq : 0];
                const double d = ((double)P.Z[i * Din + qq] - (double)P.Z[j * Din + qq]) * il;
                r2 += d * d;
            }
            double k, kp;
            kern_eval_d(P.kern, r2, var, k, kp);
            if (i == j) k += jitter + (double)P.wvar[0];
            P.K64[i * M + j] = k;
            P.K64[j * M + i] = k;
            B[j * MS + i] = k;
        }
    }
    long long t1 = clock64();
    // elimination: one barrier per pivot
    for (int j = 0; j < M; ++j) {
        __syncthreads();
        double piv = B[j * MS + j];
        if (!(piv > 0.0)) { if (tid == 0) s_fail = 1; piv = 1.0; }
        const double r = fast_rcp(piv);
        // this lane's slice of the pivot row (columns i = j+1+lane+32b) and of row j of X (columns c = lane+32b <= j)
        double pr[4], xr[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int i = j + 1 + lane + 32 * b, c = lane + 32 * b;
            pr[b] = (i < M) ? B[j * MS + i] : 0.0;
            xr[b] = (c <= j) ? X[j * MS + c] : 0.0;
        }
        for (int k = j + 1 + warp; k < M; k += NW) {
            const double t = B[j * MS + k] * r;            // Lt[k][j]
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int i = j + 1 + lane + 32 * b, c = lane + 32 * b;
                if (i >= k && i < M) B[k * MS + i] -= t * pr[b];
                if (c <= j) X[k * MS + c] -= t * xr[b];
            }
        }
    }
    __syncthreads();
    long long t2 = clock64();
    for (int i = tid; i < M; i += NT) {
        const double dpiv = B[i * MS + i];
        if (!(dpiv > 0.0)) s_fail = 1;
        s_rsq[i] = rsqrt(dpiv > 0.0 ? dpiv : 1.0);
    }
    __syncthreads();
    if (tid == 0 && s_fail) atomicExch(&acc->status, blockIdx.x + 1);
    // outputs (row-major global, coalesced): Lu = Lt D^1/2, Linv = D^-1/2 Lt^-1 (fp64, fp32, fp32 transposed)
    for (int idx = tid; idx < M * M; idx += NT) {
        const int i = idx / M, j = idx % M;
        double l = 0.0, x = 0.0;
        if (j <= i) {
            l = (i == j) ? 1.0 / s_rsq[j] : B[j * MS + i] * s_rsq[j];
            x = X[i * MS + j] * s_rsq[i];
        }
        P.Lu64[idx] = l; P.Linv64[idx] = x;
        P.Linv32[idx] = (float)x;
        // transposed copy, written coalesced: element (i, j) of LinvT is Linv[j][i]
        P.LinvT32[idx] = (i <= j) ? (float)(X[j * MS + i] * s_rsq[j]) : 0.f;
    }
    double s = 0.0;
    for (int i = tid; i < M; i += NT) s -= log(s_rsq[i]);          // sum log diag Lu = -sum log d^-1/2
    s = warp_sum_d(s);
    __shared__ double red[32];
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < NW; ++w) t += red[w];
        P.scal[0] = t; P.scal[2] = 0.0; P.scal[3] = 0.0;
        P.scal[5] = (double)(t1 - t0); P.scal[6] = (double)(t2 - t1); P.scal[7] = (double)(clock64() - t2);
    }
}


// ----------------------------------------------------------------------------------------------
// prepA, blocked variant of k_prepA_ldl (pivots four at a time) on ONE combined M x M shared-memory array:
//     C[k][c] = B[k][c] = A_k[c][k]   for c >= k   (transposed factor, row k = column k of the Schur complement)
//     C[k][c] = X[k][c] = Lt^-1[k][c] for c <  k   (unit lower inverse; its diagonal of ones is implicit)
// so that the update of row k by pivot row j is a single vector operation  C[k][c] -= t P_j[c]  over c <= j (inverse part,
// P_j[j] = 1) and c >= k (factor part): half the instructions and half the shared memory of keeping B and X apart.
// Per panel (rows j0..j0+3):
//   panel phase    -- three warps bring rows j0+1..j0+3 up to date, one pivot after the other (96-thread named barrier);
//   trailing phase -- every remaining row k gets the rank-4 update from the FINAL pivot rows, which each lane keeps in
//                     registers (one load + one store per four FMAs), two rows per warp pass.
// The arithmetic is that of the unblocked elimination with the four updates of an element summed in one pass.
// ----------------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(NT) k_prepA_c4(LayerSet ls, double jitter, Accum* acc) {
    const LayerDev& P = ls.l[blockIdx.x];
    const int M = P.M, Din = P.Din, tid = threadIdx.x;
    constexpr int MS = 129;                   // row stride: 4 blocks of 32 columns + 1 (M <= 128); columns >= M stay zero, so the
                                              // 32-wide column blocks need no bounds checks (a zero pivot-row entry is a no-op)
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int NW = NT / 32;
    static_assert(NW >= 3, "panel phase uses three warps");
    extern __shared__ double smd[];
    double* C = smd;                          // [M][MS]
    __shared__ double s_il[64];
    __shared__ double s_rsq[128];
    __shared__ int s_fail;
    const double var = (double)P.var[0];

    long long t0 = clock64();
    if (tid == 0) s_fail = 0;
    for (int q = tid; q < min(Din, 64); q += NT) s_il[q] = 1.0 / (double)P.ls[P.ard ? q : 0];
    for (int idx = tid; idx < M * MS; idx += NT) C[idx] = 0.0;
    __syncthreads();
    {   // Gram on i >= j: rows p and M-1-p folded into one line of M+1 entries so that all threads carry the same load
        const int H = (M + 1) / 2, W = M + 1;
        for (int idx = tid; idx < H * W; idx += NT) {
            const int p = idx / W, q = idx % W;
            int i, j;
            if (q <= p) { i = p; j = q; }
            else { i = M - 1 - p; j = q - p - 1; if (i == p) continue; }
            double r2 = 0.0;
            for (int qq = 0; qq < Din; ++qq) {
                const double il = qq < 64 ? s_il[qq] : 1.0 / (double)P.ls[P.ard ? qq : 0];
                const double d = ((double)P.Z[i * Din + qq] - (double)P.Z[j * Din + qq]) * il;
                r2 += d * d;
            }
            double k, kp;
            kern_eval_d(P.kern, r2, var, k, kp);
            if (i == j) k += jitter + (double)P.wvar[0];
            P.K64[i * M + j] = k;
            P.K64[j * M + i] = k;
            C[j * MS + i] = k;
        }
    }
    long long t1 = clock64();
    for (int j0 = 0; j0 < M; j0 += 4) {
        const int w = min(4, M - j0);
        __syncthreads();                                   // trailing updates of the previous panel are complete
        if (warp < 3) {
            const int rrow = j0 + 1 + warp;
            double* rr = C + rrow * MS + lane;
            for (int p = 0; p < 3; ++p) {                  // pivot j = j0 + p  (all three warps run every iteration: named barrier)
                const int j = j0 + p;
                if (p < w - 1 && rrow < j0 + w && rrow > j) {
                    const double* rj = C + j * MS + lane;
                    double piv = C[j * MS + j];
                    if (!(piv > 0.0)) piv = 1.0;           // (reported by the final pass over the pivots)
                    const double t = C[j * MS + rrow] * fast_rcp(piv);
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const int c = lane + 32 * b;
                        const double pj = (c == j) ? 1.0 : rj[32 * b];
                        if (c <= j || c >= rrow) rr[32 * b] -= t * pj;
                    }
                }
                asm volatile("bar.sync 1, 96;" ::: "memory");
            }
        }
        __syncthreads();                                   // the panel's pivot rows are final
        const int kb = j0 + w;                             // first trailing row
        if (kb >= M) continue;
        double rq[4], pr[4][4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int jq = j0 + q;
            rq[q] = 0.0;
#pragma unroll
            for (int b = 0; b < 4; ++b) pr[q][b] = 0.0;
            if (q < w) {
                double piv = C[jq * MS + jq];
                if (!(piv > 0.0)) piv = 1.0;
                rq[q] = fast_rcp(piv);
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int c = lane + 32 * b;
                    const double v = C[jq * MS + c];        // zero in the padding columns
                    pr[q][b] = (c == jq) ? 1.0 : ((c < jq || c >= kb) ? v : 0.0);
                }
            }
        }
        for (int k = kb + warp; k < M; k += 2 * NW) {      // two rows per pass: their loads are issued together
            const bool has2 = k + NW < M;
            const int k2 = has2 ? k + NW : k;
            double* rk = C + k * MS + lane;
            double* rk2 = C + k2 * MS + lane;
            double t[4], u[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                t[q] = 0.0; u[q] = 0.0;
                if (q < w) { t[q] = C[(j0 + q) * MS + k] * rq[q]; u[q] = C[(j0 + q) * MS + k2] * rq[q]; }
            }
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int c0 = 32 * b, c = c0 + lane;
                if (c0 >= kb && c0 + 31 < k) continue;       // block strictly between the two parts of row k (and of k2 > k)
                const bool a1 = c < kb || c >= k;
                const bool a2 = has2 && (c < kb || c >= k2);
                double v1 = rk[c0], v2 = rk2[c0];
                v1 -= t[0] * pr[0][b]; v1 -= t[1] * pr[1][b]; v1 -= t[2] * pr[2][b]; v1 -= t[3] * pr[3][b];
                v2 -= u[0] * pr[0][b]; v2 -= u[1] * pr[1][b]; v2 -= u[2] * pr[2][b]; v2 -= u[3] * pr[3][b];
                if (a1) rk[c0] = v1;
                if (a2) rk2[c0] = v2;
            }
        }
    }
    __syncthreads();
    long long t2 = clock64();
    // final pass over the pivots: any non-positive (or NaN) one means Kuu + jitter I is not positive definite
    for (int i = tid; i < M; i += NT) {
        const double dpiv = C[i * MS + i];
        if (!(dpiv > 0.0)) s_fail = 1;
        s_rsq[i] = rsqrt(dpiv > 0.0 ? dpiv : 1.0);
    }
    __syncthreads();
    if (tid == 0 && s_fail) atomicExch(&acc->status, blockIdx.x + 1);
    // outputs (row-major global, coalesced): Lu = Lt D^1/2, Linv = D^-1/2 Lt^-1 (fp64, fp32, fp32 transposed)
    for (int idx = tid; idx < M * M; idx += NT) {
        const int i = idx / M, j = idx % M;
        double l = 0.0, x = 0.0, xt = 0.0;
        if (j < i) { l = C[j * MS + i] * s_rsq[j]; x = C[i * MS + j] * s_rsq[i]; }
        else if (j == i) { l = 1.0 / s_rsq[i]; x = s_rsq[i]; xt = x; }
        else xt = C[j * MS + i] * s_rsq[j];                 // LinvT[i][j] = Linv[j][i], i < j
        P.Lu64[idx] = l; P.Linv64[idx] = x;
        P.Linv32[idx] = (float)x;
        P.LinvT32[idx] = (float)xt;
    }
    double s = 0.0;
    for (int i = tid; i < M; i += NT) s -= log(s_rsq[i]);          // sum log diag Lu = -sum log d^-1/2
    s = warp_sum_d(s);
    __shared__ double red[32];
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < NW; ++w) t += red[w];
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
    const size_t sm_ldl = 2 * (size_t)Mmax * (Mmax | 1) * sizeof(double);
    const size_t sm_c4 = (size_t)Mmax * 129 * sizeof(double);
    if (ls.prep_algo == 2 && Mmax <= 128) {
        if (ls.prep_threads == 256) k_prepA_c4<256><<<ls.L, 256, sm_c4, st>>>(ls, jitter, acc);
        else if (ls.prep_threads == 1024) k_prepA_c4<1024><<<ls.L, 1024, sm_c4, st>>>(ls, jitter, acc);
        else k_prepA_c4<512><<<ls.L, 512, sm_c4, st>>>(ls, jitter, acc);
    } else if (ls.prep_algo == 1 && Mmax <= 128 && sm_ldl <= 200 * 1024) {
        if (ls.prep_threads == 256) k_prepA_ldl<256><<<ls.L, 256, sm_ldl, st>>>(ls, jitter, acc);
        else if (ls.prep_threads == 1024) k_prepA_ldl<1024><<<ls.L, 1024, sm_ldl, st>>>(ls, jitter, acc);
        else k_prepA_ldl<512><<<ls.L, 512, sm_ldl, st>>>(ls, jitter, acc);
    } else
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
__global__ void k_fin_qsqrt(LayerSet ls, const StepArgs* sa, int lbase) {
    const LayerDev& P = ls.l[lbase + blockIdx.z];
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
__global__ void k_fin_qmu(LayerSet ls, const StepArgs* sa, int lbase) {
    const LayerDev& P = ls.l[lbase + blockIdx.y];
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
__global__ void k_fin_w1(LayerSet ls, int lbase) {
    const LayerDev& P = ls.l[lbase + blockIdx.y];
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
__global__ void k_fin_w2(LayerSet ls, int lbase) {
    const LayerDev& P = ls.l[lbase + blockIdx.y];
    if (!P.white) return;
    const int M = P.M;
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * M) return;
    int i = idx / M, j = idx % M;
    double s = 0.0;
    for (int k = j; k <= i; ++k) s += P.T1[i * M + k] * P.Linv64[k * M + j];
    P.Ssum64[idx] = s;
}
__global__ void k_fin_w3(LayerSet ls, int lbase) {
    const LayerDev& P = ls.l[lbase + blockIdx.y];
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
__global__ void k_fin_kbar(LayerSet ls, const StepArgs* sa, int lbase) {
    const LayerDev& P = ls.l[lbase + blockIdx.y];
    const int M = P.M, Din = P.Din;
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    double s2 = 0.0, sw = 0.0;
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
        if (i == j) sw = kb;               // d Kuu / d white-variance = I
    }
    s2 = warp_sum_d(s2); sw = warp_sum_d(sw);
    if ((threadIdx.x & 31) == 0) { atomicAdd(P.gvar, (float)s2); if (P.kwhite && sw != 0.0) atomicAdd(P.gwvar, (float)sw); }
}

// Zbar_iq += 4/l_q^2 sum_j g_ij (z_iq - z_jq) ; lsbar_q += -2/l_q^3 sum_ij g_ij (z_iq - z_jq)^2
__global__ void k_fin_kuu(LayerSet ls, int lbase) {
    const LayerDev& P = ls.l[lbase + blockIdx.y];
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


// ---- tiled / warp-per-output versions of the three long fin kernels (same arithmetic, fp64).  The one-thread-per-element
// forms above walk a dependent 100-iteration fp64 chain per thread out of L2 (57 + 17 + 17 us of a 1.2 ms step at the
// north-star shape, all of it on the tail of the critical path); these keep operands in shared memory / registers.
#define FT 32
// gq_sqrt[d] tile (ti >= tj): sum over k-tiles kt >= tj of A[ti,kt] L[kt,tj],  A = P_d + P_d^T - klw Kinv (white: no Kinv)
__global__ void __launch_bounds__(256) k_fin_qsqrt_t(LayerSet ls, const StepArgs* sa, int lbase) {
    const LayerDev& P = ls.l[lbase + blockIdx.z];
    const int M = P.M, d = blockIdx.y;
    if (d >= P.Dout) return;
    const int nt = (M + FT - 1) / FT;
    int ti = 0, rem = blockIdx.x;                  // blockIdx.x enumerates the lower-triangular tile pairs row by row
    while (rem > ti) { rem -= ti + 1; ++ti; }
    const int tj = rem;
    if (ti >= nt) return;
    __shared__ double As[FT][FT + 1], Ls[FT][FT + 1];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const double klw = sa->kl_weight;
    const float* Ld = P.q_sqrt + (size_t)d * M * M;
    const float* Pd = P.Pd + (size_t)d * M * M;
    const int i0 = ti * FT, j0 = tj * FT;
    double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    for (int kt = tj; kt < nt; ++kt) {
        const int k0 = kt * FT;
        for (int e = threadIdx.x; e < FT * FT; e += 256) {
            const int r = e / FT, c = e % FT;
            const int i = i0 + r, k = k0 + c;                 // A[i][k]
            double a = 0.0;
            if (i < M && k < M) {
                a = (double)Pd[i * M + k] + (double)Pd[k * M + i];
                if (!P.white) a -= klw * P.Kinv64[i * M + k];
            }
            As[r][c] = a;
            const int kk = k0 + r, j = j0 + c;                // L[kk][j], lower triangular
            Ls[r][c] = (kk < M && j < M && j <= kk) ? (double)Ld[kk * M + j] : 0.0;
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < FT; ++kk) {
            const double a0 = As[ty][kk], a1 = As[ty + 16][kk], l0 = Ls[kk][tx], l1 = Ls[kk][tx + 16];
            acc[0][0] += a0 * l0; acc[0][1] += a0 * l1; acc[1][0] += a1 * l0; acc[1][1] += a1 * l1;
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int i = i0 + ty + 16 * a, j = j0 + tx + 16 * b;
            if (i >= M || j >= M) continue;
            double g = 0.0;
            if (j <= i) {
                g = acc[a][b];
                if (P.white) g -= klw * (double)Ld[i * M + j];
                if (i == j) g += klw / (double)Ld[i * M + i];
            }
            P.gq_sqrt[(size_t)d * M * M + i * M + j] = (float)g;      // (strictly-upper tiles stay at the step's memset zero)
        }
}

// gq_mu: one warp per output (i, d), lanes over k
__global__ void __launch_bounds__(256) k_fin_qmu_w(LayerSet ls, const StepArgs* sa, int lbase) {
    const LayerDev& P = ls.l[lbase + blockIdx.y];
    const int M = P.M, D = P.Dout;
    const int idx = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (idx >= M * D) return;
    const int i = idx / D, d = idx % D;
    const double klw = sa->kl_weight;
    double s = 0.0;
    if (!P.white)
        for (int k = lane; k < M; k += 32) s += P.Kinv64[i * M + k] * (double)P.q_mu[k * D + d];
    s = warp_sum_d(s);
    if (lane == 0) {
        double g = (double)P.qmubar[idx];
        g -= P.white ? klw * (double)P.q_mu[idx] : klw * s;
        P.gq_mu[idx] = (float)g;
    }
}

// Zbar / lsbar from Gsym: one warp per inducing point i, lanes over j, input dimensions in chunks of 8
__global__ void __launch_bounds__(256) k_fin_kuu_w(LayerSet ls, int lbase) {
    const LayerDev& P = ls.l[lbase + blockIdx.y];
    const int M = P.M, Din = P.Din;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= M) return;
    double ls_acc = 0.0;                                   // non-ARD: all dimensions share one lengthscale
    for (int q0 = 0; q0 < Din; q0 += 8) {
        const int nq = min(8, Din - q0);
        double a[8], b[8], zi[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) { a[q] = 0.0; b[q] = 0.0; zi[q] = q < nq ? (double)P.Z[i * Din + q0 + q] : 0.0; }
        for (int j = lane; j < M; j += 32) {
            const double g = P.Gsym[i * M + j];
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (q < nq) {
                    const double dd = zi[q] - (double)P.Z[j * Din + q0 + q];
                    a[q] += g * dd; b[q] += g * dd * dd;
                }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const double as = warp_sum_d(a[q]), bs = warp_sum_d(b[q]);
            if (lane == 0 && q < nq) {
                const double l = (double)P.ls[P.ard ? q0 + q : 0];
                atomicAdd(&P.gZ[i * Din + q0 + q], (float)(4.0 * as / (l * l)));
                if (P.ard) atomicAdd(&P.gls[q0 + q], (float)(-2.0 * bs / (l * l * l)));
                else ls_acc += -2.0 * bs / (l * l * l);
            }
        }
    }
    if (lane == 0 && !P.ard) atomicAdd(&P.gls[0], (float)ls_acc);
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

// layers [l0, l1): everything here depends only on that layer's accumulators and on the KL preparation, so the step DAG
// runs it per layer on the side branch right behind the layer's row reductions (api.cu)
// part 0 = everything; the two halves are independent of each other and may run on different streams:
// part 1 = variational parameters (q_sqrt, q_mu), part 2 = kernel side (whitened chain, Kuu-bar, Z / hyper-parameters)
void launch_fin(const LayerSet& ls, int l0, int l1, Accum* acc, const StepArgs* sa, cudaStream_t st, long long* nl, int part) {
    int Mmax = 0, Dmax = 0, MDmax = 0, MDin = 0;
    bool any_white = false;
    const int nL = l1 - l0;
    for (int l = l0; l < l1; ++l) {
        Mmax = max(Mmax, ls.l[l].M); Dmax = max(Dmax, ls.l[l].Dout);
        MDmax = max(MDmax, ls.l[l].M * ls.l[l].Dout); MDin = max(MDin, ls.l[l].M * ls.l[l].Din);
        any_white |= ls.l[l].white != 0;
    }
    int nb = (Mmax * Mmax + 255) / 256;
    if (part != 2) {
        if (ls.fin_algo == 1) {
            const int nt = (Mmax + FT - 1) / FT;
            k_fin_qsqrt_t<<<dim3(nt * (nt + 1) / 2, Dmax, nL), 256, 0, st>>>(ls, sa, l0);
            k_fin_qmu_w<<<dim3((MDmax + 7) / 8, nL), 256, 0, st>>>(ls, sa, l0);
        } else {
            k_fin_qsqrt<<<dim3(nb, Dmax, nL), 256, 0, st>>>(ls, sa, l0);
            k_fin_qmu<<<dim3((MDmax + 255) / 256, nL), 256, 0, st>>>(ls, sa, l0);
        }
        *nl += 2;
    }
    if (part == 1) return;
    if (any_white) {
        k_fin_w1<<<dim3(nb, nL), 256, 0, st>>>(ls, l0);
        k_fin_w2<<<dim3(nb, nL), 256, 0, st>>>(ls, l0);
        k_fin_w3<<<dim3(nb, nL), 256, 0, st>>>(ls, l0);
        *nl += 3;
    }
    k_fin_kbar<<<dim3(nb, nL), 256, 0, st>>>(ls, sa, l0);
    if (ls.fin_algo == 1) k_fin_kuu_w<<<dim3((Mmax + 7) / 8, nL), 256, 0, st>>>(ls, l0);
    else k_fin_kuu<<<dim3((MDin + 255) / 256, nL), 256, 0, st>>>(ls, l0);
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
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(k_prepA_c4<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024))) return e;
    if ((e = cudaFuncSetAttribute(k_prepA_c4<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024))) return e;
    if ((e = cudaFuncSetAttribute(k_prepA_c4<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024))) return e;
    if ((e = cudaFuncSetAttribute(k_prepA_ldl<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024))) return e;
    if ((e = cudaFuncSetAttribute(k_prepA_ldl<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024))) return e;
    if ((e = cudaFuncSetAttribute(k_prepA_ldl<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024))) return e;
    return cudaFuncSetAttribute(k_prepA, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
}
