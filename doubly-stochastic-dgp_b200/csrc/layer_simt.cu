// fp32 SIMT row-tile kernels of one SVGP layer: forward (conditional + reparameterised draw), backward over rows,
// and the row-reduction GEMMs that produce the parameter-gradient accumulators.
// Reference: layers.py:178-219 (conditional_ND), layers.py:52-119 (conditional_SND / sample_from_conditional),
// utils.py:37-41 (reparameterize); backward = what TF autodiff derives.  The exact arithmetic
// (whitened projections, no D_out-tiled temporaries) is mirrored in tests/algo_mirror.py.
//
// Shared-memory activations are kept feature-major:  XT[feature][row]  (row contiguous) so that a thread's
// register tile (RT rows x 4 columns) reads its rows with one vector load per k.
#include "dsdgp_internal.cuh"

#define KB DSDGP_KB
#define NCH DSDGP_NCH
#define QC 32                 // input-dimension chunk of the Gram stage

// OUT(r, n) = sum_k AT[k][r] * W[k*ldw + n]      kmode 0: all k; 1: k <= n (chunk-level); 2: k >= n (chunk-level)
// AT: shared [Kpad][TR+4], zero padded to a multiple of KB rows.  W: global, row-major (n contiguous).
// epi(c0, acc) is called once per 64-column chunk; thread owns rows ty*RT+rr, columns c0+tx*4+cc.
template <int TR, class Epi>
__device__ __forceinline__ void tile_gemm(const float* __restrict__ AT, int K, const float* __restrict__ W, int ldw,
                                          int NC, int kmode, float* __restrict__ Ws, Epi epi) {
    constexpr int RT = TR / 16, TRS = TR + 4;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    for (int c0 = 0; c0 < NC; c0 += NCH) {
        const int k_begin = (kmode == 2) ? c0 : 0;
        const int k_end = (kmode == 1) ? min(K, c0 + NCH) : K;
        float acc[RT][4];
#pragma unroll
        for (int rr = 0; rr < RT; ++rr)
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) acc[rr][cc] = 0.f;
        for (int k0 = k_begin; k0 < k_end; k0 += KB) {
#pragma unroll
            for (int e = tid; e < KB * NCH; e += DSDGP_NT) {
                int kk = e >> 6, cc = e & 63, k = k0 + kk, n = c0 + cc;
                Ws[e] = (k < k_end && n < NC) ? __ldg(&W[(size_t)k * ldw + n]) : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < KB; ++kk) {
                const float4 w = *reinterpret_cast<const float4*>(&Ws[kk * NCH + tx * 4]);
                const float* ap = &AT[(k0 + kk) * TRS + ty * RT];
                float a[RT];
                if constexpr (RT == 4) { float4 t = *reinterpret_cast<const float4*>(ap); a[0] = t.x; a[1] = t.y; a[2] = t.z; a[3] = t.w; }
                else if constexpr (RT == 2) { float2 t = *reinterpret_cast<const float2*>(ap); a[0] = t.x; a[1] = t.y; }
                else { a[0] = ap[0]; }
#pragma unroll
                for (int rr = 0; rr < RT; ++rr) {
                    acc[rr][0] = fmaf(a[rr], w.x, acc[rr][0]);
                    acc[rr][1] = fmaf(a[rr], w.y, acc[rr][1]);
                    acc[rr][2] = fmaf(a[rr], w.z, acc[rr][2]);
                    acc[rr][3] = fmaf(a[rr], w.w, acc[rr][3]);
                }
            }
            __syncthreads();
        }
        epi(c0, acc);
    }
}

__host__ __device__ inline int pad16(int x) { return (x + 15) & ~15; }

size_t fwd_smem_bytes(int M, int Dout, int TR) {
    size_t Mp = pad16(M), TRS = TR + 4;
    return sizeof(float) * (3 * Mp * TRS + KB * NCH + QC * TRS + QC + 2 * (size_t)TR * Dout + TR);
}
size_t bwd_smem_bytes(int M, int Dout, int TR) {
    size_t Mp = pad16(M), TRS = TR + 4;
    return sizeof(float) * (5 * Mp * TRS + KB * NCH + QC * TRS + 2 * QC + 2 * (size_t)TR * Dout + TR + 16);
}

// Gram stage: kT[i][r] = k(z_i, x_r) (and optionally kpT = dk/dr2), x rows row0.. from Xin.
template <int TR, bool WITH_KP>
__device__ __forceinline__ void gram_stage(const LayerDev& P, const float* __restrict__ Xin, int row0, int R,
                                           float* kT, float* kpT, float* xT, float* ilc) {
    constexpr int TRS = TR + 4;
    const int M = P.M, Din = P.Din, tid = threadIdx.x, Mp = pad16(M);
    for (int e = tid; e < Mp * TRS; e += DSDGP_NT) kT[e] = 0.f;
    for (int q0 = 0; q0 < Din; q0 += QC) {
        const int qn = min(QC, Din - q0);
        __syncthreads();
        for (int e = tid; e < TR * qn; e += DSDGP_NT) {
            int r = e / qn, qq = e % qn, row = row0 + r;
            xT[qq * TRS + r] = (row < R) ? Xin[(size_t)row * Din + q0 + qq] : 0.f;
        }
        if (tid < qn) ilc[tid] = 1.0f / P.ls[P.ard ? q0 + tid : 0];
        __syncthreads();
        for (int e = tid; e < M * TR; e += DSDGP_NT) {
            int r = e % TR, i = e / TR;
            const float* z = P.Z + (size_t)i * Din + q0;
            float s = 0.f;
            for (int qq = 0; qq < qn; ++qq) {
                float d = (xT[qq * TRS + r] - __ldg(&z[qq])) * ilc[qq];
                s = fmaf(d, d, s);
            }
            kT[i * TRS + r] += s;
        }
    }
    __syncthreads();
    const float var = P.var[0];
    for (int e = tid; e < M * TR; e += DSDGP_NT) {
        int r = e % TR, i = e / TR;
        float k, kp;
        kern_eval_f(P.kern, kT[i * TRS + r], var, k, kp);
        kT[i * TRS + r] = k;
        if (WITH_KP) kpT[i * TRS + r] = kp;
    }
    __syncthreads();
}

__device__ __forceinline__ float half_warp_sum(float v) {      // over the 16 lanes sharing ty
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

// ----------------------------------------------------------------------------------------------
// forward
// ----------------------------------------------------------------------------------------------
template <int TR>
__global__ void __launch_bounds__(DSDGP_NT) k_layer_fwd(LayerDev P, FwdArgs a) {
    constexpr int RT = TR / 16, TRS = TR + 4;
    const int M = P.M, Din = P.Din, D = P.Dout, Mp = pad16(M);
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int row0 = blockIdx.x * TR, R = a.R;
    extern __shared__ __align__(16) float sm[];
    float* kT = sm;
    float* bT = kT + Mp * TRS;
    float* uT = bT + Mp * TRS;
    float* Ws = uT + Mp * TRS;
    float* xT = Ws + KB * NCH;
    float* ilc = xT + QC * TRS;
    float* macc = ilc + QC;            // [TR][D]
    float* vacc = macc + TR * D;       // [TR][D]
    float* bn = vacc + TR * D;         // [TR]

    for (int e = tid; e < 2 * Mp * TRS; e += DSDGP_NT) bT[e] = 0.f;     // bT and uT (padding rows must be 0)
    gram_stage<TR, false>(P, a.Xin, row0, R, kT, nullptr, xT, ilc);

    // b = Linv k
    tile_gemm<TR>(kT, M, P.LinvT32, M, M, 1, Ws, [&](int c0, float (&acc)[RT][4]) {
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            int n = c0 + tx * 4 + cc;
            if (n < M) {
#pragma unroll
                for (int rr = 0; rr < RT; ++rr) bT[n * TRS + ty * RT + rr] = acc[rr][cc];
            }
        }
    });
    __syncthreads();
    if (tid < TR) {
        float s = 0.f;
        for (int j = 0; j < M; ++j) { float b = bT[j * TRS + tid]; s = fmaf(b, b, s); }
        bn[tid] = s;
    }
    const float* uTp = bT;
    if (!P.white) {
        // u = Linv^T b
        tile_gemm<TR>(bT, M, P.Linv32, M, M, 2, Ws, [&](int c0, float (&acc)[RT][4]) {
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                int n = c0 + tx * 4 + cc;
                if (n < M) {
#pragma unroll
                    for (int rr = 0; rr < RT; ++rr) uT[n * TRS + ty * RT + rr] = acc[rr][cc];
                }
            }
        });
        uTp = uT;
    }
    __syncthreads();
    // U out (row-major, coalesced along i) and mean = u . q_mu
    for (int e = tid; e < TR * M; e += DSDGP_NT) {
        int i = e % M, r = e / M, row = row0 + r;
        if (row < R) a.U[(size_t)row * M + i] = uTp[i * TRS + r];
    }
    for (int e = tid; e < TR * D; e += DSDGP_NT) {
        int r = e % TR, d = e / TR;
        float s = 0.f;
        for (int i = 0; i < M; ++i) s = fmaf(uTp[i * TRS + r], __ldg(&P.q_mu[i * D + d]), s);
        macc[r * D + d] = s;
    }
    // c_d = L_d^T u ; |c_d|^2
    for (int d = 0; d < D; ++d) {
        float ss[RT];
#pragma unroll
        for (int rr = 0; rr < RT; ++rr) ss[rr] = 0.f;
        tile_gemm<TR>(uTp, M, P.q_sqrt + (size_t)d * M * M, M, M, 2, Ws, [&](int c0, float (&acc)[RT][4]) {
#pragma unroll
            for (int rr = 0; rr < RT; ++rr)
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) ss[rr] = fmaf(acc[rr][cc], acc[rr][cc], ss[rr]);
        });
#pragma unroll
        for (int rr = 0; rr < RT; ++rr) {
            float t = half_warp_sum(ss[rr]);
            if (tx == 0) vacc[(ty * RT + rr) * D + d] = t;
        }
    }
    __syncthreads();
    // mean function, variance, draw
    const float var0 = P.var[0] + P.wvar[0], jit = a.jitter;      // Kdiag of Sum(kernel, White)
    const unsigned long long seed = a.sa->seed;
    const int noff = a.sa->n_offset, soff = a.sa->s_offset;
    const int ipd = P.ipd, FS = ipd + D;                          // F rows: [X[:ipd] | samples] (layers.py:105-117)
    if (a.F && ipd > 0) {
        for (int e = tid; e < TR * ipd; e += DSDGP_NT) {
            const int q = e % ipd, r = e / ipd, row = row0 + r;
            if (row >= R) continue;
            const float xv = a.Xin[(size_t)row * Din + q];
            if (a.S_rep == 1) a.F[(size_t)row * FS + q] = xv;
            else for (int s = 0; s < a.S_rep; ++s) a.F[((size_t)s * a.N + row) * FS + q] = xv;
        }
    }
    for (int e = tid; e < TR * D; e += DSDGP_NT) {
        int d = e % D, r = e / D, row = row0 + r;
        if (row >= R) continue;
        float mean = macc[r * D + d];
        if (P.mean == DSDGP_MEAN_IDENTITY) mean += a.Xin[(size_t)row * Din + d];
        else if (P.mean == DSDGP_MEAN_LINEAR) {
            float s = P.meanB[d];
            for (int q = 0; q < Din; ++q) s = fmaf(a.Xin[(size_t)row * Din + q], __ldg(&P.meanW[q * D + d]), s);
            mean += s;
        }
        float v = var0 - bn[r] + vacc[r * D + d];
        a.Fmean[(size_t)row * D + d] = mean;
        a.Fvar[(size_t)row * D + d] = v;
        if (a.F) {
            float sd = sqrtf(fmaxf(v + jit, 1e-30f));
            if (a.S_rep == 1) {
                int s = row / a.N, n = row % a.N;
                float z = a.z ? a.z[(size_t)row * D + d] : dsdgp_normal(seed, P.idx, s + soff, n + noff, d);
                if (a.z_out) a.z_out[(size_t)row * D + d] = z;
                a.F[(size_t)row * FS + ipd + d] = fmaf(z, sd, mean);
            } else {
                for (int s = 0; s < a.S_rep; ++s) {
                    size_t o = ((size_t)s * a.N + row) * D + d;
                    float z = a.z ? a.z[o] : dsdgp_normal(seed, P.idx, s + soff, row + noff, d);
                    if (a.z_out) a.z_out[o] = z;
                    a.F[((size_t)s * a.N + row) * FS + ipd + d] = fmaf(z, sd, mean);
                }
            }
        }
    }
}

// ----------------------------------------------------------------------------------------------
// backward over rows
// ----------------------------------------------------------------------------------------------
template <int TR>
__global__ void __launch_bounds__(DSDGP_NT) k_layer_bwd(LayerDev P, BwdArgs a) {
    constexpr int RT = TR / 16, TRS = TR + 4;
    const int M = P.M, Din = P.Din, D = P.Dout, Mp = pad16(M);
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int row0 = blockIdx.x * TR, R = a.R;
    extern __shared__ __align__(16) float sm[];
    float* uT = sm;
    float* kT = uT + Mp * TRS;
    float* kpT = kT + Mp * TRS;
    float* ubT = kpT + Mp * TRS;
    float* cT = ubT + Mp * TRS;
    float* Ws = cT + Mp * TRS;
    float* xT = Ws + KB * NCH;
    float* ilc = xT + QC * TRS;
    float* lsacc = ilc + QC;
    float* mub = lsacc + QC;           // [TR][D]
    float* vb = mub + TR * D;          // [TR][D]
    float* vs = vb + TR * D;           // [TR]
    float* red = vs + TR;              // [16]
    const int ipd = P.ipd, FS = ipd + D;       // upstream gradient rows: [d/dX[:ipd] | d/dF] (layers.py:105-117)

    const float jit = a.jitter;
    const unsigned long long seed = a.sa->seed;
    const int noff = a.sa->n_offset, soff = a.sa->s_offset;

    // B1: mubar, vbar
    for (int e = tid; e < TR * D; e += DSDGP_NT) {
        int d = e % D, r = e / D, row = row0 + r;
        float m = 0.f, v = 0.f;
        if (row < R) {
            if (a.fbar) {
                float sd = sqrtf(fmaxf(a.Fvar[(size_t)row * D + d] + jit, 1e-30f));
                if (a.S_rep == 1) {
                    int s = row / a.N, n = row % a.N;
                    float fb = a.fbar[(size_t)row * FS + ipd + d];
                    float z = a.z ? a.z[(size_t)row * D + d] : dsdgp_normal(seed, P.idx, s + soff, n + noff, d);
                    m = fb; v = fb * z / (2.f * sd);
                } else {
                    float sz = 0.f;
                    for (int s = 0; s < a.S_rep; ++s) {
                        size_t o = ((size_t)s * a.N + row) * D + d;
                        float fb = a.fbar[((size_t)s * a.N + row) * FS + ipd + d];
                        float z = a.z ? a.z[o] : dsdgp_normal(seed, P.idx, s + soff, row + noff, d);
                        m += fb; sz = fmaf(fb, z, sz);
                    }
                    v = sz / (2.f * sd);
                }
                a.mubar[(size_t)row * D + d] = m;
                a.vbar[(size_t)row * D + d] = v;
            } else {
                m = a.mubar[(size_t)row * D + d];
                v = a.vbar[(size_t)row * D + d];
            }
        }
        mub[r * D + d] = m; vb[r * D + d] = v;
    }
    // B2: u tile (feature-major), zero padding rows of all [Mp][TRS] buffers that feed GEMMs
    for (int e = tid; e < Mp * TRS; e += DSDGP_NT) { uT[e] = 0.f; ubT[e] = 0.f; cT[e] = 0.f; }
    __syncthreads();
    for (int e = tid; e < TR * M; e += DSDGP_NT) {
        int i = e % M, r = e / M, row = row0 + r;
        uT[i * TRS + r] = (row < R) ? a.U[(size_t)row * M + i] : 0.f;
    }
    if (tid < TR) {
        float s = 0.f;
        for (int d = 0; d < D; ++d) s += vb[tid * D + d];
        vs[tid] = s;
    }
    // B3: k, dk/dr2
    gram_stage<TR, true>(P, a.Xin, row0, R, kT, kpT, xT, ilc);
    // B4: ubar init
    for (int e = tid; e < M * TR; e += DSDGP_NT) {
        int r = e % TR, i = e / TR;
        float s = 0.f;
        for (int d = 0; d < D; ++d) s = fmaf(mub[r * D + d], __ldg(&P.q_mu[i * D + d]), s);
        s -= P.white ? 2.f * vs[r] * uT[i * TRS + r] : vs[r] * kT[i * TRS + r];
        ubT[i * TRS + r] = s;
    }
    __syncthreads();
    // B5: ubar += sum_d L_d (2 vbar_d c_d),  c_d = L_d^T u
    for (int d = 0; d < D; ++d) {
        tile_gemm<TR>(uT, M, P.q_sqrt + (size_t)d * M * M, M, M, 2, Ws, [&](int c0, float (&acc)[RT][4]) {
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                int n = c0 + tx * 4 + cc;
                if (n < M) {
#pragma unroll
                    for (int rr = 0; rr < RT; ++rr) {
                        int r = ty * RT + rr;
                        cT[n * TRS + r] = 2.f * vb[r * D + d] * acc[rr][cc];
                    }
                }
            }
        });
        __syncthreads();
        tile_gemm<TR>(cT, M, P.q_sqrtT + (size_t)d * M * M, M, M, 1, Ws, [&](int c0, float (&acc)[RT][4]) {
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                int n = c0 + tx * 4 + cc;
                if (n < M) {
#pragma unroll
                    for (int rr = 0; rr < RT; ++rr) ubT[n * TRS + ty * RT + rr] += acc[rr][cc];
                }
            }
        });
        __syncthreads();
    }
    // B6: w (row-major to global, straight from the register tile) and kbar
    float* kbT;
    auto store_w = [&](int n, int r, float w) {
        int row = row0 + r;
        if (row < R) a.W[(size_t)row * M + n] = w;
    };
    if (!P.white) {
        // t = Linv ubar ; w = Linv^T t = K^-1 ubar ; kbar = w - vs u
        tile_gemm<TR>(ubT, M, P.LinvT32, M, M, 1, Ws, [&](int c0, float (&acc)[RT][4]) {
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                int n = c0 + tx * 4 + cc;
                if (n < M) {
#pragma unroll
                    for (int rr = 0; rr < RT; ++rr) cT[n * TRS + ty * RT + rr] = acc[rr][cc];
                }
            }
        });
        __syncthreads();
        kbT = ubT;
        tile_gemm<TR>(cT, M, P.Linv32, M, M, 2, Ws, [&](int c0, float (&acc)[RT][4]) {
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                int n = c0 + tx * 4 + cc;
                if (n < M) {
#pragma unroll
                    for (int rr = 0; rr < RT; ++rr) {
                        int r = ty * RT + rr;
                        float w = acc[rr][cc];
                        store_w(n, r, w);
                        kbT[n * TRS + r] = w - vs[r] * uT[n * TRS + r];
                    }
                }
            }
        });
    } else {
        // kbar = Linv^T bbar ; w = kbar
        kbT = cT;
        tile_gemm<TR>(ubT, M, P.Linv32, M, M, 2, Ws, [&](int c0, float (&acc)[RT][4]) {
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                int n = c0 + tx * 4 + cc;
                if (n < M) {
#pragma unroll
                    for (int rr = 0; rr < RT; ++rr) {
                        int r = ty * RT + rr;
                        store_w(n, r, acc[rr][cc]);
                        kbT[n * TRS + r] = acc[rr][cc];
                    }
                }
            }
        });
    }
    __syncthreads();
    // B7: g = 2 kbar kp (into kpT) ; s2 partial
    float s2 = 0.f;
    const float inv_var = 1.0f / P.var[0];
    for (int e = tid; e < M * TR; e += DSDGP_NT) {
        int r = e % TR, i = e / TR;
        float kb = kbT[i * TRS + r];
        s2 = fmaf(kb * kT[i * TRS + r], inv_var, s2);
        kpT[i * TRS + r] = 2.f * kb * kpT[i * TRS + r];
    }
    float sw = 0.f;                                   // sum of vbar: d Kdiag / d variance = d Kdiag / d white-variance = 1
    for (int e = tid; e < TR * D; e += DSDGP_NT) sw += vb[e];
    s2 += sw;
    s2 = warp_sum(s2); sw = warp_sum(sw);
    if ((tid & 31) == 0) { red[tid >> 5] = s2; red[8 + (tid >> 5)] = sw; }
    __syncthreads();
    if (tid == 0) {
        float t = 0.f, tw = 0.f;
        for (int w = 0; w < DSDGP_NT / 32; ++w) { t += red[w]; tw += red[8 + w]; }
        atomicAdd(P.gvar, t);
        if (P.kwhite) atomicAdd(P.gwvar, tw);
    }
    // B8: xbar, Zbar, lsbar (chunked over the input dimension)
    const float* gT = kpT;
    for (int q0 = 0; q0 < Din; q0 += QC) {
        const int qn = min(QC, Din - q0);
        __syncthreads();
        for (int e = tid; e < TR * qn; e += DSDGP_NT) {
            int r = e / qn, qq = e % qn, row = row0 + r;
            xT[qq * TRS + r] = (row < R) ? a.Xin[(size_t)row * Din + q0 + qq] : 0.f;
        }
        if (tid < qn) { ilc[tid] = 1.0f / P.ls[P.ard ? q0 + tid : 0]; lsacc[tid] = 0.f; }
        __syncthreads();
        if (a.xbar) {
            for (int e = tid; e < TR * qn; e += DSDGP_NT) {
                int r = e % TR, qq = e / TR, row = row0 + r;
                float x = xT[qq * TRS + r], s = 0.f;
                for (int i = 0; i < M; ++i) s = fmaf(gT[i * TRS + r], x - __ldg(&P.Z[(size_t)i * Din + q0 + qq]), s);
                s *= ilc[qq] * ilc[qq];
                int q = q0 + qq;
                if (P.mean == DSDGP_MEAN_IDENTITY) s += mub[r * D + q];
                else if (P.mean == DSDGP_MEAN_LINEAR) {
                    for (int d = 0; d < D; ++d) s = fmaf(mub[r * D + d], __ldg(&P.meanW[q * D + d]), s);
                }
                // input propagation: the first ipd columns of this layer's output ARE its first ipd inputs
                if (row < R && a.fbar && q < ipd) s += a.fbar[(size_t)row * FS + q];
                if (row < R) a.xbar[(size_t)row * Din + q] = s;
            }
        }
        for (int e = tid; e < M * qn; e += DSDGP_NT) {
            int qq = e % qn, i = e / qn;
            float z = P.Z[(size_t)i * Din + q0 + qq], sa = 0.f, sb = 0.f;
            for (int r = 0; r < TR; ++r) {
                float d = xT[qq * TRS + r] - z, g = gT[i * TRS + r];
                sa = fmaf(g, d, sa); sb = fmaf(g * d, d, sb);
            }
            float il = ilc[qq];
            atomicAdd(&P.gZ[(size_t)i * Din + q0 + qq], -sa * il * il);
            atomicAdd(&lsacc[qq], -sb * il * il * il);
        }
        __syncthreads();
        if (tid < 32) {        // warp 0; qn <= QC == 32
            float t = (tid < qn) ? lsacc[tid] : 0.f;
            if (P.ard) { if (tid < qn) atomicAdd(&P.gls[q0 + tid], t); }
            else { t = warp_sum(t); if (tid == 0) atomicAdd(&P.gls[0], t); }
        }
    }
}

// ----------------------------------------------------------------------------------------------
// row-reduction GEMMs:  P_d = sum_r vbar_rd u_r u_r^T ;  G = sum_r w_r u_r^T ; qmubar = sum_r u_r mubar_r^T
// grid: x = row split, y = which (0..D-1: P_d, D: G, D+1: qmubar), z = (i tile of 64) * ncolblk + (column block of 128)
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DSDGP_NT) k_layer_rowred(LayerDev P, const float* __restrict__ U, const float* __restrict__ W,
                                                           const float* __restrict__ mubar, const float* __restrict__ vbar,
                                                           int R, int rows_per_split, int ncolblk) {
    const int M = P.M, D = P.Dout;
    const int which = blockIdx.y;
    const int itile = blockIdx.z / ncolblk, cblk = blockIdx.z % ncolblk;
    const int i0 = itile * 64, j0 = cblk * 128;
    const int NC = (which == D + 1) ? D : M;
    if (j0 >= NC) return;
    const int r_lo = blockIdx.x * rows_per_split, r_hi = min(R, r_lo + rows_per_split);
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    __shared__ __align__(16) float As[KB][64 + 4];
    __shared__ __align__(16) float Bs[KB][128];
    const float* Asrc = (which == D) ? W : U;
    float acc[2][4][4];
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int rr = 0; rr < 4; ++rr)
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) acc[c][rr][cc] = 0.f;
    for (int r0 = r_lo; r0 < r_hi; r0 += KB) {
        for (int e = tid; e < KB * 64; e += DSDGP_NT) {
            int kk = e >> 6, ii = e & 63, r = r0 + kk, i = i0 + ii;
            As[kk][ii] = (r < r_hi && i < M) ? Asrc[(size_t)r * M + i] : 0.f;
        }
        for (int e = tid; e < KB * 128; e += DSDGP_NT) {
            int kk = e >> 7, jj = e & 127, r = r0 + kk, j = j0 + jj;
            float v = 0.f;
            if (r < r_hi && j < NC) {
                if (which < D) v = vbar[(size_t)r * D + which] * U[(size_t)r * M + j];
                else if (which == D) v = U[(size_t)r * M + j];
                else v = mubar[(size_t)r * D + j];
            }
            Bs[kk][jj] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < KB; ++kk) {
            float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            float a4[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                float4 w = *reinterpret_cast<const float4*>(&Bs[kk][c * 64 + tx * 4]);
#pragma unroll
                for (int rr = 0; rr < 4; ++rr) {
                    acc[c][rr][0] = fmaf(a4[rr], w.x, acc[c][rr][0]);
                    acc[c][rr][1] = fmaf(a4[rr], w.y, acc[c][rr][1]);
                    acc[c][rr][2] = fmaf(a4[rr], w.z, acc[c][rr][2]);
                    acc[c][rr][3] = fmaf(a4[rr], w.w, acc[c][rr][3]);
                }
            }
        }
        __syncthreads();
    }
    float* out; int ldo;
    if (which < D) { out = P.Pd + (size_t)which * M * M; ldo = M; }
    else if (which == D) { out = P.G; ldo = M; }
    else { out = P.qmubar; ldo = D; }
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int rr = 0; rr < 4; ++rr)
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                int i = i0 + ty * 4 + rr, j = j0 + c * 64 + tx * 4 + cc;
                if (i < M && j < NC) atomicAdd(&out[(size_t)i * ldo + j], acc[c][rr][cc]);
            }
}

// ----------------------------------------------------------------------------------------------
// launchers
// ----------------------------------------------------------------------------------------------
static int pick_tr(int R, int num_sms, size_t (*bytes)(int, int, int), int M, int D) {
    // largest tile that still gives every SM work and fits in shared memory
    const int cand[3] = {64, 32, 16};
    for (int c = 0; c < 3; ++c) {
        int tr = cand[c];
        if (bytes(M, D, tr) > 220 * 1024) continue;
        if ((R + tr - 1) / tr >= num_sms || tr == 16) return tr;
    }
    return 16;
}

template <int TR>
static void fwd_t(const LayerDev& P, const FwdArgs& a, cudaStream_t st) {
    size_t sm = fwd_smem_bytes(P.M, P.Dout, TR);
    k_layer_fwd<TR><<<(a.R + TR - 1) / TR, DSDGP_NT, sm, st>>>(P, a);
}
template <int TR>
static void bwd_t(const LayerDev& P, const BwdArgs& a, cudaStream_t st) {
    size_t sm = bwd_smem_bytes(P.M, P.Dout, TR);
    k_layer_bwd<TR><<<(a.R + TR - 1) / TR, DSDGP_NT, sm, st>>>(P, a);
}

void launch_fwd(const LayerDev& P, const FwdArgs& a, int num_sms, cudaStream_t st, long long* nl) {
    int tr = pick_tr(a.R, num_sms, fwd_smem_bytes, P.M, P.Dout);
    if (tr == 64) fwd_t<64>(P, a, st); else if (tr == 32) fwd_t<32>(P, a, st); else fwd_t<16>(P, a, st);
    *nl += 1;
}

void launch_bwd_rows(const LayerDev& P, const BwdArgs& a, int num_sms, cudaStream_t st, long long* nl) {
    int tr = pick_tr(a.R, num_sms, bwd_smem_bytes, P.M, P.Dout);
    if (tr == 64) bwd_t<64>(P, a, st); else if (tr == 32) bwd_t<32>(P, a, st); else bwd_t<16>(P, a, st);
    *nl += 1;
}

void launch_bwd_rowred(const LayerDev& P, const BwdArgs& a, int num_sms, cudaStream_t st, long long* nl) {
    int itiles = (P.M + 63) / 64, ncolblk = (P.M + 127) / 128;
    int per = (P.Dout + 2) * itiles * ncolblk;
    int nsplit = max(1, min((a.R + 255) / 256, (2 * num_sms + per - 1) / per));
    int rows_per_split = (((a.R + nsplit - 1) / nsplit) + KB - 1) / KB * KB;
    nsplit = (a.R + rows_per_split - 1) / rows_per_split;
    k_layer_rowred<<<dim3(nsplit, P.Dout + 2, itiles * ncolblk), DSDGP_NT, 0, st>>>(P, a.U, a.W, a.mubar, a.vbar, a.R,
                                                                                   rows_per_split, ncolblk);
    *nl += 1;
}

// opt in to the full 227 KB of dynamic shared memory once per device (not a stream operation)
cudaError_t layer_kernels_init() {
    const int mx = 227 * 1024;
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(k_layer_fwd<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx))) return e;
    if ((e = cudaFuncSetAttribute(k_layer_fwd<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx))) return e;
    if ((e = cudaFuncSetAttribute(k_layer_fwd<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx))) return e;
    if ((e = cudaFuncSetAttribute(k_layer_bwd<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx))) return e;
    if ((e = cudaFuncSetAttribute(k_layer_bwd<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx))) return e;
    if ((e = cudaFuncSetAttribute(k_layer_bwd<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx))) return e;
    return cudaSuccess;
}
