// tcgen05 tensor-core path of one SVGP layer (forward): 128-row tiles, rows on the UMMA M dimension.
//   G1 : b = Linv k          3xTF32 (hi*hi + lo*hi + hi*lo)   -- feeds s2 - |b|^2 (cancellation) and u
//   G1': u = Linv^T b        3xTF32                            -- non-white only
//   G2 : c_d = L_d^T u       1xTF32 (RN-rounded operands), d = 0..D-1, accumulators double-buffered in TMEM
// Weight tiles ([128 n] x [32 k] tf32, SWIZZLE_128B K-major images packed once per step by k_pack_fwd) stream from
// L2 through a TMA bulk-copy ring; activations are written by the row threads straight into UMMA operand layout;
// accumulators live in TMEM and are read back with tcgen05.ld, one row per thread, so the per-row reductions
// (|b|^2, |c_d|^2, u.q_mu) need no shuffles.  Precision choice measured in DESIGN.md ("TF32 passes").
// Reference arithmetic: layers.py:178-219; mirrored in tests/algo_mirror.py::layer_fwd.
#include "dsdgp_internal.cuh"
#include "tc_common.cuh"

#define TC_ROWS 128
#define TC_NSTAGE 5
#define TC_CHUNK_BYTES 16384
#define TC_THREADS 160

// ----------------------------------------------------------------------------------------------
// weight packing: chunk stream consumed by k_layer_fwd_tc
//   [G1: kb x {hi,lo}] [G1' (non-white): kb x {hi,lo}] [G2: d x kb (hi)]
// ----------------------------------------------------------------------------------------------
__global__ void k_pack_fwd(LayerSet ls) {
    const LayerDev& P = ls.l[blockIdx.y];
    if (!P.wpack_fwd) return;
    const int M = P.M, D = P.Dout, nkb = (M + 31) / 32;
    const int n1 = 2 * nkb, n1p = P.white ? 0 : 2 * nkb, nchunks = n1 + n1p + D * nkb;
    const size_t total = (size_t)nchunks * 4096;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        int c = (int)(e >> 12), w = (int)(e & 4095), n = w >> 5, kk = w & 31;
        float out = 0.f;
        if (c < n1 + n1p) {
            int cc = c < n1 ? c : c - n1, kb = cc >> 1, part = cc & 1, k = kb * 32 + kk;
            if (n < M && k < M) {
                double v = (c < n1) ? P.Linv64[(size_t)n * M + k] : P.Linv64[(size_t)k * M + n];
                float hi = tc::tf32_rna((float)v);
                out = part ? tc::tf32_rna((float)(v - (double)hi)) : hi;
            }
        } else {
            int cc = c - n1 - n1p, d = cc / nkb, kb = cc % nkb, k = kb * 32 + kk;
            if (n < M && k < M && n <= k) out = tc::tf32_rna(P.q_sqrt[((size_t)d * M + k) * M + n]);   // B[n][k] = L_d[k][n]
        }
        *reinterpret_cast<float*>(reinterpret_cast<char*>(P.wpack_fwd) + (size_t)c * TC_CHUNK_BYTES + tc::sw128_offset(n, kk)) = out;
    }
}

void launch_pack_fwd(const LayerSet& ls, cudaStream_t st, long long* nl) {
    k_pack_fwd<<<dim3(96, ls.L), 256, 0, st>>>(ls);
    *nl += 1;
}

// ----------------------------------------------------------------------------------------------
// forward kernel
// ----------------------------------------------------------------------------------------------
template <int DINP, int DOUTP>
__global__ void __launch_bounds__(TC_THREADS, 1) k_layer_fwd_tc(LayerDev P, FwdArgs a) {
    using namespace tc;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
    const uint32_t A_hi = sbase, A_lo = sbase + 65536, Bring = sbase + 131072;
    const uint32_t misc = Bring + TC_NSTAGE * TC_CHUNK_BYTES;
    // barriers (8 B each)
    const uint32_t bar_full = misc, bar_empty = misc + 8 * TC_NSTAGE;
    const uint32_t bar_a = bar_empty + 8 * TC_NSTAGE;        // a_ready[3]
    const uint32_t bar_acc = bar_a + 24;                      // acc_full[2]
    const uint32_t bar_acc2f = bar_acc + 16;                  // acc2_full[2]
    const uint32_t bar_acc2e = bar_acc2f + 16;                // acc2_empty[2]
    const uint32_t tmem_slot = bar_acc2e + 16;
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(sgen + (tmem_slot - sbase));
    float* mean_s = reinterpret_cast<float*>(sgen + (A_lo - sbase));       // [128][D] scratch, valid once A_lo is dead

    const int M = P.M, Din = P.Din, D = P.Dout;
    const int nkb = (M + 31) / 32, NPAD = (M + 15) & ~15;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = blockIdx.x * TC_ROWS, R = a.R;
    const uint32_t copy_bytes = (uint32_t)NPAD * 128u;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_NSTAGE; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int i = 0; i < 3; ++i) mbar_init(bar_a + 8 * i, 128);
        for (int i = 0; i < 2; ++i) { mbar_init(bar_acc + 8 * i, 1); mbar_init(bar_acc2f + 8 * i, 1); mbar_init(bar_acc2e + 8 * i, 128); }
        fence_mbar_init();
    }
    if (warp == 4) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot_gen;
    const uint32_t idesc = make_idesc_tf32(128, NPAD);
    const int n1 = 2 * nkb, n1p = P.white ? 0 : 2 * nkb, NC = n1 + n1p + D * nkb;

    if (warp == 4) {
        // ===================== control warp: TMA producer + MMA issuer (one lane) =====================
        if (lane == 0) {
            const char* wsrc = reinterpret_cast<const char*>(P.wpack_fwd);
            const int PRE = TC_NSTAGE - 1;
            for (int it = 0; it < NC + PRE; ++it) {
                if (it < NC) {
                    int s = it % TC_NSTAGE, n = it / TC_NSTAGE;
                    mbar_wait(bar_empty + 8 * s, (n & 1) ^ 1);
                    mbar_arrive_expect_tx(bar_full + 8 * s, copy_bytes);
                    tma_bulk_g2s(Bring + s * TC_CHUNK_BYTES, wsrc + (size_t)it * TC_CHUNK_BYTES, copy_bytes, bar_full + 8 * s);
                }
                int j = it - PRE;
                if (j < 0) continue;
                int s = j % TC_NSTAGE, n = j / TC_NSTAGE;
                // decode chunk
                int gemm, kb, part = 0, d = 0;
                if (j < n1) { gemm = 0; kb = j >> 1; part = j & 1; }
                else if (j < n1 + n1p) { gemm = 1; kb = (j - n1) >> 1; part = (j - n1) & 1; }
                else { gemm = 2; d = (j - n1 - n1p) / nkb; kb = (j - n1 - n1p) % nkb; }
                // activation / accumulator dependencies
                if (gemm == 0 && j == 0) mbar_wait(bar_a, 0);
                if (gemm == 1 && j == n1) mbar_wait(bar_a + 8, 0);
                if (gemm == 2 && kb == 0) {
                    if (d == 0) mbar_wait(bar_a + 16, 0);
                    if (d >= 2) mbar_wait(bar_acc2e + 8 * (d & 1), ((d >> 1) - 1) & 1);
                }
                mbar_wait(bar_full + 8 * s, n & 1);
                tc_fence_after();
                const int nks = min(4, (M - 32 * kb + 7) / 8);
                const uint32_t dcol = gemm == 0 ? 0u : gemm == 1 ? 128u : 256u + 128u * (uint32_t)(d & 1);
                for (int ks = 0; ks < nks; ++ks) {
                    uint64_t bd = make_desc_sw128_kmajor(Bring + s * TC_CHUNK_BYTES + ks * 32, 1024);
                    uint64_t ah = make_desc_sw128_kmajor(A_hi + kb * TC_CHUNK_BYTES + ks * 32, 1024);
                    uint32_t first = (kb | ks) == 0 ? 0u : 1u;
                    if (gemm < 2) {
                        if (part == 0) {
                            uint64_t al = make_desc_sw128_kmajor(A_lo + kb * TC_CHUNK_BYTES + ks * 32, 1024);
                            mma_tf32(tmem + dcol, ah, bd, idesc, first);
                            mma_tf32(tmem + dcol, al, bd, idesc, 1u);
                        } else {
                            mma_tf32(tmem + dcol, ah, bd, idesc, 1u);
                        }
                    } else {
                        mma_tf32(tmem + dcol, ah, bd, idesc, first);
                    }
                }
                mma_commit(bar_empty + 8 * s);
                if (gemm == 0 && j == n1 - 1) mma_commit(bar_acc);
                if (gemm == 1 && j == n1 + n1p - 1) mma_commit(bar_acc + 8);
                if (gemm == 2 && kb == nkb - 1) mma_commit(bar_acc2f + 8 * (d & 1));
            }
        }
    } else {
        // ===================== row warps: thread t owns row t of the tile (TMEM lane t) =====================
        const int t = threadIdx.x, row = row0 + t;
        const bool valid = row < R;
        const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
        const uint32_t rsw = (uint32_t)(t & 7);
        const uint32_t rowoff = (uint32_t)((t >> 3) * 1024 + (t & 7) * 128);
        auto a_store4 = [&](uint32_t base, int k4, float4 v) {      // k4: first of 4 consecutive k (multiple of 4)
            uint32_t off = (uint32_t)(k4 >> 5) * TC_CHUNK_BYTES + rowoff + (((uint32_t)((k4 & 31) >> 2) ^ rsw) << 4);
            *reinterpret_cast<float4*>(sgen + (base - sbase) + off) = v;
        };
        // ---- Gram: k_i = k(z_i, x) -> A_hi / A_lo (tf32 split)
        float x[DINP], il[DINP];
#pragma unroll
        for (int q = 0; q < DINP; ++q) {
            x[q] = (valid && q < Din) ? a.Xin[(size_t)row * Din + q] : 0.f;
            il[q] = q < Din ? 1.0f / P.ls[P.ard ? q : 0] : 0.f;
        }
        const float var0 = P.var[0];
        for (int i4 = 0; i4 < nkb * 32; i4 += 4) {
            float kv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                int i = i4 + u;
                float k = 0.f;
                if (i < M) {
                    float s = 0.f;
#pragma unroll
                    for (int q = 0; q < DINP; ++q) {
                        float zq = q < Din ? __ldg(&P.Z[(size_t)i * Din + q]) : 0.f;
                        float dd = (x[q] - zq) * il[q];
                        s = fmaf(dd, dd, s);
                    }
                    float kp;
                    kern_eval_f(P.kern, s, var0, k, kp);
                }
                kv[u] = k;
            }
            float4 hi, lo;
            hi.x = tf32_rna(kv[0]); hi.y = tf32_rna(kv[1]); hi.z = tf32_rna(kv[2]); hi.w = tf32_rna(kv[3]);
            lo.x = tf32_rna(kv[0] - hi.x); lo.y = tf32_rna(kv[1] - hi.y); lo.z = tf32_rna(kv[2] - hi.z); lo.w = tf32_rna(kv[3] - hi.w);
            a_store4(A_hi, i4, hi);
            a_store4(A_lo, i4, lo);
        }
        fence_proxy_async();
        mbar_arrive(bar_a);

        // ---- E1: b from TMEM columns [0, NPAD)
        float bn = 0.f;
        float meanv[DOUTP];
#pragma unroll
        for (int d = 0; d < DOUTP; ++d) meanv[d] = 0.f;
        mbar_wait(bar_acc, 0);
        tc_fence_after();
        auto consume_cols = [&](uint32_t dcol, bool is_u, bool write_lo) {
            // reads NPAD accumulator columns of this row; optionally splits them back into the A operand buffers
            for (int c0 = 0; c0 < NPAD; c0 += 16) {
                float v[16];
                __syncwarp();
                tmem_ld16(lane_addr + dcol + c0, v);
                if (!is_u) {
#pragma unroll
                    for (int u = 0; u < 16; ++u) bn = fmaf(v[u], v[u], bn);
                }
                if (is_u) {
#pragma unroll
                    for (int u = 0; u < 16; ++u) {
                        int i = c0 + u;
                        if (i < M) {
#pragma unroll
                            for (int d = 0; d < DOUTP; ++d)
                                if (d < D) meanv[d] = fmaf(v[u], __ldg(&P.q_mu[i * D + d]), meanv[d]);
                            if (valid) a.U[(size_t)row * M + i] = v[u];
                        }
                    }
                }
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    float4 hi, lo;
                    hi.x = tf32_rna(v[4 * g]); hi.y = tf32_rna(v[4 * g + 1]); hi.z = tf32_rna(v[4 * g + 2]); hi.w = tf32_rna(v[4 * g + 3]);
                    a_store4(A_hi, c0 + 4 * g, hi);
                    if (write_lo) {
                        lo.x = tf32_rna(v[4 * g] - hi.x); lo.y = tf32_rna(v[4 * g + 1] - hi.y);
                        lo.z = tf32_rna(v[4 * g + 2] - hi.z); lo.w = tf32_rna(v[4 * g + 3] - hi.w);
                        a_store4(A_lo, c0 + 4 * g, lo);
                    }
                }
            }
        };
        if (P.white) {
            // u = b
            for (int c0 = 0; c0 < NPAD; c0 += 16) {
                float v[16];
                __syncwarp();
                tmem_ld16(lane_addr + c0, v);
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    bn = fmaf(v[u], v[u], bn);
                    int i = c0 + u;
                    if (i < M) {
#pragma unroll
                        for (int d = 0; d < DOUTP; ++d)
                            if (d < D) meanv[d] = fmaf(v[u], __ldg(&P.q_mu[i * D + d]), meanv[d]);
                        if (valid) a.U[(size_t)row * M + i] = v[u];
                    }
                }
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    float4 hi;
                    hi.x = tf32_rna(v[4 * g]); hi.y = tf32_rna(v[4 * g + 1]); hi.z = tf32_rna(v[4 * g + 2]); hi.w = tf32_rna(v[4 * g + 3]);
                    a_store4(A_hi, c0 + 4 * g, hi);
                }
            }
            tc_fence_before();
            fence_proxy_async();
            mbar_arrive(bar_a + 16);
        } else {
            consume_cols(0u, false, true);
            tc_fence_before();
            fence_proxy_async();
            mbar_arrive(bar_a + 8);
            // ---- E1': u from TMEM columns [128, 128+NPAD)
            mbar_wait(bar_acc + 8, 0);
            tc_fence_after();
            consume_cols(128u, true, false);
            tc_fence_before();
            fence_proxy_async();
            mbar_arrive(bar_a + 16);
        }
        // mean scratch (A_lo is dead: G2 reads A_hi only)
#pragma unroll
        for (int d = 0; d < DOUTP; ++d)
            if (d < D) mean_s[t * D + d] = meanv[d];

        // ---- E2: |c_d|^2, variance, draw
        const float jit = a.jitter;
        const unsigned long long seed = a.sa->seed;
        const int noff = a.sa->n_offset;
        for (int d = 0; d < D; ++d) {
            const int b = d & 1;
            mbar_wait(bar_acc2f + 8 * b, (d >> 1) & 1);
            tc_fence_after();
            float s = 0.f;
            for (int c0 = 0; c0 < NPAD; c0 += 16) {
                float v[16];
                __syncwarp();
                tmem_ld16(lane_addr + 256 + 128 * b + c0, v);
#pragma unroll
                for (int u = 0; u < 16; ++u) s = fmaf(v[u], v[u], s);
            }
            tc_fence_before();
            mbar_arrive(bar_acc2e + 8 * b);
            if (valid) {
                float mean = mean_s[t * D + d];
                if (P.mean == DSDGP_MEAN_IDENTITY) mean += a.Xin[(size_t)row * Din + d];
                else if (P.mean == DSDGP_MEAN_LINEAR) {
                    float ms = P.meanB[d];
                    for (int q = 0; q < Din; ++q) ms = fmaf(a.Xin[(size_t)row * Din + q], __ldg(&P.meanW[q * D + d]), ms);
                    mean += ms;
                }
                float v = var0 - bn + s;
                a.Fmean[(size_t)row * D + d] = mean;
                a.Fvar[(size_t)row * D + d] = v;
                if (a.F) {
                    float sd = sqrtf(fmaxf(v + jit, 1e-30f));
                    if (a.S_rep == 1) {
                        int ss = row / a.N, n = row % a.N;
                        float z = a.z ? a.z[(size_t)row * D + d] : dsdgp_normal(seed, P.idx, ss, n + noff, d);
                        a.F[(size_t)row * D + d] = fmaf(z, sd, mean);
                    } else {
                        for (int ss = 0; ss < a.S_rep; ++ss) {
                            size_t o = ((size_t)ss * a.N + row) * D + d;
                            float z = a.z ? a.z[o] : dsdgp_normal(seed, P.idx, ss, row + noff, d);
                            a.F[o] = fmaf(z, sd, mean);
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) { __syncwarp(); tc_fence_after(); tmem_dealloc(tmem, 512); }
}

static size_t tc_fwd_smem() { return 1024 + 131072 + TC_NSTAGE * TC_CHUNK_BYTES + 256; }

bool tc_fwd_supported(const LayerDev& P) { return P.M <= 128 && P.M >= 8 && P.Din <= 16 && P.Dout <= 32 && P.wpack_fwd != nullptr; }

#define TC_FWD_INSTANCES(X) X(8, 1) X(8, 8) X(8, 32) X(16, 1) X(16, 8) X(16, 32)

cudaError_t layer_tc_init() {
    cudaError_t e;
#define X(a, b) if ((e = cudaFuncSetAttribute(k_layer_fwd_tc<a, b>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_fwd_smem()))) return e;
    TC_FWD_INSTANCES(X)
#undef X
    return cudaSuccess;
}

size_t tc_fwd_pack_bytes(int M, int D, int white) {
    int nkb = (M + 31) / 32;
    return (size_t)((white ? 2 : 4) * nkb + D * nkb) * TC_CHUNK_BYTES;
}

void launch_fwd_tc(const LayerDev& P, const FwdArgs& a, cudaStream_t st, long long* nl) {
    int grid = (a.R + TC_ROWS - 1) / TC_ROWS;
    int dinp = P.Din <= 8 ? 8 : 16, doutp = P.Dout <= 1 ? 1 : P.Dout <= 8 ? 8 : 32;
#define X(a_, b_) if (dinp == a_ && doutp == b_) k_layer_fwd_tc<a_, b_><<<grid, TC_THREADS, tc_fwd_smem(), st>>>(P, a);
    TC_FWD_INSTANCES(X)
#undef X
    *nl += 1;
}
