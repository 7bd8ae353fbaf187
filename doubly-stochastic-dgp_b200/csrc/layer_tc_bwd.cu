// tcgen05 tensor-core path of one SVGP layer: backward over rows (data gradient + per-row quantities the row-reduction
// GEMMs need).  Same tiling as the forward (128 rows on the UMMA M dimension, FOUR threads per row = 16 row warps, one TMA
// producer warp, one MMA-issuing warp): the tile is a chain of latency-bound SIMT phases, which 4 warps per scheduler hide
// far better than 2 (profiles/r2_*).  Per tile:
//   G2[d]: c_d = L_d^T u              1xTF32   (recomputed, as the SIMT path does)
//   G5[d]: ubar += L_d (2 vbar_d c_d) 1xTF32   accumulated over d in TMEM
//   G6   : t = Linv ubar              3xTF32   (non-white)      } w = K^-1 ubar needs the accuracy: the solve
//   G7   : w = Linv^T t               3xTF32                    } amplifies operand rounding by ~sqrt(cond K)
// then k̄, g = 2 k̄ dk/dr2, x̄, and the Z / lengthscale / variance partials exactly as k_layer_bwd (layer_simt.cu).
// Math: tests/algo_mirror.py::layer_bwdA ; reference: TF autodiff of layers.py:178-219.
#include "dsdgp_internal.cuh"
#include "tc_common.cuh"
#include "tc_pack.cuh"

#define TC_ROWS 128
#define TC_NSTAGE 2
#define TC_CHUNK_BYTES 16384
#define TC_THREADS 576
#define TC_ROWTHREADS 512
#define TC_WARP_TMA 16
#define TC_WARP_MMA 17

namespace {
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                   "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                   "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
}  // namespace

// shared-memory plan (bytes from the 1024-aligned base)
struct BwdSmem {
    uint32_t A_u, A_c, Bring, bars, Zs, qmu, mv, xs, red, total;
};
__host__ __device__ inline BwdSmem bwd_smem_plan(int M, int Din, int D) {
    BwdSmem s;
    s.A_u = 0; s.A_c = 65536; s.Bring = 131072;
    s.bars = s.Bring + TC_NSTAGE * tcp::slot_bytes(M);
    s.Zs = s.bars + 256;
    s.qmu = s.Zs + 4 * ((M * Din + 3) & ~3);
    s.mv = s.qmu + 4 * ((M * D + 3) & ~3);
    s.xs = s.mv + 4 * 128 * 2 * D;
    s.red = s.xs + 4 * 128 * Din;
    s.total = s.red + 4 * 64;
    return s;
}

// DINP == D_in and DOUTP == D_out exactly, kernel type and whitening are compile-time: the kernel is I-cache sensitive
// (measured 21% "no instruction" stalls with the generic 13k-instruction body), other shapes use layer_simt.cu.
template <int DINP, int DOUTP, int KERN, bool WHITE>
__global__ void __launch_bounds__(TC_THREADS, 1) k_layer_bwd_tc(LayerDev P, BwdArgs a) {
    using namespace tc;
    extern __shared__ uint8_t smem_raw_b[];
    const uint32_t sbase = (smem_u32(smem_raw_b) + 1023u) & ~1023u;
    uint8_t* sgen = smem_raw_b + (sbase - smem_u32(smem_raw_b));
    const int M = P.M;
    constexpr int Din = DINP, D = DOUTP;
    const BwdSmem sp = bwd_smem_plan(M, Din, D);
    const uint32_t A_u = sbase + sp.A_u, A_c = sbase + sp.A_c, Bring = sbase + sp.Bring, bars = sbase + sp.bars;
    const uint32_t bar_full = bars, bar_empty = bars + 32;
    const uint32_t slotb = tcp::slot_bytes(M);
    const uint32_t bar_au = bars + 64, bar_acc2f = bars + 72 /*[2]*/, bar_cready = bars + 88, bar_cfree = bars + 96;
    const uint32_t bar_ubar = bars + 104, bar_s6 = bars + 112, bar_acc6 = bars + 120, bar_s7 = bars + 128, bar_acc7 = bars + 136;
    const uint32_t tmem_slot = bars + 144;
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(sgen + sp.bars + 144);
    float* Zs = reinterpret_cast<float*>(sgen + sp.Zs);
    float* qmu_s = reinterpret_cast<float*>(sgen + sp.qmu);
    float* mv_s = reinterpret_cast<float*>(sgen + sp.mv);       // [128][2D]: mubar (D) | vbar (D)
    float* xs_s = reinterpret_cast<float*>(sgen + sp.xs);       // [128][Din]
    float* red_s = reinterpret_cast<float*>(sgen + sp.red);     // [64]
    float* g_s = reinterpret_cast<float*>(sgen + sp.A_c);       // [128][MP] once A_c is dead
    const int MP = M | 1;                                        // odd row stride: conflict-free column walks

    const int nkb = (M + 31) / 32, NPAD = (M + 15) & ~15;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = blockIdx.x * TC_ROWS, R = a.R;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_NSTAGE; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_au, TC_ROWTHREADS);
        mbar_init(bar_acc2f, 1); mbar_init(bar_acc2f + 8, 1);
        mbar_init(bar_cready, TC_ROWTHREADS); mbar_init(bar_cfree, 1); mbar_init(bar_ubar, 1);
        mbar_init(bar_s6, TC_ROWTHREADS); mbar_init(bar_acc6, 1); mbar_init(bar_s7, TC_ROWTHREADS); mbar_init(bar_acc7, 1);
        fence_mbar_init();
    }
    if (warp == TC_WARP_TMA) tmem_alloc(tmem_slot, 512);
    for (int e = threadIdx.x; e < M * Din; e += TC_THREADS) Zs[e] = P.Z[e];
    for (int e = threadIdx.x; e < M * D; e += TC_THREADS) qmu_s[e] = P.q_mu[e];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot_gen;

    if (warp == TC_WARP_TMA) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            const char* wsrc = reinterpret_cast<const char*>(P.wpack_fwd);
            int s = 0;
            uint32_t ph = 1;
            auto load = [&](int blk, int pat) {            // one band block = one bulk copy
                const uint32_t bytes = tcp::block_bytes(pat, M);
                mbar_wait(bar_empty + 8 * s, ph);
                mbar_arrive_expect_tx(bar_full + 8 * s, bytes);
                tma_bulk_g2s(Bring + s * slotb, wsrc + (size_t)blk * slotb, bytes, bar_full + 8 * s);
                if (++s == TC_NSTAGE) { s = 0; ph ^= 1; }
            };
            load(tcp::blk_g2(0), tcp::PAT_GE);
            if (D > 1) load(tcp::blk_g2(1), tcp::PAT_GE);
            for (int d = 0; d < D; ++d) { load(tcp::blk_g5(D, d), tcp::PAT_LE); if (d + 2 < D) load(tcp::blk_g2(d + 2), tcp::PAT_GE); }
            if (!WHITE) { load(tcp::blk_g1(0), tcp::PAT_LE); load(tcp::blk_g1(1), tcp::PAT_LE); }
            load(tcp::blk_g1p(0), tcp::PAT_GE); load(tcp::blk_g1p(1), tcp::PAT_GE);
        }
    } else if (warp == TC_WARP_MMA) {
        // ===================== MMA issuer: the whole warp runs the (uniform) control flow, one elected lane issues ========
        int s = 0;
        uint32_t ph = 0;
        const uint64_t desc_hi = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
        auto mkdesc = [&](uint32_t addr) { return desc_hi | (uint64_t)(((addr >> 4) & 0x3FFF) | (1u << 16)); };
        // one band block: D (+)= A * B^T over all its k-blocks.  with_lo: also A_lo against the same block; fresh: the
        // block starts a new accumulator
        auto do_block = [&](uint32_t dcol, uint32_t Ahi, uint32_t Alo, int pat, bool with_lo, bool fresh) {
            mbar_wait(bar_full + 8 * s, ph);
            tc_fence_after();
            const uint32_t bslot = Bring + s * slotb;
            uint32_t boff = 0;
#pragma unroll 1
            for (int q = 0; q < nkb; ++q) {
                const int kb = pat == tcp::PAT_GE ? nkb - 1 - q : q;
                const int nks = min(4, (M - 32 * kb + 7) / 8);
                const int nrows = tcp::band_rows(pat, M, kb);
                const uint32_t bbase = bslot + boff, abase = kb * TC_CHUNK_BYTES;
                boff += 128u * (uint32_t)nrows;
                const uint32_t id = make_idesc_tf32(128, nrows);
                const uint32_t dc = tmem + dcol + (uint32_t)tcp::band_row0(pat, kb);
                if (tc::elect_one()) {
#pragma unroll 1
                    for (int ks = 0; ks < nks; ++ks) {
                        const uint64_t bd = mkdesc(bbase + ks * 32);
                        mma_tf32(dc, mkdesc(Ahi + abase + ks * 32), bd, id, (fresh && q == 0 && ks == 0) ? 0u : 1u);
                        if (with_lo) mma_tf32(dc, mkdesc(Alo + abase + ks * 32), bd, id, 1u);
                    }
                }
                __syncwarp();
            }
            if (tc::elect_one()) mma_commit(bar_empty + 8 * s);
            __syncwarp();
            if (++s == TC_NSTAGE) { s = 0; ph ^= 1; }
        };
        auto commit = [&](uint32_t bar) { if (tc::elect_one()) mma_commit(bar); __syncwarp(); };
        auto g2 = [&](int d) {
            do_block(128u * (uint32_t)(d & 1), A_u, 0, tcp::PAT_GE, false, true);
            commit(bar_acc2f + 8 * (d & 1));
        };
        mbar_wait(bar_au, 0);
        tc_fence_after();
        g2(0);
        if (D > 1) g2(1);
        for (int d = 0; d < D; ++d) {
            mbar_wait(bar_cready, d & 1);
            tc_fence_after();
            do_block(256u, A_c, 0, tcp::PAT_LE, false, d == 0);
            commit(bar_cfree);
            if (d == D - 1) commit(bar_ubar);
            if (d + 2 < D) g2(d + 2);
        }
        if (!WHITE) {
            mbar_wait(bar_s6, 0);
            tc_fence_after();
            do_block(0u, A_c, A_u, tcp::PAT_LE, true, true);
            do_block(0u, A_c, A_u, tcp::PAT_LE, false, false);
            commit(bar_acc6);
        }
        mbar_wait(bar_s7, 0);
        tc_fence_after();
        do_block(128u, A_c, A_u, tcp::PAT_GE, true, true);
        do_block(128u, A_c, A_u, tcp::PAT_GE, false, false);
        commit(bar_acc7);
    } else {
        // ===================== row warps =====================
        const int t = threadIdx.x & 127, qt = threadIdx.x >> 7, row = row0 + t;      // qt: quarter 0..3 of the row's threads
        const bool valid = row < R;
        const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        const uint32_t rsw = (uint32_t)(t & 7);
        const uint32_t rowoff = (uint32_t)((t >> 3) * 1024 + (t & 7) * 128);
        auto a_store4 = [&](uint32_t base, int k4, float4 v) {
            uint32_t off = (uint32_t)(k4 >> 5) * TC_CHUNK_BYTES + rowoff + (((uint32_t)((k4 & 31) >> 2) ^ rsw) << 4);
            *reinterpret_cast<float4*>(sgen + (base - sbase) + off) = v;
        };
        auto store_hi = [&](uint32_t base, int k4, const float* v) {
            float4 hi;
            hi.x = tf32_rna(v[0]); hi.y = tf32_rna(v[1]); hi.z = tf32_rna(v[2]); hi.w = tf32_rna(v[3]);
            a_store4(base, k4, hi);
        };
        auto store_hi_lo = [&](uint32_t bhi, uint32_t blo, int k4, const float* v) {
            float4 hi, lo;
            hi.x = tf32_rna(v[0]); hi.y = tf32_rna(v[1]); hi.z = tf32_rna(v[2]); hi.w = tf32_rna(v[3]);
            lo.x = tf32_lo_trunc(v[0], hi.x); lo.y = tf32_lo_trunc(v[1], hi.y); lo.z = tf32_lo_trunc(v[2], hi.z); lo.w = tf32_lo_trunc(v[3], hi.w);
            a_store4(bhi, k4, hi);
            a_store4(blo, k4, lo);
        };
        // this quarter's accumulator columns [c_lo, c_hi), multiples of 8
        const int nch8 = NPAD >> 3, cq = nch8 >> 2, cr = nch8 & 3;
        const int c_lo = 8 * (qt * cq + min(qt, cr)), c_hi = c_lo + 8 * (cq + (qt < cr ? 1 : 0));
        const float jit = a.jitter;
        const unsigned long long seed = a.sa->seed;
        const int noff = a.sa->n_offset, soff = a.sa->s_offset;
        const float var0 = P.var[0];
        const bool dbg = a.dbg && blockIdx.x == 0 && threadIdx.x == 0;
        int dbi = 0;
#define BSTAMP() do { if (dbg) a.dbg[dbi++] = clock64(); } while (0)
        BSTAMP();   // 0

        // ---- R0: x tile, mubar / vbar (this quarter: d = qt, qt+4, ...)
        float x[DINP], il[DINP];
#pragma unroll
        for (int q = 0; q < DINP; ++q) {
            x[q] = (valid && q < Din) ? a.Xin[(size_t)row * Din + q] : 0.f;
            il[q] = q < Din ? 1.0f / P.ls[P.ard ? q : 0] : 0.f;
            if (qt == 0 && q < Din) xs_s[t * Din + q] = x[q];
        }
#pragma unroll 1
        for (int d = qt; d < D; d += 4) {
            float m = 0.f, v = 0.f;
            if (valid) {
                if (a.fbar) {
                    // z: the draws of the forward pass (injected, or the Philox draws it stored)
                    const float sd = sqrtf(fmaxf(a.Fvar[(size_t)row * D + d] + jit, 1e-30f));
                    float sz = 0.f;
#pragma unroll 1
                    for (int ss = 0; ss < a.S_rep; ++ss) {
                        const size_t o = ((size_t)ss * a.N * (a.S_rep > 1) + row) * D + d;
                        const float fb = a.fbar[o];
                        m += fb; sz = fmaf(fb, a.z[o], sz);
                    }
                    v = sz / (2.f * sd);
                    a.mubar[(size_t)row * D + d] = m;
                    a.vbar[(size_t)row * D + d] = v;
                } else {
                    m = a.mubar[(size_t)row * D + d];
                    v = a.vbar[(size_t)row * D + d];
                }
            }
            mv_s[t * 2 * D + d] = m;
            mv_s[t * 2 * D + D + d] = v;
        }
        // ---- R1: u (this quarter's columns) -> A_u as the G2 operand; loads are issued four chunks ahead of their use
        constexpr bool vec4 = true;               // M % 4 == 0 (tc_bwd_supported)
        auto load_u4 = [&](int c0) -> float4 {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid && c0 < c_hi && c0 + 4 <= M) v = *reinterpret_cast<const float4*>(a.U + (size_t)row * M + c0);
            return v;
        };
        {
            float4 q0 = load_u4(c_lo), q1 = load_u4(c_lo + 4), q2 = load_u4(c_lo + 8), q3 = load_u4(c_lo + 12);
#pragma unroll 1
            for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
                float4 n0 = load_u4(c0 + 16), n1 = load_u4(c0 + 20), n2 = load_u4(c0 + 24), n3 = load_u4(c0 + 28);
                const float4 cur[4] = {q0, q1, q2, q3};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (c0 + 4 * e < c_hi) {
                        const float v[4] = {cur[e].x, cur[e].y, cur[e].z, cur[e].w};
                        store_hi(A_u, c0 + 4 * e, v);
                    }
                }
                q0 = n0; q1 = n1; q2 = n2; q3 = n3;
            }
        }
        if (qt == 3) {       // zero the K padding beyond NPAD (columns NPAD .. 32 nkb) once
            const float z4[4] = {0.f, 0.f, 0.f, 0.f};
            for (int c0 = NPAD; c0 < nkb * 32; c0 += 4) { store_hi(A_u, c0, z4); store_hi(A_c, c0, z4); }
        }
        fence_proxy_async();
        mbar_arrive(bar_au);
        BSTAMP();   // 1: R0+R1 done
        named_bar_sync(1, TC_ROWTHREADS);          // mv_s / xs_s visible
        float mub[DOUTP], vb[DOUTP], vs = 0.f;
#pragma unroll
        for (int d = 0; d < DOUTP; ++d) {
            mub[d] = d < D ? mv_s[t * 2 * D + d] : 0.f;
            vb[d] = d < D ? mv_s[t * 2 * D + D + d] : 0.f;
            vs += vb[d];
        }
        // ---- R2 (deferred, interleaved below): r2_i for this quarter's inducing points -> TMEM scratch columns 384+
        auto gram_chunk = [&](int c0) {
            float r2[8];
            {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int i = min(c0 + u, M - 1);
                    const float4* zr = reinterpret_cast<const float4*>(Zs + i * DINP);
                    float s = 0.f;
#pragma unroll
                    for (int q4 = 0; q4 < DINP / 4; ++q4) {
                        const float4 zv = zr[q4];
                        float d0 = (x[4 * q4] - zv.x) * il[4 * q4], d1 = (x[4 * q4 + 1] - zv.y) * il[4 * q4 + 1];
                        float d2 = (x[4 * q4 + 2] - zv.z) * il[4 * q4 + 2], d3 = (x[4 * q4 + 3] - zv.w) * il[4 * q4 + 3];
                        s = fmaf(d0, d0, s); s = fmaf(d1, d1, s); s = fmaf(d2, d2, s); s = fmaf(d3, d3, s);
                    }
                    r2[u] = s;
                }
            }
            __syncwarp();
            tmem_st8(lane_addr + 384 + c0, r2);
        };
        const int nch = (c_hi - c_lo) >> 3;
        // ---- R3: d loop -- cbar_d = 2 vbar_d c_d -> A_c
        for (int d = 0; d < D; ++d) {
            for (int j = d; j < nch; j += D) gram_chunk(c_lo + 8 * j);
            BSTAMP();   // 2+3d: deferred gram done
            mbar_wait(bar_acc2f + 8 * (d & 1), (d >> 1) & 1);
            BSTAMP();   // 3+3d: G2[d] ready
            if (d > 0) mbar_wait(bar_cfree, (d - 1) & 1);
            BSTAMP();   // 4+3d: A_c free
            tc_fence_after();
            float sc = 0.f;
#pragma unroll
            for (int dd = 0; dd < DOUTP; ++dd) if (dd == d) sc = 2.f * vb[dd];
            for (int c0 = c_lo; c0 < c_hi; c0 += 8) {
                float v[8];
                __syncwarp();
                tmem_ld8(lane_addr + 128 * (d & 1) + c0, v);
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] *= sc;
                store_hi(A_c, c0, v);
                store_hi(A_c, c0 + 4, v + 4);
            }
            tc_fence_before();
            fence_proxy_async();
            mbar_arrive(bar_cready);
        }
        // ---- R4: ubar = Ubar + sum_d mubar_d m_d - (vs k | 2 vs u)  -> operands of the solve
        BSTAMP();   // 26: d loop done
        mbar_wait(bar_ubar, 0);
        tc_fence_after();
        BSTAMP();   // 27: Ubar ready
        {
            float4 ua = WHITE ? load_u4(c_lo) : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 ub4 = WHITE ? load_u4(c_lo + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
            for (int c0 = c_lo; c0 < c_hi; c0 += 8) {
                float4 na = ua, nb4 = ub4;
                if (WHITE) { na = load_u4(c0 + 8); nb4 = load_u4(c0 + 12); }
                float ub[8], r2[8];
                __syncwarp();
                tmem_ld8(lane_addr + 256 + c0, ub);
                __syncwarp();
                tmem_ld8(lane_addr + 384 + c0, r2);
                const float uloc[8] = {ua.x, ua.y, ua.z, ua.w, ub4.x, ub4.y, ub4.z, ub4.w};
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int i = c0 + u;
                    float acc = 0.f;
                    if (i < M) {
                        acc = ub[u];
                        bool done = false;
                        if constexpr (DOUTP % 4 == 0) {
                            if (D == DOUTP) {
                                const float4* qr = reinterpret_cast<const float4*>(qmu_s + i * DOUTP);
#pragma unroll
                                for (int d4 = 0; d4 < DOUTP / 4; ++d4) {
                                    float4 qv = qr[d4];
                                    acc = fmaf(mub[4 * d4], qv.x, acc); acc = fmaf(mub[4 * d4 + 1], qv.y, acc);
                                    acc = fmaf(mub[4 * d4 + 2], qv.z, acc); acc = fmaf(mub[4 * d4 + 3], qv.w, acc);
                                }
                                done = true;
                            }
                        }
                        if (!done) {
#pragma unroll
                            for (int d = 0; d < DOUTP; ++d) if (d < D) acc = fmaf(mub[d], qmu_s[i * D + d], acc);
                        }
                        if (WHITE) acc -= 2.f * vs * uloc[u];
                        else {
                            float k, kp;
                            kern_eval_fast(KERN, r2[u], var0, k, kp);
                            acc -= vs * k;
                        }
                    }
                    ub[u] = acc;
                }
                store_hi_lo(A_c, A_u, c0, ub);
                store_hi_lo(A_c, A_u, c0 + 4, ub + 4);
                ua = na; ub4 = nb4;
            }
        }
        tc_fence_before();
        fence_proxy_async();
        if (!WHITE) {
            mbar_arrive(bar_s6);
            BSTAMP();   // 28: R4 done
            // ---- R5: t = Linv ubar -> operands of the second triangular product
            mbar_wait(bar_acc6, 0);
            tc_fence_after();
            BSTAMP();   // 29: G6 done
            for (int c0 = c_lo; c0 < c_hi; c0 += 8) {
                float v[8];
                __syncwarp();
                tmem_ld8(lane_addr + c0, v);
                store_hi_lo(A_c, A_u, c0, v);
                store_hi_lo(A_c, A_u, c0 + 4, v + 4);
            }
            tc_fence_before();
            fence_proxy_async();
        }
        mbar_arrive(bar_s7);
        BSTAMP();   // 30: R5 done
        // ---- R6: w -> W (global), kbar, g = 2 kbar dk/dr2 -> g_s ; s2 partial
        mbar_wait(bar_acc7, 0);
        tc_fence_after();
        BSTAMP();   // 31: G7 done
        float s2 = 0.f;
        const float inv_var = 1.0f / var0;
        {
            float4 ua = !WHITE ? load_u4(c_lo) : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 ub4 = !WHITE ? load_u4(c_lo + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
            for (int c0 = c_lo; c0 < c_hi; c0 += 8) {
                float4 na = ua, nb4 = ub4;
                if (!WHITE) { na = load_u4(c0 + 8); nb4 = load_u4(c0 + 12); }
                float w[8], r2[8];
                __syncwarp();
                tmem_ld8(lane_addr + 128 + c0, w);
                __syncwarp();
                tmem_ld8(lane_addr + 384 + c0, r2);
                if (valid) {
                    if (c0 + 8 <= M && vec4) {
                        float4* dst = reinterpret_cast<float4*>(a.W + (size_t)row * M + c0);
                        dst[0] = make_float4(w[0], w[1], w[2], w[3]);
                        dst[1] = make_float4(w[4], w[5], w[6], w[7]);
                    } else {
#pragma unroll
                        for (int u = 0; u < 8; ++u) if (c0 + u < M) a.W[(size_t)row * M + c0 + u] = w[u];
                    }
                }
                const float uloc[8] = {ua.x, ua.y, ua.z, ua.w, ub4.x, ub4.y, ub4.z, ub4.w};
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int i = c0 + u;
                    if (i < M) {
                        float kb_ = w[u];
                        if (!WHITE) kb_ -= vs * uloc[u];
                        float k, kp;
                        kern_eval_fast(KERN, r2[u], var0, k, kp);
                        s2 = fmaf(kb_ * k, inv_var, s2);
                        g_s[t * MP + i] = 2.f * kb_ * kp;
                    }
                }
                ua = na; ub4 = nb4;
            }
        }
        BSTAMP();   // 32: R6 done
        if (qt == 0)
#pragma unroll
            for (int d = 0; d < DOUTP; ++d) s2 += vb[d];
        s2 = warp_sum(s2);
        if (lane == 0) red_s[warp] = s2;
        if (threadIdx.x < 32) red_s[32 + threadIdx.x] = 0.f;     // lengthscale accumulators
        named_bar_sync(1, TC_ROWTHREADS);
        if (threadIdx.x == 0) {
            float tot = 0.f;
            for (int w8 = 0; w8 < TC_ROWTHREADS / 32; ++w8) tot += red_s[w8];
            atomicAdd(P.gvar, tot);
        }
        // ---- R7a: xbar (this quarter: q = qt, qt+4, ...)
        if (a.xbar && valid) {
            constexpr int NQ = (DINP + 3) / 4;          // input dimensions per thread: q = qt + 4 j
            float accq[NQ], xq[NQ], ilq[NQ];
#pragma unroll
            for (int j = 0; j < NQ; ++j) { accq[j] = 0.f; xq[j] = 0.f; ilq[j] = 0.f; }
#pragma unroll
            for (int q = 0; q < DINP; ++q)
                if ((q & 3) == qt) { xq[q >> 2] = x[q]; ilq[q >> 2] = il[q]; }
            {
#pragma unroll 4
                for (int i = 0; i < M; ++i) {
                    const float g = g_s[t * MP + i];
#pragma unroll
                    for (int j = 0; j < NQ; ++j) {
                        const int q = qt + 4 * j;
                        const float z = q < DINP ? Zs[i * DINP + q] : 0.f;
                        accq[j] = fmaf(g, xq[j] - z, accq[j]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < NQ; ++j) {
                const int q = qt + 4 * j;
                if (q < Din) {
                    float s = accq[j] * ilq[j] * ilq[j];
                    if (P.mean == DSDGP_MEAN_IDENTITY) {
#pragma unroll
                        for (int d = 0; d < DOUTP; ++d) if (d == q) s += mub[d];
                    } else if (P.mean == DSDGP_MEAN_LINEAR) {
#pragma unroll
                        for (int d = 0; d < DOUTP; ++d) if (d < D) s = fmaf(mub[d], __ldg(&P.meanW[q * D + d]), s);
                    }
                    a.xbar[(size_t)row * Din + q] = s;
                }
            }
        }
        BSTAMP();   // 33: R7a done
        // ---- R7b: Z / lengthscale partials: one (i, q) pair per thread at a time, walking the 128 rows
        {
            const int i = threadIdx.x & 127, rh = threadIdx.x >> 7;      // inducing point, row quarter
            float sa[DINP], sb[DINP];
            if (i < M) {
                float zi[DINP];
#pragma unroll
                for (int q = 0; q < DINP; ++q) { sa[q] = 0.f; sb[q] = 0.f; zi[q] = q < Din ? Zs[i * Din + q] : 0.f; }
                {
#pragma unroll 4
                    for (int r = rh * 32; r < rh * 32 + 32; ++r) {
                        const float g = g_s[r * MP + i];
                        const float4* xr = reinterpret_cast<const float4*>(xs_s + r * DINP);
#pragma unroll
                        for (int q4 = 0; q4 < DINP / 4; ++q4) {
                            const float4 xv = xr[q4];
                            const float xx[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float dd = xx[e] - zi[4 * q4 + e];
                                sa[4 * q4 + e] = fmaf(g, dd, sa[4 * q4 + e]);
                                sb[4 * q4 + e] = fmaf(g * dd, dd, sb[4 * q4 + e]);
                            }
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < DINP; ++q) {
                    if (q < Din) {
                        const float ilq = il[q];
                        atomicAdd(&P.gZ[i * Din + q], -sa[q] * ilq * ilq);
                        sb[q] = -sb[q] * ilq * ilq * ilq;
                    } else sb[q] = 0.f;
                }
                if (!P.ard) {
#pragma unroll
                    for (int q = 1; q < DINP; ++q) sb[0] += sb[q];
                }
            } else {
#pragma unroll
                for (int q = 0; q < DINP; ++q) sb[q] = 0.f;
            }
            // one shared atomic per warp (per input dimension when ARD)
#pragma unroll
            for (int q = 0; q < DINP; ++q) {
                if (q < (P.ard ? Din : 1)) {
                    float tsum = warp_sum(sb[q]);
                    if (lane == 0) atomicAdd(&red_s[32 + q], tsum);
                }
            }
        }
        BSTAMP();   // 34: R7b done
        named_bar_sync(1, TC_ROWTHREADS);
        if (threadIdx.x < (P.ard ? Din : 1)) atomicAdd(&P.gls[threadIdx.x], red_s[32 + threadIdx.x]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == TC_WARP_TMA) { __syncwarp(); tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// (DINP, DOUTP, KERN, WHITE) instances
#define TC_BWD_INSTANCES(X) \
    X(8, 1, 0, false) X(8, 1, 0, true) X(8, 1, 1, false) X(8, 1, 1, true) \
    X(8, 8, 0, false) X(8, 8, 0, true) X(8, 8, 1, false) X(8, 8, 1, true)

bool tc_bwd_supported(const LayerDev& P) {
    if (!(P.M <= 128 && P.M >= 8 && (P.M & 3) == 0 && P.wpack_fwd != nullptr)) return false;
    if (P.Din != 8 || !(P.Dout == 1 || P.Dout == 8)) return false;      // compile-time shapes (see the kernel comment)
    return bwd_smem_plan(P.M, P.Din, P.Dout).total + 1024 <= 227 * 1024;
}

cudaError_t layer_tc_bwd_init() {
    cudaError_t e;
#define X(a_, b_, k_, w_) if ((e = cudaFuncSetAttribute(k_layer_bwd_tc<a_, b_, k_, w_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024))) return e;
    TC_BWD_INSTANCES(X)
#undef X
    return cudaSuccess;
}

void launch_bwd_rows_tc(const LayerDev& P, const BwdArgs& a, cudaStream_t st, long long* nl) {
    int grid = (a.R + TC_ROWS - 1) / TC_ROWS;
    size_t sm = bwd_smem_plan(P.M, P.Din, P.Dout).total + 1024;
    const bool wh = P.white != 0;
#define X(a_, b_, k_, w_) if (P.Din == a_ && P.Dout == b_ && P.kern == k_ && wh == w_) k_layer_bwd_tc<a_, b_, k_, w_><<<grid, TC_THREADS, sm, st>>>(P, a);
    TC_BWD_INSTANCES(X)
#undef X
    *nl += 1;
}
