// tcgen05 tensor-core path of one SVGP layer: backward over rows (data gradient + per-row quantities the row-reduction
// GEMMs need).  128 rows on the UMMA M dimension, FOUR threads per row (16 row warps), one TMA producer warp, one
// MMA-issuing warp.  Per tile:
//   GS[d]: y_d = S_d u,  S_d = L_d L_d^T   1xTF32   one product per output instead of c_d = L_d^T u followed by
//                                                    L_d (2 vbar_d c_d): the row threads only accumulate
//                                                    ubar += 2 vbar_d y_d in registers -- no per-d SIMT -> smem -> MMA
//                                                    round trip (round 1: 30k of the tile's 88k cycles)
//   G6   : t = Linv ubar                    3xTF32   (non-white)      } w = K^-1 ubar needs the accuracy: the solve
//   G7   : w = Linv^T t                     3xTF32                    } amplifies operand rounding by ~sqrt(cond K)
// then kbar, g = 2 kbar dk/dr2, xbar, and the Z / lengthscale / variance partials.
// Weights stream from L2 as 32-wide k-block "bands" (<= 16 KB, one TMA bulk copy each) through a ring of sub-slots: four
// dedicated ones plus, while the S_d products run, the four 16 KB blocks of the (then unused) second operand buffer, so
// up to eight bands are in flight.
// Math: tests/algo_mirror.py::layer_bwdA ; reference: TF autodiff of layers.py:178-219.
#include "dsdgp_internal.cuh"
#include "tc_common.cuh"
#include "tc_pack.cuh"

#define TC_ROWS 128
#define TC_CHUNK_BYTES 16384
#define TC_THREADS 576
#define TC_ROWTHREADS 512
#define TC_WARP_TMA 16
#define TC_WARP_MMA 17
#define BW_NR_MAX 5              // dedicated ring sub-slots (one full band = 128 * NPAD bytes each), as many as fit
#define BW_NS_MAX (BW_NR_MAX + 4) // + the sub-slots carved out of A_c while it is free (S_d phase)

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
    uint32_t A_u, A_c, ring, bars, Zs, ZsT, qmu, mv, xs, red, total;
    int nr, nsa;           // ring sub-slots: dedicated, and in total while A_c is free
};
__host__ __device__ inline BwdSmem bwd_smem_plan(int M, int Din, int D) {
    BwdSmem s;
    const uint32_t band = 128u * (uint32_t)tcp::npad_of(M);
    const uint32_t misc = 512 + 4 * ((M * Din + 3) & ~3) + 4 * Din * ((M + 3) & ~3) + 4 * ((M * D + 3) & ~3) +
                          4 * 128 * 2 * D + 4 * 128 * Din + 4 * 96;
    // as many dedicated sub-slots as fit next to the two 64 KB operand buffers (at least 2)
    int nr = (int)((227u * 1024u - 1024u - 131072u - misc) / band);
    nr = nr > BW_NR_MAX ? BW_NR_MAX : nr < 2 ? 2 : nr;
    s.nr = nr; s.nsa = nr + (int)(65536u / band);
    s.A_u = 0; s.A_c = 65536; s.ring = 131072;
    s.bars = s.ring + (uint32_t)nr * band;
    s.Zs = s.bars + 512;
    s.ZsT = s.Zs + 4 * ((M * Din + 3) & ~3);
    s.qmu = s.ZsT + 4 * Din * ((M + 3) & ~3);
    s.mv = s.qmu + 4 * ((M * D + 3) & ~3);
    s.xs = s.mv + 4 * 128 * 2 * D;
    s.red = s.xs + 4 * 128 * Din;
    s.total = s.red + 4 * 96;      // [0,32) warp partials, [32,64) ARD lengthscale sums, [64,96) 1/lengthscale
    return s;
}

// DINP == D_in and DOUTP == D_out exactly, kernel type and whitening are compile-time: the kernel is I-cache sensitive
// (measured 21% "no instruction" stalls with a generic 13k-instruction body), other shapes use layer_simt.cu.
template <int DINP, int DOUTP, int KERN, bool WHITE>
__global__ void __launch_bounds__(TC_THREADS, 1) k_layer_bwd_tc(LayerDev P, BwdArgs a) {
    using namespace tc;
    extern __shared__ uint8_t smem_raw_b[];
    const uint32_t sbase = (smem_u32(smem_raw_b) + 1023u) & ~1023u;
    uint8_t* sgen = smem_raw_b + (sbase - smem_u32(smem_raw_b));
    const int M = P.M;
    constexpr int Din = DINP, D = DOUTP;
    const BwdSmem sp = bwd_smem_plan(M, Din, D);
    const uint32_t A_u = sbase + sp.A_u, A_c = sbase + sp.A_c, ring = sbase + sp.ring, bars = sbase + sp.bars;
    const uint32_t bar_full = bars, bar_empty = bars + 8 * BW_NS_MAX;              // [BW_NS_MAX] each
    const int BW_NR = sp.nr, BW_NSA = sp.nsa;
    const uint32_t bar_au = bars + 16 * BW_NS_MAX, bar_yfull = bar_au + 8 /*[2]*/, bar_yfree = bar_au + 24 /*[2]*/;
    const uint32_t bar_s6 = bar_au + 40, bar_acc6 = bar_au + 48, bar_s7 = bar_au + 56, bar_acc7 = bar_au + 64;
    const uint32_t tmem_slot = bar_au + 72;
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(sgen + sp.bars + 16 * BW_NS_MAX + 72);
    float* Zs = reinterpret_cast<float*>(sgen + sp.Zs);         // [M][Din], scaled by 1/lengthscale
    float* ZsT = reinterpret_cast<float*>(sgen + sp.ZsT);       // [Din][M4] the same, transposed (M4 = M rounded up to 4)
    float* qmu_s = reinterpret_cast<float*>(sgen + sp.qmu);
    float* mv_s = reinterpret_cast<float*>(sgen + sp.mv);       // [128][2D]: mubar (D) | vbar (D)
    float* xs_s = reinterpret_cast<float*>(sgen + sp.xs);       // [128][Din], scaled by 1/lengthscale
    float* red_s = reinterpret_cast<float*>(sgen + sp.red);     // [96]
    float* il_s = red_s + 64;                                   // [Din] 1/lengthscale
    float* g_s = reinterpret_cast<float*>(sgen + sp.A_c);       // [128][MP] once A_c is dead
    float* zred_s = reinterpret_cast<float*>(sgen + sp.A_u);    // [4][M][Din] once A_u is dead
    // row stride of g_s: a multiple of 4 (16-byte row-wise accesses) with MP/4 odd, so that eight consecutive rows start in
    // eight different 16-byte bank groups; column walks (consecutive i) are conflict-free for any stride
    const int MP = ((M >> 2) & 1) ? M : M + 4;

    const int nkb = (M + 31) / 32, NPAD = (M + 15) & ~15;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = blockIdx.x * TC_ROWS, R = a.R;
    const uint32_t band_full = 128u * (uint32_t)NPAD;           // one k-block of a square (S_d) operand
    auto slot_addr = [&](int s) { return s < BW_NR ? ring + (uint32_t)s * band_full : A_c + (uint32_t)(s - BW_NR) * band_full; };

    // every CTA of this grid is resident (or done) once the last one has issued this: from then on the next layer's row kernel
    // (launched as a programmatic dependent) may take the SMs this grid's tail wave leaves idle; its tiles wait on tile_done
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (threadIdx.x == 0) {
        for (int s = 0; s < BW_NSA; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_au, TC_ROWTHREADS);
        mbar_init(bar_yfull, 1); mbar_init(bar_yfull + 8, 1);
        mbar_init(bar_yfree, TC_ROWTHREADS); mbar_init(bar_yfree + 8, TC_ROWTHREADS);
        mbar_init(bar_s6, TC_ROWTHREADS); mbar_init(bar_acc6, 1); mbar_init(bar_s7, TC_ROWTHREADS); mbar_init(bar_acc7, 1);
        fence_mbar_init();
    }
    if (warp == TC_WARP_TMA) tmem_alloc(tmem_slot, 512);
    const int M4 = (M + 3) & ~3;
    for (int e = threadIdx.x; e < M * Din; e += TC_THREADS) {
        const float zv = P.Z[e] * (1.0f / P.ls[P.ard ? e % Din : 0]);
        Zs[e] = zv;
        ZsT[(e % Din) * M4 + e / Din] = zv;
    }
    if (threadIdx.x < Din) il_s[threadIdx.x] = 1.0f / P.ls[P.ard ? threadIdx.x : 0];
    for (int e = threadIdx.x; e < M * D; e += TC_THREADS) qmu_s[e] = P.q_mu[e];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot_gen;

    if (warp == TC_WARP_TMA) {
        // ===================== TMA producer: one bulk copy per band, in the order the MMA warp consumes them ============
        if (lane == 0) {
            const char* wsrc = reinterpret_cast<const char*>(P.wpack_fwd);
            uint32_t par = 0;          // bit s: parity of the next use of sub-slot s
            int s = 0;
            auto load = [&](const char* src, uint32_t bytes, int nslots) {
                mbar_wait(bar_empty + 8 * s, ((par >> s) & 1u) ^ 1u);
                mbar_arrive_expect_tx(bar_full + 8 * s, bytes);
                tma_bulk_g2s(slot_addr(s), src, bytes, bar_full + 8 * s);
                par ^= 1u << s;
                if (++s == nslots) s = 0;
            };
            const char* ssrc = wsrc + tcp::s_region_offset(M, D);
            // Only the first two bands are requested before the row threads have fetched the tile's inputs: every CTA of a wave
            // starts at the same time, and a full ring of weight requests ahead of the (smaller, but latency-critical) u / x /
            // mubar loads made those wait ~8k cycles for L2 bandwidth (profiles/r2_bwd_tile_stamps.md)
            int nband = 0;
            for (int d = 0; d < D; ++d)
                for (int kb = 0; kb < nkb; ++kb) {
                    if (nband++ == 2) mbar_wait(bar_au, 0);
                    load(ssrc + (size_t)d * tcp::sfull_bytes(M) + (size_t)kb * band_full, tcp::sfull_band_tx_bytes(M, kb), BW_NSA);
                }
            s = 0;
            const uint32_t slotb = tcp::slot_bytes(M);
            auto load_block = [&](int blk, int pat) {
                for (int q = 0; q < nkb; ++q) {
                    const int kb = pat == tcp::PAT_GE ? nkb - 1 - q : q;
                    load(wsrc + (size_t)blk * slotb + tcp::band_offset(pat, M, kb), tcp::band_tx_bytes(pat, M, kb), BW_NR);
                }
            };
            if (!WHITE) { load_block(tcp::blk_g1(0), tcp::PAT_LE); load_block(tcp::blk_g1(1), tcp::PAT_LE); }
            load_block(tcp::blk_g1p(0), tcp::PAT_GE); load_block(tcp::blk_g1p(1), tcp::PAT_GE);
        }
    } else if (warp == TC_WARP_MMA) {
        // ===================== MMA issuer: the whole warp runs the (uniform) control flow, one elected lane issues ========
        uint32_t par = 0;
        int s = 0;
        const uint64_t desc_hi = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
        auto mkdesc = [&](uint32_t addr) { return desc_hi | (uint64_t)(((addr >> 4) & 0x3FFF) | (1u << 16)); };
        // tail band (tc_pack.cuh tail32): SWIZZLE_32B K-major, 8-row atoms 256 B apart
        const uint64_t desc32_hi = ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)6 << 61);
        auto mkdesc32 = [&](uint32_t addr) { return desc32_hi | (uint64_t)(((addr >> 4) & 0x3FFF) | (1u << 16)); };
        // one band: D[:, n0 .. n0+nrows) (+)= A[:, 32 kb .. ) * band^T.  with_lo: also A_lo against the same band
        auto do_band = [&](uint32_t dcol, uint32_t Ahi, uint32_t Alo, int kb, int n0, int nrows, bool with_lo, bool fresh, int nslots) {
            mbar_wait(bar_full + 8 * s, (par >> s) & 1u);
            tc_fence_after();
            const uint32_t bbase = slot_addr(s), abase = kb * TC_CHUNK_BYTES;
            const int nks = min(4, (M - 32 * kb + 7) / 8);
            const uint32_t id = make_idesc_tf32(128, nrows);
            const uint32_t dc = tmem + dcol + (uint32_t)n0;
            const bool t32 = tcp::is_tail32(M, kb);          // (then nks == 1)
            if (tc::elect_one()) {
#pragma unroll 1
                for (int ks = 0; ks < nks; ++ks) {
                    const uint64_t bd = t32 ? mkdesc32(bbase) : mkdesc(bbase + ks * 32);
                    mma_tf32(dc, mkdesc(Ahi + abase + ks * 32), bd, id, (fresh && ks == 0) ? 0u : 1u);
                    if (with_lo) mma_tf32(dc, mkdesc(Alo + abase + ks * 32), bd, id, 1u);
                }
                mma_commit(bar_empty + 8 * s);
            }
            __syncwarp();
            par ^= 1u << s;
            if (++s == nslots) s = 0;
        };
        auto commit = [&](uint32_t bar) { if (tc::elect_one()) mma_commit(bar); __syncwarp(); };
        // triangular band block: PAT_LE holds rows [32 kb, NPAD) of k-block kb, PAT_GE rows [0, min(NPAD, 32 kb + 32)); in
        // both orders the FIRST band of a block spans all NPAD accumulator columns, so it is the one that clears them
        auto do_block = [&](uint32_t dcol, int pat, bool with_lo, bool fresh) {
            for (int q = 0; q < nkb; ++q) {
                const int kb = pat == tcp::PAT_GE ? nkb - 1 - q : q;
                do_band(dcol, A_c, A_u, kb, tcp::band_row0(pat, kb), tcp::band_rows(pat, M, kb), with_lo, fresh && q == 0, BW_NR);
            }
        };
        mbar_wait(bar_au, 0);
        tc_fence_after();
        for (int d = 0; d < D; ++d) {
            if (d >= 2) { mbar_wait(bar_yfree + 8 * (d & 1), ((d >> 1) - 1) & 1); tc_fence_after(); }
            for (int kb = 0; kb < nkb; ++kb)
                do_band(128u * (uint32_t)(d & 1), A_u, 0, kb, 0, NPAD, false, kb == 0, BW_NSA);
            commit(bar_yfull + 8 * (d & 1));
        }
        s = 0;
        if (!WHITE) {
            mbar_wait(bar_s6, 0);
            tc_fence_after();
            do_block(0u, tcp::PAT_LE, true, true);
            do_block(0u, tcp::PAT_LE, false, false);
            commit(bar_acc6);
        }
        mbar_wait(bar_s7, 0);
        tc_fence_after();
        do_block(128u, tcp::PAT_GE, true, true);
        do_block(128u, tcp::PAT_GE, false, false);
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
        // this quarter's accumulator columns [c_lo, c_hi), multiples of 8, at most 32 wide
        const int nch8 = NPAD >> 3, cq = nch8 >> 2, cr = nch8 & 3;
        const int c_lo = 8 * (qt * cq + min(qt, cr)), c_hi = c_lo + 8 * (cq + (qt < cr ? 1 : 0));
        const int nch = (c_hi - c_lo) >> 3;
        const float jit = a.jitter;
        const float var0 = P.var[0];
        const bool dbg = a.dbg && blockIdx.x == 0 && threadIdx.x == 0;
        int dbi = 0;
#define BSTAMP() do { if (dbg) a.dbg[dbi++] = clock64(); } while (0)
        BSTAMP();   // 0

        // ---- R0 / R1: every global load of the tile's inputs is issued before the first dependent instruction: a dependent
        // round trip costs ~2k cycles here (measured: x 2.1k, mubar/vbar 1.9k per d, u 2-5k when they were serialised)
        // upstream dE/dF of this tile's rows: published by tile blockIdx.x of the previous launch.  Behind another backward
        // launch only fbar is new (U, x, Fvar, z date from the forward pass), behind the forward chain's tail everything is
        auto wait_upstream = [&]() {
            if (a.tile_wait) {
                if (lane == 0) {
                    const unsigned epoch = a.sa->epoch;
                    while (ld_acquire_gpu(a.tile_wait + blockIdx.x) != epoch) __nanosleep(100);
                }
                __syncwarp();
            }
        };
        if (a.wait_before_loads && a.tile_wait) {
            // possibly a long wait (the producer may be a forward tile still in flight): ONE polling thread per CTA with a long
            // back-off, so that a hundred waiting CTAs do not hammer the L2 lines the producers publish through
            const unsigned epoch = a.sa->epoch;
            if (a.wait_count > 0) {
                for (int i = threadIdx.x; i < a.wait_count; i += TC_ROWTHREADS)
                    while (ld_acquire_gpu(a.tile_wait + i) != epoch) __nanosleep(1000);
            } else if (threadIdx.x == 0) {
                while (ld_acquire_gpu(a.tile_wait + blockIdx.x) != epoch) __nanosleep(1000);
            }
            named_bar_sync(1, TC_ROWTHREADS);
        }
        auto load_u4 = [&](int c0) -> float4 {       // M % 4 == 0 (tc_bwd_supported)
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid && c0 < c_hi && c0 + 4 <= M) v = *reinterpret_cast<const float4*>(a.U + (size_t)row * M + c0);
            return v;
        };
        float4 uq[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) uq[e] = load_u4(c_lo + 4 * e);
        float xs[DINP];
#pragma unroll
        for (int q = 0; q < DINP; ++q) xs[q] = valid ? a.Xin[(size_t)row * Din + q] : 0.f;
        if (!a.wait_before_loads) wait_upstream();
        // mubar / vbar (this quarter: d = qt, qt+4, ...)
        constexpr int NDQ = (DOUTP + 3) / 4;
        float fv[NDQ], fm[NDQ], fz[NDQ];
        const bool single = a.S_rep == 1;
#pragma unroll
        for (int j = 0; j < NDQ; ++j) {
            const int d = qt + 4 * j;
            fv[j] = 0.f; fm[j] = 0.f; fz[j] = 0.f;
            if (valid && d < D) {
                const size_t o = (size_t)row * D + d;
                if (a.fbar) {
                    fv[j] = a.Fvar[o];
                    if (single) { fm[j] = a.fbar[o]; fz[j] = a.z[o]; }
                } else { fm[j] = a.mubar[o]; fv[j] = a.vbar[o]; }
            }
        }
#pragma unroll
        for (int q = 0; q < DINP; ++q) {
            xs[q] *= il_s[q];
            if (qt == 0) xs_s[t * Din + q] = xs[q];
        }
        BSTAMP();   // 1: inputs loaded
        float fsum[NDQ], fsz[NDQ];
#pragma unroll
        for (int j = 0; j < NDQ; ++j) { fsum[j] = 0.f; fsz[j] = 0.f; }
        if (!single && a.fbar && valid) {
            // layer-1 fold: the row's S samples.  Four samples of every d of this thread are loaded before the first use
            // (one sample at a time this was 2 * S dependent round trips of ~2k cycles: half of the 8-tile launch's 77 us)
#pragma unroll 1
            for (int s0 = 0; s0 < a.S_rep; s0 += 4) {
                float fb[4][NDQ], zz[4][NDQ];
#pragma unroll
                for (int e = 0; e < 4; ++e)
#pragma unroll
                    for (int j = 0; j < NDQ; ++j) {
                        const int d = qt + 4 * j;
                        const bool ok = s0 + e < a.S_rep && d < D;
                        const size_t o = ((size_t)(s0 + e) * a.N + row) * D + d;
                        fb[e][j] = ok ? a.fbar[o] : 0.f;
                        zz[e][j] = ok ? a.z[o] : 0.f;
                    }
#pragma unroll
                for (int e = 0; e < 4; ++e)
#pragma unroll
                    for (int j = 0; j < NDQ; ++j) { fsum[j] += fb[e][j]; fsz[j] = fmaf(fb[e][j], zz[e][j], fsz[j]); }
            }
        }
#pragma unroll
        for (int j = 0; j < NDQ; ++j) {
            const int d = qt + 4 * j;
            if (d < D) {
                float m = 0.f, v = 0.f;
                if (valid) {
                    if (a.fbar) {
                        // z: the draws of the forward pass (injected, or the Philox draws it stored)
                        const float sd = sqrtf(fmaxf(fv[j] + jit, 1e-30f));
                        float sz;
                        if (single) { m = fm[j]; sz = fm[j] * fz[j]; }
                        else { m = fsum[j]; sz = fsz[j]; }
                        v = sz / (2.f * sd);
                        a.mubar[(size_t)row * D + d] = m;
                        a.vbar[(size_t)row * D + d] = v;
                    } else { m = fm[j]; v = fv[j]; }
                }
                mv_s[t * 2 * D + d] = m;
                mv_s[t * 2 * D + D + d] = v;
            }
        }
        BSTAMP();   // 2: mubar / vbar done
        // ---- R1: u (this quarter's columns) -> A_u as the operand of the S_d products
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            if (c_lo + 4 * e < c_hi) {
                const float v[4] = {uq[e].x, uq[e].y, uq[e].z, uq[e].w};
                store_hi(A_u, c_lo + 4 * e, v);
            }
        }
        fence_proxy_async();
        mbar_arrive(bar_au);
        BSTAMP();   // 3: R0+R1 done
        named_bar_sync(1, TC_ROWTHREADS);          // mv_s / xs_s visible
        float mub[DOUTP], vb[DOUTP], vs = 0.f;
#pragma unroll
        for (int d = 0; d < DOUTP; ++d) {
            mub[d] = mv_s[t * 2 * D + d];
            vb[d] = mv_s[t * 2 * D + D + d];
            vs += vb[d];
        }
        // ---- R2 (interleaved below): r2_i for this quarter's inducing points -> TMEM scratch columns 384+
        auto gram_chunk = [&](int c0) {
            float r2[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = min(c0 + u, M - 1);
                const float4* zr = reinterpret_cast<const float4*>(Zs + i * DINP);
                float s = 0.f;
#pragma unroll
                for (int q4 = 0; q4 < DINP / 4; ++q4) {
                    const float4 zv = zr[q4];
                    const float d0 = xs[4 * q4] - zv.x, d1 = xs[4 * q4 + 1] - zv.y;
                    const float d2 = xs[4 * q4 + 2] - zv.z, d3 = xs[4 * q4 + 3] - zv.w;
                    s = fmaf(d0, d0, s); s = fmaf(d1, d1, s); s = fmaf(d2, d2, s); s = fmaf(d3, d3, s);
                }
                r2[u] = s;
            }
            __syncwarp();
            tmem_st8(lane_addr + 384 + c0, r2);
        };
        // ---- R3: ubar (this quarter's columns, registers) += 2 vbar_d (S_d u)
        float ub[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) ub[c] = 0.f;
#pragma unroll
        for (int d = 0; d < DOUTP; ++d) {
            for (int j = d; j < nch; j += D) gram_chunk(c_lo + 8 * j);     // every chunk exactly once over d = 0..D-1
            mbar_wait(bar_yfull + 8 * (d & 1), (d >> 1) & 1);
            tc_fence_after();
            BSTAMP();   // 2+d: y_d ready
            const float sc = 2.f * vb[d];
            {
                // this quarter's columns of y_d in two batches of two loads, one wait per batch
#pragma unroll
                for (int b2 = 0; b2 < 2; ++b2) {
                    uint32_t r[2][8];
                    __syncwarp();
#pragma unroll
                    for (int c8 = 0; c8 < 2; ++c8)
                        if (2 * b2 + c8 < nch) tmem_ld8_nw(lane_addr + 128 * (d & 1) + c_lo + 8 * (2 * b2 + c8), r[c8]);
                    tmem_ld_wait();
#pragma unroll
                    for (int c8 = 0; c8 < 2; ++c8) {
                        if (2 * b2 + c8 < nch) {
#pragma unroll
                            for (int u = 0; u < 8; ++u)
                                ub[8 * (2 * b2 + c8) + u] = fmaf(sc, __uint_as_float(r[c8][u]), ub[8 * (2 * b2 + c8) + u]);
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(bar_yfree + 8 * (d & 1));
        }
        // ---- R4: ubar += sum_d mubar_d m_d - (vs k | 2 vs u)  -> operands of the solve
        BSTAMP();   // 2+D: S_d products consumed
        {
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {
                if (c8 < nch) {
                    const int c0 = c_lo + 8 * c8;
                    float r2[8];
                    float4 ua[2];
                    if (!WHITE) {
                        __syncwarp();
                        tmem_ld8(lane_addr + 384 + c0, r2);
                    } else { ua[0] = load_u4(c0); ua[1] = load_u4(c0 + 4); }
                    float o8[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int i = c0 + u;
                        float acc = 0.f;
                        if (i < M) {
                            acc = ub[8 * c8 + u];
                            if constexpr (DOUTP % 4 == 0) {
                                const float4* qr = reinterpret_cast<const float4*>(qmu_s + i * DOUTP);
#pragma unroll
                                for (int d4 = 0; d4 < DOUTP / 4; ++d4) {
                                    const float4 qv = qr[d4];
                                    acc = fmaf(mub[4 * d4], qv.x, acc); acc = fmaf(mub[4 * d4 + 1], qv.y, acc);
                                    acc = fmaf(mub[4 * d4 + 2], qv.z, acc); acc = fmaf(mub[4 * d4 + 3], qv.w, acc);
                                }
                            } else {
#pragma unroll
                                for (int d = 0; d < DOUTP; ++d) acc = fmaf(mub[d], qmu_s[i * D + d], acc);
                            }
                            if (WHITE) {
                                const float4 uu = ua[u >> 2];
                                const float ue = (u & 3) == 0 ? uu.x : (u & 3) == 1 ? uu.y : (u & 3) == 2 ? uu.z : uu.w;
                                acc -= 2.f * vs * ue;
                            } else {
                                float k, kp;
                                kern_eval_fast(KERN, r2[u], var0, k, kp);
                                acc -= vs * k;
                            }
                        }
                        o8[u] = acc;
                    }
                    store_hi_lo(A_c, A_u, c0, o8);
                    store_hi_lo(A_c, A_u, c0 + 4, o8 + 4);
                }
            }
        }
        tc_fence_before();
        fence_proxy_async();
        if (!WHITE) {
            mbar_arrive(bar_s6);
            BSTAMP();   // R4 done
            // ---- R5: t = Linv ubar -> operands of the second triangular product
            mbar_wait(bar_acc6, 0);
            tc_fence_after();
            BSTAMP();   // G6 done
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
        BSTAMP();   // R5 done
        // ---- R6: w -> W (global), kbar, g = 2 kbar dk/dr2 -> g_s ; s2 partial ; (isotropic) lengthscale partial
        float4 uq6[8];
        if (!WHITE) {
#pragma unroll
            for (int e = 0; e < 8; ++e) uq6[e] = load_u4(c_lo + 4 * e);      // in flight while G7 runs
        }
        mbar_wait(bar_acc7, 0);
        tc_fence_after();
        BSTAMP();   // G7 done
        float s2 = 0.f, lsum = 0.f;
        const float inv_var = 1.0f / var0;
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
            if (c8 < nch) {
                const int c0 = c_lo + 8 * c8;
                float w[8], r2[8];
                __syncwarp();
                tmem_ld8(lane_addr + 128 + c0, w);
                __syncwarp();
                tmem_ld8(lane_addr + 384 + c0, r2);
                if (valid) {
                    if (c0 + 8 <= M) {
                        float4* dst = reinterpret_cast<float4*>(a.W + (size_t)row * M + c0);
                        dst[0] = make_float4(w[0], w[1], w[2], w[3]);
                        dst[1] = make_float4(w[4], w[5], w[6], w[7]);
                    } else {
#pragma unroll
                        for (int u = 0; u < 8; ++u) if (c0 + u < M) a.W[(size_t)row * M + c0 + u] = w[u];
                    }
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int i = c0 + u;
                    if (i < M) {
                        float kb_ = w[u];
                        if (!WHITE) {
                            const float4 uu = uq6[2 * c8 + (u >> 2)];
                            const float ue = (u & 3) == 0 ? uu.x : (u & 3) == 1 ? uu.y : (u & 3) == 2 ? uu.z : uu.w;
                            kb_ -= vs * ue;
                        }
                        float k, kp;
                        kern_eval_fast(KERN, r2[u], var0, k, kp);
                        s2 = fmaf(kb_ * k, inv_var, s2);
                        const float g = 2.f * kb_ * kp;
                        lsum = fmaf(g, r2[u], lsum);
                        w[u] = g;
                    } else w[u] = 0.f;
                }
                // (M % 4 == 0: whole float4s are either inside the row or outside it)
                if (c0 + 4 <= M) *reinterpret_cast<float4*>(g_s + t * MP + c0) = make_float4(w[0], w[1], w[2], w[3]);
                if (c0 + 8 <= M) *reinterpret_cast<float4*>(g_s + t * MP + c0 + 4) = make_float4(w[4], w[5], w[6], w[7]);
            }
        }
        BSTAMP();   // R6 done
        if (qt == 0) s2 += vs;                    // d Kdiag / d variance = 1
        s2 = warp_sum(s2);
        lsum = warp_sum(lsum);
        if (lane == 0) { red_s[warp] = s2; red_s[16 + warp] = lsum; }
        if (qt == 0 && P.kwhite) {                // the same sum is the gradient of a White term's variance (Kdiag += wvar)
            const float sw = warp_sum(vs);
            if (lane == 0) atomicAdd(P.gwvar, sw);
        }
        if (threadIdx.x < 32) red_s[32 + threadIdx.x] = 0.f;     // ARD lengthscale accumulators
        named_bar_sync(1, TC_ROWTHREADS);                        // g_s, red_s complete; A_u is dead (G7 has completed)
        if (threadIdx.x == 0) {
            float tot = 0.f, ltot = 0.f;
            for (int w8 = 0; w8 < TC_ROWTHREADS / 32; ++w8) { tot += red_s[w8]; ltot += red_s[16 + w8]; }
            atomicAdd(P.gvar, tot);
            // isotropic lengthscale: d r2 / d l = -2 r2 / l  ->  d/dl = -(1/l) sum g r2
            if (!P.ard) atomicAdd(P.gls, -ltot * il_s[0]);
        }
        // ---- R7a: xbar (this quarter: q = qt, qt+4, ...):  xbar_q = (1/l_q) sum_i g_i (xs_q - zs_iq) + mean-function term
        if (a.xbar && valid) {
            constexpr int NQ = (DINP + 3) / 4;
            float accq[NQ], xq[NQ];
#pragma unroll
            for (int j = 0; j < NQ; ++j) { accq[j] = 0.f; xq[j] = xs_s[t * Din + min(qt + 4 * j, Din - 1)]; }
#pragma unroll 2
            for (int i4 = 0; i4 < M; i4 += 4) {
                const float4 g4 = *reinterpret_cast<const float4*>(g_s + t * MP + i4);
#pragma unroll
                for (int j = 0; j < NQ; ++j) {
                    const int q = min(qt + 4 * j, DINP - 1);
                    const float4 z4 = *reinterpret_cast<const float4*>(ZsT + q * M4 + i4);
                    accq[j] = fmaf(g4.x, xq[j] - z4.x, accq[j]); accq[j] = fmaf(g4.y, xq[j] - z4.y, accq[j]);
                    accq[j] = fmaf(g4.z, xq[j] - z4.z, accq[j]); accq[j] = fmaf(g4.w, xq[j] - z4.w, accq[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < NQ; ++j) {
                const int q = qt + 4 * j;
                if (q < Din) {
                    float s = accq[j] * il_s[q];
                    if (P.mean == DSDGP_MEAN_IDENTITY) {
#pragma unroll
                        for (int d = 0; d < DOUTP; ++d) if (d == q) s += mub[d];
                    } else if (P.mean == DSDGP_MEAN_LINEAR) {
#pragma unroll
                        for (int d = 0; d < DOUTP; ++d) s = fmaf(mub[d], __ldg(&P.meanW[q * D + d]), s);
                    }
                    a.xbar[(size_t)row * Din + q] = s;
                }
            }
        }
        BSTAMP();   // R7a done
        // ---- R7b: Z / (ARD) lengthscale partials: thread (i, rh) walks rows [32 rh, 32 rh + 32); the four row quarters are
        // reduced in shared memory so that a tile issues ONE global atomic per (i, q) (round 1: four; the same M*Din addresses
        // are hit by every tile of the layer, and the contention showed up as the tile's slowest phase)
        {
            const int i = threadIdx.x & 127, rh = threadIdx.x >> 7;      // inducing point, row quarter
            float sa[DINP], sb[DINP];
#pragma unroll
            for (int q = 0; q < DINP; ++q) { sa[q] = 0.f; sb[q] = 0.f; }
            if (i < M) {
                float zi[DINP];
#pragma unroll
                for (int q = 0; q < DINP; ++q) zi[q] = Zs[i * Din + q];
                if (P.ard) {
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
                } else {
                    // sum_r g_r (xs_rq - zs_iq) = sum_r g_r xs_rq - zs_iq sum_r g_r
                    float gs = 0.f;
#pragma unroll 4
                    for (int r = rh * 32; r < rh * 32 + 32; ++r) {
                        const float g = g_s[r * MP + i];
                        const float4* xr = reinterpret_cast<const float4*>(xs_s + r * DINP);
                        gs += g;
#pragma unroll
                        for (int q4 = 0; q4 < DINP / 4; ++q4) {
                            const float4 xv = xr[q4];
                            sa[4 * q4] = fmaf(g, xv.x, sa[4 * q4]); sa[4 * q4 + 1] = fmaf(g, xv.y, sa[4 * q4 + 1]);
                            sa[4 * q4 + 2] = fmaf(g, xv.z, sa[4 * q4 + 2]); sa[4 * q4 + 3] = fmaf(g, xv.w, sa[4 * q4 + 3]);
                        }
                    }
#pragma unroll
                    for (int q = 0; q < DINP; ++q) sa[q] = fmaf(-gs, zi[q], sa[q]);
                }
#pragma unroll
                for (int q = 0; q < DINP; ++q) zred_s[(rh * M + i) * Din + q] = sa[q];
            }
            if (P.ard) {
                // one shared atomic per warp and input dimension
#pragma unroll
                for (int q = 0; q < DINP; ++q) {
                    const float tsum = warp_sum(sb[q]);
                    if (lane == 0) atomicAdd(&red_s[32 + q], tsum);
                }
            }
        }
        BSTAMP();   // R7b done
        named_bar_sync(1, TC_ROWTHREADS);
        for (int e = threadIdx.x; e < M * Din; e += TC_ROWTHREADS) {
            const float il = il_s[e % Din];
            const float v = (zred_s[e] + zred_s[M * Din + e]) + (zred_s[2 * M * Din + e] + zred_s[3 * M * Din + e]);
            atomicAdd(&P.gZ[e], -v * il);
        }
        if (P.ard && threadIdx.x < Din) atomicAdd(&P.gls[threadIdx.x], -red_s[32 + threadIdx.x] * il_s[threadIdx.x]);
    }
    tc_fence_before();
    __syncthreads();
    if (a.tile_done && threadIdx.x == 0) {      // xbar / mubar / vbar / W of this tile are written (barrier above): publish
        __threadfence();
        st_release_gpu(a.tile_done + blockIdx.x, a.sa->epoch);
    }
    if (warp == TC_WARP_TMA) { __syncwarp(); tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// (DINP, DOUTP, KERN, WHITE) instances
#define TC_BWD_INSTANCES(X) \
    X(8, 1, 0, false) X(8, 1, 0, true) X(8, 1, 1, false) X(8, 1, 1, true) \
    X(8, 8, 0, false) X(8, 8, 0, true) X(8, 8, 1, false) X(8, 8, 1, true)

bool tc_bwd_supported(const LayerDev& P) {
    if (!(P.M <= 128 && P.M >= 8 && (P.M & 3) == 0 && P.wpack_fwd != nullptr && P.ipd == 0)) return false;
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

void launch_bwd_rows_tc(const LayerDev& P, const BwdArgs& a, cudaStream_t st, long long* nl, bool programmatic) {
    int grid = (a.R + TC_ROWS - 1) / TC_ROWS;
    size_t sm = bwd_smem_plan(P.M, P.Din, P.Dout).total + 1024;
    const bool wh = P.white != 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = sm; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = programmatic ? 1 : 0;
#define X(a_, b_, k_, w_) if (P.Din == a_ && P.Dout == b_ && P.kern == k_ && wh == w_) cudaLaunchKernelEx(&cfg, k_layer_bwd_tc<a_, b_, k_, w_>, P, a);
    TC_BWD_INSTANCES(X)
#undef X
    *nl += 1;
}
