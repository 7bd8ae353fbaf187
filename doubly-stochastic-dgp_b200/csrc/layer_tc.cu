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
#include "tc_pack.cuh"

#define TC_ROWS 128
#define TC_NS_MAX 10         // weight ring sub-slots (barriers)
#define TC_CHUNK_BYTES 16384 // one [128 rows] x [32 k] activation k-block
#define TC_THREADS 576

// ----------------------------------------------------------------------------------------------
// weight packing (layout: tc_pack.cuh).  One thread per stored element; hi = tf32(x), lo = tf32(x - hi).
// ----------------------------------------------------------------------------------------------
// part 0: every block; 1: the four Linv blocks (need this step's factorisation); 2: the q_sqrt blocks (parameters only --
// packed on the side branch of the step DAG while the factorisation runs): G2 (L_d^T, forward) and S_d = L_d L_d^T (backward)
__global__ void k_pack_fwd(LayerSet ls, int part, Accum* acc) {
    const LayerDev& P = ls.l[blockIdx.y];
    if (!P.wpack_fwd) return;
    const int M = P.M, D = P.Dout, nkb = tcp::nkb_of(M), NPAD = tcp::npad_of(M);
    const uint32_t slot = tcp::slot_bytes(M);
    const int nblk = tcp::num_blocks(D) + D;            // triangular blocks, then the D square S_d operands
    const float g2thr = 1e-3f * sqrtf(P.var[0]);
    const int per_blk = nkb * 128 * 32;                 // (kb, n, kk) index space; rows outside a band are skipped
    const int blk0 = part == 2 ? 4 : 0, blk1 = part == 1 ? 4 : nblk;
    const size_t total = (size_t)(blk1 - blk0) * per_blk;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int blk = blk0 + (int)(e / per_blk), w = (int)(e % per_blk), kb = w >> 12, n = (w >> 5) & 127, kk = w & 31, k = kb * 32 + kk;
        if (blk >= 4 + 2 * D) {
            // S_d[n][k] = sum_{j <= min(n,k)} L_d[n][j] L_d[k][j]  (symmetric: B[n][k] = S_d[k][n] = S_d[n][k])
            if (n >= NPAD) continue;
            const int d = blk - 4 - 2 * D;
            float out = 0.f;
            if (n < M && k < M) {
                const float* Ln = P.q_sqrt + ((size_t)d * M + n) * M;
                const float* Lk = P.q_sqrt + ((size_t)d * M + k) * M;
                const int jm = min(n, k);
                float s0 = 0.f, s1 = 0.f;
                int j = 0;
                for (; j + 1 <= jm; j += 2) { s0 = fmaf(Ln[j], Lk[j], s0); s1 = fmaf(Ln[j + 1], Lk[j + 1], s1); }
                if (j <= jm) s0 = fmaf(Ln[j], Lk[j], s0);
                out = tc::tf32_rna(s0 + s1);
            }
            const bool t32 = tcp::is_tail32(M, kb);
            if (t32 && kk >= 8) continue;
            char* dst = reinterpret_cast<char*>(P.wpack_fwd) + tcp::s_region_offset(M, D) + (size_t)d * tcp::sfull_bytes(M) +
                        (size_t)kb * tcp::sfull_band_bytes(M) + (t32 ? tcp::sw32_offset(n, kk) : tc::sw128_offset(n, kk));
            *reinterpret_cast<float*>(dst) = out;
            continue;
        }
        const int pat = blk < 2 ? tcp::PAT_LE : tcp::PAT_GE;
        const int r0 = tcp::band_row0(pat, kb), nr = tcp::band_rows(pat, M, kb);
        if (n < r0 || n >= r0 + nr) continue;
        float out = 0.f;
        if (n < M && k < M) {
            if (blk < 4) {
                const int part = blk & 1;
                const double v = (blk < 2) ? P.Linv64[(size_t)n * M + k] : P.Linv64[(size_t)k * M + n];   // G1: Linv[n][k] ; G1': Linv[k][n]
                const float hi = tc::tf32_rna((float)v);
                out = part ? tc::tf32_rna((float)(v - (double)hi)) : hi;
            } else {
                const int lo = blk >= 4 + D, d = blk - 4 - (lo ? D : 0);
                if (n <= k) {                                                                               // G2: L_d[k][n]
                    const float v = P.q_sqrt[((size_t)d * M + k) * M + n];
                    const float hi = tc::tf32_rna(v);
                    out = lo ? tc::tf32_rna(v - hi) : hi;
                    if (!lo && fabsf(v) > g2thr && acc) atomicOr(&acc->g2flag[P.idx], 1);
                }
            }
        }
        const bool t32 = tcp::is_tail32(M, kb);
        if (t32 && kk >= 8) continue;
        char* dst = reinterpret_cast<char*>(P.wpack_fwd) + (size_t)blk * slot + tcp::band_offset(pat, M, kb) +
                    (t32 ? tcp::sw32_offset(n - r0, kk) : tc::sw128_offset(n - r0, kk));
        *reinterpret_cast<float*>(dst) = out;
    }
}

void launch_pack_fwd(const LayerSet& ls, int part, Accum* acc, cudaStream_t st, long long* nl) {
    k_pack_fwd<<<dim3(part == 1 ? 24 : 96, ls.L), 256, 0, st>>>(ls, part, acc);
    *nl += 1;
}

// |c_d|^2 scratch in its own [D][128] region (then G2 may run as 3xTF32) iff everything still fits into 227 KB
__host__ __device__ inline bool tc_fwd_csq_dedicated(int M, int Din, int D) {
    const size_t base = 1024 + 131072 + 2 * (size_t)tcp::slot_bytes(M) + 256 + 2048 + sizeof(float) * ((size_t)M * Din + (size_t)M * D + 8);
    return base + sizeof(float) * 128 * (size_t)D <= 226 * 1024;
}

// ----------------------------------------------------------------------------------------------
// forward kernel.  576 threads: warps 0-15 = row warps (FOUR threads per row: threads t, t+128, t+256, t+384 share TMEM
// lane t&127; a warp may only touch the TMEM lane quadrant warp%4, which is the quadrant of its rows), warp 16 = TMA
// producer, warp 17 = MMA issue.  A tile is a serial chain of short SIMT phases between MMAs, so the phases are latency
// bound: 16 row warps (4 per scheduler) instead of 8 roughly halve every phase (profiles/r2_*).  Work split of a row's
// four threads ("quarters" q = 0..3):
//   Gram, E1 (b), E1' (u)   : columns / inducing points split four ways
//   E2 (|c_d|^2)            : quarter pair p = q>>1 owns the outputs d = p, p+2, ... (= TMEM accumulator buffer p), the
//                             two quarters of a pair split the columns in halves -> two partial sums per (row, d)
//   mean, draw, stores      : quarter q owns outputs d = 2q, 2q+1 (+8k): complete dot products, no partial sums, and one
//                             Philox block + one Box-Muller pair yields both draws
// ----------------------------------------------------------------------------------------------
#define TC_ROWTHREADS 512
#define TC_WARP_TMA 16
#define TC_WARP_MMA 17
// One 128-row tile of one layer.  `tile` = row-tile index; tmem = base of this CTA's 512 TMEM columns; reinit = the
// mbarriers were used by a previous tile of this CTA (persistent chain kernel) and must be invalidated first.
template <int DINP, int DOUTP>
__device__ __forceinline__ void fwd_tile_body(const LayerDev& P, const FwdArgs& a, const int tile, const uint32_t sbase,
                                              uint8_t* sgen, const uint32_t tmem, const bool reinit) {
    using namespace tc;
    const uint32_t A_hi = sbase, A_lo = sbase + 65536, Bring = sbase + 131072;
    const uint32_t slotb = tcp::slot_bytes(P.M);
    const uint32_t misc = Bring + 2 * slotb;
    // Weights stream as 32-wide k-block bands (<= 128 NPAD bytes, one TMA bulk copy each) through a ring of sub-slots: the
    // ring area holds `nr` of them; once the 3xTF32 projections are done A_lo is dead and adds `nalo` more for the G2 stream
    // (its last bytes keep the |c_d|^2 scratch), so up to nr + nalo bands are in flight instead of two or three whole blocks.
    // The |c_d|^2 partial sums live in a dedicated [D][128] region (the two halves of a pair add with shared-memory atomics)
    // when it fits next to everything else, else as [2][D][128] partials at the end of A_lo.  Only with the dedicated region
    // can A_lo keep u's low part, i.e. can G2 run as 3xTF32 (g2x3): the flag comes from k_pack_fwd (q_sqrt not negligible).
    const uint32_t band_full = 128u * (uint32_t)tcp::npad_of(P.M);
    const bool csq_ded = tc_fwd_csq_dedicated(P.M, P.Din, P.Dout);
    const uint32_t csq_bytes = csq_ded ? 0u : 2u * (uint32_t)P.Dout * 128u * 4u;
    // G2 precision (option "g2_passes"; 0 = automatic: 2 when k_pack_fwd flagged the layer's q_sqrt as non-negligible, else 1):
    //   1: u_hi L_hi                      2: u_hi (L_hi + L_lo)                     3: (u_hi + u_lo) L_hi + u_hi L_lo
    // Measured (tools/g2_passes_experiment.py): the ELBO error of (1) comes from rounding the WEIGHTS -- the same perturbation for
    // every row, so it does not average out over the minibatch -- while u's rounding is independent per row and does; (2) is
    // within 2x of (3) at the cost of the doubled weight stream only (no second operand image, A_lo stays a ring area).
    const int g2p = a.g2_passes ? a.g2_passes : ((a.acc && a.acc->g2flag[P.idx] != 0) ? 2 : 1);
    const bool g2x3 = csq_ded && g2p == 3;            // u's low part in A_lo
    const bool g2lo = g2p == 2 || g2x3;               // stream the low parts of L_d too
    const int nr = (int)(2 * slotb / band_full);
    const int nalo = (g2x3 || csq_bytes >= 65536u) ? 0 : min(TC_NS_MAX - nr, (int)((65536u - csq_bytes) / band_full));
    auto slot_addr = [&](int sl) { return sl < nr ? Bring + (uint32_t)sl * band_full : A_lo + (uint32_t)(sl - nr) * band_full; };
    const uint32_t bar_full = misc, bar_empty = misc + 8 * TC_NS_MAX;
    const uint32_t bar_a = bar_empty + 8 * TC_NS_MAX;        // a_ready[3]
    const uint32_t bar_acc = bar_a + 24;                      // acc_full[2]
    const uint32_t bar_acc2f = bar_acc + 16;                  // acc2_full[2]
    const uint32_t bar_acc2e = bar_acc2f + 16;                // acc2_empty[2]
    float* part_s = reinterpret_cast<float*>(sgen + (misc + 256 - sbase));        // [4][128] |b|^2 partials
    float* Zs = part_s + 512;                                                       // [M][Din], pre-scaled by 1/lengthscale
    float* qmu_s = Zs + ((P.M * P.Din + 3) & ~3);                                   // [M][D], 16-byte aligned
    // scratch aliased on A_lo once it is dead (G2 reads A_hi only): |c_d|^2 partials [2][D][128]
    float* csq_p = csq_ded ? qmu_s + ((P.M * P.Dout + 3) & ~3)
                           : reinterpret_cast<float*>(sgen + (A_lo - sbase) + (65536u - min(csq_bytes, 65536u)));

    const int M = P.M, Din = P.Din, D = P.Dout;
    const int nkb = (M + 31) / 32, NPAD = (M + 15) & ~15;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = tile * TC_ROWS, R = a.R;

    // the tile's inputs are requested before anything else: the round trip (~2k cycles) overlaps the barrier set-up and the
    // Z / q_mu staging below instead of heading the Gram phase
    float xraw[DINP];
#pragma unroll
    for (int q = 0; q < DINP; ++q) xraw[q] = 0.f;
    if (warp < TC_WARP_TMA && row0 + (int)(threadIdx.x & 127) < R) {
        const int r = row0 + (int)(threadIdx.x & 127);
        if (a.fold_mean) {
            // deferred layer-1 fold (see FwdArgs): every quarter thread of the row needs all Din inputs and draws them itself
            // (the four quarters of a row sit in different warps); quarter qt stores pair qt, qt + 4, ... of the row
            const int s_idx = r / a.N, n = r - s_idx * a.N, myq = (int)(threadIdx.x >> 7);
            const unsigned long long seed = a.sa->seed;
            const int noff = a.sa->n_offset, soff = a.sa->s_offset;
#pragma unroll
            for (int q = 0; q < DINP; q += 2) {
                if (q < Din) {
                    const int nd = (q + 1 < Din) ? 2 : 1;
                    float z2[2], m2[2] = {0.f, 0.f}, sd2[2] = {0.f, 0.f};
                    for (int e = 0; e < nd; ++e) {
                        m2[e] = __ldcg(&a.fold_mean[(size_t)n * Din + q + e]);
                        sd2[e] = sqrtf(fmaxf(__ldcg(&a.fold_var[(size_t)n * Din + q + e]) + a.jitter, 1e-30f));
                    }
                    const size_t o = (size_t)r * Din + q;
                    if (a.fold_z) { z2[0] = a.fold_z[o]; z2[1] = nd == 2 ? a.fold_z[o + 1] : 0.f; }
                    else dsdgp_normal2(seed, a.fold_layer, s_idx + soff, n + noff, q, z2[0], z2[1]);
                    const bool mine = ((q >> 1) & 3) == myq;
                    for (int e = 0; e < nd; ++e) {
                        const float x = fmaf(z2[e], sd2[e], m2[e]);
                        xraw[q + e] = x;
                        if (mine) {
                            a.fold_F[o + e] = x;
                            if (a.fold_zout) a.fold_zout[o + e] = z2[e];
                        }
                    }
                }
            }
        } else {
#pragma unroll
            for (int q = 0; q < DINP; ++q)
                if (q < Din) xraw[q] = __ldcg(&a.Xin[(size_t)r * Din + q]);
        }
    }
    if (threadIdx.x == 0) {
        if (reinit)
            for (int i = 0; i < 2 * TC_NS_MAX + 9; ++i) mbar_inval(misc + 8 * i);
        for (int s = 0; s < TC_NS_MAX; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int i = 0; i < 3; ++i) mbar_init(bar_a + 8 * i, TC_ROWTHREADS);
        for (int i = 0; i < 2; ++i) { mbar_init(bar_acc + 8 * i, 1); mbar_init(bar_acc2f + 8 * i, 1); mbar_init(bar_acc2e + 8 * i, TC_ROWTHREADS / 2); }
        fence_mbar_init();
    }
    for (int e = threadIdx.x; e < M * Din; e += TC_THREADS) Zs[e] = P.Z[e] * (1.0f / P.ls[P.ard ? e % Din : 0]);
    for (int e = threadIdx.x; e < M * D; e += TC_THREADS) qmu_s[e] = P.q_mu[e];
    if (csq_ded)
        for (int e = threadIdx.x; e < D * 128; e += TC_THREADS) csq_p[e] = 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();


    if (warp == TC_WARP_TMA) {
        // ===================== TMA producer (one lane): one bulk copy per band, in the order the MMA warp consumes them =====
        if (lane == 0) {
            const char* wsrc = reinterpret_cast<const char*>(P.wpack_fwd);
            uint32_t par = 0;          // bit s: parity of the next use of sub-slot s
            int s = 0;
            bool alo_ok = false;
            auto load_block = [&](int blk, int pat, int nslots) {
                for (int q = 0; q < nkb; ++q) {
                    const int kb = pat == tcp::PAT_GE ? nkb - 1 - q : q;
                    const uint32_t bytes = tcp::band_tx_bytes(pat, M, kb);
                    if (s >= nr && !alo_ok) {      // first band into A_lo: the projections must have finished reading it
                        mbar_wait(P.white ? bar_acc : bar_acc + 8, 0);
                        alo_ok = true;
                    }
                    mbar_wait(bar_empty + 8 * s, ((par >> s) & 1u) ^ 1u);
                    mbar_arrive_expect_tx(bar_full + 8 * s, bytes);
                    tma_bulk_g2s(slot_addr(s), wsrc + (size_t)blk * slotb + tcp::band_offset(pat, M, kb), bytes, bar_full + 8 * s);
                    par ^= 1u << s;
                    if (++s == nslots) s = 0;
                }
            };
            load_block(tcp::blk_g1(0), tcp::PAT_LE, nr); load_block(tcp::blk_g1(1), tcp::PAT_LE, nr);
            if (!P.white) { load_block(tcp::blk_g1p(0), tcp::PAT_GE, nr); load_block(tcp::blk_g1p(1), tcp::PAT_GE, nr); }
            s = 0;
            for (int d = 0; d < D; ++d) {
                load_block(tcp::blk_g2(d), tcp::PAT_GE, nr + nalo);
                if (g2lo) load_block(tcp::blk_g2lo(D, d), tcp::PAT_GE, nr + nalo);
            }
        }
    } else if (warp == TC_WARP_MMA) {
        // ===================== MMA issuer: whole warp runs the uniform control flow, one elected lane issues ==========
        {
            uint32_t par = 0;
            int s = 0;
            const uint64_t desc_hi = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
            auto mkdesc = [&](uint32_t addr) { return desc_hi | (uint64_t)(((addr >> 4) & 0x3FFF) | (1u << 16)); };
            // tail band (tc_pack.cuh tail32): SWIZZLE_32B K-major, 8-row atoms 256 B apart
            const uint64_t desc32_hi = ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)6 << 61);
            auto mkdesc32 = [&](uint32_t addr) { return desc32_hi | (uint64_t)(((addr >> 4) & 0x3FFF) | (1u << 16)); };
            // one band block: D (+)= A * B^T, band by band.  mode 0: A_hi and A_lo against this block (B_hi of a 3xTF32
            // product), 1: A_hi only, accumulating (B_lo), 2: A_hi only (1xTF32 product, fresh accumulator).  In both band
            // orders the first band of a block spans all NPAD accumulator columns, so it is the one that clears them.
            auto do_block = [&](uint32_t dcol, int pat, int mode, int nslots) {
#pragma unroll 1
                for (int q = 0; q < nkb; ++q) {
                    const int kb = pat == tcp::PAT_GE ? nkb - 1 - q : q;
                    mbar_wait(bar_full + 8 * s, (par >> s) & 1u);
                    tc_fence_after();
                    const int nks = min(4, (M - 32 * kb + 7) / 8);
                    const int nrows = tcp::band_rows(pat, M, kb);
                    const uint32_t bbase = slot_addr(s), abase = kb * TC_CHUNK_BYTES;
                    const uint32_t id = make_idesc_tf32(128, nrows);
                    const uint32_t dc = tmem + dcol + (uint32_t)tcp::band_row0(pat, kb);
                    const bool t32 = tcp::is_tail32(M, kb);          // (then nks == 1)
                    if (elect_one()) {
#pragma unroll 1
                        for (int ks = 0; ks < nks; ++ks) {
                            const uint64_t bd = t32 ? mkdesc32(bbase) : mkdesc(bbase + ks * 32), ah = mkdesc(A_hi + abase + ks * 32);
                            mma_tf32(dc, ah, bd, id, (mode != 1 && q == 0 && ks == 0) ? 0u : 1u);
                            if (mode == 0) mma_tf32(dc, mkdesc(A_lo + abase + ks * 32), bd, id, 1u);
                        }
                        mma_commit(bar_empty + 8 * s);
                    }
                    __syncwarp();
                    par ^= 1u << s;
                    if (++s == nslots) s = 0;
                }
            };
            auto commit = [&](uint32_t bar) { if (elect_one()) mma_commit(bar); __syncwarp(); };
            // G1: b = Linv k   (3xTF32)
            mbar_wait(bar_a, 0);
            tc_fence_after();
            do_block(0u, tcp::PAT_LE, 0, nr);
            do_block(0u, tcp::PAT_LE, 1, nr);
            commit(bar_acc);
            if (!P.white) {
                // G1': u = Linv^T b   (3xTF32)
                mbar_wait(bar_a + 8, 0);
                tc_fence_after();
                do_block(128u, tcp::PAT_GE, 0, nr);
                do_block(128u, tcp::PAT_GE, 1, nr);
                commit(bar_acc + 8);
            }
            // G2: c_d = L_d^T u   (1xTF32), accumulators double-buffered (buffer d&1 is consumed by quarter pair d&1)
            mbar_wait(bar_a + 16, 0);
            tc_fence_after();
            s = 0;
            for (int d = 0; d < D; ++d) {
                if (d >= 2) { mbar_wait(bar_acc2e + 8 * (d & 1), ((d >> 1) - 1) & 1); tc_fence_after(); }
                // hi block: u_hi (and u_lo when g2x3) against L_hi; lo block (g2lo): u_hi against L_lo.  The band sequence must
                // be exactly the producer's (every band it loads is consumed here)
                do_block(256u + 128u * (uint32_t)(d & 1), tcp::PAT_GE, g2x3 ? 0 : 2, nr + nalo);
                if (g2lo) do_block(256u + 128u * (uint32_t)(d & 1), tcp::PAT_GE, 1, nr + nalo);
                commit(bar_acc2f + 8 * (d & 1));
            }
        }
    } else {
        // ===================== row warps =====================
        const int t = threadIdx.x & 127, qt = threadIdx.x >> 7, row = row0 + t;       // qt: quarter 0..3
        const int pair = qt >> 1, half = qt & 1;
        const bool dbg = a.dbg && tile == 0 && threadIdx.x == 0;
        int dbi = 0;
#define STAMP() do { if (dbg) a.dbg[dbi++] = clock64(); } while (0)
        STAMP();
        const bool valid = row < R;
        const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        const uint32_t rsw = (uint32_t)(t & 7);
        const uint32_t rowoff = (uint32_t)((t >> 3) * 1024 + (t & 7) * 128);
        auto a_store4 = [&](uint32_t base, int k4, float4 v) {      // k4: first of 4 consecutive k (multiple of 4)
            uint32_t off = (uint32_t)(k4 >> 5) * TC_CHUNK_BYTES + rowoff + (((uint32_t)((k4 & 31) >> 2) ^ rsw) << 4);
            *reinterpret_cast<float4*>(sgen + (base - sbase) + off) = v;
        };
        auto split_store = [&](int k4, const float* v, bool write_lo) {
            float4 hi, lo;
            hi.x = tf32_rna(v[0]); hi.y = tf32_rna(v[1]); hi.z = tf32_rna(v[2]); hi.w = tf32_rna(v[3]);
            a_store4(A_hi, k4, hi);
            if (write_lo) {
                lo.x = tf32_rna(v[0] - hi.x); lo.y = tf32_rna(v[1] - hi.y); lo.z = tf32_rna(v[2] - hi.z); lo.w = tf32_rna(v[3] - hi.w);
                a_store4(A_lo, k4, lo);
            }
        };
        // column ranges (multiples of 8): this quarter's [c_lo, c_hi) of the NPAD accumulator columns, and this thread's
        // half [h_lo, h_hi) inside its pair (E2)
        const int nch8 = NPAD >> 3;
        const int cq = nch8 >> 2, cr = nch8 & 3;
        const int c_lo = 8 * (qt * cq + min(qt, cr)), c_hi = c_lo + 8 * (cq + (qt < cr ? 1 : 0));
        const int NH = ((NPAD >> 1) + 7) & ~7;
        const int h_lo = half ? NH : 0, h_hi = half ? NPAD : NH;

        // ---- Gram: k_i = k(z_i, x) -> A_hi / A_lo (tf32 split); the quarters take alternate groups of 4 inducing points.
        // Inputs and inducing points are pre-scaled by 1/lengthscale (Zs above).
        float xs[DINP];
#pragma unroll
        for (int q = 0; q < DINP; ++q)
            xs[q] = q < Din ? xraw[q] * (1.0f / P.ls[P.ard ? q : 0]) : 0.f;
        const float var0 = P.var[0];
        const bool rbf = P.kern == DSDGP_KERN_RBF;
        const float l2var = log2f(var0);
        for (int i4 = 4 * qt; i4 < nkb * 32; i4 += 16) {
            float kv[4] = {0.f, 0.f, 0.f, 0.f};
            if (i4 < M) {
                float r2[4];
                if (Din == DINP) {
                    // vector loads of the inducing inputs (row i = DINP consecutive floats, 16-byte aligned)
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int i = min(i4 + u, M - 1);
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
                } else {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int i = min(i4 + u, M - 1);
                        float s = 0.f;
#pragma unroll
                        for (int q = 0; q < DINP; ++q) {
                            if (q < Din) {
                                const float dd = xs[q] - Zs[i * Din + q];
                                s = fmaf(dd, dd, s);
                            }
                        }
                        r2[u] = s;
                    }
                }
                if (rbf) {
                    // var * exp(-r2/2) = 2^(log2(var) - r2 * log2(e)/2): one FMA + ex2.approx per point
#pragma unroll
                    for (int u = 0; u < 4; ++u) kv[u] = (i4 + u < M) ? fast_ex2(fmaf(r2[u], -0.72134752044448170368f, l2var)) : 0.f;
                } else {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        float k, kp;
                        kern_eval_fast(P.kern, r2[u], var0, k, kp);
                        kv[u] = (i4 + u < M) ? k : 0.f;
                    }
                }
            }
            split_store(i4, kv, true);
        }
        fence_proxy_async();
        mbar_arrive(bar_a);
        STAMP();      // 1: Gram done

        // ---- E1 (b) and E1' (u): this quarter's columns
        float bn = 0.f;
        auto consume = [&](uint32_t dcol, bool acc_bn, bool write_lo) {
            for (int c0 = c_lo; c0 < c_hi; c0 += 8) {
                float v[8];
                __syncwarp();
                tmem_ld8(lane_addr + dcol + c0, v);
                if (acc_bn) {
#pragma unroll
                    for (int u = 0; u < 8; ++u) bn = fmaf(v[u], v[u], bn);
                }
                split_store(c0, v, write_lo);
                split_store(c0 + 4, v + 4, write_lo);
            }
        };
        mbar_wait(bar_acc, 0);
        tc_fence_after();
        STAMP();      // 2: G1 accumulators ready
        if (P.white) {
            consume(0u, true, g2x3);
            tc_fence_before();
            fence_proxy_async();
            mbar_arrive(bar_a + 16);
        } else {
            consume(0u, true, true);
            tc_fence_before();
            fence_proxy_async();
            mbar_arrive(bar_a + 8);
            STAMP();  // 3: E1 done
            mbar_wait(bar_acc + 8, 0);
            tc_fence_after();
            STAMP();  // 4: G1' ready
            consume(128u, false, g2x3);        // critical path: u -> A_hi (and A_lo when G2 runs as 3xTF32)
            tc_fence_before();
            fence_proxy_async();
            mbar_arrive(bar_a + 16);
        }
        STAMP();      // 5: operands for G2 published
        part_s[qt * 128 + t] = bn;
        // off the critical path (interleaved with the G2 epilogues), re-reading u from TMEM: mean_d = u . q_mu[:, d] for the
        // outputs this quarter owns (d = 2 qt + (j & 1) + 8 (j >> 1)), and the U store of this quarter's columns
        constexpr int MD = DOUTP <= 8 ? 2 : DOUTP / 4;
        float meanv[MD];
#pragma unroll
        for (int j = 0; j < MD; ++j) meanv[j] = 0.f;
        const uint32_t ucol = P.white ? 0u : 128u;
        auto deferred = [&](int c0) {
            float v[8];
            __syncwarp();
            tmem_ld8(lane_addr + ucol + c0, v);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = c0 + u;
                if (i < M) {
                    if (DOUTP == 8 && D == 8) {
                        const float2 qv = *reinterpret_cast<const float2*>(qmu_s + i * 8 + 2 * qt);
                        meanv[0] = fmaf(v[u], qv.x, meanv[0]);
                        meanv[1] = fmaf(v[u], qv.y, meanv[1]);
                    } else {
#pragma unroll
                        for (int j = 0; j < MD; ++j) {
                            const int d = 2 * qt + (j & 1) + 8 * (j >> 1);
                            if (d < D) meanv[j] = fmaf(v[u], qmu_s[i * D + d], meanv[j]);
                        }
                    }
                }
            }
            if (valid && c0 >= c_lo && c0 < c_hi) {
                if (c0 + 8 <= M && (M & 3) == 0) {
                    float4* dst = reinterpret_cast<float4*>(a.U + (size_t)row * M + c0);
                    dst[0] = make_float4(v[0], v[1], v[2], v[3]);
                    dst[1] = make_float4(v[4], v[5], v[6], v[7]);
                } else {
#pragma unroll
                    for (int u = 0; u < 8; ++u) if (c0 + u < M) a.U[(size_t)row * M + c0 + u] = v[u];
                }
            }
        };

        // ---- E2: |c_d|^2 partials: pair `pair` consumes accumulator buffer `pair` (d = pair, pair+2, ...), this thread sums
        // its half of the columns
        const int nmine = (D - pair + 1) >> 1;            // outputs this pair consumes
        int kdef = 0;
        for (int d = pair; d < D; d += 2, ++kdef) {
            for (int j = kdef; j < nch8; j += nmine) deferred(8 * j);
            mbar_wait(bar_acc2f + 8 * pair, (d >> 1) & 1);
            tc_fence_after();
            STAMP();  // G2[d] ready
            float s = 0.f;
            {
                // this half's columns (<= 64) in two batches of up to four loads, one wait per batch
                float s0 = 0.f, s1 = 0.f;
#pragma unroll
                for (int b4 = 0; b4 < 2; ++b4) {
                    uint32_t r[4][8];
                    __syncwarp();
#pragma unroll
                    for (int c8 = 0; c8 < 4; ++c8)
                        if (h_lo + 8 * (4 * b4 + c8) < h_hi) tmem_ld8_nw(lane_addr + 256 + 128 * pair + h_lo + 8 * (4 * b4 + c8), r[c8]);
                    tmem_ld_wait();
#pragma unroll
                    for (int c8 = 0; c8 < 4; ++c8) {
                        if (h_lo + 8 * (4 * b4 + c8) < h_hi) {
#pragma unroll
                            for (int u = 0; u < 8; u += 2) {
                                const float a0 = __uint_as_float(r[c8][u]), a1 = __uint_as_float(r[c8][u + 1]);
                                s0 = fmaf(a0, a0, s0); s1 = fmaf(a1, a1, s1);
                            }
                        }
                    }
                }
                s = s0 + s1;
            }
            tc_fence_before();
            mbar_arrive(bar_acc2e + 8 * pair);
            if (csq_ded) atomicAdd(&csq_p[d * 128 + t], s);
            else csq_p[(half * D + d) * 128 + t] = s;
            STAMP();  // E2[d] done
        }
        if (nmine == 0)
            for (int j = 0; j < nch8; ++j) deferred(8 * j);
        named_bar_sync(1, TC_ROWTHREADS);
        STAMP();
        // ---- finalise: quarter qt handles outputs d0 = 2 qt + 8 k and d0 + 1
        const float jit = a.jitter;
        const unsigned long long seed = a.sa->seed;
        const int noff = a.sa->n_offset, soff = a.sa->s_offset;
        const float bnt = (part_s[t] + part_s[128 + t]) + (part_s[256 + t] + part_s[384 + t]);
        if (valid) {
#pragma unroll
            for (int jp = 0; jp < MD; jp += 2) {
                const int d0 = 2 * qt + 8 * (jp >> 1);
                if (d0 >= D) break;
                const int nd = (d0 + 1 < D) ? 2 : 1;
                float mean2[2] = {meanv[jp], meanv[jp + 1]}, sd2[2] = {0.f, 0.f};
                for (int e = 0; e < nd; ++e) {
                    const int d = d0 + e;
                    float mean = mean2[e];
                    if (P.mean == DSDGP_MEAN_IDENTITY) mean += __ldcg(&a.Xin[(size_t)row * Din + d]);
                    else if (P.mean == DSDGP_MEAN_LINEAR) {
                        float ms = P.meanB[d];
                        for (int q = 0; q < Din; ++q) ms = fmaf(__ldcg(&a.Xin[(size_t)row * Din + q]), __ldg(&P.meanW[q * D + d]), ms);
                        mean += ms;
                    }
                    const float csq = csq_ded ? csq_p[d * 128 + t] : csq_p[d * 128 + t] + csq_p[(D + d) * 128 + t];
                    const float v = (var0 + P.wvar[0]) - bnt + csq;      // Kdiag incl. a White term
                    a.Fmean[(size_t)row * D + d] = mean;
                    a.Fvar[(size_t)row * D + d] = v;
                    mean2[e] = mean;
                    sd2[e] = sqrtf(fmaxf(v + jit, 1e-30f));
                }
                if (a.F && a.S_rep > 1 && !a.z) {
                    const size_t o0 = (size_t)row * D + d0;
                    dsdgp_draw_fold(seed, P.idx, soff, row + noff, d0, nd, a.S_rep, (size_t)a.N * D, a.F + o0,
                                    a.z_out ? a.z_out + o0 : nullptr, mean2[0], mean2[1], sd2[0], sd2[1]);
                } else if (a.F) {
                    const int nrep = a.S_rep;
                    for (int ss = 0; ss < nrep; ++ss) {
                        // row r = s N + n (S_rep == 1), or layer-1 dedup: row = n, one draw per sample ss
                        const int sidx = nrep == 1 ? row / a.N : ss, n = nrep == 1 ? row % a.N : row;
                        const size_t o = (nrep == 1 ? (size_t)row : (size_t)ss * a.N + row) * D + d0;
                        float z2[2];
                        if (a.z) { z2[0] = a.z[o]; z2[1] = nd == 2 ? a.z[o + 1] : 0.f; }
                        else dsdgp_normal2(seed, P.idx, sidx + soff, n + noff, d0, z2[0], z2[1]);
                        for (int e = 0; e < nd; ++e) {
                            if (a.z_out) a.z_out[o + e] = z2[e];
                            a.F[o + e] = fmaf(z2[e], sd2[e], mean2[e]);
                        }
                    }
                }
            }
        }
    }
    if (a.dbg && tile == 0 && threadIdx.x == 0) a.dbg[40] = clock64();
    tc_fence_before();
    __syncthreads();
}

// ---- one layer, one tile per CTA
template <int DINP, int DOUTP>
__global__ void __launch_bounds__(TC_THREADS, 1) k_layer_fwd_tc(LayerDev P, FwdArgs a) {
    using namespace tc;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
    __shared__ uint32_t tmem_slot_s;
    const int warp = threadIdx.x >> 5;
    if (warp == TC_WARP_TMA) tmem_alloc(smem_u32(&tmem_slot_s), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot_s;
    fwd_tile_body<DINP, DOUTP>(P, a, blockIdx.x, sbase, sgen, tmem, false);
    if (warp == TC_WARP_TMA) { __syncwarp(); tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ---- all layers in ONE persistent kernel.  Tasks (layer l, tile t) are numbered layer-major; CTA c runs tasks
// c, c+grid, c+2*grid, ... in increasing order and, before a task, spins until the tile(s) of the previous layer it
// reads have been published (flag == epoch).  A task's dependency always has a smaller task number, and every CTA
// runs its tasks in increasing order, so the smallest unfinished task is always runnable: no deadlock as long as
// all CTAs are co-resident (grid <= number of SMs, 1 CTA/SM).  Rows are independent chains through the layers, so the
// layer boundary costs no grid-wide wave quantisation: ceil(tasks/grid) tile latencies instead of 2 per layer.
template <int DINP, int DOUTP>
__global__ void __launch_bounds__(TC_THREADS, 1) k_chain_fwd_tc(const __grid_constant__ LayerSet ls, const __grid_constant__ FwdChain fc) {
    using namespace tc;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
    __shared__ uint32_t tmem_slot_s;
    const int warp = threadIdx.x >> 5;
    // the likelihood and the last layer's backward rows are programmatic dependents that wait on this kernel's tile flags:
    // they take the SMs of the CTAs that have no task left in the last round
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (warp == TC_WARP_TMA) tmem_alloc(smem_u32(&tmem_slot_s), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot_s;
    const unsigned epoch = fc.sa->epoch;
    bool first = true;
    for (int q = blockIdx.x; q < fc.base[fc.L]; q += gridDim.x) {
        int l = 0;
        while (q >= fc.base[l + 1]) ++l;
        const int t = q - fc.base[l];
        if (l > 0 && threadIdx.x == 0) {
            // dependency: the producer tile(s) of layer l-1
            const unsigned* fl = fc.flags + (size_t)(l - 1) * fc.max_tiles;
            if (fc.a[l - 1].S_rep > 1 || fc.tiles[l - 1] != fc.tiles[l]) {
                for (int tt = 0; tt < fc.tiles[l - 1]; ++tt)
                    while (ld_acquire_gpu(fl + tt) != epoch) {}
            } else {
                while (ld_acquire_gpu(fl + t) != epoch) {}
            }
        }
        __syncthreads();
        fwd_tile_body<DINP, DOUTP>(ls.l[l], fc.a[l], t, sbase, sgen, tmem, !first);
        first = false;
        if (threadIdx.x == 0) {                 // publish (all threads' global stores precede the barrier above)
            __threadfence();
            st_release_gpu(fc.flags + (size_t)l * fc.max_tiles + t, epoch);
        }
    }
    if (warp == TC_WARP_TMA) { __syncwarp(); tc_fence_after(); tmem_dealloc(tmem, 512); }
}

static size_t tc_fwd_smem_base(int M, int Din, int D) { return 1024 + 131072 + 2 * (size_t)tcp::slot_bytes(M) + 256 + 2048 + sizeof(float) * ((size_t)M * Din + (size_t)M * D + 8); }
static size_t tc_fwd_smem(int M, int Din, int D) {
    return tc_fwd_smem_base(M, Din, D) + (tc_fwd_csq_dedicated(M, Din, D) ? sizeof(float) * 128 * (size_t)D : 0);
}

bool tc_fwd_supported(const LayerDev& P) {
    return P.M <= 128 && P.M >= 8 && P.Din <= 16 && P.Dout <= 32 && P.wpack_fwd != nullptr && P.ipd == 0 &&
           tc_fwd_smem(P.M, P.Din, P.Dout) <= 226 * 1024;
}

#define TC_FWD_INSTANCES(X) X(8, 1) X(8, 8) X(8, 32) X(16, 1) X(16, 8) X(16, 32)

cudaError_t layer_tc_init() {
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(k_chain_fwd_tc<8, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024))) return e;
#define X(a, b) if ((e = cudaFuncSetAttribute(k_layer_fwd_tc<a, b>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024))) return e;
    TC_FWD_INSTANCES(X)
#undef X
    return cudaSuccess;
}

size_t tc_fwd_pack_bytes(int M, int D, int white) {
    (void)white;
    return tcp::pack_bytes(M, D);
}

bool tc_chain_fwd_supported(const LayerSet& ls) {
    for (int l = 0; l < ls.L; ++l)
        if (!tc_fwd_supported(ls.l[l]) || ls.l[l].Din > 8 || ls.l[l].Dout > 8 || ls.l[l].M != ls.l[0].M) return false;
    return ls.L >= 2;      // (same M everywhere: the shared-memory carve-up, hence the mbarrier addresses, must not move)
}

// The chain kernel's CTAs wait on flags published by other CTAs of the same launch: every CTA must be resident at once.  The
// launch is therefore COOPERATIVE (the driver refuses it instead of letting it hang when the grid cannot be co-resident: MPS
// with a thread-percentage limit, green contexts, another long-lived kernel holding SMs) and the grid is sized from the
// occupancy query, not from the SM count alone.  Returns false if the launch was refused -- the caller falls back to one
// launch per layer.
bool launch_chain_fwd_tc(const LayerSet& ls, const FwdChain& fc, int num_sms, cudaStream_t st, long long* nl) {
    size_t sm = 0;
    for (int l = 0; l < ls.L; ++l) sm = max(sm, tc_fwd_smem(ls.l[l].M, ls.l[l].Din, ls.l[l].Dout));
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_chain_fwd_tc<8, 8>, TC_THREADS, sm) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        return false;
    }
    const int grid = min(num_sms * per_sm, fc.base[fc.L]);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = sm; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative; attr[0].val.cooperative = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, k_chain_fwd_tc<8, 8>, ls, fc) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    *nl += 1;
    return true;
}

void launch_fwd_tc(const LayerDev& P, const FwdArgs& a, cudaStream_t st, long long* nl) {
    int grid = (a.R + TC_ROWS - 1) / TC_ROWS;
    int dinp = P.Din <= 8 ? 8 : 16, doutp = P.Dout <= 1 ? 1 : P.Dout <= 8 ? 8 : 32;
#define X(a_, b_) if (dinp == a_ && doutp == b_) k_layer_fwd_tc<a_, b_><<<grid, TC_THREADS, tc_fwd_smem(P.M, P.Din, P.Dout), st>>>(P, a);
    TC_FWD_INSTANCES(X)
#undef X
    *nl += 1;
}
