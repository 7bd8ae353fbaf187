// tcgen05 row-reduction GEMMs (K = rows): the parameter-gradient accumulators of one layer,
//   P_d = sum_r vbar_rd u_r u_r^T (d < D) ;  G = sum_r w_r u_r^T ;  qmubar = sum_r u_r mubar_r^T .
// Out[i][j] = sum_r A[r][i] B[r][j]: both operands are "MN-major" for the tensor core.  For 32-bit operands the only
// MN-major shared-memory layout tcgen05 accepts is SWIZZLE_128B_BASE32B (cutlass sm100_common.inl:92): blocks of
// 32 features (128 B rows), 4 k-rows per swizzle atom, 32-byte granules XOR-ed with the k-row index (Swizzle<2,5,2>);
// the row tile is written as [feature block][128 rows][32 features] in that pattern; one UMMA k-step = 8 rows.
// Grid (1-D): output groups x row splits -- up to three P_d's (+ qmubar on the last) per P-group CTA (TMEM holds 512
// columns) and one group for G.  Each CTA keeps its accumulators in TMEM over all its row tiles and flushes once with
// vector reductions.  P_d and qmubar are 1xTF32 (gradients; tolerance in tests/test_gpu_parity.py); G feeds the Kuu
// adjoint whose kernel-hyper-parameter sums cancel heavily, so it is accumulated as 3xTF32: the G CTA runs
// W_hi U_hi + W_hi U_lo and then W_lo U_hi into ONE accumulator (two sub-iterations per row tile).
// Math: tests/algo_mirror.py::layer_bwdB.
#include "dsdgp_internal.cuh"
#include "tc_common.cuh"

#define RR_THREADS 544          // 16 row warps (four threads per row) + 1 MMA warp
#define RR_ROWTHREADS 512
#define RR_WARP_MMA 16
#define RR_TILE_BYTES 65536     // [4 blocks][128 rows][32 tf32]

namespace {
__device__ __forceinline__ bool elect_one_rr() {
    uint32_t p;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(p));
    return p != 0;
}
// MN-major SWIZZLE_128B_BASE32B descriptor (layout_type 1): 32 MN-elements (128 B) contiguous, 4 k-rows 128 B apart per
// atom, next group of 4 k-rows SBO = 512 B away, next 32-element MN block LBO bytes away.
__device__ __forceinline__ uint64_t make_desc_sw128_mnmajor(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((512 >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;
    return d;
}
}  // namespace

__global__ void __launch_bounds__(RR_THREADS, 1) k_layer_rowred_tc(LayerDev P, const float* __restrict__ U,
                                                                 const float* __restrict__ W, const float* __restrict__ mubar,
                                                                 const float* __restrict__ vbar, int R, int ngroups_d,
                                                                 int nsplit_p, int nsplit_g, long long* dbgp) {
    using namespace tc;
    extern __shared__ uint8_t smem_raw_r[];
    const uint32_t sbase = (smem_u32(smem_raw_r) + 1023u) & ~1023u;
    uint8_t* sgen = smem_raw_r + (sbase - smem_u32(smem_raw_r));
    const uint32_t A_t = sbase, B_t = sbase + RR_TILE_BYTES;           // B: two buffers
    const uint32_t bars = sbase + 3 * RR_TILE_BYTES;
    const uint32_t bar_aready = bars, bar_afree = bars + 8, bar_bready = bars + 16 /*[2]*/, bar_bfree = bars + 32 /*[2]*/,
                   bar_done = bars + 48, tmem_slot = bars + 56;
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(sgen + 3 * RR_TILE_BYTES + 56);

    const int M = P.M, D = P.Dout, NPAD = (M + 15) & ~15;
    const int warp = threadIdx.x >> 5;
    // 1-D grid: ngroups_d P-groups of nsplit_p row splits each, then ONE G group of nsplit_g splits.  The G group does the
    // three 3xTF32 passes itself (W_hi U_hi + W_hi U_lo, then W_lo U_hi, all into one TMEM accumulator): a row tile costs it
    // ~1.5x a P-group's, hence its own (larger) number of splits.
    const bool is_g = (int)blockIdx.x >= ngroups_d * nsplit_p;
    const int group = is_g ? ngroups_d : (int)blockIdx.x / nsplit_p;
    const int split = is_g ? (int)blockIdx.x - ngroups_d * nsplit_p : (int)blockIdx.x % nsplit_p;
    const int gsz = is_g ? nsplit_g : nsplit_p;
    const int d0 = group * 3;
    const int nd = is_g ? 0 : min(3, D - d0);             // P_d's of this CTA
    const bool has_q = !is_g && group == ngroups_d - 1;   // qmubar rides with the last P group
    const int nb = is_g ? 1 : nd + (has_q ? 1 : 0);       // B operands per row tile
    const int ntiles = (R + 127) / 128;
    const int my_tiles = (split < ntiles) ? (ntiles - 1 - split) / gsz + 1 : 0;
    const int n_it = is_g ? 2 * my_tiles : my_tiles;      // G: two sub-iterations per row tile (A = W_hi, then A = W_lo)

    if (threadIdx.x == 0) {
        mbar_init(bar_aready, RR_ROWTHREADS); mbar_init(bar_afree, 1);
        mbar_init(bar_bready, RR_ROWTHREADS); mbar_init(bar_bready + 8, RR_ROWTHREADS);
        mbar_init(bar_bfree, 1); mbar_init(bar_bfree + 8, 1);
        mbar_init(bar_done, 1);
        fence_mbar_init();
    }
    if (warp == RR_WARP_MMA) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot_gen;
    if (my_tiles == 0) {       // nothing to do (more CTAs than row tiles)
        __syncthreads();
        if (warp == RR_WARP_MMA) { __syncwarp(); tmem_dealloc(tmem, 512); }
        return;
    }

    if (warp == RR_WARP_MMA) {
        // ===================== MMA issuer =====================
        const uint32_t id_full = make_idesc_tf32(128, NPAD, 1, 1), id_q = make_idesc_tf32(128, 16, 1, 1);
        int bcount = 0;          // B operands consumed so far (buffer = bcount & 1, phase = (bcount >> 1) & 1)
        for (int it = 0; it < n_it; ++it) {
            mbar_wait(bar_aready, it & 1);
            const int nbi = is_g ? ((it & 1) ? 1 : 2) : nb;
            for (int b = 0; b < nbi; ++b, ++bcount) {
                const int buf = bcount & 1;
                mbar_wait(bar_bready + 8 * buf, (bcount >> 1) & 1);
                tc_fence_after();
                const bool isq = has_q && b == nd;
                const uint32_t dcol = is_g ? 0u : (isq ? 384u : 128u * (uint32_t)b);
                if (elect_one_rr()) {
                    const uint64_t ad0 = make_desc_sw128_mnmajor(A_t, 16384);
                    const uint64_t bd0 = make_desc_sw128_mnmajor(B_t + buf * RR_TILE_BYTES, 16384);
                    const uint32_t idd = isq ? id_q : id_full;
#pragma unroll
                    for (int ks = 0; ks < 16; ++ks)      // one k-step = 8 rows = 1024 B: +64 in the (>>4) start-address field
                        mma_tf32(tmem + dcol, ad0 + (uint64_t)(ks * 64), bd0 + (uint64_t)(ks * 64), idd,
                                 (it == 0 && ks == 0 && (!is_g || b == 0)) ? 0u : 1u);
                    mma_commit(bar_bfree + 8 * buf);
                    if (b == nbi - 1) mma_commit(bar_afree);
                }
                __syncwarp();
            }
        }
        if (elect_one_rr()) mma_commit(bar_done);
        __syncwarp();
    } else {
        // ===================== row warps: four threads per row =====================
        const int t = threadIdx.x & 127, qt = threadIdx.x >> 7;
        const uint32_t rsw = (uint32_t)(t & 3);                        // k-row inside the 4-row swizzle atom
        const uint32_t rowoff = (uint32_t)(t * 128);                   // rows are 128 B apart; atoms (4 rows) 512 B
        auto store4 = [&](uint32_t base, int k4, float4 v) {           // k4: feature index, multiple of 4
            uint32_t gran = (uint32_t)((k4 & 31) >> 3);                // 32-byte granule inside the 128-byte row
            uint32_t off = (uint32_t)(k4 >> 5) * 16384u + rowoff + ((gran ^ rsw) << 5) + (uint32_t)((k4 & 7) << 2);
            *reinterpret_cast<float4*>(sgen + (base - sbase) + off) = v;
        };
        // this thread's feature range [c_lo, c_hi): a quarter of the NPAD features in units of 4 (<= 32 wide); the flush
        // reads TMEM in units of 8 columns: [f_lo, f_hi)
        const int n4 = NPAD >> 2, q4 = n4 >> 2, r4 = n4 & 3;
        const int c_lo = 4 * (qt * q4 + min(qt, r4)), c_hi = c_lo + 4 * (q4 + (qt < r4 ? 1 : 0));
        const int n8 = NPAD >> 3, q8 = n8 >> 2, r8 = n8 & 3;
        const int f_lo = 8 * (qt * q8 + min(qt, r8)), f_hi = f_lo + 8 * (q8 + (qt < r8 ? 1 : 0));
        const float* Asrc = is_g ? W : U;
        int bcount = 0;
        const bool dbg = dbgp && blockIdx.x == 0 && threadIdx.x == 0;
        int dbi = 0;
#define RSTAMP() do { if (dbg && dbi < 38) dbgp[dbi++] = clock64(); } while (0)
        RSTAMP();
        for (int it = 0; it < n_it; ++it) {
            const int tit = is_g ? (it >> 1) : it, sub = is_g ? (it & 1) : 0;
            const int tile = split + tit * gsz;
            const int row = tile * 128 + t;
            const bool valid = row < R;
            const int nbi = is_g ? (sub ? 1 : 2) : nb;
            // pull the next tile's row slices towards L1 while this tile is converted and multiplied: the row threads are
            // long-scoreboard bound on exactly these loads (46% of the stall samples, profiles/r1b_big3_ncu_stalls.txt)
            if (tit + 1 < my_tiles && sub == 0) {
                const int nrow = row + 128 * gsz;
                if (nrow < R) {
                    const char* pu = reinterpret_cast<const char*>(U + (size_t)nrow * M + c_lo);
                    const int nbytes = 4 * (min(c_hi, M) - c_lo);      // stay inside the row (c_hi is padded to 16 features)
                    for (int o = 0; o < nbytes; o += 128) asm volatile("prefetch.global.L1 [%0];" ::"l"(pu + o));
                    if (is_g) {
                        const char* pw = reinterpret_cast<const char*>(W + (size_t)nrow * M + c_lo);
                        for (int o = 0; o < nbytes; o += 128) asm volatile("prefetch.global.L1 [%0];" ::"l"(pw + o));
                    } else if (qt == 0) {
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const char*>(vbar + (size_t)nrow * D)));
                    }
                }
            }
            // this half's slice of the row(s), tf32-rounded, kept in registers
            float4 uv[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int c0 = c_lo + 4 * c;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid && c0 < c_hi) {
                    if (c0 + 4 <= M && (M & 3) == 0) v = *reinterpret_cast<const float4*>(U + (size_t)row * M + c0);
                    else {
                        if (c0 < M) v.x = U[(size_t)row * M + c0];
                        if (c0 + 1 < M) v.y = U[(size_t)row * M + c0 + 1];
                        if (c0 + 2 < M) v.z = U[(size_t)row * M + c0 + 2];
                        if (c0 + 3 < M) v.w = U[(size_t)row * M + c0 + 3];
                    }
                }
                uv[c] = v;
            }
            float scv[3] = {1.f, 1.f, 1.f};
            if (!is_g) {
#pragma unroll
                for (int b = 0; b < 3; ++b) scv[b] = (valid && b < nd) ? vbar[(size_t)row * D + d0 + b] : 0.f;
            }
            RSTAMP();     // loads issued / landed
            // ---- A operand
            if (it > 0) mbar_wait(bar_afree, (it - 1) & 1);
            RSTAMP();     // A free
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int c0 = c_lo + 4 * c;
                if (c0 < c_hi) {
                    float4 v = uv[c];
                    if (is_g) {
                        v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (valid) {
                            if (c0 + 4 <= M && (M & 3) == 0) v = *reinterpret_cast<const float4*>(Asrc + (size_t)row * M + c0);
                            else {
                                if (c0 < M) v.x = Asrc[(size_t)row * M + c0];
                                if (c0 + 1 < M) v.y = Asrc[(size_t)row * M + c0 + 1];
                                if (c0 + 2 < M) v.z = Asrc[(size_t)row * M + c0 + 2];
                                if (c0 + 3 < M) v.w = Asrc[(size_t)row * M + c0 + 3];
                            }
                        }
                    }
                    float4 hi = make_float4(tf32_rna(v.x), tf32_rna(v.y), tf32_rna(v.z), tf32_rna(v.w));
                    if (sub == 1) hi = make_float4(tf32_lo_trunc(v.x, hi.x), tf32_lo_trunc(v.y, hi.y), tf32_lo_trunc(v.z, hi.z), tf32_lo_trunc(v.w, hi.w));
                    store4(A_t, c0, hi);
                }
            }
            if (qt == 3 && it == 0) {   // zero the feature padding [NPAD, 128) of A and of both B buffers once
                for (int c0 = NPAD; c0 < 128; c0 += 4) {
                    store4(A_t, c0, make_float4(0.f, 0.f, 0.f, 0.f));
                    store4(B_t, c0, make_float4(0.f, 0.f, 0.f, 0.f));
                    store4(B_t + RR_TILE_BYTES, c0, make_float4(0.f, 0.f, 0.f, 0.f));
                }
            }
            fence_proxy_async();
            mbar_arrive(bar_aready);
            RSTAMP();     // A stored
            // ---- B operands
            for (int b = 0; b < nbi; ++b, ++bcount) {
                const int buf = bcount & 1;
                if (bcount >= 2) mbar_wait(bar_bfree + 8 * buf, ((bcount >> 1) - 1) & 1);
                const uint32_t Bb = B_t + buf * RR_TILE_BYTES;
                const bool isq = has_q && b == nd;
                if (isq) {
                    // B[r][d] = mubar[r][d] (16 columns, zero padded)
                    if (qt == 0) {
#pragma unroll
                        for (int c0 = 0; c0 < 16; c0 += 4) {
                            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (valid) {
                                if (c0 < D) v.x = mubar[(size_t)row * D + c0];
                                if (c0 + 1 < D) v.y = mubar[(size_t)row * D + c0 + 1];
                                if (c0 + 2 < D) v.z = mubar[(size_t)row * D + c0 + 2];
                                if (c0 + 3 < D) v.w = mubar[(size_t)row * D + c0 + 3];
                            }
                            store4(Bb, c0, make_float4(tf32_rna(v.x), tf32_rna(v.y), tf32_rna(v.z), tf32_rna(v.w)));
                        }
                    }
                } else {
                    const float sc = b == 0 ? scv[0] : b == 1 ? scv[1] : scv[2];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const int c0 = c_lo + 4 * c;
                        if (c0 < c_hi) {
                            float4 v = uv[c];
                            float4 hi = make_float4(tf32_rna(v.x * sc), tf32_rna(v.y * sc), tf32_rna(v.z * sc), tf32_rna(v.w * sc));
                            if (is_g && b == 1) hi = make_float4(tf32_lo_trunc(v.x, hi.x), tf32_lo_trunc(v.y, hi.y), tf32_lo_trunc(v.z, hi.z), tf32_lo_trunc(v.w, hi.w));   // U_lo
                            store4(Bb, c0, hi);
                        }
                    }
                }
                fence_proxy_async();
                mbar_arrive(bar_bready + 8 * buf);
            }
        }
        RSTAMP();         // all operands stored
        // ---- flush: TMEM lane i = output row i; this quarter's columns
        mbar_wait(bar_done, 0);
        tc_fence_after();
        RSTAMP();         // MMAs done
        const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        const int i = t;
        for (int b = 0; b < nb; ++b) {
            const bool isq = has_q && b == nd;
            float* out; int ldo, ncols;
            if (is_g) { out = P.G; ldo = M; ncols = M; }
            else if (isq) { out = P.qmubar; ldo = D; ncols = D; }
            else { out = P.Pd + (size_t)(d0 + b) * M * M; ldo = M; ncols = M; }
            const uint32_t dcol = isq ? 384u : 128u * (uint32_t)b;
            const int lo = isq ? (qt ? 16 : 0) : f_lo, hi = isq ? 16 : f_hi;
            for (int c0 = lo; c0 < hi; c0 += 8) {
                float v[8];
                __syncwarp();
                tmem_ld8(lane_addr + dcol + c0, v);
                if (i < M) {
                    float* dst = &out[(size_t)i * ldo + c0];
                    if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0 && c0 + 8 <= ncols) {       // 16-byte aligned: two vector reductions instead of eight scalar
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
                    } else {
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            if (c0 + u < ncols) atomicAdd(dst + u, v[u]);
                    }
                }
            }
        }
    }
    if (dbgp && blockIdx.x == 0 && threadIdx.x == 0) dbgp[39] = clock64();
    tc_fence_before();
    __syncthreads();
    if (warp == RR_WARP_MMA) { __syncwarp(); tc_fence_after(); tmem_dealloc(tmem, 512); }
}

bool tc_rowred_supported(const LayerDev& P) { return P.M <= 128 && P.M >= 8 && P.Dout <= 16; }

cudaError_t rowred_tc_init() {
    return cudaFuncSetAttribute(k_layer_rowred_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * RR_TILE_BYTES + 1024 + 256);
}

void launch_bwd_rowred_tc(const LayerDev& P, const BwdArgs& a, int num_sms, cudaStream_t st, long long* nl) {
    const int ngroups_d = (P.Dout + 2) / 3;
    const int ntiles = (a.R + 127) / 128;
    // the G group's row tile costs ~1.5x a P group's (five operand tiles and two row sets instead of four and one)
    int nsplit_p = max(1, min(ntiles, (int)(num_sms / (ngroups_d + 1.5))));
    int nsplit_g = max(1, min(ntiles, num_sms - ngroups_d * nsplit_p));
    k_layer_rowred_tc<<<ngroups_d * nsplit_p + nsplit_g, RR_THREADS, 3 * RR_TILE_BYTES + 1024 + 256, st>>>(
        P, a.U, a.W, a.mubar, a.vbar, a.R, ngroups_d, nsplit_p, nsplit_g, a.dbg_rr);
    *nl += 1;
}
