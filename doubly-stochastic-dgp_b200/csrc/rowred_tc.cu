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
#include <type_traits>

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
        // ===================== row warps: warp w owns rows 8w .. 8w+7 of every row tile, lane l owns features 4l .. 4l+3 ========
        // (round 1 had one thread per row: its 16-byte loads from rows 4M bytes apart cost one L1 wavefront per row and chunk,
        // and its MN-major stores -- rows 128 B apart -- were 8-way bank conflicts: half of all shared-memory wavefronts.  With
        // lanes along the feature dimension a row is read as one coalesced 4M-byte segment and a warp's store covers whole
        // 128-byte rows of the operand image.)
        const int lane = threadIdx.x & 31;
        float* vs_w = reinterpret_cast<float*>(sgen + 3 * RR_TILE_BYTES + 256) + warp * 256;      // [8][D] vbar, D <= 16
        float* ms_w = vs_w + 128;                                                              // [8][D] mubar
        const int f0 = 4 * lane;
        const bool fact = f0 < NPAD;                                   // lanes beyond the padded width only zero the padding
        const bool vec = (M & 3) == 0;
        const uint32_t foff = (uint32_t)(f0 >> 5) * 16384u + (uint32_t)((f0 & 7) << 2);
        const uint32_t gran = (uint32_t)((f0 & 31) >> 3);
        // element (row r = 8 warp + j, features f0..f0+3) of an operand image; r & 3 == j & 3, so the swizzle term is one of
        // four per-lane constants and j * 128 an immediate
        uint8_t* const wbase = sgen + foff + (uint32_t)warp * 1024u;
        const uint32_t swz[4] = {(gran ^ 0u) << 5, (gran ^ 1u) << 5, (gran ^ 2u) << 5, (gran ^ 3u) << 5};
        auto store4 = [&](uint32_t base, int j, float4 v) {
            *reinterpret_cast<float4*>(wbase + (base - sbase) + j * 128 + swz[j & 3]) = v;
        };
        auto load4s = [&](uint32_t base, int j) -> float4 {
            return *reinterpret_cast<const float4*>(wbase + (base - sbase) + j * 128 + swz[j & 3]);
        };
        auto load_row4 = [&](const float* __restrict__ src, int grow) -> float4 {     // features f0..f0+3 of global row grow
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (grow < R && f0 < M) {
                const float* p = src + (size_t)grow * M + f0;
                if (vec) v = *reinterpret_cast<const float4*>(p);
                else {
                    v.x = p[0];
                    if (f0 + 1 < M) v.y = p[1];
                    if (f0 + 2 < M) v.z = p[2];
                    if (f0 + 3 < M) v.w = p[3];
                }
            }
            return v;
        };
        // rows r0 .. r0+7 of src, this lane's four features.  Fast path (whole rows inside the matrix, M % 4 == 0): one base
        // pointer, eight loads at a constant stride -- the generic path's per-row address arithmetic and branches were ~15 % of
        // the kernel's instructions
        auto load_rows8 = [&](const float* __restrict__ src, int r0, float4 (&dst)[8]) {
            if (vec && r0 + 8 <= R) {                  // warp-uniform: lanes past the last feature are predicated off, not diverged
                const float4* p = reinterpret_cast<const float4*>(src + (size_t)r0 * M + (f0 < M ? f0 : 0));
                const int st = M >> 2;
#pragma unroll
                for (int j = 0; j < 8; ++j) dst[j] = f0 < M ? p[j * st] : make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) dst[j] = fact ? load_row4(src, r0 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        // this lane's (up to four) elements of a warp's [8 rows][D] slice of vbar / mubar: element idx = lane + 32 k is row
        // idx / D, output idx % D; stored transposed ([d][8]) so that the scale of row j for output d is slice[8 d + j]
        int sl_row[4], sl_dst[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int idx = lane + 32 * k;
            sl_row[k] = idx < 8 * D ? idx / D : 1 << 20;
            sl_dst[k] = idx < 8 * D ? (idx % D) * 8 + idx / D : 0;
        }
        int bcount = 0;
        const bool dbg = dbgp && blockIdx.x == 0 && threadIdx.x == 0;
        int dbi = 0;
#define RSTAMP() do { if (dbg && dbi < 20) dbgp[dbi++] = clock64(); } while (0)
        RSTAMP();
        // the loop over row tiles, compiled twice (P groups / G group) so that neither variant carries the other's registers
        auto tiles_loop = [&](auto GC) {
        constexpr bool G = decltype(GC)::value;
        // P groups: the rows of the NEXT tile (and its vbar / mubar slice) are loaded into registers while the current tile is
        // converted and multiplied -- a dependent global round trip costs 2-4k cycles here, about one tile's worth of work.
        // G group: nx holds the current tile's W rows (both operands are needed at once; its loads rely on the L1 prefetch).
        float4 nx[8];
        float vnx[4], mnx[4];
        auto load_slice = [&](int r0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const bool ok = r0 + sl_row[k] < R;
                vnx[k] = ok ? vbar[(size_t)r0 * D + lane + 32 * k] : 0.f;
                mnx[k] = (ok && has_q) ? mubar[(size_t)r0 * D + lane + 32 * k] : 0.f;
            }
        };
        if (!G) {
            const int r0 = split * 128 + 8 * warp;
#pragma unroll
            load_rows8(U, r0, nx);
            load_slice(r0);
        }
        for (int it = 0; it < n_it; ++it) {
            const int tit = G ? (it >> 1) : it, sub = G ? (it & 1) : 0;
            const int tile = split + tit * gsz;
            const int row0 = tile * 128 + 8 * warp;
            const int nbi = G ? (sub ? 1 : 2) : nb;
            float4 uv[8];
            if (G) {
                if (tit + 1 < my_tiles && sub == 0 && (lane & 7) == 0 && f0 < M) {      // one lane per 128-byte line of a row
                    const int nrow0 = row0 + 128 * gsz;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (nrow0 + j < R) {
                            asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const char*>(U + (size_t)(nrow0 + j) * M + f0)));
                            asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const char*>(W + (size_t)(nrow0 + j) * M + f0)));
                        }
                    }
                }
                load_rows8(U, row0, uv);
                load_rows8(W, row0, nx);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) uv[j] = nx[j];
                // this warp's rows of vbar / mubar -> its private slice of shared memory (read back as broadcasts below)
                __syncwarp();
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (sl_row[k] < 8) { vs_w[sl_dst[k]] = vnx[k]; if (has_q) ms_w[sl_dst[k]] = mnx[k]; }
                }
                __syncwarp();
                if (tit + 1 < my_tiles) {
                    const int nrow0 = row0 + 128 * gsz;
                    load_rows8(U, nrow0, nx);
                    load_slice(nrow0);
                }
            }
            RSTAMP();     // loads issued
            // ---- A operand: U (P groups) or W hi / lo (G group)
            if (it > 0) mbar_wait(bar_afree, (it - 1) & 1);
            RSTAMP();     // A free
            if (fact) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 v = G ? nx[j] : uv[j];
                    float4 hi = make_float4(tf32_rna(v.x), tf32_rna(v.y), tf32_rna(v.z), tf32_rna(v.w));
                    if (sub == 1) hi = make_float4(tf32_lo_trunc(v.x, hi.x), tf32_lo_trunc(v.y, hi.y), tf32_lo_trunc(v.z, hi.z), tf32_lo_trunc(v.w, hi.w));
                    store4(A_t, j, hi);
                }
            } else if (it == 0) {       // zero the feature padding [NPAD, 128) of A and of both B buffers once
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    store4(A_t, j, make_float4(0.f, 0.f, 0.f, 0.f));
                    store4(B_t, j, make_float4(0.f, 0.f, 0.f, 0.f));
                    store4(B_t + RR_TILE_BYTES, j, make_float4(0.f, 0.f, 0.f, 0.f));
                }
            }
            fence_proxy_async();
            mbar_arrive(bar_aready);
            RSTAMP();     // A stored
            // ---- B operands
            for (int b = 0; b < nbi; ++b, ++bcount) {
                const int buf = bcount & 1;
                const uint32_t Bb = B_t + buf * RR_TILE_BYTES;
                const bool isq = has_q && b == nd;
                if (dbg && it == 1) dbgp[20 + 4 * b] = clock64();
                if (bcount >= 2) mbar_wait(bar_bfree + 8 * buf, ((bcount >> 1) - 1) & 1);
                if (dbg && it == 1) dbgp[21 + 4 * b] = clock64();
                if (isq) {
                    // B[r][d] = mubar[r][d] (16 columns, zero padded): lanes 0..3
                    if (lane < 4) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (f0 < D) v.x = ms_w[8 * f0 + j];
                            if (f0 + 1 < D) v.y = ms_w[8 * (f0 + 1) + j];
                            if (f0 + 2 < D) v.z = ms_w[8 * (f0 + 2) + j];
                            if (f0 + 3 < D) v.w = ms_w[8 * (f0 + 3) + j];
                            store4(Bb, j, make_float4(tf32_rna(v.x), tf32_rna(v.y), tf32_rna(v.z), tf32_rna(v.w)));
                        }
                    }
                } else if (fact) {
                    const float* scp = vs_w + 8 * (d0 + b);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        // P groups: the row comes back from the A image (tf32(u), exactly what the tensor core multiplies on the
                        // A side), which frees the registers that hold the prefetched next tile; G group: full-precision u
                        const float4 v = G ? uv[j] : load4s(A_t, j);
                        const float sc = G ? 1.f : scp[j];
                        float4 hi = make_float4(tf32_rna(v.x * sc), tf32_rna(v.y * sc), tf32_rna(v.z * sc), tf32_rna(v.w * sc));
                        if (G && b == 1) hi = make_float4(tf32_lo_trunc(v.x, hi.x), tf32_lo_trunc(v.y, hi.y), tf32_lo_trunc(v.z, hi.z), tf32_lo_trunc(v.w, hi.w));   // U_lo
                        store4(Bb, j, hi);
                    }
                }
                if (dbg && it == 1) dbgp[22 + 4 * b] = clock64();
                fence_proxy_async();
                mbar_arrive(bar_bready + 8 * buf);
                if (dbg && it == 1) dbgp[23 + 4 * b] = clock64();
            }
        }
        };
        if (is_g) tiles_loop(std::true_type{}); else tiles_loop(std::false_type{});
        RSTAMP();         // all operands stored
        // ---- flush.  TMEM lane i = output row i.  The accumulator is staged in shared memory (the operand buffers are dead)
        // so that the reductions go out row by row: a warp's vector reduction then covers one contiguous row segment instead
        // of 32 different rows.
        mbar_wait(bar_done, 0);
        tc_fence_after();
        RSTAMP();         // MMAs done
        {
            const int t = threadIdx.x & 127, qt = threadIdx.x >> 7;
            const int n8 = NPAD >> 3, q8 = n8 >> 2, r8 = n8 & 3;
            const int f_lo = 8 * (qt * q8 + min(qt, r8)), f_hi = f_lo + 8 * (q8 + (qt < r8 ? 1 : 0));
            const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
            const int SST = NPAD + 4;                  // staging row stride (floats): SST/4 odd -> conflict-free 16-byte accesses
            float* stage = reinterpret_cast<float*>(sgen);      // [128][SST] over A_t / B_t (<= 67.6 KB of the 192 KB)
            for (int b = 0; b < nb; ++b) {
                const bool isq = has_q && b == nd;
                float* out; int ldo, ncols;
                if (is_g) { out = P.G; ldo = M; ncols = M; }
                else if (isq) { out = P.qmubar; ldo = D; ncols = D; }
                else { out = P.Pd + (size_t)(d0 + b) * M * M; ldo = M; ncols = M; }
                const uint32_t dcol = isq ? 384u : 128u * (uint32_t)b;
                const int lo = isq ? (qt ? 16 : 0) : f_lo, hi = isq ? 16 : f_hi;
                if (b > 0) named_bar_sync(2, RR_ROWTHREADS);      // the previous accumulator's staging has been read
                for (int c0 = lo; c0 < hi; c0 += 8) {
                    float v[8];
                    __syncwarp();
                    tmem_ld8(lane_addr + dcol + c0, v);
                    *reinterpret_cast<float4*>(stage + t * SST + c0) = make_float4(v[0], v[1], v[2], v[3]);
                    *reinterpret_cast<float4*>(stage + t * SST + c0 + 4) = make_float4(v[4], v[5], v[6], v[7]);
                }
                named_bar_sync(2, RR_ROWTHREADS);
                if (f0 < ncols) {
                    for (int i = warp; i < M; i += RR_ROWTHREADS / 32) {
                        const float4 v = *reinterpret_cast<const float4*>(stage + i * SST + f0);
                        float* dst = &out[(size_t)i * ldo + f0];
                        if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0 && f0 + 4 <= ncols) {
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
                        } else {
                            atomicAdd(dst, v.x);
                            if (f0 + 1 < ncols) atomicAdd(dst + 1, v.y);
                            if (f0 + 2 < ncols) atomicAdd(dst + 2, v.z);
                            if (f0 + 3 < ncols) atomicAdd(dst + 3, v.w);
                        }
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
    return cudaFuncSetAttribute(k_layer_rowred_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * RR_TILE_BYTES + 1024 + 256 + 16 * 1024);
}

void launch_bwd_rowred_tc(const LayerDev& P, const BwdArgs& a, int num_sms, cudaStream_t st, long long* nl) {
    const int ngroups_d = (P.Dout + 2) / 3;
    const int ntiles = (a.R + 127) / 128;
    // the G group's row tile costs ~1.5x a P group's (five operand tiles and two row sets instead of four and one)
    int nsplit_p = max(1, min(ntiles, (int)(num_sms / (ngroups_d + 1.5))));
    int nsplit_g = max(1, min(ntiles, num_sms - ngroups_d * nsplit_p));
    k_layer_rowred_tc<<<ngroups_d * nsplit_p + nsplit_g, RR_THREADS, 3 * RR_TILE_BYTES + 1024 + 256 + 16 * 1024, st>>>(
        P, a.U, a.W, a.mubar, a.vbar, a.R, ngroups_d, nsplit_p, nsplit_g, a.dbg_rr);
    *nl += 1;
}
