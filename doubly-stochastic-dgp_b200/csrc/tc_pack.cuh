// Geometry of the packed tf32 weight "band blocks" consumed by the tensor-core kernels.
// A weight operand B[n][k] (n, k < M) is triangular.  It is stored as the sequence of its 32-wide k-blocks IN MMA ORDER,
// each k-block holding only the row band that is non-zero, every band a [rows] x [32 k] SWIZZLE_128B K-major image
// (128 B per row, 8-row atoms of 1024 B).  One block = one contiguous TMA bulk copy.
//   pattern GE (k >= n : G1', G2, G7): k-blocks nkb-1 .. 0, rows [0, min(NPAD, 32 kb + 32))
//   pattern LE (k <= n : G1,  G5, G6): k-blocks 0 .. nkb-1, rows [32 kb, NPAD)
// Block order in wpack: [G1 hi][G1 lo][G1' hi][G1' lo][G2 hi d=0..D-1][G2 lo d=0..D-1], each at a stride of slot_bytes (the
// lo parts are only streamed when the layer's q_sqrt is large enough for 1xTF32 to matter); then the square operands
// S_d = L_d L_d^T (backward: y_d = S_d u), d = 0..D-1, each nkb full bands of NPAD rows (sfull_bytes), k-blocks 0..nkb-1.
#pragma once
#include <stdint.h>

namespace tcp {
enum { PAT_LE = 0, PAT_GE = 1 };

__host__ __device__ inline int nkb_of(int M) { return (M + 31) / 32; }
__host__ __device__ inline int npad_of(int M) { return (M + 15) & ~15; }
// rows of k-block kb's band, and the first row
__host__ __device__ inline int band_rows(int pat, int M, int kb) {
    const int NPAD = npad_of(M);
    return pat == PAT_GE ? (NPAD < 32 * kb + 32 ? NPAD : 32 * kb + 32) : NPAD - 32 * kb;
}
__host__ __device__ inline int band_row0(int pat, int kb) { return pat == PAT_GE ? 0 : 32 * kb; }
// Tail band: when the last k-block holds at most 8 real k's (M = 100: k = 96..99) it is ONE UMMA k-step.  Its image is then
// stored with 32-byte rows (SWIZZLE_32B K-major: 8-row atoms of 256 B) instead of 128-byte rows: a quarter of the bytes for
// the band that is mostly padding (M = 100: G2 38 -> 27.5 KB per output, S_d 56 -> 45.5 KB).  The band keeps its place in
// the packed buffer (offsets are still computed with 128-byte rows); only the transferred size and the descriptor change.
__host__ __device__ inline bool tail32(int M) { return M - 32 * (nkb_of(M) - 1) <= 8; }
__host__ __device__ inline bool is_tail32(int M, int kb) { return kb == nkb_of(M) - 1 && tail32(M); }
__host__ __device__ inline uint32_t band_tx_bytes(int pat, int M, int kb) {
    return (is_tail32(M, kb) ? 32u : 128u) * (uint32_t)band_rows(pat, M, kb);
}
__host__ __device__ inline uint32_t sfull_band_tx_bytes(int M, int kb) { return (is_tail32(M, kb) ? 32u : 128u) * (uint32_t)npad_of(M); }
// byte offset of element (row, k < 8) in a SWIZZLE_32B K-major band (256-byte aligned base): Swizzle<1,4,3>
__host__ __device__ inline uint32_t sw32_offset(int row, int k) {
    return (uint32_t)((row >> 3) * 256 + (row & 7) * 32 + ((((k >> 2) & 1) ^ ((row >> 2) & 1)) << 4) + ((k & 3) << 2));
}
// byte offset of k-block kb's band inside its block (bands are stored in MMA order)
__host__ __device__ inline uint32_t band_offset(int pat, int M, int kb) {
    const int nkb = nkb_of(M);
    uint32_t off = 0;
    if (pat == PAT_GE) { for (int b = nkb - 1; b > kb; --b) off += 128u * (uint32_t)band_rows(pat, M, b); }
    else { for (int b = 0; b < kb; ++b) off += 128u * (uint32_t)band_rows(pat, M, b); }
    return off;
}
__host__ __device__ inline uint32_t block_bytes(int pat, int M) { 
    uint32_t t = 0;
    for (int b = 0; b < nkb_of(M); ++b) t += 128u * (uint32_t)band_rows(pat, M, b);
    return t;
}
__host__ __device__ inline uint32_t slot_bytes(int M) {
    uint32_t a = block_bytes(PAT_LE, M), b = block_bytes(PAT_GE, M);
    return a > b ? a : b;
}
// block indices
__host__ __device__ inline int blk_g1(int part) { return part; }                 // part: 0 hi, 1 lo
__host__ __device__ inline int blk_g1p(int part) { return 2 + part; }
__host__ __device__ inline int blk_g2(int d) { return 4 + d; }
__host__ __device__ inline int blk_g2lo(int D, int d) { return 4 + D + d; }
__host__ __device__ inline int num_blocks(int D) { return 4 + 2 * D; }
// square (non-triangular) operands: one band = all NPAD rows of one k-block
__host__ __device__ inline uint32_t sfull_band_bytes(int M) { return 128u * (uint32_t)npad_of(M); }
__host__ __device__ inline uint32_t sfull_bytes(int M) { return (uint32_t)nkb_of(M) * sfull_band_bytes(M); }
__host__ __device__ inline size_t s_region_offset(int M, int D) { return (size_t)num_blocks(D) * slot_bytes(M); }
__host__ __device__ inline size_t pack_bytes(int M, int D) { return s_region_offset(M, D) + (size_t)D * sfull_bytes(M); }
}  // namespace tcp
