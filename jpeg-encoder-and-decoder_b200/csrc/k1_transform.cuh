/*
 * K1 -- fused colour-convert -> (4:2:2 average) -> level shift -> FDCT -> quantise -> zig-zag.
 *
 * Replaces, in one pass over the pixels (citations relative to /root/reference/):
 *   convert_blocks + blockToCoords  src/preprocess.c:101-174,199-211
 *   level_shift                      src/preprocess.c:176-196
 *   subsample_422 (a stub there)     src/downsample.c:24-27   (semantics: SURVEY.md A.3)
 *   dct_block                        src/dct.c:36-59
 *   quantise_lum / quantise_chr      src/quantise.c:52-72
 *   zig_zag_block                    src/zig_zag.c:48-59
 *
 * Data layout in HBM
 *   in : three u8 planes r,g,b (W*H each, top-down raster) as bmp_GetRed/Green/Blue
 *        deliver them                                                   3 B / pixel
 *   out: int16 zig-zag coefficients, component-planar [Y | Cb | Cr][block][64] with
 *        blocks in raster order (the reference's zig_zag_Y/Cb/Cr order)  6 B / pixel (4:4:4)
 *   => algorithmic traffic 9 B/pixel (4:4:4), 7 B/pixel (4:2:2); HBM-bound by design.
 *
 * Work decomposition: a tile = 32 horizontally consecutive 8x8 blocks (256 x 8 pixels) of one
 * block row, processed by 128 threads in three phases:
 *   phase 1 (all threads): thread (k, block j) converts rows k and 7-k of block j (2 x 8 pixels
 *           of the three planes) to FP32 Y/Cb/Cr held as ROW PAIRS (f[k][x], f[7-k][x]): bytes
 *           become floats by PRMT + one FADD2, the colour maths is FFMA2 on the pairs; float2
 *           stores to a [comp][k][x][slot] layout that writer and reader touch conflict-free.
 *   phase 2 (warp c = component c): one thread per 8x8 block keeps all 64 samples in
 *           registers as 32 FP32x2 pairs: row pass = 4 packed 8-point butterflies, column pass
 *           = 64 scalar adds (its first stage, which combines the two rows of a pair and writes
 *           results paired over the horizontal frequency) + 4 packed butterfly tails, 32 FFMA2
 *           quantise, static zig-zag pack to 32 words, stored as 8 x 16 B to a 128-byte-row
 *           shared tile with an XOR swizzle.  (FADD2/FMUL2/FFMA2 are sm_100a's two-wide FP32
 *           instructions: bit-identical results, half the issue slots.)
 *   phase 3 (all threads): the tile's coefficients are contiguous in HBM per component
 *           (32 blocks x 128 B = 4 KB); they are copied out with fully coalesced 16 B stores.
 *
 * Two kernels share the phases:
 *   k1_transform_tma (default, width % 16 == 0): persistent CTAs; the raw pixel rows of the
 *           NEXT tiles are fetched by the TMA unit (cp.async.bulk global->shared, completion
 *           on an mbarrier) into a ring while the current tile is transformed, so no warp
 *           ever waits on an HBM load.
 *   k1_transform (any width % 8 == 0): one CTA per tile, pixels loaded with 8-byte LDGs.
 *
 * Bug-compatible addressing: the last block column of every block row is
 * read one pixel row higher; its row -1 (top-right block) comes from `oob` (the bytes the
 * reference finds in front of its plane buffers) or, for an MCU-row partition of a larger
 * image, from the halo row stored in front of the band.
 */
#ifndef JPGB200_K1_TRANSFORM_CUH
#define JPGB200_K1_TRANSFORM_CUH

#include "jpgb200_math.cuh"

namespace jpgb200 {

struct K1Params {
    const uint8_t *r, *g, *b; /* planes of frame 0 */
    int16_t *zz_y, *zz_cb, *zz_cr;
    long long plane_stride; /* bytes between consecutive frames' planes (batch) */
    long long zz_stride;    /* int16 elements between consecutive frames' coefficient sets */
    int W, H, BW;           /* pixels, pixels, W/8 */
    int has_halo;           /* row -1 of the band is valid memory (MCU-row partition) */
    int tiles_per_row, n_frames;
    unsigned long long magic_pf, magic_tpr; /* ceil(2^40 / tiles per frame), ceil(2^40 / tiles per row); 0 = divide */
    uint32_t *clear; /* small scratch array the entropy stage accumulates into: cleared here, for free */
    int clear_n;
    float mul[2][64];       /* quantiser multipliers [lum|chr][v*8+u] */
    uint8_t oob[24];        /* r[8] g[8] b[8] */
};

#define K1_THREADS 128
#define K1_TILE_BLOCKS 32

__device__ __forceinline__ uint2 ld_oob(const uint8_t *o)
{
    uint2 v;
    v.x = (uint32_t)o[0] | ((uint32_t)o[1] << 8) | ((uint32_t)o[2] << 16) | ((uint32_t)o[3] << 24);
    v.y = (uint32_t)o[4] | ((uint32_t)o[5] << 8) | ((uint32_t)o[6] << 16) | ((uint32_t)o[7] << 24);
    return v;
}

/* FP32 samples between phase 1 and phase 2.  A block is held as ROW PAIRS
 * P[k][x] = (f[k][x], f[7-k][x]), k = 0..3 (the pairing fdct8x8_p wants), one float2 each --
 * the unit the packed arithmetic produces and consumes, so neither side shuffles registers.
 * [k][x][slot]: the 32 phase-1 threads of a (k) warp and the 32 phase-2 threads of a component
 * warp both touch 32 consecutive float2 -> conflict-free. */
typedef float2 K1F[4][8][32];   /* per component row : 8 KB */
typedef uint4 K1Out[32][8];     /* [slot][chunk ^ (slot & 7)] per component row : 4 KB */

/* the 8 pixels of rows k and 7-k of one plane -> 8 row pairs */
__device__ __forceinline__ void k1_pairs8(const uint2 a, const uint2 b, f32x2 *o)
{
#pragma unroll
    for (int x = 0; x < 4; x++) {
        o[x] = bytes_to_f32x2(a.x, b.x, x);
        o[4 + x] = bytes_to_f32x2(a.y, b.y, x);
    }
}

/* sums of horizontally adjacent pixel pairs of 16 pixels (rows k and 7-k) -> 8 row pairs */
__device__ __forceinline__ void k1_pairsums16(const uint4 a, const uint4 b, f32x2 *o)
{
    const uint32_t wa[4] = {a.x, a.y, a.z, a.w}, wb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int w = 0; w < 4; w++) {
        const uint32_t sa = (wa[w] & 0x00ff00ffu) + ((wa[w] >> 8) & 0x00ff00ffu);
        const uint32_t sb = (wb[w] & 0x00ff00ffu) + ((wb[w] >> 8) & 0x00ff00ffu);
        o[2 * w] = halves_to_f32x2(sa, sb, 0);
        o[2 * w + 1] = halves_to_f32x2(sa, sb, 1);
    }
}

__device__ __forceinline__ void k1_store_pairs(K1F &F, int k, int slot, const f32x2 *o)
{
#pragma unroll
    for (int x = 0; x < 8; x++) F[k][x][slot] = o[x];
}

/* phase 1, full-resolution components: thread (k, j) converts rows k and 7-k of block j.
 * ra/rb etc. = the 8 bytes of row k / row 7-k of the three planes. */
template <int RATIO, bool STD = false>
__device__ __forceinline__ void k1_convert_full(K1F *sF, int k, int j, const uint2 (&pa)[3], const uint2 (&pb)[3])
{
    f32x2 r[8], g[8], b[8], o[8];
    k1_pairs8(pa[0], pb[0], r);
    k1_pairs8(pa[1], pb[1], g);
    k1_pairs8(pa[2], pb[2], b);
#pragma unroll
    for (int x = 0; x < 8; x++) o[x] = colour_y_p(r[x], g[x], b[x]);
    k1_store_pairs(sF[0], k, j, o);
    if (RATIO == 0) {
#pragma unroll
        for (int x = 0; x < 8; x++) o[x] = STD ? colour_cb_std_p(r[x], g[x], b[x]) : colour_cb_p(r[x], g[x], b[x]);
        k1_store_pairs(sF[1], k, j, o);
#pragma unroll
        for (int x = 0; x < 8; x++) o[x] = colour_cr_p(r[x], g[x], b[x]);
        k1_store_pairs(sF[2], k, j, o);
    }
}

/* phase 1, 4:2:2 chroma: thread (k, lane) with lane = which*16 + jc averages the 16 pixels of
 * rows k and 7-k under chroma block jc and stores Cb (which = 0, slot jc) or Cr (slot 16+jc) */
template <bool STD = false>
__device__ __forceinline__ void k1_convert_chroma422(K1F *sF, int k, int lane, const uint4 (&pa)[3], const uint4 (&pb)[3])
{
    f32x2 r[8], g[8], b[8], o[8];
    k1_pairsums16(pa[0], pb[0], r);
    k1_pairsums16(pa[1], pb[1], g);
    k1_pairsums16(pa[2], pb[2], b);
    if (lane < 16) {
#pragma unroll
        for (int x = 0; x < 8; x++) o[x] = STD ? colour_cb2_std_p(r[x], g[x], b[x]) : colour_cb2_p(r[x], g[x], b[x]);
    } else {
#pragma unroll
        for (int x = 0; x < 8; x++) o[x] = colour_cr2_p(r[x], g[x], b[x]);
    }
    k1_store_pairs(sF[1], k, lane, o);
}

/* phase 2: one thread per 8x8 block (warp = component row, lane = slot) */
template <int RATIO>
__device__ __forceinline__ void k1_phase2(const K1F *sF, K1Out *sOut, const float (*sMul)[64], int tid)
{
    const int c = tid >> 5, lane = tid & 31;
    const int ncomp_rows = (RATIO == 0) ? 3 : 2;
    if (c < ncomp_rows) {
        f32x2 P[4][8], Q[8][4];
#pragma unroll
        for (int k = 0; k < 4; k++)
#pragma unroll
            for (int x = 0; x < 8; x++) P[k][x] = sF[c][k][x][lane];
        fdct8x8_p(P, Q);
        const float4 *m4 = reinterpret_cast<const float4 *>(sMul[c == 0 ? 0 : 1]);
        float f[64];
#pragma unroll
        for (int v = 0; v < 8; v++)
#pragma unroll
            for (int hh = 0; hh < 2; hh++) {
                const float4 m = m4[v * 2 + hh];
                const f32x2 q0 = p_fma(Q[v][2 * hh], make_float2(m.x, m.y), p_bc(JPGB_ROUND_MAGIC));
                const f32x2 q1 = p_fma(Q[v][2 * hh + 1], make_float2(m.z, m.w), p_bc(JPGB_ROUND_MAGIC));
                f[v * 8 + 4 * hh + 0] = q0.x;
                f[v * 8 + 4 * hh + 1] = q0.y;
                f[v * 8 + 4 * hh + 2] = q1.x;
                f[v * 8 + 4 * hh + 3] = q1.y;
            }
        uint32_t o[32];
        zigzag_pack(f, o);
#pragma unroll
        for (int j = 0; j < 8; j++)
            sOut[c][lane][j ^ (lane & 7)] = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
    }
}

/* phase 3: coalesced copy-out of the tile's coefficients */
template <int RATIO>
__device__ __forceinline__ void k1_phase3(const K1Out *sOut, const K1Params &p, int frame, int by, int tb0, int tid)
{
    int16_t *zy = p.zz_y + (long long)frame * p.zz_stride;
    int16_t *zcb = p.zz_cb + (long long)frame * p.zz_stride;
    int16_t *zcr = p.zz_cr + (long long)frame * p.zz_stride;
    const int nchunks = (RATIO == 0) ? 3 * 256 : 2 * 256;
#pragma unroll
    for (int i = tid; i < nchunks; i += K1_THREADS) {
        const int crow = i >> 8, rem = i & 255, j = rem >> 3, ch = rem & 7;
        int slot;
        uint4 *dst;
        bool ok;
        if (RATIO == 0 || crow == 0) {
            /* j = tile-local block of a full-resolution component */
            slot = j;
            ok = (tb0 + j) < p.BW;
            int16_t *base = (crow == 0) ? zy : (crow == 1 ? zcb : zcr);
            dst = reinterpret_cast<uint4 *>(base + ((long long)by * p.BW + tb0 + j) * 64) + ch;
        } else {
            /* 4:2:2 chroma: j<16 -> Cb block j, else Cr block j-16 (slot == j) */
            slot = j;
            const int cj = j & 15, cbw = p.BW >> 1;
            ok = ((tb0 >> 1) + cj) < cbw;
            int16_t *base = (j < 16) ? zcb : zcr;
            dst = reinterpret_cast<uint4 *>(base + ((long long)by * cbw + (tb0 >> 1) + cj) * 64) + ch;
        }
        if (ok) *dst = sOut[crow][slot][ch ^ (slot & 7)];
    }
}

/* ---------------------------------------------------------------------------------------------
 * k1_transform: one CTA per tile, direct 8-byte global loads (any width % 8 == 0)
 * --------------------------------------------------------------------------------------------- */
/* STD = the standards-conforming front door (jpgb200_encode_planar_std): textbook Cb and no
 * shift of the last block column -- the reference's two coefficient-path bugs left out */
template <int RATIO, bool STD = false>
__global__ void __launch_bounds__(K1_THREADS, 4) k1_transform(const K1Params p)
{
    __shared__ K1F sF[3];
    __shared__ K1Out sOut[3];
    __shared__ float sMul[2][64];

    const int tid = threadIdx.x;
    const int by = blockIdx.y;
    const int tb0 = blockIdx.x * K1_TILE_BLOCKS;
    const int frame = blockIdx.z;
    const uint8_t *pr = p.r + (long long)frame * p.plane_stride;
    const uint8_t *pg = p.g + (long long)frame * p.plane_stride;
    const uint8_t *pb = p.b + (long long)frame * p.plane_stride;

    sMul[tid >> 6][tid & 63] = p.mul[tid >> 6][tid & 63];
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0)
        for (int i = tid; i < p.clear_n; i += K1_THREADS) p.clear[i] = 0u;
    {
        const int k = tid >> 5, lane = tid & 31;
        const uint8_t *pl[3] = {pr, pg, pb};
        /* 8 bytes of image row by*8+rr of block column bx (one row higher for the last column) */
        auto ld8 = [&](int c, int rr, int bx) -> uint2 {
            if (bx >= p.BW) return make_uint2(0u, 0u);
            const long long row = (long long)by * 8 + rr - ((!STD && bx == p.BW - 1) ? 1 : 0);
            if (row >= 0 || p.has_halo) return __ldg(reinterpret_cast<const uint2 *>(pl[c] + row * p.W + (long long)bx * 8));
            return ld_oob(p.oob + 8 * c);
        };
        uint2 pa[3], pb2[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            pa[c] = ld8(c, k, tb0 + lane);
            pb2[c] = ld8(c, 7 - k, tb0 + lane);
        }
        k1_convert_full<RATIO, STD>(sF, k, lane, pa, pb2);
        if (RATIO == 1) {
            const int jc = lane & 15;
            uint4 ca[3], cb[3];
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const uint2 l0 = ld8(c, k, tb0 + 2 * jc), l1 = ld8(c, k, tb0 + 2 * jc + 1);
                const uint2 h0 = ld8(c, 7 - k, tb0 + 2 * jc), h1 = ld8(c, 7 - k, tb0 + 2 * jc + 1);
                ca[c] = make_uint4(l0.x, l0.y, l1.x, l1.y);
                cb[c] = make_uint4(h0.x, h0.y, h1.x, h1.y);
            }
            k1_convert_chroma422<STD>(sF, k, lane, ca, cb);
        }
    }
    __syncthreads();
    k1_phase2<RATIO>(sF, sOut, sMul, tid);
    __syncthreads();
    k1_phase3<RATIO>(sOut, p, frame, by, tb0, tid);
}

/* ---------------------------------------------------------------------------------------------
 * k1_transform420: 4:2:0 (HORIZONTAL_VERTICAL_SUBSAMPLING, a printf stub in the reference,
 * src/downsample.c:29-32).  One CTA per tile of 16 MCUs (32 x 2 luma
 * blocks, 16 Cb, 16 Cr), direct loads.  Luma coefficients are stored in MCU order (Y00 Y01 Y10
 * Y11 per MCU: the order of the scan), chroma one block per MCU.  Chroma = mean of each 2x2 group,
 * computed from exact integer sums with quarter-weight colour coefficients.
 * --------------------------------------------------------------------------------------------- */
template <bool STD>
__global__ void __launch_bounds__(K1_THREADS, 4) k1_transform420(const K1Params p)
{
    __shared__ K1F sF[3];      /* luma block row 0, luma block row 1, chroma (slots 0..15 Cb, 16..31 Cr) */
    __shared__ K1Out sOut[3];
    __shared__ float sMul[2][64];

    const int tid = threadIdx.x, k = tid >> 5, lane = tid & 31;
    const int my = blockIdx.y;                   /* MCU row */
    const int tb0 = blockIdx.x * K1_TILE_BLOCKS; /* first luma block column of the tile */
    const int frame = blockIdx.z;
    const uint8_t *pl[3] = {p.r + (long long)frame * p.plane_stride, p.g + (long long)frame * p.plane_stride,
                            p.b + (long long)frame * p.plane_stride};
    sMul[tid >> 6][tid & 63] = p.mul[tid >> 6][tid & 63];
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0)
        for (int i = tid; i < p.clear_n; i += K1_THREADS) p.clear[i] = 0u;

    /* 8 bytes of image row y of block column bx (one row higher for the last column) */
    auto ld8 = [&](int c, long long y, int bx) -> uint2 {
        if (bx >= p.BW) return make_uint2(0u, 0u);
        const long long row = y - ((!STD && bx == p.BW - 1) ? 1 : 0);
        if (row >= 0 || p.has_halo) return __ldg(reinterpret_cast<const uint2 *>(pl[c] + row * p.W + (long long)bx * 8));
        return ld_oob(p.oob + 8 * c);
    };
    /* luma: rows k and 7-k of block (r, lane) */
#pragma unroll
    for (int r = 0; r < 2; r++) {
        uint2 pa[3], pb[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            pa[c] = ld8(c, (long long)my * 16 + 8 * r + k, tb0 + lane);
            pb[c] = ld8(c, (long long)my * 16 + 8 * r + 7 - k, tb0 + lane);
        }
        k1_convert_full<1, STD>(sF + r, k, lane, pa, pb); /* the luma half of the 4:2:2 converter */
    }
    /* chroma rows k and 7-k of MCU column jc = pixel rows (2k, 2k+1) and (14-2k, 15-2k), 16 wide */
    {
        const int jc = lane & 15;
        f32x2 sum[3][8];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            f32x2 o0[8], o1[8];
#pragma unroll
            for (int d = 0; d < 2; d++) {
                const long long ya = (long long)my * 16 + 2 * k + d, yb = (long long)my * 16 + 14 - 2 * k + d;
                const uint2 a0 = ld8(c, ya, tb0 + 2 * jc), a1 = ld8(c, ya, tb0 + 2 * jc + 1);
                const uint2 b0 = ld8(c, yb, tb0 + 2 * jc), b1 = ld8(c, yb, tb0 + 2 * jc + 1);
                k1_pairsums16(make_uint4(a0.x, a0.y, a1.x, a1.y), make_uint4(b0.x, b0.y, b1.x, b1.y), d ? o1 : o0);
            }
#pragma unroll
            for (int x = 0; x < 8; x++) sum[c][x] = p_add(o0[x], o1[x]);
        }
        f32x2 o[8];
        if (lane < 16) {
#pragma unroll
            for (int x = 0; x < 8; x++)
                o[x] = STD ? colour_cb4_std_p(sum[0][x], sum[1][x], sum[2][x]) : colour_cb4_p(sum[0][x], sum[1][x], sum[2][x]);
        } else {
#pragma unroll
            for (int x = 0; x < 8; x++) o[x] = colour_cr4_p(sum[0][x], sum[1][x], sum[2][x]);
        }
        k1_store_pairs(sF[2], k, lane, o);
    }
    __syncthreads();
    /* phase 2: warps 0,1 luma rows (luma table), warp 2 chroma */
    if (k < 3) {
        f32x2 P[4][8], Q[8][4];
#pragma unroll
        for (int kk = 0; kk < 4; kk++)
#pragma unroll
            for (int x = 0; x < 8; x++) P[kk][x] = sF[k][kk][x][lane];
        fdct8x8_p(P, Q);
        const float4 *m4 = reinterpret_cast<const float4 *>(sMul[k < 2 ? 0 : 1]);
        float f[64];
#pragma unroll
        for (int v = 0; v < 8; v++)
#pragma unroll
            for (int hh = 0; hh < 2; hh++) {
                const float4 m = m4[v * 2 + hh];
                const f32x2 q0 = p_fma(Q[v][2 * hh], make_float2(m.x, m.y), p_bc(JPGB_ROUND_MAGIC));
                const f32x2 q1 = p_fma(Q[v][2 * hh + 1], make_float2(m.z, m.w), p_bc(JPGB_ROUND_MAGIC));
                f[v * 8 + 4 * hh + 0] = q0.x;
                f[v * 8 + 4 * hh + 1] = q0.y;
                f[v * 8 + 4 * hh + 2] = q1.x;
                f[v * 8 + 4 * hh + 3] = q1.y;
            }
        uint32_t o[32];
        zigzag_pack(f, o);
#pragma unroll
        for (int j = 0; j < 8; j++) sOut[k][lane][j ^ (lane & 7)] = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
    }
    __syncthreads();
    /* phase 3: luma block (r, j) -> MCU (my, (tb0+j)/2), sub-block 2r + (j&1); chroma slot s -> MCU tb0/2 + (s&15) */
    {
        int16_t *zy = p.zz_y + (long long)frame * p.zz_stride;
        int16_t *zcb = p.zz_cb + (long long)frame * p.zz_stride;
        int16_t *zcr = p.zz_cr + (long long)frame * p.zz_stride;
        const int mw = p.BW >> 1;
        for (int i = tid; i < 3 * 256; i += K1_THREADS) {
            const int crow = i >> 8, rem = i & 255, j = rem >> 3, ch = rem & 7;
            uint4 *dst;
            bool ok;
            if (crow < 2) {
                const int bx = tb0 + j;
                ok = bx < p.BW;
                const long long m = (long long)my * mw + (bx >> 1);
                dst = reinterpret_cast<uint4 *>(zy + (m * 4 + 2 * crow + (bx & 1)) * 64) + ch;
            } else {
                const int mx = (tb0 >> 1) + (j & 15);
                ok = mx < mw;
                dst = reinterpret_cast<uint4 *>((j < 16 ? zcb : zcr) + ((long long)my * mw + mx) * 64) + ch;
            }
            if (ok) *dst = sOut[crow][j][ch ^ (j & 7)];
        }
    }
}

/* ---------------------------------------------------------------------------------------------
 * k1_transform_tma: persistent CTAs, raw pixel rows prefetched by TMA bulk copies
 * --------------------------------------------------------------------------------------------- */
#define K1_STAGES 2
#define K1_ROW_BYTES (K1_TILE_BLOCKS * 8) /* 256 */

struct K1Stage {
    uint8_t raw[3][8][K1_ROW_BYTES]; /* [plane][row][x]                              6144 B */
    uint8_t alt[3][8][16];           /* last block column, one row higher: x = W-16.. 384 B */
};

struct K1SmemTma {
    K1Stage stage[K1_STAGES];
    K1F sF[3];
    K1Out sOut[3];
    float sMul[2][64];
    unsigned long long full[K1_STAGES]; /* mbarriers */
};

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra WAIT_%=;\n"
        "}\n" ::"r"(a),
        "r"(parity)
        : "memory");
}
/* TMA bulk copy global -> shared, completion counted in bytes on an mbarrier (UBLKCP) */
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst), b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d),
                 "l"(src), "r"(bytes), "r"(b)
                 : "memory");
}

struct K1Tile {
    int frame, by, tb0;
};
__device__ __forceinline__ K1Tile k1_tile(const K1Params &p, long long t)
{
    K1Tile o;
    const unsigned tpr = (unsigned)p.tiles_per_row, per_frame = tpr * (unsigned)(p.H >> 3);
    const unsigned tt = (unsigned)t; /* tile counts stay far below 2^32 */
    if (p.magic_pf) {
        /* exact for n * d < 2^40 (checked by the host): two wide multiplies instead of two divisions */
        o.frame = (int)(((unsigned long long)tt * p.magic_pf) >> 40);
        const unsigned rem = tt - (unsigned)o.frame * per_frame;
        o.by = (int)(((unsigned long long)rem * p.magic_tpr) >> 40);
        o.tb0 = (int)(rem - (unsigned)o.by * tpr) * K1_TILE_BLOCKS;
        return o;
    }
    o.frame = (int)(tt / per_frame);
    const unsigned rem = tt - (unsigned)o.frame * per_frame;
    o.by = (int)(rem / tpr);
    o.tb0 = (int)(rem - (unsigned)o.by * tpr) * K1_TILE_BLOCKS;
    return o;
}

#ifdef JPGB200_AB /* an A/B form: built only with -DJPGB200_AB (build.py: JPGB200_AB=1) */
/* warp 0: arm the stage's mbarrier and issue the tile's bulk copies (one row per lane) */
__device__ __forceinline__ void k1_issue_tile(const K1Params &p, K1Stage &st, unsigned long long *bar, long long t,
                                              int lane)
{
    const K1Tile tl = k1_tile(p, t);
    const int x0 = tl.tb0 * 8;
    const int row_bytes = min(K1_ROW_BYTES, p.W - x0);
    const bool has_last = (tl.tb0 + K1_TILE_BLOCKS >= p.BW); /* tile holds block column BW-1 */
    const int alt_rows = has_last ? ((tl.by == 0 && !p.has_halo) ? 7 : 8) : 0;
    if (lane == 0) mbar_expect_tx(bar, 24u * (unsigned)row_bytes + 3u * 16u * (unsigned)alt_rows);
    __syncwarp();
    if (lane < 24) {
        const int pl = lane >> 3, y = lane & 7;
        const uint8_t *base = (pl == 0 ? p.r : (pl == 1 ? p.g : p.b)) + (long long)tl.frame * p.plane_stride;
        const long long row = (long long)tl.by * 8 + y;
        tma_load_1d(&st.raw[pl][y][0], base + row * p.W + x0, (unsigned)row_bytes, bar);
        if (has_last && (row - 1 >= 0 || p.has_halo))
            tma_load_1d(&st.alt[pl][y][0], base + (row - 1) * p.W + (p.W - 16), 16u, bar);
    }
}

template <int RATIO>
__global__ void __launch_bounds__(K1_THREADS, 4) k1_transform_tma(const K1Params p, long long n_tiles)
{
    extern __shared__ __align__(128) unsigned char k1_smem[];
    K1SmemTma &S = *reinterpret_cast<K1SmemTma *>(k1_smem);
    const int tid = threadIdx.x, lane = tid & 31;

    S.sMul[tid >> 6][tid & 63] = p.mul[tid >> 6][tid & 63];
    if (blockIdx.x == 0)
        for (int i = tid; i < p.clear_n; i += K1_THREADS) p.clear[i] = 0u;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < K1_STAGES; s++) mbar_init(&S.full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    /* prologue: fill the ring */
    if (tid < 32) {
#pragma unroll
        for (int s = 0; s < K1_STAGES; s++) {
            const long long t = (long long)blockIdx.x + (long long)s * gridDim.x;
            if (t < n_tiles) k1_issue_tile(p, S.stage[s], &S.full[s], t, lane);
        }
    }

    int it = 0;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x, it++) {
        const int s = it % K1_STAGES;
        const unsigned parity = (unsigned)(it / K1_STAGES) & 1u;
        const K1Tile tl = k1_tile(p, t);
        mbar_wait(&S.full[s], parity);

        /* ---- phase 1: raw bytes (shared) -> FP32 Y/Cb/Cr (shared) ---- */
        {
            const K1Stage &st = S.stage[s];
            const int k = tid >> 5;
            /* block column BW-1 is read one pixel row higher: its rows come from
             * `alt`, row -1 of the image from the caller's out-of-bounds bytes */
            const int jl = p.BW - 1 - tl.tb0; /* its tile-local index, if inside this tile */
            const bool top_oob = (tl.by == 0 && !p.has_halo);
            auto ld8 = [&](int c, int rr, int j) -> uint2 {
                if (j == jl) {
                    if (top_oob && rr == 0) return ld_oob(p.oob + 8 * c);
                    return *reinterpret_cast<const uint2 *>(&st.alt[c][rr][8]);
                }
                return *reinterpret_cast<const uint2 *>(&st.raw[c][rr][j * 8]);
            };
            uint2 pa[3], pb[3];
#pragma unroll
            for (int c = 0; c < 3; c++) {
                pa[c] = ld8(c, k, lane);
                pb[c] = ld8(c, 7 - k, lane);
            }
            k1_convert_full<RATIO>(S.sF, k, lane, pa, pb);
            if (RATIO == 1) {
                const int jc = lane & 15;
                uint4 ca[3], cb[3];
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const uint2 l0 = ld8(c, k, 2 * jc), l1 = ld8(c, k, 2 * jc + 1);
                    const uint2 h0 = ld8(c, 7 - k, 2 * jc), h1 = ld8(c, 7 - k, 2 * jc + 1);
                    ca[c] = make_uint4(l0.x, l0.y, l1.x, l1.y);
                    cb[c] = make_uint4(h0.x, h0.y, h1.x, h1.y);
                }
                k1_convert_chroma422(S.sF, k, lane, ca, cb);
            }
        }
        __syncthreads();
        /* the stage's raw bytes are consumed: refill it with the tile K1_STAGES iterations ahead.
         * Issued by warp 3, which has no 8x8 blocks to transform in phase 2 (4:4:4: 3 component
         * warps, 4:2:2: 2), so the TMA descriptors cost nothing on the critical path. */
        if (tid >= 96) {
            const long long tn = t + (long long)K1_STAGES * gridDim.x;
            if (tn < n_tiles) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                k1_issue_tile(p, S.stage[s], &S.full[s], tn, lane);
            }
        }
        k1_phase2<RATIO>(S.sF, S.sOut, S.sMul, tid);
        __syncthreads();
        k1_phase3<RATIO>(S.sOut, p, tl.frame, tl.by, tl.tb0, tid);
    }
}
#endif /* JPGB200_AB */

}  // namespace jpgb200
#endif
