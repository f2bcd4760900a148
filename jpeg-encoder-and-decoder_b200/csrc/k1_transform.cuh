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
 * Work decomposition
 *   One CTA of 128 threads owns a tile of 32 horizontally consecutive 8x8 blocks
 *   (256 x 8 pixels) of one block row.
 *   phase 1 (all 128 threads): thread (row y, 16-pixel segment s) loads 2 x 8 bytes per
 *           plane (coalesced: a warp covers 2 rows x 256 contiguous bytes), converts to
 *           FP32 Y/Cb/Cr and stores float4s to shared memory in a [comp][y][half][slot][4]
 *           layout that both the writer and the phase-2 reader touch conflict-free.
 *   phase 2 (warp c = component c): one thread per 8x8 block keeps all 64 samples in
 *           registers: 16 x 8-point butterflies, 64 FFMA quantise, static zig-zag pack to
 *           32 words, stored as 8 x 16 B to a 128-byte-row shared tile with an XOR swizzle.
 *   phase 3 (all threads): the tile's coefficients are contiguous in HBM per component
 *           (32 blocks x 128 B = 4 KB); they are copied out with fully coalesced 16 B stores.
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
    float mul[2][64];       /* quantiser multipliers [lum|chr][v*8+u] */
    uint8_t oob[24];        /* r[8] g[8] b[8] */
};

#define K1_THREADS 128
#define K1_TILE_BLOCKS 32

__device__ __forceinline__ void unpack8(uint2 v, float *f)
{
    f[0] = (float)(v.x & 0xffu);
    f[1] = (float)((v.x >> 8) & 0xffu);
    f[2] = (float)((v.x >> 16) & 0xffu);
    f[3] = (float)(v.x >> 24);
    f[4] = (float)(v.y & 0xffu);
    f[5] = (float)((v.y >> 8) & 0xffu);
    f[6] = (float)((v.y >> 16) & 0xffu);
    f[7] = (float)(v.y >> 24);
}

/* pairwise byte sums of 8 pixels -> 4 floats (exact small integers) */
__device__ __forceinline__ void unpack8_pairsum(uint2 v, float *f)
{
    f[0] = (float)((v.x & 0xffu) + ((v.x >> 8) & 0xffu));
    f[1] = (float)(((v.x >> 16) & 0xffu) + (v.x >> 24));
    f[2] = (float)((v.y & 0xffu) + ((v.y >> 8) & 0xffu));
    f[3] = (float)(((v.y >> 16) & 0xffu) + (v.y >> 24));
}

__device__ __forceinline__ uint2 ld_oob(const uint8_t *o)
{
    uint2 v;
    v.x = (uint32_t)o[0] | ((uint32_t)o[1] << 8) | ((uint32_t)o[2] << 16) | ((uint32_t)o[3] << 24);
    v.y = (uint32_t)o[4] | ((uint32_t)o[5] << 8) | ((uint32_t)o[6] << 16) | ((uint32_t)o[7] << 24);
    return v;
}

/* slot of tile-local block j in the shared layouts: even blocks first, then odd ones, so
 * that the 16 segment-threads of a row write consecutive slots */
__device__ __forceinline__ int k1_slot(int j) { return (j >> 1) | ((j & 1) << 4); }
__device__ __forceinline__ int k1_block_of_slot(int s) { return ((s & 15) << 1) | (s >> 4); }

template <int RATIO>
__global__ void __launch_bounds__(K1_THREADS, 4) k1_transform(const K1Params p)
{
    /* [comp][y][half][slot] float4 : 3*8*2*32*16 B = 24 KB */
    __shared__ float4 sF[3][8][2][32];
    /* [comp-row 0..2][slot] rows of 128 B, 16-byte chunks XOR-swizzled by (slot & 7) : 12 KB */
    __shared__ uint4 sOut[3][32][8];
    __shared__ float sMul[2][64];

    const int tid = threadIdx.x;
    const int by = blockIdx.y;
    const int tb0 = blockIdx.x * K1_TILE_BLOCKS;
    const int frame = blockIdx.z;
    const uint8_t *pr = p.r + (long long)frame * p.plane_stride;
    const uint8_t *pg = p.g + (long long)frame * p.plane_stride;
    const uint8_t *pb = p.b + (long long)frame * p.plane_stride;

    sMul[tid >> 6][tid & 63] = p.mul[tid >> 6][tid & 63];

    /* ---------------- phase 1: load, convert, stage as FP32 ---------------- */
    {
        const int y = tid >> 4, seg = tid & 15;
        float cs_r[8], cs_g[8], cs_b[8]; /* 4:2:2 pair sums of both halves */
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int bx = tb0 + 2 * seg + e;
            uint2 vr = make_uint2(0, 0), vg = vr, vb = vr;
            if (bx < p.BW) {
                long long row = (long long)by * 8 + y - (bx == p.BW - 1 ? 1 : 0);
                if (row >= 0 || p.has_halo) {
                    long long idx = row * p.W + (long long)bx * 8;
                    vr = __ldg(reinterpret_cast<const uint2 *>(pr + idx));
                    vg = __ldg(reinterpret_cast<const uint2 *>(pg + idx));
                    vb = __ldg(reinterpret_cast<const uint2 *>(pb + idx));
                } else {
                    vr = ld_oob(p.oob);
                    vg = ld_oob(p.oob + 8);
                    vb = ld_oob(p.oob + 16);
                }
            }
            float fr[8], fg[8], fb[8], o[8];
            unpack8(vr, fr);
            unpack8(vg, fg);
            unpack8(vb, fb);
            const int slot = seg + 16 * e;
#pragma unroll
            for (int i = 0; i < 8; i++) o[i] = colour_y(fr[i], fg[i], fb[i]);
            sF[0][y][0][slot] = make_float4(o[0], o[1], o[2], o[3]);
            sF[0][y][1][slot] = make_float4(o[4], o[5], o[6], o[7]);
            if (RATIO == 0) {
#pragma unroll
                for (int i = 0; i < 8; i++) o[i] = colour_cb(fr[i], fg[i], fb[i]);
                sF[1][y][0][slot] = make_float4(o[0], o[1], o[2], o[3]);
                sF[1][y][1][slot] = make_float4(o[4], o[5], o[6], o[7]);
#pragma unroll
                for (int i = 0; i < 8; i++) o[i] = colour_cr(fr[i], fg[i], fb[i]);
                sF[2][y][0][slot] = make_float4(o[0], o[1], o[2], o[3]);
                sF[2][y][1][slot] = make_float4(o[4], o[5], o[6], o[7]);
            } else {
                unpack8_pairsum(vr, cs_r + 4 * e);
                unpack8_pairsum(vg, cs_g + 4 * e);
                unpack8_pairsum(vb, cs_b + 4 * e);
            }
        }
        if (RATIO == 1) {
            /* one chroma block row (8 averaged samples) per thread: Cb -> slot seg,
             * Cr -> slot 16+seg of component row 1 */
            float o[8];
#pragma unroll
            for (int i = 0; i < 8; i++) o[i] = colour_cb2(cs_r[i], cs_g[i], cs_b[i]);
            sF[1][y][0][seg] = make_float4(o[0], o[1], o[2], o[3]);
            sF[1][y][1][seg] = make_float4(o[4], o[5], o[6], o[7]);
#pragma unroll
            for (int i = 0; i < 8; i++) o[i] = colour_cr2(cs_r[i], cs_g[i], cs_b[i]);
            sF[1][y][0][16 + seg] = make_float4(o[0], o[1], o[2], o[3]);
            sF[1][y][1][16 + seg] = make_float4(o[4], o[5], o[6], o[7]);
        }
    }
    __syncthreads();

    /* ---------------- phase 2: one thread per 8x8 block ---------------- */
    {
        const int c = tid >> 5, lane = tid & 31;
        const int ncomp_rows = (RATIO == 0) ? 3 : 2;
        if (c < ncomp_rows) {
            float f[64];
#pragma unroll
            for (int y = 0; y < 8; y++) {
                float4 a = sF[c][y][0][lane], b4 = sF[c][y][1][lane];
                f[y * 8 + 0] = a.x;
                f[y * 8 + 1] = a.y;
                f[y * 8 + 2] = a.z;
                f[y * 8 + 3] = a.w;
                f[y * 8 + 4] = b4.x;
                f[y * 8 + 5] = b4.y;
                f[y * 8 + 6] = b4.z;
                f[y * 8 + 7] = b4.w;
            }
            fdct8x8(f);
            const float4 *m4 = reinterpret_cast<const float4 *>(sMul[c == 0 ? 0 : 1]);
#pragma unroll
            for (int i = 0; i < 16; i++) {
                float4 m = m4[i];
                f[4 * i + 0] = jfma(f[4 * i + 0], m.x, JPGB_ROUND_MAGIC);
                f[4 * i + 1] = jfma(f[4 * i + 1], m.y, JPGB_ROUND_MAGIC);
                f[4 * i + 2] = jfma(f[4 * i + 2], m.z, JPGB_ROUND_MAGIC);
                f[4 * i + 3] = jfma(f[4 * i + 3], m.w, JPGB_ROUND_MAGIC);
            }
            uint32_t o[32];
            zigzag_pack(f, o);
#pragma unroll
            for (int j = 0; j < 8; j++)
                sOut[c][lane][j ^ (lane & 7)] = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
        }
    }
    __syncthreads();

    /* ---------------- phase 3: coalesced copy-out ---------------- */
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
                slot = k1_slot(j);
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
}

}  // namespace jpgb200
#endif
