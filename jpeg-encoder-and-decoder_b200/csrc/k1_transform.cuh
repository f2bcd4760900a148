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
 *   phase 1 (all threads): thread (row y, 16-pixel segment s) converts 16 pixels of the three
 *           planes to FP32 Y/Cb/Cr and stores float4s to shared memory in a
 *           [comp][y][half][slot][4] layout that writer and reader both touch conflict-free.
 *   phase 2 (warp c = component c): one thread per 8x8 block keeps all 64 samples in
 *           registers: 16 x 8-point butterflies, 64 FFMA quantise, static zig-zag pack to
 *           32 words, stored as 8 x 16 B to a 128-byte-row shared tile with an XOR swizzle.
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
    uint32_t *clear; /* small scratch array the entropy stage accumulates into: cleared here, for free */
    int clear_n;
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

typedef float4 K1F[8][2][32];   /* [y][half][slot] per component        : 8 KB  */
typedef uint4 K1Out[32][8];     /* [slot][chunk ^ (slot & 7)] per component row : 4 KB */

/* phase 1 body for one thread: 2 x 8 pixels of row y (tile-local blocks 2*seg, 2*seg+1) */
template <int RATIO>
__device__ __forceinline__ void k1_convert(K1F *sF, int y, int seg, const uint2 (&vr)[2], const uint2 (&vg)[2],
                                           const uint2 (&vb)[2])
{
    float cs_r[8], cs_g[8], cs_b[8]; /* 4:2:2 pair sums of both halves */
#pragma unroll
    for (int e = 0; e < 2; e++) {
        float fr[8], fg[8], fb[8], o[8];
        unpack8(vr[e], fr);
        unpack8(vg[e], fg);
        unpack8(vb[e], fb);
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
            unpack8_pairsum(vr[e], cs_r + 4 * e);
            unpack8_pairsum(vg[e], cs_g + 4 * e);
            unpack8_pairsum(vb[e], cs_b + 4 * e);
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

/* phase 2: one thread per 8x8 block */
template <int RATIO>
__device__ __forceinline__ void k1_phase2(const K1F *sF, K1Out *sOut, const float (*sMul)[64], int tid)
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

/* ---------------------------------------------------------------------------------------------
 * k1_transform: one CTA per tile, direct 8-byte global loads (any width % 8 == 0)
 * --------------------------------------------------------------------------------------------- */
template <int RATIO>
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
        const int y = tid >> 4, seg = tid & 15;
        uint2 vr[2], vg[2], vb[2];
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int bx = tb0 + 2 * seg + e;
            vr[e] = vg[e] = vb[e] = make_uint2(0, 0);
            if (bx < p.BW) {
                long long row = (long long)by * 8 + y - (bx == p.BW - 1 ? 1 : 0);
                if (row >= 0 || p.has_halo) {
                    long long idx = row * p.W + (long long)bx * 8;
                    vr[e] = __ldg(reinterpret_cast<const uint2 *>(pr + idx));
                    vg[e] = __ldg(reinterpret_cast<const uint2 *>(pg + idx));
                    vb[e] = __ldg(reinterpret_cast<const uint2 *>(pb + idx));
                } else {
                    vr[e] = ld_oob(p.oob);
                    vg[e] = ld_oob(p.oob + 8);
                    vb[e] = ld_oob(p.oob + 16);
                }
            }
        }
        k1_convert<RATIO>(sF, y, seg, vr, vg, vb);
    }
    __syncthreads();
    k1_phase2<RATIO>(sF, sOut, sMul, tid);
    __syncthreads();
    k1_phase3<RATIO>(sOut, p, frame, by, tb0, tid);
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
    o.frame = (int)(tt / per_frame);
    const unsigned rem = tt - (unsigned)o.frame * per_frame;
    o.by = (int)(rem / tpr);
    o.tb0 = (int)(rem - (unsigned)o.by * tpr) * K1_TILE_BLOCKS;
    return o;
}

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
            const int y = tid >> 4, seg = tid & 15;
            uint2 vr[2], vg[2], vb[2];
            const uint4 qr = *reinterpret_cast<const uint4 *>(&st.raw[0][y][seg * 16]);
            const uint4 qg = *reinterpret_cast<const uint4 *>(&st.raw[1][y][seg * 16]);
            const uint4 qb = *reinterpret_cast<const uint4 *>(&st.raw[2][y][seg * 16]);
            vr[0] = make_uint2(qr.x, qr.y);
            vr[1] = make_uint2(qr.z, qr.w);
            vg[0] = make_uint2(qg.x, qg.y);
            vg[1] = make_uint2(qg.z, qg.w);
            vb[0] = make_uint2(qb.x, qb.y);
            vb[1] = make_uint2(qb.z, qb.w);
            /* block column BW-1 is read one pixel row higher */
            const int jl = p.BW - 1 - tl.tb0; /* its tile-local index, if inside this tile */
            if ((jl >> 1) == seg && jl >= 0 && jl < K1_TILE_BLOCKS) {
                const int e = jl & 1;
                uint2 ar, ag, ab;
                if (tl.by == 0 && y == 0 && !p.has_halo) {
                    ar = ld_oob(p.oob);
                    ag = ld_oob(p.oob + 8);
                    ab = ld_oob(p.oob + 16);
                } else {
                    ar = *reinterpret_cast<const uint2 *>(&st.alt[0][y][8]);
                    ag = *reinterpret_cast<const uint2 *>(&st.alt[1][y][8]);
                    ab = *reinterpret_cast<const uint2 *>(&st.alt[2][y][8]);
                }
                if (e == 0) {
                    vr[0] = ar;
                    vg[0] = ag;
                    vb[0] = ab;
                } else {
                    vr[1] = ar;
                    vg[1] = ag;
                    vb[1] = ab;
                }
            }
            k1_convert<RATIO>(S.sF, y, seg, vr, vg, vb);
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

}  // namespace jpgb200
#endif
