/*
 * K1, software-pipelined form (k1_transform_pipe).
 *
 * Same arithmetic and HBM layout as the other two forms.  The warp-autonomous kernel
 * (k1_warp.cuh) is bound by instruction issue, and a quarter of its transform instructions are
 * issued for idle lanes (24 block transforms per 32-lane warp) while every warp also pays for
 * its own fetch and copy-out.  Here the CTA works on one tile of 32 blocks at a time and its four
 * warps specialise:
 *
 *   all four warps   phase 1 of tile i : warp k converts rows k and 7-k of all 32 blocks
 *   -- barrier A --
 *   warps 0..2       phase 2 of tile i : warp c transforms the 32 blocks of component c, every
 *                    lane busy; results wait in registers
 *   warp 3           meanwhile: cp.async prefetch of tile i+2's raw rows into a three-slot ring,
 *                    and phase 3 of tile i-1 (its coefficients, still in shared memory, go out
 *                    with coalesced 16-byte stores)
 *   -- barrier C --  (named barrier: warp 3 arrives when tile i-1 has left shared memory)
 *   warps 0..2       store tile i's coefficients into shared memory
 *   -- barrier B --  (tile i+1's raw rows have landed)
 *
 * so the fetch and the copy-out cost the transform warps nothing and no lane idles in phase 2.
 */
#ifndef JPGB200_K1_PIPE_CUH
#define JPGB200_K1_PIPE_CUH

#include <cuda.h>

#include "k1_transform.cuh"
#include "k1_warp.cuh"

namespace jpgb200 {

#ifndef K1P_STAGES
#define K1P_STAGES 2 /* ring of raw tiles: prefetch distance K1P_STAGES-1 */
#endif
#define K1P_MIN_CTAS 4

struct K1PSmem {
    K1Out sOut[3]; /* first: the TMA store's 128-byte swizzle wants 1 KB alignment */
    uint8_t raw[K1P_STAGES][3][8][K1_ROW_BYTES]; /* [slot][plane][row][x] */
    K1F sF[3];
    float mul[2][64];
};

/* The coefficients as a 2-D tensor [block][64 x int16] for TMA stores: a tile's component row
 * (32 blocks x 128 B, or 16 for a 4:2:2 chroma half) leaves shared memory with ONE instruction.
 * K1Out's XOR of the 16-byte chunk index with (row & 7) is exactly CU_TENSOR_MAP_SWIZZLE_128B. */
struct K1PStore {
    CUtensorMap box32, box16;
    long long row_cb, row_cr; /* first Cb / Cr block (tensor row); Y starts at 0 */
    long long rows_per_frame; /* zz_stride / 64 */
    int enabled;
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap *tm, const void *smem, int c0, int c1)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tm), "r"(c0), "r"(c1),
                 "r"(sa)
                 : "memory");
}

/* warp 3: asynchronous fetch of one tile's raw rows, 16 bytes per copy.  The block of column BW-1
 * is fetched one pixel row higher; row -1 of an image without halo is left alone. */
template <bool STD>
__device__ __forceinline__ void k1p_fetch(const K1Params &p, uint8_t (*raw)[8][K1_ROW_BYTES], const K1Tile &tl, int lane)
{
    /* lane = (row parity, 16-byte column): 4 rows x 3 planes = 12 copies, one address computation */
    const int part = lane & 15, ysub = lane >> 4;
    const int bx = tl.tb0 + 2 * part;
    if (bx < p.BW) {
        const long long off = (long long)tl.frame * p.plane_stride + ((long long)tl.by * 8 + ysub) * p.W + (long long)bx * 8;
        const int jl = STD ? -1 : p.BW - 1 - bx; /* 0 or 1: the chunk holds the last block column */
        const long long step = 2ll * p.W;
        if (jl == 0 || jl == 1) {
            const bool top = (tl.by == 0 && ysub == 0 && !p.has_halo); /* row -1 does not exist */
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const uint8_t *src = (c == 0 ? p.r : (c == 1 ? p.g : p.b)) + off;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    uint8_t *dst = &raw[c][2 * q + ysub][part * 16];
                    if (jl == 1) cp_async8(dst, src + q * step);
                    if (!(top && q == 0)) cp_async8(dst + 8 * jl, src + q * step + 8 * jl - p.W);
                }
            }
        } else {
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const uint8_t *src = (c == 0 ? p.r : (c == 1 ? p.g : p.b)) + off;
#pragma unroll
                for (int q = 0; q < 4; q++) cp_async16(&raw[c][2 * q + ysub][part * 16], src + q * step);
            }
        }
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
}

/* warp 3: copy-out of a finished tile */
template <int RATIO>
__device__ __forceinline__ void k1p_phase3(const K1Out *sOut, const K1Params &p, const K1Tile &tl, int lane)
{
    int16_t *zy = p.zz_y + (long long)tl.frame * p.zz_stride;
    int16_t *zcb = p.zz_cb + (long long)tl.frame * p.zz_stride;
    int16_t *zcr = p.zz_cr + (long long)tl.frame * p.zz_stride;
    const int ch = lane & 7, jsub = lane >> 3;
    /* full-resolution component rows: 32 blocks = 4 KB contiguous */
#pragma unroll
    for (int c = 0; c < (RATIO == 0 ? 3 : 1); c++) {
        uint4 *dst = reinterpret_cast<uint4 *>((c == 0 ? zy : (c == 1 ? zcb : zcr)) + ((long long)tl.by * p.BW + tl.tb0) * 64);
        const int nvalid = min(K1_TILE_BLOCKS, p.BW - tl.tb0);
#pragma unroll
        for (int t = 0; t < 8; t++) {
            const int j = t * 4 + jsub;
            if (j < nvalid) dst[j * 8 + ch] = sOut[c][j][ch ^ (j & 7)];
        }
    }
    if (RATIO == 1) {
        /* 4:2:2 chroma row: slots 0..15 Cb, 16..31 Cr, 2 KB contiguous each */
        const int cbw = p.BW >> 1, cx0 = tl.tb0 >> 1;
        const int nvalid = min(16, cbw - cx0);
#pragma unroll
        for (int t = 0; t < 8; t++) {
            const int j = t * 4 + jsub, cj = j & 15;
            uint4 *dst = reinterpret_cast<uint4 *>((j < 16 ? zcb : zcr) + ((long long)tl.by * cbw + cx0) * 64);
            if (cj < nvalid) dst[cj * 8 + ch] = sOut[1][j][ch ^ (j & 7)];
        }
    }
}

/* warp 3: a whole tile with three (4:2:2: three smaller) TMA tensor stores; tiles cut short by the
 * image's right edge take the manual path (a full box would spill into the next block row) */
template <int RATIO>
__device__ __forceinline__ void k1p_store(const K1Out *sOut, const K1Params &p, const K1PStore &st, const K1Tile &tl, int lane)
{
    const bool full = (tl.tb0 + K1_TILE_BLOCKS <= p.BW);
    if (!st.enabled || !full) {
        k1p_phase3<RATIO>(sOut, p, tl, lane);
        return;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); /* the transform warps' writes -> async proxy */
    if (lane == 0) {
        const long long f0 = (long long)tl.frame * st.rows_per_frame;
        const long long ry = f0 + (long long)tl.by * p.BW + tl.tb0;
        tma_store_2d(&st.box32, &sOut[0], 0, (int)ry);
        if (RATIO == 0) {
            tma_store_2d(&st.box32, &sOut[1], 0, (int)(st.row_cb + ry));
            tma_store_2d(&st.box32, &sOut[2], 0, (int)(st.row_cr + ry));
        } else {
            const long long rc = f0 + (long long)tl.by * (p.BW >> 1) + (tl.tb0 >> 1);
            tma_store_2d(&st.box16, &sOut[1][0], 0, (int)(st.row_cb + rc));
            tma_store_2d(&st.box16, &sOut[1][16], 0, (int)(st.row_cr + rc));
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); /* shared memory has been read */
    }
    __syncwarp();
}

#ifdef JPGB200_AB /* an A/B form: built only with -DJPGB200_AB (build.py: JPGB200_AB=1) */
template <int RATIO, bool STD = false>
__global__ void __launch_bounds__(K1_THREADS, K1P_MIN_CTAS) k1_transform_pipe(const K1Params p, long long n_tiles,
                                                                              const __grid_constant__ K1PStore st)
{
    extern __shared__ unsigned char k1_smem_raw[];
    /* 1 KB alignment for the swizzled TMA source */
    unsigned char *k1_smem = k1_smem_raw + ((1024u - ((unsigned)__cvta_generic_to_shared(k1_smem_raw) & 1023u)) & 1023u);
    K1PSmem &S = *reinterpret_cast<K1PSmem *>(k1_smem);
    /* Warp w of every CTA runs on scheduler w % 4: rotate the roles with the CTA index so that the
     * service role (and the idle transform role of 4:2:2) is spread over the SM's four schedulers */
    const int tid = threadIdx.x, lane = tid & 31, w = ((tid >> 5) + (int)(blockIdx.x & 3u)) & 3;
    constexpr int NP2 = (RATIO == 0) ? 3 : 2; /* transform warps with work */

    S.mul[tid >> 6][tid & 63] = p.mul[tid >> 6][tid & 63];
    if (blockIdx.x == 0)
        for (int i = tid; i < p.clear_n; i += K1_THREADS) p.clear[i] = 0u;
    K1Tile tl = k1_tile(p, blockIdx.x), prev = tl;
    if (w == 3) {
        /* prologue: the first K1P_STAGES-1 tiles of the ring */
#pragma unroll
        for (int d = 0; d < K1P_STAGES - 1; d++) {
            const long long td = (long long)blockIdx.x + (long long)d * gridDim.x;
            if (td < n_tiles)
                k1p_fetch<STD>(p, S.raw[d], k1_tile(p, td), lane);
            else
                asm volatile("cp.async.commit_group;\n" ::: "memory");
        }
        asm volatile("cp.async.wait_group %0;\n" ::"n"(K1P_STAGES - 2) : "memory");
    }
    __syncthreads();

    int it = 0;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x, it++) {
        const int s = it % K1P_STAGES;
        /* ---- phase 1: warp k, lane j ---- */
        {
            const int k = w;
            const int jl = STD ? -1 : p.BW - 1 - tl.tb0;
            const bool top_oob = (tl.by == 0 && !p.has_halo);
            auto ld8 = [&](int c, int rr, int j) -> uint2 {
                if (top_oob && rr == 0 && j == jl) return ld_oob(p.oob + 8 * c);
                return *reinterpret_cast<const uint2 *>(&S.raw[s][c][rr][j * 8]);
            };
            uint2 pa[3], pb[3];
#pragma unroll
            for (int c = 0; c < 3; c++) {
                pa[c] = ld8(c, k, lane);
                pb[c] = ld8(c, 7 - k, lane);
            }
            k1_convert_full<RATIO, STD>(S.sF, k, lane, pa, pb);
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
                k1_convert_chroma422<STD>(S.sF, k, lane, ca, cb);
            }
        }
        __syncthreads(); /* A: the FP32 samples of tile i are complete, its raw rows consumed */

        const long long tn = t + gridDim.x;
        if (w == 3) {
            /* the fetch of tile i+K1P_STAGES-1 first (its loads are in flight during the copy-out) */
            const long long tn2 = t + (long long)(K1P_STAGES - 1) * gridDim.x;
            if (tn2 < n_tiles)
                k1p_fetch<STD>(p, S.raw[(it + K1P_STAGES - 1) % K1P_STAGES], k1_tile(p, tn2), lane);
            else
                asm volatile("cp.async.commit_group;\n" ::: "memory");
            if (it > 0) k1p_store<RATIO>(S.sOut, p, st, prev, lane);
            asm volatile("bar.arrive 1, 128;\n" ::: "memory"); /* C: tile i-1 has left shared memory */
            asm volatile("cp.async.wait_group %0;\n" ::"n"(K1P_STAGES - 2) : "memory"); /* tile i+1's rows are in */
            prev = tl;
            if (tn < n_tiles) tl = k1_tile(p, tn);
        } else {
            /* ---- phase 2: warp c, lane = block ---- */
            uint32_t o[32];
            if (w < NP2) {
                f32x2 P[4][8], Q[8][4];
#pragma unroll
                for (int k = 0; k < 4; k++)
#pragma unroll
                    for (int x = 0; x < 8; x++) P[k][x] = S.sF[w][k][x][lane];
                fdct8x8_p(P, Q);
                const float4 *m4 = reinterpret_cast<const float4 *>(S.mul[w == 0 ? 0 : 1]);
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
                zigzag_pack(f, o);
            }
            asm volatile("bar.sync 1, 128;\n" ::: "memory"); /* C */
            if (w < NP2) {
#pragma unroll
                for (int j = 0; j < 8; j++)
                    S.sOut[w][lane][j ^ (lane & 7)] = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
            }
            if (tn < n_tiles) tl = k1_tile(p, tn);
        }
        __syncthreads(); /* B: tile i's coefficients are in shared memory, tile i+1's raw rows have landed */
    }
    if (w == 3 && it > 0) {
        k1p_store<RATIO>(S.sOut, p, st, prev, lane);
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}
#endif /* JPGB200_AB */

}  // namespace jpgb200
#endif
