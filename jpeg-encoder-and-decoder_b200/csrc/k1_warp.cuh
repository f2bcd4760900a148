/*
 * K1, warp-autonomous form (k1_transform_ws) -- the default whenever width % 16 == 0.
 *
 * Same arithmetic and the same HBM layout as k1_transform_tma (k1_transform.cuh); what changes
 * is who waits for whom.  The CTA-barrier form parks three warps at __syncthreads() while the
 * slowest finishes, and leaves the fourth warp idle during the transform; ncu showed it
 * latency-bound (issue active 54 %, 1.6 of 3.6 warps per scheduler stalled at the barrier), not
 * bound by any pipe or by HBM.  Here a CTA is four independent warps:
 *
 *   tile   = 32 horizontally consecutive 8x8 blocks of one block row
 *   warp w = the strip of 8 blocks [8w, 8w+8) of the tile, start to finish:
 *     fetch    its 3 x 8 rows x 64 raw bytes, two tiles ahead, with asynchronous 16-byte copies
 *              (LDGSTS, cp.async groups) into a private three-slot ring -- measured against TMA
 *              bulk copies shared by the CTA: those need one UBLKCP per 256-byte row, issued
 *              serially through the uniform datapath, and kept warps waiting on the mbarrier for
 *              27 % of their time (profiles/README.md)
 *     phase 1  lane (k, j): rows k and 7-k of block j -> FP32 Y/Cb/Cr row pairs (PRMT + FADD2,
 *              FFMA2), written to the warp's private scratch
 *     phase 2  lane (c, j), 24 of 32 lanes: the 8x8 block of component c -- packed row pass,
 *              scalar first stage + packed tail of the column pass, FFMA2 quantise, zig-zag pack
 *              -- written back over the scratch as 144-byte rows
 *     phase 3  the strip's 24 x 128 B of coefficients (1 KB contiguous in HBM per component)
 *              copied out with coalesced 16-byte stores
 *   Only __syncwarp() separates the phases; warps never wait for each other.
 *
 * 4:2:2 uses the same structure with 8 luma + 4 Cb + 4 Cr blocks per strip (16 lanes in phase 2).
 */
#ifndef JPGB200_K1_WARP_CUH
#define JPGB200_K1_WARP_CUH

#include "k1_transform.cuh"

namespace jpgb200 {

#define K1W_STAGES 3   /* per-warp ring of raw strips: prefetch distance 2 */
#define K1W_MIN_CTAS 4 /* 128 registers: no spills; measured 82.5 us against 87.2 us with 5 CTAs at 96 registers */

/* raw bytes of one strip: [plane][row][64 B].  Rows k and k+1 lie 16 banks apart, so the
 * (k, j) lanes of phase 1 read them conflict-free. */
struct K1WRaw {
    uint8_t b[3][8][64];
};

/* per-warp scratch: phase 1 -> 2 as float2 [c][k][x'][j] with x' = x ^ ((k ^ c) & 1) (the XOR
 * keeps both the (k, j) writers and the (c, j) readers of a half-warp on disjoint banks);
 * phase 2 -> 3 as 24 rows of 144 bytes over the same memory */
struct K1WScratch {
    float2 f[3 * 4 * 8 * 8]; /* 6144 B */
};

struct K1WSmem {
    K1WRaw raw[4][K1W_STAGES];
    K1WScratch scratch[4];
    float mul[2][64];
};

__device__ __forceinline__ int k1w_idx(int c, int k, int x, int j) { return (((c * 4 + k) * 8) + (x ^ ((k ^ c) & 1))) * 8 + j; }

__device__ __forceinline__ void cp_async8(void *smem, const void *gmem)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gmem) : "memory");
}

/* One warp fetches its strip of tile t: lane (y, part) copies 16 bytes (two blocks) of row y of
 * each plane, asynchronously (LDGSTS: no registers, completion by cp.async group).  The block of
 * column BW-1 is fetched one pixel row higher; row -1 of an image without halo
 * is left alone -- phase 1 substitutes the caller's out-of-bounds bytes for it. */
template <bool STD>
__device__ __forceinline__ void k1w_fetch(const K1Params &p, K1WRaw &raw, long long t, int w, int lane)
{
    const K1Tile tl = k1_tile(p, t);
    const int y = lane >> 2, part = lane & 3;
    const int bx = tl.tb0 + 8 * w + 2 * part; /* first of the chunk's two block columns */
    if (bx < p.BW) {
        const long long row = (long long)tl.by * 8 + y;
        const long long off = (long long)tl.frame * p.plane_stride + row * p.W + (long long)bx * 8;
        const int jl = STD ? -1 : p.BW - 1 - bx; /* 0 or 1: the chunk holds the last block column */
        const uint8_t *src[3] = {p.r + off, p.g + off, p.b + off};
        if (jl == 0 || jl == 1) {
            const bool up_ok = (row - 1 >= 0) || p.has_halo;
#pragma unroll
            for (int c = 0; c < 3; c++) {
                if (jl == 1) cp_async8(&raw.b[c][y][part * 16], src[c]);
                if (up_ok) cp_async8(&raw.b[c][y][part * 16 + 8 * jl], src[c] + 8 * jl - p.W);
            }
        } else {
#pragma unroll
            for (int c = 0; c < 3; c++) cp_async16(&raw.b[c][y][part * 16], src[c]);
        }
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
}

#ifdef JPGB200_AB /* an A/B form: built only with -DJPGB200_AB (build.py: JPGB200_AB=1) */
template <int RATIO, bool STD = false>
__global__ void __launch_bounds__(K1_THREADS, K1W_MIN_CTAS) k1_transform_ws(const K1Params p, long long n_tiles)
{
    extern __shared__ __align__(128) unsigned char k1_smem[];
    K1WSmem &S = *reinterpret_cast<K1WSmem *>(k1_smem);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    constexpr int NCOMP = (RATIO == 0) ? 3 : 2;  /* component rows in the scratch */
    constexpr int NACT = NCOMP * 8;              /* phase-2 lanes */

    S.mul[tid >> 6][tid & 63] = p.mul[tid >> 6][tid & 63];
    if (blockIdx.x == 0)
        for (int i = tid; i < p.clear_n; i += K1_THREADS) p.clear[i] = 0u;
    __syncthreads(); /* the only CTA-wide barrier: the quantiser table */

    /* prologue: the first K1W_STAGES-1 strips of this warp */
#pragma unroll
    for (int s = 0; s < K1W_STAGES - 1; s++) {
        const long long t = (long long)blockIdx.x + (long long)s * gridDim.x;
        if (t < n_tiles)
            k1w_fetch<STD>(p, S.raw[w][s], t, w, lane);
        else
            asm volatile("cp.async.commit_group;\n" ::: "memory");
    }

    float2 *F = S.scratch[w].f;
    uint4 *rows = reinterpret_cast<uint4 *>(S.scratch[w].f); /* [24][9] uint4 = 144-byte rows */
    const int k = lane >> 3, jj = lane & 7;
    int it = 0;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x, it++) {
        const int s = it % K1W_STAGES;
        const K1Tile tl = k1_tile(p, t);
        {
            /* the slot consumed in the previous iteration is refilled with the strip two tiles ahead */
            const long long tn = t + (long long)(K1W_STAGES - 1) * gridDim.x;
            if (tn < n_tiles)
                k1w_fetch<STD>(p, S.raw[w][(it + K1W_STAGES - 1) % K1W_STAGES], tn, w, lane);
            else
                asm volatile("cp.async.commit_group;\n" ::: "memory");
        }
        asm volatile("cp.async.wait_group %0;\n" ::"n"(K1W_STAGES - 1) : "memory");
        __syncwarp();

        /* ---- phase 1 ---- */
        {
            const K1WRaw &raw = S.raw[w][s];
            const int jl = STD ? -1 : p.BW - 1 - tl.tb0 - 8 * w; /* strip-local index of block column BW-1 */
            const bool top_oob = (tl.by == 0 && !p.has_halo);
            auto ld8 = [&](int c, int rr, int j) -> uint2 {
                if (RATIO == 1 && top_oob && rr == 0 && j == jl) return ld_oob(p.oob + 8 * c);
                return *reinterpret_cast<const uint2 *>(&raw.b[c][rr][j * 8]);
            };
            f32x2 r[8], g[8], b[8];
            {
                uint2 a0 = ld8(0, k, jj), a1 = ld8(1, k, jj), a2 = ld8(2, k, jj);
                if (RATIO == 0 && top_oob && k == 0 && jj == jl) { /* image row -1 of block column BW-1 */
                    a0 = ld_oob(p.oob);
                    a1 = ld_oob(p.oob + 8);
                    a2 = ld_oob(p.oob + 16);
                }
                k1_pairs8(a0, ld8(0, 7 - k, jj), r);
                k1_pairs8(a1, ld8(1, 7 - k, jj), g);
                k1_pairs8(a2, ld8(2, 7 - k, jj), b);
            }
#pragma unroll
            for (int x = 0; x < 8; x++) F[k1w_idx(0, k, x, jj)] = colour_y_p(r[x], g[x], b[x]);
            if (RATIO == 0) {
#pragma unroll
                for (int x = 0; x < 8; x++)
                    F[k1w_idx(1, k, x, jj)] = STD ? colour_cb_std_p(r[x], g[x], b[x]) : colour_cb_p(r[x], g[x], b[x]);
#pragma unroll
                for (int x = 0; x < 8; x++) F[k1w_idx(2, k, x, jj)] = colour_cr_p(r[x], g[x], b[x]);
            } else {
                /* lane (k, which, jc): chroma block jc of the strip = luma blocks 2jc, 2jc+1 */
                const int which = jj >> 2, j0 = 2 * (jj & 3);
                uint4 ca[3], cb[3];
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const uint2 l0 = ld8(c, k, j0), l1 = ld8(c, k, j0 + 1);
                    const uint2 h0 = ld8(c, 7 - k, j0), h1 = ld8(c, 7 - k, j0 + 1);
                    ca[c] = make_uint4(l0.x, l0.y, l1.x, l1.y);
                    cb[c] = make_uint4(h0.x, h0.y, h1.x, h1.y);
                }
                k1_pairsums16(ca[0], cb[0], r);
                k1_pairsums16(ca[1], cb[1], g);
                k1_pairsums16(ca[2], cb[2], b);
                if (which == 0) {
#pragma unroll
                    for (int x = 0; x < 8; x++)
                        F[k1w_idx(1, k, x, jj)] = STD ? colour_cb2_std_p(r[x], g[x], b[x]) : colour_cb2_p(r[x], g[x], b[x]);
                } else {
#pragma unroll
                    for (int x = 0; x < 8; x++) F[k1w_idx(1, k, x, jj)] = colour_cr2_p(r[x], g[x], b[x]);
                }
            }
        }
        __syncwarp();

        /* ---- phase 2 ---- */
        uint32_t o[32];
        if (lane < NACT) {
            const int c = lane >> 3;
            f32x2 P[4][8], Q[8][4];
#pragma unroll
            for (int kk = 0; kk < 4; kk++)
#pragma unroll
                for (int x = 0; x < 8; x++) P[kk][x] = F[k1w_idx(c, kk, x, jj)];
            fdct8x8_p(P, Q);
            const float4 *m4 = reinterpret_cast<const float4 *>(S.mul[c == 0 ? 0 : 1]);
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
        __syncwarp(); /* every lane has its samples in registers: the scratch can be overwritten */
        if (lane < NACT) {
#pragma unroll
            for (int j = 0; j < 8; j++) rows[lane * 9 + j] = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
        }
        __syncwarp();

        /* ---- phase 3: row r = c*8 + j; a component's 8 rows are 1 KB contiguous in HBM ---- */
        {
            int16_t *zy = p.zz_y + (long long)tl.frame * p.zz_stride;
            int16_t *zcb = p.zz_cb + (long long)tl.frame * p.zz_stride;
            int16_t *zcr = p.zz_cr + (long long)tl.frame * p.zz_stride;
            const int ch = lane & 7, rsub = lane >> 3;
#pragma unroll
            for (int i = 0; i < NACT / 4; i++) {
                const int row = i * 4 + rsub, c = row >> 3, j = row & 7;
                uint4 *dst;
                bool ok;
                if (RATIO == 0 || c == 0) {
                    const int bx = tl.tb0 + 8 * w + j;
                    ok = bx < p.BW;
                    int16_t *base = (c == 0) ? zy : (c == 1 ? zcb : zcr);
                    dst = reinterpret_cast<uint4 *>(base + ((long long)tl.by * p.BW + bx) * 64) + ch;
                } else {
                    /* 4:2:2 chroma: rows 8..11 Cb, 12..15 Cr */
                    const int cbw = p.BW >> 1, cx = ((tl.tb0 + 8 * w) >> 1) + (j & 3);
                    ok = cx < cbw;
                    int16_t *base = (j < 4) ? zcb : zcr;
                    dst = reinterpret_cast<uint4 *>(base + ((long long)tl.by * cbw + cx) * 64) + ch;
                }
                if (ok) *dst = rows[row * 9 + ch];
            }
        }
        __syncwarp(); /* the scratch is free for the next tile's phase 1 */
    }
}
#endif /* JPGB200_AB */

}  // namespace jpgb200
#endif
