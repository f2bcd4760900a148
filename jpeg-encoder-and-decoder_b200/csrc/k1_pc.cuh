/*
 * K1, producer/consumer form (k1_transform_pc): PIPELINES of eight specialised warps, no CTA-wide barrier --
 * every hand-over is an mbarrier.  One CTA of sixteen warps per SM runs two pipelines (K1C_DUAL, below).
 *
 *   convert warps 0..3  (k = role)  fetch rows k and 7-k of the tile's 32 blocks two tiles ahead
 *                       (cp.async, private ring), turn them into FP32 row pairs (PRMT + FADD2,
 *                       FFMA2) and fill sample buffer b          -> arrive  samples_full[b]
 *   transform warps 4..6 (c = role - 4)  wait samples_full[b], pull their 32 blocks (one per
 *                       lane, 64 samples) into registers          -> arrive  samples_free[b]
 *                       packed FDCT, FFMA2 quantise, zig-zag pack; wait coef_free[o], store the
 *                       rows (128-byte swizzle)                   -> arrive  coef_full[o]
 *   store warp 7        wait coef_full[o], three TMA tensor stores (32 blocks x 128 B per
 *                       component), wait until shared memory is read -> arrive  coef_free[o]
 *
 * Sample and coefficient buffers are double; because a transform warp copies its samples into
 * registers before it computes, the convert warps run one to two tiles ahead and a transform warp
 * goes from one tile's arithmetic straight into the next: the FMA pipes are the resource this
 * form tries to keep busy (the CTA-barrier and warp-autonomous forms leave them idle 45 % of the
 * time, profiles/README.md).  Moving the FDCT's row pass (all of it: 79.7 us; its first twelve
 * operations: 78.6 us) into the convert warps, which own whole rows, was measured against doing
 * it all in the transform warps (76.5 us): both roles are already equally busy.  Three CTAs per SM
 * (two convert warps with two row pairs each, one sample buffer, 96 registers) ran at 95 us.
 *
 * WHERE a warp runs matters more than any of that.  A warp is bound to one of the SM's four schedulers (and its
 * FMA pipe) for life: warp q of a CTA to scheduler q % 4.  As two CTAs of eight warps, both pipelines of an SM put
 * their three transform warps on schedulers 0..2 and their store warps on scheduler 3: per pair of tiles the FMA
 * pipes of schedulers 0..2 carry 1472 cycles of packed arithmetic each, scheduler 3's 384 (ncu: issue active
 * 46-61 % per scheduler, profiles/r2_final_ncu_full.csv) -- 76.4 us per 8K frame.  Measured on the way
 * (profiles/README.md): two transform warps of the SAME pipeline on one scheduler are ruinous (they work in phase:
 * 90-95 us), one CTA of sixteen warps with the two-CTA layout is 72.5 us, and the layout of K1C_DUAL_MAP below --
 * no scheduler holds two transform warps of one pipeline, the two that hold two transform warps (one of each
 * pipeline) hold one convert warp, the other two hold three -- is 68.5 us: 1120 / 1280 / 1280 / 1120 cycles.
 */
#ifndef JPGB200_K1_PC_CUH
#define JPGB200_K1_PC_CUH

#include "k1_pipe.cuh"

namespace jpgb200 {

#define K1C_THREADS 256
#ifndef K1C_STAGES
#define K1C_STAGES 3
#endif
#ifndef K1C_SLEEP_NS
#define K1C_SLEEP_NS 200
#endif
#ifndef K1C_MIN_CTAS
#define K1C_MIN_CTAS 2
#endif
/* K1C_DUAL: ONE CTA of sixteen warps per SM running the two eight-warp pipelines that otherwise live in two CTAs.
 * Same code, same tiles (pipeline q of CTA x takes the tile sequence of CTA 2x + q of the two-CTA form), same shared
 * memory per pipeline -- but the place of every warp on the SM's four schedulers (warp q -> scheduler q % 4) is now
 * ours to choose: with two CTAs both put their three transform warps on schedulers 0..2 (FMA-pipe cycles per tile
 * pair 1472 / 1472 / 1472 / 384); here it is 1120 / 1280 / 1280 / 1120 with the two pipelines' transform warps never
 * more than one of each per scheduler. */
#ifndef K1C_DUAL
#define K1C_DUAL 1 /* measured: 68.5-68.9 us per 8K frame against 76.4 us for two CTAs of eight warps (-DK1C_DUAL=0) */
#endif
#if K1C_DUAL
#define K1C_CTA_THREADS 512
#define K1C_LB_CTAS 1
#define K1C_PIPES 2
/* nibble q = (pipeline << 3) | role of warp q; roles 0..3 convert warp k, 4..6 transform warp c, 7 store warp
 *   S0: q0 = A.C0  q4 = A.C1  q8  = B.C0  q12 = A.T0      S1: q1 = A.St  q5 = A.C2  q9  = B.T0  q13 = A.T1
 *   S2: q2 = B.St  q6 = B.C1  q10 = B.T1  q14 = A.T2      S3: q3 = A.C3  q7 = B.C2  q11 = B.C3  q15 = B.T2 */
#ifndef K1C_DUAL_MAP
#define K1C_DUAL_MAP 0xE654BDC8A9213F70ull
#endif
/* 4:2:2 has two transform warps per pipeline (luma; Cb + Cr), role 6 idles: one transform warp and two convert warps
 * (one of each pipeline) per scheduler
 *   S0: q0 = A.C0  q4 = B.C0  q8  = A.St  q12 = A.T0      S1: q1 = A.C1  q5 = B.C1  q9  = (A.6)  q13 = A.T1
 *   S2: q2 = A.C2  q6 = B.C2  q10 = B.St  q14 = B.T0      S3: q3 = A.C3  q7 = B.C3  q11 = (B.6)  q15 = B.T1 */
#ifndef K1C_DUAL_MAP_422
#define K1C_DUAL_MAP_422 0xDC54EF67BA983210ull
#endif
#else
#define K1C_CTA_THREADS K1C_THREADS
#define K1C_LB_CTAS K1C_MIN_CTAS
#define K1C_PIPES 1
#endif

struct alignas(1024) K1CSmem {
    K1Out coef[2][3];                       /* 24 KB, 1 KB aligned (TMA swizzle) */
    K1F samples[2][3];                      /* 48 KB */
    uint8_t raw[4][K1C_STAGES][3][2][K1_ROW_BYTES]; /* [convert warp][slot][plane][row k | row 7-k][x]  18 KB */
    float mul[2][64];
    unsigned long long samples_full[2], samples_free[2], coef_full[2], coef_free[2];
};

/* wait with back-off: a warp that is ahead of its partners sleeps instead of spending the issue
 * slots of its scheduler on try_wait (the transform warps, which are the critical path, rarely
 * wait and use the plain spin) */
__device__ __forceinline__ void mbar_wait_relaxed(unsigned long long *bar, unsigned parity)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    unsigned done;
    for (;;) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
        if (done) break;
        __nanosleep(K1C_SLEEP_NS);
    }
}

__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
}

/* Tile coordinates of a persistent CTA's tile sequence t, t + D, t + 2D, ... (D = gridDim.x) without a division per
 * tile: the step is a fixed (block rows, tiles) pair.  Falls back to k1_tile when a step spans whole frames. */
struct K1TileIter {
    int frame, by, tbx, d_by, d_tb, tpr, rows;
    long long t;
    unsigned step;
    bool fast;
    __device__ __forceinline__ void init(const K1Params &p, long long t0, unsigned step_)
    {
        t = t0;
        step = step_;
        tpr = p.tiles_per_row;
        rows = p.H >> 3;
        const K1Tile o = k1_tile(p, t0);
        frame = o.frame;
        by = o.by;
        tbx = o.tb0 / K1_TILE_BLOCKS;
        d_by = (int)(step / (unsigned)tpr);
        d_tb = (int)(step - (unsigned)d_by * (unsigned)tpr);
        fast = d_by < rows - 1;
    }
    __device__ __forceinline__ K1Tile get() const
    {
        K1Tile o;
        o.frame = frame;
        o.by = by;
        o.tb0 = tbx * K1_TILE_BLOCKS;
        return o;
    }
    __device__ __forceinline__ void next(const K1Params &p)
    {
        t += step;
        if (fast) {
            tbx += d_tb;
            by += d_by;
            if (tbx >= tpr) {
                tbx -= tpr;
                by++;
            }
            if (by >= rows) { /* d_by < rows - 1: at most one frame boundary per step */
                by -= rows;
                frame++;
            }
        } else {
            const K1Tile o = k1_tile(p, t);
            frame = o.frame;
            by = o.by;
            tbx = o.tb0 / K1_TILE_BLOCKS;
        }
    }
};

/* convert warp k: its two rows of one tile, 16 bytes per copy (3 per lane).  The block of column
 * BW-1 is fetched one pixel row higher; row -1 of an image without halo is left
 * alone and replaced by the caller's out-of-bounds bytes at conversion time. */
template <bool STD>
__device__ __forceinline__ void k1c_fetch(const K1Params &p, uint8_t (*raw)[2][K1_ROW_BYTES], const K1Tile &tl, int k, int lane)
{
    const int part = lane & 15, sel = lane >> 4;
    const int y = sel ? 7 - k : k;
    const int bx = tl.tb0 + 2 * part;
    if (bx < p.BW) {
        const long long row = (long long)tl.by * 8 + y;
        const long long off = (long long)tl.frame * p.plane_stride + row * p.W + (long long)bx * 8;
        const int jl = STD ? -1 : p.BW - 1 - bx;
        if (jl == 0 || jl == 1) {
            const bool up_ok = (row - 1 >= 0) || p.has_halo;
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const uint8_t *src = (c == 0 ? p.r : (c == 1 ? p.g : p.b)) + off;
                uint8_t *dst = &raw[c][sel][part * 16];
                if (jl == 1) cp_async8(dst, src);
                if (up_ok) cp_async8(dst + 8 * jl, src + 8 * jl - p.W);
            }
        } else {
#pragma unroll
            for (int c = 0; c < 3; c++)
                cp_async16(&raw[c][sel][part * 16], (c == 0 ? p.r : (c == 1 ? p.g : p.b)) + off);
        }
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
}

template <int RATIO, bool STD = false>
__global__ void __launch_bounds__(K1C_CTA_THREADS, K1C_LB_CTAS) k1_transform_pc(const K1Params p, long long n_tiles,
                                                                                 const __grid_constant__ K1PStore st)
{
    extern __shared__ unsigned char k1_smem_raw[];
    unsigned char *k1_smem = k1_smem_raw + ((1024u - ((unsigned)__cvta_generic_to_shared(k1_smem_raw) & 1023u)) & 1023u);
    K1CSmem *const SP = reinterpret_cast<K1CSmem *>(k1_smem);
    const int tid = threadIdx.x, lane = tid & 31;
#if K1C_DUAL
    const unsigned role = (unsigned)(((RATIO == 1 ? K1C_DUAL_MAP_422 : K1C_DUAL_MAP) >> (4 * (tid >> 5))) & 15ull);
    const int w = (int)(role & 7u), pipe = (int)(role >> 3);
#else
    const int w = tid >> 5, pipe = 0;
#endif
    K1CSmem &S = SP[pipe];
    /* this pipeline's place in the persistent tile sequence: first tile and stride */
    const unsigned vb = blockIdx.x * K1C_PIPES + (unsigned)pipe, vg = gridDim.x * K1C_PIPES;
    constexpr int NT = (RATIO == 0) ? 3 : 2; /* transform warps with work */

    if (tid < 128 * K1C_PIPES) SP[tid >> 7].mul[(tid >> 6) & 1][tid & 63] = p.mul[(tid >> 6) & 1][tid & 63];
    if (blockIdx.x == 0)
        for (int i = tid; i < p.clear_n; i += K1C_CTA_THREADS) p.clear[i] = 0u;
    if (tid < K1C_PIPES) {
#pragma unroll
        for (int b = 0; b < 2; b++) {
            mbar_init(&SP[tid].samples_full[b], 4);
            mbar_init(&SP[tid].samples_free[b], NT);
            mbar_init(&SP[tid].coef_full[b], NT);
            mbar_init(&SP[tid].coef_free[b], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (w < 4) {
        /* ================= convert warp k ================= */
        /* (setmaxnreg.dec on this warpgroup -- 72 / 56 / 40 registers, to make room for a CTA of the neighbouring frames'
         * entropy kernels beside K1 -- was measured: K1 80.8 / 88.8 / 91.1 us against 78.3 us, the batch slower each time) */
        const int k = w;
        uint8_t(*ring)[3][2][K1_ROW_BYTES] = S.raw[k];
#pragma unroll
        for (int d = 0; d < K1C_STAGES - 1; d++) {
            const long long td = (long long)vb + (long long)d * vg;
            if (td < n_tiles)
                k1c_fetch<STD>(p, ring[d], k1_tile(p, td), k, lane);
            else
                asm volatile("cp.async.commit_group;\n" ::: "memory");
        }
        int it = 0;
        K1TileIter cur, pre; /* this tile and the one fetched now, K1C_STAGES - 1 tiles ahead */
        cur.init(p, vb, vg);
        pre.init(p, (long long)vb + (long long)(K1C_STAGES - 1) * vg, vg);
        for (long long t = vb; t < n_tiles; t += vg, it++, cur.next(p), pre.next(p)) {
            const int b = it & 1;
            const unsigned use = (unsigned)(it >> 1);
            const K1Tile tl = cur.get();
            {
                if (pre.t < n_tiles)
                    k1c_fetch<STD>(p, ring[(it + K1C_STAGES - 1) % K1C_STAGES], pre.get(), k, lane);
                else
                    asm volatile("cp.async.commit_group;\n" ::: "memory");
            }
            asm volatile("cp.async.wait_group %0;\n" ::"n"(K1C_STAGES - 1) : "memory");
            __syncwarp();
            const uint8_t(*raw)[2][K1_ROW_BYTES] = ring[it % K1C_STAGES];
            const int jl = STD ? -1 : p.BW - 1 - tl.tb0;
            const bool top_oob = (tl.by == 0 && !p.has_halo && k == 0);
            auto ld8 = [&](int c, int sel, int j) -> uint2 {
                if (top_oob && sel == 0 && j == jl) return ld_oob(p.oob + 8 * c);
                return *reinterpret_cast<const uint2 *>(&raw[c][sel][j * 8]);
            };
            uint2 pa[3], pb[3];
#pragma unroll
            for (int c = 0; c < 3; c++) {
                pa[c] = ld8(c, 0, lane);
                pb[c] = ld8(c, 1, lane);
            }
            uint4 ca[3], cb[3];
            if (RATIO == 1) {
                const int jc = lane & 15;
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const uint2 l0 = ld8(c, 0, 2 * jc), l1 = ld8(c, 0, 2 * jc + 1);
                    const uint2 h0 = ld8(c, 1, 2 * jc), h1 = ld8(c, 1, 2 * jc + 1);
                    ca[c] = make_uint4(l0.x, l0.y, l1.x, l1.y);
                    cb[c] = make_uint4(h0.x, h0.y, h1.x, h1.y);
                }
            }
            mbar_wait_relaxed(&S.samples_free[b], (use & 1u) ^ 1u); /* the transform warps have taken tile it-2 */
            k1_convert_full<RATIO, STD>(S.samples[b], k, lane, pa, pb);
            if (RATIO == 1) k1_convert_chroma422<STD>(S.samples[b], k, lane, ca, cb);
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.samples_full[b]);
        }
    } else if (w < 4 + NT) {
        /* ================= transform warp c ================= */
        const int c = w - 4;
        const float4 *m4 = reinterpret_cast<const float4 *>(S.mul[c == 0 ? 0 : 1]);
        int it = 0;
        for (long long t = vb; t < n_tiles; t += vg, it++) {
            const int b = it & 1;
            const unsigned use = (unsigned)(it >> 1);
            f32x2 P[4][8], Q[8][4];
            mbar_wait(&S.samples_full[b], use & 1u);
#pragma unroll
            for (int k = 0; k < 4; k++)
#pragma unroll
                for (int x = 0; x < 8; x++) P[k][x] = S.samples[b][c][k][x][lane];
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.samples_free[b]);
            fdct8x8_p(P, Q);
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
            mbar_wait(&S.coef_free[b], (use & 1u) ^ 1u); /* tile it-2 has left the buffer */
#pragma unroll
            for (int j = 0; j < 8; j++)
                S.coef[b][c][lane][j ^ (lane & 7)] = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); /* generic writes -> the TMA's proxy */
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.coef_full[b]);
        }
    } else if (w == 7) {
        /* ================= store warp ================= */
        int it = 0;
        K1TileIter cur;
        cur.init(p, vb, vg);
        for (long long t = vb; t < n_tiles; t += vg, it++, cur.next(p)) {
            const int b = it & 1;
            const unsigned use = (unsigned)(it >> 1);
            const K1Tile tl = cur.get();
            mbar_wait_relaxed(&S.coef_full[b], use & 1u);
            k1p_store<RATIO>(S.coef[b], p, st, tl, lane); /* returns when shared memory has been read */
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.coef_free[b]);
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

/* ---- K1, producer/consumer form with FOUR transform warps (k1_transform_pc4, 4:4:4) ---------------------------
 * k1_transform_pc gives each of its three transform warps one component of every tile: with two CTAs per SM that is
 * six transform warps on four schedulers (2, 2, 1, 1), and the store warp idles.  Here the tile's three components
 * are UNITS handed round-robin to four transform warps (unit u = 3 * tile + component goes to warp u % 4): eight
 * transform warps per SM, two per scheduler, each with 3/4 of a tile's arithmetic per tile time.  There is no store
 * warp: the warp that transformed a unit issues its TMA tensor store (one 32-block box) itself and hands the
 * coefficient buffer back -- after the next unit's samples are in its registers, when the store has long read
 * shared memory -- through one mbarrier per (buffer, component).
 * MEASURED (profiles/README.md, round 2): bit-identical, the four schedulers evenly loaded (issue active 58.4-59.7 %
 * against 45.6-60.6 %), but 79.9 us per 8K frame against 77.1 us: the component index is no longer a per-warp
 * constant (14 % more instructions) and the transform warps gain nothing from two more of their kind -- each one
 * runs at its own pace of ~3.6 cycles per instruction either way.  An A/B form (JPGB200_AB=1), not in the product. */
#ifdef JPGB200_AB
struct K1C4Smem {
    K1Out coef[2][3];
    K1F samples[2][3];
    uint8_t raw[4][K1C_STAGES][3][2][K1_ROW_BYTES];
    float mul[2][64];
    unsigned long long samples_full[2], samples_free[2], coef_free[2][3];
};

/* one component's 32 blocks of a tile cut short by the image's right edge: coalesced 16-byte stores by one warp */
__device__ __forceinline__ void k1c4_copyout(const K1Out &sOut, const K1Params &p, const K1Tile &tl, int c, int lane)
{
    int16_t *z = (c == 0 ? p.zz_y : (c == 1 ? p.zz_cb : p.zz_cr)) + (long long)tl.frame * p.zz_stride;
    uint4 *dst = reinterpret_cast<uint4 *>(z + ((long long)tl.by * p.BW + tl.tb0) * 64);
    const int ch = lane & 7, jsub = lane >> 3;
    const int nvalid = min(K1_TILE_BLOCKS, p.BW - tl.tb0);
#pragma unroll
    for (int t = 0; t < 8; t++) {
        const int j = t * 4 + jsub;
        if (j < nvalid) dst[j * 8 + ch] = sOut[j][ch ^ (j & 7)];
    }
}

template <bool STD = false>
__global__ void __launch_bounds__(K1C_THREADS, K1C_MIN_CTAS) k1_transform_pc4(const K1Params p, long long n_tiles,
                                                                               const __grid_constant__ K1PStore st)
{
    extern __shared__ unsigned char k1_smem_raw[];
    unsigned char *k1_smem = k1_smem_raw + ((1024u - ((unsigned)__cvta_generic_to_shared(k1_smem_raw) & 1023u)) & 1023u);
    K1C4Smem &S = *reinterpret_cast<K1C4Smem *>(k1_smem);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;

    if (tid < 128) S.mul[tid >> 6][tid & 63] = p.mul[tid >> 6][tid & 63];
    if (blockIdx.x == 0)
        for (int i = tid; i < p.clear_n; i += K1C_THREADS) p.clear[i] = 0u;
    if (tid == 0) {
#pragma unroll
        for (int b = 0; b < 2; b++) {
            mbar_init(&S.samples_full[b], 4);
            mbar_init(&S.samples_free[b], 3);
#pragma unroll
            for (int c = 0; c < 3; c++) mbar_init(&S.coef_free[b][c], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (w < 4) {
        /* ================= convert warp k: as in k1_transform_pc ================= */
        const int k = w;
        uint8_t(*ring)[3][2][K1_ROW_BYTES] = S.raw[k];
#pragma unroll
        for (int d = 0; d < K1C_STAGES - 1; d++) {
            const long long td = (long long)blockIdx.x + (long long)d * gridDim.x;
            if (td < n_tiles)
                k1c_fetch<STD>(p, ring[d], k1_tile(p, td), k, lane);
            else
                asm volatile("cp.async.commit_group;\n" ::: "memory");
        }
        int it = 0;
        for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x, it++) {
            const int b = it & 1;
            const unsigned use = (unsigned)(it >> 1);
            const K1Tile tl = k1_tile(p, t);
            {
                const long long tn = t + (long long)(K1C_STAGES - 1) * gridDim.x;
                if (tn < n_tiles)
                    k1c_fetch<STD>(p, ring[(it + K1C_STAGES - 1) % K1C_STAGES], k1_tile(p, tn), k, lane);
                else
                    asm volatile("cp.async.commit_group;\n" ::: "memory");
            }
            asm volatile("cp.async.wait_group %0;\n" ::"n"(K1C_STAGES - 1) : "memory");
            __syncwarp();
            const uint8_t(*raw)[2][K1_ROW_BYTES] = ring[it % K1C_STAGES];
            const int jl = STD ? -1 : p.BW - 1 - tl.tb0;
            const bool top_oob = (tl.by == 0 && !p.has_halo && k == 0);
            auto ld8 = [&](int c, int sel, int j) -> uint2 {
                if (top_oob && sel == 0 && j == jl) return ld_oob(p.oob + 8 * c);
                return *reinterpret_cast<const uint2 *>(&raw[c][sel][j * 8]);
            };
            uint2 pa[3], pb[3];
#pragma unroll
            for (int c = 0; c < 3; c++) {
                pa[c] = ld8(c, 0, lane);
                pb[c] = ld8(c, 1, lane);
            }
            mbar_wait_relaxed(&S.samples_free[b], (use & 1u) ^ 1u); /* the transform warps have taken tile it-2 */
            k1_convert_full<0, STD>(S.samples[b], k, lane, pa, pb);
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.samples_full[b]);
        }
    } else {
        /* ================= transform warp x: units x, x + 4, x + 8, ... ================= */
        const int x = w - 4;
        const long long my_tiles = (n_tiles - (long long)blockIdx.x + gridDim.x - 1) / gridDim.x;
        const int n_units = (int)(my_tiles * 3);
        int pend_b = -1, pend_c = 0; /* a TMA store of this warp that may still be reading shared memory */
        for (int u = x; u < n_units; u += 4) {
            const int it = (int)((unsigned)u / 3u);
            const int c = u - 3 * it;
            const int b = it & 1;
            const unsigned use = (unsigned)(it >> 1);
            const float4 *m4 = reinterpret_cast<const float4 *>(S.mul[c == 0 ? 0 : 1]);
            f32x2 P[4][8], Q[8][4];
            mbar_wait(&S.samples_full[b], use & 1u);
#pragma unroll
            for (int k = 0; k < 4; k++)
#pragma unroll
                for (int xx = 0; xx < 8; xx++) P[k][xx] = S.samples[b][c][k][xx][lane];
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&S.samples_free[b]);
                if (pend_b >= 0) {
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    mbar_arrive(&S.coef_free[pend_b][pend_c]);
                }
            }
            pend_b = -1;
            fdct8x8_p(P, Q);
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
            mbar_wait(&S.coef_free[b][c], (use & 1u) ^ 1u); /* tile it-2's unit has left the buffer */
#pragma unroll
            for (int j = 0; j < 8; j++)
                S.coef[b][c][lane][j ^ (lane & 7)] = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
            const K1Tile tl = k1_tile(p, (long long)blockIdx.x + (long long)it * gridDim.x);
            if (st.enabled && tl.tb0 + K1_TILE_BLOCKS <= p.BW) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); /* generic writes -> the TMA's proxy */
                __syncwarp();
                if (lane == 0) {
                    const long long ry = (long long)tl.frame * st.rows_per_frame + (long long)tl.by * p.BW + tl.tb0;
                    tma_store_2d(&st.box32, &S.coef[b][c], 0, (int)(ry + (c == 0 ? 0ll : (c == 1 ? st.row_cb : st.row_cr))));
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                pend_b = b;
                pend_c = c;
            } else {
                __syncwarp();
                k1c4_copyout(S.coef[b][c], p, tl, c, lane);
                __syncwarp();
                if (lane == 0) mbar_arrive(&S.coef_free[b][c]);
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}
#endif /* JPGB200_AB */

/* ---- K1 with TWELVE warps per CTA (k1_transform_pc12, 4:4:4) -------------------------------------------------
 * In k1_transform_pc both roles -- four convert warps with two pixel rows each, three transform warps with one
 * component each -- need about the same ~1300 cycles per tile, each warp at its own dependency-limited pace; giving only
 * the transform role a fourth warp (k1_transform_pc4) therefore changes nothing.  Here BOTH roles get more warps:
 *   convert warps 0..7   warp k owns pixel row k of the tile's 32 blocks: lane j fetches the 8 bytes of its block's
 *                        row from each plane itself (cp.async, private ring: no hand-over inside the warp), converts
 *                        them as four pairs of neighbouring pixels (the same IEEE operations per sample) and stores
 *                        the eight samples of each component into half k < 4 ? .x : .y of the row pair min(k, 7-k)
 *   transform warps 8..11 the tile's components as units handed round-robin (unit u = 3 * tile + component -> warp u % 4),
 *                        each warp issuing the TMA tensor store of its own unit (as in k1_transform_pc4)
 * 384 threads, two CTAs per SM: the launch bound is 80 registers per thread and setmaxnreg re-divides them -- the two
 * convert warpgroups keep 48, the transform warpgroup takes 144 (no spills; 56 / 128 spills 32 registers).
 * MEASURED (profiles/README.md, round 2): bit-identical and SLOWER, 85.3 us per 8K frame against 77.3 us.  The convert
 * warps now wait for the transform warps (20 % of all stall samples), but the kernel executes 60.4 M warp
 * instructions instead of 42.6 M (eight warps repeat the per-tile address and tile-index arithmetic, the component
 * index of the transform warps is no longer a constant), issue is active 68 % of the time, and the 4-byte sample
 * stores collide two ways in shared memory (4.1 M bank conflicts against 1.0 M, LSU wavefronts at 67 % of peak).
 * K1 behaves like an instruction-issue-bound kernel as soon as more warps are given to it: the 47 % of idle issue
 * slots of the eight-warp form are not there to be filled.  An A/B form (JPGB200_AB=1, form 8), not in the product. */
#ifdef JPGB200_AB
#define K1D_THREADS 384
#define K1D_REG_CONVERT 48
#define K1D_REG_TRANSFORM 144

struct K1DSmem {
    K1Out coef[2][3];
    K1F samples[2][3];
    uint8_t raw[8][K1C_STAGES][3][K1_ROW_BYTES]; /* [convert warp = row][slot][plane][x] */
    float mul[2][64];
    unsigned long long samples_full[2], samples_free[2], coef_free[2][3];
};

/* convert warp k, lane j: the 8 bytes of row k of block tb0 + j from every plane (the block of column BW-1 one
 * pixel row higher; row -1 of an image without halo is left alone: the caller's out-of-bounds bytes) */
template <bool STD>
__device__ __forceinline__ void k1d_fetch(const K1Params &p, uint8_t (*raw)[K1_ROW_BYTES], const K1Tile &tl, int k, int lane)
{
    const int bx = tl.tb0 + lane;
    if (bx < p.BW) {
        long long row = (long long)tl.by * 8 + k;
        bool ok = true;
        if (!STD && bx == p.BW - 1) {
            row -= 1;
            ok = (row >= 0) || p.has_halo;
        }
        if (ok) {
            const long long off = (long long)tl.frame * p.plane_stride + row * p.W + (long long)bx * 8;
#pragma unroll
            for (int c = 0; c < 3; c++) cp_async8(&raw[c][lane * 8], (c == 0 ? p.r : (c == 1 ? p.g : p.b)) + off);
        }
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
}

/* bytes 2i, 2i+1 of a word pair (lo, hi = the row's 8 bytes) -> two exact FP32 values */
__device__ __forceinline__ f32x2 k1d_pair(const uint2 w, int i)
{
    const uint32_t src = (i < 2) ? w.x : w.y;
    const int b0 = (2 * i) & 3;
    f32x2 v;
    v.x = __uint_as_float(__byte_perm(JPGB_BYTE_MAGIC, src, 0x3104u + (uint32_t)b0));
    v.y = __uint_as_float(__byte_perm(JPGB_BYTE_MAGIC, src, 0x3104u + (uint32_t)b0 + 1u));
    return p_add(v, p_bc(-8388608.0f));
}

/* one component's 32 blocks of a tile cut short by the image's right edge: coalesced 16-byte stores by one warp */
__device__ __forceinline__ void k1d_copyout(const K1Out &sOut, const K1Params &p, const K1Tile &tl, int c, int lane)
{
    int16_t *z = (c == 0 ? p.zz_y : (c == 1 ? p.zz_cb : p.zz_cr)) + (long long)tl.frame * p.zz_stride;
    uint4 *dst = reinterpret_cast<uint4 *>(z + ((long long)tl.by * p.BW + tl.tb0) * 64);
    const int ch = lane & 7, jsub = lane >> 3;
    const int nvalid = min(K1_TILE_BLOCKS, p.BW - tl.tb0);
#pragma unroll
    for (int t = 0; t < 8; t++) {
        const int j = t * 4 + jsub;
        if (j < nvalid) dst[j * 8 + ch] = sOut[j][ch ^ (j & 7)];
    }
}

template <bool STD = false>
__global__ void __launch_bounds__(K1D_THREADS, K1C_MIN_CTAS) k1_transform_pc12(const K1Params p, long long n_tiles,
                                                                                const __grid_constant__ K1PStore st)
{
    extern __shared__ unsigned char k1_smem_raw[];
    unsigned char *k1_smem = k1_smem_raw + ((1024u - ((unsigned)__cvta_generic_to_shared(k1_smem_raw) & 1023u)) & 1023u);
    K1DSmem &S = *reinterpret_cast<K1DSmem *>(k1_smem);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;

    if (tid < 128) S.mul[tid >> 6][tid & 63] = p.mul[tid >> 6][tid & 63];
    if (blockIdx.x == 0)
        for (int i = tid; i < p.clear_n; i += K1D_THREADS) p.clear[i] = 0u;
    if (tid == 0) {
#pragma unroll
        for (int b = 0; b < 2; b++) {
            mbar_init(&S.samples_full[b], 8);
            mbar_init(&S.samples_free[b], 3);
#pragma unroll
            for (int c = 0; c < 3; c++) mbar_init(&S.coef_free[b][c], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (w < 8) {
        /* ================= convert warp k: one pixel row ================= */
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(K1D_REG_CONVERT));
        const int k = w;
        const int kk = k < 4 ? k : 7 - k, half = k < 4 ? 0 : 1;
        uint8_t(*ring)[3][K1_ROW_BYTES] = S.raw[k];
#pragma unroll
        for (int d = 0; d < K1C_STAGES - 1; d++) {
            const long long td = (long long)blockIdx.x + (long long)d * gridDim.x;
            if (td < n_tiles)
                k1d_fetch<STD>(p, ring[d], k1_tile(p, td), k, lane);
            else
                asm volatile("cp.async.commit_group;\n" ::: "memory");
        }
        int it = 0;
        for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x, it++) {
            const int b = it & 1;
            const unsigned use = (unsigned)(it >> 1);
            const K1Tile tl = k1_tile(p, t);
            {
                const long long tn = t + (long long)(K1C_STAGES - 1) * gridDim.x;
                if (tn < n_tiles)
                    k1d_fetch<STD>(p, ring[(it + K1C_STAGES - 1) % K1C_STAGES], k1_tile(p, tn), k, lane);
                else
                    asm volatile("cp.async.commit_group;\n" ::: "memory");
            }
            asm volatile("cp.async.wait_group %0;\n" ::"n"(K1C_STAGES - 1) : "memory"); /* this lane's own copies */
            const uint8_t(*raw)[K1_ROW_BYTES] = ring[it % K1C_STAGES];
            uint2 px[3];
            const bool oob = !STD && (tl.by == 0 && !p.has_halo && k == 0 && tl.tb0 + lane == p.BW - 1);
#pragma unroll
            for (int c = 0; c < 3; c++) px[c] = oob ? ld_oob(p.oob + 8 * c) : *reinterpret_cast<const uint2 *>(&raw[c][lane * 8]);
            mbar_wait_relaxed(&S.samples_free[b], (use & 1u) ^ 1u); /* the transform warps have taken tile it-2 */
            float *sy = reinterpret_cast<float *>(&S.samples[b][0][kk][0][lane]) + half;
            float *scb = reinterpret_cast<float *>(&S.samples[b][1][kk][0][lane]) + half;
            float *scr = reinterpret_cast<float *>(&S.samples[b][2][kk][0][lane]) + half;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const f32x2 r = k1d_pair(px[0], i), g = k1d_pair(px[1], i), bl = k1d_pair(px[2], i);
                const f32x2 y = colour_y_p(r, g, bl);
                const f32x2 cb = STD ? colour_cb_std_p(r, g, bl) : colour_cb_p(r, g, bl);
                const f32x2 cr = colour_cr_p(r, g, bl);
                /* [x][slot] of float2: sample x of this row sits 64 floats after sample x - 1 */
                sy[(2 * i) * 64] = y.x;
                sy[(2 * i + 1) * 64] = y.y;
                scb[(2 * i) * 64] = cb.x;
                scb[(2 * i + 1) * 64] = cb.y;
                scr[(2 * i) * 64] = cr.x;
                scr[(2 * i + 1) * 64] = cr.y;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.samples_full[b]);
        }
    } else {
        /* ================= transform warp x: units x, x + 4, x + 8, ... ================= */
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(K1D_REG_TRANSFORM));
        const int x = w - 8;
        const long long my_tiles = (n_tiles - (long long)blockIdx.x + gridDim.x - 1) / gridDim.x;
        const int n_units = (int)(my_tiles * 3);
        int pend_b = -1, pend_c = 0; /* a TMA store of this warp that may still be reading shared memory */
        for (int u = x; u < n_units; u += 4) {
            const int it = (int)((unsigned)u / 3u);
            const int c = u - 3 * it;
            const int b = it & 1;
            const unsigned use = (unsigned)(it >> 1);
            const float4 *m4 = reinterpret_cast<const float4 *>(S.mul[c == 0 ? 0 : 1]);
            f32x2 P[4][8], Q[8][4];
            mbar_wait(&S.samples_full[b], use & 1u);
#pragma unroll
            for (int k = 0; k < 4; k++)
#pragma unroll
                for (int xx = 0; xx < 8; xx++) P[k][xx] = S.samples[b][c][k][xx][lane];
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&S.samples_free[b]);
                if (pend_b >= 0) {
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    mbar_arrive(&S.coef_free[pend_b][pend_c]);
                }
            }
            pend_b = -1;
            fdct8x8_p(P, Q);
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
            mbar_wait(&S.coef_free[b][c], (use & 1u) ^ 1u); /* tile it-2's unit has left the buffer */
#pragma unroll
            for (int j = 0; j < 8; j++)
                S.coef[b][c][lane][j ^ (lane & 7)] = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
            const K1Tile tl = k1_tile(p, (long long)blockIdx.x + (long long)it * gridDim.x);
            if (st.enabled && tl.tb0 + K1_TILE_BLOCKS <= p.BW) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); /* generic writes -> the TMA's proxy */
                __syncwarp();
                if (lane == 0) {
                    const long long ry = (long long)tl.frame * st.rows_per_frame + (long long)tl.by * p.BW + tl.tb0;
                    tma_store_2d(&st.box32, &S.coef[b][c], 0, (int)(ry + (c == 0 ? 0ll : (c == 1 ? st.row_cb : st.row_cr))));
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                pend_b = b;
                pend_c = c;
            } else {
                __syncwarp();
                k1d_copyout(S.coef[b][c], p, tl, c, lane);
                __syncwarp();
                if (lane == 0) mbar_arrive(&S.coef_free[b][c]);
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}
#endif /* JPGB200_AB */

/* ---- K1, producer/consumer form fed by 2-D TMA tile loads (k1_transform_pct) --------------------------------
 * k1_transform_pc's convert warps fetch their own two rows of the tile with cp.async: 27 LDGSTS and their
 * address arithmetic per lane and tile, issued on the schedulers the transform warps need.  Here the (otherwise
 * idle) store warp issues ONE tensor load per plane and tile -- cp.async.bulk.tensor.2d, box 256 bytes x 8 rows,
 * SASS UTMALDG -- into a CTA-shared ring; completion is counted in bytes on an mbarrier, the convert warps hand
 * the slot back through a second one.  The block of column BW-1 is read one pixel row higher: the
 * tile that holds it gets a second, 16-byte-wide box per plane starting one row up; rows outside the image (row -1
 * of an image without halo) are zero-filled by the TMA unit and replaced by the caller's out-of-bounds bytes.
 * MEASURED (profiles/README.md): 83.5 us per 8K frame against 76.7 us for the cp.async-fed form -- issued by
 * convert warp 0 instead: 87.8 us, with a four-slot ring: 86.3 us.  The shared ring couples the four convert
 * warps through two mbarriers per tile where the cp.async form gives every warp a private ring and no hand-over
 * at all.  Selectable (jpgb200_set_plain_transform(ctx, 6)), bit-identical, not the default. */
#ifndef K1T_STAGES
#define K1T_STAGES 3
#endif

struct K1TLoad {
    CUtensorMap plane[3]; /* [rows of all frames][W] u8, box 256 x 8 */
    CUtensorMap alt[3];   /* same tensor, box 16 x 8 */
    int rows_per_frame;   /* plane_stride / W */
    int y_bias;           /* 1 when the tensors start one row above the band (halo) */
};

struct K1TSmem {
    K1Out coef[2][3];                       /* 24 KB, 1 KB aligned (TMA swizzle) */
    K1F samples[2][3];                      /* 48 KB */
    uint8_t raw[K1T_STAGES][3][8][K1_ROW_BYTES]; /* 18 KB, [slot][plane][row][x]: 128-byte aligned TMA destinations */
    uint8_t alt[K1T_STAGES][3][8][16];
    float mul[2][64];
    unsigned long long samples_full[2], samples_free[2], coef_full[2], coef_free[2];
    unsigned long long raw_full[K1T_STAGES], raw_free[K1T_STAGES];
};

__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *tm, int c0, int c1, unsigned long long *bar)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst), b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(d),
                 "l"(tm), "r"(c0), "r"(c1), "r"(b)
                 : "memory");
}

/* one lane: arm the slot's mbarrier and issue the tile's tensor loads */
template <bool STD>
__device__ __forceinline__ void k1t_issue(const K1Params &p, const K1TLoad &ld, K1TSmem &S, int s, const K1Tile &tl)
{
    const bool has_last = !STD && (tl.tb0 + K1_TILE_BLOCKS >= p.BW);
    const int y = tl.frame * ld.rows_per_frame + tl.by * 8 + ld.y_bias;
    mbar_expect_tx(&S.raw_full[s], 3u * 8u * K1_ROW_BYTES + (has_last ? 3u * 8u * 16u : 0u));
#pragma unroll
    for (int c = 0; c < 3; c++) {
        tma_load_2d(&S.raw[s][c][0][0], &ld.plane[c], tl.tb0 * 8, y, &S.raw_full[s]);
        if (has_last) tma_load_2d(&S.alt[s][c][0][0], &ld.alt[c], p.W - 16, y - 1, &S.raw_full[s]);
    }
}

template <int RATIO, bool STD = false>
__global__ void __launch_bounds__(K1C_THREADS, K1C_MIN_CTAS) k1_transform_pct(const K1Params p, long long n_tiles,
                                                                               const __grid_constant__ K1PStore st,
                                                                               const __grid_constant__ K1TLoad ld)
{
    extern __shared__ unsigned char k1_smem_raw[];
    unsigned char *k1_smem = k1_smem_raw + ((1024u - ((unsigned)__cvta_generic_to_shared(k1_smem_raw) & 1023u)) & 1023u);
    K1TSmem &S = *reinterpret_cast<K1TSmem *>(k1_smem);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    constexpr int NT = (RATIO == 0) ? 3 : 2; /* transform warps with work */

    if (tid < 128) S.mul[tid >> 6][tid & 63] = p.mul[tid >> 6][tid & 63];
    if (blockIdx.x == 0)
        for (int i = tid; i < p.clear_n; i += K1C_THREADS) p.clear[i] = 0u;
    if (tid == 0) {
#pragma unroll
        for (int b = 0; b < 2; b++) {
            mbar_init(&S.samples_full[b], 4);
            mbar_init(&S.samples_free[b], NT);
            mbar_init(&S.coef_full[b], NT);
            mbar_init(&S.coef_free[b], 1);
        }
#pragma unroll
        for (int s = 0; s < K1T_STAGES; s++) {
            mbar_init(&S.raw_full[s], 1);
            mbar_init(&S.raw_free[s], 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (w < 4) {
        /* ================= convert warp k ================= */
        const int k = w;
        int it = 0;
        for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x, it++) {
            const int b = it & 1;
            const unsigned use = (unsigned)(it >> 1);
            const int s = it % K1T_STAGES;

            const unsigned suse = (unsigned)(it / K1T_STAGES);
            const K1Tile tl = k1_tile(p, t);
            mbar_wait(&S.raw_full[s], suse & 1u);
            const int jl = STD ? -1 : p.BW - 1 - tl.tb0;
            const bool top_oob = (tl.by == 0 && !p.has_halo && k == 0);
            auto ld8 = [&](int c, int sel, int j) -> uint2 {
                const int rr = sel ? 7 - k : k;
                if (j == jl) {
                    if (top_oob && sel == 0) return ld_oob(p.oob + 8 * c);
                    return *reinterpret_cast<const uint2 *>(&S.alt[s][c][rr][8]);
                }
                return *reinterpret_cast<const uint2 *>(&S.raw[s][c][rr][j * 8]);
            };
            uint2 pa[3], pb[3];
#pragma unroll
            for (int c = 0; c < 3; c++) {
                pa[c] = ld8(c, 0, lane);
                pb[c] = ld8(c, 1, lane);
            }
            uint4 ca[3], cb[3];
            if (RATIO == 1) {
                const int jc = lane & 15;
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const uint2 l0 = ld8(c, 0, 2 * jc), l1 = ld8(c, 0, 2 * jc + 1);
                    const uint2 h0 = ld8(c, 1, 2 * jc), h1 = ld8(c, 1, 2 * jc + 1);
                    ca[c] = make_uint4(l0.x, l0.y, l1.x, l1.y);
                    cb[c] = make_uint4(h0.x, h0.y, h1.x, h1.y);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.raw_free[s]); /* this warp's rows are in registers */
            mbar_wait_relaxed(&S.samples_free[b], (use & 1u) ^ 1u); /* the transform warps have taken tile it-2 */
            k1_convert_full<RATIO, STD>(S.samples[b], k, lane, pa, pb);
            if (RATIO == 1) k1_convert_chroma422<STD>(S.samples[b], k, lane, ca, cb);
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.samples_full[b]);
        }
    } else if (w < 4 + NT) {
        /* ================= transform warp c ================= */
        const int c = w - 4;
        const float4 *m4 = reinterpret_cast<const float4 *>(S.mul[c == 0 ? 0 : 1]);
        int it = 0;
        for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x, it++) {
            const int b = it & 1;
            const unsigned use = (unsigned)(it >> 1);
            f32x2 P[4][8], Q[8][4];
            mbar_wait(&S.samples_full[b], use & 1u);
#pragma unroll
            for (int k = 0; k < 4; k++)
#pragma unroll
                for (int x = 0; x < 8; x++) P[k][x] = S.samples[b][c][k][x][lane];
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.samples_free[b]);
            fdct8x8_p(P, Q);
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
            mbar_wait(&S.coef_free[b], (use & 1u) ^ 1u); /* tile it-2 has left the buffer */
#pragma unroll
            for (int j = 0; j < 8; j++)
                S.coef[b][c][lane][j ^ (lane & 7)] = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); /* generic writes -> the TMA's proxy */
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.coef_full[b]);
        }
    } else if (w == 7) {
        /* ================= load + store warp ================= */
        if (lane == 0) {
#pragma unroll
            for (int d = 0; d < K1T_STAGES - 1; d++) {
                const long long td = (long long)blockIdx.x + (long long)d * gridDim.x;
                if (td < n_tiles) k1t_issue<STD>(p, ld, S, d, k1_tile(p, td));
            }
        }
        int it = 0;
        for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x, it++) {
            const int b = it & 1;
            const unsigned use = (unsigned)(it >> 1);
            const K1Tile tl = k1_tile(p, t);
            {
                /* the load of tile it + STAGES - 1 goes into the slot tile it - 1 used */
                const long long tn = t + (long long)(K1T_STAGES - 1) * gridDim.x;
                const int itn = it + K1T_STAGES - 1, sn = itn % K1T_STAGES;
                if (tn < n_tiles && lane == 0) {
                    if (itn >= K1T_STAGES) mbar_wait(&S.raw_free[sn], (unsigned)(itn / K1T_STAGES - 1) & 1u);
                    k1t_issue<STD>(p, ld, S, sn, k1_tile(p, tn));
                }
                __syncwarp();
            }
            mbar_wait_relaxed(&S.coef_full[b], use & 1u);
            k1p_store<RATIO>(S.coef[b], p, st, tl, lane); /* returns when shared memory has been read */
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.coef_free[b]);
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

}  // namespace jpgb200
#endif
