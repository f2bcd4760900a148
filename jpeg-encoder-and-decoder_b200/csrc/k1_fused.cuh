/*
 * K1E (k1e_fused): the transform AND the chunk encode in one persistent kernel -- the quantised
 * coefficients never leave the SM.
 *
 * What the reference chains as dct -> quantise -> zig_zag -> dpcm -> huffman_encode
 * (src/jpg_encode.c:36-50, src/huffman.c:182-235) over whole-image arrays is one pass here:
 *
 *   warps 0..3   convert   (k1_pc.cuh)  rows k and 7-k of the tile -> FP32 row pairs
 *   warps 4..6   transform (k1_pc.cuh)  FDCT, quantise, zig-zag -> coefficient tile in shared memory
 *                                        (+ the tile's last DC per component, published for the
 *                                        next tile's first DC difference)
 *   warp  7      store     only when the caller wants the coefficients in HBM (parity tap)
 *   warps 8..11  entropy   the tile is a CHUNK of the entropy stage (32 MCUs for 4:4:4, 16 for
 *                          4:2:2): every block is coded once into a thread-private region, the
 *                          lengths are scanned in stream order, the codes relocated into the chunk
 *                          stream and copied to the chunk's slot; chunk total -> two-level sums.
 *                          KF/KS (tail.cuh) turn the chunk streams into the stuffed JFIF scan.
 *
 * The three warpgroups re-divide the CTA's registers with setmaxnreg (launch bound 80 per thread:
 * 2 CTAs x 384 threads per SM): the transform warps, which hold a whole 8x8 block per lane, grow,
 * the other two groups shrink.
 *
 * DC prediction across tiles: tile t's first block of a component needs the DC of tile t-1's last
 * block.  Tiles are dealt round-robin to co-resident persistent CTAs, so tile t-1 is transformed by
 * the neighbouring CTA at the same time; its transform warps publish {epoch, DC} words and the
 * entropy warps of tile t read them (one hop, no chain: a DC does not depend on earlier tiles).
 *
 * Same bytes as K1 -> KE: the chunk boundaries differ (tiles instead of 64 MCUs), the concatenated
 * stream does not.
 */
#ifndef JPGB200_K1_FUSED_CUH
#define JPGB200_K1_FUSED_CUH

#include "entropy.cuh"
#include "k1_pc.cuh"

namespace jpgb200 {

#define K1E_PRIV 10                /* words of private stream per entry */
#define K1E_PSTRIDE (K1E_PRIV + 1) /* odd: lanes hit distinct banks */
#define K1E_STAGE_WORDS 768        /* chunk stream staging: 3 KB (a natural 8K tile needs ~0.5 KB) */

template <int RATIO>
struct K1ECfg {
    static const int NT = (RATIO == 0) ? 3 : 2;        /* transform warps with work */
    static const int NEW = (RATIO == 0) ? 4 : 3;       /* entropy warps with work */
    static const int NE = 32 * NEW;                    /* entropy threads = entries per full tile */
    static const int BPT = (RATIO == 0) ? 96 : 64;     /* blocks per tile */
    /* entropy groups: each is a warpgroup that codes every NG-th tile of the CTA on its own (a tile's
     * coding is a long dependent chain per lane: what hides its latency is a second tile in flight) */
    static const int NG = (RATIO == 0) ? 2 : 1;
    static const int THREADS = 256 + 128 * NG;
    /* registers per thread after setmaxnreg; the launch bound is 65536 / (2 * THREADS) rounded down to 8 */
    static const int REG_CONVERT = (RATIO == 0) ? 48 : 72;
    static const int REG_TRANSFORM = (RATIO == 0) ? 96 : 112;
    static const int REG_ENTROPY = 56;
};

struct K1EGroup {
    uint32_t stage[K1E_STAGE_WORDS]; /* the chunk stream, numeric (MSB-first) words; all zero between tiles */
    uint32_t priv[128 * K1E_PSTRIDE];
    uint32_t off[128];               /* entry lengths in stream order */
    uint32_t flag[4];                /* [0..1] per tile parity: an entry overflowed its private region */
};

struct K1ESmem {
    K1Out coef[2][3];                                /* 24 KB, 1 KB aligned (TMA swizzle) */
    K1F samples[2][3];                               /* 48 KB */
    uint8_t raw[4][K1C_STAGES][3][2][K1_ROW_BYTES];  /* 18 KB */
    float mul[2][64];
    HuffPack huff;
    K1EGroup grp[2];
    unsigned long long samples_full[2], samples_free[2], coef_full[2], coef_free[2];
};

struct K1EParams {
    uint32_t *tmp;         /* [n_chunks][slot_words] chunk streams, big-endian words */
    long long slot_words;
    uint32_t *cta_bits;    /* [n_chunks] */
    uint32_t *cta_tail;    /* [n_chunks] last 32 bits of every chunk stream */
    uint32_t *super_bits;  /* [n_chunks / 64 + 1], zero at launch */
    unsigned long long *dc_pub; /* [n_chunks][3] {epoch << 32 | (uint16) last DC} */
    unsigned int epoch;    /* distinguishes this launch's words from older ones */
    const HuffPack *huff;
    int store_coefs;       /* 1: warp 7 also writes the coefficients to HBM (K1Params.zz_*) */
    int dc_pred[3];        /* predictor of the image's first blocks (0) */
};

/* named barrier of entropy group 0 / 1 (ids 2 / 3; immediate ids keep the kernel's barrier count at 4) */
__device__ __forceinline__ void k1e_bar(int grp, int count)
{
    if (grp == 0)
        asm volatile("bar.sync 2, %0;" ::"r"(count) : "memory");
    else
        asm volatile("bar.sync 3, %0;" ::"r"(count) : "memory");
}

__device__ __forceinline__ void k1e_publish_dc(unsigned long long *w, unsigned epoch, uint32_t word0)
{
    const unsigned long long v = ((unsigned long long)epoch << 32) | (unsigned long long)(word0 & 0xffffu);
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(w), "l"(v) : "memory");
}
__device__ __forceinline__ int k1e_wait_dc(const unsigned long long *w, unsigned epoch)
{
    unsigned long long v;
    for (;;) {
        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(w) : "memory");
        if ((unsigned)(v >> 32) == epoch) break;
        __nanosleep(100);
    }
    return (int)(short)(v & 0xffffull);
}

/* coefficient k of a swizzled row: 16-byte chunk (k >> 3) sits at chunk ^ (slot & 7) */
__device__ __forceinline__ int k1e_coef(unsigned rowbase, unsigned sw4, int k)
{
    short v;
    const unsigned a = rowbase | (((unsigned)k << 1) ^ sw4);
    asm volatile("ld.shared.s16 %0, [%1];" : "=h"(v) : "r"(a));
    return (int)v;
}

/* codes of one entry: the DC difference (do_dc), the coefficients whose bits are set in mm0 | mm1<<32, the
 * end-of-block code (eob, unless position 63 is coded).  Tables are HuffPack words. */
template <class Sink>
__device__ __forceinline__ void k1e_code(unsigned rowbase, unsigned sw4, bool do_dc, int diff, uint32_t mm0, uint32_t mm1, int prev,
                                         bool eob, const uint32_t *sAc, const uint32_t *sDc, Sink &s)
{
    if (do_dc) {
        const int cat = diff == 0 ? 0 : msb_index((uint32_t)abs(diff)) + 1;
        const uint32_t e = sDc[cat];
        const uint32_t extra = (uint32_t)(diff - (diff < 0 ? 1 : 0)) & ((1u << cat) - 1u);
        s.put((e >> 5) | extra, (int)(e & 31u));
    }
#pragma unroll
    for (int half = 0; half < 2; half++) {
        uint32_t mr = __brev(half ? mm1 : mm0); /* bit 31 = position 0 */
        while (mr) {
            const int pos = __clz((int)mr);
            mr ^= 0x80000000u >> pos;
            const int k = half * 32 + pos;
            const int v = k1e_coef(rowbase, sw4, k);
            int run = k - prev - 1;
            prev = k;
            if (run > 15) { /* rare: one ZRL code per 16 zeros */
                const uint32_t zrl = sAc[0xF0];
                do {
                    s.put(zrl >> 5, (int)(zrl & 31u));
                    run -= 16;
                } while (run > 15);
            }
            const int cat = msb_index((uint32_t)abs(v)) + 1;
            const uint32_t e = sAc[(run << 4) | cat];
            const uint32_t extra = (uint32_t)(v - (v < 0 ? 1 : 0)) & ((1u << cat) - 1u);
            s.put((e >> 5) | extra, (int)(e & 31u));
        }
    }
    if (eob && prev != 63) {
        const uint32_t eobc = sAc[0x00];
        s.put(eobc >> 5, (int)(eobc & 31u));
    }
}

template <int RATIO, bool STD = false>
__global__ void __launch_bounds__(K1ECfg<RATIO>::THREADS, 2) k1e_fused(const K1Params p, long long n_tiles, const __grid_constant__ K1PStore st,
                                                            const K1EParams e)
{
    typedef K1ECfg<RATIO> C;
    extern __shared__ unsigned char k1_smem_raw[];
    unsigned char *k1_smem = k1_smem_raw + ((1024u - ((unsigned)__cvta_generic_to_shared(k1_smem_raw) & 1023u)) & 1023u);
    K1ESmem &S = *reinterpret_cast<K1ESmem *>(k1_smem);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    constexpr int NT = C::NT;

    if (tid < 128) S.mul[tid >> 6][tid & 63] = p.mul[tid >> 6][tid & 63];
    for (int i = tid; i < (int)(sizeof(S.grp) / 4); i += C::THREADS) reinterpret_cast<uint32_t *>(S.grp)[i] = 0u;
    for (int i = tid; i < (int)(sizeof(HuffPack) / 4); i += C::THREADS)
        reinterpret_cast<uint32_t *>(&S.huff)[i] = reinterpret_cast<const uint32_t *>(e.huff)[i];
    if (tid == 0) {
#pragma unroll
        for (int b = 0; b < 2; b++) {
            mbar_init(&S.samples_full[b], 4);
            mbar_init(&S.samples_free[b], NT);
            mbar_init(&S.coef_full[b], NT);
            mbar_init(&S.coef_free[b], 1u + (e.store_coefs ? 1u : 0u));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (w < 4) {
        /* ================= convert warp k (as k1_transform_pc) ================= */
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(C::REG_CONVERT));
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
            mbar_wait_relaxed(&S.samples_free[b], (use & 1u) ^ 1u);
            k1_convert_full<RATIO, STD>(S.samples[b], k, lane, pa, pb);
            if (RATIO == 1) k1_convert_chroma422<STD>(S.samples[b], k, lane, ca, cb);
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.samples_full[b]);
        }
    } else if (w < 8) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(C::REG_TRANSFORM));
        if (w < 4 + NT) {
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
                {
                    /* the tile's last DC of this warp's component(s), for the next tile's first difference */
                    const K1Tile tl = k1_tile(p, t);
                    const int nvalid = min(K1_TILE_BLOCKS, p.BW - tl.tb0);
                    unsigned long long *pub = e.dc_pub + t * 3;
                    if (RATIO == 0 || c == 0) {
                        if (lane == nvalid - 1) k1e_publish_dc(pub + c, e.epoch, o[0]);
                    } else {
                        const int nvc = nvalid >> 1;
                        if (lane == nvc - 1) k1e_publish_dc(pub + 1, e.epoch, o[0]);
                        if (lane == 16 + nvc - 1) k1e_publish_dc(pub + 2, e.epoch, o[0]);
                    }
                }
                mbar_wait(&S.coef_free[b], (use & 1u) ^ 1u);
#pragma unroll
                for (int j = 0; j < 8; j++)
                    S.coef[b][c][lane][j ^ (lane & 7)] = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&S.coef_full[b]);
            }
        } else if (w == 7 && e.store_coefs) {
            /* ================= store warp (parity tap) ================= */
            int it = 0;
            for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x, it++) {
                const int b = it & 1;
                const unsigned use = (unsigned)(it >> 1);
                const K1Tile tl = k1_tile(p, t);
                mbar_wait_relaxed(&S.coef_full[b], use & 1u);
                k1p_store<RATIO>(S.coef[b], p, st, tl, lane);
                __syncwarp();
                if (lane == 0) mbar_arrive(&S.coef_free[b]);
            }
            if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(C::REG_ENTROPY));
        const int grp = (w - 8) >> 2, we = (w - 8) & 3;
        if (we >= C::NEW) return;
        /* ================= entropy group `grp`: tiles grp, grp + NG, ... of this CTA ================= */
        constexpr int NE = C::NE, NG = C::NG;
        const int te = tid - 256 - 128 * grp;
        K1EGroup &G = S.grp[grp];
        /* entry of this thread (component-major, so that warps are component-pure) */
        int crow, row, part, e_local;
        bool first_of_comp;
        const bool luma = te < 64;
        if (luma) {
            crow = 0;
            row = te >> 1;
            part = te & 1;
            e_local = (RATIO == 0) ? 4 * row + part : 6 * (row >> 1) + 2 * (row & 1) + part;
            first_of_comp = (te == 0);
        } else if (RATIO == 0) {
            crow = 1 + ((te - 64) >> 5);
            row = (te - 64) & 31;
            part = 0;
            e_local = 4 * row + 1 + crow;
            first_of_comp = (row == 0);
        } else {
            crow = 1;
            row = te - 64; /* slots 0..15 Cb, 16..31 Cr */
            part = 0;
            e_local = 6 * (row & 15) + 4 + (row >> 4);
            first_of_comp = ((row & 15) == 0);
        }
        const int comp = (RATIO == 0) ? crow : (luma ? 0 : 1 + (row >> 4));
        const unsigned sw4 = (unsigned)(row & 7) << 4;
        const uint32_t *sAc = S.huff.ac[luma ? 0 : 1], *sDc = S.huff.dc[luma ? 0 : 1];
        uint32_t *priv = G.priv + te * K1E_PSTRIDE;

        int it = grp;
        for (long long t = (long long)blockIdx.x + (long long)grp * gridDim.x; t < n_tiles; t += (long long)NG * gridDim.x, it += NG) {
            const int b = it & 1;
            const unsigned use = (unsigned)(it >> 1);
            const K1Tile tl = k1_tile(p, t);
            const int nvalid = min(K1_TILE_BLOCKS, p.BW - tl.tb0);
            const bool valid = luma ? (row < nvalid) : ((RATIO == 0) ? (row < nvalid) : ((row & 15) < (nvalid >> 1)));
            const long long chunk = t;
            /* the predecessor's DC can be fetched while the transform warps are still busy */
            int pred = 0;
            if (first_of_comp && part == 0) pred = (chunk == 0) ? (comp == 0 ? e.dc_pred[0] : (comp == 1 ? e.dc_pred[1] : e.dc_pred[2])) : k1e_wait_dc(e.dc_pub + (chunk - 1) * 3 + comp, e.epoch);
            mbar_wait(&S.coef_full[b], use & 1u);
            const uint4 *rowp = S.coef[b][crow][row];
            const unsigned rowbase = (unsigned)__cvta_generic_to_shared(rowp);

            /* masks of the non-zero AC coefficients; a luma block is coded by two adjacent lanes */
            uint32_t mm0, mm1;
            int prev = 0;
            bool eob = true, active = valid;
            if (luma) {
                const int j0 = part * 4;
                const uint32_t mine_half = (nz_mask16(rowp[(j0 + 0) ^ (row & 7)], rowp[(j0 + 1) ^ (row & 7)]) |
                                            (nz_mask16(rowp[(j0 + 2) ^ (row & 7)], rowp[(j0 + 3) ^ (row & 7)]) << 16)) &
                                           (part == 0 ? ~1u : ~0u);
                const uint32_t other_half = __shfl_xor_sync(0xffffffffu, mine_half, 1);
                const unsigned long long M = part == 0 ? ((unsigned long long)mine_half | ((unsigned long long)other_half << 32))
                                                       : ((unsigned long long)other_half | ((unsigned long long)mine_half << 32));
                const int Sp = ec_split(M);
                const unsigned long long lowmask = Sp >= 63 ? ~0ull : ((2ull << Sp) - 1ull);
                unsigned long long mine;
                if (part == 0) {
                    mine = M & lowmask;
                    eob = (Sp >= 63);
                } else {
                    mine = M & ~lowmask;
                    const unsigned long long before = M & lowmask;
                    prev = before ? 63 - __clzll((long long)before) : 0;
                    if (Sp >= 63) active = false; /* part 0 codes the whole block */
                }
                mm0 = (uint32_t)mine;
                mm1 = (uint32_t)(mine >> 32);
            } else {
                const int s7 = row & 7;
                mm0 = (nz_mask16(rowp[0 ^ s7], rowp[1 ^ s7]) | (nz_mask16(rowp[2 ^ s7], rowp[3 ^ s7]) << 16)) & ~1u;
                mm1 = nz_mask16(rowp[4 ^ s7], rowp[5 ^ s7]) | (nz_mask16(rowp[6 ^ s7], rowp[7 ^ s7]) << 16);
            }
            int diff = 0;
            if (part == 0) {
                if (!first_of_comp) pred = (int)(short)(S.coef[b][crow][row - 1][(row - 1) & 7].x & 0xffffu);
                diff = (int)(short)(rowp[row & 7].x & 0xffffu) - pred;
            }
            uint32_t nbits = 0;
            if (active) {
                PrivSink s;
                s.acc = 0;
                s.nacc = 0;
                s.wp = priv;
                s.end = priv + K1E_PRIV;
                s.nwords = 0;
                k1e_code(rowbase, sw4, part == 0, diff, mm0, mm1, prev, eob, sAc, sDc, s);
                nbits = s.finish();
                if (nbits > (uint32_t)K1E_PRIV * 32u) G.flag[b] = 1u;
            }
            G.off[e_local] = nbits;
            k1e_bar(grp, NE); /* B1: every entry is coded */

            /* exclusive scan in stream order, redundantly per warp: lane = MCU */
            uint32_t loff, total;
            if (RATIO == 0) {
                const uint4 l4 = reinterpret_cast<const uint4 *>(G.off)[lane];
                const uint32_t mt = l4.x + l4.y + l4.z + l4.w;
                uint32_t inc = mt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t u = __shfl_up_sync(0xffffffffu, inc, d);
                    if (lane >= d) inc += u;
                }
                total = __shfl_sync(0xffffffffu, inc, 31);
                const uint32_t ex0 = inc - mt, ex1 = ex0 + l4.x, ex2 = ex1 + l4.y, ex3 = ex2 + l4.z;
                const int src = 16 * (we & 1) + (lane >> 1);
                const uint32_t a0 = __shfl_sync(0xffffffffu, ex0, src), a1 = __shfl_sync(0xffffffffu, ex1, src);
                loff = we < 2 ? (part ? a1 : a0) : (we == 2 ? ex2 : ex3);
            } else {
                uint32_t v[6];
                {
                    const uint2 *o2 = reinterpret_cast<const uint2 *>(G.off) + 3 * (lane & 15);
                    const uint2 x0 = o2[0], x1 = o2[1], x2 = o2[2];
                    const bool on = lane < 16;
                    v[0] = on ? x0.x : 0u;
                    v[1] = on ? x0.y : 0u;
                    v[2] = on ? x1.x : 0u;
                    v[3] = on ? x1.y : 0u;
                    v[4] = on ? x2.x : 0u;
                    v[5] = on ? x2.y : 0u;
                }
                const uint32_t mt = v[0] + v[1] + v[2] + v[3] + v[4] + v[5];
                uint32_t inc = mt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t u = __shfl_up_sync(0xffffffffu, inc, d);
                    if (lane >= d) inc += u;
                }
                total = __shfl_sync(0xffffffffu, inc, 31);
                uint32_t ex[6];
                ex[0] = inc - mt;
#pragma unroll
                for (int j = 1; j < 6; j++) ex[j] = ex[j - 1] + v[j - 1];
                if (we < 2) {
                    const int src = 8 * we + (lane >> 2);
                    const uint32_t a0 = __shfl_sync(0xffffffffu, ex[0], src), a1 = __shfl_sync(0xffffffffu, ex[1], src);
                    const uint32_t a2 = __shfl_sync(0xffffffffu, ex[2], src), a3 = __shfl_sync(0xffffffffu, ex[3], src);
                    const int j = lane & 3;
                    loff = j == 0 ? a0 : (j == 1 ? a1 : (j == 2 ? a2 : a3));
                } else {
                    const uint32_t a5 = __shfl_sync(0xffffffffu, ex[5], lane & 15);
                    loff = lane < 16 ? ex[4] : a5;
                }
            }
            const uint32_t nwords = (total + 31u) >> 5;
            const bool staged = (G.flag[b] == 0u) && nwords < (uint32_t)K1E_STAGE_WORDS; /* word [nwords] may receive an empty OR */
            uint32_t *slot = e.tmp + chunk * e.slot_words;
            if (staged) {
                if (te == 0) mbar_arrive(&S.coef_free[b]); /* the coefficient rows are dead */
                if (active) {
                    const uint32_t sh = loff & 31u, nw = (nbits + 31u) >> 5;
                    const unsigned a0 = (unsigned)__cvta_generic_to_shared(G.stage + (loff >> 5));
                    for (uint32_t i = 0; i < nw; i++) {
                        const uint32_t v = priv[i];
                        asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a0 + 4u * i), "r"(v >> sh) : "memory");
                        if (sh) asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a0 + 4u * i + 4u), "r"(v << (32u - sh)) : "memory");
                    }
                }
                k1e_bar(grp, NE); /* B2: the chunk stream is complete */
                /* copy-out; the words holding the stream's last 32 bits belong to thread 0, which needs them whole */
                const uint32_t wt = total > 32u ? (total - 32u) >> 5 : 0u;
                for (uint32_t i = (uint32_t)te; i < wt; i += NE) {
                    slot[i] = __byte_perm(G.stage[i], 0, 0x0123);
                    G.stage[i] = 0u;
                }
                if (te == 0) {
                    e.cta_tail[chunk] = stream_tail32(G.stage, total, false);
                    for (uint32_t i = wt; i < nwords; i++) {
                        slot[i] = __byte_perm(G.stage[i], 0, 0x0123);
                        G.stage[i] = 0u;
                    }
                    e.cta_bits[chunk] = total;
                    atomicAdd(&e.super_bits[chunk >> EC_SUPER_SHIFT], total);
                }
            } else {
                /* rare (a block longer than its private region, or a very long chunk): straight into the slot */
                for (uint32_t i = (uint32_t)te; i < nwords; i += NE) slot[i] = 0u;
                k1e_bar(grp, NE);
                if (active) {
                    PackSink<false> s;
                    s.acc = 0;
                    s.nacc = (int)(loff & 31u);
                    s.wp = slot + (loff >> 5);
                    s.first = true;
                    k1e_code(rowbase, sw4, part == 0, diff, mm0, mm1, prev, eob, sAc, sDc, s);
                    s.finish();
                }
                k1e_bar(grp, NE);
                if (te == 0) {
                    mbar_arrive(&S.coef_free[b]);
                    G.flag[b] = 0u;
                    e.cta_tail[chunk] = stream_tail32(slot, total, true);
                    e.cta_bits[chunk] = total;
                    atomicAdd(&e.super_bits[chunk >> EC_SUPER_SHIFT], total);
                }
            }
        }
    }
}

}  // namespace jpgb200
#endif
