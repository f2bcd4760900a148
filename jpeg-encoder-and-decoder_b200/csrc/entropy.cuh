/*
 * Entropy stage: K3 (per-block bit lengths + per-chunk offsets), K4 (device-wide exclusive
 * scan), K5 (parallel bit pack).  The reference has no bit writer at all (SURVEY.md F1,
 * F2): what it pins is the symbol alphabet -- categories = bit length of |v|, EOB 0x00
 * when zeros trail the last non-zero coefficient, ZRL 0xF0 per 16 zeros
 * (src/huffman.c:182-235) -- and the code tables (src/headers/tables.h:37-72).  The
 * emitted stream follows T.81 F.1.2: true DC differences and (run<<4)|size symbols.
 *
 * Scan order is the interleaved MCU order (T.81 A.2): 4:4:4 -> Y Cb Cr per MCU,
 * 4:2:2 -> Y(2m) Y(2m+1) Cb(m) Cr(m).  One thread encodes one 8x8 block; a CTA owns a chunk
 * of EC_MCUS consecutive MCUs, whose coefficients are contiguous per component in HBM and
 * are staged into shared memory with asynchronous coalesced 16-byte copies (rows of 128 B
 * padded to 144 B, so that per-thread 16-byte row reads are bank-conflict free and a
 * coefficient's address is linear in its zig-zag index).
 */
#ifndef JPGB200_ENTROPY_CUH
#define JPGB200_ENTROPY_CUH

#include <cuda_fp16.h>
#include <stdint.h>

namespace jpgb200 {

#define EC_MCUS 64
#define EC_ROW_CHUNKS 9 /* 8 data chunks of 16 B + 1 pad */
#define EC_SUPER_SHIFT 6 /* chunks per super-chunk = 64 */

/* pack tables */
struct HuffDev {
    uint32_t ac[2][256]; /* (code << 8) | code_length ; [lum|chr][(run<<4)|size] */
    uint32_t dc[2][16];  /* (code << 8) | code_length ; [lum|chr][category]     */
};
/* KE's pack tables: ((code << category) << 5) | (code_length + category): the code already
 * shifted to make room for the additional bits, and the total length, in one word */
struct HuffPack {
    /* [lum|chr][run * 8 + value + 4], |value| <= 3 (99 % of the non-zero coefficients of the headline frames): the
     * ac word of (run, category of value) with value's additional bits already merged.  value = 0 never occurs:
     * entry [4] holds the end-of-block code and [12] the ZRL code (KE keeps lut + dc in shared memory and can
     * leave the rarely used ac table in global memory: KE_AC_GLOBAL) */
    uint32_t lut[2][128];
    uint32_t dc[2][16];
    uint32_t ac[2][256];
};
#define HP_LUT_EOB 4
#define HP_LUT_ZRL 12
/* count tables: total bits of one symbol incl. its additional bits; the AC table is
 * extended to runs of up to 62 zeros, folding in the ZRL codes a long run needs */
struct HuffBits {
    uint8_t ac[2][1024]; /* [lum|chr][(run<<4)|size], run 0..62 */
    uint8_t dc[2][16];
};

/* per-encode device state, written by the scan kernels and read by later kernels */
struct EncState {
    unsigned long long total_bits; /* entropy-coded bits before padding and stuffing */
    unsigned long long nbytes;     /* ceil(total_bits / 8)                             */
    unsigned long long n_tiles;    /* stuffing tiles                                   */
    unsigned long long total_ff;   /* number of 0xFF bytes that get a stuffed 0x00     */
    unsigned long long out_len;    /* header + stuffed scan + EOI                      */
    int error;                     /* 1: output capacity too small                     */
    int pad;
    unsigned long long ticket;     /* K6 tile dispenser                                 */
    unsigned long long ticket35;   /* K35 chunk dispenser (reset by K35 itself)         */
};

struct EcParams {
    const int16_t *zz; /* [Y blocks | Cb blocks | Cr blocks][64] */
    long long coff1, coff2; /* first Cb / Cr block inside zz */
    long long n_mcu;
    int dc_pred[3];        /* predictor of the first block per component (0 for a whole image) */
};

template <int RATIO>
struct EcCfg {
    static const int BPM = (RATIO == 0) ? 3 : (RATIO == 1 ? 4 : 6); /* 2 = 4:2:0: 4 luma blocks per MCU, stored in MCU order */
    static const int NT = EC_MCUS * BPM;
    static const int STAGE_WORDS = (RATIO == 0) ? 3072 : 1664; /* chunk stream staging (static smem <= 48 KB) */
};

/* Row (= thread) <-> block mapping inside a chunk.  Rows are grouped by COMPONENT, not by
 * MCU: luma blocks carry ~3x the non-zero coefficients of chroma blocks, and the per-block
 * loops run as long as the busiest lane of a warp, so component-pure warps execute far fewer
 * divergent iterations.  s_local is the block's position in scan (MCU interleaved) order
 * inside the chunk; gabs its block index inside zz. */
template <int RATIO>
struct EcMap {
    int comp, s_local;
    long long gabs;
    bool valid, first_of_comp;
    __device__ __forceinline__ EcMap(const EcParams &p, long long m0, int row)
    {
        if (RATIO == 0) {
            comp = row >> 6;
            const int m = row & 63;
            s_local = m * 3 + comp;
            gabs = m0 + m + (comp == 0 ? 0ll : (comp == 1 ? p.coff1 : p.coff2));
            valid = (m0 + m) < p.n_mcu;
            first_of_comp = (m == 0);
        } else {
            if (row < 128) {
                comp = 0;
                s_local = (row >> 1) * 4 + (row & 1);
                gabs = 2 * m0 + row;
                valid = (m0 + (row >> 1)) < p.n_mcu;
                first_of_comp = (row == 0);
            } else {
                comp = 1 + ((row - 128) >> 6);
                const int m = row & 63;
                s_local = m * 4 + 1 + comp;
                gabs = m0 + m + (comp == 1 ? p.coff1 : p.coff2);
                valid = (m0 + m) < p.n_mcu;
                first_of_comp = (m == 0);
            }
        }
    }
};

/* 16-byte asynchronous global->shared copy (LDGSTS): no register staging, so a thread keeps
 * all of its copies in flight at once */
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}

/* asynchronous fill of the chunk's coefficient rows.  Rows are component-major (EcMap) and a
 * component's blocks of the chunk are contiguous in HBM, so each component is one linear run of
 * 16-byte chunks: chunk i -> row i/8, column i%8 (rows padded to 144 B in shared memory) */
template <int RATIO, int NTH = EcCfg<RATIO>::NT>
__device__ __forceinline__ void ec_fill(uint4 (*sC)[EC_ROW_CHUNKS], const EcParams &p, long long m0, int tid)
{
    constexpr int NT = NTH;
    const long long left = p.n_mcu - m0;
    const int nv = left < EC_MCUS ? (int)left : EC_MCUS; /* valid MCUs of this chunk */
    const uint4 *z = reinterpret_cast<const uint4 *>(p.zz);
    {
        constexpr int YPM = (RATIO == 0) ? 1 : (RATIO == 1 ? 2 : 4); /* luma blocks per MCU */
        const int n0 = YPM * nv * 8;
        const uint4 *src = z + YPM * m0 * 8;
#pragma unroll
        for (int k = 0; k < (YPM * 64 * 8 + NT - 1) / NT; k++) {
            const int i = k * NT + tid;
            if (i < n0) cp_async16(&sC[i >> 3][i & 7], src + i);
        }
    }
    {
        const int n1 = nv * 8, r1 = (RATIO == 0 ? 64 : (RATIO == 1 ? 128 : 256));
        const uint4 *s1 = z + (p.coff1 + m0) * 8, *s2 = z + (p.coff2 + m0) * 8;
#pragma unroll
        for (int k = 0; k < (64 * 8 + NT - 1) / NT; k++) {
            const int i = k * NT + tid;
            if (i < n1) {
                cp_async16(&sC[r1 + (i >> 3)][i & 7], s1 + i);
                cp_async16(&sC[r1 + 64 + (i >> 3)][i & 7], s2 + i);
            }
        }
    }
}


/* ---- KE's coefficient rows: 128 bytes each, NOT padded; the 16-byte chunk j of row r sits at chunk j ^ (r & 7)
 * (the swizzle of K1's coefficient tile): per-lane 16-byte row reads stay bank-conflict free and a CTA needs 3 KB
 * less shared memory than with 144-byte rows -- the sixth resident CTA per SM (KE lives on occupancy). */
#define KE_ROW_CHUNKS 8

__device__ __forceinline__ void cp_async16_sa(unsigned sa, const void *gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem) : "memory");
}

/* Piece i = k * NT + tid of a component run goes to row i / 8, chunk (i & 7) ^ (row & 7).  NT is a multiple of 64,
 * so row & 7 and i & 7 do not depend on k: a thread's shared address is ONE base plus k * NT * 16 bytes, and the
 * global address likewise -- no per-copy address arithmetic. */
template <int RATIO, int NTH>
__device__ __forceinline__ void ke_fill(uint4 (*sC)[KE_ROW_CHUNKS], const EcParams &p, long long m0, int tid)
{
    constexpr int NT = NTH;
    static_assert(NT % 64 == 0, "ke_fill: the swizzle term must not depend on the iteration");
    const long long left = p.n_mcu - m0;
    const int nv = left < EC_MCUS ? (int)left : EC_MCUS; /* valid MCUs of this chunk */
    const uint4 *z = reinterpret_cast<const uint4 *>(p.zz);
    const int r0 = tid >> 3;
    const unsigned sa0 = (unsigned)__cvta_generic_to_shared(sC) + (unsigned)r0 * 128u + (unsigned)(((tid & 7) ^ (r0 & 7)) << 4);
    {
        constexpr int YPM = (RATIO == 0) ? 1 : (RATIO == 1 ? 2 : 4); /* luma blocks per MCU */
        const int n0 = YPM * nv * 8;
        const uint4 *src = z + YPM * m0 * 8 + tid;
#pragma unroll
        for (int k = 0; k < (YPM * 64 * 8 + NT - 1) / NT; k++)
            if (k * NT + tid < n0) cp_async16_sa(sa0 + (unsigned)(k * NT * 16), src + k * NT);
    }
    {
        constexpr int r1 = (RATIO == 0 ? 64 : (RATIO == 1 ? 128 : 256)); /* r1 and r1 + 64 are multiples of 8 */
        const int n1 = nv * 8;
        const uint4 *s1 = z + (p.coff1 + m0) * 8 + tid, *s2 = z + (p.coff2 + m0) * 8 + tid;
#pragma unroll
        for (int k = 0; k < (64 * 8 + NT - 1) / NT; k++)
            if (k * NT + tid < n1) {
                cp_async16_sa(sa0 + (unsigned)(r1 * 128 + k * NT * 16), s1 + k * NT);
                cp_async16_sa(sa0 + (unsigned)((r1 + 64) * 128 + k * NT * 16), s2 + k * NT);
            }
    }
}

/* tables -> shared memory, 16 bytes per copy, NTH threads */
template <int NTH, int BYTES>
__device__ __forceinline__ void ec_load_bytes(void *sT, const void *g, int tid)
{
    static_assert(BYTES % 16 == 0, "table size");
    const unsigned sa = (unsigned)__cvta_generic_to_shared(sT) + (unsigned)tid * 16u;
    const uint4 *src = reinterpret_cast<const uint4 *>(g) + tid;
#pragma unroll
    for (int i = 0; i < (BYTES / 16 + NTH - 1) / NTH; i++)
        if (i * NTH + tid < BYTES / 16) cp_async16_sa(sa + (unsigned)(i * NTH * 16), src + i * NTH);
}
template <int NTH, class T>
__device__ __forceinline__ void ec_load_table(T *sT, const T *g, int tid)
{
    ec_load_bytes<NTH, (int)sizeof(T)>(sT, g, tid);
}

/* DC predictor: the previous block of the same component (raster == scan order per
 * component) -- the row above in shared memory, the caller's predictor for the first block
 * of the image/band, or the last block of the previous chunk read from HBM */
template <int RATIO>
__device__ __forceinline__ int ec_dc_pred(const uint4 (*sC)[EC_ROW_CHUNKS], const EcParams &p, long long m0, int tid,
                                          const EcMap<RATIO> &mp)
{
    if (!mp.first_of_comp) return (int)(short)(sC[tid - 1][0].x & 0xffffu);
    if (m0 == 0) return p.dc_pred[mp.comp];
    return (int)p.zz[(mp.gabs - 1) * 64];
}

/* 0xFFFF per non-zero int16 half.  The pair is compared as two fp16 numbers with an
 * unordered != 0: every non-zero bit pattern except 0x8000 (-32768, which a quantised
 * coefficient never reaches) is a non-zero number, a denormal or a NaN, and "unordered"
 * makes NaN patterns (small negative integers) compare true.  One HSET2 per two coefficients. */
__device__ __forceinline__ uint32_t nz_halves(uint32_t w)
{
    __half2 h;
    *reinterpret_cast<uint32_t *>(&h) = w;
    __half2 zero;
    *reinterpret_cast<uint32_t *>(&zero) = 0u;
    return __hneu2_mask(h, zero);
}

/* bit i = (coefficient i of these 16 != 0) */
__device__ __forceinline__ uint32_t nz_mask16(const uint4 a, const uint4 b)
{
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) x |= nz_halves(w[i]) & ((1u << (2 * i)) | (1u << (17 + 2 * i)));
    return (x | (x >> 16)) & 0xffffu;
}

/* masks of the non-zero AC coefficients of a row: m[0] = positions 1..31, m[1] = 32..63 */
__device__ __forceinline__ void ec_nz_masks(const uint4 *row, uint32_t &m0, uint32_t &m1)
{
    m0 = (nz_mask16(row[0], row[1]) | (nz_mask16(row[2], row[3]) << 16)) & ~1u; /* bit 0 is the DC */
    m1 = nz_mask16(row[4], row[5]) | (nz_mask16(row[6], row[7]) << 16);
}

__device__ __forceinline__ int ec_coef(const uint4 *row, int k) { return (int)reinterpret_cast<const short *>(row)[k]; }
/* coefficient k of a swizzled row (sw4 = (row & 7) << 4): byte offset (2 k) ^ sw4 */
__device__ __forceinline__ int ke_coef(const uint4 *row, unsigned sw4, int k)
{
    return (int)*reinterpret_cast<const short *>(reinterpret_cast<const unsigned char *>(row) + (((unsigned)k << 1) ^ sw4));
}

/* index of the most significant set bit (v != 0): category - 1 */
__device__ __forceinline__ int msb_index(uint32_t v)
{
    int r;
    asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}

/* ---- block coder, counting form (K3) ------------------------------------------------------
 * The symbol loop visits only the set bits of the non-zero mask: its trip count is the
 * block's number of non-zero AC coefficients (in a warp: the maximum over 32 blocks of one
 * component), not 63.  One table lookup per symbol returns code length + additional bits +
 * any ZRL codes of a long run. */
__device__ __forceinline__ uint32_t ec_count_block(const uint4 *row, int diff, const uint8_t *bAc, const uint8_t *bDc)
{
    uint32_t bits = bDc[diff == 0 ? 0 : msb_index((uint32_t)abs(diff)) + 1];
    uint32_t mm[2];
    ec_nz_masks(row, mm[0], mm[1]);
    int prev = 0;
#pragma unroll
    for (int half = 0; half < 2; half++) {
        uint32_t m = mm[half];
        while (m) {
            const int k = half * 32 + (__ffs((int)m) - 1);
            m &= m - 1u;
            const int v = ec_coef(row, k);
            const int idx = ((k - prev - 1) << 4) + msb_index((uint32_t)abs(v)) + 1;
            prev = k;
            bits += bAc[idx];
        }
    }
    if (prev != 63) bits += bAc[0]; /* EOB */
    return bits;
}

/* ---- bit writer: big-endian 32-bit words; the first and last word a block touches can be
 * shared with its neighbours and are OR-ed, interior words are owned exclusively */
template <bool SHARED>
struct PackSink {
    unsigned long long acc;
    int nacc;
    uint32_t *wp;
    bool first;
    __device__ __forceinline__ void or_word(uint32_t be)
    {
        if (SHARED) {
            const unsigned sa = (unsigned)__cvta_generic_to_shared(wp);
            asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(sa), "r"(be) : "memory");
        } else {
            atomicOr(wp, be);
        }
    }
    __device__ __forceinline__ void emit(uint32_t word)
    {
        const uint32_t be = __byte_perm(word, 0, 0x0123);
        if (first) {
            or_word(be);
            first = false;
        } else {
            *wp = be;
        }
        wp++;
    }
    __device__ __forceinline__ void put(uint32_t v, int n)
    {
        acc = (acc << n) | (unsigned long long)v;
        nacc += n;
        if (nacc >= 32) {
            nacc -= 32;
            emit((uint32_t)(acc >> nacc));
        }
    }
    __device__ __forceinline__ void finish()
    {
        if (nacc > 0) or_word(__byte_perm((uint32_t)(acc << (32 - nacc)), 0, 0x0123));
    }
};

/* ---- bit writer into KE's staged chunk stream (shared memory, NUMERIC words): the first and last word a block
 * touches are shared with its neighbours and OR-ed, interior words are owned and stored */
struct StageSink {
    unsigned long long acc;
    int nacc;
    uint32_t *wp;
    bool first;
    __device__ __forceinline__ void or_word(uint32_t v)
    {
        const unsigned sa = (unsigned)__cvta_generic_to_shared(wp);
        asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(sa), "r"(v) : "memory");
    }
    __device__ __forceinline__ void put(uint32_t v, int n)
    {
        acc = (acc << n) | (unsigned long long)v;
        nacc += n;
        if (nacc >= 32) {
            nacc -= 32;
            const uint32_t word = (uint32_t)(acc >> nacc);
            if (first) {
                or_word(word);
                first = false;
            } else {
                *wp = word;
            }
            wp++;
        }
    }
    __device__ __forceinline__ void finish()
    {
        if (nacc > 0) or_word((uint32_t)(acc << (32 - nacc)));
    }
};

/* ---- block coder, packing form (K5) ------------------------------------------------------- */
template <class Sink>
__device__ __forceinline__ void ec_pack_block(const uint4 *row, int diff, const uint32_t *sAc, const uint32_t *sDc,
                                              Sink &s)
{
    {
        const int cat = diff == 0 ? 0 : msb_index((uint32_t)abs(diff)) + 1;
        const uint32_t e = sDc[cat];
        const uint32_t extra = (uint32_t)(diff - (diff < 0 ? 1 : 0)) & ((1u << cat) - 1u);
        s.put(((e >> 8) << cat) | extra, (int)(e & 0xffu) + cat);
    }
    uint32_t mm[2];
    ec_nz_masks(row, mm[0], mm[1]);
    int prev = 0;
#pragma unroll
    for (int half = 0; half < 2; half++) {
        uint32_t m = mm[half];
        while (m) {
            const int k = half * 32 + (__ffs((int)m) - 1);
            m &= m - 1u;
            const int v = ec_coef(row, k);
            int run = k - prev - 1;
            prev = k;
            if (run > 15) { /* rare: one ZRL code per 16 zeros */
                const uint32_t zrl = sAc[0xF0];
                do {
                    s.put(zrl >> 8, (int)(zrl & 0xffu));
                    run -= 16;
                } while (run > 15);
            }
            const int cat = msb_index((uint32_t)abs(v)) + 1;
            const uint32_t e = sAc[(run << 4) | cat];
            const uint32_t extra = (uint32_t)(v - (v < 0 ? 1 : 0)) & ((1u << cat) - 1u);
            s.put(((e >> 8) << cat) | extra, (int)(e & 0xffu) + cat);
        }
    }
    if (prev != 63) {
        const uint32_t eob = sAc[0x00];
        s.put(eob >> 8, (int)(eob & 0xffu));
    }
}

/* ---- block coder in two parts ---------------------------------------------------------------
 * A luma block carries about three times the non-zero coefficients of a chroma block, and a
 * thread's coding loop runs as long as its block is busy; so KE codes a luma block with TWO
 * threads.  Both see the same non-zero mask M (bit k = coefficient k != 0, k = 1..63) and derive
 * the same split position S from it: part 0 codes the DC difference and the coefficients 1..S,
 * part 1 the coefficients S+1..63 and the end-of-block code -- its first run length starts at
 * the last non-zero of part 0.  The two bit strings simply follow each other in the stream.
 * S is the first of a few fixed positions that gives part 0 at least half of the non-zeros. */
__device__ __forceinline__ int ec_split(unsigned long long M)
{
    const int n = __popcll(M);
    if (n < 4) return 63; /* not worth splitting: part 0 codes everything */
    const int half = (n + 1) >> 1;
    if (__popcll(M & 0x3eull) >= half) return 5;
    if (__popcll(M & 0x3feull) >= half) return 9;
    if (__popcll(M & 0x7ffeull) >= half) return 14;
    if (__popcll(M & 0xffffffeull) >= half) return 27;
    return 40;
}

/* The masks of one entry.  WHOLE = true: one thread codes the entire block (chroma); otherwise part 0 or 1 of a
 * split block, coded by two ADJACENT lanes which each build one half of the non-zero mask and swap them: that
 * shuffle is executed by all 32 lanes of the (component-pure) warp with a full mask, valid block or not, so it
 * must be called OUTSIDE any `valid` branch.  Returns false when this entry has nothing to code. */
template <bool WHOLE>
__device__ __forceinline__ bool ec_entry_masks(const uint4 *row, int sw, int part, uint32_t (&mm)[2], int &prev, bool &eob)
{
    prev = 0;
    eob = true;
    if (WHOLE) {
        mm[0] = (nz_mask16(row[0 ^ sw], row[1 ^ sw]) | (nz_mask16(row[2 ^ sw], row[3 ^ sw]) << 16)) & ~1u; /* bit 0 is the DC */
        mm[1] = nz_mask16(row[4 ^ sw], row[5 ^ sw]) | (nz_mask16(row[6 ^ sw], row[7 ^ sw]) << 16);
        return true;
    }
    const uint32_t mine_half = part == 0 ? ((nz_mask16(row[0 ^ sw], row[1 ^ sw]) | (nz_mask16(row[2 ^ sw], row[3 ^ sw]) << 16)) & ~1u)
                                         : (nz_mask16(row[4 ^ sw], row[5 ^ sw]) | (nz_mask16(row[6 ^ sw], row[7 ^ sw]) << 16));
    const uint32_t other_half = __shfl_xor_sync(0xffffffffu, mine_half, 1);
    const unsigned long long M = part == 0 ? ((unsigned long long)mine_half | ((unsigned long long)other_half << 32))
                                           : ((unsigned long long)other_half | ((unsigned long long)mine_half << 32));
    const int S = ec_split(M);
    const unsigned long long lowmask = S >= 63 ? ~0ull : ((2ull << S) - 1ull);
    unsigned long long mine;
    bool active = true;
    if (part == 0) {
        mine = M & lowmask;
        eob = (S >= 63);
    } else {
        mine = M & ~lowmask;
        const unsigned long long before = M & lowmask;
        prev = before ? 63 - __clzll((long long)before) : 0;
        if (S >= 63) active = false; /* part 0 codes the whole block */
    }
    mm[0] = (uint32_t)mine;
    mm[1] = (uint32_t)(mine >> 32);
    return active;
}

/* codes of one entry from its masks: the DC difference (do_dc), the coefficients whose bits are set in mm, the
 * end-of-block code (eob, unless position 63 is coded).  Tables are HuffPack words; lut = HuffPack::lut[table]. */
template <class Sink>
__device__ __forceinline__ void ec_pack_masked(const uint4 *row, unsigned sw4, bool do_dc, int diff, const uint32_t (&mm)[2], int prev,
                                               bool eob, const uint32_t *sAc, const uint32_t *sDc, const uint32_t *lut, Sink &s)
{
    if (do_dc) {
        const int cat = diff == 0 ? 0 : msb_index((uint32_t)abs(diff)) + 1;
        const uint32_t e = sDc[cat];
        const uint32_t extra = (uint32_t)(diff - (diff < 0 ? 1 : 0)) & ((1u << cat) - 1u);
        s.put((e >> 5) | extra, (int)(e & 31u));
    }
#pragma unroll
    for (int half = 0; half < 2; half++) {
        uint32_t mr = __brev(mm[half]); /* bit 31 = position 0: one FLO per coefficient, no BREV */
        while (mr) {
            const int pos = __clz((int)mr);
            mr ^= 0x80000000u >> pos;
            const int k = half * 32 + pos;
            const int v = ke_coef(row, sw4, k);
            int run = k - prev - 1;
            prev = k;
            const unsigned u = (unsigned)(v + 4);
            uint32_t e; /* (code << category | additional bits) << 5 | total length */
            if (run < 16 && u - 1u < 7u) { /* |v| <= 3: one table word holds code, additional bits and length */
                e = lut[((unsigned)run << 3) + u];
            } else {
                if (run > 15) { /* rare: one ZRL code per 16 zeros */
                    const uint32_t zrl = lut[HP_LUT_ZRL];
#pragma unroll 1
                    do {
                        s.put(zrl >> 5, (int)(zrl & 31u));
                        run -= 16;
                    } while (run > 15);
                }
                const int cat = msb_index((uint32_t)abs(v)) + 1;
                const uint32_t extra = (uint32_t)(v - (v < 0 ? 1 : 0)) & ((1u << cat) - 1u);
                e = sAc[(run << 4) | cat] | (extra << 5);
            }
            /* ONE put for both kinds of symbol.  (Fetching the NEXT coefficient's position and value before this one is
             * coded -- software pipelining of find-bit -> address -> load against table word -> shift -> accumulate --
             * was measured: 73.1 us against 72.2 us.  So was a form of this loop with the position arithmetic folded
             * into one shift and one XOR on 32-bit shared addresses -- three instructions fewer per symbol, LDS.S16
             * instead of LDS.U16 + PRMT: 57.9-58.1 us against 57.9-58.0 us, no difference; what did count was the
             * bit writer's own chain, see PrivSinkBF.) */
            s.put(e >> 5, (int)(e & 31u));
        }
    }
    if (eob && prev != 63) {
        const uint32_t eobc = lut[HP_LUT_EOB];
        s.put(eobc >> 5, (int)(eobc & 31u));
    }
}

/* exclusive scan of one uint32 per thread over the CTA; returns this thread's offset and
 * the CTA total through `total` */
template <int NT>
__device__ __forceinline__ uint32_t cta_exclusive_scan(uint32_t v, uint32_t *sWarp, int tid, uint32_t &total)
{
    const int lane = tid & 31, wid = tid >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) sWarp[wid] = inc;
    __syncthreads();
    constexpr int NW = NT / 32;
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < NW; i++) {
        const uint32_t s = sWarp[i];
        if (i < wid) base += s;
        tot += s;
    }
    total = tot;
    return base + inc - v;
}

/* chained-scan state word: flag (2 bits) | epoch (22 bits) | value (40 bits) */
#define CS_FLAG_A 1ull
#define CS_FLAG_P 2ull
#define CS_VALUE_BITS 40
#define CS_EPOCH_BITS 22
#define CS_VALUE_MASK ((1ull << CS_VALUE_BITS) - 1ull)
#define CS_EPOCH_MASK ((1ull << CS_EPOCH_BITS) - 1ull)

__device__ __forceinline__ unsigned long long cs_pack(unsigned long long flag, unsigned long long epoch,
                                                      unsigned long long value)
{
    return (flag << 62) | ((epoch & CS_EPOCH_MASK) << CS_VALUE_BITS) | (value & CS_VALUE_MASK);
}

/* Executed by all 32 lanes of one warp.  Returns the sum of the values of tiles [0, t).
 * publish_a: announce this tile's own value first so that successors need not wait for its
 * prefix; the inclusive prefix is published at the end when publish_p is set.
 * Every lane inspects W consecutive predecessors per step (window of 32*W tiles, all loads in
 * flight together), so the "prefix known" front advances 32*W tiles per memory round trip even
 * when a whole wave of tiles starts at once. */
template <int W>
__device__ __forceinline__ unsigned long long cs_lookback(volatile unsigned long long *ts, unsigned long long t,
                                                          unsigned long long epoch, unsigned long long own, int lane,
                                                          bool publish_a, bool publish_p)
{
    if (t == 0) {
        if (lane == 0 && publish_p) ts[0] = cs_pack(CS_FLAG_P, epoch, own);
        return 0ull;
    }
    if (lane == 0 && publish_a) ts[t] = cs_pack(CS_FLAG_A, epoch, own);
    unsigned long long excl = 0;
    long long j = (long long)t - 1;
    for (;;) {
        bool done = false;
        for (;;) {
            /* lane inspects tiles j - lane*W - k, k = 0..W-1 (nearest first) */
            unsigned long long st[W];
#pragma unroll
            for (int k = 0; k < W; k++) {
                const long long idx = j - (long long)lane * W - k;
                st[k] = idx >= 0 ? ts[idx] : cs_pack(CS_FLAG_P, epoch, 0ull); /* before tile 0: prefix 0 */
            }
            bool has_p = false, empty_before_p = false;
            unsigned long long sum = 0ull;
#pragma unroll
            for (int k = 0; k < W; k++) {
                const bool cur = ((st[k] >> CS_VALUE_BITS) & CS_EPOCH_MASK) == (epoch & CS_EPOCH_MASK);
                const unsigned flag = cur ? (unsigned)(st[k] >> 62) : 0u;
                if (!has_p) {
                    if (flag == 0u) empty_before_p = true;
                    sum += st[k] & CS_VALUE_MASK;
                    if (flag == 2u) has_p = true;
                }
            }
            const unsigned pm = __ballot_sync(0xffffffffu, has_p);
            const unsigned em = __ballot_sync(0xffffffffu, empty_before_p);
            const unsigned upto = pm ? (0xffffffffu >> (32 - __ffs((int)pm))) : 0xffffffffu; /* lanes 0..first P */
            if ((em & upto) == 0u) {
                unsigned long long c = ((upto >> lane) & 1u) ? sum : 0ull;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
                excl += c;
                done = pm != 0u;
                break;
            }
        }
        if (done) break;
        j -= 32 * W;
    }
    if (lane == 0 && publish_p) ts[t] = cs_pack(CS_FLAG_P, epoch, excl + own);
    return excl;
}

/* the last 32 bits of a stream of `total` bits (right-aligned; fewer if the stream is shorter) */
__device__ __forceinline__ uint32_t stream_tail32(const uint32_t *words, uint32_t total, bool big_endian)
{
    if (total == 0u) return 0u;
    const uint32_t L = total > 32u ? total - 32u : 0u, w = L >> 5, sh = L & 31u, nw = (total + 31u) >> 5;
    uint32_t hi = words[w], lo = (w + 1u < nw) ? words[w + 1u] : 0u;
    if (big_endian) {
        hi = __byte_perm(hi, 0, 0x0123);
        lo = __byte_perm(lo, 0, 0x0123);
    }
    const uint32_t v = sh ? ((hi << sh) | (lo >> (32u - sh))) : hi;
    return total >= 32u ? v : (v >> (32u - total));
}

/* ---- bit writer into a thread-private region: numeric (MSB-first) 32-bit words; words past
 * the region are dropped but still counted, so the bit length is always right */
struct PrivSink {
    unsigned long long acc;
    int nacc;
    uint32_t *wp, *end;
    uint32_t nwords;
    __device__ __forceinline__ void put(uint32_t v, int n)
    {
        acc = (acc << n) | (unsigned long long)v;
        nacc += n;
        if (nacc >= 32) {
            nacc -= 32;
            if (wp < end) *wp = (uint32_t)(acc >> nacc);
            wp++;
            nwords++;
        }
    }
    /* returns the number of bits written */
    __device__ __forceinline__ uint32_t finish()
    {
        const uint32_t bits = nwords * 32u + (uint32_t)nacc;
        if (nacc > 0 && wp < end) *wp = (uint32_t)(acc << (32 - nacc));
        return bits;
    }
};

/* ---- KE's private bit writer, branch-free: in a warp some lane completes a word at practically every symbol, so a
 * branch around the flush is taken by the warp every time and only adds its own instructions (and a convergence
 * barrier pair).  Here the word is stored UNCONDITIONALLY -- to the region while it has room and a word was
 * completed, else to the pad word behind the region -- and the bookkeeping is selects. */
struct PrivSinkBF {
    unsigned long long acc;
    unsigned nacc;   /* bits of acc not stored yet: < 32 */
    unsigned wa, ea; /* shared-window addresses: the next word of the region, the region's pad word (always writable) */
    /* 32-bit shared addresses and arithmetic selects only: full = (nacc + n) >> 5 is 0 or 1, the destination is
     * min(wa | (full - 1), ea) -- the pad word when no word was completed or the region is used up.  (58.0 us per 8K
     * frame against 60.7 us with generic pointers and compare-and-select; a PREDICATED store of completed words
     * only -- two instructions fewer -- was measured at 59.2 us.) */
    __device__ __forceinline__ void put(uint32_t v, int n)
    {
        acc = (acc << n) | (unsigned long long)v;
        const unsigned t = nacc + (unsigned)n; /* < 64 */
        const unsigned full = t >> 5;
        nacc = t & 31u;
        const unsigned dst = min(wa | (full - 1u), ea);
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(dst), "r"((uint32_t)(acc >> nacc)) : "memory"); /* meaningful only when `full` */
        wa += full << 2;
    }
    /* returns the number of bits written; `base` = shared-window address of the region's first word */
    __device__ __forceinline__ uint32_t finish(unsigned base)
    {
        const uint32_t bits = ((wa - base) << 3) + nacc;
        if (nacc > 0 && wa < ea) asm volatile("st.shared.b32 [%0], %1;" ::"r"(wa), "r"((uint32_t)(acc << (32u - nacc))) : "memory");
        return bits;
    }
};

/* ---- KE (k_encode_chunks): every block is coded ONCE -----------------------------------------
 * One CTA per chunk of EC_MCUS MCUs, coefficients read from HBM once, one thread per block:
 *   code     : the block's codes go, at bit 0, into a thread-private region of shared memory
 *              (EC_PRIV words); the thread learns its bit length as a by-product -- there is no
 *              separate counting pass
 *   scan     : CTA scan of the lengths in scan (MCU-interleaved) order -> the block's offset
 *              inside the CHUNK's own stream (bit 0 = the chunk's first bit), chunk total T
 *   relocate : the private words are funnel-shifted to that offset and OR-ed into the chunk
 *              stream, assembled in shared memory over the (now dead) coefficient rows, then
 *              copied to the chunk's slot of a temporary buffer with coalesced stores -- a chunk
 *              owns its slot, so nothing in HBM is shared between CTAs
 *   totals   : T per chunk and, atomically, per 64 chunks: the two-level sums from which KC
 *              derives every chunk's global bit offset (the device-wide exclusive prefix sum)
 * A block longer than its private region (> 40 bytes: busy luma blocks of noisy images) is coded a second
 * time, straight into the staged stream at its offset (StageSink) -- only that block, not its chunk.  A chunk
 * whose stream is longer than the staging area (> 24 KB for 192 blocks) clears its slot and packs straight
 * into it at the offsets just computed (atomics on block-boundary words). */
struct KeParams {
    const int16_t *zz;
    long long coff1, coff2, n_mcu;
    int dc_pred[3];
    uint32_t *tmp;         /* [n_chunks][slot_words] chunk-local streams, big-endian words */
    long long slot_words;
    uint32_t *cta_bits;    /* [n_chunks] */
    uint32_t *cta_tail;    /* [n_chunks] the chunk stream's last 32 bits: the next chunk's leading bits (KF, KS) */
    uint32_t *super_bits;  /* [n_chunks / 64 + 1], cleared by K1 */
    uint32_t *local_off;   /* optional parity tap: [n_mcu*BPM] offset of every block in its chunk */
    /* batches of small frames in one launch: blockIdx.y = frame, every array advances by these strides */
    long long zz_fstride, tmp_fstride; /* int16 elements, words */
    int cta_fstride, super_fstride;    /* elements */
    const int *dc_pred_dev;            /* optional: the predictors in device memory (multi-GPU bands) */
    unsigned int *recode_count;        /* optional: += the number of blocks the tight form codes twice (the host picks KE's form from it) */
};

/* KE's threads: "entries" in component-major order -- two per luma block (part 0, part 1), one per
 * chroma block -- so that warps are component-pure; e_local is the entry's position in stream
 * order inside the chunk (4:4:4: Y.0 Y.1 Cb Cr per MCU; 4:2:2: Y0.0 Y0.1 Y1.0 Y1.1 Cb Cr). */
template <int RATIO, bool ROOMY = false>
struct KeCfg {
    static const int NY = (RATIO == 0) ? 64 : (RATIO == 1 ? 128 : 256); /* luma blocks per chunk */
    static const int ROWS = NY + 128;                    /* coefficient rows (blocks) per chunk */
#ifndef KE_SPLIT
#define KE_SPLIT 0 /* 1: two lanes per luma block.  Measured on the headline frames: 96.5 us split, 91.0 us one lane per block
                      (the second lane halves the luma loop but pays the per-entry work -- masks, scan, relocation -- again);
                      iid noise: 85 k against 92 k Mpixel/s */
#endif
    static const bool SPLIT = (RATIO != 2) && KE_SPLIT;
    static const int NT = (SPLIT ? 2 : 1) * NY + 128;    /* threads = entries: 256 / 384 / 384 */
    static const int EPM = (RATIO == 0) ? 4 : 6;         /* entries per MCU */
#ifndef KE_TABLES_GLOBAL
#define KE_TABLES_GLOBAL 0 /* 1: the code tables are read from global memory (L1) instead of a per-CTA copy in shared memory */
#endif
#ifndef KE_AC_GLOBAL
#define KE_AC_GLOBAL 1 /* 1: only lut + dc (1152 bytes) are copied to shared memory, the (run, category) table stays in global memory */
#endif
    /* words of private stream per entry.  KE lives on occupancy (shared memory per CTA): 4:4:4 with 5 words and the
     * (run, category) table left in global memory is 31.9 KB -> SEVEN CTAs per SM (72.6 us per 8K frame against
     * 77.0 us with 10 words and all tables in shared memory: six CTAs); 4:2:2 with 8 words is 45.2 KB -> five CTAs.
     * A block that outgrows its region is coded a second time, alone (see below).
     * ROOMY: 10 words for every sampling (4:4:4: six CTAs per SM) -- the form for busy content, where most luma blocks
     * would outgrow 5 words and be coded twice; the host switches between the two forms from the re-code count. */
#ifdef KE_PRIV
    static const int PRIV = KE_PRIV;
#else
    static const int PRIV = ROOMY ? 10 : ((RATIO == 0) ? 5 : (RATIO == 1 ? 8 : 10));
#endif
    static const int PSTRIDE = PRIV + 1;                 /* odd: lanes hit distinct banks */
    static const int ROW_BYTES = ROWS * KE_ROW_CHUNKS * 16;
    static const int STAGE_WORDS = ROW_BYTES / 4;        /* the chunk stream reuses the coefficient rows */
    static const int OFF_PRIV = ROW_BYTES;
    static const int OFF_H = OFF_PRIV + NT * PSTRIDE * 4;
    /* the roomy form has the shared memory for the whole table set (and busy content uses the (run, category) table
     * for most of its symbols); the tight form keeps lut + dc only */
    static const bool AC_GLOBAL = KE_AC_GLOBAL && !(ROOMY && RATIO != 2);
    static const int TABLE_BYTES = KE_TABLES_GLOBAL ? 0 : (AC_GLOBAL ? (int)(sizeof(uint32_t) * (2 * 128 + 2 * 16)) : (int)sizeof(HuffPack));
    static const int OFF_OFF = OFF_H + TABLE_BYTES;
    static const int OFF_EX = OFF_OFF + NT * 4;
    static const int OFF_WARP = OFF_EX + NT * 4;
    static const int OFF_FLAG = OFF_WARP + (NT / 32) * 4;
#ifndef KE_PAD_BYTES
#define KE_PAD_BYTES 0 /* occupancy experiments */
#endif
    static const int BYTES = OFF_FLAG + 16 + KE_PAD_BYTES;
    /* resident CTAs per SM the shared memory allows (228 KB per SM, 1 KB reserved per CTA): the register budget follows */
    static const int CTAS = (233472 / (BYTES + 1024)) * NT > 2048 ? 2048 / NT : 233472 / (BYTES + 1024);
};
template <int RATIO, bool ROOMY = false>
using KeSmem = KeCfg<RATIO, ROOMY>;

template <int RATIO>
struct KeMap {
    int row, part, comp, e_local, blk_local; /* blk_local: the block's position in scan order, -1 for part 1 */
    bool valid, first_of_comp, luma;
    long long gabs; /* block index inside zz */
    __device__ __forceinline__ KeMap(const EcParams &p, long long m0, int tid)
    {
        constexpr int NY = KeCfg<RATIO>::NY;
        if (RATIO != 2 && !KeCfg<RATIO>::SPLIT) {
            /* one lane per block, component-major */
            part = 0;
            if (tid < NY) {
                luma = true;
                comp = 0;
                row = tid;
                const int m = (RATIO == 0) ? row : (row >> 1);
                e_local = blk_local = (RATIO == 0) ? 3 * m : 4 * m + (row & 1);
                gabs = ((RATIO == 0) ? m0 : 2 * m0) + row;
                valid = (m0 + m) < p.n_mcu;
                first_of_comp = (row == 0);
            } else {
                luma = false;
                const int c = (tid - NY) >> 6, m = (tid - NY) & 63;
                comp = 1 + c;
                row = NY + 64 * c + m;
                e_local = blk_local = (RATIO == 0) ? 3 * m + 1 + c : 4 * m + 2 + c;
                gabs = m0 + m + (c == 0 ? p.coff1 : p.coff2);
                valid = (m0 + m) < p.n_mcu;
                first_of_comp = (m == 0);
            }
        } else if (RATIO == 2) {
            /* 4:2:0: one lane per block; luma rows are already in stream order (MCU-major) */
            part = 0;
            if (tid < NY) {
                luma = true;
                comp = 0;
                row = tid;
                const int m = row >> 2;
                e_local = blk_local = 6 * m + (row & 3);
                gabs = 4 * m0 + row;
                valid = (m0 + m) < p.n_mcu;
                first_of_comp = (row == 0);
            } else {
                luma = false;
                const int c = (tid - NY) >> 6, m = (tid - NY) & 63;
                comp = 1 + c;
                row = NY + 64 * c + m;
                e_local = blk_local = 6 * m + 4 + c;
                gabs = m0 + m + (c == 0 ? p.coff1 : p.coff2);
                valid = (m0 + m) < p.n_mcu;
                first_of_comp = (m == 0);
            }
        } else if (tid < 2 * NY) {
            luma = true;
            comp = 0;
            row = tid >> 1;
            part = tid & 1;
            const int m = (RATIO == 0) ? row : (row >> 1);
            e_local = (RATIO == 0) ? 4 * m + part : 6 * m + 2 * (row & 1) + part;
            blk_local = part ? -1 : ((RATIO == 0) ? 3 * m : 4 * m + (row & 1));
            gabs = ((RATIO == 0) ? m0 : 2 * m0) + row;
            valid = (m0 + m) < p.n_mcu;
            first_of_comp = (row == 0);
        } else {
            luma = false;
            part = 0;
            const int c = (tid - 2 * NY) >> 6, m = (tid - 2 * NY) & 63;
            comp = 1 + c;
            row = NY + 64 * c + m;
            e_local = (RATIO == 0) ? 4 * m + 2 + c : 6 * m + 4 + c;
            blk_local = (RATIO == 0) ? 3 * m + 1 + c : 4 * m + 2 + c;
            gabs = m0 + m + (c == 0 ? p.coff1 : p.coff2);
            valid = (m0 + m) < p.n_mcu;
            first_of_comp = (m == 0);
        }
    }
};

template <int RATIO, bool ROOMY = false>
__global__ void __launch_bounds__(KeCfg<RATIO, ROOMY>::NT, KeCfg<RATIO, ROOMY>::CTAS) k_encode_chunks(const KeParams q0, const HuffPack *huff)
{
    /* frame of a multi-frame launch (blockIdx.y): every array advances by its stride.  The parameters are read where
     * they are used (constant bank) -- a copy of the struct would cost every warp its set-up. */
    const KeParams &q = q0;
    const long long fr = blockIdx.y;
    const int16_t *const zz_f = q0.zz + fr * q0.zz_fstride;
    uint32_t *const tmp_f = q0.tmp + fr * q0.tmp_fstride;
    uint32_t *const cta_bits_f = q0.cta_bits + fr * q0.cta_fstride;
    uint32_t *const cta_tail_f = q0.cta_tail + fr * q0.cta_fstride;
    uint32_t *const super_bits_f = q0.super_bits + fr * q0.super_fstride;
    constexpr int BPM = EcCfg<RATIO>::BPM, NT = KeCfg<RATIO, ROOMY>::NT;
    typedef KeCfg<RATIO, ROOMY> L;
    extern __shared__ __align__(16) unsigned char ke_smem[];
    uint4(*sC)[KE_ROW_CHUNKS] = reinterpret_cast<uint4(*)[KE_ROW_CHUNKS]>(ke_smem);
    uint32_t *sP = reinterpret_cast<uint32_t *>(ke_smem + L::OFF_PRIV);
    const HuffPack &sH = KE_TABLES_GLOBAL ? *huff : *reinterpret_cast<const HuffPack *>(ke_smem + L::OFF_H);
    uint32_t *sOff = reinterpret_cast<uint32_t *>(ke_smem + L::OFF_OFF);
    uint32_t *sEx = reinterpret_cast<uint32_t *>(ke_smem + L::OFF_EX);
    uint32_t *sWarp = reinterpret_cast<uint32_t *>(ke_smem + L::OFF_WARP);
    const int tid = threadIdx.x;
    const long long chunk = blockIdx.x;
    const long long m0 = chunk * EC_MCUS;

    EcParams p;
    p.zz = zz_f;
    p.coff1 = q.coff1;
    p.coff2 = q.coff2;
    p.n_mcu = q.n_mcu;
    p.dc_pred[0] = q.dc_pred_dev ? q.dc_pred_dev[0] : q.dc_pred[0];
    p.dc_pred[1] = q.dc_pred_dev ? q.dc_pred_dev[1] : q.dc_pred[1];
    p.dc_pred[2] = q.dc_pred_dev ? q.dc_pred_dev[2] : q.dc_pred[2];

    ke_fill<RATIO, NT>(sC, p, m0, tid);
    if (!KE_TABLES_GLOBAL) ec_load_bytes<NT, L::TABLE_BYTES>(ke_smem + L::OFF_H, huff, tid); /* lut and dc come first in HuffPack */
    const KeMap<RATIO> mp(p, m0, tid);
    cp_async_wait_all();
    __syncthreads();

    /* code every entry once, privately */
    uint32_t nbits = 0;
    int diff = 0;
    uint32_t *priv = sP + tid * L::PSTRIDE;
    /* masks first, by every lane: the two lanes of a split luma block swap their halves by shuffle */
    uint32_t mm[2] = {0u, 0u};
    int prev0 = 0;
    bool eob = true, active;
    const bool split = mp.luma && L::SPLIT; /* warp-uniform: warps are component-pure */
    const int sw = mp.row & 7;
    const unsigned sw4 = (unsigned)sw << 4;
    if (split)
        active = ec_entry_masks<false>(sC[mp.row], sw, mp.part, mm, prev0, eob);
    else
        active = ec_entry_masks<true>(sC[mp.row], sw, 0, mm, prev0, eob);
    active = active && mp.valid;
    const uint32_t *tAc = (L::AC_GLOBAL ? huff->ac : sH.ac)[mp.luma ? 0 : 1], *tDc = sH.dc[mp.luma ? 0 : 1], *tLut = sH.lut[mp.luma ? 0 : 1];
    if (mp.valid) {
        if (mp.part == 0) {
            int pred;
            if (!mp.first_of_comp)
                pred = (int)(short)(sC[mp.row - 1][(mp.row - 1) & 7].x & 0xffffu);
            else if (m0 == 0)
                pred = mp.comp == 0 ? p.dc_pred[0] : (mp.comp == 1 ? p.dc_pred[1] : p.dc_pred[2]);
            else
                pred = (int)p.zz[(mp.gabs - 1) * 64];
            diff = (int)(short)(sC[mp.row][sw].x & 0xffffu) - pred;
        }
    }
    if (active) {
        PrivSinkBF s;
        const unsigned pa = (unsigned)__cvta_generic_to_shared(priv);
        s.acc = 0;
        s.nacc = 0;
        s.wa = pa;
        s.ea = pa + 4u * (unsigned)L::PRIV; /* PSTRIDE = PRIV + 1: the pad word */
        ec_pack_masked(sC[mp.row], sw4, mp.part == 0, diff, mm, prev0, eob, tAc, tDc, tLut, s);
        nbits = s.finish(pa);
    }
    sOff[mp.e_local] = nbits; /* stream order */
    __syncthreads(); /* every entry is coded: the coefficient rows are dead from here on */
    uint32_t total;
    {
        /* thread t scans stream position t; the result goes to a second array, so the lengths need
         * no barrier of their own before they are overwritten */
        const uint32_t ex = cta_exclusive_scan<NT>(sOff[tid], sWarp, tid, total);
        sEx[tid] = ex;
    }
    __syncthreads();
    if (q.local_off && mp.valid && mp.blk_local >= 0) q.local_off[m0 * BPM + mp.blk_local] = sEx[mp.e_local];
    if (tid == 0) {
        cta_bits_f[chunk] = total;
        atomicAdd(&super_bits_f[chunk >> EC_SUPER_SHIFT], total);
    }
    const uint32_t nwords = (total + 31u) >> 5;
    /* The chunk stream is assembled in shared memory at the END of the (dead) coefficient rows, 16-byte aligned and
     * zero-padded to whole 16-byte groups: the rows in front of it -- the luma rows, which are the ones that
     * outgrow a private region -- stay readable. */
    const uint32_t ngroups = (nwords + 4u) >> 2; /* covers word [nwords], which may receive an empty OR */
    const bool staged = ngroups * 4u <= (uint32_t)L::STAGE_WORDS;
    uint32_t *slot = tmp_f + chunk * q.slot_words;
    const uint32_t loff = sEx[mp.e_local];
    if (staged) {
        const uint32_t stage0 = (uint32_t)L::STAGE_WORDS - ngroups * 4u; /* first word of the stream */
        uint32_t *sS = reinterpret_cast<uint32_t *>(ke_smem) + stage0;
        uint4 *sS4 = reinterpret_cast<uint4 *>(sS);
        /* a block that outgrew its private region is coded a second time, straight into the staged stream: from
         * its row if the stream does not reach it, else from HBM/L2 */
        const bool again = active && nbits > (uint32_t)L::PRIV * 32u;
        const bool row_alive = ((uint32_t)mp.row + 1u) * 32u <= stage0;
        for (uint32_t i = tid; i < ngroups; i += NT) sS4[i] = make_uint4(0u, 0u, 0u, 0u);
        /* the count the host picks KE's form from: blocks whose string outgrows the TIGHT form's region, in either form */
        /* The count the host picks KE's form from: blocks whose string outgrows the TIGHT form's region.  The tight
         * form counts them where it codes them again (below: nothing on the common path -- a ballot here was measured
         * at 1.9 us per 8K frame); the roomy form, whose own overflows are fewer, counts them here. */
        if (ROOMY && KeCfg<RATIO, false>::PRIV != L::PRIV) {
            const unsigned busy = __ballot_sync(0xffffffffu, active && nbits > (uint32_t)KeCfg<RATIO, false>::PRIV * 32u);
            if (busy && (tid & 31) == 0 && q.recode_count) atomicAdd(q.recode_count, (unsigned)__popc(busy));
        }
        __syncthreads();
        if (active && !again) {
            const uint32_t sh = loff & 31u, nw = (nbits + 31u) >> 5;
            unsigned a0 = (unsigned)__cvta_generic_to_shared(sS + (loff >> 5));
            uint32_t prevw = 0u;
            for (uint32_t i = 0; i < nw; i++, a0 += 4u) {
                const uint32_t v = priv[i];
                asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a0), "r"(__funnelshift_r(v, prevw, sh)) : "memory");
                prevw = v;
            }
            if (sh) asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a0), "r"(prevw << (32u - sh)) : "memory");
        } else if (again) {
            if (!(ROOMY && KeCfg<RATIO, false>::PRIV != L::PRIV) && q.recode_count) {
                const unsigned am = __activemask();
                if ((tid & 31) == __ffs((int)am) - 1) atomicAdd(q.recode_count, (unsigned)__popc(am));
            }
            StageSink s;
            s.acc = 0;
            s.nacc = (int)(loff & 31u);
            s.wp = sS + (loff >> 5);
            s.first = true;
            /* the row goes to thread-local memory first (eight 16-byte loads in flight together): a dead row
             * would otherwise cost one dependent L2 round trip per coefficient */
            uint4 loc[8];
            const uint4 *src = row_alive ? sC[mp.row] : reinterpret_cast<const uint4 *>(p.zz + mp.gabs * 64);
            const int swz = row_alive ? sw : 0;
#pragma unroll
            for (int j = 0; j < 8; j++) loc[j] = src[j ^ swz];
            ec_pack_masked(loc, 0u, mp.part == 0, diff, mm, prev0, eob, tAc, tDc, tLut, s);
            s.finish();
        }
        __syncthreads();
        uint4 *slot4 = reinterpret_cast<uint4 *>(slot);
        for (uint32_t i = tid; i < ((nwords + 3u) >> 2); i += NT) {
            const uint4 v = sS4[i];
            slot4[i] = make_uint4(__byte_perm(v.x, 0, 0x0123), __byte_perm(v.y, 0, 0x0123), __byte_perm(v.z, 0, 0x0123),
                                  __byte_perm(v.w, 0, 0x0123));
        }
        if (tid == 0) cta_tail_f[chunk] = stream_tail32(sS, total, false);
        return;
    }
    /* rare: straight into the slot */
    for (uint32_t i = tid; i < nwords; i += NT) slot[i] = 0u;
    __syncthreads();
    if (active) {
        PackSink<false> s;
        s.acc = 0;
        s.nacc = (int)(loff & 31u);
        s.wp = slot + (loff >> 5);
        s.first = true;
        ec_pack_masked(sC[mp.row], sw4, mp.part == 0, diff, mm, prev0, eob, tAc, tDc, tLut, s);
        s.finish();
    }
    __syncthreads();
    if (tid == 0) cta_tail_f[chunk] = stream_tail32(slot, total, true);
}

}  // namespace jpgb200
#endif
