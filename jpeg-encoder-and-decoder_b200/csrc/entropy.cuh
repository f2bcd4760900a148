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

/* pack tables */
struct HuffDev {
    uint32_t ac[2][256]; /* (code << 8) | code_length ; [lum|chr][(run<<4)|size] */
    uint32_t dc[2][16];  /* (code << 8) | code_length ; [lum|chr][category]     */
};
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
    int is_last_band;      /* pad the final byte after the last block                         */
    uint32_t *local_off;   /* [n_mcu*BPM] bit offset of each block inside its chunk            */
    uint32_t *cta_bits;    /* [n_chunks] bits per chunk                                        */
    const unsigned long long *cta_base; /* [n_chunks] exclusive scan of cta_bits (K5)         */
    unsigned long long bit_base;        /* global bit offset of this band (multi-GPU)         */
    uint32_t *ubuf;        /* un-stuffed big-endian bit stream, all-zero on entry (K5)         */
    const EncState *st;    /* error flag set by K4 when the stream would not fit ubuf          */
};

template <int RATIO>
struct EcCfg {
    static const int BPM = (RATIO == 0) ? 3 : 4;
    static const int NT = EC_MCUS * BPM;
    static const int STAGE_WORDS = (RATIO == 0) ? 3072 : 2048; /* K5 staging buffer */
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

/* asynchronous fill of the chunk's coefficient rows: 8 chunks of 16 B per thread, a warp's
 * 32 copies cover 512 contiguous bytes of one component */
template <int RATIO>
__device__ __forceinline__ void ec_fill(uint4 (*sC)[EC_ROW_CHUNKS], const EcParams &p, long long m0, int tid)
{
    constexpr int NT = EcCfg<RATIO>::NT;
    const uint4 *z = reinterpret_cast<const uint4 *>(p.zz);
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int i = k * NT + tid, row = i >> 3, q = i & 7;
        const EcMap<RATIO> mp(p, m0, row);
        if (mp.valid) cp_async16(&sC[row][q], z + mp.gabs * 8 + q);
    }
}

/* tables -> shared memory, 16 bytes per thread */
template <class T>
__device__ __forceinline__ void ec_load_table(T *sT, const T *g, int tid)
{
    static_assert(sizeof(T) % 16 == 0, "table size");
    if (tid < (int)(sizeof(T) / 16))
        cp_async16(reinterpret_cast<uint4 *>(sT) + tid, reinterpret_cast<const uint4 *>(g) + tid);
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

/* ---- K3: bit length of every block, exclusive offsets inside the chunk, chunk total ------- */
template <int RATIO>
__global__ void __launch_bounds__(EcCfg<RATIO>::NT) k3_block_bits(const EcParams p, const HuffBits *hbits)
{
    constexpr int BPM = EcCfg<RATIO>::BPM, NT = EcCfg<RATIO>::NT;
    __shared__ uint4 sC[NT][EC_ROW_CHUNKS];
    __shared__ __align__(16) HuffBits sB;
    __shared__ uint32_t sBits[NT];
    __shared__ uint32_t sWarp[NT / 32];
    const int tid = threadIdx.x;
    const long long m0 = (long long)blockIdx.x * EC_MCUS;

    ec_fill<RATIO>(sC, p, m0, tid);
    ec_load_table(&sB, hbits, tid);
    const EcMap<RATIO> mp(p, m0, tid);
    cp_async_wait_all();
    __syncthreads();

    uint32_t nbits = 0;
    if (mp.valid) {
        const int pred = ec_dc_pred<RATIO>(sC, p, m0, tid, mp);
        const int diff = (int)(short)(sC[tid][0].x & 0xffffu) - pred;
        nbits = ec_count_block(sC[tid], diff, sB.ac[mp.comp ? 1 : 0], sB.dc[mp.comp ? 1 : 0]);
    }
    sBits[mp.s_local] = nbits; /* scan (MCU-interleaved) order */
    __syncthreads();
    uint32_t total;
    const uint32_t off = cta_exclusive_scan<NT>(sBits[tid], sWarp, tid, total);
    if (m0 * BPM + tid < p.n_mcu * BPM) p.local_off[m0 * BPM + tid] = off;
    if (tid == 0) p.cta_bits[blockIdx.x] = total;
}

/* ---- K5: pack every block's codes at its global bit offset -------------------------------
 * Invariant: the un-stuffed buffer is all-zero on entry -- it is cleared once at allocation
 * and K6 zeroes every 16-byte line right after consuming it.
 * The chunk's stream is assembled in shared memory (at the same bit phase inside a 32-bit
 * word as its destination) and copied out with coalesced stores; only the chunk's first and
 * last word, which neighbouring chunks also write, are OR-ed into HBM.  A chunk too long for
 * the staging buffer packs straight into HBM with atomics on block-boundary words. */
template <int RATIO>
__global__ void __launch_bounds__(EcCfg<RATIO>::NT) k5_pack(const EcParams p, const HuffDev *huff)
{
    constexpr int BPM = EcCfg<RATIO>::BPM, NT = EcCfg<RATIO>::NT, STAGE = EcCfg<RATIO>::STAGE_WORDS;
    __shared__ uint4 sC[NT][EC_ROW_CHUNKS];
    __shared__ __align__(16) HuffDev sH;
    __shared__ uint32_t sS[STAGE];
    const int tid = threadIdx.x;
    const long long m0 = (long long)blockIdx.x * EC_MCUS;

    /* all global reads of the prologue are issued back to back, none depends on another */
    ec_fill<RATIO>(sC, p, m0, tid);
    ec_load_table(&sH, huff, tid);
    const EcMap<RATIO> mp(p, m0, tid);
    const int err = p.st->error;
    const unsigned long long g0 = p.bit_base + p.cta_base[blockIdx.x];
    uint32_t tbits = p.cta_bits[blockIdx.x];
    const uint32_t my_off = mp.valid ? p.local_off[m0 * BPM + mp.s_local] : 0u;
    if (err) return;

    const bool last_chunk = (m0 + EC_MCUS >= p.n_mcu);
    if (last_chunk && p.is_last_band) tbits += (uint32_t)((8u - (uint32_t)((g0 + tbits) & 7ull)) & 7u); /* 1-padding */
    const uint32_t b0 = (uint32_t)(g0 & 31ull);
    const uint32_t nwords = (b0 + tbits + 31u) >> 5;
    uint32_t *gw = p.ubuf + (g0 >> 5);
    const bool staged = nwords <= (uint32_t)STAGE;
    if (staged)
        for (uint32_t i = tid; i < nwords; i += NT) sS[i] = 0u;
    cp_async_wait_all();
    __syncthreads();

    if (mp.valid) {
        const int pred = ec_dc_pred<RATIO>(sC, p, m0, tid, mp);
        const int diff = (int)(short)(sC[tid][0].x & 0xffffu) - pred;
        const uint32_t loff = b0 + my_off;
        const bool very_last = p.is_last_band && (m0 * BPM + mp.s_local == p.n_mcu * BPM - 1);
        if (staged) {
            PackSink<true> s;
            s.acc = 0;
            s.nacc = (int)(loff & 31u);
            s.wp = sS + (loff >> 5);
            s.first = true;
            ec_pack_block(sC[tid], diff, sH.ac[mp.comp ? 1 : 0], sH.dc[mp.comp ? 1 : 0], s);
            if (very_last) {
                const int rem = s.nacc & 7; /* complete the final byte with 1-bits (T.81 F.1.2.3) */
                if (rem) s.put((1u << (8 - rem)) - 1u, 8 - rem);
            }
            s.finish();
        } else {
            PackSink<false> s;
            s.acc = 0;
            s.nacc = (int)(loff & 31u);
            s.wp = gw + (loff >> 5);
            s.first = true;
            ec_pack_block(sC[tid], diff, sH.ac[mp.comp ? 1 : 0], sH.dc[mp.comp ? 1 : 0], s);
            if (very_last) {
                const int rem = s.nacc & 7;
                if (rem) s.put((1u << (8 - rem)) - 1u, 8 - rem);
            }
            s.finish();
        }
    }
    if (!staged) return;
    __syncthreads();
    for (uint32_t i = tid; i < nwords; i += NT) {
        const uint32_t v = sS[i];
        if (i == 0 || i + 1 == nwords) {
            if (v) atomicOr(gw + i, v);
        } else {
            gw[i] = v;
        }
    }
}

/* ---- K4: single-CTA exclusive scan of per-chunk totals (uint32 -> uint64) ----------------
 * 8 items per thread: 8192 chunks (one 8K 4:4:4 frame) per pass. */
#define SCAN_THREADS 1024
#define SCAN_ITEMS 8
__device__ __forceinline__ unsigned long long scan_u32_to_u64(const uint32_t *in, unsigned long long *out,
                                                              long long n, uint32_t *sWarp)
{
    const int tid = threadIdx.x;
    unsigned long long carry = 0;
    for (long long base = 0; base < n; base += SCAN_THREADS * SCAN_ITEMS) {
        const long long i0 = base + (long long)tid * SCAN_ITEMS;
        uint32_t v[SCAN_ITEMS], sum = 0;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            v[k] = (i0 + k < n) ? in[i0 + k] : 0u;
            sum += v[k];
        }
        uint32_t total;
        uint32_t ex = cta_exclusive_scan<SCAN_THREADS>(sum, sWarp, tid, total);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            if (i0 + k < n) out[i0 + k] = carry + ex;
            ex += v[k];
        }
        carry += total;
        __syncthreads();
    }
    return carry;
}

/* chunk bits -> chunk bases, total bits, byte and tile counts */
__global__ void __launch_bounds__(SCAN_THREADS) k4_scan_chunks(const uint32_t *cta_bits, unsigned long long *cta_base,
                                                               long long n, EncState *st, int tile_bytes,
                                                               unsigned long long ubuf_cap_bytes,
                                                               unsigned long long bit_base)
{
    __shared__ uint32_t sWarp[SCAN_THREADS / 32];
    const unsigned long long total = scan_u32_to_u64(cta_bits, cta_base, n, sWarp);
    if (threadIdx.x == 0) {
        /* the workspace is sized for the worst case, so this only trips on corrupted lengths;
         * it keeps every later kernel inside its buffers */
        const bool fits = ((bit_base + total + 7) >> 3) + 64ull <= ubuf_cap_bytes;
        st->total_bits = fits ? total : 0ull;
        st->nbytes = fits ? (total + 7) >> 3 : 0ull;
        st->n_tiles = (st->nbytes + (unsigned long long)tile_bytes - 1) / (unsigned long long)tile_bytes;
        st->total_ff = 0;
        st->out_len = 0;
        st->error = fits ? 0 : 1;
        st->ticket = 0ull;
    }
}

}  // namespace jpgb200
#endif
