/*
 * Entropy stage: K3 (per-block bit lengths + per-CTA offsets), K4 (device-wide exclusive
 * scan), K5 (parallel bit pack).  The reference has no bit writer at all (SURVEY.md F1,
 * F2): what it pins is the symbol alphabet -- categories = bit length of |v|, EOB 0x00
 * when zeros trail the last non-zero coefficient, ZRL 0xF0 per 16 zeros
 * (src/huffman.c:182-235) -- and the code tables (src/headers/tables.h:37-72).  The
 * emitted stream follows T.81 F.1.2: true DC differences and (run<<4)|size symbols.
 *
 * Scan order is the interleaved MCU order (T.81 A.2): 4:4:4 -> Y Cb Cr per MCU,
 * 4:2:2 -> Y(2m) Y(2m+1) Cb(m) Cr(m).  One thread encodes one 8x8 block; a CTA owns
 * EC_MCUS consecutive MCUs, whose coefficients are contiguous per component in HBM and are
 * staged through shared memory with coalesced 16-byte loads (rows of 128 B, XOR-swizzled so
 * that the per-thread row reads are conflict-free).
 */
#ifndef JPGB200_ENTROPY_CUH
#define JPGB200_ENTROPY_CUH

#include <stdint.h>

namespace jpgb200 {

#define EC_MCUS 64

struct HuffDev {
    uint32_t ac[2][256]; /* (code << 8) | code_length ; [lum|chr][(run<<4)|size] */
    uint32_t dc[2][16];  /* (code << 8) | code_length ; [lum|chr][category]     */
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
};

struct EcParams {
    const int16_t *zz_y, *zz_cb, *zz_cr;
    long long n_mcu;
    int dc_pred[3];        /* predictor of the first block per component (0 for a whole image) */
    int is_last_band;      /* pad the final byte after the last block                         */
    uint32_t *local_off;   /* [n_mcu*BPM] bit offset of each block inside its CTA chunk        */
    uint32_t *cta_bits;    /* [n_cta] bits per CTA chunk                                       */
    const unsigned long long *cta_base; /* [n_cta] exclusive scan of cta_bits (K5)            */
    unsigned long long bit_base;        /* global bit offset of this band (multi-GPU)         */
    uint32_t *ubuf;        /* un-stuffed big-endian bit stream, zero-initialised (K5)          */
    const EncState *st;    /* error flag set by K4 when the stream would not fit ubuf          */
};

template <int RATIO>
struct EcCfg {
    static const int BPM = (RATIO == 0) ? 3 : 4;
    static const int NT = EC_MCUS * BPM;
};

/* Thread <-> block mapping inside a CTA.  Threads (= shared-memory rows) are grouped by
 * COMPONENT, not by MCU: luma blocks carry ~3x the non-zero coefficients of chroma blocks,
 * and the per-block loops run as long as the busiest lane of a warp, so component-pure warps
 * execute far fewer divergent iterations.  s_local is the block's position in scan (MCU
 * interleaved) order inside the CTA's chunk. */
template <int RATIO>
struct EcMap {
    int comp, s_local;
    long long gblk; /* block index inside its component's coefficient array */
    bool valid, first_of_comp;
    __device__ __forceinline__ EcMap(const EcParams &p, long long m0, int tid)
    {
        if (RATIO == 0) {
            comp = tid >> 6;
            const int m = tid & 63;
            s_local = m * 3 + comp;
            gblk = m0 + m;
            valid = (m0 + m) < p.n_mcu;
            first_of_comp = (m == 0);
        } else {
            if (tid < 128) {
                comp = 0;
                s_local = (tid >> 1) * 4 + (tid & 1);
                gblk = 2 * m0 + tid;
                valid = (m0 + (tid >> 1)) < p.n_mcu;
                first_of_comp = (tid == 0);
            } else {
                comp = 1 + ((tid - 128) >> 6);
                const int m = tid & 63;
                s_local = m * 4 + 1 + comp;
                gblk = m0 + m;
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

/* coalesced asynchronous fill of the CTA's coefficient rows (row == thread index,
 * component-major); 16-byte chunk q of a row lives at chunk (q ^ (row & 7)) */
template <int RATIO>
__device__ __forceinline__ void ec_fill(uint4 (*sC)[8], const EcParams &p, long long m0, int tid)
{
    constexpr int NT = EcCfg<RATIO>::NT;
    long long left = p.n_mcu - m0;
    const int nv = left < EC_MCUS ? (int)left : EC_MCUS; /* valid MCUs in this CTA */
    if (RATIO == 0) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const int16_t *z = (c == 0) ? p.zz_y : (c == 1 ? p.zz_cb : p.zz_cr);
            const uint4 *src = reinterpret_cast<const uint4 *>(z) + m0 * 8;
            for (int i = tid; i < nv * 8; i += NT) {
                const int row = c * 64 + (i >> 3), q = i & 7;
                cp_async16(&sC[row][q ^ (row & 7)], src + i);
            }
        }
    } else {
        {
            const uint4 *src = reinterpret_cast<const uint4 *>(p.zz_y) + m0 * 16;
            for (int i = tid; i < nv * 16; i += NT) {
                const int row = i >> 3, q = i & 7;
                cp_async16(&sC[row][q ^ (row & 7)], src + i);
            }
        }
#pragma unroll
        for (int c = 1; c < 3; c++) {
            const uint4 *src = reinterpret_cast<const uint4 *>(c == 1 ? p.zz_cb : p.zz_cr) + m0 * 8;
            for (int i = tid; i < nv * 8; i += NT) {
                const int row = 64 + c * 64 + (i >> 3), q = i & 7;
                cp_async16(&sC[row][q ^ (row & 7)], src + i);
            }
        }
    }
}

/* DC predictor: the previous block of the same component (raster == scan order per
 * component) -- the row above in shared memory, the caller's predictor for the first block
 * of the image/band, or the last block of the previous chunk read from HBM */
template <int RATIO>
__device__ __forceinline__ int ec_dc_pred(const uint4 (*sC)[8], const EcParams &p, long long m0, int tid,
                                          const EcMap<RATIO> &mp)
{
    if (!mp.first_of_comp) return (int)(short)(sC[tid - 1][(tid - 1) & 7].x & 0xffffu);
    if (m0 == 0) return p.dc_pred[mp.comp];
    const int16_t *z = (mp.comp == 0) ? p.zz_y : (mp.comp == 1 ? p.zz_cb : p.zz_cr);
    return (int)z[(mp.gblk - 1) * 64];
}

struct CountSink {
    uint32_t bits;
    __device__ __forceinline__ void put(uint32_t, int n) { bits += (uint32_t)n; }
};

/* Bit writer: big-endian 32-bit words.  SHARED selects shared-memory atomics for the two
 * words a block can share with its neighbours; interior words are owned exclusively. */
template <bool SHARED>
struct PackSink {
    unsigned long long acc;
    int nacc;
    uint32_t *wp;
    bool first;
    __device__ __forceinline__ void or_word(uint32_t be)
    {
        if (SHARED)
            atomicOr(wp, be);
        else
            atomicOr(wp, be);
    }
    __device__ __forceinline__ void emit(uint32_t word)
    {
        const uint32_t be = __byte_perm(word, 0, 0x0123);
        if (first) {
            or_word(be); /* word shared with the preceding block */
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
        if (nacc > 0) {
            const uint32_t word = (uint32_t)(acc << (32 - nacc));
            or_word(__byte_perm(word, 0, 0x0123)); /* word shared with the following block */
        }
    }
};

/* int16 coefficient k of a swizzled 128-byte row */
__device__ __forceinline__ int ec_coef(const uint4 *row, int key, int k)
{
    const char *base = reinterpret_cast<const char *>(row);
    return (int)*reinterpret_cast<const short *>(base + ((((k >> 3) ^ key) << 4) | ((k & 7) << 1)));
}

/* One block: DC difference + AC run/size symbols (T.81 F.1.2.1-F.1.2.2).
 * row = this block's 128-byte shared-memory row (8 swizzled 16-byte chunks, key = row & 7);
 * sAc/sDc = (code<<8)|len tables of this component.
 * The 63 AC positions are first reduced to a bit mask of non-zero coefficients (2 predicated
 * instructions per coefficient, no branches); the symbol loop then visits only set bits, so
 * its trip count is the block's number of non-zero coefficients rather than 63 -- in a warp,
 * the maximum over 32 blocks of the same component. */
template <class Sink>
__device__ __forceinline__ void ec_encode_block(const uint4 *row, int key, int diff, const uint32_t *sAc,
                                                const uint32_t *sDc, Sink &s)
{
    {
        const int cat = 32 - __clz(abs(diff));
        const uint32_t e = sDc[cat];
        const uint32_t extra = (uint32_t)(diff - (diff < 0 ? 1 : 0)) & ((1u << cat) - 1u);
        s.put(((e >> 8) << cat) | extra, (int)(e & 0xffu) + cat);
    }
    uint32_t m[2] = {0u, 0u};
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const uint4 q = row[j ^ key];
        const uint32_t ws[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int k = 8 * j + 2 * i;
            if (ws[i] & 0x0000ffffu) m[k >> 5] |= 1u << (k & 31);
            if (ws[i] & 0xffff0000u) m[(k + 1) >> 5] |= 1u << ((k + 1) & 31);
        }
    }
    m[0] &= ~1u; /* position 0 is the DC */
    const uint32_t zrl = sAc[0xF0], eob = sAc[0x00];
    int prev = 0;
#pragma unroll
    for (int half = 0; half < 2; half++) {
        uint32_t mm = m[half];
        while (mm) {
            const int k = half * 32 + (__ffs((int)mm) - 1);
            mm &= mm - 1u;
            int run = k - prev - 1;
            prev = k;
            const int v = ec_coef(row, key, k);
            while (run > 15) {
                s.put(zrl >> 8, (int)(zrl & 0xffu));
                run -= 16;
            }
            const int cat = 32 - __clz(abs(v));
            const uint32_t e = sAc[(run << 4) | cat];
            const uint32_t extra = (uint32_t)(v - (v < 0 ? 1 : 0)) & ((1u << cat) - 1u);
            s.put(((e >> 8) << cat) | extra, (int)(e & 0xffu) + cat);
        }
    }
    if (prev != 63) s.put(eob >> 8, (int)(eob & 0xffu));
}

/* Huffman tables -> shared memory (same layout as HuffDev: 2176 bytes = 136 x 16) */
__device__ __forceinline__ void ec_load_tables(HuffDev *sH, const HuffDev *h, int tid)
{
    if (tid < (int)(sizeof(HuffDev) / 16))
        cp_async16(reinterpret_cast<uint4 *>(sH) + tid, reinterpret_cast<const uint4 *>(h) + tid);
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

/* ---- K3: bit length of every block, exclusive offsets inside the CTA chunk, chunk total -- */
template <int RATIO>
__global__ void __launch_bounds__(EcCfg<RATIO>::NT) k3_block_bits(const EcParams p, const HuffDev *huff)
{
    constexpr int BPM = EcCfg<RATIO>::BPM, NT = EcCfg<RATIO>::NT;
    __shared__ uint4 sC[NT][8];
    __shared__ __align__(16) HuffDev sH;
    __shared__ uint32_t sBits[NT];
    __shared__ uint32_t sWarp[NT / 32];
    const int tid = threadIdx.x;
    const long long m0 = (long long)blockIdx.x * EC_MCUS;

    ec_fill<RATIO>(sC, p, m0, tid);
    ec_load_tables(&sH, huff, tid);
    cp_async_wait_all();
    __syncthreads();

    const EcMap<RATIO> mp(p, m0, tid);
    uint32_t nbits = 0;
    if (mp.valid) {
        const int pred = ec_dc_pred<RATIO>(sC, p, m0, tid, mp);
        const int diff = (int)(short)(sC[tid][tid & 7].x & 0xffffu) - pred;
        CountSink s{0};
        ec_encode_block(sC[tid], tid & 7, diff, sH.ac[mp.comp ? 1 : 0], sH.dc[mp.comp ? 1 : 0], s);
        nbits = s.bits;
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
#define K5_STAGE_WORDS 3072 /* 12 KB = 98304 bits: ~4x the typical chunk at q75 */

template <int RATIO>
__global__ void __launch_bounds__(EcCfg<RATIO>::NT) k5_pack(const EcParams p, const HuffDev *huff)
{
    constexpr int BPM = EcCfg<RATIO>::BPM, NT = EcCfg<RATIO>::NT;
    __shared__ uint4 sC[NT][8];
    __shared__ __align__(16) HuffDev sH;
    __shared__ uint32_t sS[K5_STAGE_WORDS];
    const int tid = threadIdx.x;
    const long long m0 = (long long)blockIdx.x * EC_MCUS;

    /* all global reads of the prologue are issued back to back, none depends on another */
    ec_fill<RATIO>(sC, p, m0, tid);
    ec_load_tables(&sH, huff, tid);
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
    const bool staged = nwords <= K5_STAGE_WORDS;
    if (staged)
        for (uint32_t i = tid; i < nwords; i += NT) sS[i] = 0u;
    cp_async_wait_all();
    __syncthreads();

    if (mp.valid) {
        const int pred = ec_dc_pred<RATIO>(sC, p, m0, tid, mp);
        const int diff = (int)(short)(sC[tid][tid & 7].x & 0xffffu) - pred;
        const uint32_t loff = b0 + my_off;
        const bool very_last = p.is_last_band && (m0 * BPM + mp.s_local == p.n_mcu * BPM - 1);
        if (staged) {
            PackSink<true> s;
            s.acc = 0;
            s.nacc = (int)(loff & 31u);
            s.wp = sS + (loff >> 5);
            s.first = true;
            ec_encode_block(sC[tid], tid & 7, diff, sH.ac[mp.comp ? 1 : 0], sH.dc[mp.comp ? 1 : 0], s);
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
            ec_encode_block(sC[tid], tid & 7, diff, sH.ac[mp.comp ? 1 : 0], sH.dc[mp.comp ? 1 : 0], s);
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
 * 8 items per thread: 8192 chunks (one 8K 4:4:4 frame) per pass.  `each(i, base, v)` sees
 * every element with its exclusive prefix while both are still in registers. */
#define SCAN_THREADS 1024
#define SCAN_ITEMS 8
template <class F>
__device__ __forceinline__ unsigned long long scan_u32_to_u64(const uint32_t *in, unsigned long long *out,
                                                              long long n, uint32_t *sWarp, F each)
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
            if (i0 + k < n) {
                if (out) out[i0 + k] = carry + ex;
                each(i0 + k, carry + ex, v[k]);
            }
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
    const unsigned long long total =
        scan_u32_to_u64(cta_bits, cta_base, n, sWarp, [](long long, unsigned long long, uint32_t) {});
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
