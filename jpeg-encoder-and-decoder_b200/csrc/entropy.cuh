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

/* coalesced fill of the CTA's coefficient rows; row index == thread index of the block */
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
                const int row = (i >> 3) * 3 + c, q = i & 7;
                sC[row][q ^ (row & 7)] = __ldg(src + i);
            }
        }
    } else {
        {
            const uint4 *src = reinterpret_cast<const uint4 *>(p.zz_y) + m0 * 16;
            for (int i = tid; i < nv * 16; i += NT) {
                const int blk = i >> 3, q = i & 7;
                const int row = (blk >> 1) * 4 + (blk & 1);
                sC[row][q ^ (row & 7)] = __ldg(src + i);
            }
        }
#pragma unroll
        for (int c = 1; c < 3; c++) {
            const uint4 *src = reinterpret_cast<const uint4 *>(c == 1 ? p.zz_cb : p.zz_cr) + m0 * 8;
            for (int i = tid; i < nv * 8; i += NT) {
                const int row = (i >> 3) * 4 + 1 + c, q = i & 7;
                sC[row][q ^ (row & 7)] = __ldg(src + i);
            }
        }
    }
}

/* component (0 Y, 1 Cb, 2 Cr) of the block handled by thread tid, and the DC predictor */
template <int RATIO>
__device__ __forceinline__ int ec_dc_pred(const uint4 (*sC)[8], const EcParams &p, long long m0, int tid, int &comp)
{
    constexpr int BPM = EcCfg<RATIO>::BPM;
    const int mcu = tid / BPM, k = tid - mcu * BPM;
    int prev_row;
    bool in_cta;
    long long gblk; /* global block index of the predecessor inside its component */
    if (RATIO == 0) {
        comp = k;
        prev_row = tid - 3;
        in_cta = mcu > 0;
        gblk = m0 - 1;
    } else {
        comp = (k < 2) ? 0 : k - 1;
        if (k == 1) {
            prev_row = tid - 1;
            in_cta = true;
            gblk = 0;
        } else if (k == 0) {
            prev_row = tid - 3;
            in_cta = mcu > 0;
            gblk = 2 * m0 - 1;
        } else {
            prev_row = tid - 4;
            in_cta = mcu > 0;
            gblk = m0 - 1;
        }
    }
    if (in_cta) return (int)(short)(sC[prev_row][0 ^ (prev_row & 7)].x & 0xffffu);
    if (m0 == 0) return p.dc_pred[comp];
    const int16_t *z = (comp == 0) ? p.zz_y : (comp == 1 ? p.zz_cb : p.zz_cr);
    return (int)z[gblk * 64];
}

struct CountSink {
    uint32_t bits;
    __device__ __forceinline__ void put(uint32_t, int n) { bits += (uint32_t)n; }
};

struct PackSink {
    unsigned long long acc;
    int nacc;
    uint32_t *wp;
    bool first;
    __device__ __forceinline__ void emit(uint32_t word)
    {
        const uint32_t be = __byte_perm(word, 0, 0x0123);
        if (first) {
            atomicOr(wp, be); /* word shared with the preceding block */
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
            atomicOr(wp, __byte_perm(word, 0, 0x0123)); /* word shared with the following block */
        }
    }
};

/* One block: DC difference + AC run/size symbols (T.81 F.1.2.1-F.1.2.2).
 * row = this block's 128-byte shared-memory row (8 swizzled 16-byte chunks, key = row & 7);
 * sAc/sDc = (code<<8)|len tables of this component.  The chunk loop is deliberately not
 * unrolled: 8 coefficient bodies of code instead of 63 keep the kernel inside the
 * instruction cache. */
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
    const uint32_t zrl = sAc[0xF0], eob = sAc[0x00];
    int run = -1; /* the masked DC slot below is walked as one zero */
#pragma unroll 1
    for (int j = 0; j < 8; j++) {
        uint4 q = row[j ^ key];
        if (j == 0) q.x &= 0xffff0000u; /* DC handled above */
        if ((q.x | q.y | q.z | q.w) == 0u) {
            run += 8;
            continue;
        }
        const uint32_t ws[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t ww = ws[i];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int v = h ? ((int)ww >> 16) : ((int)(ww << 16) >> 16);
                if (v == 0) {
                    run++;
                } else {
                    while (run > 15) {
                        s.put(zrl >> 8, (int)(zrl & 0xffu));
                        run -= 16;
                    }
                    const int cat = 32 - __clz(abs(v));
                    const uint32_t e = sAc[(run << 4) | cat];
                    const uint32_t extra = (uint32_t)(v - (v < 0 ? 1 : 0)) & ((1u << cat) - 1u);
                    s.put(((e >> 8) << cat) | extra, (int)(e & 0xffu) + cat);
                    run = 0;
                }
            }
        }
    }
    if (run > 0) s.put(eob >> 8, (int)(eob & 0xffu));
}

template <int RATIO>
__device__ __forceinline__ void ec_load_tables(uint32_t (*sAc)[256], uint32_t (*sDc)[16], const HuffDev *h, int tid)
{
    constexpr int NT = EcCfg<RATIO>::NT;
    for (int i = tid; i < 512; i += NT) sAc[i >> 8][i & 255] = h->ac[i >> 8][i & 255];
    if (tid < 32) sDc[tid >> 4][tid & 15] = h->dc[tid >> 4][tid & 15];
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
    __shared__ uint32_t sAc[2][256];
    __shared__ uint32_t sDc[2][16];
    __shared__ uint32_t sWarp[NT / 32];
    const int tid = threadIdx.x;
    const long long m0 = (long long)blockIdx.x * EC_MCUS;

    ec_load_tables<RATIO>(sAc, sDc, huff, tid);
    ec_fill<RATIO>(sC, p, m0, tid);
    __syncthreads();

    const bool valid = (m0 + tid / BPM) < p.n_mcu;
    uint32_t nbits = 0;
    if (valid) {
        int comp;
        const int pred = ec_dc_pred<RATIO>(sC, p, m0, tid, comp);
        const int diff = (int)(short)(sC[tid][tid & 7].x & 0xffffu) - pred;
        CountSink s{0};
        ec_encode_block(sC[tid], tid & 7, diff, sAc[comp ? 1 : 0], sDc[comp ? 1 : 0], s);
        nbits = s.bits;
    }
    uint32_t total;
    const uint32_t off = cta_exclusive_scan<NT>(nbits, sWarp, tid, total);
    if (valid) p.local_off[m0 * BPM + tid] = off;
    if (tid == 0) p.cta_bits[blockIdx.x] = total;
}

/* ---- K5: pack every block's codes at its global bit offset ------------------------------- */
template <int RATIO>
__global__ void __launch_bounds__(EcCfg<RATIO>::NT) k5_pack(const EcParams p, const HuffDev *huff)
{
    constexpr int BPM = EcCfg<RATIO>::BPM, NT = EcCfg<RATIO>::NT;
    __shared__ uint4 sC[NT][8];
    __shared__ uint32_t sAc[2][256];
    __shared__ uint32_t sDc[2][16];
    const int tid = threadIdx.x;
    const long long m0 = (long long)blockIdx.x * EC_MCUS;

    ec_load_tables<RATIO>(sAc, sDc, huff, tid);
    ec_fill<RATIO>(sC, p, m0, tid);
    __syncthreads();

    const bool valid = (m0 + tid / BPM) < p.n_mcu;
    if (!valid || p.st->error) return;
    int comp;
    const int pred = ec_dc_pred<RATIO>(sC, p, m0, tid, comp);
    const int diff = (int)(short)(sC[tid][tid & 7].x & 0xffffu) - pred;
    const long long s_idx = m0 * BPM + tid;
    const unsigned long long off = p.bit_base + p.cta_base[blockIdx.x] + p.local_off[s_idx];
    PackSink s;
    s.acc = 0;
    s.nacc = (int)(off & 31);
    s.wp = p.ubuf + (off >> 5);
    s.first = true;
    ec_encode_block(sC[tid], tid & 7, diff, sAc[comp ? 1 : 0], sDc[comp ? 1 : 0], s);
    if (p.is_last_band && s_idx == p.n_mcu * BPM - 1) {
        /* complete the final byte with 1-bits (T.81 F.1.2.3) */
        const int rem = s.nacc & 7;
        if (rem) s.put((1u << (8 - rem)) - 1u, 8 - rem);
    }
    s.finish();
}

/* ---- K4: single-CTA exclusive scan of per-chunk totals (uint32 -> uint64) ---------------- */
#define SCAN_THREADS 1024
__device__ __forceinline__ unsigned long long scan_u32_to_u64(const uint32_t *in, unsigned long long *out,
                                                              long long n, uint32_t *sWarp)
{
    const int tid = threadIdx.x;
    unsigned long long carry = 0;
    for (long long base = 0; base < n; base += SCAN_THREADS) {
        const long long i = base + tid;
        const uint32_t v = (i < n) ? in[i] : 0u;
        uint32_t total;
        const uint32_t ex = cta_exclusive_scan<SCAN_THREADS>(v, sWarp, tid, total);
        if (i < n) out[i] = carry + ex;
        carry += total;
        __syncthreads();
    }
    return carry;
}

/* mode 0: chunk bits -> chunk bases, total bits, byte and tile counts */
__global__ void __launch_bounds__(SCAN_THREADS) k4_scan_chunks(const uint32_t *cta_bits, unsigned long long *cta_base,
                                                               long long n, EncState *st, int tile_bytes,
                                                               unsigned long long ubuf_cap_bytes)
{
    __shared__ uint32_t sWarp[SCAN_THREADS / 32];
    const unsigned long long total = scan_u32_to_u64(cta_bits, cta_base, n, sWarp);
    if (threadIdx.x == 0) {
        /* the workspace is sized for the worst case, so this only trips on corrupted lengths;
         * it keeps every later kernel inside its buffers */
        const bool fits = ((total + 7) >> 3) + 64ull <= ubuf_cap_bytes;
        st->total_bits = fits ? total : 0ull;
        st->nbytes = fits ? (total + 7) >> 3 : 0ull;
        st->n_tiles = (st->nbytes + (unsigned long long)tile_bytes - 1) / (unsigned long long)tile_bytes;
        st->total_ff = 0;
        st->out_len = 0;
        st->error = fits ? 0 : 1;
    }
}

/* zero the words the pack kernel will OR into */
__global__ void __launch_bounds__(256) k_zero_ubuf(uint4 *ubuf, const EncState *st)
{
    const unsigned long long n16 = (st->total_bits >> 7) + 2;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride)
        ubuf[i] = make_uint4(0, 0, 0, 0);
}

}  // namespace jpgb200
#endif
