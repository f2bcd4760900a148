/*
 * Tail of the entropy stage: chunk streams -> the byte-stuffed JFIF scan, with no intermediate
 * un-stuffed stream in HBM.
 *
 *   KF k_count_ff : a warp per chunk.  The chunk's global bit offset g0 is added up from KE's
 *                   two-level sums (per chunk, per 64 chunks: at most n/64 + 63 numbers, no scan
 *                   kernel, no chain between CTAs).  The chunk OWNS the stream bytes whose last
 *                   bit lies in its bit range, [g0 >> 3, (g0 + T) >> 3); their leading g0 & 7 bits
 *                   are the previous chunk's last bits.  The warp funnel-shifts its chunk to the
 *                   byte grid, counts the 0xFF bytes among the owned ones and leaves the count per
 *                   chunk and (atomically) per 64 chunks.
 *   KS k_scatter  : a warp per chunk again: g0 as above, the number of stuffed zeros before the
 *                   chunk from the second pair of two-level sums; bytes are written straight to
 *                   their final position header_len + byte index + zeros before, each 0xFF
 *                   followed by 0x00 (T.81 F.1.2.3).  Chunk 0 writes the JFIF header, the last
 *                   chunk completes the final byte with 1-bits, writes EOI and the length word.
 * This is BASELINE.json's "second length/offset scan" for the 0xFF00 stuffing; the reference
 * emits no byte stream at all, the oracle for these bytes is
 * oracle/jpgb200_oracle.c (bw_byte), cross-checked by PIL/OpenCV decodes in tests/.
 *
 * MCU-row bands (jpgb200_band_*): bit_base is the band's bit phase (0..7), `carry` holds the
 * previous band's trailing bits, and a band that is not the last leaves its trailing partial byte
 * to the next band.
 */
#ifndef JPGB200_TAIL_CUH
#define JPGB200_TAIL_CUH

#include "entropy.cuh"

namespace jpgb200 {

#ifndef TAIL_THREADS
#define TAIL_THREADS 256
#endif
#define TAIL_WARPS (TAIL_THREADS / 32)

/* what the band's own kernels read instead of host-side launch parameters */
struct MgState {
    int pred[3];              /* DC predictor of the band's first blocks */
    unsigned int phase;       /* bit phase of the band's first bit in the scan, 0..7 */
    unsigned int carry;       /* the previous band's trailing bits (low `phase` bits) */
    unsigned int pad;
    unsigned long long bits;     /* band bits with the true predictor */
    unsigned long long bit_base; /* bit offset of the band in the scan */
    unsigned long long out_off;  /* byte offset of the band's segment in the final stream */
    unsigned long long seg_len;  /* bytes of the band's segment (stuffed scan bytes, + EOI on the last band) */
    unsigned long long total_len; /* whole stream: header + scan + EOI */
};

struct TailParams {
    const uint32_t *tmp; /* [n_chunks][slot_words] chunk streams, big-endian words (KE) */
    long long slot_words, n_chunks;
    const uint32_t *cta_bits, *super_bits, *cta_tail;
    uint32_t *cta_ff, *super_ff;
    unsigned long long *cta_base; /* [n_chunks] global bit offset of every chunk: written by KF, read by KS */
    unsigned int bit_phase; /* global bit offset of chunk 0, 0..7 */
    unsigned int carry;     /* its low bit_phase bits precede the band's first bit */
    int is_last_band;
    /* KS only */
    const uint8_t *header;
    int header_len;
    uint8_t *out;
    unsigned long long cap;
    unsigned long long *len_out;
    int write_eoi;
    EncState *st; /* optional */
    /* batches of small frames in one launch: blockIdx.y = frame */
    long long tmp_fstride, out_fstride; /* words, bytes */
    int cta_fstride, super_fstride;     /* elements */
    /* multi-GPU bands: bit phase, carry and output offset come from device memory (mgpu.cuh) */
    const MgState *mg;
    unsigned long long out_off;
    int mg_local; /* 1: KS writes into a local segment buffer at offset (out_off & 15); k_mg_push moves it to rank 0 */
};

__device__ __forceinline__ TailParams tail_frame(const TailParams &q0)
{
    TailParams q = q0;
    const long long f = blockIdx.y;
    q.tmp += f * q.tmp_fstride;
    q.cta_bits += f * q.cta_fstride;
    q.cta_tail += f * q.cta_fstride;
    q.cta_ff += f * q.cta_fstride;
    q.cta_base += f * q.cta_fstride;
    q.super_bits += f * q.super_fstride;
    q.super_ff += f * q.super_fstride;
    if (q.out) q.out += f * q.out_fstride;
    if (q.len_out) q.len_out += f;
    if (q.mg) {
        q.bit_phase = q.mg->phase;
        q.carry = q.mg->carry;
        q.out_off = q.mg_local ? (q.mg->out_off & 15ull) : q.mg->out_off;
    }
    return q;
}

__device__ __forceinline__ uint32_t be32(uint32_t w) { return __byte_perm(w, 0, 0x0123); }

/* sum over the warp of the two-level sums before `chunk` */
__device__ __forceinline__ unsigned long long tail_prefix(const uint32_t *cta, const uint32_t *sup, long long chunk, int lane)
{
    unsigned long long part = 0ull;
    const int c = (int)chunk, s = c >> EC_SUPER_SHIFT; /* chunk counts stay far below 2^31 */
    for (int i = lane; i < s; i += 32) part += sup[i];
    for (int i = (s << EC_SUPER_SHIFT) + lane; i < c; i += 32) part += cta[i];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
    return part;
}

/* geometry of one chunk on the byte grid */
struct TailGeom {
    uint32_t T, nw;             /* bits and words of the chunk stream */
    int d;                      /* local bit offset of the first owned byte: -(g0 & 7) */
    unsigned long long bA;      /* index of the first owned byte in the un-stuffed stream */
    uint32_t nB;                /* owned bytes */
    uint32_t pad;               /* 1-bits completing the final byte (last chunk of the last band) */
    __device__ __forceinline__ TailGeom(const TailParams &q, long long chunk, unsigned long long g0, uint32_t T_)
    {
        T = T_;
        nw = (T + 31u) >> 5;
        d = -(int)(g0 & 7ull);
        bA = g0 >> 3;
        const unsigned long long end = g0 + T;
        unsigned long long bB = end >> 3;
        pad = 0;
        if (chunk == q.n_chunks - 1 && q.is_last_band && (end & 7ull)) {
            pad = 8u - (uint32_t)(end & 7ull);
            bB += 1;
        }
        nB = (uint32_t)(bB - bA);
    }
    /* Owned bytes 4i..4i+3 = the 32 stream bits at local offset d + 32 i, assembled from the raw
     * (big-endian) words i-1 and i of the slot -- word -1 being the 32 bits in front of the chunk.
     * Bits at or beyond T read 0, the pad 1.  The words can be fetched before g0 is known: only the
     * shift depends on it. */
    __device__ __forceinline__ uint32_t group(uint32_t i, uint32_t raw_prev, uint32_t raw_cur) const
    {
        const uint32_t wc = i < nw ? be32(raw_cur) : 0u;
        uint32_t v = wc;
        if (d) {
            const uint32_t wp = (i == 0u) ? raw_prev : ((i - 1u) < nw ? be32(raw_prev) : 0u);
            const uint32_t sh = 32u + (uint32_t)d;
            v = (wp << sh) | (wc >> (32u - sh));
        }
        if (pad) {
            const int L = d + 32 * (int)i;
            const int lo_j = (int)T - L, hi_j = (int)(T + pad) - L; /* window bits [lo_j, hi_j) are pad */
            if (hi_j > 0 && lo_j < 32) {
                const uint32_t m_lo = lo_j <= 0 ? 0xffffffffu : (0xffffffffu >> lo_j);
                const uint32_t m_hi = hi_j >= 32 ? 0u : (0xffffffffu >> hi_j);
                v |= m_lo & ~m_hi;
            }
        }
        return v;
    }
};

/* Batched fetch: TAIL_BATCH rounds of 32 consecutive raw words, all requested before any is used.
 * Word i-1 of group i is the neighbouring lane's word (shuffle); for lane 0 it is the last word
 * of the previous round, or, in front of the chunk, the previous chunk's last 32 bits (numeric). */
#define TAIL_BATCH 4
__device__ __forceinline__ void tail_fetch(const TailParams &q, const uint32_t *slot, uint32_t base, int lane,
                                           uint32_t (&rc)[TAIL_BATCH])
{
#pragma unroll
    for (int r = 0; r < TAIL_BATCH; r++) {
        const uint32_t i = base + 32u * (uint32_t)r + (uint32_t)lane;
        rc[r] = (long long)i < q.slot_words ? slot[i] : 0u;
    }
}
/* raw_prev of this lane for round r; `carry_word` = word (base + 32 r - 1), maintained by the caller */
__device__ __forceinline__ uint32_t tail_prev(uint32_t cur, uint32_t carry_word, int lane)
{
    const uint32_t up = __shfl_up_sync(0xffffffffu, cur, 1);
    return lane == 0 ? carry_word : up;
}

/* 0xff per byte of v that is 0xFF, restricted to its first (most significant) nvalid bytes */
__device__ __forceinline__ uint32_t ff_mask(uint32_t v, int nvalid)
{
    uint32_t m = __vcmpeq4(v, 0xffffffffu);
    if (nvalid < 4) m &= nvalid <= 0 ? 0u : (0xffffffffu << (8 * (4 - nvalid)));
    return m;
}

/* number of 0xFF bytes in v, exactly: a byte of ~v is zero iff bit 7 of ((~v & 0x7f..) + 0x7f..) | ~v is clear */
__device__ __forceinline__ uint32_t ff_count4(uint32_t v)
{
    const uint32_t t = ~v;
    const uint32_t y = (t & 0x7f7f7f7fu) + 0x7f7f7f7fu;
    return (uint32_t)__popc(~(y | t) & 0x80808080u);
}

/* ---- KF ---- */
#ifndef KF_WIDE
#define KF_WIDE 1 /* 0: every round through the general path (32 groups of 4 bytes per round) */
#endif
__global__ void __launch_bounds__(TAIL_THREADS) k_count_ff(const TailParams q0)
{
    const TailParams q = tail_frame(q0);
    const int lane = threadIdx.x & 31;
    const long long chunk = (long long)blockIdx.x * TAIL_WARPS + (threadIdx.x >> 5);
    if (chunk >= q.n_chunks) return;
    const uint32_t *slot = q.tmp + chunk * q.slot_words;
    /* everything this warp needs is requested before anything is waited for */
#if KF_WIDE
    /* the first WIDE round's words: 16 bytes per lane (slots are 16-byte aligned and never shorter than 128 words) */
    uint4 w4 = reinterpret_cast<const uint4 *>(slot)[lane];
#else
    uint32_t rc[TAIL_BATCH];
    tail_fetch(q, slot, 0u, lane, rc);
#endif
    uint32_t carry_word = chunk == 0 ? q.carry : q.cta_tail[chunk - 1];
    const uint32_t T = q.cta_bits[chunk];
    const unsigned long long g0 = (unsigned long long)q.bit_phase + tail_prefix(q.cta_bits, q.super_bits, chunk, lane);
    const TailGeom g(q, chunk, g0, T);
    const uint32_t groups = (g.nB + 3u) >> 2;
    uint32_t c = 0;
    uint32_t base = 0;
#if KF_WIDE
    /* WIDE rounds: 128 groups (512 owned bytes) per round, four consecutive groups per lane from one 16-byte load.
     * A round qualifies when all its bytes are owned and the chunk's final byte (the only one that can hold pad bits)
     * lies behind it: then every group is the plain funnel of two stream words -- W(i-1), W(i) in numeric form, W(-1)
     * being the 32 bits in front of the chunk -- and bounds, validity masks and pad handling fall away. */
    {
        const uint32_t nd = (uint32_t)(-g.d); /* 0..7 */
        uint32_t carry_num = carry_word;      /* numeric word in front of the round */
        while ((base + 128u) * 4u < g.nB) {
            const uint32_t W0 = be32(w4.x), W1 = be32(w4.y), W2 = be32(w4.z), W3 = be32(w4.w);
            const uint32_t up = __shfl_up_sync(0xffffffffu, W3, 1);
            const uint32_t Wm = lane == 0 ? carry_num : up;
            carry_num = __shfl_sync(0xffffffffu, W3, 31);
            base += 128u;
            if ((base + 128u) * 4u < g.nB) w4 = reinterpret_cast<const uint4 *>(slot)[(base >> 2) + lane]; /* the next round's */
            c += ff_count4(__funnelshift_r(W0, Wm, nd)) + ff_count4(__funnelshift_r(W1, W0, nd)) +
                 ff_count4(__funnelshift_r(W2, W1, nd)) + ff_count4(__funnelshift_r(W3, W2, nd));
        }
        if (base) carry_word = be32(carry_num); /* the general path takes the word in front of a later round raw */
    }
    uint32_t rc[TAIL_BATCH];
    if (base < groups) tail_fetch(q, slot, base, lane, rc);
    const uint32_t base0 = base;
#else
    const uint32_t base0 = 0u;
#endif
    for (; base < groups; base += 32u * TAIL_BATCH) {
        if (base != base0) tail_fetch(q, slot, base, lane, rc);
#pragma unroll
        for (int r = 0; r < TAIL_BATCH; r++) {
            const uint32_t i = base + 32u * (uint32_t)r + (uint32_t)lane;
            if (base + 32u * (uint32_t)r >= groups) break; /* warp-uniform */
            const uint32_t rp = tail_prev(rc[r], carry_word, lane);
            carry_word = __shfl_sync(0xffffffffu, rc[r], 31);
            if (i < groups) {
                const uint32_t v = g.group(i, rp, rc[r]);
                const int left = (int)(g.nB - 4u * i);
                c += (uint32_t)__popc(ff_mask(v, left >= 4 ? 4 : left)) >> 3;
            }
        }
    }
#pragma unroll
    for (int dd = 16; dd > 0; dd >>= 1) c += __shfl_xor_sync(0xffffffffu, c, dd);
    if (lane == 0) {
        q.cta_ff[chunk] = c;
        q.cta_base[chunk] = g0;
        if (c) atomicAdd(&q.super_ff[chunk >> EC_SUPER_SHIFT], c);
    }
}

/* ---- KS ---- */
#ifndef KS_INTERIOR
#define KS_INTERIOR 1 /* 0: every round through the general geometry (bounds, validity masks, pad) */
#endif
__global__ void __launch_bounds__(TAIL_THREADS) k_scatter(const TailParams q0)
{
    const TailParams q = tail_frame(q0);
    const int lane = threadIdx.x & 31;
    const long long chunk = (long long)blockIdx.x * TAIL_WARPS + (threadIdx.x >> 5);
    if (chunk >= q.n_chunks) return;
    const uint32_t *slot = q.tmp + chunk * q.slot_words;
    uint32_t rc[TAIL_BATCH];
    tail_fetch(q, slot, 0u, lane, rc);
    uint32_t carry_word = chunk == 0 ? q.carry : q.cta_tail[chunk - 1];
    const uint32_t T = q.cta_bits[chunk];
    const unsigned long long g0 = q.cta_base[chunk];
    const uint32_t my_ff = q.cta_ff[chunk];
    const unsigned long long ff_before = tail_prefix(q.cta_ff, q.super_ff, chunk, lane);
    const TailGeom g(q, chunk, g0, T);
    const unsigned long long trailer = q.write_eoi ? 2ull : 0ull;
    unsigned long long P = q.out_off + (unsigned long long)q.header_len + g.bA + ff_before;
    const bool fits = P + g.nB + my_ff + trailer <= q.cap;
    const bool last = (chunk == q.n_chunks - 1);
    if (chunk == 0 && (unsigned long long)q.header_len <= q.cap)
        for (int i = lane; i < q.header_len; i += 32) q.out[i] = q.header[i];
    if (fits) {
        const uint32_t groups = (g.nB + 3u) >> 2;
        for (uint32_t base = 0; base < groups; base += 32u * TAIL_BATCH) {
            if (base) tail_fetch(q, slot, base, lane, rc);
#pragma unroll
            for (int r = 0; r < TAIL_BATCH; r++) {
                const uint32_t i = base + 32u * (uint32_t)r + (uint32_t)lane;
                if (base + 32u * (uint32_t)r >= groups) break; /* warp-uniform */
                const uint32_t rp = tail_prev(rc[r], carry_word, lane);
                carry_word = __shfl_sync(0xffffffffu, rc[r], 31);
                const uint32_t round0 = base + 32u * (uint32_t)r;
#if KS_INTERIOR
                if ((round0 + 32u) * 4u < g.nB) {
                    /* warp-uniform: 128 owned bytes and the chunk's final byte (the only one that can hold pad bits)
                     * behind them -- every group is the plain funnel of two stream words, all four bytes count */
                    const uint32_t wp = i == 0u ? rp : be32(rp);
                    const uint32_t v = __funnelshift_r(be32(rc[r]), wp, (uint32_t)(-g.d));
                    const uint32_t m = __vcmpeq4(v, 0xffffffffu);
                    unsigned busy = __ballot_sync(0xffffffffu, m != 0u);
                    if (busy == 0u) {
                        /* no 0xFF in these 128 bytes (40 % of the rounds): they go out as they are, with aligned 32-bit
                         * stores -- words straddling two lanes are assembled by shuffle */
                        uint8_t *d = q.out + P;
                        const unsigned a = (unsigned)((unsigned long long)(uintptr_t)d & 3ull);
                        const uint32_t own = __byte_perm(v, 0, 0x0123); /* memory order */
                        const uint32_t prv = __shfl_up_sync(0xffffffffu, own, 1);
                        if (a == 0u) {
                            reinterpret_cast<uint32_t *>(d)[lane] = own;
                        } else {
                            /* a = 1..3: lane 0 writes the 4 - a bytes in front of the first aligned word, lane 31 the a
                             * bytes behind the last one -- predicated single stores, no loops with a lane-dependent trip count */
                            if (lane > 0) {
                                reinterpret_cast<uint32_t *>(d - a)[lane] = __funnelshift_r(prv, own, 8u * (4u - a));
                            } else {
#pragma unroll
                                for (unsigned t = 0; t < 3u; t++)
                                    if (t < 4u - a) d[t] = (uint8_t)(own >> (8u * t));
                            }
                            if (lane == 31) {
#pragma unroll
                                for (unsigned t = 0; t < 3u; t++)
                                    if (t < a) d[128u - a + t] = (uint8_t)(own >> (8u * (4u - a + t)));
                            }
                        }
                        P += 128ull;
                        continue;
                    }
                    /* some lane holds a 0xFF (typically one or two): every lane writes its four bytes at 4 * lane plus the
                     * zeros stuffed by the lanes below it -- summed over the FEW busy lanes, not scanned over all 32 */
                    const uint32_t e = (uint32_t)__popc(m) >> 3;
                    uint32_t below = 0u, extra = 0u;
                    do {
                        const int j = __ffs((int)busy) - 1;
                        busy &= busy - 1u;
                        const uint32_t ej = __shfl_sync(0xffffffffu, e, j);
                        below += lane > j ? ej : 0u;
                        extra += ej;
                    } while (busy);
                    uint8_t *dst = q.out + P + 4u * (uint32_t)lane + below;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const uint32_t byte = (v >> (24 - 8 * k)) & 0xffu;
                        *dst++ = (uint8_t)byte;
                        if (byte == 0xffu) *dst++ = 0;
                    }
                    P += 128ull + extra;
                    continue;
                }
#endif
                /* the chunk's last round(s): bounds, validity masks, pad */
                uint32_t v = 0;
                int nvalid = 0;
                if (i < groups) {
                    v = g.group(i, rp, rc[r]);
                    const int left = (int)(g.nB - 4u * i);
                    nvalid = left >= 4 ? 4 : left;
                }
                const uint32_t m = ff_mask(v, nvalid);
                const uint32_t n_out = (uint32_t)nvalid + ((uint32_t)__popc(m) >> 3);
                uint32_t inc = n_out;
#pragma unroll
                for (int dd = 1; dd < 32; dd <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, dd);
                    if (lane >= dd) inc += t;
                }
                const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
                uint8_t *dst = q.out + P + (inc - n_out);
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (k < nvalid) {
                        const uint32_t byte = (v >> (24 - 8 * k)) & 0xffu;
                        *dst++ = (uint8_t)byte;
                        if (byte == 0xffu) *dst++ = 0;
                    }
                }
                P += total;
            }
        }
    }
    if (last && lane == 0) {
        const unsigned long long nbytes = g.bA + g.nB;
        const unsigned long long total_ff = ff_before + my_ff;
        const unsigned long long out_len = q.out_off + (unsigned long long)q.header_len + nbytes + total_ff + trailer;
        const bool ok = out_len <= q.cap;
        if (ok && q.write_eoi) {
            q.out[out_len - 2] = 0xFF;
            q.out[out_len - 1] = 0xD9; /* EOI */
        }
        if (EncState *st = q.st) {
            st->nbytes = nbytes;
            st->total_ff = total_ff;
            st->out_len = out_len;
            st->error = ok ? 0 : 1;
        }
        if (q.len_out) *q.len_out = ok ? out_len : 0ull;
    }
}

/* trailing partial byte of a band: the last r bits of its stream, left-aligned in a byte */
__global__ void k_band_tail(const uint32_t *tmp, long long slot_words, long long n_chunks, const uint32_t *cta_bits,
                            unsigned int r, unsigned long long *out_byte)
{
    const uint32_t T = cta_bits[n_chunks - 1];
    const uint32_t *slot = tmp + (n_chunks - 1) * slot_words;
    uint32_t byte = 0;
    if (r && T >= r) {
        const uint32_t L = T - r, w = L >> 5, sh = L & 31u, nw = (T + 31u) >> 5;
        const uint32_t hi = be32(slot[w]);
        const uint32_t lo = (w + 1 < nw) ? be32(slot[w + 1]) : 0u;
        const uint32_t v = sh ? ((hi << sh) | (lo >> (32u - sh))) : hi;
        byte = (v >> 24) & (0xffu << (8u - r)) & 0xffu;
    }
    *out_byte = byte;
}

}  // namespace jpgb200
#endif
