/*
 * K35 -- the whole entropy stage of one image in ONE launch: per-block bit lengths, a
 * device-wide exclusive prefix sum of bit offsets, the parallel bit pack, and 0xFF -> 0xFF00
 * byte stuffing placed by a second length/offset scan (BASELINE.json north_star's three
 * stages), fused so that the coefficients are read from HBM once and the un-stuffed stream
 * never leaves shared memory.
 *
 * One CTA = one chunk of EC_MCUS consecutive MCUs (taken in order through an atomic ticket):
 *   1. asynchronous fill of the chunk's coefficient rows (shared memory)
 *   2. count pass: bits per block (ec_count_block) -> CTA scan -> chunk total T
 *   3. chained scan #1 over chunks ("decoupled look-back", 64-bit state words tagged with the
 *      encode's epoch): exclusive bit offset g0 of the chunk
 *   4. pack pass: every block writes its codes into a shared-memory staging buffer at the
 *      stream's true bit phase.  A chunk OWNS the bytes whose last bit lies in its bit range;
 *      the leading bits of its first byte belong to the previous chunk's last block, which one
 *      lane of the lightest warp re-encodes (no waiting on the neighbour, no atomics in HBM)
 *   5. count 0xFF among the owned bytes -> chained scan #2: bytes inserted before this chunk
 *   6. scatter with the stuffed zeros into a second shared buffer held at the destination's
 *      16-byte phase, then 16-byte coalesced stores straight into the JFIF stream.
 * Chunk 0 also writes the marker segments, the last chunk the EOI and the length word.
 * A chunk whose stream exceeds the staging buffer is processed in batches of whole blocks.
 */
#ifndef JPGB200_ENTROPY_FUSED_CUH
#define JPGB200_ENTROPY_FUSED_CUH

#include "entropy.cuh"

namespace jpgb200 {

#define EF_CAP_BYTES 4096 /* un-stuffed bytes staged per batch */
#define EF_CAP_BITS (EF_CAP_BYTES * 8 - 64)

struct EfParams {
    const int16_t *zz; /* [Y blocks | Cb blocks | Cr blocks][64] */
    long long coff1, coff2;
    long long n_mcu, n_chunks;
    int dc_pred[3];
    int header_len;
    EncState *st;
    unsigned long long *bit_state, *ff_state; /* [n_chunks] chained-scan states */
    unsigned long long epoch;
    const uint8_t *header;
    uint8_t *out;
    unsigned long long cap;
    unsigned long long *len_out;
    unsigned long long *dbg_block_off; /* optional: global bit offset of every block (+ total) */
};

/* keeps the most recent bits: used to recover the tail of the previous chunk's last block */
struct TailSink {
    unsigned long long acc;
    __device__ __forceinline__ void put(uint32_t v, int n) { acc = (acc << n) | (unsigned long long)v; }
};

/* number of 0xFF bytes among the first `nvalid` (0..16) bytes of v */
__device__ __forceinline__ int ef_count_ff16(const uint4 v, int nvalid)
{
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    int c = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint32_t m = __vcmpeq4(w[i], 0xffffffffu);
        const int nv = nvalid - 4 * i;
        if (nv < 4) m &= (nv <= 0) ? 0u : ((1u << (8 * nv)) - 1u);
        c += __popc(m) >> 3;
    }
    return c;
}

/* dynamic shared memory layout (all offsets multiples of 16) */
template <int RATIO>
struct EfSmem {
    static const int NT = EcCfg<RATIO>::NT;
    static const int OFF_B = (NT + 2) * EC_ROW_CHUNKS * 16;
    static const int OFF_H = OFF_B + (int)sizeof(HuffBits);
    static const int OFF_S = OFF_H + (int)sizeof(HuffDev);
    static const int OFF_O = OFF_S + EF_CAP_BYTES + 32;
    static const int OFF_OFF = OFF_O + 2 * EF_CAP_BYTES + 48;
    static const int BYTES = OFF_OFF + (NT + 1) * 4 + 12;
};

#ifdef JPGB200_AB /* a cross-check: built only with -DJPGB200_AB (build.py: JPGB200_AB=1) */
template <int RATIO>
__global__ void __launch_bounds__(EcCfg<RATIO>::NT) k35_entropy(const EfParams q, const HuffBits *hbits, const HuffDev *huff)
{
    constexpr int BPM = EcCfg<RATIO>::BPM, NT = EcCfg<RATIO>::NT, NW = NT / 32;
    extern __shared__ __align__(16) unsigned char ef_smem[];
    typedef uint4 Row[EC_ROW_CHUNKS];
    Row *sC = reinterpret_cast<Row *>(ef_smem); /* NT+2 rows; rows NT, NT+1: Cb and Cr block of the previous chunk's last MCU */
    unsigned char *sp = ef_smem + EfSmem<RATIO>::OFF_B;
    HuffBits &sB = *reinterpret_cast<HuffBits *>(sp);
    HuffDev &sH = *reinterpret_cast<HuffDev *>(ef_smem + EfSmem<RATIO>::OFF_H);
    uint32_t *sS = reinterpret_cast<uint32_t *>(ef_smem + EfSmem<RATIO>::OFF_S);
    uint8_t *sO = ef_smem + EfSmem<RATIO>::OFF_O;
    uint32_t *sOff = reinterpret_cast<uint32_t *>(ef_smem + EfSmem<RATIO>::OFF_OFF);
    __shared__ uint32_t sWarp[NW];
    __shared__ unsigned long long sTicket, sG0, sFF;
    __shared__ int sHi;
    __shared__ uint32_t sCarry;

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) sTicket = atomicAdd(&q.st->ticket35, 1ull);
    __syncthreads();
    const long long chunk = (long long)sTicket;
    const long long m0 = chunk * EC_MCUS;

    /* EcParams view for the shared helpers */
    EcParams p;
    p.zz = q.zz;
    p.coff1 = q.coff1;
    p.coff2 = q.coff2;
    p.n_mcu = q.n_mcu;
    p.dc_pred[0] = q.dc_pred[0];
    p.dc_pred[1] = q.dc_pred[1];
    p.dc_pred[2] = q.dc_pred[2];

    /* ---- 1. asynchronous fill ---- */
    ec_fill<RATIO>(reinterpret_cast<uint4(*)[EC_ROW_CHUNKS]>(sC), p, m0, tid);
    if (m0 > 0 && tid < 16) {
        const long long blk = ((tid < 8) ? q.coff1 : q.coff2) + m0 - 1;
        cp_async16(&sC[NT + (tid >> 3)][tid & 7], reinterpret_cast<const uint4 *>(q.zz) + blk * 8 + (tid & 7));
    }
    ec_load_table<EcCfg<RATIO>::NT>(&sB, hbits, tid);
    ec_load_table<EcCfg<RATIO>::NT>(&sH, huff, tid);
    const EcMap<RATIO> mp(p, m0, tid);
    /* DC predictors of the previous chunk's last Cb and Cr block (m0 is a multiple of EC_MCUS,
     * so block m0-2 exists whenever m0 > 0) */
    int tail_pred_cb = 0, tail_pred_cr = 0;
    if (tid == NT - 1 && m0 > 0) {
        tail_pred_cb = (int)q.zz[(q.coff1 + m0 - 2) * 64];
        tail_pred_cr = (int)q.zz[(q.coff2 + m0 - 2) * 64];
    }
    cp_async_wait_all();
    __syncthreads();

    /* ---- 2. count pass ---- */
    uint32_t nbits = 0;
    int diff = 0;
    if (mp.valid) {
        const int pred = ec_dc_pred<RATIO>(reinterpret_cast<const uint4(*)[EC_ROW_CHUNKS]>(sC), p, m0, tid, mp);
        diff = (int)(short)(sC[tid][0].x & 0xffffu) - pred;
        nbits = ec_count_block(sC[tid], diff, sB.ac[mp.comp ? 1 : 0], sB.dc[mp.comp ? 1 : 0]);
    }
    sOff[mp.s_local] = nbits; /* scan (MCU-interleaved) order */
    __syncthreads();
    uint32_t total;
    {
        const uint32_t v = sOff[tid];
        const uint32_t ex = cta_exclusive_scan<NT>(v, sWarp, tid, total);
        __syncthreads();
        sOff[tid] = ex;
        if (tid == 0) sOff[NT] = total;
    }

    /* ---- 3. chained scan #1: global bit offset of the chunk ---- */
    if (wid == NW - 1) {
        const unsigned long long g = cs_lookback<16>(q.bit_state, (unsigned long long)chunk, q.epoch, total, lane, true, true);
        if (lane == 0) sG0 = g;
    }
    __syncthreads();
    const unsigned long long g0 = sG0;
    long long left = q.n_mcu - m0;
    const int nblk = (int)((left < EC_MCUS ? left : EC_MCUS) * BPM); /* valid blocks, contiguous in scan order */
    const bool last_chunk = (chunk == q.n_chunks - 1);
    if (q.dbg_block_off) {
        if (tid < nblk) q.dbg_block_off[m0 * BPM + tid] = g0 + sOff[tid];
        if (last_chunk && tid == 0) q.dbg_block_off[q.n_mcu * BPM] = g0 + total;
    }

    /* ---- 4-6. batches of whole blocks: pack, count 0xFF, place, stuff, store ---- */
    const int my_s = mp.s_local;
    const uint32_t my_off = sOff[my_s];
    int b_lo = 0;
    unsigned long long ff_run = 0; /* bytes inserted before the current batch (valid after the first look-back) */
    bool have_ff_excl = false, published = false;
    unsigned long long ff_own = 0;
    while (b_lo < nblk) {
        const uint32_t base_bits = sOff[b_lo];
        const unsigned long long gb = g0 + base_bits; /* global bit offset of the batch */
        const int phase = (int)(gb & 7ull);
        /* largest b_hi with phase + (sOff[b_hi] - base) <= EF_CAP_BITS */
        if (tid == 0) sHi = nblk;
        __syncthreads();
        if (tid >= b_lo && tid < nblk) {
            const bool fit_here = phase + (sOff[tid + 1] - base_bits) <= (uint32_t)EF_CAP_BITS;
            const bool fit_prev = (tid == b_lo) || (phase + (sOff[tid] - base_bits) <= (uint32_t)EF_CAP_BITS);
            if (!fit_here && fit_prev) sHi = tid; /* first block that no longer fits */
        }
        const uint32_t nzero = (uint32_t)(EF_CAP_BYTES / 4 + 8);
        for (uint32_t i = tid; i < nzero; i += NT) sS[i] = 0u;
        __syncthreads();
        int b_hi = sHi;
        if (b_hi == b_lo) b_hi = b_lo + 1; /* a single block always fits (<= 1658 bits) */
        const uint32_t tb = sOff[b_hi] - base_bits;
        const bool very_last_batch = last_chunk && (b_hi == nblk);

        if (mp.valid && my_s >= b_lo && my_s < b_hi) {
            const uint32_t loff = (uint32_t)phase + (my_off - base_bits);
            PackSink<true> s;
            s.acc = 0;
            s.nacc = (int)(loff & 31u);
            s.wp = sS + (loff >> 5);
            s.first = true;
            ec_pack_block(sC[tid], diff, sH.ac[mp.comp ? 1 : 0], sH.dc[mp.comp ? 1 : 0], s);
            if (very_last_batch && my_s == nblk - 1) {
                const int rem = s.nacc & 7; /* complete the final byte with 1-bits (T.81 F.1.2.3) */
                if (rem) s.put((1u << (8 - rem)) - 1u, 8 - rem);
            }
            s.finish();
        }
        if (b_lo == 0 && tid == NT - 1) {
            /* Tail of the previous chunk's stream = leading bits of our first owned byte.  Its
             * last MCU's Cb and Cr blocks carry at least 8 bits; re-encoding them here (a lane
             * of the lightest warp, overlapped with the luma warps' packing) avoids any wait on
             * the neighbouring chunk. */
            uint32_t carry = 0u;
            if (m0 > 0 && phase) {
                TailSink ts;
                ts.acc = 0ull;
                ec_pack_block(sC[NT], (int)(short)(sC[NT][0].x & 0xffffu) - tail_pred_cb, sH.ac[1], sH.dc[1], ts);
                ec_pack_block(sC[NT + 1], (int)(short)(sC[NT + 1][0].x & 0xffffu) - tail_pred_cr, sH.ac[1], sH.dc[1], ts);
                carry = (uint32_t)((ts.acc & ((1ull << phase) - 1ull)) << (8 - phase));
            }
            sCarry = carry;
        }
        __syncthreads();
        if (tid == 0 && phase) sS[0] |= sCarry; /* leading bits of the first owned byte (little-endian word, byte 0) */
        __syncthreads();
        /* owned bytes of this batch: those whose last bit lies inside it */
        int n_own = (int)(((uint32_t)phase + tb) >> 3);
        if (very_last_batch && (((uint32_t)phase + tb) & 7u)) n_own += 1; /* the padded final byte */
        /* carry for the next batch: its first byte's leading bits */
        const uint32_t next_carry = reinterpret_cast<const uint8_t *>(sS)[((uint32_t)phase + tb) >> 3];

        /* 0xFF count, 32 contiguous bytes per thread */
        uint4 v[2];
        int nv[2], c = 0;
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int bk = tid * 32 + 16 * k;
            nv[k] = n_own - bk;
            nv[k] = nv[k] < 0 ? 0 : (nv[k] > 16 ? 16 : nv[k]);
            v[k] = make_uint4(0, 0, 0, 0);
            if (nv[k] > 0) v[k] = reinterpret_cast<const uint4 *>(sS)[tid * 2 + k];
            c += ef_count_ff16(v[k], nv[k]);
        }
        uint32_t batch_ff;
        const uint32_t ff_before = cta_exclusive_scan<NT>((uint32_t)c, sWarp, tid, batch_ff);
        const bool single = (b_lo == 0 && b_hi == nblk);
        if (!have_ff_excl) {
            /* chained scan #2.  A single-batch chunk announces its count at once; a multi-batch
             * chunk only learns its total at the end and publishes the prefix then. */
            if (wid == NW - 1) {
                const unsigned long long e =
                    cs_lookback<16>(q.ff_state, (unsigned long long)chunk, q.epoch, batch_ff, lane, single, single);
                if (lane == 0) sFF = e;
            }
            __syncthreads();
            ff_run = sFF;
            have_ff_excl = true;
            published = single;
        }
        ff_own += batch_ff;

        /* destination of the batch's first owned byte */
        const unsigned long long dst = (unsigned long long)q.header_len + (gb >> 3) + ff_run;
        const int n_out = n_own + (int)batch_ff;
        const bool fits = dst + (unsigned long long)n_out + 2ull <= q.cap;
        if (fits) {
            const int dphase = (int)(((unsigned long long)(uintptr_t)q.out + dst) & 15ull);
            int pos = dphase + tid * 32 + (int)ff_before;
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const uint32_t w[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    if (i < nv[k]) {
                        const uint8_t byte = (uint8_t)(w[i >> 2] >> (8 * (i & 3)));
                        sO[pos++] = byte;
                        if (byte == 0xFF) sO[pos++] = 0x00;
                    }
                }
            }
            __syncthreads();
            const int end = dphase + n_out;
            uint8_t *gbase = q.out + dst - dphase; /* 16-byte aligned image of sO[0] */
            const int first_full = (dphase + 15) >> 4, last_full = end >> 4;
            for (int l = first_full + tid; l < last_full; l += NT)
                reinterpret_cast<uint4 *>(gbase)[l] = reinterpret_cast<const uint4 *>(sO)[l];
            if (tid < 16) { /* ragged head */
                const int head_end = (first_full << 4) < end ? (first_full << 4) : end;
                const int i = dphase + tid;
                if (i < head_end) gbase[i] = sO[i];
            } else if (tid < 32) { /* ragged tail */
                const int tail_start = (last_full << 4) > (first_full << 4) ? (last_full << 4) : (first_full << 4);
                const int i = tail_start + (tid - 16);
                if (i >= dphase && i < end && last_full >= first_full) gbase[i] = sO[i];
            }
        }
        if (tid == 0) sCarry = next_carry;
        ff_run += batch_ff;
        b_lo = b_hi;
        __syncthreads();
    }
    /* a multi-batch chunk publishes its inclusive 0xFF prefix now */
    if (tid == 0 && !published) q.ff_state[chunk] = cs_pack(CS_FLAG_P, q.epoch, ff_run);

    if (chunk == 0)
        for (int i = tid; i < q.header_len; i += NT) q.out[i] = q.header[i];
    if (last_chunk && tid == 0) {
        const unsigned long long total_bits = g0 + total;
        const unsigned long long nbytes = (total_bits + 7ull) >> 3;
        const unsigned long long out_len = (unsigned long long)q.header_len + nbytes + ff_run + 2ull;
        const bool ok = out_len <= q.cap;
        if (ok) {
            q.out[out_len - 2] = 0xFF;
            q.out[out_len - 1] = 0xD9; /* EOI */
        }
        q.st->total_bits = total_bits;
        q.st->nbytes = nbytes;
        q.st->total_ff = ff_run;
        q.st->out_len = out_len;
        q.st->error = ok ? 0 : 1;
        q.st->ticket35 = 0ull; /* every ticket of this launch has been taken */
        if (q.len_out) *q.len_out = ok ? out_len : 0ull;
    }
}
#endif /* JPGB200_AB */

}  // namespace jpgb200
#endif
