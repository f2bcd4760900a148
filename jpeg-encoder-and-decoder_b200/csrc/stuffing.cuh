/*
 * K6 -- 0xFF -> 0xFF00 byte stuffing (T.81 F.1.2.3) by a second length/offset scan:
 *   k6_stuff: one persistent kernel.  Per 4 KB tile of the un-stuffed stream: count 0xFF
 *             bytes, obtain the number of 0xFF bytes in all earlier tiles by a single-pass
 *             chained scan (decoupled look-back), then write the tile at (tile offset +
 *             #0xFF before it) inserting the zero bytes; output is staged in shared memory
 *             at the same 16-byte phase as the destination so the copy-out is 16-byte
 *             vectorised and coalesced.  Tile 0 also writes the JFIF header, the last tile
 *             the EOI marker and the final length.
 * The reference emits no byte stream; the oracle for these bytes is
 * oracle/jpgb200_oracle.c (bw_byte), checked by PIL/OpenCV decodes in tests/.
 */
#ifndef JPGB200_STUFFING_CUH
#define JPGB200_STUFFING_CUH

#include "entropy.cuh"

namespace jpgb200 {

#define K6_THREADS 256
#define K6_ITEMS 4
#define K6_TILE (K6_THREADS * 16 * K6_ITEMS) /* 16 KB of un-stuffed stream per tile: 64 contiguous bytes per thread */

/* number of 0xFF bytes among the first `nvalid` (0..16) bytes of v */
__device__ __forceinline__ int count_ff16(const uint4 v, int nvalid)
{
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    int c = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint32_t m = __vcmpeq4(w[i], 0xffffffffu); /* 0xff per matching byte */
        const int nv = nvalid - 4 * i;
        if (nv < 4) m &= (nv <= 0) ? 0u : ((1u << (8 * nv)) - 1u);
        c += __popc(m) >> 3;
    }
    return c;
}

/* single-pass chained scan over tiles ("decoupled look-back"): tile_state[t] holds, in one
 * 64-bit word, a 2-bit flag, the encode's epoch and a count: A = this tile's own 0xFF count is
 * known, P = the inclusive prefix up to and including this tile is known.  States of earlier
 * encodes carry another epoch and read as "empty", so nothing has to be cleared.  Tiles are
 * handed out by an atomic ticket, so every predecessor of a waiting tile is already running on
 * a resident CTA and the wait terminates.  Warp 0 looks back 32 tiles per step. */
#define K6_FLAG_A 1ull
#define K6_FLAG_P 2ull
#define K6_VALUE_BITS 40
#define K6_EPOCH_BITS 22
#define K6_VALUE_MASK ((1ull << K6_VALUE_BITS) - 1ull)
#define K6_EPOCH_MASK ((1ull << K6_EPOCH_BITS) - 1ull)

__device__ __forceinline__ unsigned long long k6_pack(unsigned long long flag, unsigned long long epoch,
                                                      unsigned long long value)
{
    return (flag << 62) | ((epoch & K6_EPOCH_MASK) << K6_VALUE_BITS) | (value & K6_VALUE_MASK);
}

/* executed by all 32 lanes of warp 0; returns the number of 0xFF bytes in tiles [0, t) */
__device__ __forceinline__ unsigned long long k6_lookback(volatile unsigned long long *ts, unsigned long long t,
                                                          unsigned long long epoch, uint32_t own, int lane)
{
    if (lane == 0) ts[t] = k6_pack(t == 0 ? K6_FLAG_P : K6_FLAG_A, epoch, own);
    unsigned long long excl = 0;
    if (t == 0) return 0ull;
    long long j = (long long)t - 1;
    for (;;) {
        const long long idx = j - lane;
        bool done = false;
        for (;;) {
            unsigned flag = 2u; /* tiles before the first one: prefix 0 */
            unsigned long long val = 0ull;
            if (idx >= 0) {
                const unsigned long long s = ts[idx];
                const bool cur = ((s >> K6_VALUE_BITS) & K6_EPOCH_MASK) == (epoch & K6_EPOCH_MASK);
                flag = cur ? (unsigned)(s >> 62) : 0u;
                val = s & K6_VALUE_MASK;
            }
            const unsigned pm = __ballot_sync(0xffffffffu, flag == 2u);
            const unsigned em = __ballot_sync(0xffffffffu, flag == 0u);
            const unsigned upto = pm ? (0xffffffffu >> (32 - __ffs((int)pm))) : 0xffffffffu; /* lanes 0..first P */
            if ((em & upto) == 0u) {
                unsigned long long c = ((upto >> lane) & 1u) ? val : 0ull;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
                excl += c;
                done = pm != 0u;
                break;
            }
        }
        if (done) break;
        j -= 32;
    }
    if (lane == 0) ts[t] = k6_pack(K6_FLAG_P, epoch, excl + own);
    return excl;
}

__global__ void __launch_bounds__(K6_THREADS) k6_stuff(uint4 *ubuf, EncState *st, unsigned long long *tile_state,
                                                       const uint8_t *header, int header_len, uint8_t *out,
                                                       unsigned long long cap, unsigned long long *len_out,
                                                       unsigned long long epoch, long long nbytes_override,
                                                       int write_eoi)
{
    __shared__ uint32_t sWarp[K6_THREADS / 32];
    __shared__ __align__(16) uint8_t sT[2 * K6_TILE + 32];
    __shared__ unsigned long long sTicket, sExcl;
    const int tid = threadIdx.x;
    if (st->error) {
        if (blockIdx.x == 0 && tid == 0 && len_out) *len_out = 0ull;
        return;
    }
    /* an MCU-row band stuffs only the bytes it owns (jpgb200_band_finish) */
    const unsigned long long nbytes = nbytes_override >= 0 ? (unsigned long long)nbytes_override : st->nbytes;
    const unsigned long long ntiles = (nbytes + K6_TILE - 1) / K6_TILE;
    const unsigned long long trailer = write_eoi ? 2ull : 0ull;
    while (true) {
        if (tid == 0) sTicket = atomicAdd(&st->ticket, 1ull);
        __syncthreads();
        const unsigned long long t = sTicket;
        if (t >= ntiles) break;
        const unsigned long long b0 = t * K6_TILE + (unsigned long long)tid * (16 * K6_ITEMS);
        int nvalid[K6_ITEMS];
        uint4 v[K6_ITEMS];
        int c = 0;
#pragma unroll
        for (int k = 0; k < K6_ITEMS; k++) {
            const unsigned long long bk = b0 + 16ull * k;
            nvalid[k] = 0;
            v[k] = make_uint4(0, 0, 0, 0);
            if (bk < nbytes) {
                const unsigned long long left = nbytes - bk;
                nvalid[k] = left >= 16 ? 16 : (int)left;
                v[k] = ubuf[bk >> 4];
                ubuf[bk >> 4] = make_uint4(0, 0, 0, 0); /* leave the buffer clean for the next encode's K5 */
            }
        }
#pragma unroll
        for (int k = 0; k < K6_ITEMS; k++) c += count_ff16(v[k], nvalid[k]);
        uint32_t tile_ffs;
        const uint32_t ff_before = cta_exclusive_scan<K6_THREADS>((uint32_t)c, sWarp, tid, tile_ffs);
        if (tid < 32) {
            const unsigned long long e = k6_lookback(tile_state, t, epoch, tile_ffs, tid);
            if (tid == 0) sExcl = e;
        }
        __syncthreads();
        const unsigned long long excl = sExcl;
        /* destination of the tile's first byte, and its phase inside a 16-byte line */
        const unsigned long long g0 = (unsigned long long)header_len + t * K6_TILE + excl;
        const unsigned long long tile_in = (nbytes - t * K6_TILE) >= K6_TILE ? K6_TILE : (nbytes - t * K6_TILE);
        const int n_out = (int)tile_in + (int)tile_ffs;
        const bool fits = g0 + (unsigned long long)n_out + trailer <= cap;
        if (fits) {
            const unsigned long long gaddr = (unsigned long long)(uintptr_t)out + g0;
            const int phase = (int)(gaddr & 15ull);
            int pos = phase + tid * (16 * K6_ITEMS) + (int)ff_before;
#pragma unroll
            for (int k = 0; k < K6_ITEMS; k++) {
                const uint32_t w[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    if (i < nvalid[k]) {
                        const uint8_t byte = (uint8_t)(w[i >> 2] >> (8 * (i & 3)));
                        sT[pos++] = byte;
                        if (byte == 0xFF) sT[pos++] = 0x00;
                    }
                }
            }
            __syncthreads();
            const int end = phase + n_out;
            uint8_t *gbase = out + g0 - phase; /* 16-byte aligned address of sT[0]'s image */
            const int first_full = (phase + 15) >> 4, last_full = end >> 4;
            for (int l = first_full + tid; l < last_full; l += K6_THREADS)
                reinterpret_cast<uint4 *>(gbase)[l] = reinterpret_cast<const uint4 *>(sT)[l];
            if (tid < 16) { /* ragged head */
                const int head_end = (first_full << 4) < end ? (first_full << 4) : end;
                const int i = phase + tid;
                if (i < head_end) gbase[i] = sT[i];
            } else if (tid < 32) { /* ragged tail */
                const int tail_start = (last_full << 4) > (first_full << 4) ? (last_full << 4) : (first_full << 4);
                const int i = tail_start + (tid - 16);
                if (i >= phase && i < end && last_full >= first_full) gbase[i] = sT[i];
            }
            if (t == 0)
                for (int i = tid; i < header_len; i += K6_THREADS) out[i] = header[i];
        }
        if (t == ntiles - 1 && tid == 0) {
            const unsigned long long total_ff = excl + (unsigned long long)tile_ffs;
            const unsigned long long out_len = (unsigned long long)header_len + nbytes + total_ff + trailer;
            const bool ok = out_len <= cap;
            if (ok && write_eoi) {
                out[out_len - 2] = 0xFF;
                out[out_len - 1] = 0xD9; /* EOI */
            }
            st->total_ff = total_ff;
            st->out_len = out_len;
            if (!ok) st->error = 1;
            if (len_out) *len_out = ok ? out_len : 0ull;
        }
        __syncthreads();
    }
}

/* OR the previous band's trailing bits into the first byte of this band's stream */
__global__ void k_or_first_byte(uint32_t *ubuf, uint32_t carry) { ubuf[0] |= (carry & 0xffu); }

}  // namespace jpgb200
#endif
