/*
 * K6 -- 0xFF -> 0xFF00 byte stuffing (T.81 F.1.2.3) by a second length/offset scan:
 *   k6_count   : 0xFF bytes per 4 KB tile of the un-stuffed stream, and per 64 tiles
 *   k6_scatter : a tile's insertion offset = sum of the counts before it (at most
 *                n_tiles/64 + 63 numbers, added up by the CTA itself -- no scan kernel, no
 *                serial chain between CTAs); then the stuffed bytes go out.
 * (A single-pass chained scan was measured first: with one wave of tiles per frame its
 * look-back latency dominated -- 40 us per 8K frame against 25 us for this form.)
 * The reference emits no byte stream; the oracle for these bytes is
 * oracle/jpgb200_oracle.c (bw_byte), checked by PIL/OpenCV decodes in tests/.
 */
#ifndef JPGB200_STUFFING_CUH
#define JPGB200_STUFFING_CUH

#include "entropy.cuh"
#include "entropy_fused.cuh"

namespace jpgb200 {

#define K6_THREADS 256
#define K6_TILE 4096     /* bytes of un-stuffed stream per tile: 16 per thread */
#define K6_SUPER_SHIFT 6 /* tiles per super-tile = 64 */

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

/* K6a: 0xFF bytes per tile, and (atomically) per 64 tiles -- the two-level sums from which
 * K6b derives every tile's insertion offset without a scan kernel or a serial chain */
__global__ void __launch_bounds__(K6_THREADS) k6_count(const uint4 *ubuf, const EncState *st, uint32_t *tile_ff,
                                                       uint32_t *super_ff, long long nbytes_override)
{
    __shared__ uint32_t sWarp[K6_THREADS / 32];
    if (st->error) return;
    const unsigned long long nbytes = nbytes_override >= 0 ? (unsigned long long)nbytes_override : st->nbytes;
    const unsigned long long ntiles = (nbytes + K6_TILE - 1) / K6_TILE;
    const int tid = threadIdx.x;
    for (unsigned long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const unsigned long long b0 = t * K6_TILE + (unsigned long long)tid * 16;
        int c = 0;
        if (b0 < nbytes) {
            const unsigned long long left = nbytes - b0;
            c = count_ff16(ubuf[b0 >> 4], left >= 16 ? 16 : (int)left);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
        if ((tid & 31) == 0) sWarp[tid >> 5] = (uint32_t)c;
        __syncthreads();
        if (tid == 0) {
            uint32_t total = 0;
#pragma unroll
            for (int i = 0; i < K6_THREADS / 32; i++) total += sWarp[i];
            tile_ff[t] = total;
            if (total) atomicAdd(&super_ff[t >> K6_SUPER_SHIFT], total);
        }
        __syncthreads();
    }
}

/* K6b: every tile writes its bytes at (tile offset + #0xFF before it), inserting the zero
 * bytes; output is staged in shared memory at the same 16-byte phase as the destination so the
 * copy-out is 16-byte vectorised and coalesced.  Tile 0 also writes the JFIF header, the last
 * tile the EOI marker and the length word.  Every line of the un-stuffed buffer is cleared
 * after it has been read (K5's all-zero invariant). */
__global__ void __launch_bounds__(K6_THREADS) k6_scatter(uint4 *ubuf, EncState *st, const uint32_t *tile_ff,
                                                         const uint32_t *super_ff, const uint8_t *header,
                                                         int header_len, uint8_t *out, unsigned long long cap,
                                                         unsigned long long *len_out, long long nbytes_override,
                                                         int write_eoi)
{
    __shared__ uint32_t sWarp[K6_THREADS / 32];
    __shared__ unsigned long long sPart[K6_THREADS / 32];
    __shared__ __align__(16) uint8_t sT[2 * K6_TILE + 32];
    const int tid = threadIdx.x;
    if (st->error) {
        if (blockIdx.x == 0 && tid == 0 && len_out) *len_out = 0ull;
        return;
    }
    /* an MCU-row band stuffs only the bytes it owns (jpgb200_band_finish) */
    const unsigned long long nbytes = nbytes_override >= 0 ? (unsigned long long)nbytes_override : st->nbytes;
    const unsigned long long ntiles = (nbytes + K6_TILE - 1) / K6_TILE;
    const unsigned long long trailer = write_eoi ? 2ull : 0ull;
    for (unsigned long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const unsigned long long b0 = t * K6_TILE + (unsigned long long)tid * 16;
        int nvalid = 0;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (b0 < nbytes) {
            const unsigned long long left = nbytes - b0;
            nvalid = left >= 16 ? 16 : (int)left;
            v = ubuf[b0 >> 4];
            ubuf[b0 >> 4] = make_uint4(0, 0, 0, 0);
        }
        /* 0xFF bytes in all earlier tiles, from the two-level sums */
        unsigned long long part = 0ull;
        {
            const unsigned long long sup = t >> K6_SUPER_SHIFT;
            for (unsigned long long i = tid; i < sup; i += K6_THREADS) part += super_ff[i];
            for (unsigned long long i = (sup << K6_SUPER_SHIFT) + tid; i < t; i += K6_THREADS) part += tile_ff[i];
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
        if ((tid & 31) == 0) sPart[tid >> 5] = part;
        const int c = count_ff16(v, nvalid);
        uint32_t tile_ffs;
        const uint32_t ff_before = cta_exclusive_scan<K6_THREADS>((uint32_t)c, sWarp, tid, tile_ffs); /* has a barrier */
        unsigned long long excl = 0ull;
#pragma unroll
        for (int i = 0; i < K6_THREADS / 32; i++) excl += sPart[i];
        /* destination of the tile's first byte, and its phase inside a 16-byte line */
        const unsigned long long g0 = (unsigned long long)header_len + t * K6_TILE + excl;
        const unsigned long long tile_in = (nbytes - t * K6_TILE) >= K6_TILE ? K6_TILE : (nbytes - t * K6_TILE);
        const int n_out = (int)tile_in + (int)tile_ffs;
        const bool fits = g0 + (unsigned long long)n_out + trailer <= cap;
        if (fits) {
            const unsigned long long gaddr = (unsigned long long)(uintptr_t)out + g0;
            const int phase = (int)(gaddr & 15ull);
            int pos = phase + tid * 16 + (int)ff_before;
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int i = 0; i < 16; i++) {
                if (i < nvalid) {
                    const uint8_t byte = (uint8_t)(w[i >> 2] >> (8 * (i & 3)));
                    sT[pos++] = byte;
                    if (byte == 0xFF) sT[pos++] = 0x00;
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
