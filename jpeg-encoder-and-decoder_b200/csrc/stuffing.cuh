/*
 * K6 -- 0xFF -> 0xFF00 byte stuffing (T.81 F.1.2.3) by a second length/offset scan:
 *   k6_count   : 0xFF bytes per 4 KB tile of the un-stuffed stream
 *   k6_scan    : single-CTA exclusive scan of the tile counts; also writes the JFIF header
 *                in front of the scan, the EOI behind it and the final length
 *   k6_scatter : every tile writes its bytes at (tile offset + #0xFF before it), inserting
 *                the zero bytes; output is staged in shared memory at the same 16-byte phase
 *                as the destination so the copy-out is 16-byte vectorised and coalesced.
 * The reference emits no byte stream; the oracle for these bytes is
 * oracle/jpgb200_oracle.c (bw_byte), checked by PIL/OpenCV decodes in tests/.
 */
#ifndef JPGB200_STUFFING_CUH
#define JPGB200_STUFFING_CUH

#include "entropy.cuh"

namespace jpgb200 {

#define K6_THREADS 256
#define K6_TILE 4096 /* bytes of un-stuffed stream per tile: 16 per thread */

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

__global__ void __launch_bounds__(K6_THREADS) k6_count(const uint4 *ubuf, const EncState *st, uint32_t *tile_ff)
{
    __shared__ uint32_t sWarp[K6_THREADS / 32];
    const unsigned long long nbytes = st->nbytes, ntiles = st->n_tiles;
    const int tid = threadIdx.x;
    for (unsigned long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const unsigned long long b0 = t * K6_TILE + (unsigned long long)tid * 16;
        int c = 0;
        if (b0 < nbytes) {
            const unsigned long long left = nbytes - b0;
            c = count_ff16(ubuf[b0 >> 4], left >= 16 ? 16 : (int)left);
        }
        uint32_t total;
        cta_exclusive_scan<K6_THREADS>((uint32_t)c, sWarp, tid, total);
        if (tid == 0) tile_ff[t] = total;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) k6_scan(const uint32_t *tile_ff, unsigned long long *tile_base,
                                                        EncState *st, const uint8_t *header, int header_len,
                                                        uint8_t *out, unsigned long long cap,
                                                        unsigned long long *len_out)
{
    __shared__ uint32_t sWarp[SCAN_THREADS / 32];
    const unsigned long long total_ff = scan_u32_to_u64(tile_ff, tile_base, (long long)st->n_tiles, sWarp);
    const unsigned long long out_len = (unsigned long long)header_len + st->nbytes + total_ff + 2ull;
    const bool ok = out_len <= cap && st->error == 0;
    if (ok) {
        for (int i = threadIdx.x; i < header_len; i += SCAN_THREADS) out[i] = header[i];
        if (threadIdx.x == 0) {
            out[out_len - 2] = 0xFF;
            out[out_len - 1] = 0xD9; /* EOI */
        }
    }
    if (threadIdx.x == 0) {
        st->total_ff = total_ff;
        st->out_len = out_len;
        st->error = ok ? 0 : 1;
        if (len_out) *len_out = ok ? out_len : 0ull;
    }
}

__global__ void __launch_bounds__(K6_THREADS) k6_scatter(const uint4 *ubuf, const EncState *st,
                                                         const unsigned long long *tile_base, uint8_t *out,
                                                         int header_len)
{
    __shared__ uint32_t sWarp[K6_THREADS / 32];
    __shared__ __align__(16) uint8_t sT[2 * K6_TILE + 32];
    if (st->error) return;
    const unsigned long long nbytes = st->nbytes, ntiles = st->n_tiles;
    const int tid = threadIdx.x;
    for (unsigned long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const unsigned long long b0 = t * K6_TILE + (unsigned long long)tid * 16;
        int nvalid = 0;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (b0 < nbytes) {
            const unsigned long long left = nbytes - b0;
            nvalid = left >= 16 ? 16 : (int)left;
            v = ubuf[b0 >> 4];
        }
        const int c = count_ff16(v, nvalid);
        uint32_t tile_ffs;
        const uint32_t ff_before = cta_exclusive_scan<K6_THREADS>((uint32_t)c, sWarp, tid, tile_ffs);
        /* destination of the tile's first byte, and its phase inside a 16-byte line */
        const unsigned long long g0 = (unsigned long long)header_len + t * K6_TILE + tile_base[t];
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
        const unsigned long long tile_in = (nbytes - t * K6_TILE) >= K6_TILE ? K6_TILE : (nbytes - t * K6_TILE);
        const int n_out = (int)tile_in + (int)tile_ffs;
        const int end = phase + n_out;
        uint8_t *gbase = out + g0 - phase; /* 16-byte aligned address of sT[0]'s image */
        /* whole 16-byte lines */
        const int first_full = (phase + 15) >> 4, last_full = end >> 4;
        for (int l = first_full + tid; l < last_full; l += K6_THREADS)
            reinterpret_cast<uint4 *>(gbase)[l] = reinterpret_cast<const uint4 *>(sT)[l];
        /* ragged head and tail */
        if (tid < 16) {
            const int head_end = (first_full << 4) < end ? (first_full << 4) : end;
            const int i = phase + tid;
            if (i < head_end) gbase[i] = sT[i];
        } else if (tid < 32) {
            const int tail_start = (last_full << 4) > (first_full << 4) ? (last_full << 4) : (first_full << 4);
            const int i = tail_start + (tid - 16);
            if (i >= phase && i < end && last_full >= first_full) gbase[i] = sT[i];
        }
        __syncthreads();
    }
}

}  // namespace jpgb200
#endif
