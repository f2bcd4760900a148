/*
 * One very large image across the GPUs of a box (BASELINE.json configs[4]): the image is
 * cut into bands of whole MCU rows, one band per rank.  The entropy-coded scan is ONE bit string with DC
 * prediction running across the bands, so the ranks exchange, with two small ncclAllGathers and without a
 * single host synchronisation in between:
 *
 *   exchange 1  {last DC of Y/Cb/Cr, first DC of Y/Cb/Cr, band bits coded with a ZERO predictor}
 *               The true predictor of a band is the previous band's last DC; it changes nothing but
 *               the three DC codes at the band's start, so every rank can correct every band's bit
 *               count from the DC code lengths alone (k_mg_fix0) -- no second coding pass, only
 *               the band's first chunk is coded again with the right predictor.  With all bit
 *               counts known, every band knows its bit offset in the scan and its bit phase.
 *   exchange 2  {bytes owned, 0xFF bytes among them, trailing partial byte, first 8 bits}
 *               A band OWNS the scan bytes whose last bit lies in its bit range; the leading bits of its
 *               first byte are the previous band's trailing bits.  Whether that one byte is 0xFF (and
 *               needs a stuffed zero) is the only thing the neighbour's bits can change, and every
 *               rank resolves it for every band from the exchanged bytes (k_mg_fix1).  With all stuffed
 *               lengths known, every band knows its byte offset in the final stream.
 *   stitch      KS writes the band's stuffed bytes STRAIGHT into rank 0's output buffer through a
 *               peer-mapped pointer (CUDA IPC over NVLink); rank 0 adds the header, the last rank
 *               the EOI marker.  A last 8-byte allgather is the completion barrier.
 *
 * The reference has no counterpart (its loader cannot even index such an image: int fileSize,
 * src/bitmap.c:27); the bytes are those of a single-GPU encode of the whole image (tests/).
 */
#ifndef JPGB200_MGPU_CUH
#define JPGB200_MGPU_CUH

#include "tail.cuh"

namespace jpgb200 {

#define MG_MAX_RANKS 64

struct MgMsg1 { /* exchange 1, per rank */
    int last_dc[3];
    int first_dc[3];
    unsigned long long bits0; /* band bits with a zero predictor */
};
struct MgMsg2 { /* exchange 2, per rank */
    unsigned long long owned; /* scan bytes owned by the band */
    unsigned long long ff;    /* 0xFF bytes among them, the band's first byte not judged */
    unsigned int tail;        /* trailing partial byte, left-aligned */
    unsigned int head;        /* first 8 bits of the band's stream */
};

__device__ __forceinline__ int mg_dc_len(const HuffPack *h, int chroma, int diff)
{
    const int cat = diff == 0 ? 0 : msb_index((uint32_t)abs(diff)) + 1;
    return (int)(h->dc[chroma][cat] & 31u);
}

/* exchange 1 payload: after K1 and KE(pred = 0) */
__global__ void k_mg_pack1(const int16_t *zz, long long n0, long long n1, long long n2, const uint32_t *super_bits, int n_super,
                           MgMsg1 *msg)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    msg->first_dc[0] = zz[0];
    msg->first_dc[1] = zz[n0 * 64];
    msg->first_dc[2] = zz[(n0 + n1) * 64];
    msg->last_dc[0] = zz[(n0 - 1) * 64];
    msg->last_dc[1] = zz[(n0 + n1 - 1) * 64];
    msg->last_dc[2] = zz[(n0 + n1 + n2 - 1) * 64];
    unsigned long long b = 0;
    for (int i = 0; i < n_super; i++) b += super_bits[i];
    msg->bits0 = b;
}

/* after exchange 1: every band's true bit count, this band's predictor, bit offset and phase; the first
 * chunk's bits leave the two-level sums, because that chunk is coded again with the true predictor */
__global__ void k_mg_fix0(const MgMsg1 *all, int rank, int world, const HuffPack *huff, MgState *st, uint32_t *super_bits,
                          const uint32_t *cta_bits, unsigned long long *all_bits /*[world]*/)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    unsigned long long base = 0;
    for (int j = 0; j < world; j++) {
        long long d = 0;
        if (j > 0)
            for (int c = 0; c < 3; c++) {
                const int f = all[j].first_dc[c], p = all[j - 1].last_dc[c];
                d += mg_dc_len(huff, c ? 1 : 0, f - p) - mg_dc_len(huff, c ? 1 : 0, f);
            }
        const unsigned long long bits = (unsigned long long)((long long)all[j].bits0 + d);
        all_bits[j] = bits;
        if (j == rank) {
            st->bits = bits;
            st->bit_base = base;
            st->phase = (unsigned)(base & 7ull);
            for (int c = 0; c < 3; c++) st->pred[c] = j > 0 ? all[j - 1].last_dc[c] : 0;
        }
        base += bits;
    }
    st->carry = 0u; /* KF runs before the neighbour's trailing bits are known */
    super_bits[0] -= cta_bits[0];
}

/* exchange 2 payload: after the first chunk's second coding and KF(carry = 0) */
__global__ void k_mg_pack2(const MgState *st, int is_last, const uint32_t *tmp, long long slot_words, long long n_chunks,
                           const uint32_t *cta_bits, const uint32_t *super_ff, int n_super, MgMsg2 *msg)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const unsigned long long end = (unsigned long long)st->phase + st->bits;
    unsigned long long own = end >> 3;
    if (is_last && (end & 7ull)) own += 1; /* the 1-padded final byte */
    unsigned long long ff = 0;
    for (int i = 0; i < n_super; i++) ff += super_ff[i];
    msg->owned = own;
    msg->ff = ff;
    msg->head = be32(tmp[0]) >> 24;
    unsigned int tail = 0;
    const unsigned r = (unsigned)(end & 7ull);
    if (!is_last && r) {
        const uint32_t T = cta_bits[n_chunks - 1];
        const uint32_t *slot = tmp + (n_chunks - 1) * slot_words;
        if (T >= r) {
            const uint32_t L = T - r, w = L >> 5, sh = L & 31u, nw = (T + 31u) >> 5;
            const uint32_t hi = be32(slot[w]);
            const uint32_t lo = (w + 1 < nw) ? be32(slot[w + 1]) : 0u;
            const uint32_t v = sh ? ((hi << sh) | (lo >> (32u - sh))) : hi;
            tail = (v >> 24) & (0xffu << (8u - r)) & 0xffu;
        }
    }
    msg->tail = tail;
}

/* after exchange 2: the neighbour's trailing bits, the one 0xFF decision they can change, every band's
 * stuffed length -> this band's byte offset; rank 0 also learns the total length */
__global__ void k_mg_fix1(const MgMsg2 *all, const unsigned long long *all_bits, int rank, int world, int header_len, MgState *st,
                          uint32_t *cta_ff, uint32_t *super_ff, unsigned long long *total_len_out)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    unsigned long long base_bits = 0, off = (unsigned long long)header_len;
    for (int j = 0; j < world; j++) {
        const unsigned phase = (unsigned)(base_bits & 7ull);
        unsigned fix = 0, carry = 0;
        if (j > 0 && phase) {
            carry = (all[j - 1].tail & 0xffu) >> (8u - phase);
            if (all[j].owned >= 1ull && (((carry << (8u - phase)) | (all[j].head >> phase)) & 0xffu) == 0xffu) fix = 1;
        }
        if (j == rank) {
            st->carry = carry;
            st->out_off = j == 0 ? 0ull : off; /* rank 0's KS adds the header itself */
            st->seg_len = all[j].owned + all[j].ff + fix + (j == world - 1 ? 2ull : 0ull);
            if (fix) {
                cta_ff[0] += 1u;
                super_ff[0] += 1u;
            }
        }
        off += all[j].owned + all[j].ff + fix;
        base_bits += all_bits[j];
    }
    st->total_len = off + 2ull;
    if (total_len_out) *total_len_out = off + 2ull;
}

/* the band's segment, written by KS into local memory at the 16-byte phase of its destination, goes to
 * rank 0's buffer with 16-byte stores through the peer pointer (single-byte stores over NVLink are slow) */
__global__ void __launch_bounds__(256) k_mg_push(const MgState *st, const uint8_t *seg, uint8_t *peer, unsigned long long cap)
{
    const unsigned long long off = st->out_off, len = st->seg_len;
    if (off + len > cap) return;
    const uint8_t *src = seg + (off & 15ull);
    uint8_t *dst = peer + off;
    const unsigned long long head = min(len, (16ull - (off & 15ull)) & 15ull);
    const unsigned long long body = (len - head) >> 4, tail0 = head + (body << 4);
    const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (unsigned long long)gridDim.x * blockDim.x;
    if (tid < head) dst[tid] = src[tid];
    const uint4 *s4 = reinterpret_cast<const uint4 *>(src + head);
    uint4 *d4 = reinterpret_cast<uint4 *>(dst + head);
    for (unsigned long long i = tid; i < body; i += nth) d4[i] = s4[i];
    if (tid < len - tail0) dst[tail0 + tid] = src[tail0 + tid];
}

}  // namespace jpgb200
#endif
