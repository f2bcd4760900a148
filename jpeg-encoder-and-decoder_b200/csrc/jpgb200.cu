/*
 * libjpgb200.so -- context, launch sequencing and the extern "C" boundary declared in
 * include/jpgb200.h.  Build: see jpeg-encoder-and-decoder_b200/build.py
 *   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
 *
 * Launch sequence of one encode (one stream, no host synchronisation in between):
 *   K1 k1_transform   pixels -> int16 zig-zag coefficients           (HBM-bound, 9 B/px)
 *   KE k_encode_chunks  per-block code lengths and the bit pack, chunk-locally: coefficients are
 *                     read once; every chunk's stream goes to its own slot, its length into
 *                     two-level sums (the device-wide exclusive prefix sum of bit offsets)
 *   KF k_count_ff     a warp per chunk: global bit offset from the two-level sums, the chunk's
 *                     bytes on the global byte grid, 0xFF count -> second pair of two-level sums
 *   KS k_scatter      a warp per chunk: 0xFF00-stuffed bytes straight to their final position
 *                     (the second length/offset scan), JFIF header, EOI and length word
 * There is no CPU arithmetic path in this library.
 */
#include <cuda_runtime.h>

#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/jpgb200.h"
#include "entropy.cuh"
#include "entropy_fused.cuh"
#include "jpgb200_math.cuh"
#include "jpgb200_tables.h"
#include "k1_transform.cuh"
#include "k1_warp.cuh"
#include "k1_pipe.cuh"
#include "k1_pc.cuh"
#include "k1_fused.cuh"
#include "tail.cuh"
#include "mgpu.cuh"
#include "taps.cuh"

using namespace jpgb200;

struct ncclUniqueIdBytes {
    char internal[128];
};

struct HuffSpecs {
    uint8_t nr[4][17];
    uint8_t vals[4][256];
    int nvals[4];
};

enum { KI_K1 = 0, KI_K35, KI_KE, KI_KF, KI_KS, KI_K1E };
static const char *kKernelNames[JPGB200_NUM_KERNELS] = {"k1_transform", "k35_entropy", "k_encode_chunks", "k_count_ff",
                                                        "k_scatter", "k1e_fused"};

struct EvPair {
    int k;
    cudaEvent_t a, b;
};

#ifndef K1_TWELVE_DEFAULT
#define K1_TWELVE_DEFAULT false /* 4:4:4 K1 with eight one-row convert warps and four transform warps per CTA (k1_transform_pc12) */
#endif
#ifndef K1_FOUR_DEFAULT
#define K1_FOUR_DEFAULT false /* 4:4:4 K1 with four transform warps per CTA (k1_transform_pc4) */
#endif
struct jpgb200_ctx {
    int device;
    int sm_count;
    cudaStream_t own_stream, stream, copy_in, copy_out;
    HuffDev *d_huff;
    HuffPack *d_hpack;
    HuffBits *d_hbits;
    int *d_scan_order;
    uint8_t oob[24];
    /* workspaces, grown on demand */
    uint8_t *d_planes[2];
    size_t planes_cap;
    int16_t *d_zz;
    size_t zz_cap;
    uint32_t *d_local_off;
    size_t local_cap;
    uint32_t *d_cta_bits;
    uint32_t *d_super;   /* [n_super_bits | n_super_ff] two-level sums, cleared by K1 */
    size_t n_super_bits, n_super_ff, super_frames;
    unsigned long long *d_cta_base;
    size_t cta_cap;
    uint32_t *d_cta_ff;   /* [n_chunks] 0xFF bytes owned by every chunk (KF) */
    uint32_t *d_cta_tail; /* [n_chunks] last 32 bits of every chunk stream (KE) */
    uint32_t *d_tmp; /* chunk-local streams written by KE, one slot per chunk */
    size_t tmp_cap;
    unsigned long long *d_bit_state, *d_ff_state;
    size_t chain_cap;
    bool fused_attr_set, k1_attr_set, ke_attr_set, k1p_attr_set;
    bool k1_plain; /* force the non-TMA K1 (tests) */
    bool k1_pipe;  /* the software-pipelined K1 (k1_pipe.cuh) */
    bool k1_pc;    /* the producer/consumer K1 (k1_pc.cuh): the default */
    bool k1_ws;    /* the warp-autonomous K1 (k1_warp.cuh) */
    bool k1c_attr_set;
    bool k1_tmaload, k1t_attr_set; /* K1 fed by 2-D TMA tile loads (k1_transform_pct, JPGB200_K1=pct) */
    /* KE's form follows the content: the tight form (7 CTAs per SM at 4:4:4) codes blocks that outgrow their private
     * region a second time; when a sample of the device-side re-code counter says that happens to more than 5 % of
     * the blocks, the roomy form (10-word regions) takes over until the rate -- counted the same way by both forms -- falls
     * below 2 % (jpgb200_set_entropy_form) */
    bool ke_roomy;
    int ke_form_mode;             /* 0 = automatic, 1 = always tight, 2 = always roomy */
    unsigned int *d_recode, *h_recode;
    cudaEvent_t ev_recode;
    bool recode_pending;
    unsigned int recode_last;
    unsigned long long ke_blocks, ke_blocks_sampled, ke_blocks_last;
    int ke_since_sample;
    bool k1_four, k1c4_attr_set;
    bool k1_twelve, k1d_attr_set; /* 4:4:4: eight one-row convert warps + four transform warps per CTA (k1_transform_pc12) */   /* 4:4:4: four transform warps per CTA taking component units in turn (k1_transform_pc4) */
    K1TLoad k1t_maps;
    bool no_tma_store; /* k1_pipe: copy the coefficients out with ordinary stores (A/B runs) */
    CUtensorMap tm32, tm16; /* tensor maps of the coefficient buffer, cached (two entries: the pipelined batch alternates buffers) */
    int16_t *tm_ptr;
    long long tm_rows;
    CUtensorMap tm32_b, tm16_b;
    int16_t *tm_ptr_b;
    long long tm_rows_b;
    bool k1_cta;   /* force the CTA-barrier TMA K1 instead of the warp-autonomous one (tests, A/B runs) */
    EncState *d_state;
    uint8_t *d_header;
    int hdr_w, hdr_h, hdr_q, hdr_ratio, hdr_len;
    uint8_t *d_out[2];
    size_t out_cap;
    unsigned long long *d_len;  /* device */
    unsigned long long *h_len;  /* pinned host, 2 slots, also mapped for kernels */
    int *d_dc;
    size_t dc_cap;
    int *d_tile_sum;
    size_t tile_sum_cap;
    int *d_freq;
    /* geometry of the last transform */
    int w, h, q, ratio;
    long long n[3];
    bool have_bits;
    /* MCU-row band in progress (jpgb200_band_*) */
    unsigned long long band_bits;
    int band_phase, band_last, band_ratio, band_pred[3];
    long long band_n[3];
    bool std_mode;       /* standards-conforming K1: textbook Cb, no last-column shift */
    int sof_w, sof_h;    /* true image size for SOF0 when the planes were padded (0: same as the planes) */
    /* batch overlap: the tail (KF, KS) of frame f runs on a second stream beside KE of frame f+1;
     * the entropy workspaces exist twice and alternate with the frame parity */
    struct EntWs {
        uint32_t *tmp, *cta_bits, *cta_ff, *cta_tail, *super;
        unsigned long long *cta_base;
        size_t tmp_cap, cta_cap, n_super_bits, n_super_ff, super_frames;
    } alt;
    cudaStream_t tail_stream;
    cudaEvent_t ev_ke, ev_ks[2];
    cudaEvent_t ev_in[2], ev_done[2], ev_out[2]; /* jpgb200_encode_batch_host: per plane/output slot */
    bool overlap_on;     /* jpgb200_encode_batch_device overlaps tails (JPGB200_OVERLAP=0 turns it off) */
    bool overlap_now;    /* inside such a batch */
    bool in_batch;       /* inside a batch call: the per-block offset tap is not written */
    bool have_pending;   /* a frame's tail waits to be launched behind the next frame's K1 */
    TailParams pending_q;
    int pending_parity;
    int ws_parity;
    bool opt_huff;       /* per-image optimal Huffman tables (T.81 K.2) instead of the tables.h ones */
    HuffSpecs specs;     /* tables in use: DHT segments and code tables derive from them */
    HuffDev std_huff;    /* the tables.h tables, kept for switching back */
    HuffBits std_hbits;
    HuffPack std_hpack;
    bool legacy_entropy; /* staged KE/KF/KS pipeline (default) instead of the single-launch K35 */
    bool fused_k1e;      /* K1E: transform + chunk encode in one kernel, coefficients stay on chip (default) */
    int k1e_per_sm[4];   /* resident CTAs per SM of k1e_fused<RATIO, STD>, 0 = not yet queried */
    int k1_ctas_per_sm;  /* persistent K1 CTAs per SM: 0 = as many as fit (2); the pipelined batch uses 1 and leaves room for KE */
    bool pipelined;      /* jpgb200_encode_batch_device: K1 of frame f+1 runs beside KE of frame f (three streams) */
    cudaStream_t k1_stream, ke_stream;
    cudaEvent_t ev_k1[2], ev_ked[2], ev_start;
    bool group_small;    /* jpgb200_encode_batch_device: multi-frame launches for small frames (default on) */
    bool k1e_tap;        /* K1E also writes the coefficients to HBM (jpgb200_set_fused_tap) */
    unsigned long long *d_dc_pub; /* K1E: last DC per tile and component, {epoch, DC} words */
    size_t dc_pub_cap;
    unsigned k1e_epoch;
    long long geo_chunks, geo_slot_words; /* chunk geometry of the streams now in d_tmp (KE: 64 MCUs; K1E: tiles) */
    /* quantiser cache */
    int q_cached;
    float mul[2][64];
    /* profiling */
    int profiling; /* 0 off, 1 every kernel, 2 only K1 (the roofline kernel: fewer event records in a timed region),
                      3 only every 8th K1 launch (an event record costs ~2.7 us of stream time) */
    long long k1_seen, ke_seen;
    std::vector<EvPair> pending;
    std::vector<cudaEvent_t> free_events;
    double ms[JPGB200_NUM_KERNELS];
    long long launches[JPGB200_NUM_KERNELS];
    long long launch_total;
    unsigned long long epoch; /* K6 chained-scan generation */
    char err[256];
};

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            snprintf(ctx->err, sizeof ctx->err, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                     __FILE__, __LINE__);                                                                \
            return JPGB200_ERR_CUDA;                                                                     \
        }                                                                                                \
    } while (0)

/* ---- small host helpers ---------------------------------------------------------------- */

extern "C" const char *jpgb200_strerror(int code)
{
    switch (code) {
    case JPGB200_OK: return "ok";
    case JPGB200_ERR_NO_DEVICE: return "no usable CUDA device (this library has no CPU fallback)";
    case JPGB200_ERR_BAD_DIMS: return "width/height must be multiples of 8 (16 wide for 4:2:2)";
    case JPGB200_ERR_BAD_QUALITY: return "quality must be in 24..97";
    case JPGB200_ERR_CUDA: return "CUDA error (see jpgb200_last_error)";
    case JPGB200_ERR_OUT_TOO_SMALL: return "output buffer too small";
    case JPGB200_ERR_BAD_ARG: return "bad argument";
    case JPGB200_ERR_NOMEM: return "out of memory";
    case JPGB200_ERR_NCCL: return "NCCL error or libnccl.so.2 not found (see jpgb200_last_error)";
    default: return "unknown error";
    }
}

extern "C" const char *jpgb200_kernel_name(int k) { return (k >= 0 && k < JPGB200_NUM_KERNELS) ? kKernelNames[k] : ""; }

extern "C" int jpgb200_quant_tables(int quality, int lum[64], int chr[64])
{
    if (quality < 1 || quality > 100) return JPGB200_ERR_BAD_QUALITY;
    int rc = JPGB200_OK;
    for (int i = 0; i < 64; i++) {
        lum[i] = scale_quant_entry(kBaseLum[i], quality);
        chr[i] = scale_quant_entry(kBaseChr[i], quality);
        if (lum[i] < 1 || lum[i] > 255 || chr[i] < 1 || chr[i] > 255) rc = JPGB200_ERR_BAD_QUALITY;
    }
    return rc;
}

extern "C" long long jpgb200_num_blocks(int w, int h, int ratio, long long n[3])
{
    if (w <= 0 || h <= 0 || (h % (ratio == 2 ? 16 : 8)) || ratio < 0 || ratio > 2) return JPGB200_ERR_BAD_DIMS;
    if (w % (ratio ? 16 : 8)) return JPGB200_ERR_BAD_DIMS;
    long long ny = (long long)(w / 8) * (h / 8);
    long long nc = ratio == 2 ? ny / 4 : (ratio ? ny / 2 : ny);
    if (n) {
        n[0] = ny;
        n[1] = nc;
        n[2] = nc;
    }
    return ny + 2 * nc;
}

/* BITS / HUFFVAL of the four tables (0 DC lum, 1 DC chr, 2 AC lum, 3 AC chr) */
static void standard_specs(HuffSpecs *sp)
{
    memset(sp, 0, sizeof *sp);
    for (int t = 0; t < 4; t++) {
        HuffSpec s = huff_spec(t);
        memcpy(sp->nr[t], s.nr, 17);
        memcpy(sp->vals[t], s.vals, (size_t)s.nvals);
        sp->nvals[t] = s.nvals;
    }
}

/* Per-image optimal table: T.81 Annex K.2 (figures K.1-K.4), the procedure the reference's
 * construct_huffman_table sets out to follow and never finishes (src/huffman.c:76-180,
 * SURVEY.md F2).  freq[257], freq[256] = 1 reserves the all-ones code (src/huffman.c:54). */
static void optimal_spec(const int32_t *freq_in, uint8_t nr[17], uint8_t vals[256], int *nvals)
{
    long long freq[257];
    int size[257], next[257], bits[34];
    for (int i = 0; i < 257; i++) {
        freq[i] = freq_in[i];
        size[i] = 0;
        next[i] = -1;
    }
    freq[256] = 1;
    for (;;) {
        int a = -1, b = -1; /* the two least frequent values; on ties the larger value */
        for (int i = 0; i < 257; i++)
            if (freq[i] > 0 && (a < 0 || freq[i] <= freq[a])) a = i;
        for (int i = 0; i < 257; i++)
            if (i != a && freq[i] > 0 && (b < 0 || freq[i] <= freq[b])) b = i;
        if (b < 0) break;
        freq[a] += freq[b];
        freq[b] = 0;
        for (size[a]++; next[a] >= 0;) {
            a = next[a];
            size[a]++;
        }
        next[a] = b;
        for (size[b]++; next[b] >= 0;) {
            b = next[b];
            size[b]++;
        }
    }
    memset(bits, 0, sizeof bits);
    for (int i = 0; i < 257; i++)
        if (size[i] > 0) bits[size[i] > 32 ? 32 : size[i]]++;
    int i = 32;
    for (; i > 16; i--)
        while (bits[i] > 0) { /* K.3: no code longer than 16 bits */
            int j = i - 2;
            while (bits[j] == 0) j--;
            bits[i] -= 2;
            bits[i - 1] += 1;
            bits[j + 1] += 2;
            bits[j] -= 1;
        }
    while (bits[i] == 0) i--;
    bits[i]--; /* drop the reserved code point */
    memset(nr, 0, 17);
    for (int l = 1; l <= 16; l++) nr[l] = (uint8_t)bits[l];
    int k = 0;
    for (int l = 1; l <= 32; l++)
        for (int j = 0; j < 256; j++)
            if (size[j] == l) vals[k++] = (uint8_t)j;
    *nvals = k;
}

/* canonical codes from the specs (T.81 Annex C) */
static void build_huff(const HuffSpecs *specs, HuffDev *h)
{
    memset(h, 0, sizeof *h);
    for (int t = 0; t < 4; t++) {
        struct {
            const uint8_t *nr, *vals;
        } sp = {specs->nr[t], specs->vals[t]};
        unsigned code = 0;
        int k = 0;
        for (int l = 1; l <= 16; l++) {
            for (int i = 0; i < sp.nr[l]; i++) {
                const int sym = sp.vals[k++];
                const uint32_t e = ((uint32_t)code << 8) | (uint32_t)l;
                if (t < 2)
                    h->dc[t][sym] = e;
                else
                    h->ac[t - 2][sym] = e;
                code++;
            }
            code <<= 1;
        }
    }
}

/* KE's tables: code pre-shifted by the category, total length in the low 5 bits */
static void build_huff_pack(const HuffDev *h, HuffPack *k)
{
    memset(k, 0, sizeof *k);
    for (int t = 0; t < 2; t++) {
        for (int c = 0; c < 16; c++) {
            const uint32_t e = h->dc[t][c], len = e & 0xff;
            if (len) k->dc[t][c] = (((e >> 8) << c) << 5) | (len + (uint32_t)c);
        }
        for (int sym = 0; sym < 256; sym++) {
            const uint32_t e = h->ac[t][sym], len = e & 0xff;
            const uint32_t c = (sym == 0xF0) ? 0u : (uint32_t)(sym & 15);
            if (len) k->ac[t][sym] = (((e >> 8) << c) << 5) | (len + c);
        }
        /* small values: the word of (run, category) with the additional bits of the value merged */
        for (int run = 0; run < 16; run++)
            for (int v = -3; v <= 3; v++) {
                if (v == 0) continue;
                const int av = v < 0 ? -v : v, c = av == 1 ? 1 : 2;
                const uint32_t extra = (uint32_t)(v - (v < 0 ? 1 : 0)) & ((1u << c) - 1u);
                const uint32_t w = k->ac[t][(run << 4) | c];
                k->lut[t][run * 8 + v + 4] = w ? (w | (extra << 5)) : 0u;
            }
        k->lut[t][HP_LUT_EOB] = k->ac[t][0x00]; /* value 0 never indexes the lut: two of its spare entries */
        k->lut[t][HP_LUT_ZRL] = k->ac[t][0xF0];
    }
}

/* total bits per symbol (code + additional bits), AC extended to runs of 0..62 zeros */
static void build_huff_bits(const HuffDev *h, HuffBits *b)
{
    memset(b, 0, sizeof *b);
    for (int t = 0; t < 2; t++) {
        for (int c = 0; c < 16; c++)
            if (h->dc[t][c] & 0xff) b->dc[t][c] = (uint8_t)((h->dc[t][c] & 0xff) + c);
        const int zrl = (int)(h->ac[t][0xF0] & 0xff);
        b->ac[t][0] = (uint8_t)(h->ac[t][0x00] & 0xff); /* EOB */
        for (int run = 0; run < 63; run++)
            for (int c = 1; c <= 10; c++) {
                const int len = (int)(h->ac[t][((run & 15) << 4) | c] & 0xff);
                if (len) b->ac[t][(run << 4) | c] = (uint8_t)((run >> 4) * zrl + len + c);
            }
    }
}

static size_t put16(uint8_t *o, size_t p, unsigned v)
{
    o[p] = (uint8_t)(v >> 8);
    o[p + 1] = (uint8_t)v;
    return p + 2;
}

/* JFIF marker segments SOI..SOS (T.81 Annex B).  The DQT carries, at zig-zag
 * position of (v,u), the divisor the reference actually applies there: T'[u][v]. */
static int jfif_header_specs(int w, int h, int quality, int ratio, const HuffSpecs *specs, uint8_t *o, size_t cap, size_t *len,
                             bool check_dims = true);

extern "C" int jpgb200_jfif_header(int w, int h, int quality, int ratio, uint8_t *o, size_t cap, size_t *len)
{
    HuffSpecs sp;
    standard_specs(&sp);
    return jfif_header_specs(w, h, quality, ratio, &sp, o, cap, len);
}

static int jfif_header_specs(int w, int h, int quality, int ratio, const HuffSpecs *specs, uint8_t *o, size_t cap, size_t *len,
                             bool check_dims)
{
    int lum[64], chr[64];
    if (check_dims && jpgb200_num_blocks(w, h, ratio, NULL) < 0) return JPGB200_ERR_BAD_DIMS;
    if (w < 1 || h < 1 || w > 65535 || h > 65535 || ratio < 0 || ratio > 2) return JPGB200_ERR_BAD_DIMS;
    if (quality < 24 || quality > 97 || jpgb200_quant_tables(quality, lum, chr)) return JPGB200_ERR_BAD_QUALITY;
    if (cap < 640) return JPGB200_ERR_OUT_TOO_SMALL;
    size_t p = 0;
    p = put16(o, p, 0xFFD8);
    p = put16(o, p, 0xFFE0);
    p = put16(o, p, 16);
    memcpy(o + p, "JFIF\0", 5);
    p += 5;
    o[p++] = 1;
    o[p++] = 1;
    o[p++] = 0;
    p = put16(o, p, 1);
    p = put16(o, p, 1);
    o[p++] = 0;
    o[p++] = 0;
    for (int t = 0; t < 2; t++) {
        const int *tab = t ? chr : lum;
        p = put16(o, p, 0xFFDB);
        p = put16(o, p, 67);
        o[p++] = (uint8_t)t;
        for (int v = 0; v < 8; v++)
            for (int u = 0; u < 8; u++) o[p + kZigZag[v * 8 + u]] = (uint8_t)tab[u * 8 + v];
        p += 64;
    }
    p = put16(o, p, 0xFFC0);
    p = put16(o, p, 17);
    o[p++] = 8;
    p = put16(o, p, (unsigned)h);
    p = put16(o, p, (unsigned)w);
    o[p++] = 3;
    const uint8_t sof[9] = {1, (uint8_t)(ratio == 1 ? 0x21 : (ratio == 2 ? 0x22 : 0x11)), 0, 2, 0x11, 1, 3, 0x11, 1};
    memcpy(o + p, sof, 9);
    p += 9;
    static const int order[4] = {0, 2, 1, 3};
    static const uint8_t tcth[4] = {0x00, 0x10, 0x01, 0x11};
    for (int i = 0; i < 4; i++) {
        const int t = order[i];
        p = put16(o, p, 0xFFC4);
        p = put16(o, p, (unsigned)(2 + 1 + 16 + specs->nvals[t]));
        o[p++] = tcth[i];
        memcpy(o + p, specs->nr[t] + 1, 16);
        p += 16;
        memcpy(o + p, specs->vals[t], (size_t)specs->nvals[t]);
        p += (size_t)specs->nvals[t];
    }
    p = put16(o, p, 0xFFDA);
    p = put16(o, p, 12);
    const uint8_t sos[10] = {3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 63, 0};
    memcpy(o + p, sos, 10);
    p += 10;
    if (len) *len = p;
    return JPGB200_OK;
}

/* ---- context --------------------------------------------------------------------------- */

extern "C" int jpgb200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

extern "C" int jpgb200_create(jpgb200_ctx **out, int device)
{
    if (!out) return JPGB200_ERR_BAD_ARG;
    *out = NULL;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return JPGB200_ERR_NO_DEVICE;
    jpgb200_ctx *ctx = new jpgb200_ctx();
    memset(ctx->oob, 0, sizeof ctx->oob);
    ctx->device = device;
    ctx->q_cached = -1;
    ctx->legacy_entropy = true; /* measured faster than the fused K35 so far */
    ctx->k1_pc = true;          /* K1 default: the producer/consumer form (fastest measured) */
    ctx->k1_four = K1_FOUR_DEFAULT;
    ctx->k1_twelve = K1_TWELVE_DEFAULT;
    ctx->group_small = true;
    {
        const char *pz = getenv("JPGB200_PIPELINED"); /* "0": K1 and KE of a batch strictly one after the other */
        ctx->pipelined = !(pz && !strcmp(pz, "0"));
        const char *cz = getenv("JPGB200_K1_CTAS");
        ctx->k1_ctas_per_sm = cz ? atoi(cz) : 0;
    }
    ctx->fused_k1e = false;     /* transform + chunk encode in one kernel (k1_fused.cuh): measured slower than K1 -> KE so far */
    {
        const char *fz = getenv("JPGB200_FUSED"); /* "1": K1E instead of K1 and KE as separate kernels (A/B runs) */
        if (fz && !strcmp(fz, "1")) ctx->fused_k1e = true;
        const char *e = getenv("JPGB200_ENTROPY"); /* "fused" | "staged": experiment switch for bench runs */
        if (e && !strcmp(e, "fused") && jpgb200_has_ab_forms()) ctx->legacy_entropy = false;
        const char *k = getenv("JPGB200_K1"); /* "plain" | "tma" */
        const char *ov = getenv("JPGB200_OVERLAP");
        ctx->overlap_on = !(ov && !strcmp(ov, "0"));
        if (k && !strcmp(k, "plain")) ctx->k1_plain = true, ctx->k1_pc = false;
        if (k && !strcmp(k, "pc_notma")) ctx->no_tma_store = true;
        if (k && !strcmp(k, "pct")) ctx->k1_tmaload = true;
        const char *kf = getenv("JPGB200_KE_FORM"); /* "tight" | "roomy": pin KE's form (default: follows the content) */
        ctx->ke_form_mode = kf ? (!strcmp(kf, "tight") ? 1 : (!strcmp(kf, "roomy") ? 2 : 0)) : 0;
        ctx->ke_roomy = ctx->ke_form_mode == 2;
        if (k && !strcmp(k, "pc4")) ctx->k1_four = true;
        if (k && !strcmp(k, "pc12")) ctx->k1_twelve = true;
        if (k && !strcmp(k, "pc8")) ctx->k1_twelve = false;
        if (k && !strcmp(k, "pc3")) ctx->k1_four = false;
        if (jpgb200_has_ab_forms()) {
            if (k && !strcmp(k, "cta")) ctx->k1_cta = true, ctx->k1_pc = false;
            if (k && !strcmp(k, "pipe")) ctx->k1_pipe = true, ctx->k1_pc = false;
            if (k && !strcmp(k, "pipe_notma")) ctx->k1_pipe = ctx->no_tma_store = true, ctx->k1_pc = false;
            if (k && !strcmp(k, "ws")) ctx->k1_ws = true, ctx->k1_pc = false;
        }
    }
    ctx->hdr_w = -1;
    ctx->err[0] = 0;
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        delete ctx;
        return JPGB200_ERR_NO_DEVICE;
    }
    ctx->sm_count = prop.multiProcessorCount;
    standard_specs(&ctx->specs);
    HuffDev hh;
    HuffBits hb;
    build_huff(&ctx->specs, &hh);
    build_huff_bits(&hh, &hb);
    HuffPack hp;
    build_huff_pack(&hh, &hp);
    ctx->std_huff = hh;
    ctx->std_hbits = hb;
    ctx->std_hpack = hp;
    int rc = [&]() -> int {
        CK(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
        {
            int lo = 0, hi = 0;
            CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
            /* Stream priorities of the pipelined batch, measured on 8K frames (profiles/README.md): K1 and KE EQUAL and
             * the tails above them: 152.2 us per frame = K1 + KE + 2.5 us.  K1 above KE (the first design: "K1's few
             * persistent CTAs go first") lets K1 of frame f+1 move in before the chunk encode of frame f has drained:
             * 156.2 us; with the tails at the lowest priority as well 158.2 us (K1 of frame f+2 waits for the tails of
             * frame f: it clears the two-level sums they read); KE above the tails: 160 us.  A third set of entropy
             * workspaces, which takes the tails off K1's dependencies, was slower too (157.7 us). */
            const char *p1 = getenv("JPGB200_K1_PRIO"), *p2 = getenv("JPGB200_KE_PRIO"), *tp = getenv("JPGB200_TAIL_PRIO"); /* A/B runs */
            CK(cudaStreamCreateWithPriority(&ctx->tail_stream, cudaStreamNonBlocking, (tp && !strcmp(tp, "lo")) ? lo : hi));
            CK(cudaStreamCreateWithPriority(&ctx->k1_stream, cudaStreamNonBlocking, (p1 && !strcmp(p1, "hi")) ? hi : lo));
            CK(cudaStreamCreateWithPriority(&ctx->ke_stream, cudaStreamNonBlocking, (p2 && !strcmp(p2, "hi")) ? hi : lo));
            for (int sl = 0; sl < 2; sl++) {
                CK(cudaEventCreateWithFlags(&ctx->ev_k1[sl], cudaEventDisableTiming));
                CK(cudaEventCreateWithFlags(&ctx->ev_ked[sl], cudaEventDisableTiming));
            }
            CK(cudaEventCreateWithFlags(&ctx->ev_start, cudaEventDisableTiming));
        }
        CK(cudaEventCreateWithFlags(&ctx->ev_ke, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ctx->ev_ks[0], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ctx->ev_ks[1], cudaEventDisableTiming));
        for (int sl = 0; sl < 2; sl++) {
            CK(cudaEventCreateWithFlags(&ctx->ev_in[sl], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&ctx->ev_done[sl], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&ctx->ev_out[sl], cudaEventDisableTiming));
        }
        ctx->stream = ctx->own_stream;
        CK(cudaMalloc(&ctx->d_huff, sizeof(HuffDev)));
        CK(cudaMemcpy(ctx->d_huff, &hh, sizeof hh, cudaMemcpyHostToDevice));
        CK(cudaMalloc(&ctx->d_hpack, sizeof(HuffPack)));
        CK(cudaMemcpy(ctx->d_hpack, &hp, sizeof hp, cudaMemcpyHostToDevice));
        CK(cudaMalloc(&ctx->d_hbits, sizeof(HuffBits)));
        CK(cudaMemcpy(ctx->d_hbits, &hb, sizeof hb, cudaMemcpyHostToDevice));
        CK(cudaMalloc(&ctx->d_scan_order, 64 * sizeof(int)));
        CK(cudaMemcpy(ctx->d_scan_order, kZigZag, 64 * sizeof(int), cudaMemcpyHostToDevice));
        CK(cudaMalloc(&ctx->d_state, sizeof(EncState)));
        CK(cudaMemset(ctx->d_state, 0, sizeof(EncState)));
        CK(cudaMalloc(&ctx->d_header, 1024));
        CK(cudaMalloc(&ctx->d_len, 2 * sizeof(unsigned long long)));
        CK(cudaHostAlloc(&ctx->h_len, 2 * sizeof(unsigned long long), cudaHostAllocMapped));
        CK(cudaMalloc(&ctx->d_freq, 4 * 257 * sizeof(int)));
        CK(cudaMalloc(&ctx->d_recode, sizeof(unsigned int)));
        CK(cudaMemset(ctx->d_recode, 0, sizeof(unsigned int)));
        CK(cudaHostAlloc(&ctx->h_recode, sizeof(unsigned int), cudaHostAllocDefault));
        *ctx->h_recode = 0u;
        CK(cudaEventCreateWithFlags(&ctx->ev_recode, cudaEventDisableTiming));
        return JPGB200_OK;
    }();
    if (rc != JPGB200_OK) {
        fprintf(stderr, "jpgb200_create: %s\n", ctx->err);
        delete ctx;
        return rc;
    }
    *out = ctx;
    return JPGB200_OK;
}

extern "C" void jpgb200_destroy(jpgb200_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    void *ptrs[] = {ctx->d_huff,     ctx->d_hpack, ctx->d_hbits, ctx->d_scan_order, ctx->d_planes[0], ctx->d_planes[1], ctx->d_zz,   ctx->d_local_off,
                    ctx->d_cta_bits, ctx->d_super, ctx->d_cta_base,   ctx->d_cta_ff,    ctx->d_cta_tail,  ctx->d_tmp,
                    ctx->d_state,    ctx->d_bit_state, ctx->d_ff_state, ctx->d_header,     ctx->d_out[0],    ctx->d_out[1],    ctx->d_len,  ctx->d_dc,
                    ctx->d_tile_sum, ctx->d_freq,      ctx->d_dc_pub};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    if (ctx->h_len) cudaFreeHost(ctx->h_len);
    if (ctx->d_recode) cudaFree(ctx->d_recode);
    if (ctx->h_recode) cudaFreeHost(ctx->h_recode);
    if (ctx->ev_recode) cudaEventDestroy(ctx->ev_recode);
    for (auto &e : ctx->pending) {
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
    }
    for (auto &e : ctx->free_events) cudaEventDestroy(e);
    cudaStreamDestroy(ctx->own_stream);
    cudaStreamDestroy(ctx->copy_in);
    cudaStreamDestroy(ctx->copy_out);
    cudaStreamDestroy(ctx->tail_stream);
    cudaStreamDestroy(ctx->k1_stream);
    cudaStreamDestroy(ctx->ke_stream);
    cudaEventDestroy(ctx->ev_start);
    for (int sl = 0; sl < 2; sl++) {
        cudaEventDestroy(ctx->ev_k1[sl]);
        cudaEventDestroy(ctx->ev_ked[sl]);
    }
    cudaEventDestroy(ctx->ev_ke);
    cudaEventDestroy(ctx->ev_ks[0]);
    cudaEventDestroy(ctx->ev_ks[1]);
    for (int sl = 0; sl < 2; sl++) {
        cudaEventDestroy(ctx->ev_in[sl]);
        cudaEventDestroy(ctx->ev_done[sl]);
        cudaEventDestroy(ctx->ev_out[sl]);
    }
    {
        void *alts[] = {ctx->alt.tmp, ctx->alt.cta_bits, ctx->alt.cta_ff, ctx->alt.cta_tail, ctx->alt.super, ctx->alt.cta_base};
        for (void *q : alts)
            if (q) cudaFree(q);
    }
    delete ctx;
}

extern "C" const char *jpgb200_last_error(const jpgb200_ctx *ctx) { return ctx ? ctx->err : ""; }

extern "C" int jpgb200_set_stream(jpgb200_ctx *ctx, void *s)
{
    if (!ctx) return JPGB200_ERR_BAD_ARG;
    ctx->stream = s ? (cudaStream_t)s : ctx->own_stream;
    return JPGB200_OK;
}

extern "C" int jpgb200_sync(jpgb200_ctx *ctx)
{
    if (!ctx) return JPGB200_ERR_BAD_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    return JPGB200_OK;
}

extern "C" int jpgb200_has_ab_forms(void)
{
#ifdef JPGB200_AB
    return 1;
#else
    return 0;
#endif
}

extern "C" int jpgb200_set_plain_transform(jpgb200_ctx *ctx, int on)
{
    if (!ctx) return JPGB200_ERR_BAD_ARG;
    if (((on >= 2 && on <= 4) || on == 7 || on == 8) && !jpgb200_has_ab_forms()) return JPGB200_ERR_BAD_ARG; /* forms 2..4, 7 and 8 are A/B builds only */
    ctx->k1_pc = (on == 0 || on == 5 || on == 6 || on == 7 || on == 8);
    ctx->k1_twelve = (on == 8) || (on == 0 && K1_TWELVE_DEFAULT); /* 8: twelve warps per CTA (4:4:4), 5: eight */
    ctx->k1_tmaload = (on == 6); /* 5: the cp.async-fed form explicitly, 6: the TMA-load-fed form */
    ctx->k1_four = (on == 7) || (on == 0 && K1_FOUR_DEFAULT); /* 7: four transform warps (4:4:4, A/B builds), 5: three */
    ctx->k1_plain = (on == 1);
    ctx->k1_cta = (on == 2);
    ctx->k1_pipe = (on == 3);
    ctx->k1_ws = (on == 4);
    return JPGB200_OK;
}

extern "C" int jpgb200_set_entropy_form(jpgb200_ctx *ctx, int mode)
{
    if (!ctx || mode < 0 || mode > 2) return JPGB200_ERR_BAD_ARG;
    ctx->ke_form_mode = mode;
    if (mode) ctx->ke_roomy = (mode == 2);
    return JPGB200_OK;
}

extern "C" int jpgb200_entropy_form(jpgb200_ctx *ctx)
{
    return ctx ? (ctx->ke_roomy ? 2 : 1) : JPGB200_ERR_BAD_ARG;
}

extern "C" int jpgb200_set_staged_entropy(jpgb200_ctx *ctx, int on)
{
    if (!ctx) return JPGB200_ERR_BAD_ARG;
    if (!on && !jpgb200_has_ab_forms()) return JPGB200_ERR_BAD_ARG; /* K35 is an A/B build only */
    ctx->legacy_entropy = on != 0;
    return JPGB200_OK;
}

extern "C" int jpgb200_set_fused_transform(jpgb200_ctx *ctx, int on)
{
    if (!ctx) return JPGB200_ERR_BAD_ARG;
    ctx->fused_k1e = on != 0;
    return JPGB200_OK;
}

extern "C" int jpgb200_set_pipelined_batch(jpgb200_ctx *ctx, int on)
{
    if (!ctx) return JPGB200_ERR_BAD_ARG;
    ctx->pipelined = on != 0;
    return JPGB200_OK;
}

extern "C" int jpgb200_set_frame_groups(jpgb200_ctx *ctx, int on)
{
    if (!ctx) return JPGB200_ERR_BAD_ARG;
    ctx->group_small = on != 0;
    return JPGB200_OK;
}

extern "C" int jpgb200_set_fused_tap(jpgb200_ctx *ctx, int on)
{
    if (!ctx) return JPGB200_ERR_BAD_ARG;
    ctx->k1e_tap = on != 0;
    return JPGB200_OK;
}

extern "C" int jpgb200_set_batch_overlap(jpgb200_ctx *ctx, int on)
{
    if (!ctx) return JPGB200_ERR_BAD_ARG;
    ctx->overlap_on = on != 0;
    return JPGB200_OK;
}

extern "C" int jpgb200_set_standard_mode(jpgb200_ctx *ctx, int on)
{
    if (!ctx) return JPGB200_ERR_BAD_ARG;
    ctx->std_mode = on != 0;
    return JPGB200_OK;
}

extern "C" int jpgb200_set_optimized_huffman(jpgb200_ctx *ctx, int on)
{
    if (!ctx) return JPGB200_ERR_BAD_ARG;
    CK(cudaSetDevice(ctx->device));
    const bool was = ctx->opt_huff;
    ctx->opt_huff = on != 0;
    if (was && !ctx->opt_huff) { /* back to the tables.h tables */
        standard_specs(&ctx->specs);
        CK(cudaMemcpyAsync(ctx->d_huff, &ctx->std_huff, sizeof(HuffDev), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->d_hbits, &ctx->std_hbits, sizeof(HuffBits), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->d_hpack, &ctx->std_hpack, sizeof(HuffPack), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->hdr_w = -1;
    }
    return JPGB200_OK;
}

extern "C" int jpgb200_huffman_specs(const jpgb200_ctx *ctx, uint8_t nr[4][17], uint8_t vals[4][256], int nvals[4])
{
    if (!ctx || !nr || !vals || !nvals) return JPGB200_ERR_BAD_ARG;
    memcpy(nr, ctx->specs.nr, sizeof ctx->specs.nr);
    memcpy(vals, ctx->specs.vals, sizeof ctx->specs.vals);
    memcpy(nvals, ctx->specs.nvals, sizeof ctx->specs.nvals);
    return JPGB200_OK;
}

extern "C" int jpgb200_set_oob_bytes(jpgb200_ctx *ctx, const uint8_t oob[24])
{
    if (!ctx || !oob) return JPGB200_ERR_BAD_ARG;
    memcpy(ctx->oob, oob, 24);
    return JPGB200_OK;
}

/* ---- profiling --------------------------------------------------------------------------- */

static cudaEvent_t get_event(jpgb200_ctx *ctx)
{
    if (!ctx->free_events.empty()) {
        cudaEvent_t e = ctx->free_events.back();
        ctx->free_events.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

struct LaunchScope {
    jpgb200_ctx *ctx;
    int k;
    cudaEvent_t a, b;
    bool on;
    LaunchScope(jpgb200_ctx *c, int kk) : ctx(c), k(kk)
    {
        const bool head = (kk == KI_K1 || kk == KI_K1E); /* the transform kernel */
        on = ctx->profiling == 1 || (ctx->profiling == 2 && head) || (ctx->profiling == 3 && head && (ctx->k1_seen++ & 7) == 0) ||
             (ctx->profiling == 3 && kk == KI_KE && (ctx->ke_seen++ & 7) == 4); /* the chunk encode: in other frames than K1 */
        if (on) {
            a = get_event(ctx);
            b = get_event(ctx);
            cudaEventRecord(a, ctx->stream);
        }
    }
    ~LaunchScope()
    {
        ctx->launch_total++;
        if (on) {
            ctx->launches[k]++; /* launches[] counts the TIMED launches: ms[k] / launches[k] is their mean */
            cudaEventRecord(b, ctx->stream);
            ctx->pending.push_back({k, a, b});
        }
    }
};

static int resolve_profile(jpgb200_ctx *ctx)
{
    if (ctx->pending.empty()) return JPGB200_OK;
    CK(cudaStreamSynchronize(ctx->stream));
    for (auto &e : ctx->pending) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e.a, e.b));
        ctx->ms[e.k] += (double)ms;
        ctx->free_events.push_back(e.a);
        ctx->free_events.push_back(e.b);
    }
    ctx->pending.clear();
    return JPGB200_OK;
}

extern "C" int jpgb200_set_profiling(jpgb200_ctx *ctx, int on)
{
    if (!ctx) return JPGB200_ERR_BAD_ARG;
    int rc = resolve_profile(ctx);
    ctx->profiling = on < 0 ? 0 : (on > 3 ? 1 : on);
    return rc;
}

extern "C" int jpgb200_reset_profile(jpgb200_ctx *ctx)
{
    if (!ctx) return JPGB200_ERR_BAD_ARG;
    int rc = resolve_profile(ctx);
    for (int i = 0; i < JPGB200_NUM_KERNELS; i++) {
        ctx->ms[i] = 0.0;
        ctx->launches[i] = 0;
    }
    return rc;
}

extern "C" int jpgb200_kernel_times(jpgb200_ctx *ctx, double *ms, long long *launches)
{
    if (!ctx) return JPGB200_ERR_BAD_ARG;
    int rc = resolve_profile(ctx);
    for (int i = 0; i < JPGB200_NUM_KERNELS; i++) {
        if (ms) ms[i] = ctx->ms[i];
        if (launches) launches[i] = ctx->launches[i];
    }
    return rc;
}

extern "C" long long jpgb200_launch_count(const jpgb200_ctx *ctx) { return ctx ? ctx->launch_total : 0; }

/* ---- workspaces -------------------------------------------------------------------------- */

template <class T>
static int grow(jpgb200_ctx *ctx, T **p, size_t *cap, size_t need_bytes)
{
    if (*cap >= need_bytes && *p) return JPGB200_OK;
    CK(cudaStreamSynchronize(ctx->stream));
    if (*p) CK(cudaFree(*p));
    *p = NULL;
    *cap = 0;
    size_t want = need_bytes + need_bytes / 8 + 4096;
    cudaError_t e = cudaMalloc((void **)p, want);
    if (e != cudaSuccess) {
        snprintf(ctx->err, sizeof ctx->err, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
        return JPGB200_ERR_NOMEM;
    }
    *cap = want;
    return JPGB200_OK;
}

static int validate(int w, int h, int quality, int ratio, long long n[3])
{
    if (jpgb200_num_blocks(w, h, ratio, n) < 0) return JPGB200_ERR_BAD_DIMS;
    if (w > 65535 || h > 65535) return JPGB200_ERR_BAD_DIMS; /* SOF0 fields are 16 bit */
    if (quality < 24 || quality > 97) return JPGB200_ERR_BAD_QUALITY;
    return JPGB200_OK;
}

static long long n_chunks(const long long n[3]) { return (n[1] + EC_MCUS - 1) / EC_MCUS; }
/* words of one chunk's slot: worst case 1658 bits per block (20-bit DC + 63 x 26-bit AC), 16-byte multiple */
static long long slot_words(const long long n[3])
{
    const long long bpm = n[0] / n[1] + 2; /* 3, 4 or 6 blocks per MCU */
    return (((long long)EC_MCUS * bpm * 1658 + 31) / 32 + 4) & ~3ll;
}

/* K1E's chunks are K1's tiles: 32 blocks of one block row per full-resolution component */
static long long tile_chunks(int w, int h) { return (long long)((w / 8 + K1_TILE_BLOCKS - 1) / K1_TILE_BLOCKS) * (h / 8); }
static long long tile_slot_words(int ratio)
{
    const long long blocks = ratio == 0 ? 96 : 64;
    return ((blocks * 1658 + 31) / 32 + 4) & ~3ll;
}

/* w, h > 0: also size the chunk workspaces for K1E's tile geometry; nframes > 1: room for that many frames in
 * every per-frame array (multi-frame launches of small frames) */
static int ensure_workspace(jpgb200_ctx *ctx, const long long n[3], int nframes, int w = 0, int h = 0, int ratio = 0)
{
    const long long nb = n[0] + n[1] + n[2];
    const size_t G = (size_t)(nframes < 1 ? 1 : nframes);
    int rc;
    if ((rc = grow(ctx, &ctx->d_zz, &ctx->zz_cap, (size_t)nb * 128 * G))) return rc;
    if ((rc = grow(ctx, &ctx->d_local_off, &ctx->local_cap, (size_t)nb * 4))) return rc;
    size_t nc = (size_t)n_chunks(n);
    size_t tmp_words = nc * (size_t)slot_words(n);
    if (w > 0 && ratio != 2) {
        const size_t nct = (size_t)tile_chunks(w, h);
        if (nct * (size_t)tile_slot_words(ratio) > tmp_words) tmp_words = nct * (size_t)tile_slot_words(ratio);
        if (nct > nc) nc = nct;
        unsigned long long *before = ctx->d_dc_pub;
        if ((rc = grow(ctx, &ctx->d_dc_pub, &ctx->dc_pub_cap, nct * 3 * 8))) return rc;
        if (ctx->d_dc_pub != before) CK(cudaMemsetAsync(ctx->d_dc_pub, 0, ctx->dc_pub_cap, ctx->stream)); /* epoch 0 is never used */
    }
    const size_t nct_all = nc * G; /* chunk entries over all frames */
    if (ctx->cta_cap < nct_all * 8 || !ctx->d_cta_bits) {
        size_t c1 = 0, c2 = 0;
        if (ctx->d_cta_bits) {
            CK(cudaStreamSynchronize(ctx->stream));
            CK(cudaFree(ctx->d_cta_bits));
            CK(cudaFree(ctx->d_cta_base));
            ctx->d_cta_bits = NULL;
            ctx->d_cta_base = NULL;
        }
        if ((rc = grow(ctx, &ctx->d_cta_bits, &c1, nct_all * 4))) return rc;
        if ((rc = grow(ctx, &ctx->d_cta_base, &c2, nct_all * 8))) return rc;
        if (ctx->d_cta_ff) CK(cudaFree(ctx->d_cta_ff));
        if (ctx->d_cta_tail) CK(cudaFree(ctx->d_cta_tail));
        ctx->d_cta_ff = NULL;
        ctx->d_cta_tail = NULL;
        size_t c3 = 0, c4 = 0;
        if ((rc = grow(ctx, &ctx->d_cta_ff, &c3, nct_all * 4))) return rc;
        if ((rc = grow(ctx, &ctx->d_cta_tail, &c4, nct_all * 4))) return rc;
        ctx->cta_cap = nct_all * 8;
    }
    if (ctx->chain_cap < nc || !ctx->d_bit_state) {
        size_t c1 = 0, c2 = 0;
        if (ctx->d_bit_state) {
            CK(cudaStreamSynchronize(ctx->stream));
            CK(cudaFree(ctx->d_bit_state));
            CK(cudaFree(ctx->d_ff_state));
            ctx->d_bit_state = NULL;
            ctx->d_ff_state = NULL;
        }
        if ((rc = grow(ctx, &ctx->d_bit_state, &c1, nc * 8))) return rc;
        if ((rc = grow(ctx, &ctx->d_ff_state, &c2, nc * 8))) return rc;
        CK(cudaMemsetAsync(ctx->d_bit_state, 0, c1, ctx->stream));
        CK(cudaMemsetAsync(ctx->d_ff_state, 0, c2, ctx->stream));
        ctx->chain_cap = c1 / 8;
    }
    if ((rc = grow(ctx, &ctx->d_tmp, &ctx->tmp_cap, tmp_words * G * 4))) return rc;
    {
        /* two-level sums: [n_super_bits | n_super_ff] per frame, frames back to back */
        const size_t nsb = nc / 64 + 2, nsf = nc / 64 + 2;
        if (!ctx->d_super || ctx->n_super_bits < nsb || ctx->n_super_ff < nsf || ctx->super_frames < G) {
            if (ctx->d_super) {
                CK(cudaStreamSynchronize(ctx->stream));
                CK(cudaFree(ctx->d_super));
                ctx->d_super = NULL;
            }
            size_t c3 = 0;
            if ((rc = grow(ctx, &ctx->d_super, &c3, (nsb + nsf) * G * 4))) return rc;
            CK(cudaMemsetAsync(ctx->d_super, 0, c3, ctx->stream));
            ctx->n_super_bits = nsb;
            ctx->n_super_ff = nsf;
            ctx->super_frames = G;
        }
    }
    return JPGB200_OK;
}

static int ensure_tables(jpgb200_ctx *ctx, int w, int h, int quality, int ratio)
{
    const bool padded = ctx->sof_w > 0;
    if (padded) { /* SOF0 carries the true size, the planes hold whole MCUs */
        w = ctx->sof_w;
        h = ctx->sof_h;
    }
    if (ctx->q_cached != quality) {
        int lum[64], chr[64];
        if (jpgb200_quant_tables(quality, lum, chr)) return JPGB200_ERR_BAD_QUALITY;
        make_quant_multipliers(lum, ctx->mul[0]);
        make_quant_multipliers(chr, ctx->mul[1]);
        ctx->q_cached = quality;
    }
    if (ctx->hdr_w != w || ctx->hdr_h != h || ctx->hdr_q != quality || ctx->hdr_ratio != ratio) {
        uint8_t hdr[1024];
        size_t len = 0;
        int rc = jfif_header_specs(w, h, quality, ratio, &ctx->specs, hdr, sizeof hdr, &len, !padded);
        if (rc) return rc;
        CK(cudaMemcpyAsync(ctx->d_header, hdr, len, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream)); /* hdr is a stack buffer */
        ctx->hdr_w = w;
        ctx->hdr_h = h;
        ctx->hdr_q = quality;
        ctx->hdr_ratio = ratio;
        ctx->hdr_len = (int)len;
    }
    return JPGB200_OK;
}

/* ---- launch sequencing --------------------------------------------------------------------- */

/* tensor maps of the coefficient buffer for K1's TMA stores (k1_pipe.cuh); cached per buffer */
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static void make_store_maps(jpgb200_ctx *ctx, K1PStore *st, int16_t *zz, const long long n[3], long long zz_stride, int nframes)
{
    memset(st, 0, sizeof *st);
    st->row_cb = n[0];
    st->row_cr = n[0] + n[1];
    st->rows_per_frame = zz_stride / 64;
    const long long rows = (nframes > 1 ? (long long)(nframes - 1) * st->rows_per_frame : 0) + n[0] + n[1] + n[2];
    if (ctx->no_tma_store || (zz_stride % 64) || rows >= (1ll << 31)) return;
    if (ctx->tm_ptr == zz && ctx->tm_rows == rows) {
        st->box32 = ctx->tm32;
        st->box16 = ctx->tm16;
        st->enabled = 1;
        return;
    }
    if (ctx->tm_ptr_b == zz && ctx->tm_rows_b == rows) {
        st->box32 = ctx->tm32_b;
        st->box16 = ctx->tm16_b;
        st->enabled = 1;
        return;
    }
    /* the older entry moves to the second slot */
    ctx->tm32_b = ctx->tm32;
    ctx->tm16_b = ctx->tm16;
    ctx->tm_ptr_b = ctx->tm_ptr;
    ctx->tm_rows_b = ctx->tm_rows;
    ctx->tm_ptr = NULL;
    static EncodeTiledFn fn = NULL;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *pfn = NULL;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &pfn, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)pfn;
    }
    if (!fn) return;
    const cuuint64_t gdim[2] = {64, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {128};
    const cuuint32_t estr[2] = {1, 1};
    const cuuint32_t box32[2] = {64, 32}, box16[2] = {64, 16};
    if (fn(&ctx->tm32, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, zz, gdim, gstride, box32, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return;
    if (fn(&ctx->tm16, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, zz, gdim, gstride, box16, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return;
    ctx->tm_ptr = zz;
    ctx->tm_rows = rows;
    st->box32 = ctx->tm32;
    st->box16 = ctx->tm16;
    st->enabled = 1;
}

/* tensor maps of the three planes for K1's 2-D TMA tile loads (k1_transform_pct) */
static bool make_load_maps(jpgb200_ctx *ctx, K1TLoad *ld, const uint8_t *d_r, const uint8_t *d_g, const uint8_t *d_b, int w, int h,
                           int nframes, long long plane_stride, int has_halo)
{
    (void)ctx;
    memset(ld, 0, sizeof *ld);
    if (nframes > 1 && (plane_stride % w)) return false;
    static EncodeTiledFn fn = NULL;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *pfn = NULL;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &pfn, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)pfn;
    }
    if (!fn) return false;
    ld->rows_per_frame = nframes > 1 ? (int)(plane_stride / w) : 0;
    ld->y_bias = has_halo ? 1 : 0;
    const long long rows = (long long)(nframes > 1 ? (nframes - 1) : 0) * ld->rows_per_frame + h + (has_halo ? 1 : 0);
    if (rows >= (1ll << 31)) return false;
    const uint8_t *base[3] = {d_r, d_g, d_b};
    const cuuint64_t gdim[2] = {(cuuint64_t)w, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)w};
    const cuuint32_t estr[2] = {1, 1};
    const cuuint32_t box[2] = {K1_ROW_BYTES, 8}, boxa[2] = {16, 8};
    for (int c = 0; c < 3; c++) {
        void *g = (void *)(base[c] - (has_halo ? w : 0));
        if (fn(&ld->plane[c], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, g, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return false;
        if (fn(&ld->alt[c], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, g, gdim, gstride, boxa, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return false;
    }
    return true;
}

static int launch_k1(jpgb200_ctx *ctx, const uint8_t *d_r, const uint8_t *d_g, const uint8_t *d_b, int w, int h,
                     int ratio, const long long n[3], int nframes, long long plane_stride, int16_t *zz,
                     long long zz_stride, int has_halo)
{
    K1Params p;
    p.r = d_r;
    p.g = d_g;
    p.b = d_b;
    p.zz_y = zz;
    p.zz_cb = zz + n[0] * 64;
    p.zz_cr = zz + (n[0] + n[1]) * 64;
    p.plane_stride = plane_stride;
    p.zz_stride = zz_stride;
    p.W = w;
    p.H = h;
    p.BW = w / 8;
    p.has_halo = has_halo;
    p.clear = ctx->d_super; /* K3 and K6 accumulate their two-level sums into it */
    p.clear_n = (int)(ctx->n_super_bits + ctx->n_super_ff);
    memcpy(p.mul, ctx->mul, sizeof p.mul);
    memcpy(p.oob, ctx->oob, 24);
    p.tiles_per_row = (p.BW + K1_TILE_BLOCKS - 1) / K1_TILE_BLOCKS;
    p.n_frames = nframes;
    {
        const unsigned long long per_frame = (unsigned long long)p.tiles_per_row * (unsigned long long)(h / 8);
        const unsigned long long total = per_frame * (unsigned long long)nframes;
        const bool ok = total * per_frame < (1ull << 40) && total < (1ull << 31);
        p.magic_pf = ok ? ((1ull << 40) + per_frame - 1) / per_frame : 0ull;
        p.magic_tpr = ok ? ((1ull << 40) + (unsigned long long)p.tiles_per_row - 1) / (unsigned long long)p.tiles_per_row : 0ull;
    }
    const bool aligned16 = (w % 16 == 0) && (((uintptr_t)d_r | (uintptr_t)d_g | (uintptr_t)d_b) % 16 == 0) &&
                           (plane_stride % 16 == 0);
    if (ratio == 2) {
        /* 4:2:0: one CTA per tile of 16 MCUs (k1_transform420) */
        dim3 grid((unsigned)p.tiles_per_row, (unsigned)(h / 16), (unsigned)nframes);
        LaunchScope ls(ctx, KI_K1);
        if (ctx->std_mode)
            k1_transform420<true><<<grid, K1_THREADS, 0, ctx->stream>>>(p);
        else
            k1_transform420<false><<<grid, K1_THREADS, 0, ctx->stream>>>(p);
    } else if (aligned16 && ctx->k1_pc && ctx->k1_tmaload && !ctx->k1_plain && w >= 256 &&
               make_load_maps(ctx, &ctx->k1t_maps, d_r, d_g, d_b, w, h, nframes, plane_stride, has_halo)) {
        /* the same roles fed by 2-D TMA tile loads issued by the store warp (k1_pc.cuh: k1_transform_pct) */
        const size_t smem = sizeof(K1TSmem) + 1024;
        if (!ctx->k1t_attr_set) {
            CK(cudaFuncSetAttribute((k1_transform_pct<0, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CK(cudaFuncSetAttribute((k1_transform_pct<1, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CK(cudaFuncSetAttribute((k1_transform_pct<0, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CK(cudaFuncSetAttribute((k1_transform_pct<1, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ctx->k1t_attr_set = true;
        }
        const long long n_tiles = (long long)p.tiles_per_row * (h / 8) * nframes;
        const long long resident = (long long)ctx->sm_count * (ctx->k1_ctas_per_sm > 0 ? ctx->k1_ctas_per_sm : K1C_MIN_CTAS);
        const unsigned grid = (unsigned)(n_tiles < resident ? n_tiles : resident);
        K1PStore st;
        make_store_maps(ctx, &st, zz, n, zz_stride, nframes);
        LaunchScope ls(ctx, KI_K1);
        if (ctx->std_mode) {
            if (ratio == 0)
                k1_transform_pct<0, true><<<grid, K1C_THREADS, smem, ctx->stream>>>(p, n_tiles, st, ctx->k1t_maps);
            else
                k1_transform_pct<1, true><<<grid, K1C_THREADS, smem, ctx->stream>>>(p, n_tiles, st, ctx->k1t_maps);
        } else if (ratio == 0)
            k1_transform_pct<0, false><<<grid, K1C_THREADS, smem, ctx->stream>>>(p, n_tiles, st, ctx->k1t_maps);
        else
            k1_transform_pct<1, false><<<grid, K1C_THREADS, smem, ctx->stream>>>(p, n_tiles, st, ctx->k1t_maps);
#ifdef JPGB200_AB
    } else if (aligned16 && ctx->k1_pc && ctx->k1_four && ratio == 0 && !ctx->k1_plain) {
        /* 4:4:4: four convert + four transform warps, component units handed round-robin (k1_pc.cuh: k1_transform_pc4) */
        const size_t smem = sizeof(K1C4Smem) + 1024;
        if (!ctx->k1c4_attr_set) {
            CK(cudaFuncSetAttribute((k1_transform_pc4<false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CK(cudaFuncSetAttribute((k1_transform_pc4<true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ctx->k1c4_attr_set = true;
        }
        const long long n_tiles = (long long)p.tiles_per_row * (h / 8) * nframes;
        const long long resident = (long long)ctx->sm_count * (ctx->k1_ctas_per_sm > 0 ? ctx->k1_ctas_per_sm : K1C_MIN_CTAS);
        const unsigned grid = (unsigned)(n_tiles < resident ? n_tiles : resident);
        K1PStore st;
        make_store_maps(ctx, &st, zz, n, zz_stride, nframes);
        LaunchScope ls(ctx, KI_K1);
        if (ctx->std_mode)
            k1_transform_pc4<true><<<grid, K1C_THREADS, smem, ctx->stream>>>(p, n_tiles, st);
        else
            k1_transform_pc4<false><<<grid, K1C_THREADS, smem, ctx->stream>>>(p, n_tiles, st);
#endif
#ifdef JPGB200_AB
    } else if (aligned16 && ctx->k1_pc && ctx->k1_twelve && ratio == 0 && !ctx->k1_plain) {
        /* 4:4:4: eight one-row convert warps + four transform warps, registers re-divided with setmaxnreg (k1_transform_pc12) */
        const size_t smem = sizeof(K1DSmem) + 1024;
        if (!ctx->k1d_attr_set) {
            CK(cudaFuncSetAttribute((k1_transform_pc12<false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CK(cudaFuncSetAttribute((k1_transform_pc12<true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ctx->k1d_attr_set = true;
        }
        const long long n_tiles = (long long)p.tiles_per_row * (h / 8) * nframes;
        const long long resident = (long long)ctx->sm_count * (ctx->k1_ctas_per_sm > 0 ? ctx->k1_ctas_per_sm : K1C_MIN_CTAS);
        const unsigned grid = (unsigned)(n_tiles < resident ? n_tiles : resident);
        K1PStore st;
        make_store_maps(ctx, &st, zz, n, zz_stride, nframes);
        LaunchScope ls(ctx, KI_K1);
        if (ctx->std_mode)
            k1_transform_pc12<true><<<grid, K1D_THREADS, smem, ctx->stream>>>(p, n_tiles, st);
        else
            k1_transform_pc12<false><<<grid, K1D_THREADS, smem, ctx->stream>>>(p, n_tiles, st);
#endif
    } else if (aligned16 && ctx->k1_pc && !ctx->k1_plain) {
        /* eight specialised warps per CTA handing tiles over through mbarriers (k1_pc.cuh) */
        const size_t smem = K1C_PIPES * sizeof(K1CSmem) + 1024;
        if (!ctx->k1c_attr_set) {
            CK(cudaFuncSetAttribute((k1_transform_pc<0, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CK(cudaFuncSetAttribute((k1_transform_pc<1, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CK(cudaFuncSetAttribute((k1_transform_pc<0, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CK(cudaFuncSetAttribute((k1_transform_pc<1, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ctx->k1c_attr_set = true;
        }
        const long long n_tiles = (long long)p.tiles_per_row * (h / 8) * nframes;
        const long long resident = (long long)ctx->sm_count * (ctx->k1_ctas_per_sm > 0 ? ctx->k1_ctas_per_sm : K1C_MIN_CTAS);
        /* pipelines (eight warps each) in flight; a CTA holds K1C_PIPES of them */
        const unsigned grid = (unsigned)(((n_tiles < resident ? n_tiles : resident) + K1C_PIPES - 1) / K1C_PIPES);
        K1PStore st;
        make_store_maps(ctx, &st, zz, n, zz_stride, nframes);
        LaunchScope ls(ctx, KI_K1);
        if (ctx->std_mode) {
            if (ratio == 0)
                k1_transform_pc<0, true><<<grid, K1C_CTA_THREADS, smem, ctx->stream>>>(p, n_tiles, st);
            else
                k1_transform_pc<1, true><<<grid, K1C_CTA_THREADS, smem, ctx->stream>>>(p, n_tiles, st);
        } else if (ratio == 0)
            k1_transform_pc<0, false><<<grid, K1C_CTA_THREADS, smem, ctx->stream>>>(p, n_tiles, st);
        else
            k1_transform_pc<1, false><<<grid, K1C_CTA_THREADS, smem, ctx->stream>>>(p, n_tiles, st);
#ifdef JPGB200_AB
    } else if (aligned16 && ctx->k1_pipe && !ctx->k1_plain) {
        /* persistent CTAs, warps specialised and software-pipelined over tiles (k1_pipe.cuh) */
        if (!ctx->k1p_attr_set) {
            CK(cudaFuncSetAttribute((k1_transform_pipe<0, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K1PSmem) + 1024));
            CK(cudaFuncSetAttribute((k1_transform_pipe<1, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K1PSmem) + 1024));
            CK(cudaFuncSetAttribute((k1_transform_pipe<0, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K1PSmem) + 1024));
            CK(cudaFuncSetAttribute((k1_transform_pipe<1, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K1PSmem) + 1024));
            ctx->k1p_attr_set = true;
        }
        const long long n_tiles = (long long)p.tiles_per_row * (h / 8) * nframes;
        const long long resident = (long long)ctx->sm_count * K1P_MIN_CTAS;
        const unsigned grid = (unsigned)(n_tiles < resident ? n_tiles : resident);
        K1PStore st;
        make_store_maps(ctx, &st, zz, n, zz_stride, nframes);
        const size_t smem = sizeof(K1PSmem) + 1024;
        LaunchScope ls(ctx, KI_K1);
        if (ctx->std_mode) {
            if (ratio == 0)
                k1_transform_pipe<0, true><<<grid, K1_THREADS, smem, ctx->stream>>>(p, n_tiles, st);
            else
                k1_transform_pipe<1, true><<<grid, K1_THREADS, smem, ctx->stream>>>(p, n_tiles, st);
        } else if (ratio == 0)
            k1_transform_pipe<0, false><<<grid, K1_THREADS, smem, ctx->stream>>>(p, n_tiles, st);
        else
            k1_transform_pipe<1, false><<<grid, K1_THREADS, smem, ctx->stream>>>(p, n_tiles, st);
    } else if (aligned16 && !ctx->k1_plain && (!ctx->k1_cta || ctx->std_mode)) {
        /* persistent CTAs of four autonomous warps fed by TMA bulk copies (k1_warp.cuh) */
        if (!ctx->k1_attr_set) {
            CK(cudaFuncSetAttribute(k1_transform_tma<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K1SmemTma)));
            CK(cudaFuncSetAttribute(k1_transform_tma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K1SmemTma)));
            CK(cudaFuncSetAttribute(k1_transform_ws<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K1WSmem)));
            CK(cudaFuncSetAttribute(k1_transform_ws<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K1WSmem)));
            CK(cudaFuncSetAttribute((k1_transform_ws<0, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K1WSmem)));
            CK(cudaFuncSetAttribute((k1_transform_ws<1, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K1WSmem)));
            ctx->k1_attr_set = true;
        }
        const long long n_tiles = (long long)p.tiles_per_row * (h / 8) * nframes;
        const long long resident = (long long)ctx->sm_count * K1W_MIN_CTAS;
        const unsigned grid = (unsigned)(n_tiles < resident ? n_tiles : resident);
        LaunchScope ls(ctx, KI_K1);
        if (ctx->std_mode) {
            if (ratio == 0)
                k1_transform_ws<0, true><<<grid, K1_THREADS, sizeof(K1WSmem), ctx->stream>>>(p, n_tiles);
            else
                k1_transform_ws<1, true><<<grid, K1_THREADS, sizeof(K1WSmem), ctx->stream>>>(p, n_tiles);
        } else if (ratio == 0)
            k1_transform_ws<0><<<grid, K1_THREADS, sizeof(K1WSmem), ctx->stream>>>(p, n_tiles);
        else
            k1_transform_ws<1><<<grid, K1_THREADS, sizeof(K1WSmem), ctx->stream>>>(p, n_tiles);
    } else if (aligned16 && !ctx->k1_plain) {
        /* persistent CTAs fed by TMA bulk copies, phases separated by CTA barriers */
        if (!ctx->k1_attr_set) {
            CK(cudaFuncSetAttribute(k1_transform_ws<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K1WSmem)));
            CK(cudaFuncSetAttribute(k1_transform_ws<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K1WSmem)));
            CK(cudaFuncSetAttribute(k1_transform_tma<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K1SmemTma)));
            CK(cudaFuncSetAttribute(k1_transform_tma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K1SmemTma)));
            ctx->k1_attr_set = true;
        }
        const long long n_tiles = (long long)p.tiles_per_row * (h / 8) * nframes;
        const long long resident = (long long)ctx->sm_count * 4;
        const unsigned grid = (unsigned)(n_tiles < resident ? n_tiles : resident);
        LaunchScope ls(ctx, KI_K1);
        if (ratio == 0)
            k1_transform_tma<0><<<grid, K1_THREADS, sizeof(K1SmemTma), ctx->stream>>>(p, n_tiles);
        else
            k1_transform_tma<1><<<grid, K1_THREADS, sizeof(K1SmemTma), ctx->stream>>>(p, n_tiles);
#endif /* JPGB200_AB */
    } else {
        dim3 grid((unsigned)p.tiles_per_row, (unsigned)(h / 8), (unsigned)nframes);
        LaunchScope ls(ctx, KI_K1);
        if (ctx->std_mode) {
            if (ratio == 0)
                k1_transform<0, true><<<grid, K1_THREADS, 0, ctx->stream>>>(p);
            else
                k1_transform<1, true><<<grid, K1_THREADS, 0, ctx->stream>>>(p);
        } else if (ratio == 0)
            k1_transform<0><<<grid, K1_THREADS, 0, ctx->stream>>>(p);
        else
            k1_transform<1><<<grid, K1_THREADS, 0, ctx->stream>>>(p);
    }
    CK(cudaGetLastError());
    return JPGB200_OK;
}

/* K1E: transform + chunk encode in one persistent kernel (k1_fused.cuh).  The chunk streams, their
 * lengths and the two-level sums land where KE would leave them (tile geometry), KF/KS follow. */
template <int RATIO, bool STD>
static int launch_k1e_t(jpgb200_ctx *ctx, const K1Params &p, long long n_tiles, const K1PStore &st, const K1EParams &e)
{
    const size_t smem = sizeof(K1ESmem) + 1024;
    int &per_sm = ctx->k1e_per_sm[RATIO * 2 + (STD ? 1 : 0)];
    if (per_sm == 0) {
        CK(cudaFuncSetAttribute((k1e_fused<RATIO, STD>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (k1e_fused<RATIO, STD>), K1ECfg<RATIO>::THREADS, smem));
        if (per_sm < 1) {
            per_sm = 0;
            snprintf(ctx->err, sizeof ctx->err, "k1e_fused does not fit an SM");
            return JPGB200_ERR_CUDA;
        }
    }
    /* every CTA must be resident: tile t waits for the DC that tile t-1's CTA publishes */
    const long long resident = (long long)ctx->sm_count * per_sm;
    const unsigned grid = (unsigned)(n_tiles < resident ? n_tiles : resident);
    LaunchScope ls(ctx, KI_K1E);
    k1e_fused<RATIO, STD><<<grid, K1ECfg<RATIO>::THREADS, smem, ctx->stream>>>(p, n_tiles, st, e);
    return JPGB200_OK;
}

static bool k1e_usable(const jpgb200_ctx *ctx, const uint8_t *d_r, const uint8_t *d_g, const uint8_t *d_b, int w, int ratio,
                       long long plane_stride)
{
    return ctx->fused_k1e && ctx->legacy_entropy && !ctx->opt_huff && ratio != 2 && (w % 16 == 0) &&
           (((uintptr_t)d_r | (uintptr_t)d_g | (uintptr_t)d_b) % 16 == 0) && (plane_stride % 16 == 0);
}

static int launch_k1e(jpgb200_ctx *ctx, const uint8_t *d_r, const uint8_t *d_g, const uint8_t *d_b, int w, int h, int ratio,
                      const long long n[3], int16_t *zz_tap, int has_halo, const int *pred)
{
    K1Params p;
    p.r = d_r;
    p.g = d_g;
    p.b = d_b;
    p.zz_y = zz_tap;
    p.zz_cb = zz_tap ? zz_tap + n[0] * 64 : NULL;
    p.zz_cr = zz_tap ? zz_tap + (n[0] + n[1]) * 64 : NULL;
    p.plane_stride = 0;
    p.zz_stride = 0;
    p.W = w;
    p.H = h;
    p.BW = w / 8;
    p.has_halo = has_halo;
    p.clear = NULL; /* the two-level sums are cleared by a memset in front of the kernel: it accumulates into them */
    p.clear_n = 0;
    memcpy(p.mul, ctx->mul, sizeof p.mul);
    memcpy(p.oob, ctx->oob, 24);
    p.tiles_per_row = (p.BW + K1_TILE_BLOCKS - 1) / K1_TILE_BLOCKS;
    p.n_frames = 1;
    {
        const unsigned long long per_frame = (unsigned long long)p.tiles_per_row * (unsigned long long)(h / 8);
        const bool ok = per_frame * per_frame < (1ull << 40) && per_frame < (1ull << 31);
        p.magic_pf = ok ? ((1ull << 40) + per_frame - 1) / per_frame : 0ull;
        p.magic_tpr = ok ? ((1ull << 40) + (unsigned long long)p.tiles_per_row - 1) / (unsigned long long)p.tiles_per_row : 0ull;
    }
    const long long n_tiles = tile_chunks(w, h);
    ctx->geo_chunks = n_tiles;
    ctx->geo_slot_words = tile_slot_words(ratio);
    K1EParams e;
    e.tmp = ctx->d_tmp;
    e.slot_words = ctx->geo_slot_words;
    e.cta_bits = ctx->d_cta_bits;
    e.cta_tail = ctx->d_cta_tail;
    e.super_bits = ctx->d_super;
    e.dc_pub = ctx->d_dc_pub;
    if (++ctx->k1e_epoch == 0u) ctx->k1e_epoch = 1u;
    e.epoch = ctx->k1e_epoch;
    e.huff = ctx->d_hpack;
    e.store_coefs = zz_tap ? 1 : 0;
    for (int c = 0; c < 3; c++) e.dc_pred[c] = pred ? pred[c] : 0;
    K1PStore st;
    memset(&st, 0, sizeof st);
    if (zz_tap) make_store_maps(ctx, &st, zz_tap, n, 0, 1);
    CK(cudaMemsetAsync(ctx->d_super, 0, (ctx->n_super_bits + ctx->n_super_ff) * 4, ctx->stream));
    int rc;
    if (ctx->std_mode)
        rc = ratio == 0 ? launch_k1e_t<0, true>(ctx, p, n_tiles, st, e) : launch_k1e_t<1, true>(ctx, p, n_tiles, st, e);
    else
        rc = ratio == 0 ? launch_k1e_t<0, false>(ctx, p, n_tiles, st, e) : launch_k1e_t<1, false>(ctx, p, n_tiles, st, e);
    if (rc) return rc;
    CK(cudaGetLastError());
    ctx->have_bits = false;
    return JPGB200_OK;
}

/* KE in the form the content asks for.  Every fourth launch a copy of the device-side re-code counter is queued
 * behind the kernel; its arrival is polled (never waited for) before a later launch and moves ctx->ke_roomy. */
static int launch_ke(jpgb200_ctx *ctx, const KeParams &e0, int ratio, dim3 grid, cudaStream_t st, unsigned long long blocks)
{
    if (!ctx->ke_attr_set) {
        CK(cudaFuncSetAttribute((k_encode_chunks<0, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, KeSmem<0, false>::BYTES));
        CK(cudaFuncSetAttribute((k_encode_chunks<1, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, KeSmem<1, false>::BYTES));
        CK(cudaFuncSetAttribute((k_encode_chunks<2, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, KeSmem<2, false>::BYTES));
        CK(cudaFuncSetAttribute((k_encode_chunks<0, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, KeSmem<0, true>::BYTES));
        CK(cudaFuncSetAttribute((k_encode_chunks<1, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, KeSmem<1, true>::BYTES));
        ctx->ke_attr_set = true;
    }
    if (ctx->recode_pending && cudaEventQuery(ctx->ev_recode) == cudaSuccess) {
        const unsigned int now = *ctx->h_recode;
        const unsigned long long nb = ctx->ke_blocks_sampled - ctx->ke_blocks_last;
        const double rate = nb ? (double)(unsigned int)(now - ctx->recode_last) / (double)nb : 0.0;
        ctx->recode_last = now;
        ctx->ke_blocks_last = ctx->ke_blocks_sampled;
        ctx->recode_pending = false;
        if (ctx->ke_form_mode == 0) {
            if (!ctx->ke_roomy && rate > 0.05) ctx->ke_roomy = true;
            else if (ctx->ke_roomy && rate < 0.02) ctx->ke_roomy = false;
        }
    }
    KeParams e = e0;
    e.recode_count = ctx->d_recode;
    const bool roomy = ctx->ke_roomy && ratio != 2; /* 4:2:0 has one form */
    {
        LaunchScope ls(ctx, KI_KE);
        if (ratio == 0) {
            if (roomy)
                k_encode_chunks<0, true><<<grid, KeCfg<0, true>::NT, KeSmem<0, true>::BYTES, st>>>(e, ctx->d_hpack);
            else
                k_encode_chunks<0, false><<<grid, KeCfg<0, false>::NT, KeSmem<0, false>::BYTES, st>>>(e, ctx->d_hpack);
        } else if (ratio == 1) {
            if (roomy)
                k_encode_chunks<1, true><<<grid, KeCfg<1, true>::NT, KeSmem<1, true>::BYTES, st>>>(e, ctx->d_hpack);
            else
                k_encode_chunks<1, false><<<grid, KeCfg<1, false>::NT, KeSmem<1, false>::BYTES, st>>>(e, ctx->d_hpack);
        } else
            k_encode_chunks<2, false><<<grid, KeCfg<2, false>::NT, KeSmem<2, false>::BYTES, st>>>(e, ctx->d_hpack);
    }
    CK(cudaGetLastError());
    ctx->ke_blocks += blocks;
    if (ctx->ke_form_mode == 0 && !ctx->recode_pending && ++ctx->ke_since_sample >= 4) {
        ctx->ke_since_sample = 0;
        CK(cudaMemcpyAsync(ctx->h_recode, ctx->d_recode, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
        CK(cudaEventRecord(ctx->ev_recode, st));
        ctx->ke_blocks_sampled = ctx->ke_blocks;
        ctx->recode_pending = true;
    }
    return JPGB200_OK;
}

/* KE: every chunk's bit lengths and its packed, chunk-local stream; two-level sums of the lengths */
static int launch_bits(jpgb200_ctx *ctx, const int16_t *zz, int ratio, const long long n[3], const int *pred = NULL,
                       bool clear_supers = false)
{
    if (clear_supers) CK(cudaMemsetAsync(ctx->d_super, 0, (ctx->n_super_bits + ctx->n_super_ff) * 4, ctx->stream));
    KeParams e;
    e.zz = zz;
    e.coff1 = n[0];
    e.coff2 = n[0] + n[1];
    e.n_mcu = n[1];
    for (int c = 0; c < 3; c++) e.dc_pred[c] = pred ? pred[c] : 0;
    e.tmp = ctx->d_tmp;
    e.slot_words = slot_words(n);
    e.cta_bits = ctx->d_cta_bits;
    e.cta_tail = ctx->d_cta_tail;
    e.super_bits = ctx->d_super;
    e.local_off = ctx->overlap_now || ctx->in_batch ? NULL : ctx->d_local_off; /* parity tap: not inside batches */
    e.zz_fstride = e.tmp_fstride = 0;
    e.cta_fstride = e.super_fstride = 0;
    e.dc_pred_dev = NULL;
    const unsigned nc = (unsigned)n_chunks(n);
    ctx->geo_chunks = n_chunks(n);
    ctx->geo_slot_words = slot_words(n);
    {
        const int rc = launch_ke(ctx, e, ratio, dim3(nc), ctx->stream, (unsigned long long)(n[0] + n[1] + n[2]));
        if (rc) return rc;
    }
    CK(cudaGetLastError());
    ctx->have_bits = true;
    return JPGB200_OK;
}

static int run_tail(jpgb200_ctx *ctx, const TailParams &q)
{
    const unsigned grid = (unsigned)((q.n_chunks + TAIL_WARPS - 1) / TAIL_WARPS);
    {
        LaunchScope ls(ctx, KI_KF);
        k_count_ff<<<grid, TAIL_THREADS, 0, ctx->stream>>>(q);
    }
    CK(cudaGetLastError());
    {
        LaunchScope ls(ctx, KI_KS);
        k_scatter<<<grid, TAIL_THREADS, 0, ctx->stream>>>(q);
    }
    CK(cudaGetLastError());
    return JPGB200_OK;
}

/* batch overlap: the previous frame's tail goes to the tail stream now, i.e. after everything the
 * main stream holds so far (the previous frame's KE and, when called after K1, this frame's K1 --
 * which fills the register files and is better left alone) */
static int run_pending_tail(jpgb200_ctx *ctx)
{
    if (!ctx->have_pending) return JPGB200_OK;
    ctx->have_pending = false;
    cudaStream_t main_stream = ctx->stream;
    CK(cudaEventRecord(ctx->ev_ke, main_stream));
    CK(cudaStreamWaitEvent(ctx->tail_stream, ctx->ev_ke, 0));
    ctx->stream = ctx->tail_stream;
    const int rc = run_tail(ctx, ctx->pending_q);
    ctx->stream = main_stream;
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev_ks[ctx->pending_parity], ctx->tail_stream));
    return JPGB200_OK;
}

/* KF + KS: chunk streams -> stuffed bytes at d_out (+ header, EOI, length word) */
static int launch_tail(jpgb200_ctx *ctx, const long long n[3], unsigned bit_phase, unsigned carry, int is_last_band,
                       int header_len, int write_eoi, uint8_t *d_out, size_t cap, unsigned long long *len_out)
{
    TailParams q;
    q.tmp = ctx->d_tmp;
    (void)n;
    q.slot_words = ctx->geo_slot_words;
    q.n_chunks = ctx->geo_chunks;
    q.cta_bits = ctx->d_cta_bits;
    q.super_bits = ctx->d_super;
    q.cta_ff = ctx->d_cta_ff;
    q.cta_tail = ctx->d_cta_tail;
    q.cta_base = ctx->d_cta_base;
    q.super_ff = ctx->d_super + ctx->n_super_bits;
    q.bit_phase = bit_phase;
    q.carry = carry;
    q.is_last_band = is_last_band;
    q.header = ctx->d_header;
    q.header_len = header_len;
    q.out = d_out;
    q.cap = (unsigned long long)cap;
    q.len_out = len_out;
    q.write_eoi = write_eoi;
    q.st = ctx->d_state;
    q.tmp_fstride = q.out_fstride = 0;
    q.cta_fstride = q.super_fstride = 0;
    q.mg = NULL;
    q.out_off = 0ull;
    q.mg_local = 0;
    if (ctx->overlap_now) { /* deferred: it will run beside the next frame's KE (run_pending_tail) */
        ctx->pending_q = q;
        ctx->pending_parity = ctx->ws_parity;
        ctx->have_pending = true;
        return JPGB200_OK;
    }
    return run_tail(ctx, q);
}

/* exchange the entropy workspaces in use with the alternate set */
static void swap_entropy_ws(jpgb200_ctx *ctx)
{
    jpgb200_ctx::EntWs cur = {ctx->d_tmp,      ctx->d_cta_bits, ctx->d_cta_ff,     ctx->d_cta_tail,   ctx->d_super,
                              ctx->d_cta_base, ctx->tmp_cap,    ctx->cta_cap,      ctx->n_super_bits, ctx->n_super_ff,
                              ctx->super_frames};
    ctx->d_tmp = ctx->alt.tmp;
    ctx->d_cta_bits = ctx->alt.cta_bits;
    ctx->d_cta_ff = ctx->alt.cta_ff;
    ctx->d_cta_tail = ctx->alt.cta_tail;
    ctx->d_super = ctx->alt.super;
    ctx->d_cta_base = ctx->alt.cta_base;
    ctx->tmp_cap = ctx->alt.tmp_cap;
    ctx->cta_cap = ctx->alt.cta_cap;
    ctx->n_super_bits = ctx->alt.n_super_bits;
    ctx->n_super_ff = ctx->alt.n_super_ff;
    ctx->super_frames = ctx->alt.super_frames;
    ctx->alt = cur;
    ctx->ws_parity ^= 1;
}

static int launch_pack(jpgb200_ctx *ctx, const int16_t *zz, int ratio, const long long n[3], uint8_t *d_out, size_t cap,
                       unsigned long long *len_out)
{
    (void)zz;
    (void)ratio;
    return launch_tail(ctx, n, 0u, 0u, 1, ctx->hdr_len, 1, d_out, cap, len_out); /* deferred inside a batch overlap */
}

#ifdef JPGB200_AB
/* K35: the whole entropy stage in one launch (entropy_fused.cuh) */
static int launch_fused(jpgb200_ctx *ctx, const int16_t *zz, int ratio, const long long n[3], uint8_t *d_out, size_t cap,
                        unsigned long long *len_out)
{
    if (!ctx->fused_attr_set) {
        CK(cudaFuncSetAttribute(k35_entropy<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, EfSmem<0>::BYTES));
        CK(cudaFuncSetAttribute(k35_entropy<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, EfSmem<1>::BYTES));
        ctx->fused_attr_set = true;
    }
    if (((ctx->epoch + 1) & CS_EPOCH_MASK) == 0ull) { /* epoch wrap: forget old states */
        CK(cudaMemsetAsync(ctx->d_bit_state, 0, ctx->chain_cap * 8, ctx->stream));
        CK(cudaMemsetAsync(ctx->d_ff_state, 0, ctx->chain_cap * 8, ctx->stream));
        ctx->epoch++;
    }
    EfParams q;
    q.zz = zz;
    q.coff1 = n[0];
    q.coff2 = n[0] + n[1];
    q.n_mcu = n[1];
    q.n_chunks = n_chunks(n);
    q.dc_pred[0] = q.dc_pred[1] = q.dc_pred[2] = 0;
    q.header_len = ctx->hdr_len;
    q.st = ctx->d_state;
    q.bit_state = ctx->d_bit_state;
    q.ff_state = ctx->d_ff_state;
    q.epoch = ++ctx->epoch;
    q.header = ctx->d_header;
    q.out = d_out;
    q.cap = (unsigned long long)cap;
    q.len_out = len_out;
    q.dbg_block_off = NULL;
    {
        LaunchScope ls(ctx, KI_K35);
        if (ratio == 0)
            k35_entropy<0><<<(unsigned)q.n_chunks, EcCfg<0>::NT, EfSmem<0>::BYTES, ctx->stream>>>(q, ctx->d_hbits, ctx->d_huff);
        else
            k35_entropy<1><<<(unsigned)q.n_chunks, EcCfg<1>::NT, EfSmem<1>::BYTES, ctx->stream>>>(q, ctx->d_hbits, ctx->d_huff);
    }
    CK(cudaGetLastError());
    return JPGB200_OK;
}
#else
static int launch_fused(jpgb200_ctx *ctx, const int16_t *, int, const long long *, uint8_t *, size_t, unsigned long long *)
{
    snprintf(ctx->err, sizeof ctx->err, "K35 is an A/B form: build with JPGB200_AB=1");
    return JPGB200_ERR_BAD_ARG;
}
#endif /* JPGB200_AB */

/* code tables and header of ctx->specs -> device (synchronises: the sources are stack buffers) */
static int upload_tables(jpgb200_ctx *ctx)
{
    HuffDev hh;
    HuffBits hb;
    HuffPack hp;
    build_huff(&ctx->specs, &hh);
    build_huff_bits(&hh, &hb);
    build_huff_pack(&hh, &hp);
    CK(cudaMemcpyAsync(ctx->d_huff, &hh, sizeof hh, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_hbits, &hb, sizeof hb, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_hpack, &hp, sizeof hp, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->hdr_w = -1; /* the cached header carries the old DHT segments */
    return JPGB200_OK;
}

static int device_histograms(jpgb200_ctx *ctx, int dc_mode, int sym_mode, int32_t *freq);

/* histograms of the coefficients just produced (KH) -> host-side T.81 K.2
 * -> DHT segments + code tables for this image.  One host synchronisation per image. */
static int optimise_tables(jpgb200_ctx *ctx, int w, int h, int quality, int ratio)
{
    int32_t freq[4 * 257];
    int rc = device_histograms(ctx, JPGB200_DC_SPEC_DIFF, JPGB200_SYM_SPEC, freq);
    if (rc) return rc;
    static const int src[4] = {0, 2, 1, 3}; /* histogram order: lum_DC, lum_AC, chrom_DC, chrom_AC */
    for (int t = 0; t < 4; t++) optimal_spec(freq + 257 * src[t], ctx->specs.nr[t], ctx->specs.vals[t], &ctx->specs.nvals[t]);
    if ((rc = upload_tables(ctx))) return rc;
    return ensure_tables(ctx, w, h, quality, ratio);
}

static int encode_core(jpgb200_ctx *ctx, const uint8_t *d_r, const uint8_t *d_g, const uint8_t *d_b, int w, int h,
                       int quality, int ratio, uint8_t *d_out, size_t cap, unsigned long long *len_out)
{
    long long n[3];
    int rc = validate(w, h, quality, ratio, n);
    if (rc) return rc;
    if ((rc = ensure_workspace(ctx, n, 1, w, h, ratio))) return rc;
    if ((rc = ensure_tables(ctx, w, h, quality, ratio))) return rc;
    ctx->w = w;
    ctx->h = h;
    ctx->q = quality;
    ctx->ratio = ratio;
    memcpy(ctx->n, n, sizeof n);
    if (d_out && k1e_usable(ctx, d_r, d_g, d_b, w, ratio, 0)) {
        /* the default: pixels -> chunk streams in one kernel, coefficients stay on chip */
        if ((rc = launch_k1e(ctx, d_r, d_g, d_b, w, h, ratio, n, ctx->k1e_tap ? ctx->d_zz : NULL, 0, NULL))) return rc;
        if (ctx->overlap_now && (rc = run_pending_tail(ctx))) return rc;
        return launch_pack(ctx, ctx->d_zz, ratio, n, d_out, cap, len_out);
    }
    if ((rc = launch_k1(ctx, d_r, d_g, d_b, w, h, ratio, n, 1, 0, ctx->d_zz, 0, 0))) return rc;
    if (ctx->overlap_now && (rc = run_pending_tail(ctx))) return rc; /* the previous frame's tail, beside this frame's KE */
    if (d_out) {
        ctx->have_bits = false;
        if (ctx->opt_huff) {
            if ((rc = optimise_tables(ctx, w, h, quality, ratio))) return rc;
            /* the histogram kernels ran after K1 cleared the two-level sums: still clear */
        }
        if (ctx->legacy_entropy || ratio == 2) { /* the single-launch K35 knows 4:4:4 and 4:2:2 only */
            if ((rc = launch_bits(ctx, ctx->d_zz, ratio, n))) return rc;
            return launch_pack(ctx, ctx->d_zz, ratio, n, d_out, cap, len_out);
        }
        return launch_fused(ctx, ctx->d_zz, ratio, n, d_out, cap, len_out);
    }
    /* parity tap: coefficients plus the staged K3/K4 bit offsets */
    return launch_bits(ctx, ctx->d_zz, ratio, n);
}

/* A group of G small frames in four launches (K1, KE, KF, KS over all of them: blockIdx.y = frame for the
 * entropy kernels, K1's tile index runs over the frames): what a 1080p frame needs is a sixteenth of an 8K
 * frame's work, and four launches per frame leave the GPU waiting for the next launch most of the time. */
/* pipe != NULL: the group's three stages go to the pipelined batch's three streams (K1 | KE | KF, KS), handed over
 * through the parity's events, and its coefficients to half `par` of a buffer of 2 * zz_group frames */
struct GroupPipe {
    int par, zz_group;
};

static int encode_group(jpgb200_ctx *ctx, const uint8_t *d_planes, int G, int w, int h, int quality, int ratio, uint8_t *d_out,
                        size_t out_stride, unsigned long long *d_lens, const GroupPipe *pipe = NULL)
{
    long long n[3];
    int rc = validate(w, h, quality, ratio, n);
    if (rc) return rc;
    if ((rc = ensure_workspace(ctx, n, pipe ? 2 * pipe->zz_group : G))) return rc;
    if ((rc = ensure_tables(ctx, w, h, quality, ratio))) return rc;
    ctx->w = w;
    ctx->h = h;
    ctx->q = quality;
    ctx->ratio = ratio;
    memcpy(ctx->n, n, sizeof n);
    const size_t npx = (size_t)w * h;
    const long long nb = n[0] + n[1] + n[2], nc = n_chunks(n), sw = slot_words(n);
    const int S = (int)(ctx->n_super_bits + ctx->n_super_ff);
    /* K1 over all frames; it also clears every frame's two-level sums */
    const size_t keep_b = ctx->n_super_bits, keep_f = ctx->n_super_ff;
    ctx->n_super_ff = (size_t)S * (size_t)G - keep_b; /* launch_k1 clears n_super_bits + n_super_ff words */
    int16_t *zz = ctx->d_zz + (pipe ? (size_t)pipe->par * (size_t)pipe->zz_group * (size_t)nb * 64 : 0);
    if (pipe) ctx->stream = ctx->k1_stream;
    rc = launch_k1(ctx, d_planes, d_planes + npx, d_planes + 2 * npx, w, h, ratio, n, G, (long long)(3 * npx), zz, nb * 64, 0);
    ctx->n_super_ff = keep_f;
    if (rc) return rc;
    if (pipe) {
        CK(cudaEventRecord(ctx->ev_k1[pipe->par], ctx->k1_stream));
        CK(cudaStreamWaitEvent(ctx->ke_stream, ctx->ev_k1[pipe->par], 0));
        ctx->stream = ctx->ke_stream;
    }
    ctx->geo_chunks = nc;
    ctx->geo_slot_words = sw;
    KeParams e;
    e.zz = zz;
    e.coff1 = n[0];
    e.coff2 = n[0] + n[1];
    e.n_mcu = n[1];
    e.dc_pred[0] = e.dc_pred[1] = e.dc_pred[2] = 0;
    e.tmp = ctx->d_tmp;
    e.slot_words = sw;
    e.cta_bits = ctx->d_cta_bits;
    e.cta_tail = ctx->d_cta_tail;
    e.super_bits = ctx->d_super;
    e.local_off = NULL;
    e.zz_fstride = nb * 64;
    e.tmp_fstride = nc * sw;
    e.cta_fstride = (int)nc;
    e.super_fstride = S;
    e.dc_pred_dev = NULL;
    {
        const int rc = launch_ke(ctx, e, ratio, dim3((unsigned)nc, (unsigned)G), ctx->stream, (unsigned long long)nb * (unsigned long long)G);
        if (rc) return rc;
    }
    CK(cudaGetLastError());
    if (pipe) {
        CK(cudaEventRecord(ctx->ev_ked[pipe->par], ctx->ke_stream));
        CK(cudaStreamWaitEvent(ctx->tail_stream, ctx->ev_ked[pipe->par], 0));
        ctx->stream = ctx->tail_stream;
    }
    TailParams q;
    q.tmp = ctx->d_tmp;
    q.slot_words = sw;
    q.n_chunks = nc;
    q.cta_bits = ctx->d_cta_bits;
    q.super_bits = ctx->d_super;
    q.cta_ff = ctx->d_cta_ff;
    q.cta_tail = ctx->d_cta_tail;
    q.cta_base = ctx->d_cta_base;
    q.super_ff = ctx->d_super + ctx->n_super_bits;
    q.bit_phase = 0u;
    q.carry = 0u;
    q.is_last_band = 1;
    q.header = ctx->d_header;
    q.header_len = ctx->hdr_len;
    q.out = d_out;
    q.cap = (unsigned long long)out_stride;
    q.len_out = d_lens;
    q.write_eoi = 1;
    q.st = NULL;
    q.tmp_fstride = nc * sw;
    q.out_fstride = (long long)out_stride;
    q.cta_fstride = (int)nc;
    q.super_fstride = S;
    q.mg = NULL;
    q.out_off = 0ull;
    q.mg_local = 0;
    const dim3 tgrid((unsigned)((nc + TAIL_WARPS - 1) / TAIL_WARPS), (unsigned)G);
    {
        LaunchScope ls(ctx, KI_KF);
        k_count_ff<<<tgrid, TAIL_THREADS, 0, ctx->stream>>>(q);
    }
    CK(cudaGetLastError());
    {
        LaunchScope ls(ctx, KI_KS);
        k_scatter<<<tgrid, TAIL_THREADS, 0, ctx->stream>>>(q);
    }
    CK(cudaGetLastError());
    if (pipe) CK(cudaEventRecord(ctx->ev_ks[pipe->par], ctx->tail_stream));
    ctx->have_bits = false;
    return JPGB200_OK;
}

/* Batch of small frames, pipelined like the large ones below: K1 of group g+1, KE of group g and the tails of group g-1
 * on three streams, two coefficient buffers (of one group each) and two sets of entropy workspaces by group parity. */
static int encode_groups_pipelined(jpgb200_ctx *ctx, const uint8_t *d_planes, int n_frames, int G, int w, int h, int quality, int ratio,
                                   uint8_t *d_out, size_t out_stride, unsigned long long *d_lens)
{
    long long n[3];
    int rc = validate(w, h, quality, ratio, n);
    if (rc) return rc;
    const size_t npx = (size_t)w * h;
    cudaStream_t user = ctx->stream;
    for (int k = 0; k < 2; k++) { /* both sets of workspaces exist before anything is enqueued (growing one synchronises) */
        swap_entropy_ws(ctx);
        if ((rc = ensure_workspace(ctx, n, 2 * G))) return rc;
    }
    if ((rc = ensure_tables(ctx, w, h, quality, ratio))) return rc;
    CK(cudaEventRecord(ctx->ev_start, user));
    CK(cudaStreamWaitEvent(ctx->k1_stream, ctx->ev_start, 0));
    CK(cudaStreamWaitEvent(ctx->ke_stream, ctx->ev_start, 0));
    CK(cudaStreamWaitEvent(ctx->tail_stream, ctx->ev_start, 0));
    ctx->in_batch = true;
    int g = 0;
    for (int f0 = 0; f0 < n_frames && rc == JPGB200_OK; f0 += G, g++) {
        swap_entropy_ws(ctx);
        GroupPipe pipe = {ctx->ws_parity, G};
        if (g >= 2) { /* this parity's coefficients were last read by KE of group g-2, its sums by that group's tails */
            CK(cudaStreamWaitEvent(ctx->k1_stream, ctx->ev_ked[pipe.par], 0));
            CK(cudaStreamWaitEvent(ctx->k1_stream, ctx->ev_ks[pipe.par], 0));
        }
        const int gg = n_frames - f0 < G ? n_frames - f0 : G;
        rc = encode_group(ctx, d_planes + (size_t)f0 * 3 * npx, gg, w, h, quality, ratio, d_out + (size_t)f0 * out_stride, out_stride,
                          d_lens ? d_lens + f0 : NULL, &pipe);
    }
    ctx->stream = user;
    ctx->in_batch = false;
    ctx->have_bits = false;
    cudaStreamWaitEvent(user, ctx->ev_ks[0], 0);
    cudaStreamWaitEvent(user, ctx->ev_ks[1], 0);
    cudaStreamWaitEvent(user, ctx->ev_ked[0], 0);
    cudaStreamWaitEvent(user, ctx->ev_ked[1], 0);
    return rc;
}

/* Batch, pipelined over three streams: K1 of frame f+1 (persistent CTAs) is in flight together with KE of frame f and
 * the tails of frame f-1.  Two coefficient buffers and two sets of entropy workspaces alternate with the frame parity.
 * K1's CTAs fill the register files, so the kernels take turns -- what the pipeline removes are the bubbles: KE's CTAs
 * start on every SM the moment K1's persistent CTAs retire from it and the next K1 moves in as KE drains.  Measured on
 * 8K frames: 151.6 us per frame = K1 + KE + 2 us with K1 and KE at the same stream priority and the tails above them
 * (jpgb200_create); with K1 held to ONE CTA per SM, so that KE's CTAs run beside it, both kernels slow down by more
 * than the overlap gives. */
static int encode_batch_pipelined(jpgb200_ctx *ctx, const uint8_t *d_planes, int n_frames, int w, int h, int quality, int ratio,
                                  uint8_t *d_out, size_t out_stride, unsigned long long *d_lens)
{
    long long n[3];
    int rc = validate(w, h, quality, ratio, n);
    if (rc) return rc;
    const size_t npx = (size_t)w * h;
    const long long nb = n[0] + n[1] + n[2];
    cudaStream_t user = ctx->stream;
    /* both parities' workspaces exist before anything is enqueued (growing one synchronises) */
    for (int k = 0; k < 2; k++) {
        swap_entropy_ws(ctx);
        if ((rc = ensure_workspace(ctx, n, 2))) return rc; /* d_zz: two frames */
    }
    if ((rc = ensure_tables(ctx, w, h, quality, ratio))) return rc;
    ctx->w = w;
    ctx->h = h;
    ctx->q = quality;
    ctx->ratio = ratio;
    memcpy(ctx->n, n, sizeof n);
    CK(cudaEventRecord(ctx->ev_start, user));
    CK(cudaStreamWaitEvent(ctx->k1_stream, ctx->ev_start, 0));
    CK(cudaStreamWaitEvent(ctx->ke_stream, ctx->ev_start, 0));
    CK(cudaStreamWaitEvent(ctx->tail_stream, ctx->ev_start, 0));
    ctx->in_batch = true;
    for (int f = 0; f < n_frames && rc == JPGB200_OK; f++) {
        swap_entropy_ws(ctx);
        const int par = ctx->ws_parity;
        int16_t *zz = ctx->d_zz + (size_t)par * (size_t)nb * 64;
        const uint8_t *pl = d_planes + (size_t)f * 3 * npx;
        /* K1(f): its coefficient buffer was last read by KE(f-2), its two-level sums by the tails of frame f-2 */
        if (f >= 2) {
            CK(cudaStreamWaitEvent(ctx->k1_stream, ctx->ev_ked[par], 0));
            CK(cudaStreamWaitEvent(ctx->k1_stream, ctx->ev_ks[par], 0));
        }
        ctx->stream = ctx->k1_stream;
        rc = launch_k1(ctx, pl, pl + npx, pl + 2 * npx, w, h, ratio, n, 1, 0, zz, 0, 0);
        if (rc) break;
        CK(cudaEventRecord(ctx->ev_k1[par], ctx->k1_stream));
        /* KE(f) */
        CK(cudaStreamWaitEvent(ctx->ke_stream, ctx->ev_k1[par], 0));
        ctx->stream = ctx->ke_stream;
        rc = launch_bits(ctx, zz, ratio, n);
        if (rc) break;
        CK(cudaEventRecord(ctx->ev_ked[par], ctx->ke_stream));
        /* KF, KS (f) */
        CK(cudaStreamWaitEvent(ctx->tail_stream, ctx->ev_ked[par], 0));
        ctx->stream = ctx->tail_stream;
        ctx->overlap_now = false;
        rc = launch_tail(ctx, n, 0u, 0u, 1, ctx->hdr_len, 1, d_out + (size_t)f * out_stride, out_stride, d_lens ? d_lens + f : ctx->d_len);
        if (rc) break;
        CK(cudaEventRecord(ctx->ev_ks[par], ctx->tail_stream));
    }
    ctx->stream = user;
    ctx->in_batch = false;
    ctx->have_bits = false;
    /* the caller's stream order is kept: it continues when the last tails are done */
    cudaStreamWaitEvent(user, ctx->ev_ks[0], 0);
    cudaStreamWaitEvent(user, ctx->ev_ks[1], 0);
    cudaStreamWaitEvent(user, ctx->ev_ked[0], 0);
    cudaStreamWaitEvent(user, ctx->ev_ked[1], 0);
    return rc;
}

/* frames per multi-frame launch group, 1 = frame-by-frame.  Small frames: about one 8K frame's worth of pixels per
 * launch (a 1080p frame is launch-bound on its own).  Large frames fill the GPU on their own, but every launch still
 * pays for filling and draining 148 SMs (K1's CTAs take a whole SM each: they start when the last chunk-encode CTA of
 * the frame before has left theirs) -- with eight 8K frames per launch the batch runs at 244 k Mpixel/s against 229 k
 * frame by frame (2: 240 k, 4: 243 k; profiles/README.md).  A batch is cut into at least three groups, which is what
 * the pipelined group path wants; with pipelining off the groups run one after the other (bench.py's per-kernel pass).
 * JPGB200_GROUP_LARGE=n overrides the group size of large frames (A/B). */
static int group_frames(const jpgb200_ctx *ctx, int n_frames, int w, int h)
{
    if (!ctx->group_small || !ctx->legacy_entropy || ctx->opt_huff || ctx->fused_k1e || n_frames < 2) return 1;
    const long long npx = (long long)w * h;
    if (npx > 4096ll * 2304ll) {
        static int g_large = -1;
        if (g_large < 0) {
            const char *e = getenv("JPGB200_GROUP_LARGE");
            g_large = e ? atoi(e) : 0;
            if (g_large < 0) g_large = 0;
        }
        int g = g_large ? g_large : 8;
        if (g > n_frames / 3) g = n_frames / 3;
        return g < 2 ? 1 : g;
    }
    long long g = (7680ll * 4320ll) / npx;
    if (g > 256) g = 256;
    if (g > n_frames) g = n_frames;
    return g < 2 ? 1 : (int)g;
}

/* ---- public encode entry points -------------------------------------------------------------- */

extern "C" int jpgb200_encode_device(jpgb200_ctx *ctx, const uint8_t *d_r, const uint8_t *d_g, const uint8_t *d_b, int w,
                                     int h, int quality, int ratio, uint8_t *d_out, size_t cap,
                                     unsigned long long *d_len)
{
    if (!ctx || !d_r || !d_g || !d_b || !d_out) return JPGB200_ERR_BAD_ARG;
    CK(cudaSetDevice(ctx->device));
    return encode_core(ctx, d_r, d_g, d_b, w, h, quality, ratio, d_out, cap, d_len ? d_len : ctx->d_len);
}

extern "C" int jpgb200_encode_batch_device(jpgb200_ctx *ctx, const uint8_t *d_planes, int n_frames, int w, int h,
                                           int quality, int ratio, uint8_t *d_out, size_t out_stride,
                                           unsigned long long *d_lens)
{
    if (!ctx || !d_planes || !d_out || n_frames < 0) return JPGB200_ERR_BAD_ARG;
    CK(cudaSetDevice(ctx->device));
    const size_t npx = (size_t)w * h;
    {
        /* small frames: groups of frames per launch (identical bytes, a fraction of the launches) */
        const int G = group_frames(ctx, n_frames, w, h);
        if (G > 1 && ctx->pipelined && (w % 16 == 0) && (n_frames + G - 1) / G >= 3)
            return encode_groups_pipelined(ctx, d_planes, n_frames, G, w, h, quality, ratio, d_out, out_stride, d_lens);
        if (G > 1) {
            int rc = JPGB200_OK;
            for (int f0 = 0; f0 < n_frames && rc == JPGB200_OK; f0 += G) {
                const int g = n_frames - f0 < G ? n_frames - f0 : G;
                rc = encode_group(ctx, d_planes + (size_t)f0 * 3 * npx, g, w, h, quality, ratio, d_out + (size_t)f0 * out_stride, out_stride,
                                  d_lens ? d_lens + f0 : NULL);
            }
            return rc;
        }
    }
    if (ctx->pipelined && ctx->legacy_entropy && !ctx->opt_huff && !ctx->fused_k1e && n_frames >= 3 && (w % 16 == 0))
        return encode_batch_pipelined(ctx, d_planes, n_frames, w, h, quality, ratio, d_out, out_stride, d_lens);
    /* Overlap (default): frame f's tail (KF, KS: small, latency-bound) runs on a second stream
     * beside frame f+1's KE (it is launched behind frame f+1's K1, which fills the register files
     * and is better left alone: launched right behind KE it slowed K1 from 77 to 98 us).  The entropy
     * workspaces alternate with the frame parity; the caller's stream order is kept (it waits for
     * the last tails before this call's work counts as done). */
    const bool overlap = ctx->overlap_on && ctx->legacy_entropy && !ctx->opt_huff && n_frames >= 2;
    int rc = JPGB200_OK;
    ctx->overlap_now = overlap;
    ctx->in_batch = true;
    for (int f = 0; f < n_frames && rc == JPGB200_OK; f++) {
        const uint8_t *pl = d_planes + (size_t)f * 3 * npx;
        if (overlap) {
            swap_entropy_ws(ctx);
            /* this parity's workspaces were last read by the tail of frame f-2 */
            if (f >= 2 && cudaStreamWaitEvent(ctx->stream, ctx->ev_ks[ctx->ws_parity], 0) != cudaSuccess) rc = JPGB200_ERR_CUDA;
        }
        if (rc == JPGB200_OK)
            rc = encode_core(ctx, pl, pl + npx, pl + 2 * npx, w, h, quality, ratio, d_out + (size_t)f * out_stride, out_stride,
                             d_lens ? d_lens + f : ctx->d_len);
    }
    ctx->in_batch = false;
    ctx->have_bits = false; /* no offset tap after a batch */
    if (overlap) {
        if (rc == JPGB200_OK) rc = run_pending_tail(ctx); /* the last frame's tail */
        ctx->have_pending = false;
        ctx->overlap_now = false;
        cudaStreamWaitEvent(ctx->stream, ctx->ev_ks[0], 0);
        cudaStreamWaitEvent(ctx->stream, ctx->ev_ks[1], 0);
    }
    return rc;
}

static int ensure_io(jpgb200_ctx *ctx, size_t npx, size_t cap, int slots)
{
    int rc;
    for (int s = 0; s < slots; s++) {
        size_t c = ctx->planes_cap;
        if ((rc = grow(ctx, &ctx->d_planes[s], &c, 3 * npx))) return rc;
        if (s == slots - 1) ctx->planes_cap = c;
    }
    for (int s = 0; s < slots; s++) {
        size_t c = ctx->out_cap;
        if ((rc = grow(ctx, &ctx->d_out[s], &c, cap))) return rc;
        if (s == slots - 1) ctx->out_cap = c;
    }
    return JPGB200_OK;
}

static int finish_host(jpgb200_ctx *ctx, uint8_t *d_out, uint8_t *out, size_t cap, size_t *len)
{
    unsigned long long l = 0;
    CK(cudaMemcpyAsync(&l, ctx->d_len, sizeof l, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (l == 0 || l > cap) return JPGB200_ERR_OUT_TOO_SMALL;
    CK(cudaMemcpyAsync(out, d_out, (size_t)l, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (len) *len = (size_t)l;
    return JPGB200_OK;
}

extern "C" int jpgb200_encode_planar(jpgb200_ctx *ctx, const uint8_t *r, const uint8_t *g, const uint8_t *b, int w, int h,
                                     int quality, int ratio, uint8_t *out, size_t cap, size_t *len)
{
    if (!ctx || !r || !g || !b || !out) return JPGB200_ERR_BAD_ARG;
    long long n[3];
    int rc = validate(w, h, quality, ratio, n);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    const size_t npx = (size_t)w * h;
    /* slots are re-grown together so both plane buffers always share planes_cap */
    if (ctx->planes_cap < 3 * npx || ctx->out_cap < cap || !ctx->d_planes[1] || !ctx->d_out[1]) {
        if ((rc = ensure_io(ctx, npx, cap, 2))) return rc;
    }
    uint8_t *dp = ctx->d_planes[0];
    CK(cudaMemcpyAsync(dp, r, npx, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(dp + npx, g, npx, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(dp + 2 * npx, b, npx, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = encode_core(ctx, dp, dp + npx, dp + 2 * npx, w, h, quality, ratio, ctx->d_out[0], cap, ctx->d_len))) return rc;
    return finish_host(ctx, ctx->d_out[0], out, cap, len);
}

/* any width/height: the planes are extended to whole MCUs by edge replication on the device
 * (T.81 A.2.4 leaves the padding samples to the encoder); SOF0 carries the true size */
extern "C" int jpgb200_encode_planar_any(jpgb200_ctx *ctx, const uint8_t *r, const uint8_t *g, const uint8_t *b, int w, int h,
                                         int quality, int ratio, uint8_t *out, size_t cap, size_t *len)
{
    if (!ctx || !r || !g || !b || !out) return JPGB200_ERR_BAD_ARG;
    if (w < 1 || h < 1 || w > 65535 || h > 65535 || ratio < 0 || ratio > 2) return JPGB200_ERR_BAD_DIMS;
    if (quality < 24 || quality > 97) return JPGB200_ERR_BAD_QUALITY;
    const int mw = ratio ? 16 : 8, mh = ratio == 2 ? 16 : 8;
    const int pw = (w + mw - 1) / mw * mw, ph = (h + mh - 1) / mh * mh;
    CK(cudaSetDevice(ctx->device));
    const size_t npx = (size_t)w * h, pnpx = (size_t)pw * ph;
    int rc;
    if (ctx->planes_cap < 3 * pnpx || ctx->out_cap < cap || !ctx->d_planes[1] || !ctx->d_out[1]) {
        if ((rc = ensure_io(ctx, pnpx, cap, 2))) return rc;
    }
    uint8_t *ds = ctx->d_planes[1], *dp = ctx->d_planes[0];
    CK(cudaMemcpyAsync(ds, r, npx, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ds + npx, g, npx, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ds + 2 * npx, b, npx, cudaMemcpyHostToDevice, ctx->stream));
    ks_pad_planes<<<(unsigned)ctx->sm_count * 8, 256, 0, ctx->stream>>>(ds, dp, w, h, pw, ph);
    ctx->launch_total++;
    CK(cudaGetLastError());
    ctx->sof_w = w;
    ctx->sof_h = h;
    ctx->hdr_w = -1;
    rc = encode_core(ctx, dp, dp + pnpx, dp + 2 * pnpx, pw, ph, quality, ratio, ctx->d_out[0], cap, ctx->d_len);
    ctx->sof_w = ctx->sof_h = 0;
    ctx->hdr_w = -1; /* the cached header carries the true size, not the plane size */
    if (rc) return rc;
    return finish_host(ctx, ctx->d_out[0], out, cap, len);
}

extern "C" int jpgb200_encode_rgb(jpgb200_ctx *ctx, const uint8_t *rgb, int w, int h, int quality, int ratio, uint8_t *out,
                                  size_t cap, size_t *len)
{
    if (!ctx || !rgb || !out) return JPGB200_ERR_BAD_ARG;
    long long n[3];
    int rc = validate(w, h, quality, ratio, n);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    const size_t npx = (size_t)w * h;
    if (ctx->planes_cap < 3 * npx || ctx->out_cap < cap || !ctx->d_planes[1] || !ctx->d_out[1]) {
        if ((rc = ensure_io(ctx, npx, cap, 2))) return rc;
    }
    /* interleaved upload into slot 1, de-interleave into slot 0 */
    CK(cudaMemcpyAsync(ctx->d_planes[1], rgb, 3 * npx, cudaMemcpyHostToDevice, ctx->stream));
    uint8_t *dp = ctx->d_planes[0];
    ks_deinterleave<<<(unsigned)ctx->sm_count * 8, 256, 0, ctx->stream>>>(ctx->d_planes[1], dp, dp + npx, dp + 2 * npx,
                                                                         (long long)npx);
    ctx->launch_total++;
    CK(cudaGetLastError());
    if ((rc = encode_core(ctx, dp, dp + npx, dp + 2 * npx, w, h, quality, ratio, ctx->d_out[0], cap, ctx->d_len))) return rc;
    return finish_host(ctx, ctx->d_out[0], out, cap, len);
}

/* Host batch: H2D of frame f+1 and D2H of frame f-1 overlap the kernels of frame f.
 * The stream length reaches the host through mapped pinned memory written by k_scatter, so the
 * only host wait per frame is on that frame's own completion event. */
extern "C" int jpgb200_encode_batch_host(jpgb200_ctx *ctx, const uint8_t *const *frames, int n_frames, int w, int h,
                                         int quality, int ratio, uint8_t *const *outs, size_t cap, size_t *lens)
{
    if (!ctx || !frames || !outs || n_frames < 0) return JPGB200_ERR_BAD_ARG;
    long long n[3];
    int rc = validate(w, h, quality, ratio, n);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    const size_t npx = (size_t)w * h;
    if (ctx->planes_cap < 3 * npx || ctx->out_cap < cap || !ctx->d_planes[1] || !ctx->d_out[1]) {
        if ((rc = ensure_io(ctx, npx, cap, 2))) return rc;
    }
    if ((rc = ensure_workspace(ctx, n, 1))) return rc;
    if ((rc = ensure_tables(ctx, w, h, quality, ratio))) return rc;
    unsigned long long *d_hlen = NULL;
    CK(cudaHostGetDevicePointer((void **)&d_hlen, ctx->h_len, 0));
    cudaEvent_t *ev_in = ctx->ev_in, *ev_done = ctx->ev_done, *ev_out = ctx->ev_out;
    auto drain = [&](int f) -> int {
        const int s = f & 1;
        CK(cudaEventSynchronize(ev_done[s]));
        const unsigned long long l = ctx->h_len[s];
        if (l == 0 || l > cap) return JPGB200_ERR_OUT_TOO_SMALL;
        CK(cudaMemcpyAsync(outs[f], ctx->d_out[s], (size_t)l, cudaMemcpyDeviceToHost, ctx->copy_out));
        CK(cudaEventRecord(ev_out[s], ctx->copy_out));
        if (lens) lens[f] = (size_t)l;
        return JPGB200_OK;
    };
    rc = JPGB200_OK;
    for (int f = 0; f < n_frames && rc == JPGB200_OK; f++) {
        const int s = f & 1;
        /* slot s is free once frame f-2's kernels (planes) and D2H (out) are done */
        if (f >= 2) {
            CK(cudaStreamWaitEvent(ctx->copy_in, ev_done[s], 0));
            CK(cudaStreamWaitEvent(ctx->stream, ev_out[s], 0));
        }
        CK(cudaMemcpyAsync(ctx->d_planes[s], frames[f], 3 * npx, cudaMemcpyHostToDevice, ctx->copy_in));
        CK(cudaEventRecord(ev_in[s], ctx->copy_in));
        CK(cudaStreamWaitEvent(ctx->stream, ev_in[s], 0));
        uint8_t *dp = ctx->d_planes[s];
        ctx->in_batch = true;
        rc = encode_core(ctx, dp, dp + npx, dp + 2 * npx, w, h, quality, ratio, ctx->d_out[s], cap, d_hlen + s);
        ctx->in_batch = false;
        ctx->have_bits = false;
        if (rc) break;
        CK(cudaEventRecord(ev_done[s], ctx->stream));
        if (f >= 1) rc = drain(f - 1);
    }
    if (rc == JPGB200_OK && n_frames >= 1) rc = drain(n_frames - 1);
    cudaStreamSynchronize(ctx->copy_out);
    cudaStreamSynchronize(ctx->stream);
    return rc;
}

/* The copies of jpgb200_encode_batch_host without its kernels: one cudaMemcpyAsync per frame and direction,
 * host->device on the copy-in stream, device->host on the copy-out stream, both directions at once.  What
 * this takes is the PCIe ceiling of the end-to-end number on this host (bench.py: e2e.pcie_ceiling). */
extern "C" int jpgb200_copy_probe(jpgb200_ctx *ctx, const uint8_t *const *frames, int n_frames, size_t in_bytes,
                                  uint8_t *const *outs, size_t out_bytes)
{
    if (!ctx || !frames || !outs || n_frames < 0) return JPGB200_ERR_BAD_ARG;
    CK(cudaSetDevice(ctx->device));
    int rc;
    if (ctx->planes_cap < in_bytes || ctx->out_cap < out_bytes || !ctx->d_planes[1] || !ctx->d_out[1]) {
        if ((rc = ensure_io(ctx, (in_bytes + 2) / 3, out_bytes > ctx->out_cap ? out_bytes : ctx->out_cap, 2))) return rc;
    }
    for (int f = 0; f < n_frames; f++) {
        CK(cudaMemcpyAsync(ctx->d_planes[f & 1], frames[f], in_bytes, cudaMemcpyHostToDevice, ctx->copy_in));
        CK(cudaMemcpyAsync(outs[f], ctx->d_out[f & 1], out_bytes, cudaMemcpyDeviceToHost, ctx->copy_out));
    }
    CK(cudaStreamSynchronize(ctx->copy_in));
    CK(cudaStreamSynchronize(ctx->copy_out));
    return JPGB200_OK;
}

/* ---- MCU-row partition of one large image across GPUs -----------------------------
 * Each rank runs the staged pipeline on its band; three tiny host-side exchanges (done by the
 * caller with NCCL allgathers) carry {band bits, last DCs}, {trailing partial byte} and
 * {stuffed length} between ranks.  A band OWNS the stream bytes whose last bit lies in its bit
 * range, so every stuffed segment is byte-aligned and segments simply concatenate. */
extern "C" int jpgb200_band_transform(jpgb200_ctx *ctx, const uint8_t *d_r, const uint8_t *d_g, const uint8_t *d_b,
                                      int w, int band_h, int has_halo, int quality, int ratio, int last_dc[3])
{
    if (!ctx || !d_r || !d_g || !d_b || !last_dc) return JPGB200_ERR_BAD_ARG;
    if (ratio == 2) return JPGB200_ERR_BAD_ARG; /* the MCU-row partition is built for 8-row MCUs */
    long long n[3];
    int rc = validate(w, band_h, quality, ratio, n);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    if ((rc = ensure_workspace(ctx, n, 1))) return rc;
    if (ctx->q_cached != quality) {
        int lum[64], chr[64];
        if (jpgb200_quant_tables(quality, lum, chr)) return JPGB200_ERR_BAD_QUALITY;
        make_quant_multipliers(lum, ctx->mul[0]);
        make_quant_multipliers(chr, ctx->mul[1]);
        ctx->q_cached = quality;
    }
    ctx->w = w;
    ctx->h = band_h;
    ctx->q = quality;
    ctx->ratio = ratio;
    memcpy(ctx->n, n, sizeof n);
    memcpy(ctx->band_n, n, sizeof n);
    ctx->band_ratio = ratio;
    ctx->have_bits = false;
    if ((rc = launch_k1(ctx, d_r, d_g, d_b, w, band_h, ratio, n, 1, 0, ctx->d_zz, 0, has_halo ? 1 : 0))) return rc;
    int16_t dc[3];
    CK(cudaMemcpyAsync(&dc[0], ctx->d_zz + (n[0] - 1) * 64, 2, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(&dc[1], ctx->d_zz + (n[0] + n[1] - 1) * 64, 2, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(&dc[2], ctx->d_zz + (n[0] + n[1] + n[2] - 1) * 64, 2, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int c = 0; c < 3; c++) last_dc[c] = dc[c];
    return JPGB200_OK;
}

extern "C" int jpgb200_band_bits(jpgb200_ctx *ctx, const int pred_dc[3], unsigned long long *band_bits)
{
    if (!ctx || !pred_dc || !band_bits) return JPGB200_ERR_BAD_ARG;
    CK(cudaSetDevice(ctx->device));
    for (int c = 0; c < 3; c++) ctx->band_pred[c] = pred_dc[c];
    int rc = launch_bits(ctx, ctx->d_zz, ctx->band_ratio, ctx->band_n, ctx->band_pred);
    if (rc) return rc;
    const size_t ns = (size_t)(n_chunks(ctx->band_n) / 64 + 1);
    std::vector<uint32_t> sup(ns);
    CK(cudaMemcpyAsync(sup.data(), ctx->d_super, ns * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    EncState st;
    st.total_bits = 0;
    for (size_t i = 0; i < ns; i++) st.total_bits += sup[i];
    ctx->band_bits = st.total_bits;
    *band_bits = st.total_bits;
    return JPGB200_OK;
}

extern "C" int jpgb200_band_pack(jpgb200_ctx *ctx, int bit_phase, int is_last, unsigned long long *owned_bytes,
                                 uint8_t *tail_byte)
{
    if (!ctx || !owned_bytes || !tail_byte || bit_phase < 0 || bit_phase > 7) return JPGB200_ERR_BAD_ARG;
    CK(cudaSetDevice(ctx->device));
    const unsigned long long end = (unsigned long long)bit_phase + ctx->band_bits;
    unsigned long long own = end >> 3;
    if (is_last && (end & 7ull)) own += 1; /* the 1-padded final byte */
    *tail_byte = 0;
    if (!is_last && (end & 7ull)) {
        /* the band's trailing partial byte, which the next band owns */
        k_band_tail<<<1, 1, 0, ctx->stream>>>(ctx->d_tmp, slot_words(ctx->band_n), n_chunks(ctx->band_n), ctx->d_cta_bits,
                                              (unsigned)(end & 7ull), ctx->d_len + 1);
        ctx->launch_total++;
        CK(cudaGetLastError());
        unsigned long long t = 0;
        CK(cudaMemcpyAsync(&t, ctx->d_len + 1, sizeof t, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        *tail_byte = (uint8_t)t;
    }
    ctx->band_phase = bit_phase;
    ctx->band_last = is_last ? 1 : 0;
    *owned_bytes = own;
    return JPGB200_OK;
}

extern "C" int jpgb200_band_finish(jpgb200_ctx *ctx, int carry_byte, uint8_t *d_segment, size_t cap,
                                   unsigned long long *segment_len)
{
    if (!ctx || !d_segment || !segment_len) return JPGB200_ERR_BAD_ARG;
    CK(cudaSetDevice(ctx->device));
    const unsigned long long end = (unsigned long long)ctx->band_phase + ctx->band_bits;
    unsigned long long own = end >> 3;
    if (ctx->band_last && (end & 7ull)) own += 1;
    /* K1 cleared the two-level sums at band_transform; KE used the bit half, KF uses the 0xFF half.
     * carry_byte holds the previous band's trailing bits left-aligned: bring them to the low end. */
    const unsigned carry = ctx->band_phase ? ((unsigned)(carry_byte & 0xff) >> (8 - ctx->band_phase)) : 0u;
    int rc = launch_tail(ctx, ctx->band_n, (unsigned)ctx->band_phase, carry, ctx->band_last, 0, 0, d_segment, cap, ctx->d_len);
    if (rc) return rc;
    unsigned long long l = 0;
    CK(cudaMemcpyAsync(&l, ctx->d_len, sizeof l, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (own && l == 0) return JPGB200_ERR_OUT_TOO_SMALL;
    *segment_len = l;
    return JPGB200_OK;
}


/* ---- one image across the GPUs of a box: C entry over NCCL + peer memory (mgpu.cuh) ----------------------
 * NCCL is resolved at run time (dlopen of libnccl.so.2: inside a PyTorch process that is the library torch
 * already loaded, in a plain C host the system's), so libjpgb200.so itself does not depend on it. */
namespace {
typedef int (*nccl_get_unique_id_t)(void *);
typedef int (*nccl_comm_init_rank_t)(void **, int, ncclUniqueIdBytes, int);
typedef int (*nccl_comm_destroy_t)(void *);
typedef int (*nccl_all_gather_t)(const void *, void *, size_t, int, void *, cudaStream_t);
typedef const char *(*nccl_get_error_string_t)(int);
struct NcclApi {
    void *lib;
    nccl_get_unique_id_t get_unique_id;
    nccl_comm_init_rank_t comm_init_rank;
    nccl_comm_destroy_t comm_destroy;
    nccl_all_gather_t all_gather;
    nccl_get_error_string_t error_string;
};
NcclApi *nccl_api()
{
    static NcclApi api = {};
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (h) {
            api.lib = h;
            api.get_unique_id = (nccl_get_unique_id_t)dlsym(h, "ncclGetUniqueId");
            api.comm_init_rank = (nccl_comm_init_rank_t)dlsym(h, "ncclCommInitRank");
            api.comm_destroy = (nccl_comm_destroy_t)dlsym(h, "ncclCommDestroy");
            api.all_gather = (nccl_all_gather_t)dlsym(h, "ncclAllGather");
            api.error_string = (nccl_get_error_string_t)dlsym(h, "ncclGetErrorString");
            if (!api.get_unique_id || !api.comm_init_rank || !api.comm_destroy || !api.all_gather) api.lib = NULL;
        }
    }
    return api.lib ? &api : NULL;
}
}  // namespace

extern "C" int jpgb200_nccl_unique_id(uint8_t id[128])
{
    NcclApi *a = nccl_api();
    if (!a || !id) return JPGB200_ERR_NCCL;
    return a->get_unique_id(id) == 0 ? JPGB200_OK : JPGB200_ERR_NCCL;
}

extern "C" int jpgb200_nccl_init(void **comm, int world, int rank, const uint8_t id[128], int device)
{
    NcclApi *a = nccl_api();
    if (!a || !comm || !id) return JPGB200_ERR_NCCL;
    if (cudaSetDevice(device) != cudaSuccess) return JPGB200_ERR_NO_DEVICE;
    ncclUniqueIdBytes u;
    memcpy(u.internal, id, 128);
    return a->comm_init_rank(comm, world, u, rank) == 0 ? JPGB200_OK : JPGB200_ERR_NCCL;
}

extern "C" int jpgb200_nccl_destroy(void *comm)
{
    NcclApi *a = nccl_api();
    if (!a) return JPGB200_ERR_NCCL;
    return (!comm || a->comm_destroy(comm) == 0) ? JPGB200_OK : JPGB200_ERR_NCCL;
}

struct jpgb200_mgpu {
    jpgb200_ctx *ctx;
    void *comm;
    int rank, world;
    size_t out_cap;
    uint8_t *out_local; /* rank 0: the output buffer */
    uint8_t *out_peer;  /* every rank: rank 0's buffer as seen from this GPU */
    uint8_t *seg;       /* ranks > 0: the band's segment before it is pushed to rank 0 */
    size_t seg_cap;
    bool opened;
    /* device exchange buffers */
    MgMsg1 *msg1, *all1;
    MgMsg2 *msg2, *all2;
    MgState *state;
    unsigned long long *all_bits, *total_len, *done_send, *done_recv;
    cudaIpcMemHandle_t *handles; /* [world] device */
};

#define NK(call)                                                                                                              \
    do {                                                                                                                      \
        int r_ = (call);                                                                                                      \
        if (r_ != 0) {                                                                                                        \
            snprintf(ctx->err, sizeof ctx->err, "%s failed: %s (%s:%d)", #call, api->error_string ? api->error_string(r_) : "nccl", \
                     __FILE__, __LINE__);                                                                                     \
            return JPGB200_ERR_NCCL;                                                                                          \
        }                                                                                                                     \
    } while (0)

extern "C" int jpgb200_mgpu_create(jpgb200_mgpu **out, jpgb200_ctx *ctx, void *nccl_comm, int rank, int world, size_t out_cap)
{
    if (!out || !ctx || !nccl_comm || world < 1 || world > MG_MAX_RANKS || rank < 0 || rank >= world) return JPGB200_ERR_BAD_ARG;
    NcclApi *api = nccl_api();
    if (!api) {
        snprintf(ctx->err, sizeof ctx->err, "libnccl.so.2 not found");
        return JPGB200_ERR_NCCL;
    }
    CK(cudaSetDevice(ctx->device));
    jpgb200_mgpu *mg = new jpgb200_mgpu();
    memset(mg, 0, sizeof *mg);
    mg->ctx = ctx;
    mg->comm = nccl_comm;
    mg->rank = rank;
    mg->world = world;
    mg->out_cap = out_cap;
    *out = mg;
    CK(cudaMalloc(&mg->msg1, sizeof(MgMsg1)));
    CK(cudaMalloc(&mg->all1, sizeof(MgMsg1) * world));
    CK(cudaMalloc(&mg->msg2, sizeof(MgMsg2)));
    CK(cudaMalloc(&mg->all2, sizeof(MgMsg2) * world));
    CK(cudaMalloc(&mg->state, sizeof(MgState)));
    CK(cudaMemset(mg->state, 0, sizeof(MgState)));
    CK(cudaMalloc(&mg->all_bits, 8 * world));
    CK(cudaMalloc(&mg->total_len, 8));
    CK(cudaMalloc(&mg->done_send, 8));
    CK(cudaMemset(mg->done_send, 0, 8));
    CK(cudaMalloc(&mg->done_recv, 8 * world));
    CK(cudaMalloc(&mg->handles, sizeof(cudaIpcMemHandle_t) * world));
    /* rank 0 owns the output; its IPC handle travels to everyone through the communicator itself */
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof mine);
    if (rank == 0) {
        CK(cudaMalloc(&mg->out_local, out_cap));
        CK(cudaIpcGetMemHandle(&mine, mg->out_local));
    }
    cudaIpcMemHandle_t *d_mine = NULL;
    CK(cudaMalloc(&d_mine, sizeof mine));
    CK(cudaMemcpyAsync(d_mine, &mine, sizeof mine, cudaMemcpyHostToDevice, ctx->stream));
    NK(api->all_gather(d_mine, mg->handles, sizeof mine, 1 /* ncclUint8 */, mg->comm, ctx->stream));
    cudaIpcMemHandle_t h0;
    CK(cudaMemcpyAsync(&h0, mg->handles, sizeof h0, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaFree(d_mine));
    if (rank == 0) {
        mg->out_peer = mg->out_local;
    } else {
        void *p = NULL;
        CK(cudaIpcOpenMemHandle(&p, h0, cudaIpcMemLazyEnablePeerAccess));
        mg->out_peer = (uint8_t *)p;
        mg->opened = true;
    }
    return JPGB200_OK;
}

extern "C" int jpgb200_mgpu_destroy(jpgb200_mgpu *mg)
{
    if (!mg) return JPGB200_OK;
    jpgb200_ctx *ctx = mg->ctx;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (mg->opened) cudaIpcCloseMemHandle(mg->out_peer);
    void *ptrs[] = {mg->msg1, mg->all1, mg->msg2, mg->all2, mg->state, mg->all_bits, mg->total_len, mg->done_send, mg->done_recv,
                    mg->handles, mg->out_local, mg->seg};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    delete mg;
    return JPGB200_OK;
}

extern "C" uint8_t *jpgb200_mgpu_output(jpgb200_mgpu *mg) { return mg ? mg->out_local : NULL; }

/* Encode this rank's band (d_r/d_g/d_b: its first row; the row above must be readable when rank > 0) and
 * stitch the stream in rank 0's buffer.  Everything is enqueued on the context's stream: K1, KE, exchange 1,
 * the first chunk again, KF, exchange 2, KS through the peer pointer, completion allgather.  Nothing is
 * synchronised with the host; jpgb200_mgpu_result waits and reads the length. */
extern "C" int jpgb200_mgpu_encode(jpgb200_mgpu *mg, const uint8_t *d_r, const uint8_t *d_g, const uint8_t *d_b, int w, int band_h,
                                   int image_h, int quality, int ratio)
{
    if (!mg || !d_r || !d_g || !d_b) return JPGB200_ERR_BAD_ARG;
    if (ratio != 0 && ratio != 1) return JPGB200_ERR_BAD_ARG; /* the MCU-row partition is built for 8-row MCUs */
    jpgb200_ctx *ctx = mg->ctx;
    NcclApi *api = nccl_api();
    if (!api) return JPGB200_ERR_NCCL;
    long long n[3], n_img[3];
    int rc = validate(w, band_h, quality, ratio, n);
    if (rc) return rc;
    if ((rc = validate(w, image_h, quality, ratio, n_img))) return rc;
    CK(cudaSetDevice(ctx->device));
    if ((rc = ensure_workspace(ctx, n, 1))) return rc;
    if ((rc = ensure_tables(ctx, w, image_h, quality, ratio))) return rc; /* the header describes the WHOLE image */
    ctx->w = w;
    ctx->h = band_h;
    ctx->q = quality;
    ctx->ratio = ratio;
    memcpy(ctx->n, n, sizeof n);
    const int rank = mg->rank, world = mg->world, is_last = rank == world - 1;
    cudaStream_t st = ctx->stream;
    const bool was_batch = ctx->in_batch;
    ctx->in_batch = true; /* no per-block offset tap */
    /* K1 + KE with a zero predictor */
    if ((rc = launch_k1(ctx, d_r, d_g, d_b, w, band_h, ratio, n, 1, 0, ctx->d_zz, 0, rank > 0 ? 1 : 0))) return rc;
    rc = launch_bits(ctx, ctx->d_zz, ratio, n);
    if (rc) {
        ctx->in_batch = was_batch;
        return rc;
    }
    const long long nc = n_chunks(n), sw = slot_words(n);
    const int nsb = (int)(nc / 64 + 1);
    k_mg_pack1<<<1, 32, 0, st>>>(ctx->d_zz, n[0], n[1], n[2], ctx->d_super, nsb, mg->msg1);
    ctx->launch_total++;
    NK(api->all_gather(mg->msg1, mg->all1, sizeof(MgMsg1), 1, mg->comm, st)); /* exchange 1 */
    k_mg_fix0<<<1, 32, 0, st>>>(mg->all1, rank, world, ctx->d_hpack, mg->state, ctx->d_super, ctx->d_cta_bits, mg->all_bits);
    ctx->launch_total++;
    {
        /* the band's first chunk again, now with the true predictor (device memory) */
        KeParams e;
        e.zz = ctx->d_zz;
        e.coff1 = n[0];
        e.coff2 = n[0] + n[1];
        e.n_mcu = n[1];
        e.dc_pred[0] = e.dc_pred[1] = e.dc_pred[2] = 0;
        e.tmp = ctx->d_tmp;
        e.slot_words = sw;
        e.cta_bits = ctx->d_cta_bits;
        e.cta_tail = ctx->d_cta_tail;
        e.super_bits = ctx->d_super;
        e.local_off = NULL;
        e.zz_fstride = e.tmp_fstride = 0;
        e.cta_fstride = e.super_fstride = 0;
        e.dc_pred_dev = mg->state->pred;
        e.recode_count = NULL;
        LaunchScope ls(ctx, KI_KE);
        if (ratio == 0)
            k_encode_chunks<0, false><<<1, KeCfg<0>::NT, KeSmem<0>::BYTES, st>>>(e, ctx->d_hpack);
        else
            k_encode_chunks<1, false><<<1, KeCfg<1>::NT, KeSmem<1>::BYTES, st>>>(e, ctx->d_hpack);
    }
    ctx->in_batch = was_batch;
    CK(cudaGetLastError());
    TailParams q;
    q.tmp = ctx->d_tmp;
    q.slot_words = sw;
    q.n_chunks = nc;
    q.cta_bits = ctx->d_cta_bits;
    q.super_bits = ctx->d_super;
    q.cta_ff = ctx->d_cta_ff;
    q.cta_tail = ctx->d_cta_tail;
    q.cta_base = ctx->d_cta_base;
    q.super_ff = ctx->d_super + ctx->n_super_bits;
    q.bit_phase = 0u;
    q.carry = 0u;
    q.is_last_band = is_last;
    q.header = ctx->d_header;
    q.header_len = rank == 0 ? ctx->hdr_len : 0;
    if (rank > 0) {
        /* worst case of a band's segment: every block at its longest, every byte stuffed */
        const size_t need = (size_t)(n[0] + n[1] + n[2]) * 1658 / 8 * 2 + 64;
        const size_t want = need < mg->out_cap + 64 ? need : mg->out_cap + 64;
        if (mg->seg_cap < want) {
            CK(cudaStreamSynchronize(st));
            if (mg->seg) CK(cudaFree(mg->seg));
            mg->seg = NULL;
            CK(cudaMalloc(&mg->seg, want));
            mg->seg_cap = want;
        }
    }
    q.out = rank == 0 ? mg->out_local : mg->seg;
    q.cap = (unsigned long long)(rank == 0 ? mg->out_cap : mg->seg_cap);
    q.len_out = NULL;
    q.write_eoi = is_last;
    q.st = NULL;
    q.tmp_fstride = q.out_fstride = 0;
    q.cta_fstride = q.super_fstride = 0;
    q.mg = mg->state;
    q.out_off = 0ull;
    q.mg_local = rank > 0 ? 1 : 0;
    const unsigned tgrid = (unsigned)((nc + TAIL_WARPS - 1) / TAIL_WARPS);
    {
        LaunchScope ls(ctx, KI_KF);
        k_count_ff<<<tgrid, TAIL_THREADS, 0, st>>>(q);
    }
    CK(cudaGetLastError());
    k_mg_pack2<<<1, 32, 0, st>>>(mg->state, is_last, ctx->d_tmp, sw, nc, ctx->d_cta_bits, ctx->d_super + ctx->n_super_bits, nsb, mg->msg2);
    ctx->launch_total++;
    NK(api->all_gather(mg->msg2, mg->all2, sizeof(MgMsg2), 1, mg->comm, st)); /* exchange 2 */
    k_mg_fix1<<<1, 32, 0, st>>>(mg->all2, mg->all_bits, rank, world, ctx->hdr_len, mg->state, ctx->d_cta_ff,
                                ctx->d_super + ctx->n_super_bits, mg->total_len);
    ctx->launch_total++;
    {
        LaunchScope ls(ctx, KI_KS);
        k_scatter<<<tgrid, TAIL_THREADS, 0, st>>>(q); /* rank 0: in place; others: a local segment at the destination's 16-byte phase */
    }
    CK(cudaGetLastError());
    if (rank > 0) {
        /* the segment goes straight into rank 0's buffer through the peer pointer, 16 bytes per store */
        k_mg_push<<<(unsigned)ctx->sm_count * 4, 256, 0, st>>>(mg->state, mg->seg, mg->out_peer, (unsigned long long)mg->out_cap);
        ctx->launch_total++;
        CK(cudaGetLastError());
    }
    /* completion: when this allgather has run on rank 0's stream, every rank's KS has finished */
    NK(api->all_gather(mg->done_send, mg->done_recv, 8, 1, mg->comm, st));
    return JPGB200_OK;
}

/* waits for the stream; rank 0 gets the length of the stitched stream (0: the output buffer was too small) */
extern "C" int jpgb200_mgpu_result(jpgb200_mgpu *mg, unsigned long long *total_len)
{
    if (!mg) return JPGB200_ERR_BAD_ARG;
    jpgb200_ctx *ctx = mg->ctx;
    CK(cudaSetDevice(ctx->device));
    unsigned long long l = 0;
    CK(cudaMemcpyAsync(&l, mg->total_len, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (total_len) *total_len = l <= mg->out_cap ? l : 0ull;
    return l <= mg->out_cap ? JPGB200_OK : JPGB200_ERR_OUT_TOO_SMALL;
}

/* ---- parity taps -------------------------------------------------------------------------------- */

extern "C" int jpgb200_coefficients(jpgb200_ctx *ctx, const uint8_t *r, const uint8_t *g, const uint8_t *b, int w, int h,
                                    int quality, int ratio, int16_t *zz_out)
{
    if (!ctx || !r || !g || !b || !zz_out) return JPGB200_ERR_BAD_ARG;
    long long n[3];
    int rc = validate(w, h, quality, ratio, n);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    const size_t npx = (size_t)w * h;
    if (ctx->planes_cap < 3 * npx || !ctx->d_planes[1] || !ctx->d_out[1]) {
        if ((rc = ensure_io(ctx, npx, ctx->out_cap ? ctx->out_cap : 4096, 2))) return rc;
    }
    uint8_t *dp = ctx->d_planes[0];
    CK(cudaMemcpyAsync(dp, r, npx, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(dp + npx, g, npx, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(dp + 2 * npx, b, npx, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = encode_core(ctx, dp, dp + npx, dp + 2 * npx, w, h, quality, ratio, NULL, 0, NULL))) return rc;
    CK(cudaMemcpyAsync(zz_out, ctx->d_zz, (size_t)(n[0] + n[1] + n[2]) * 128, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return JPGB200_OK;
}

extern "C" int jpgb200_last_coefficients(jpgb200_ctx *ctx, int16_t *zz_out)
{
    if (!ctx || !zz_out || !ctx->d_zz || ctx->n[0] == 0) return JPGB200_ERR_BAD_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(zz_out, ctx->d_zz, (size_t)(ctx->n[0] + ctx->n[1] + ctx->n[2]) * 128, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return JPGB200_OK;
}

static int run_dc_sequence(jpgb200_ctx *ctx, int mode)
{
    const long long nb = ctx->n[0] + ctx->n[1] + ctx->n[2];
    int rc;
    if ((rc = grow(ctx, &ctx->d_dc, &ctx->dc_cap, (size_t)nb * 4))) return rc;
    long long off = 0;
    for (int c = 0; c < 3; c++) {
        const long long nc = ctx->n[c];
        const unsigned tiles = (unsigned)((nc + K2_TILE - 1) / K2_TILE);
        if ((rc = grow(ctx, &ctx->d_tile_sum, &ctx->tile_sum_cap, (size_t)tiles * 4))) return rc;
        const int16_t *z = ctx->d_zz + off * 64;
        if (mode == JPGB200_DC_REF_RECURRENCE) {
            k2_tile_sums<<<tiles, K2_THREADS, 0, ctx->stream>>>(z, nc, ctx->d_tile_sum);
            k2_scan_tiles<<<1, 1024, 0, ctx->stream>>>(ctx->d_tile_sum, (long long)tiles);
            ctx->launch_total += 2;
        }
        k2_apply<<<tiles, K2_THREADS, 0, ctx->stream>>>(z, nc, ctx->d_tile_sum, mode, 0, ctx->d_dc + off);
        ctx->launch_total++;
        CK(cudaGetLastError());
        off += nc;
    }
    return JPGB200_OK;
}

extern "C" int jpgb200_dc_sequence(jpgb200_ctx *ctx, int mode, int32_t *d_out)
{
    if (!ctx || !d_out || (mode != 0 && mode != 1)) return JPGB200_ERR_BAD_ARG;
    if (!ctx->d_zz || ctx->n[0] == 0) return JPGB200_ERR_BAD_ARG;
    CK(cudaSetDevice(ctx->device));
    int rc = run_dc_sequence(ctx, mode);
    if (rc) return rc;
    const long long nb = ctx->n[0] + ctx->n[1] + ctx->n[2];
    CK(cudaMemcpyAsync(d_out, ctx->d_dc, (size_t)nb * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return JPGB200_OK;
}

extern "C" int jpgb200_histograms(jpgb200_ctx *ctx, int dc_mode, int sym_mode, int32_t *freq)
{
    if (!ctx || !freq || (dc_mode != 0 && dc_mode != 1) || (sym_mode != 0 && sym_mode != 1)) return JPGB200_ERR_BAD_ARG;
    if (!ctx->d_zz || ctx->n[0] == 0) return JPGB200_ERR_BAD_ARG;
    CK(cudaSetDevice(ctx->device));
    return device_histograms(ctx, dc_mode, sym_mode, freq);
}

static int device_histograms(jpgb200_ctx *ctx, int dc_mode, int sym_mode, int32_t *freq)
{
    int rc = run_dc_sequence(ctx, dc_mode);
    if (rc) return rc;
    CK(cudaMemsetAsync(ctx->d_freq, 0, 4 * 257 * sizeof(int), ctx->stream));
    long long off = 0;
    for (int c = 0; c < 3; c++) {
        const long long nc = ctx->n[c];
        kh_histograms<<<(unsigned)((nc + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_zz + off * 64, ctx->d_dc + off, nc,
                                                                            c != 0, sym_mode, ctx->d_freq);
        ctx->launch_total++;
        CK(cudaGetLastError());
        off += nc;
    }
    CK(cudaMemcpyAsync(freq, ctx->d_freq, 4 * 257 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int t = 0; t < 4; t++) freq[t * 257 + 256] = 1; /* the reserved code point, src/huffman.c:54 */
    return JPGB200_OK;
}

extern "C" int jpgb200_block_bit_offsets(jpgb200_ctx *ctx, unsigned long long *offsets)
{
    if (!ctx || !offsets || !ctx->have_bits) return JPGB200_ERR_BAD_ARG;
    CK(cudaSetDevice(ctx->device));
    const int bpm = ctx->ratio == 2 ? 6 : (ctx->ratio ? 4 : 3);
    const long long nb = ctx->n[1] * bpm, nc = n_chunks(ctx->n);
    std::vector<uint32_t> loc((size_t)nb), cbits((size_t)nc);
    CK(cudaMemcpyAsync(loc.data(), ctx->d_local_off, (size_t)nb * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(cbits.data(), ctx->d_cta_bits, (size_t)nc * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    /* tap only: the device-side prefix of the chunk totals is computed inside K5 */
    std::vector<unsigned long long> base((size_t)nc + 1);
    base[0] = 0;
    for (long long c = 0; c < nc; c++) base[(size_t)c + 1] = base[(size_t)c] + cbits[(size_t)c];
    const long long per = (long long)EC_MCUS * bpm;
    for (long long s = 0; s < nb; s++) offsets[s] = base[(size_t)(s / per)] + loc[(size_t)s];
    offsets[nb] = base[(size_t)nc];
    return JPGB200_OK;
}

/* ---- staged FP64 block API ------------------------------------------------------------------------ */

static int stage_buf(jpgb200_ctx *ctx, size_t bytes, void **p)
{
    cudaError_t e = cudaMalloc(p, bytes ? bytes : 8);
    if (e != cudaSuccess) {
        snprintf(ctx->err, sizeof ctx->err, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        return JPGB200_ERR_NOMEM;
    }
    return JPGB200_OK;
}

extern "C" int jpgb200_stage_dct(jpgb200_ctx *ctx, double *blocks, long long n)
{
    if (!ctx || !blocks || n <= 0) return JPGB200_ERR_BAD_ARG;
    CK(cudaSetDevice(ctx->device));
    void *d = NULL;
    int rc = stage_buf(ctx, (size_t)n * 512, &d);
    if (rc) return rc;
    CK(cudaMemcpyAsync(d, blocks, (size_t)n * 512, cudaMemcpyHostToDevice, ctx->stream));
    kf_dct_blocks<<<(unsigned)n, 64, 0, ctx->stream>>>((double *)d, n);
    ctx->launch_total++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(blocks, d, (size_t)n * 512, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaFree(d));
    return JPGB200_OK;
}

extern "C" int jpgb200_stage_quantise(jpgb200_ctx *ctx, double *blocks, long long n, const int table[64])
{
    if (!ctx || !blocks || !table || n <= 0) return JPGB200_ERR_BAD_ARG;
    CK(cudaSetDevice(ctx->device));
    void *d = NULL, *t = NULL;
    int rc = stage_buf(ctx, (size_t)n * 512, &d);
    if (rc) return rc;
    if ((rc = stage_buf(ctx, 256, &t))) return rc;
    CK(cudaMemcpyAsync(d, blocks, (size_t)n * 512, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(t, table, 256, cudaMemcpyHostToDevice, ctx->stream));
    kf_quantise_blocks<<<(unsigned)n, 64, 0, ctx->stream>>>((double *)d, n, (const int *)t);
    ctx->launch_total++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(blocks, d, (size_t)n * 512, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaFree(d));
    CK(cudaFree(t));
    return JPGB200_OK;
}

extern "C" int jpgb200_stage_zigzag(jpgb200_ctx *ctx, const double *blocks, long long n, int *zz)
{
    if (!ctx || !blocks || !zz || n <= 0) return JPGB200_ERR_BAD_ARG;
    CK(cudaSetDevice(ctx->device));
    void *d = NULL, *z = NULL;
    int rc = stage_buf(ctx, (size_t)n * 512, &d);
    if (rc) return rc;
    if ((rc = stage_buf(ctx, (size_t)n * 256, &z))) return rc;
    CK(cudaMemcpyAsync(d, blocks, (size_t)n * 512, cudaMemcpyHostToDevice, ctx->stream));
    kf_zigzag_blocks<<<(unsigned)n, 64, 0, ctx->stream>>>((const double *)d, n, ctx->d_scan_order, (int *)z);
    ctx->launch_total++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(zz, z, (size_t)n * 256, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaFree(d));
    CK(cudaFree(z));
    return JPGB200_OK;
}

/* RAII-free scratch helper for the staged entry points (they are not on the hot path) */
struct Scratch {
    jpgb200_ctx *ctx;
    std::vector<void *> ptrs;
    explicit Scratch(jpgb200_ctx *c) : ctx(c) {}
    ~Scratch()
    {
        for (void *p : ptrs) cudaFree(p);
    }
    int get(size_t bytes, void **p)
    {
        int rc = stage_buf(ctx, bytes, p);
        if (rc == JPGB200_OK) ptrs.push_back(*p);
        return rc;
    }
};

extern "C" int jpgb200_stage_preprocess(jpgb200_ctx *ctx, const uint8_t *r, const uint8_t *g, const uint8_t *b, int w,
                                        int h, double *Y, double *Cb, double *Cr)
{
    if (!ctx || !r || !g || !b || !Y || !Cb || !Cr) return JPGB200_ERR_BAD_ARG;
    if (w <= 0 || h <= 0 || (w % 8) || (h % 8)) return JPGB200_ERR_BAD_DIMS;
    CK(cudaSetDevice(ctx->device));
    const size_t npx = (size_t)w * h;
    Scratch sc(ctx);
    void *dp = NULL, *dd = NULL, *doob = NULL;
    int rc;
    if ((rc = sc.get(3 * npx, &dp)) || (rc = sc.get(3 * npx * 8, &dd)) || (rc = sc.get(32, &doob))) return rc;
    uint8_t *pl = (uint8_t *)dp;
    double *blk = (double *)dd;
    CK(cudaMemcpyAsync(pl, r, npx, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(pl + npx, g, npx, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(pl + 2 * npx, b, npx, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(doob, ctx->oob, 24, cudaMemcpyHostToDevice, ctx->stream));
    kf_preprocess_blocks<<<(unsigned)((npx + 255) / 256), 256, 0, ctx->stream>>>(pl, pl + npx, pl + 2 * npx, w, h,
                                                                               (const uint8_t *)doob, blk, blk + npx,
                                                                               blk + 2 * npx);
    ctx->launch_total++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(Y, blk, npx * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(Cb, blk + npx, npx * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(Cr, blk + 2 * npx, npx * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return JPGB200_OK;
}

extern "C" int jpgb200_stage_subsample422(jpgb200_ctx *ctx, const double *in, double *out, int width_blocks,
                                          int height_blocks)
{
    if (!ctx || !in || !out || width_blocks <= 0 || height_blocks <= 0 || (width_blocks & 1)) return JPGB200_ERR_BAD_ARG;
    CK(cudaSetDevice(ctx->device));
    const size_t nin = (size_t)width_blocks * height_blocks, nout = nin / 2;
    Scratch sc(ctx);
    void *di = NULL, *dout = NULL;
    int rc;
    if ((rc = sc.get(nin * 512, &di)) || (rc = sc.get(nout * 512, &dout))) return rc;
    CK(cudaMemcpyAsync(di, in, nin * 512, cudaMemcpyHostToDevice, ctx->stream));
    kf_subsample422_blocks<<<(unsigned)nout, 64, 0, ctx->stream>>>((const double *)di, (double *)dout, width_blocks);
    ctx->launch_total++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, dout, nout * 512, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return JPGB200_OK;
}

extern "C" int jpgb200_stage_dpcm(jpgb200_ctx *ctx, int *dc, long long n)
{
    if (!ctx || !dc || n <= 0) return JPGB200_ERR_BAD_ARG;
    CK(cudaSetDevice(ctx->device));
    const unsigned tiles = (unsigned)((n + K2_TILE - 1) / K2_TILE);
    Scratch sc(ctx);
    void *din = NULL, *dout = NULL, *dts = NULL;
    int rc;
    if ((rc = sc.get((size_t)n * 4, &din)) || (rc = sc.get((size_t)n * 4, &dout)) || (rc = sc.get((size_t)tiles * 4, &dts)))
        return rc;
    CK(cudaMemcpyAsync(din, dc, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    k2s_tile_sums<<<tiles, K2_THREADS, 0, ctx->stream>>>((const int *)din, n, (int *)dts);
    k2_scan_tiles<<<1, 1024, 0, ctx->stream>>>((int *)dts, (long long)tiles);
    k2s_apply<<<tiles, K2_THREADS, 0, ctx->stream>>>((const int *)din, n, (const int *)dts, (int *)dout);
    ctx->launch_total += 3;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(dc, dout, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return JPGB200_OK;
}

extern "C" int jpgb200_stage_histograms(jpgb200_ctx *ctx, const int *zz, long long n, int sym_mode, int *freq_dc,
                                        int *freq_ac)
{
    if (!ctx || !zz || !freq_dc || !freq_ac || n <= 0) return JPGB200_ERR_BAD_ARG;
    CK(cudaSetDevice(ctx->device));
    Scratch sc(ctx);
    void *dz = NULL, *df = NULL;
    int rc;
    if ((rc = sc.get((size_t)n * 256, &dz)) || (rc = sc.get(2 * 257 * 4, &df))) return rc;
    int *f = (int *)df;
    CK(cudaMemcpyAsync(dz, zz, (size_t)n * 256, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(f, freq_dc, 257 * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(f + 257, freq_ac, 257 * 4, cudaMemcpyHostToDevice, ctx->stream));
    khs_histograms<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((const int *)dz, n, sym_mode, f, f + 257);
    ctx->launch_total++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(freq_dc, f, 257 * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(freq_ac, f + 257, 257 * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return JPGB200_OK;
}

/* Entropy-code caller-supplied quantised coefficients (host, int16 [Y|Cb|Cr][block][64]) into a
 * complete JFIF stream: K3..K6 (or K35) only.  "Given identical coefficients the bytes are
 * identical" is checked through this entry point. */
extern "C" int jpgb200_encode_coefficients(jpgb200_ctx *ctx, const int16_t *zz, int w, int h, int quality, int ratio,
                                           uint8_t *out, size_t cap, size_t *len)
{
    if (!ctx || !zz || !out) return JPGB200_ERR_BAD_ARG;
    long long n[3];
    int rc = validate(w, h, quality, ratio, n);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    if ((rc = ensure_workspace(ctx, n, 1))) return rc;
    if ((rc = ensure_tables(ctx, w, h, quality, ratio))) return rc;
    if (ctx->out_cap < cap || !ctx->d_out[1] || !ctx->d_planes[1]) {
        if ((rc = ensure_io(ctx, ctx->planes_cap ? ctx->planes_cap / 3 : 64, cap, 2))) return rc;
    }
    ctx->w = w;
    ctx->h = h;
    ctx->q = quality;
    ctx->ratio = ratio;
    memcpy(ctx->n, n, sizeof n);
    CK(cudaMemcpyAsync(ctx->d_zz, zz, (size_t)(n[0] + n[1] + n[2]) * 128, cudaMemcpyHostToDevice, ctx->stream));
    {
        /* the entropy kernels index their tables with the categories: refuse what a quantiser cannot produce */
        int bad = 0;
        CK(cudaMemsetAsync(ctx->d_freq, 0, sizeof(int), ctx->stream));
        const long long nb = n[0] + n[1] + n[2];
        k_validate_coefs<<<(unsigned)((nb + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_zz, n[0], n[1], n[2], ctx->d_freq);
        ctx->launch_total++;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(&bad, ctx->d_freq, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        if (bad) {
            snprintf(ctx->err, sizeof ctx->err, "coefficients out of range: AC must be in [-1023, 1023], DC differences in [-2047, 2047]");
            return JPGB200_ERR_BAD_ARG;
        }
    }
    if (ctx->opt_huff && (rc = optimise_tables(ctx, w, h, quality, ratio))) return rc;
    if (ctx->legacy_entropy || ratio == 2) {
        if ((rc = launch_bits(ctx, ctx->d_zz, ratio, n, NULL, true))) return rc;
        if ((rc = launch_pack(ctx, ctx->d_zz, ratio, n, ctx->d_out[0], cap, ctx->d_len))) return rc;
    } else {
        if ((rc = launch_fused(ctx, ctx->d_zz, ratio, n, ctx->d_out[0], cap, ctx->d_len))) return rc;
    }
    return finish_host(ctx, ctx->d_out[0], out, cap, len);
}

/* ---- device helpers -------------------------------------------------------------------------------- */

extern "C" int jpgb200_malloc(jpgb200_ctx *ctx, void **dptr, size_t bytes)
{
    if (!ctx || !dptr) return JPGB200_ERR_BAD_ARG;
    CK(cudaSetDevice(ctx->device));
    cudaError_t e = cudaMalloc(dptr, bytes);
    if (e != cudaSuccess) {
        snprintf(ctx->err, sizeof ctx->err, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        return JPGB200_ERR_NOMEM;
    }
    return JPGB200_OK;
}

extern "C" int jpgb200_free(jpgb200_ctx *ctx, void *dptr)
{
    if (!ctx) return JPGB200_ERR_BAD_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaFree(dptr));
    return JPGB200_OK;
}

extern "C" int jpgb200_host_alloc(void **hptr, size_t bytes)
{
    if (!hptr) return JPGB200_ERR_BAD_ARG;
    return cudaHostAlloc(hptr, bytes, cudaHostAllocDefault) == cudaSuccess ? JPGB200_OK : JPGB200_ERR_NOMEM;
}

/* flags: 1 = write-combined (a buffer the host only writes: frames on their way to the device) */
extern "C" int jpgb200_host_alloc_ex(void **hptr, size_t bytes, int flags)
{
    if (!hptr) return JPGB200_ERR_BAD_ARG;
    return cudaHostAlloc(hptr, bytes, (flags & 1) ? cudaHostAllocWriteCombined : cudaHostAllocDefault) == cudaSuccess ? JPGB200_OK
                                                                                                                    : JPGB200_ERR_NOMEM;
}

/* NUMA node of the device's PCIe function (-1: not reported, e.g. inside a VM) and the number of NUMA
 * nodes this process can see: what a caller needs to place its pinned buffers */
extern "C" int jpgb200_device_numa(int device, int *gpu_node, int *visible_nodes)
{
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) return JPGB200_ERR_NO_DEVICE;
    for (char *c = bus; *c; c++)
        if (*c >= 'A' && *c <= 'Z') *c = (char)(*c - 'A' + 'a');
    char path[128];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
    int node = -1;
    if (FILE *f = fopen(path, "r")) {
        if (fscanf(f, "%d", &node) != 1) node = -1;
        fclose(f);
    }
    int nodes = 0;
    for (int i = 0; i < 64; i++) {
        snprintf(path, sizeof path, "/sys/devices/system/node/node%d", i);
        if (FILE *f = fopen((std::string(path) + "/cpulist").c_str(), "r")) {
            nodes++;
            fclose(f);
        }
    }
    if (gpu_node) *gpu_node = node;
    if (visible_nodes) *visible_nodes = nodes;
    return JPGB200_OK;
}

extern "C" int jpgb200_host_free(void *hptr) { return cudaFreeHost(hptr) == cudaSuccess ? JPGB200_OK : JPGB200_ERR_CUDA; }

extern "C" int jpgb200_memcpy_h2d(jpgb200_ctx *ctx, void *dptr, const void *hptr, size_t bytes)
{
    if (!ctx) return JPGB200_ERR_BAD_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(dptr, hptr, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return JPGB200_OK;
}

extern "C" int jpgb200_memcpy_d2h(jpgb200_ctx *ctx, void *hptr, const void *dptr, size_t bytes)
{
    if (!ctx) return JPGB200_ERR_BAD_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(hptr, dptr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return JPGB200_OK;
}

extern "C" int jpgb200_synth_device(jpgb200_ctx *ctx, uint8_t *d_planes, int w, int h, uint64_t seed, uint32_t frame,
                                    int kind)
{
    if (!ctx || !d_planes || w <= 0 || h <= 0 || (w % 8)) return JPGB200_ERR_BAD_ARG;
    CK(cudaSetDevice(ctx->device));
    dim3 grid((unsigned)ctx->sm_count * 8, 1, 3);
    ks_synth<<<grid, 256, 0, ctx->stream>>>(d_planes, w, h, (unsigned long long)seed, frame, kind, 0);
    ctx->launch_total++;
    CK(cudaGetLastError());
    return JPGB200_OK;
}

/* rows row0 .. row0 + rows - 1 of the same synthetic image into planes [3][rows][w]: a rank of the MCU-row
 * partition generates its band (and the halo row above it) without ever holding the whole image */
extern "C" int jpgb200_synth_device_rows(jpgb200_ctx *ctx, uint8_t *d_planes, int w, int rows, int row0, uint64_t seed,
                                         uint32_t frame, int kind)
{
    if (!ctx || !d_planes || w <= 0 || rows <= 0 || row0 < 0 || (w % 8)) return JPGB200_ERR_BAD_ARG;
    CK(cudaSetDevice(ctx->device));
    dim3 grid((unsigned)ctx->sm_count * 8, 1, 3);
    ks_synth<<<grid, 256, 0, ctx->stream>>>(d_planes, w, rows, (unsigned long long)seed, frame, kind, row0);
    ctx->launch_total++;
    CK(cudaGetLastError());
    return JPGB200_OK;
}
