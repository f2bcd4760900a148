/*
 * The reference's per-stage API -- void stage(JpgData), src/jpg_encode.c:32-50 -- with every
 * stage's arithmetic executed by a GPU kernel of libjpgb200 ("staged mode").  It exists for
 * drop-in fidelity and debugging: callers that walk the stages themselves (as the reference's
 * driver does, src/jpg_driver.c:135-144) find JpegData filled exactly as the reference fills
 * it.  encode_bmp_to_jpeg() does NOT go through here; it uses the fused path.
 *
 * Differences from the reference, all deliberate:
 *  - 4:2:2 is implemented (chroma_subsample halves num_blocks_Cb/Cr); 4:2:0 is rejected
 *  - quantise() always scales the pristine tables (the reference compounds on re-entry)
 *  - huffman_encode() terminates: histograms like the reference's loops, then the JPEG file
 *  - no per-block printf
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "headers/bitmap.h"
#include "headers/block.h"
#include "headers/dct.h"
#include "headers/downsample.h"
#include "headers/dpcm.h"
#include "headers/huffman.h"
#include "headers/preprocess.h"
#include "headers/quantise.h"
#include "headers/zig_zag.h"
#include "host_ctx.h"

static void report(const char *where, int rc)
{
    if (rc != JPGB200_OK) fprintf(stderr, "%s: %s (%s)\n", where, jpgb200_strerror(rc), host_ctx_error());
}

/* gather / scatter between the reference's array-of-handles and the contiguous layout the
 * kernels take */
static double *gather(Block *blocks, int n)
{
    double *buf = (double *)malloc((size_t)n * 64 * sizeof(double));
    if (buf)
        for (int i = 0; i < n; i++) memcpy(buf + (size_t)i * 64, block_values(blocks[i]), 64 * sizeof(double));
    return buf;
}
static void scatter(Block *blocks, int n, const double *buf)
{
    for (int i = 0; i < n; i++) memcpy(block_values(blocks[i]), buf + (size_t)i * 64, 64 * sizeof(double));
}
static Block *alloc_blocks(int n)
{
    Block *a = (Block *)malloc(sizeof(Block) * (size_t)n);
    for (int i = 0; a && i < n; i++) a[i] = new_block();
    return a;
}

/* ---- preprocess_jpeg: src/preprocess.c:25-70 ------------------------------------------------ */
void preprocess_jpeg(JpgData j)
{
    jpgb200_ctx *ctx = host_ctx();
    BmpImage bmp = bmp_OpenBitmap(j->input_filename);
    if (!ctx || !bmp || bmp_GetErrorCode(bmp) != BMP_SUCCESS) {
        if (bmp) bmp_GetLastError(bmp);
        bmp_DestroyBitmap(bmp);
        return;
    }
    j->width = bmp_GetWidth(bmp);
    j->height = bmp_GetHeight(bmp);
    const int need_w = (j->sample_ratio == NO_CHROMA_SUBSAMPLING) ? 8 : 16;
    if (j->width % need_w || j->height % 8 || j->sample_ratio == HORIZONTAL_VERTICAL_SUBSAMPLING) {
        report("preprocess_jpeg", JPGB200_ERR_BAD_DIMS); /* the reference's padding path is broken (B.8) */
        bmp_DestroyBitmap(bmp);
        return;
    }
    const int n = (j->width / 8) * (j->height / 8);
    j->num_blocks_Y = j->num_blocks_Cb = j->num_blocks_Cr = n; /* src/preprocess.c:45-47 */
    j->Y = alloc_blocks(n);
    j->Cb = alloc_blocks(n);
    j->Cr = alloc_blocks(n);
    double *buf = (double *)malloc((size_t)n * 64 * 3 * sizeof(double));
    if (j->Y && j->Cb && j->Cr && buf) {
        const int rc = jpgb200_stage_preprocess(ctx, bmp_GetRed(bmp), bmp_GetGreen(bmp), bmp_GetBlue(bmp), j->width,
                                                j->height, buf, buf + (size_t)n * 64, buf + (size_t)n * 128);
        report("preprocess_jpeg", rc);
        if (rc == JPGB200_OK) {
            scatter(j->Y, n, buf);
            scatter(j->Cb, n, buf + (size_t)n * 64);
            scatter(j->Cr, n, buf + (size_t)n * 128);
        }
    }
    free(buf);
    bmp_DestroyBitmap(bmp);
}

/* ---- chroma_subsample: src/downsample.c:9-32 (stubs there) ------------------------------------ */
static void halve(jpgb200_ctx *ctx, Block **blocks, int *count, int bw, int bh)
{
    const int n = *count;
    double *in = gather(*blocks, n), *out = (double *)malloc((size_t)(n / 2) * 64 * sizeof(double));
    if (in && out) {
        const int rc = jpgb200_stage_subsample422(ctx, in, out, bw, bh);
        report("chroma_subsample", rc);
        if (rc == JPGB200_OK) {
            for (int i = n / 2; i < n; i++) destroy_block((*blocks)[i]);
            scatter(*blocks, n / 2, out);
            *count = n / 2;
        }
    }
    free(in);
    free(out);
}

void chroma_subsample(JpgData j)
{
    jpgb200_ctx *ctx = host_ctx();
    if (!ctx || j->sample_ratio != HORIZONTAL_SUBSAMPLING) return;
    halve(ctx, &j->Cb, &j->num_blocks_Cb, j->width / 8, j->height / 8);
    halve(ctx, &j->Cr, &j->num_blocks_Cr, j->width / 8, j->height / 8);
}

/* ---- dct: src/dct.c:19-59 ------------------------------------------------------------------- */
static void dct_many(jpgb200_ctx *ctx, Block *blocks, int n)
{
    double *buf = gather(blocks, n);
    if (buf) {
        const int rc = jpgb200_stage_dct(ctx, buf, n);
        report("dct", rc);
        if (rc == JPGB200_OK) scatter(blocks, n, buf);
    }
    free(buf);
}

void dct(JpgData j)
{
    jpgb200_ctx *ctx = host_ctx();
    if (!ctx) return;
    dct_many(ctx, j->Y, j->num_blocks_Y);
    dct_many(ctx, j->Cb, j->num_blocks_Cb);
    dct_many(ctx, j->Cr, j->num_blocks_Cr);
}

void dct_block(Block b)
{
    jpgb200_ctx *ctx = host_ctx();
    if (ctx) report("dct_block", jpgb200_stage_dct(ctx, block_values(b), 1));
}

/* ---- quantise: src/quantise.c:30-86 ----------------------------------------------------------- */
static int cur_lum[64], cur_chr[64], tables_ready = 0;

static void base_tables(void)
{
    if (!tables_ready) {
        jpgb200_quant_tables(50, cur_lum, cur_chr); /* s = 100: the Annex-K base tables */
        tables_ready = 1;
    }
}

static void quant_many(jpgb200_ctx *ctx, Block *blocks, int n, const int *table)
{
    double *buf = gather(blocks, n);
    if (buf) {
        const int rc = jpgb200_stage_quantise(ctx, buf, n, table);
        report("quantise", rc);
        if (rc == JPGB200_OK) scatter(blocks, n, buf);
    }
    free(buf);
}

void quantise(JpgData j)
{
    jpgb200_ctx *ctx = host_ctx();
    if (!ctx) return;
    if (jpgb200_quant_tables(j->quality, cur_lum, cur_chr) != JPGB200_OK) {
        report("quantise", JPGB200_ERR_BAD_QUALITY);
        return;
    }
    tables_ready = 1;
    quant_many(ctx, j->Y, j->num_blocks_Y, cur_lum);
    quant_many(ctx, j->Cb, j->num_blocks_Cb, cur_chr);
    quant_many(ctx, j->Cr, j->num_blocks_Cr, cur_chr);
}

void quantise_lum(Block b)
{
    jpgb200_ctx *ctx = host_ctx();
    base_tables();
    if (ctx) report("quantise_lum", jpgb200_stage_quantise(ctx, block_values(b), 1, cur_lum));
}

void quantise_chr(Block b)
{
    jpgb200_ctx *ctx = host_ctx();
    base_tables();
    if (ctx) report("quantise_chr", jpgb200_stage_quantise(ctx, block_values(b), 1, cur_chr));
}

/* ---- zig_zag: src/zig_zag.c:17-59 -------------------------------------------------------------- */
static int **zz_many(jpgb200_ctx *ctx, Block *blocks, int n)
{
    int **rows = (int **)malloc(sizeof(int *) * (size_t)n);
    double *buf = gather(blocks, n);
    int *flat = (int *)malloc((size_t)n * 64 * sizeof(int));
    if (rows && buf && flat) {
        const int rc = jpgb200_stage_zigzag(ctx, buf, n, flat);
        report("zig_zag", rc);
        for (int i = 0; i < n; i++) {
            rows[i] = (int *)calloc(64, sizeof(int)); /* one allocation per block, like the reference */
            if (rows[i] && rc == JPGB200_OK) memcpy(rows[i], flat + (size_t)i * 64, 64 * sizeof(int));
        }
    }
    free(buf);
    free(flat);
    return rows;
}

void zig_zag(JpgData j)
{
    jpgb200_ctx *ctx = host_ctx();
    if (!ctx) return;
    j->zig_zag_Y = zz_many(ctx, j->Y, j->num_blocks_Y);
    j->zig_zag_Cb = zz_many(ctx, j->Cb, j->num_blocks_Cb);
    j->zig_zag_Cr = zz_many(ctx, j->Cr, j->num_blocks_Cr);
}

void zig_zag_block(Block b, int *zz)
{
    jpgb200_ctx *ctx = host_ctx();
    if (ctx) report("zig_zag_block", jpgb200_stage_zigzag(ctx, block_values(b), 1, zz));
}

/* ---- dpcm: src/dpcm.c:6-21 --------------------------------------------------------------------- */
static void dpcm_comp(jpgb200_ctx *ctx, int **zz, int n)
{
    int *dc = (int *)malloc(sizeof(int) * (size_t)n);
    if (!dc) return;
    for (int i = 0; i < n; i++) dc[i] = zz[i][0];
    const int rc = jpgb200_stage_dpcm(ctx, dc, n);
    report("dpcm", rc);
    if (rc == JPGB200_OK)
        for (int i = 0; i < n; i++) zz[i][0] = dc[i];
    free(dc);
}

void dpcm(JpgData j)
{
    jpgb200_ctx *ctx = host_ctx();
    if (!ctx) return;
    dpcm_comp(ctx, j->zig_zag_Y, j->num_blocks_Y);
    dpcm_comp(ctx, j->zig_zag_Cb, j->num_blocks_Cb);
    dpcm_comp(ctx, j->zig_zag_Cr, j->num_blocks_Cr);
}

/* ---- huffman_encode: src/huffman.c:23-74,182-222 ---------------------------------------------- */
static void init_tables(JpgData j)
{
    HuffmanData *t[4] = {&j->lum_DC, &j->lum_AC, &j->chrom_DC, &j->chrom_AC};
    for (int k = 0; k < 4; k++) {
        memset(t[k], 0, sizeof *t[k]);
        t[k]->freq[256] = 1; /* reserved code point, src/huffman.c:54 */
        for (int i = 0; i <= 256; i++) t[k]->others[i] = -1;
    }
}

static void freq_comp(jpgb200_ctx *ctx, int **zz, int n, HuffmanData *dc, HuffmanData *ac)
{
    int *flat = (int *)malloc((size_t)n * 64 * sizeof(int));
    if (!flat) return;
    for (int i = 0; i < n; i++) memcpy(flat + (size_t)i * 64, zz[i], 64 * sizeof(int));
    report("huffman_encode", jpgb200_stage_histograms(ctx, flat, n, JPGB200_SYM_REF, dc->freq, ac->freq));
    free(flat);
}

void calculate_freq_block_DC(HuffmanData *h, int *image_data)
{
    jpgb200_ctx *ctx = host_ctx();
    int scratch[257] = {0};
    if (ctx) report("calculate_freq_block_DC", jpgb200_stage_histograms(ctx, image_data, 1, JPGB200_SYM_REF, h->freq, scratch));
}

void calculate_freq_block_AC(HuffmanData *h, int *image_data)
{
    jpgb200_ctx *ctx = host_ctx();
    int scratch[257] = {0};
    if (ctx) report("calculate_freq_block_AC", jpgb200_stage_histograms(ctx, image_data, 1, JPGB200_SYM_REF, scratch, h->freq));
}

/* undo the dpcm() recurrence to recover the quantised DCs: dc[i] = d[i] + d[i-1] */
static void pack_comp(int16_t *dst, int **zz, int n)
{
    int prev = 0;
    for (int i = 0; i < n; i++) {
        for (int k = 1; k < 64; k++) dst[(size_t)i * 64 + k] = (int16_t)zz[i][k];
        dst[(size_t)i * 64] = (int16_t)(zz[i][0] + (i ? prev : 0));
        prev = zz[i][0];
    }
}

void huffman_encode(JpgData j)
{
    jpgb200_ctx *ctx = host_ctx();
    if (!ctx) return;
    init_tables(j);
    freq_comp(ctx, j->zig_zag_Y, j->num_blocks_Y, &j->lum_DC, &j->lum_AC);
    freq_comp(ctx, j->zig_zag_Cb, j->num_blocks_Cb, &j->chrom_DC, &j->chrom_AC);
    freq_comp(ctx, j->zig_zag_Cr, j->num_blocks_Cr, &j->chrom_DC, &j->chrom_AC);
    if (!j->output_filename) return;
    /* the byte stream the reference never gets to: standard tables, true DC differences */
    const size_t nb = (size_t)j->num_blocks_Y + j->num_blocks_Cb + j->num_blocks_Cr;
    int16_t *zz = (int16_t *)malloc(nb * 64 * sizeof(int16_t));
    /* worst case: every block at its longest (1658 bits) and every byte stuffed */
    const size_t cap = 1024 + nb * ((1658 + 7) / 8) * 2 + 2;
    unsigned char *out = (unsigned char *)malloc(cap);
    if (zz && out) {
        pack_comp(zz, j->zig_zag_Y, j->num_blocks_Y);
        pack_comp(zz + (size_t)j->num_blocks_Y * 64, j->zig_zag_Cb, j->num_blocks_Cb);
        pack_comp(zz + ((size_t)j->num_blocks_Y + j->num_blocks_Cb) * 64, j->zig_zag_Cr, j->num_blocks_Cr);
        size_t len = 0;
        const int rc = jpgb200_encode_coefficients(ctx, zz, j->width, j->height, j->quality, j->sample_ratio, out, cap, &len);
        report("huffman_encode", rc);
        if (rc == JPGB200_OK) {
            FILE *f = fopen(j->output_filename, "wb");
            if (f) {
                fwrite(out, 1, len, f);
                fclose(f);
            } else {
                fprintf(stderr, "huffman_encode: cannot write %s\n", j->output_filename);
            }
        }
    }
    free(zz);
    free(out);
}

/* The stage chain of the reference's encode_bmp_to_jpeg, for callers who want JpegData filled. */
JpgData encode_bmp_to_jpeg_staged(const char *input, const char *output, int quality, int sample_ratio)
{
    JpgData j = (JpgData)calloc(1, sizeof(JpegData));
    if (!j) return NULL;
    j->sample_ratio = sample_ratio;
    j->quality = quality;
    j->output_filename = (char *)output;
    j->input_filename = (char *)input;
    preprocess_jpeg(j);
    if (!j->Y) return j;
    chroma_subsample(j);
    dct(j);
    quantise(j);
    zig_zag(j);
    dpcm(j);
    huffman_encode(j);
    return j;
}
