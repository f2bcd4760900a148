/*
 * oracle/ref_harness.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Driver that links the UNMODIFIED reference sources where they lie under
 * /root/reference/src (see oracle/Makefile; nothing is copied into this repo)
 * and exposes the reference's own stage functions for (a) pinning the oracle and
 * (b) timing the reference's CPU implementation on the GPU box's host cores.
 * The built binary goes to oracle/_ref/ (git-ignored, shipped by gpurun).
 *
 * It never calls huffman_encode(): construct_huffman_table() does not
 * terminate.  The histogram loops of src/huffman.c:29-42 are
 * replayed through the extern calculate_freq_block_DC/AC instead.
 *
 * One process per image: scale_table() rewrites global tables in place
 * (src/quantise.c:34-35,83).
 *
 *   ref_harness dump <in.bmp> <quality> <out.bin>
 *   ref_harness time <in.bmp> <quality> <out.json>
 *   ref_harness testdct <out.txt>
 *
 * dump file layout (little endian):
 *   int32 magic 0x4A524546, w, h, quality, nblocks
 *   uint8 oob[24]                (bytes in front of r_new, g_new, b_new)
 *   uint8 r[w*h], g[w*h], b[w*h] (planes as bmp_GetRed/Green/Blue return them)
 *   int32 zz[3][nblocks][64]     (after zig_zag, before dpcm)
 *   int32 dpcm[3][nblocks]       (zz[*][0] after dpcm)
 *   int32 freq[4][257]           (lum_DC, lum_AC, chrom_DC, chrom_AC)
 */
#define _POSIX_C_SOURCE 199309L
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "headers/bitmap.h"
#include "headers/block.h"
#include "headers/dct.h"
#include "headers/downsample.h"
#include "headers/dpcm.h"
#include "headers/jpg_encode.h"
#include "headers/preprocess.h"
#include "headers/quantise.h"
#include "headers/zig_zag.h"

/* external linkage in src/huffman.c:182,187 */
void calculate_freq_block_DC(HuffmanData *huffman_data, int *image_data);
void calculate_freq_block_AC(HuffmanData *huffman_data, int *image_data);
/* src/jpg_driver.c:54, compiled with -Dmain=ref_driver_main */
void test_dct(void);

/* ---- malloc interposer (-Wl,--wrap=malloc): remember the 8 bytes in front of the
 * last three allocations of exactly `watch_size` bytes: r_new, g_new, b_new of
 * src/preprocess.c:127-129. ------------------------------------------------- */
void *__real_malloc(size_t n);
static size_t watch_size = 0;
static unsigned char oob_ring[3][8];
static int oob_count = 0;

void *__wrap_malloc(size_t n)
{
    void *p = __real_malloc(n);
    if (p && watch_size && n == watch_size) {
        memmove(oob_ring[0], oob_ring[1], 8);
        memmove(oob_ring[1], oob_ring[2], 8);
        memcpy(oob_ring[2], (unsigned char *)p - 8, 8);
        oob_count++;
    }
    return p;
}

static double now(void)
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

static FILE *quiet_stdout(void)
{
    /* the reference prints one line per block (src/zig_zag.c:50) */
    fflush(stdout);
    return freopen("/dev/null", "w", stdout);
}

static void freq_pass(JpgData j)
{
    HuffmanData *t[4] = {&j->lum_DC, &j->lum_AC, &j->chrom_DC, &j->chrom_AC};
    for (int k = 0; k < 4; k++) {
        memset(t[k]->freq, 0, sizeof t[k]->freq);
        t[k]->freq[256] = 1; /* src/huffman.c:54 */
    }
    for (int i = 0; i < j->num_blocks_Y; i++) {
        calculate_freq_block_DC(&j->lum_DC, j->zig_zag_Y[i]);
        calculate_freq_block_AC(&j->lum_AC, j->zig_zag_Y[i]);
    }
    for (int i = 0; i < j->num_blocks_Cb; i++) {
        calculate_freq_block_DC(&j->chrom_DC, j->zig_zag_Cb[i]);
        calculate_freq_block_AC(&j->chrom_AC, j->zig_zag_Cb[i]);
    }
    for (int i = 0; i < j->num_blocks_Cr; i++) {
        calculate_freq_block_DC(&j->chrom_DC, j->zig_zag_Cr[i]);
        calculate_freq_block_AC(&j->chrom_AC, j->zig_zag_Cr[i]);
    }
}

static int bmp_dims(const char *path, int *w, int *h)
{
    unsigned char hdr[32];
    FILE *f = fopen(path, "rb");
    if (!f) return -1;
    if (fread(hdr, 1, 32, f) != 32) {
        fclose(f);
        return -1;
    }
    fclose(f);
    memcpy(w, hdr + 18, 4);
    memcpy(h, hdr + 22, 4);
    return 0;
}

static int cmd_dump(const char *in, int quality, const char *outp)
{
    int w, h;
    if (bmp_dims(in, &w, &h)) return 2;
    FILE *out = fopen(outp, "wb");
    if (!out) return 3;
    quiet_stdout();

    /* planes exactly as the reference's loader produces them */
    BmpImage bmp = bmp_OpenBitmap(in);
    int np = bmp_GetNumPixels(bmp);
    int32_t hdr[5] = {0x4A524546, w, h, quality, (w / 8) * (h / 8)};

    JpgData j = calloc(1, sizeof(JpegData));
    j->input_filename = (char *)in;
    j->quality = quality;
    j->sample_ratio = NO_CHROMA_SUBSAMPLING;

    watch_size = (size_t)w * (size_t)h;
    preprocess_jpeg(j);
    watch_size = 0;

    fwrite(hdr, 4, 5, out);
    fwrite(oob_ring, 1, 24, out);
    fwrite(bmp_GetRed(bmp), 1, (size_t)np, out);
    fwrite(bmp_GetGreen(bmp), 1, (size_t)np, out);
    fwrite(bmp_GetBlue(bmp), 1, (size_t)np, out);
    bmp_DestroyBitmap(bmp);

    chroma_subsample(j);
    dct(j);
    quantise(j);
    zig_zag(j);
    int n = j->num_blocks_Y;
    for (int i = 0; i < n; i++) fwrite(j->zig_zag_Y[i], 4, 64, out);
    for (int i = 0; i < n; i++) fwrite(j->zig_zag_Cb[i], 4, 64, out);
    for (int i = 0; i < n; i++) fwrite(j->zig_zag_Cr[i], 4, 64, out);
    dpcm(j);
    for (int i = 0; i < n; i++) fwrite(&j->zig_zag_Y[i][0], 4, 1, out);
    for (int i = 0; i < n; i++) fwrite(&j->zig_zag_Cb[i][0], 4, 1, out);
    for (int i = 0; i < n; i++) fwrite(&j->zig_zag_Cr[i][0], 4, 1, out);
    freq_pass(j);
    fwrite(j->lum_DC.freq, 4, 257, out);
    fwrite(j->lum_AC.freq, 4, 257, out);
    fwrite(j->chrom_DC.freq, 4, 257, out);
    fwrite(j->chrom_AC.freq, 4, 257, out);
    fclose(out);
    return 0;
}

static int cmd_time(const char *in, int quality, const char *outp)
{
    int w, h;
    double t[8];
    if (bmp_dims(in, &w, &h)) return 2;
    FILE *out = fopen(outp, "w");
    if (!out) return 3;
    quiet_stdout();
    JpgData j = calloc(1, sizeof(JpegData));
    j->input_filename = (char *)in;
    j->quality = quality;
    j->sample_ratio = NO_CHROMA_SUBSAMPLING;
    t[0] = now();
    preprocess_jpeg(j);
    t[1] = now();
    chroma_subsample(j);
    dct(j);
    t[2] = now();
    quantise(j);
    t[3] = now();
    zig_zag(j);
    t[4] = now();
    dpcm(j);
    t[5] = now();
    freq_pass(j);
    t[6] = now();
    double total = t[6] - t[0];
    fprintf(out,
            "{\"w\": %d, \"h\": %d, \"quality\": %d, \"seconds\": %.6f, \"mpixel_per_s\": %.6f, "
            "\"preprocess_s\": %.6f, \"dct_s\": %.6f, \"quantise_s\": %.6f, \"zigzag_s\": %.6f, "
            "\"dpcm_s\": %.6f, \"freq_s\": %.6f}\n",
            w, h, quality, total, (double)w * h / total / 1e6, t[1] - t[0], t[2] - t[1], t[3] - t[2],
            t[4] - t[3], t[5] - t[4], t[6] - t[5]);
    fclose(out);
    return 0;
}

static int cmd_testdct(const char *outp)
{
    /* src/jpg_driver.c:54-150 prints the known-answer vector to stdout */
    fflush(stdout);
    if (!freopen(outp, "w", stdout)) return 3;
    test_dct();
    fflush(stdout);
    return 0;
}

int main(int argc, char **argv)
{
    if (argc >= 5 && !strcmp(argv[1], "dump")) return cmd_dump(argv[2], atoi(argv[3]), argv[4]);
    if (argc >= 5 && !strcmp(argv[1], "time")) return cmd_time(argv[2], atoi(argv[3]), argv[4]);
    if (argc >= 3 && !strcmp(argv[1], "testdct")) return cmd_testdct(argv[2]);
    fprintf(stderr, "usage: ref_harness dump|time <in.bmp> <quality> <out>  |  testdct <out>\n");
    return 1;
}
