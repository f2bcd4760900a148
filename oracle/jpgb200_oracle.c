/*
 * oracle/jpgb200_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, FP64) of the encode hot path of
 * matthewT53/JPEG-Encoder-and-Decoder.  See jpgb200_oracle.h for what is
 * pinned by the compiled reference and what is "parity unpinned" (T.81 only).
 * File:line citations are relative to /root/reference/.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off so that every FP64
 * operation rounds exactly where the reference's -O0 build rounds).
 */
#include "jpgb200_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#ifndef M_PI
#define M_PI (3.14159265358979323846)
#endif

/* ---- constant tables ------------------------------------------------------ */

/* Annex K.1 / K.2 base tables as carried by src/quantise.c:8-25 */
static const int base_lum[64] = {16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,
                                 14, 13, 16, 24, 40,  57,  69,  56,  14, 17, 22, 29, 51,  87,  80,  62,
                                 18, 22, 37, 56, 68,  109, 103, 77,  24, 35, 55, 64, 81,  104, 113, 92,
                                 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
static const int base_chr[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99,
                                 24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
                                 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                                 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};

/* src/zig_zag.c:6-15: scan_order[row][col] */
static const int scan_order[64] = {0,  1,  5,  6,  14, 15, 27, 28, 2,  4,  7,  13, 16, 26, 29, 42,
                                   3,  8,  12, 17, 25, 30, 41, 43, 9,  11, 18, 24, 31, 40, 44, 53,
                                   10, 19, 23, 32, 39, 45, 52, 54, 20, 22, 33, 38, 46, 51, 55, 60,
                                   21, 34, 37, 47, 50, 56, 59, 61, 35, 36, 48, 49, 57, 58, 62, 63};

/* Huffman specs, src/headers/tables.h:37-72 (== T.81 tables K.3-K.6) */
static const uint8_t dc_lum_nr[17] = {0, 0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
static const uint8_t dc_chr_nr[17] = {0, 0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
static const uint8_t dc_vals[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
static const uint8_t ac_lum_nr[17] = {0, 0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7D};
static const uint8_t ac_lum_vals[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71,
    0x14, 0x32, 0x81, 0x91, 0xA1, 0x08, 0x23, 0x42, 0xB1, 0xC1, 0x15, 0x52, 0xD1, 0xF0, 0x24, 0x33, 0x62, 0x72,
    0x82, 0x09, 0x0A, 0x16, 0x17, 0x18, 0x19, 0x1A, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2A, 0x34, 0x35, 0x36, 0x37,
    0x38, 0x39, 0x3A, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4A, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59,
    0x5A, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6A, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7A, 0x83,
    0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8A, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9A, 0xA2, 0xA3,
    0xA4, 0xA5, 0xA6, 0xA7, 0xA8, 0xA9, 0xAA, 0xB2, 0xB3, 0xB4, 0xB5, 0xB6, 0xB7, 0xB8, 0xB9, 0xBA, 0xC2, 0xC3,
    0xC4, 0xC5, 0xC6, 0xC7, 0xC8, 0xC9, 0xCA, 0xD2, 0xD3, 0xD4, 0xD5, 0xD6, 0xD7, 0xD8, 0xD9, 0xDA, 0xE1, 0xE2,
    0xE3, 0xE4, 0xE5, 0xE6, 0xE7, 0xE8, 0xE9, 0xEA, 0xF1, 0xF2, 0xF3, 0xF4, 0xF5, 0xF6, 0xF7, 0xF8, 0xF9, 0xFA};
static const uint8_t ac_chr_nr[17] = {0, 0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77};
static const uint8_t ac_chr_vals[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22,
    0x32, 0x81, 0x08, 0x14, 0x42, 0x91, 0xA1, 0xB1, 0xC1, 0x09, 0x23, 0x33, 0x52, 0xF0, 0x15, 0x62, 0x72, 0xD1,
    0x0A, 0x16, 0x24, 0x34, 0xE1, 0x25, 0xF1, 0x17, 0x18, 0x19, 0x1A, 0x26, 0x27, 0x28, 0x29, 0x2A, 0x35, 0x36,
    0x37, 0x38, 0x39, 0x3A, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4A, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58,
    0x59, 0x5A, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6A, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7A,
    0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8A, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9A,
    0xA2, 0xA3, 0xA4, 0xA5, 0xA6, 0xA7, 0xA8, 0xA9, 0xAA, 0xB2, 0xB3, 0xB4, 0xB5, 0xB6, 0xB7, 0xB8, 0xB9, 0xBA,
    0xC2, 0xC3, 0xC4, 0xC5, 0xC6, 0xC7, 0xC8, 0xC9, 0xCA, 0xD2, 0xD3, 0xD4, 0xD5, 0xD6, 0xD7, 0xD8, 0xD9, 0xDA,
    0xE2, 0xE3, 0xE4, 0xE5, 0xE6, 0xE7, 0xE8, 0xE9, 0xEA, 0xF2, 0xF3, 0xF4, 0xF5, 0xF6, 0xF7, 0xF8, 0xF9, 0xFA};

static void huff_spec(int table, const uint8_t **nr, const uint8_t **vals, int *nvals)
{
    switch (table) {
    case 0: *nr = dc_lum_nr; *vals = dc_vals; *nvals = 12; break;
    case 1: *nr = dc_chr_nr; *vals = dc_vals; *nvals = 12; break;
    case 2: *nr = ac_lum_nr; *vals = ac_lum_vals; *nvals = 162; break;
    default: *nr = ac_chr_nr; *vals = ac_chr_vals; *nvals = 162; break;
    }
}

/* ---- quantisation tables: src/quantise.c:74-86 ------------------------------ */

static int scale_one(int t, int quality)
{
    int s = (quality < 50) ? 5000 / quality : 200 - 2 * quality; /* :81 */
    return (int)floor((s * t + 50) / 100);                       /* :82, integer division */
}

int orc_quant_tables(int quality, int lum[64], int chr[64])
{
    int ok = 0;
    if (quality < 1 || quality > 100) return -1;
    for (int i = 0; i < 64; i++) {
        lum[i] = scale_one(base_lum[i], quality);
        chr[i] = scale_one(base_chr[i], quality);
        if (lum[i] < 1 || lum[i] > 255 || chr[i] < 1 || chr[i] > 255) ok = -1;
    }
    return ok;
}

int orc_block_counts(int w, int h, int sample_ratio, long n[3])
{
    if (w <= 0 || h <= 0 || (h % 8)) return -1;
    if (sample_ratio == 0) {
        if (w % 8) return -1;
        n[0] = n[1] = n[2] = (long)(w / 8) * (h / 8); /* src/preprocess.c:45-47 */
    } else if (sample_ratio == 1) {
        if (w % 16) return -1;
        n[0] = (long)(w / 8) * (h / 8);
        n[1] = n[2] = n[0] / 2; /* ours: SURVEY.md B.9 */
    } else if (sample_ratio == 2) {
        /* 4:2:0 (ours; subsample_420 is a printf stub, src/downsample.c:29): 16x16 MCUs */
        if ((w % 16) || (h % 16)) return -1;
        n[0] = (long)(w / 8) * (h / 8);
        n[1] = n[2] = n[0] / 4;
    } else {
        return -1;
    }
    return 0;
}

/* ---- single-block stages ---------------------------------------------------- */

static double cos_tab[8][8]; /* cos_tab[x][u] = cos(((2x+1) u pi)/16), src/dct.c:49-50 */
static double alpha_tab[8];
static int tabs_ready = 0;

static void init_tabs(void)
{
    if (tabs_ready) return;
    for (int x = 0; x < 8; x++)
        for (int u = 0; u < 8; u++) cos_tab[x][u] = cos(((2 * x + 1) * u * M_PI) / 16);
    for (int u = 0; u < 8; u++) alpha_tab[u] = (u == 0 ? 1 / sqrt(2) : 1); /* src/dct.c:13 */
    tabs_ready = 1;
}

/* src/dct.c:36-59.  values[y*8+x] as in src/block.c:25-33; result F(u,v) is
 * stored at values[v*8+u] (set_value_block(b,u,v,...)). Summation order: x
 * outer, y inner; product order (f*cos_x)*cos_y. */
void orc_dct_block(double values[64])
{
    double tmp[64];
    init_tabs();
    memcpy(tmp, values, sizeof tmp);
    for (int u = 0; u < 8; u++) {
        for (int v = 0; v < 8; v++) {
            double s = 0.0;
            for (int x = 0; x < 8; x++)
                for (int y = 0; y < 8; y++) s += tmp[y * 8 + x] * cos_tab[x][u] * cos_tab[y][v];
            values[v * 8 + u] = 0.25 * alpha_tab[u] * alpha_tab[v] * s;
        }
    }
}

/* src/quantise.c:52-72: set(b,i,j, round(get(b,i,j)/table[i][j])): coefficient
 * with horizontal frequency i and vertical frequency j uses table[i][j]. */
void orc_quantise_block(double values[64], const int table[64])
{
    for (int i = 0; i < 8; i++)
        for (int j = 0; j < 8; j++) values[j * 8 + i] = round(values[j * 8 + i] / table[i * 8 + j]);
}

/* src/zig_zag.c:48-59: zz[scan_order[i][j]] = (int) get(b, j, i) */
void orc_zigzag_block(const double values[64], int zz[64])
{
    for (int i = 0; i < 8; i++)
        for (int j = 0; j < 8; j++) zz[scan_order[i * 8 + j]] = (int)values[i * 8 + j];
}

/* ---- block extraction: src/preprocess.c:101-211 ----------------------------- */

/* Fetch plane sample for 0-based block column bx, block row by, in-block (x,y),
 * reproducing blockToCoords (:199-211): the last block column of every block
 * row reads one pixel row higher; row -1 of the top-right block comes from the
 * 8 bytes in front of the buffer. */
/* standards-conforming arithmetic instead of the reference's two coefficient-path bugs: no
 * last-column shift, textbook Cb (the product's "robust front door") */
static int g_standard = 0;
void orc_set_standard(int on) { g_standard = on != 0; }

static inline int plane_sample(const uint8_t *p, int w, int bw, int bx, int by, int x, int y,
                               const uint8_t *oob)
{
    long row = (long)by * 8 + y;
    if (bx == bw - 1 && !g_standard) row -= 1;
    if (row < 0) return oob ? oob[x] : 0;
    return p[row * w + (long)bx * 8 + x];
}

/* Level-shifted Y/Cb/Cr of one full-resolution block (src/preprocess.c:160-162,186-188) */
static void convert_block(const uint8_t *r, const uint8_t *g, const uint8_t *b, int w, int bw, int bx,
                          int by, const uint8_t *oob, double *Y, double *Cb, double *Cr)
{
    for (int y = 0; y < 8; y++) {
        for (int x = 0; x < 8; x++) {
            int rv = plane_sample(r, w, bw, bx, by, x, y, oob ? oob : NULL);
            int gv = plane_sample(g, w, bw, bx, by, x, y, oob ? oob + 8 : NULL);
            int bv = plane_sample(b, w, bw, bx, by, x, y, oob ? oob + 16 : NULL);
            double yv = 0.299 * rv + 0.587 * gv + 0.114 * bv;
            double cb = g_standard ? 128 + (-0.168736 * rv - 0.331264 * gv + 0.5 * bv)
                                   : 128 - (0.168736 * rv - 0.331264 * gv + 0.5 * bv);
            double cr = 128 + (0.5 * rv - 0.418688 * gv - 0.081312 * bv);
            Y[y * 8 + x] = yv - 128;
            Cb[y * 8 + x] = cb - 128;
            Cr[y * 8 + x] = cr - 128;
        }
    }
}

static void finish_block(double *blk, const int *table, int16_t *zz, double *dct_out)
{
    int z[64];
    orc_dct_block(blk);
    if (dct_out) memcpy(dct_out, blk, 64 * sizeof(double));
    orc_quantise_block(blk, table);
    orc_zigzag_block(blk, z);
    for (int k = 0; k < 64; k++) zz[k] = (int16_t)z[k];
}

typedef struct {
    const uint8_t *r, *g, *b, *oob;
    int w, bw, sample_ratio;
    const int *lum, *chr;
    const long *n;
    int16_t *zz;
    double *dct_out;
    long lo, hi; /* range of 4:4:4 block indices, or 4:2:2 chroma block indices */
} CoefJob;

static void *coef_worker(void *arg)
{
    CoefJob *J = (CoefJob *)arg;
    const long *n = J->n;
    int bw = J->bw, w = J->w;
    int16_t *zy = J->zz, *zcb = J->zz + n[0] * 64, *zcr = J->zz + (n[0] + n[1]) * 64;
    double *dy = J->dct_out, *dcb = dy ? dy + n[0] * 64 : NULL, *dcr = dy ? dy + (n[0] + n[1]) * 64 : NULL;
    if (J->sample_ratio == 0) {
        for (long bi = J->lo; bi < J->hi; bi++) {
            int by = (int)(bi / bw), bx = (int)(bi % bw);
            double Y[64], Cb[64], Cr[64];
            convert_block(J->r, J->g, J->b, w, bw, bx, by, J->oob, Y, Cb, Cr);
            finish_block(Y, J->lum, zy + bi * 64, dy ? dy + bi * 64 : NULL);
            finish_block(Cb, J->chr, zcb + bi * 64, dcb ? dcb + bi * 64 : NULL);
            finish_block(Cr, J->chr, zcr + bi * 64, dcr ? dcr + bi * 64 : NULL);
        }
    } else if (J->sample_ratio == 2) {
        /* 4:2:0 (ours): units are MCUs of 2x2 luma blocks.  Luma blocks are stored in MCU order
         * (Y00 Y01 Y10 Y11 per MCU: the scan order), chroma is the mean of each 2x2 group of the
         * full-resolution level-shifted samples the reference's blocks would hold. */
        int mw = bw / 2;
        for (long mi = J->lo; mi < J->hi; mi++) {
            int my = (int)(mi / mw), mx = (int)(mi % mw);
            double Yb[4][64], Cbf[4][64], Crf[4][64], Cb[64], Cr[64];
            for (int q = 0; q < 4; q++)
                convert_block(J->r, J->g, J->b, w, bw, 2 * mx + (q & 1), 2 * my + (q >> 1), J->oob, Yb[q], Cbf[q], Crf[q]);
            for (int y = 0; y < 8; y++)
                for (int x = 0; x < 8; x++) {
                    const int q = (y >> 2) * 2 + (x >> 2), sy = (2 * y) & 7, sx = (2 * x) & 7;
                    Cb[y * 8 + x] = 0.25 * (Cbf[q][sy * 8 + sx] + Cbf[q][sy * 8 + sx + 1] + Cbf[q][(sy + 1) * 8 + sx] +
                                            Cbf[q][(sy + 1) * 8 + sx + 1]);
                    Cr[y * 8 + x] = 0.25 * (Crf[q][sy * 8 + sx] + Crf[q][sy * 8 + sx + 1] + Crf[q][(sy + 1) * 8 + sx] +
                                            Crf[q][(sy + 1) * 8 + sx + 1]);
                }
            for (int q = 0; q < 4; q++) finish_block(Yb[q], J->lum, zy + (mi * 4 + q) * 64, dy ? dy + (mi * 4 + q) * 64 : NULL);
            finish_block(Cb, J->chr, zcb + mi * 64, dcb ? dcb + mi * 64 : NULL);
            finish_block(Cr, J->chr, zcr + mi * 64, dcr ? dcr + mi * 64 : NULL);
        }
    } else {
        /* 4:2:2 (ours): form full-resolution level-shifted chroma
         * as the reference's blocks would hold it, then average horizontal pairs. */
        int cbw = bw / 2;
        for (long ci = J->lo; ci < J->hi; ci++) {
            int by = (int)(ci / cbw), cx = (int)(ci % cbw);
            double Y0[64], Y1[64], Cb0[64], Cb1[64], Cr0[64], Cr1[64], Cb[64], Cr[64];
            convert_block(J->r, J->g, J->b, w, bw, 2 * cx, by, J->oob, Y0, Cb0, Cr0);
            convert_block(J->r, J->g, J->b, w, bw, 2 * cx + 1, by, J->oob, Y1, Cb1, Cr1);
            for (int y = 0; y < 8; y++) {
                for (int x = 0; x < 4; x++) {
                    Cb[y * 8 + x] = 0.5 * (Cb0[y * 8 + 2 * x] + Cb0[y * 8 + 2 * x + 1]);
                    Cr[y * 8 + x] = 0.5 * (Cr0[y * 8 + 2 * x] + Cr0[y * 8 + 2 * x + 1]);
                    Cb[y * 8 + 4 + x] = 0.5 * (Cb1[y * 8 + 2 * x] + Cb1[y * 8 + 2 * x + 1]);
                    Cr[y * 8 + 4 + x] = 0.5 * (Cr1[y * 8 + 2 * x] + Cr1[y * 8 + 2 * x + 1]);
                }
            }
            long yi = (long)by * bw + 2 * cx;
            finish_block(Y0, J->lum, zy + yi * 64, dy ? dy + yi * 64 : NULL);
            finish_block(Y1, J->lum, zy + (yi + 1) * 64, dy ? dy + (yi + 1) * 64 : NULL);
            finish_block(Cb, J->chr, zcb + ci * 64, dcb ? dcb + ci * 64 : NULL);
            finish_block(Cr, J->chr, zcr + ci * 64, dcr ? dcr + ci * 64 : NULL);
        }
    }
    return NULL;
}

static int g_threads = 1;
void orc_set_threads(int t) { g_threads = t < 1 ? 1 : (t > 256 ? 256 : t); }

int orc_coefficients(const uint8_t *r, const uint8_t *g, const uint8_t *b, int w, int h, int quality,
                     int sample_ratio, const uint8_t *oob, int16_t *zz, double *dct_out)
{
    int lum[64], chr[64];
    long n[3];
    if (orc_block_counts(w, h, sample_ratio, n)) return -2;
    if (orc_quant_tables(quality, lum, chr)) return -3;
    init_tabs();
    long units = n[1]; /* 4:4:4: blocks; 4:2:2: chroma blocks (= MCUs) */
    int T = g_threads;
    if (T > units) T = (int)units;
    CoefJob jobs[256];
    pthread_t th[256];
    for (int t = 0; t < T; t++) {
        CoefJob J = {r, g, b, oob, w, w / 8, sample_ratio, lum, chr, n, zz, dct_out, units * t / T, units * (t + 1) / T};
        jobs[t] = J;
    }
    if (T == 1) {
        coef_worker(&jobs[0]);
    } else {
        for (int t = 0; t < T; t++) pthread_create(&th[t], NULL, coef_worker, &jobs[t]);
        for (int t = 0; t < T; t++) pthread_join(th[t], NULL);
    }
    return 0;
}

/* ---- DC sequence: src/dpcm.c:6-21 ------------------------------------------- */

void orc_dc_sequence(const int16_t *zz, long n, int mode, int32_t *d)
{
    for (long i = 0; i < n; i++) {
        int32_t dc = zz[i * 64];
        if (i == 0)
            d[i] = dc;
        else if (mode == ORC_DC_REF_RECURRENCE)
            d[i] = dc - d[i - 1]; /* in place and forward => recurrence (F6) */
        else
            d[i] = dc - zz[(i - 1) * 64];
    }
}

/* ---- symbols: src/huffman.c:182-235 ----------------------------------------- */

static int get_class(int value)
{
    int c = 0;
    value = abs(value);
    while (value > 0) {
        value >>= 1;
        c++;
    }
    return c;
}

static void freq_block_ac(int32_t *freq, const int16_t *zz, int sym_mode)
{
    int last = 0, run = 0;
    for (int i = 63; i > 0; i--)
        if (zz[i] != 0) {
            last = i;
            break;
        }
    for (int i = 1; i < 64; i++) {
        if (i == last + 1) {
            freq[0x00]++;
            break;
        }
        if (zz[i] == 0) {
            run++;
            if (run == 16) {
                freq[0xF0]++;
                run = 0;
            }
        } else {
            int c = get_class(zz[i]);
            freq[sym_mode == ORC_SYM_REF ? (run | c) : ((run << 4) | c)]++;
            run = 0;
        }
    }
}

void orc_histograms(const int16_t *zz, const long n[3], int dc_mode, int sym_mode, int32_t *freq)
{
    memset(freq, 0, 4 * 257 * sizeof(int32_t));
    for (int t = 0; t < 4; t++) freq[t * 257 + 256] = 1; /* src/huffman.c:54 */
    const int16_t *base = zz;
    for (int c = 0; c < 3; c++) {
        int32_t *fdc = freq + (c == 0 ? 0 : 2) * 257, *fac = freq + (c == 0 ? 1 : 3) * 257;
        int32_t *d = malloc(sizeof(int32_t) * (size_t)(n[c] > 0 ? n[c] : 1));
        orc_dc_sequence(base, n[c], dc_mode, d);
        for (long i = 0; i < n[c]; i++) {
            int cls = get_class(d[i]);
            if (cls > 256) cls = 256;
            fdc[cls]++;
            freq_block_ac(fac, base + i * 64, sym_mode);
        }
        free(d);
        base += n[c] * 64;
    }
}

/* ---- canonical Huffman codes (T.81 Annex C) from the tables.h specs --------- */

void orc_huffman_table(int table, uint16_t code[256], uint8_t len[256])
{
    const uint8_t *nr, *vals;
    int nvals, k = 0;
    unsigned c = 0;
    huff_spec(table, &nr, &vals, &nvals);
    memset(code, 0, 256 * sizeof(uint16_t));
    memset(len, 0, 256);
    for (int l = 1; l <= 16; l++) {
        for (int i = 0; i < nr[l]; i++) {
            code[vals[k]] = (uint16_t)c;
            len[vals[k]] = (uint8_t)l;
            c++;
            k++;
        }
        c <<= 1;
    }
    (void)nvals;
}

/* ---- bit writer with stuffing (T.81 F.1.2.3) ------------------------------- */

typedef struct {
    uint8_t *out;
    size_t cap, pos;
    uint64_t acc; /* bits are appended at the low end */
    int nacc;
    uint64_t total_bits;
    int overflow;
    int raw; /* 1: no byte stuffing (orc_encode_scan_raw) */
} BitW;

static void bw_byte(BitW *w, uint8_t v)
{
    if (w->pos + 2 > w->cap) {
        w->overflow = 1;
        return;
    }
    w->out[w->pos++] = v;
    if (v == 0xFF && !w->raw) w->out[w->pos++] = 0x00;
}

static void bw_put(BitW *w, uint32_t bits, int n)
{
    if (n == 0) return;
    w->acc = (w->acc << n) | (bits & ((1u << n) - 1u));
    w->nacc += n;
    w->total_bits += (uint64_t)n;
    while (w->nacc >= 8) {
        bw_byte(w, (uint8_t)(w->acc >> (w->nacc - 8)));
        w->nacc -= 8;
    }
}

static void bw_flush(BitW *w)
{
    if (w->nacc > 0) {
        int pad = 8 - w->nacc;
        uint8_t v = (uint8_t)(((w->acc << pad) | ((1u << pad) - 1u)) & 0xFF);
        bw_byte(w, v);
        w->nacc = 0;
    }
}

typedef struct {
    uint16_t code[4][256];
    uint8_t len[4][256];
} HuffSet;

static void huffset_init(HuffSet *h)
{
    for (int t = 0; t < 4; t++) orc_huffman_table(t, h->code[t], h->len[t]);
}

/* canonical codes (T.81 Annex C) of caller-supplied specs */
static void huffset_from_specs(HuffSet *h, const OrcHuffSpecs *sp)
{
    memset(h, 0, sizeof *h);
    for (int t = 0; t < 4; t++) {
        unsigned c = 0;
        int k = 0;
        for (int l = 1; l <= 16; l++) {
            for (int i = 0; i < sp->nr[t][l]; i++) {
                h->code[t][sp->vals[t][k]] = (uint16_t)c;
                h->len[t][sp->vals[t][k]] = (uint8_t)l;
                c++;
                k++;
            }
            c <<= 1;
        }
    }
}

void orc_specs_standard(OrcHuffSpecs *sp)
{
    memset(sp, 0, sizeof *sp);
    for (int t = 0; t < 4; t++) {
        const uint8_t *nr, *vals;
        int nvals;
        huff_spec(t, &nr, &vals, &nvals);
        memcpy(sp->nr[t], nr, 17);
        memcpy(sp->vals[t], vals, (size_t)nvals);
        sp->nvals[t] = nvals;
    }
}

/* ---- optimal tables: T.81 Annex K.2, the procedure construct_huffman_table
 * (src/huffman.c:76-180) attempts and never finishes.  freq[257] with
 * freq[256] = 1 reserving the all-ones code point (src/huffman.c:54); returns BITS (nr[1..16])
 * and HUFFVAL.  Figures K.1 (code sizes), K.2 (count), K.3 (limit to 16 bits), K.4 (sort). */
static void optimal_table(const int32_t freq_in[257], uint8_t nr[17], uint8_t vals[256], int *nvals)
{
    long freq[257];
    int codesize[257], others[257], bits[34];
    for (int i = 0; i < 257; i++) {
        freq[i] = freq_in[i];
        codesize[i] = 0;
        others[i] = -1;
    }
    freq[256] = 1;
    for (;;) {
        /* least frequency > 0, ties -> largest value; then the next least */
        int v1 = -1, v2 = -1;
        for (int i = 0; i < 257; i++)
            if (freq[i] > 0 && (v1 < 0 || freq[i] <= freq[v1])) v1 = i;
        for (int i = 0; i < 257; i++)
            if (i != v1 && freq[i] > 0 && (v2 < 0 || freq[i] <= freq[v2])) v2 = i;
        if (v2 < 0) break;
        freq[v1] += freq[v2];
        freq[v2] = 0;
        codesize[v1]++;
        while (others[v1] >= 0) {
            v1 = others[v1];
            codesize[v1]++;
        }
        others[v1] = v2;
        codesize[v2]++;
        while (others[v2] >= 0) {
            v2 = others[v2];
            codesize[v2]++;
        }
    }
    memset(bits, 0, sizeof bits);
    for (int i = 0; i < 257; i++)
        if (codesize[i] > 0) bits[codesize[i] > 32 ? 32 : codesize[i]]++;
    int i = 32;
    while (i > 16) {
        while (bits[i] > 0) {
            int j = i - 2;
            while (bits[j] == 0) j--;
            bits[i] -= 2;
            bits[i - 1] += 1;
            bits[j + 1] += 2;
            bits[j] -= 1;
        }
        i--;
    }
    while (bits[i] == 0) i--;
    bits[i]--; /* the reserved code point */
    memset(nr, 0, 17);
    for (int l = 1; l <= 16; l++) nr[l] = (uint8_t)bits[l];
    int k = 0;
    for (int l = 1; l <= 32; l++)
        for (int j = 0; j < 256; j++)
            if (codesize[j] == l) vals[k++] = (uint8_t)j;
    *nvals = k;
}

/* freq[4][257] in the order lum_DC, lum_AC, chrom_DC, chrom_AC (orc_histograms, SPEC modes);
 * specs in the order 0 DC lum, 1 DC chr, 2 AC lum, 3 AC chr */
void orc_specs_optimal(const int32_t *freq, OrcHuffSpecs *sp)
{
    static const int src[4] = {0, 2, 1, 3};
    memset(sp, 0, sizeof *sp);
    for (int t = 0; t < 4; t++) optimal_table(freq + 257 * src[t], sp->nr[t], sp->vals[t], &sp->nvals[t]);
}

/* additional bits: v>0 -> low cat bits of v; v<0 -> low cat bits of v-1 (T.81 F.1.2.1) */
static inline uint32_t extra_bits(int v, int cat) { return (uint32_t)(v < 0 ? v - 1 : v) & ((1u << cat) - 1u); }

/* Encode one block; if w==NULL only count bits. Returns bits. */
static uint32_t encode_block(BitW *w, const HuffSet *h, const int16_t *zz, int pred, int is_chroma)
{
    const uint16_t *dcc = h->code[is_chroma ? 1 : 0], *acc = h->code[is_chroma ? 3 : 2];
    const uint8_t *dcl = h->len[is_chroma ? 1 : 0], *acl = h->len[is_chroma ? 3 : 2];
    uint32_t bits = 0;
    int diff = zz[0] - pred;
    int cat = get_class(diff);
    bits += dcl[cat] + cat;
    if (w) {
        bw_put(w, dcc[cat], dcl[cat]);
        bw_put(w, extra_bits(diff, cat), cat);
    }
    int run = 0;
    for (int k = 1; k < 64; k++) {
        int v = zz[k];
        if (v == 0) {
            run++;
            continue;
        }
        while (run > 15) {
            bits += acl[0xF0];
            if (w) bw_put(w, acc[0xF0], acl[0xF0]);
            run -= 16;
        }
        int c = get_class(v);
        int sym = (run << 4) | c;
        bits += acl[sym] + c;
        if (w) {
            bw_put(w, acc[sym], acl[sym]);
            bw_put(w, extra_bits(v, c), c);
        }
        run = 0;
    }
    if (run > 0) {
        bits += acl[0x00];
        if (w) bw_put(w, acc[0x00], acl[0x00]);
    }
    return bits;
}

/* Walk blocks in interleaved MCU order (T.81 A.2): 4:4:4 -> Y Cb Cr per MCU;
 * 4:2:2 -> Y(2m) Y(2m+1) Cb(m) Cr(m). */
static uint64_t walk_scan_h(BitW *w, const HuffSet *hs, const int16_t *zz, int wpx, int hpx, int sample_ratio,
                            uint32_t *bits_out);
static uint64_t walk_scan(BitW *w, const int16_t *zz, int wpx, int hpx, int sample_ratio, uint32_t *bits_out)
{
    HuffSet h;
    huffset_init(&h);
    return walk_scan_h(w, &h, zz, wpx, hpx, sample_ratio, bits_out);
}

static uint64_t walk_scan_h(BitW *w, const HuffSet *hs, const int16_t *zz, int wpx, int hpx, int sample_ratio,
                            uint32_t *bits_out)
{
    const HuffSet h = *hs;
    long n[3];
    if (orc_block_counts(wpx, hpx, sample_ratio, n)) return 0;
    const int16_t *zy = zz, *zcb = zz + n[0] * 64, *zcr = zz + (n[0] + n[1]) * 64;
    int pred[3] = {0, 0, 0};
    uint64_t total = 0;
    long s = 0;
    long nmcu = n[1];
    for (long m = 0; m < nmcu; m++) {
        int ny = (sample_ratio == 0) ? 1 : (sample_ratio == 1 ? 2 : 4); /* 4:2:0: luma is stored in MCU order */
        for (int k = 0; k < ny; k++) {
            const int16_t *blk = zy + (m * ny + k) * 64;
            uint32_t bb = encode_block(w, &h, blk, pred[0], 0);
            pred[0] = blk[0];
            if (bits_out) bits_out[s] = bb;
            s++;
            total += bb;
        }
        {
            const int16_t *blk = zcb + m * 64;
            uint32_t bb = encode_block(w, &h, blk, pred[1], 1);
            pred[1] = blk[0];
            if (bits_out) bits_out[s] = bb;
            s++;
            total += bb;
        }
        {
            const int16_t *blk = zcr + m * 64;
            uint32_t bb = encode_block(w, &h, blk, pred[2], 1);
            pred[2] = blk[0];
            if (bits_out) bits_out[s] = bb;
            s++;
            total += bb;
        }
    }
    return total;
}

size_t orc_encode_scan(const int16_t *zz, int w, int h, int sample_ratio, uint8_t *out, size_t cap,
                       uint64_t *total_bits)
{
    BitW bw = {out, cap, 0, 0, 0, 0, 0, 0};
    uint64_t t = walk_scan(&bw, zz, w, h, sample_ratio, NULL);
    bw_flush(&bw);
    if (total_bits) *total_bits = t;
    return bw.overflow ? 0 : bw.pos;
}

size_t orc_encode_scan_specs(const int16_t *zz, int w, int h, int sample_ratio, const OrcHuffSpecs *sp, uint8_t *out,
                             size_t cap, uint64_t *total_bits)
{
    HuffSet hs;
    huffset_from_specs(&hs, sp);
    BitW bw = {out, cap, 0, 0, 0, 0, 0, 0};
    uint64_t t = walk_scan_h(&bw, &hs, zz, w, h, sample_ratio, NULL);
    bw_flush(&bw);
    if (total_bits) *total_bits = t;
    return bw.overflow ? 0 : bw.pos;
}

/* The same bit string without byte stuffing and without 1-padding (a trailing partial byte
 * is zero-filled): what the GPU's un-stuffed buffer holds; used by the band-partition tests. */
size_t orc_encode_scan_raw(const int16_t *zz, int w, int h, int sample_ratio, uint8_t *out, size_t cap,
                           uint64_t *total_bits)
{
    BitW bw = {out, cap, 0, 0, 0, 0, 0, 1};
    uint64_t t = walk_scan(&bw, zz, w, h, sample_ratio, NULL);
    if (bw.nacc > 0) {
        bw_byte(&bw, (uint8_t)((bw.acc << (8 - bw.nacc)) & 0xFF));
        bw.nacc = 0;
    }
    if (total_bits) *total_bits = t;
    return bw.overflow ? 0 : bw.pos;
}

void orc_block_bits(const int16_t *zz, int w, int h, int sample_ratio, uint32_t *bits)
{
    walk_scan(NULL, zz, w, h, sample_ratio, bits);
}

/* ---- JFIF markers (T.81 Annex B) ------------------------------ */

static size_t put16(uint8_t *o, size_t p, unsigned v)
{
    o[p] = (uint8_t)(v >> 8);
    o[p + 1] = (uint8_t)v;
    return p + 2;
}

size_t orc_jfif_header(int w, int h, int quality, int sample_ratio, uint8_t *o, size_t cap)
{
    OrcHuffSpecs sp;
    orc_specs_standard(&sp);
    return orc_jfif_header_specs(w, h, quality, sample_ratio, &sp, o, cap);
}

size_t orc_jfif_header_specs(int w, int h, int quality, int sample_ratio, const OrcHuffSpecs *sp, uint8_t *o, size_t cap)
{
    int lum[64], chr[64];
    size_t p = 0;
    if (cap < 700) return 0;
    if (orc_quant_tables(quality, lum, chr)) return 0;
    p = put16(o, p, 0xFFD8);
    /* APP0 JFIF 1.01, aspect 1:1 */
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
    /* DQT: position scan_order[v][u] holds T'[u][v] -- the reference divides the
     * coefficient (u horizontal, v vertical) by table[u][v] (src/quantise.c:58). */
    for (int t = 0; t < 2; t++) {
        const int *tab = t ? chr : lum;
        p = put16(o, p, 0xFFDB);
        p = put16(o, p, 67);
        o[p++] = (uint8_t)t;
        uint8_t zzq[64];
        for (int v = 0; v < 8; v++)
            for (int u = 0; u < 8; u++) zzq[scan_order[v * 8 + u]] = (uint8_t)tab[u * 8 + v];
        memcpy(o + p, zzq, 64);
        p += 64;
    }
    /* SOF0 */
    p = put16(o, p, 0xFFC0);
    p = put16(o, p, 17);
    o[p++] = 8;
    p = put16(o, p, (unsigned)h);
    p = put16(o, p, (unsigned)w);
    o[p++] = 3;
    o[p++] = 1;
    o[p++] = (uint8_t)(sample_ratio == 1 ? 0x21 : (sample_ratio == 2 ? 0x22 : 0x11));
    o[p++] = 0;
    o[p++] = 2;
    o[p++] = 0x11;
    o[p++] = 1;
    o[p++] = 3;
    o[p++] = 0x11;
    o[p++] = 1;
    /* DHT x4: class/id 0x00 DC lum, 0x10 AC lum, 0x01 DC chr, 0x11 AC chr */
    static const int order[4] = {0, 2, 1, 3};
    static const uint8_t tcth[4] = {0x00, 0x10, 0x01, 0x11};
    for (int i = 0; i < 4; i++) {
        const uint8_t *nr = sp->nr[order[i]], *vals = sp->vals[order[i]];
        const int nvals = sp->nvals[order[i]];
        p = put16(o, p, 0xFFC4);
        p = put16(o, p, (unsigned)(2 + 1 + 16 + nvals));
        o[p++] = tcth[i];
        memcpy(o + p, nr + 1, 16);
        p += 16;
        memcpy(o + p, vals, (size_t)nvals);
        p += (size_t)nvals;
    }
    /* SOS */
    p = put16(o, p, 0xFFDA);
    p = put16(o, p, 12);
    o[p++] = 3;
    o[p++] = 1;
    o[p++] = 0x00;
    o[p++] = 2;
    o[p++] = 0x11;
    o[p++] = 3;
    o[p++] = 0x11;
    o[p++] = 0;
    o[p++] = 63;
    o[p++] = 0;
    return p;
}

size_t orc_encode_jfif(const uint8_t *r, const uint8_t *g, const uint8_t *b, int w, int h,
                       int quality, int sample_ratio, const uint8_t *oob, uint8_t *out, size_t cap)
{
    long n[3];
    if (orc_block_counts(w, h, sample_ratio, n)) return 0;
    int16_t *zz = malloc(sizeof(int16_t) * 64 * (size_t)(n[0] + n[1] + n[2]));
    if (!zz) return 0;
    size_t p = 0;
    if (orc_coefficients(r, g, b, w, h, quality, sample_ratio, oob, zz, NULL) == 0) {
        p = orc_jfif_header(w, h, quality, sample_ratio, out, cap);
        if (p) {
            size_t s = orc_encode_scan(zz, w, h, sample_ratio, out + p, cap - p - 2, NULL);
            if (s == 0)
                p = 0;
            else {
                p += s;
                out[p++] = 0xFF;
                out[p++] = 0xD9;
            }
        }
    }
    free(zz);
    return p;
}

/* ---- BMP loader: src/bitmap.c:41-152 ----------------------------------------- */

int orc_bmp_load(const char *path, uint8_t **r, uint8_t **g, uint8_t **b, int *w, int *h)
{
    FILE *fp = fopen(path, "rb");
    if (!fp) return 1;
    fseek(fp, 0, SEEK_END);
    long fs = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    uint8_t *buf = malloc((size_t)fs + 1);
    if (!buf || fread(buf, 1, (size_t)fs, fp) != (size_t)fs) {
        fclose(fp);
        free(buf);
        return 2;
    }
    fclose(fp);
    int32_t width, height;
    int16_t depth;
    memcpy(&width, buf + 18, 4);  /* :70 */
    memcpy(&height, buf + 22, 4); /* :73 */
    memcpy(&depth, buf + 28, 2);  /* :76 */
    long np = (long)width * height;
    *r = malloc((size_t)np);
    *g = malloc((size_t)np);
    *b = malloc((size_t)np);
    long n = 0;
    for (int i = 1; i <= height; i++) {
        long off = fs - (long)i * width * (depth / 8); /* :130: rows taken from the END of the file */
        for (int j = 0; j < width * 3; j += 3) {
            (*r)[n] = buf[off + j]; /* file order is B,G,R: "red" receives blue (:132-134) */
            (*g)[n] = buf[off + j + 1];
            (*b)[n] = buf[off + j + 2];
            n++;
        }
    }
    *w = width;
    *h = height;
    free(buf);
    return 0;
}

void orc_free(void *p) { free(p); }

/* ---- synthetic inputs; mirrored bit-for-bit in csrc/synth.cuh - */

static inline uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

static inline uint64_t synth_key(uint64_t seed, uint32_t frame, int channel)
{
    return splitmix64(seed ^ splitmix64(((uint64_t)frame << 8) | (uint64_t)(channel & 0xff)));
}

static inline uint8_t synth_px(uint64_t key, int kind, int x, int y)
{
    uint64_t hp = splitmix64(key ^ (((uint64_t)(uint32_t)y << 32) | (uint32_t)x));
    if (kind == 1) return (uint8_t)(hp & 0xff);
    uint64_t hc = splitmix64(key ^ 0xC3115EEDULL ^ (((uint64_t)(uint32_t)(y >> 4) << 32) | (uint32_t)(x >> 4)));
    int base = (int)(hc & 0xff);
    int noise = (int)((hp >> 8) % 41u) - 20;
    int v = (base * 4) / 5 + 25 + noise;
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

void orc_synth_plane(uint8_t *dst, int w, int h, uint64_t seed, uint32_t frame, int channel, int kind)
{
    uint64_t key = synth_key(seed, frame, channel);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) dst[(long)y * w + x] = synth_px(key, kind, x, y);
}

/* ---- timing helper for bench.py's cpu_baseline ("port") ---------------------- */

double orc_time_encode(const uint8_t *r, const uint8_t *g, const uint8_t *b, int w, int h, int quality,
                       int sample_ratio)
{
    long n[3];
    struct timespec t0, t1;
    if (orc_block_counts(w, h, sample_ratio, n)) return -1.0;
    size_t nb = (size_t)(n[0] + n[1] + n[2]);
    int16_t *zz = malloc(sizeof(int16_t) * 64 * nb);
    uint8_t *out = malloc((size_t)w * h * 3 + 4096);
    clock_gettime(CLOCK_MONOTONIC, &t0);
    orc_coefficients(r, g, b, w, h, quality, sample_ratio, NULL, zz, NULL);
    size_t len = orc_encode_scan(zz, w, h, sample_ratio, out, (size_t)w * h * 3 + 4096, NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    (void)len;
    free(zz);
    free(out);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
