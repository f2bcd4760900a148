/*
 * oracle/jpgb200_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the JPEG encode hot path of
 * matthewT53/JPEG-Encoder-and-Decoder.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library;
 * the product path (libjpgb200.so) never links or calls it.
 *
 * Parity status
 *   - coefficient path (BMP quirks, colour, block addressing, FP64 DCT,
 *     transposed quantisation, zig-zag, dpcm recurrence, run|size histograms):
 *     PINNED against the unmodified reference compiled into oracle/_ref/
 *     (see oracle/Makefile, oracle/ref_harness.c) and against the committed
 *     golden vectors in tests/golden/.
 *   - 4:2:2 averaging, true DC differences, (run<<4)|size symbols, bit packing,
 *     byte stuffing and JFIF markers: PARITY UNPINNED by the reference (it has
 *     no bitstream writer).  They follow ITU-T T.81 and are
 *     checked by two independent decoders (PIL, OpenCV) in tests/.
 */
#ifndef JPGB200_ORACLE_H
#define JPGB200_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_DC_REF_RECURRENCE 0 /* d[i] = dc[i] - d[i-1]   (src/dpcm.c:10-12)      */
#define ORC_DC_SPEC_DIFF 1      /* d[i] = dc[i] - dc[i-1]  (T.81 F.1.2.1)           */
#define ORC_SYM_REF 0           /* AC symbol = run | size  (src/huffman.c:218)      */
#define ORC_SYM_SPEC 1          /* AC symbol = (run<<4)|size (T.81 F.1.2.2)         */

/* Scaled tables T'[i][j], row-major, exactly as scale_table leaves the globals
 * after the FIRST call (src/quantise.c:74-86).  Returns 0, or -1 if the quality
 * produces a zero divisor or an entry above 255 (usable range 24..97). */
int orc_quant_tables(int quality, int lum[64], int chr[64]);

/* Number of 8x8 blocks per component for (w,h,sample_ratio): n[0]=Y n[1]=Cb n[2]=Cr */
int orc_block_counts(int w, int h, int sample_ratio, long n[3]);

/* Quantised zig-zag coefficients, layout [Y blocks | Cb blocks | Cr blocks][64].
 * r,g,b are the planes as bmp_GetRed/Green/Blue deliver them.
 * oob holds the 8 bytes the reference reads in front of each of the three plane
 * buffers for block W/8-1 (oob[0..7]=r, [8..15]=g, [16..23]=b); NULL = zeros.
 * If dct_out != NULL it receives the un-quantised DCT doubles values[v*8+u]. */
int orc_coefficients(const uint8_t *r, const uint8_t *g, const uint8_t *b, int w, int h, int quality,
                     int sample_ratio, const uint8_t *oob, int16_t *zz, double *dct_out);

/* worker threads used by orc_coefficients (default 1: the reference is single-threaded) */
void orc_set_threads(int t);
/* 1: standards-conforming arithmetic (textbook Cb, no last-column shift) for the checks of the
 * product's standard mode; 0 (default): the reference's behaviour */
void orc_set_standard(int on);

/* single-block entry points mirroring dct_block / quantise_lum|chr / zig_zag_block */
void orc_dct_block(double values[64]);
void orc_quantise_block(double values[64], const int table[64]);
void orc_zigzag_block(const double values[64], int zz[64]);

/* DC sequence of one component (n blocks, stride 64 int16) -> d[n] */
void orc_dc_sequence(const int16_t *zz, long n, int mode, int32_t *d);

/* Histograms freq[4][257] in the order lum_DC, lum_AC, chrom_DC, chrom_AC
 * (src/huffman.c:23-48,182-222).  dc_mode/sym_mode select REF or SPEC. */
void orc_histograms(const int16_t *zz, const long n[3], int dc_mode, int sym_mode, int32_t *freq);

/* Canonical Huffman code tables derived from the Annex-K specs the reference
 * carries in src/headers/tables.h:37-72.  table: 0 DC lum, 1 DC chr, 2 AC lum,
 * 3 AC chr.  code[256], len[256] (len 0 = symbol absent). */
void orc_huffman_table(int table, uint16_t code[256], uint8_t len[256]);

/* Huffman table specifications (BITS = nr[1..16], HUFFVAL = vals) in the order 0 DC lum,
 * 1 DC chr, 2 AC lum, 3 AC chr: the standard ones of src/headers/tables.h, or per-image optimal
 * ones built by T.81 Annex K.2 from the SPEC-mode histograms (freq[4][257] in orc_histograms'
 * order) -- what construct_huffman_table (src/huffman.c:76-180) set out to do. */
typedef struct {
    uint8_t nr[4][17];
    uint8_t vals[4][256];
    int nvals[4];
} OrcHuffSpecs;
void orc_specs_standard(OrcHuffSpecs *sp);
void orc_specs_optimal(const int32_t *freq, OrcHuffSpecs *sp);
size_t orc_encode_scan_specs(const int16_t *zz, int w, int h, int sample_ratio, const OrcHuffSpecs *sp, uint8_t *out,
                             size_t cap, uint64_t *total_bits);
size_t orc_jfif_header_specs(int w, int h, int quality, int sample_ratio, const OrcHuffSpecs *sp, uint8_t *out,
                             size_t cap);

/* Entropy-coded segment only (SPEC DC + SPEC symbols, interleaved MCUs, 1-padding,
 * 0xFF00 stuffing).  Returns bytes written, or 0 if cap is too small.
 * If total_bits != NULL it receives the un-stuffed, un-padded bit count. */
size_t orc_encode_scan(const int16_t *zz, int w, int h, int sample_ratio, uint8_t *out, size_t cap,
                       uint64_t *total_bits);

/* Same bit string, neither stuffed nor 1-padded (trailing partial byte zero-filled). */
size_t orc_encode_scan_raw(const int16_t *zz, int w, int h, int sample_ratio, uint8_t *out, size_t cap,
                           uint64_t *total_bits);

/* Per-block code lengths in bits, scan (MCU-interleaved) order. */
void orc_block_bits(const int16_t *zz, int w, int h, int sample_ratio, uint32_t *bits);

/* JFIF header (SOI..SOS) for the given parameters.  Returns its length (623). */
size_t orc_jfif_header(int w, int h, int quality, int sample_ratio, uint8_t *out, size_t cap);

/* Whole file: header, scan, EOI. */
size_t orc_encode_jfif(const uint8_t *r, const uint8_t *g, const uint8_t *b, int w, int h,
                       int quality, int sample_ratio, const uint8_t *oob, uint8_t *out, size_t cap);

/* BMP loader with the reference's quirks (src/bitmap.c:41-152).  Planes are
 * malloc'd; free with orc_free. Returns 0 on success. */
int orc_bmp_load(const char *path, uint8_t **r, uint8_t **g, uint8_t **b, int *w, int *h);
void orc_free(void *p);

/* Deterministic synthetic planes shared with the CUDA generator
 * (kind 0 = "natural", 1 = "noise"). */
void orc_synth_plane(uint8_t *dst, int w, int h, uint64_t seed, uint32_t frame, int channel, int kind);

/* Time the FP64 path (coefficients + dc + histograms) on one image; returns seconds. */
double orc_time_encode(const uint8_t *r, const uint8_t *g, const uint8_t *b, int w, int h, int quality,
                       int sample_ratio);

#ifdef __cplusplus
}
#endif
#endif
