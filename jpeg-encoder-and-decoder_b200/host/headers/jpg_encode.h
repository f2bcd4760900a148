/*
 * Encoder interface -- drop-in for the reference's src/headers/jpg_encode.h:13-99.
 * Same constants, same public JpegData layout (callers may read width/height/num_blocks_*
 * and the per-stage results), same two entry points.  The work behind them runs on a B200
 * through libjpgb200 (include/jpgb200.h); there is no CPU arithmetic path.
 */
#ifndef JPG_ENC
#define JPG_ENC

#include "block.h"

#define NO_CHROMA_SUBSAMPLING 0           /* 4:4:4 */
#define HORIZONTAL_SUBSAMPLING 1          /* 4:2:2 */
#define HORIZONTAL_VERTICAL_SUBSAMPLING 2 /* 4:2:0: a stub in the reference (src/downsample.c:29), rejected here */

typedef struct _jpeg_data *JpgData;
typedef unsigned char Byte;

typedef struct _huffman_data {
    int freq[257]; /* histogram, index = run | size (reference) */
    int code_len[257];
    int others[257];
    int bits[32];
    int huffval[256];
} HuffmanData;

typedef struct _jpeg_data {
    char *output_filename;
    char *input_filename;
    int width;
    int height;
    int sample_ratio;
    int quality;
    int num_blocks_Y;
    int num_blocks_Cb;
    int num_blocks_Cr;
    Block *Y;
    Block *Cb;
    Block *Cr;
    int **zig_zag_Y;
    int **zig_zag_Cb;
    int **zig_zag_Cr;
    HuffmanData lum_DC;
    HuffmanData lum_AC;
    HuffmanData chrom_DC;
    HuffmanData chrom_AC;
} JpegData;

/* BMP file in, JPEG file out (src/jpg_encode.c:19-52).  quality 24..97; sample_ratio one of the
 * constants above.  Errors are reported on stderr, as the void signature leaves no other way. */
void encode_bmp_to_jpeg(const char *input_filename, const char *output_filename, int quality, int sample_ratio);

/* Declared by the reference (src/headers/jpg_encode.h:99) but never defined, and unusable as
 * declared (no dimensions).  Kept for link compatibility: it reports that and returns. */
void encode_rgb_to_jpeg(Byte *colours, const char *output, int quality, int sample_ratio);

/* The usable form: interleaved RGB24, top-down.  Returns 0 or a JPGB200_ERR_* code. */
int encode_rgb_to_jpeg_ex(const Byte *rgb, int width, int height, const char *output, int quality, int sample_ratio);

/* Same as encode_bmp_to_jpeg with a status code (0 = ok, negative = JPGB200_ERR_*, positive = BMP_* error). */
int encode_bmp_to_jpeg_status(const char *input_filename, const char *output_filename, int quality, int sample_ratio);

/* Not in the reference: strict BMP parsing, any size, standards-conforming arithmetic (no
 * bug-compatible quirks) -- same status codes as encode_bmp_to_jpeg_status. */
int encode_bmp_to_jpeg_robust(const char *input_filename, const char *output_filename, int quality, int sample_ratio);

#endif
