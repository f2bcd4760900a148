/*
 * Encoder entry points of the reference (src/jpg_encode.c:19-58) on top of libjpgb200.
 * encode_bmp_to_jpeg = BMP load (bitmap.c) -> jpgb200_encode_planar (the fused GPU path:
 * K1 transform, entropy stage, stuffing, JFIF markers) -> fwrite.  The per-stage functions
 * the reference chains here are available in stages.c for callers of the stage-level API.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "headers/bitmap.h"
#include "headers/jpg_encode.h"
#include "host_ctx.h"

static int write_file(const char *path, const unsigned char *data, size_t len)
{
    FILE *f = fopen(path, "wb");
    if (!f) {
        fprintf(stderr, "jpg_encode: cannot open %s for writing\n", path);
        return JPGB200_ERR_BAD_ARG;
    }
    const size_t w = fwrite(data, 1, len, f);
    fclose(f);
    return w == len ? JPGB200_OK : JPGB200_ERR_BAD_ARG;
}

/* Output capacity: a first guess of 3 B/pixel (a natural image needs ~0.3) and, when the library answers
 * JPGB200_ERR_OUT_TOO_SMALL (noise at the highest qualities codes to 4 B/pixel and more), the true worst case:
 * every block at its longest (1658 bits: a 20-bit DC code + 63 x 26-bit AC codes) and every byte stuffed. */
static size_t out_cap(int w, int h, int attempt)
{
    const size_t pw = (size_t)(w + 15) / 16 * 16, ph = (size_t)(h + 15) / 16 * 16;
    if (attempt == 0) return pw * ph * 3 + 65536;
    return 1024 + (pw / 8) * (ph / 8) * 3 * ((1658 + 7) / 8) * 2 + 2;
}

int encode_bmp_to_jpeg_status(const char *input, const char *output, int quality, int sample_ratio)
{
    jpgb200_ctx *ctx = host_ctx();
    if (!ctx) return JPGB200_ERR_NO_DEVICE;
    BmpImage bmp = bmp_OpenBitmap(input);
    if (!bmp) return BMP_FAILED_ALLOCATE_BUFFER;
    if (bmp_GetErrorCode(bmp) != BMP_SUCCESS) {
        const int e = bmp_GetErrorCode(bmp);
        bmp_GetLastError(bmp);
        bmp_DestroyBitmap(bmp);
        return e;
    }
    const int w = bmp_GetWidth(bmp), h = bmp_GetHeight(bmp);
    unsigned char *out = NULL;
    size_t len = 0;
    int rc = JPGB200_ERR_OUT_TOO_SMALL;
    for (int attempt = 0; attempt < 2 && rc == JPGB200_ERR_OUT_TOO_SMALL; attempt++) {
        const size_t cap = out_cap(w, h, attempt);
        free(out);
        out = (unsigned char *)malloc(cap);
        rc = out ? jpgb200_encode_planar(ctx, bmp_GetRed(bmp), bmp_GetGreen(bmp), bmp_GetBlue(bmp), w, h, quality, sample_ratio, out,
                                         cap, &len)
                 : JPGB200_ERR_NOMEM;
    }
    bmp_DestroyBitmap(bmp);
    if (rc == JPGB200_OK) rc = write_file(output, out, len);
    free(out);
    return rc;
}

/* The robust front door (the reference's README.md:16-24 "current issues"):
 * strict BMP parsing, any width and height, standards-conforming arithmetic -- the file decodes to
 * the picture in the BMP in any JPEG decoder. */
int encode_bmp_to_jpeg_robust(const char *input, const char *output, int quality, int sample_ratio)
{
    jpgb200_ctx *ctx = host_ctx();
    if (!ctx) return JPGB200_ERR_NO_DEVICE;
    BmpImage bmp = bmp_OpenBitmapStrict(input);
    if (!bmp) return BMP_FAILED_ALLOCATE_BUFFER;
    if (bmp_GetErrorCode(bmp) != BMP_SUCCESS) {
        const int e = bmp_GetErrorCode(bmp);
        bmp_GetLastError(bmp);
        bmp_DestroyBitmap(bmp);
        return e;
    }
    const int w = bmp_GetWidth(bmp), h = bmp_GetHeight(bmp);
    unsigned char *out = NULL;
    size_t len = 0;
    int rc = JPGB200_ERR_OUT_TOO_SMALL;
    for (int attempt = 0; attempt < 2 && rc == JPGB200_ERR_OUT_TOO_SMALL; attempt++) {
        const size_t cap = out_cap(w, h, attempt);
        free(out);
        out = (unsigned char *)malloc(cap);
        if (!out) {
            rc = JPGB200_ERR_NOMEM;
            break;
        }
        jpgb200_set_standard_mode(ctx, 1);
        rc = jpgb200_encode_planar_any(ctx, bmp_GetRed(bmp), bmp_GetGreen(bmp), bmp_GetBlue(bmp), w, h, quality, sample_ratio,
                                       out, cap, &len);
        jpgb200_set_standard_mode(ctx, 0);
    }
    bmp_DestroyBitmap(bmp);
    if (rc == JPGB200_OK) rc = write_file(output, out, len);
    free(out);
    return rc;
}

void encode_bmp_to_jpeg(const char *input, const char *output, int quality, int sample_ratio)
{
    const int rc = encode_bmp_to_jpeg_status(input, output, quality, sample_ratio);
    if (rc < 0) fprintf(stderr, "encode_bmp_to_jpeg: %s (%s)\n", jpgb200_strerror(rc), host_ctx_error());
}

int encode_rgb_to_jpeg_ex(const Byte *rgb, int width, int height, const char *output, int quality, int sample_ratio)
{
    jpgb200_ctx *ctx = host_ctx();
    if (!ctx) return JPGB200_ERR_NO_DEVICE;
    unsigned char *out = NULL;
    size_t len = 0;
    int rc = JPGB200_ERR_OUT_TOO_SMALL;
    for (int attempt = 0; attempt < 2 && rc == JPGB200_ERR_OUT_TOO_SMALL; attempt++) {
        const size_t cap = out_cap(width, height, attempt);
        free(out);
        out = (unsigned char *)malloc(cap);
        rc = out ? jpgb200_encode_rgb(ctx, rgb, width, height, quality, sample_ratio, out, cap, &len) : JPGB200_ERR_NOMEM;
    }
    if (rc == JPGB200_OK) rc = write_file(output, out, len);
    free(out);
    return rc;
}

void encode_rgb_to_jpeg(Byte *colours, const char *output, int quality, int sample_ratio)
{
    (void)colours;
    (void)output;
    (void)quality;
    (void)sample_ratio;
    fprintf(stderr, "encode_rgb_to_jpeg: the reference declares this without image dimensions and never "
                    "defines it; use encode_rgb_to_jpeg_ex(rgb, width, height, ...)\n");
}
