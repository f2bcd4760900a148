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
    const size_t cap = (size_t)w * h * 3 + 65536;
    unsigned char *out = (unsigned char *)malloc(cap);
    size_t len = 0;
    int rc = out ? jpgb200_encode_planar(ctx, bmp_GetRed(bmp), bmp_GetGreen(bmp), bmp_GetBlue(bmp), w, h, quality,
                                         sample_ratio, out, cap, &len)
                 : JPGB200_ERR_NOMEM;
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
    const size_t cap = (size_t)(w + 16) * (h + 8) * 3 + 65536;
    unsigned char *out = (unsigned char *)malloc(cap);
    size_t len = 0;
    int rc = out ? JPGB200_OK : JPGB200_ERR_NOMEM;
    if (rc == JPGB200_OK) {
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
    const size_t cap = (size_t)width * height * 3 + 65536;
    unsigned char *out = (unsigned char *)malloc(cap);
    size_t len = 0;
    int rc = out ? jpgb200_encode_rgb(ctx, rgb, width, height, quality, sample_ratio, out, cap, &len) : JPGB200_ERR_NOMEM;
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
