/*
 * BMP loader with the reference's observable behaviour (src/bitmap.c:41-152), written afresh:
 *  - width, height, bits per pixel are taken from byte offsets 18, 22 and 28; the pixel-data
 *    offset field is read but not used
 *  - pixel row i (1-based, from the top of the image) starts at
 *        file_size - i * width * (bpp / 8)
 *    i.e. rows are counted back from the END of the file, with no 4-byte row padding, so any
 *    trailing bytes shift every row (tiger.bmp has 2)
 *  - the three bytes of a pixel go, in file order, to the planes called red, green, blue
 *    (a real BMP stores B,G,R, so "red" holds blue)
 * The file is read once (the reference reads it twice); the result is identical.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "headers/bitmap.h"

#define NAME_MAX_LEN 500

struct _bitmap {
    char filename[NAME_MAX_LEN];
    int fileSize, offsetRGB, width, height;
    short bitDepth;
    int numPixels, error;
    Byte *red, *green, *blue;
};

static int split_planes(BmpImage b, const Byte *file)
{
    const long np = (long)b->width * b->height;
    const int bytes_pp = b->bitDepth / 8;
    if (np <= 0) return BMP_READ_FAILED;
    b->red = (Byte *)malloc((size_t)np);
    b->green = (Byte *)malloc((size_t)np);
    b->blue = (Byte *)malloc((size_t)np);
    if (!b->red || !b->green || !b->blue) return BMP_FAILED_ALLOCATE_BUFFER;
    long n = 0;
    for (int row = 1; row <= b->height; row++) {
        const long start = (long)b->fileSize - (long)row * b->width * bytes_pp;
        if (start < 0 || start + 3L * b->width > b->fileSize) return BMP_READ_FAILED; /* the reference would read out of bounds */
        const Byte *px = file + start;
        for (int x = 0; x < b->width; x++, px += 3, n++) {
            b->red[n] = px[0];
            b->green[n] = px[1];
            b->blue[n] = px[2];
        }
    }
    return BMP_SUCCESS;
}

BmpImage bmp_OpenBitmap(const char *file)
{
    BmpImage b = (BmpImage)calloc(1, sizeof(struct _bitmap));
    if (!b) return NULL;
    strncpy(b->filename, file, NAME_MAX_LEN - 1);
    FILE *fp = fopen(file, "rb");
    if (!fp) {
        b->error = BMP_FILE_DOESNT_EXIST; /* the reference calls fclose(NULL) here and crashes */
        return b;
    }
    fseek(fp, 0, SEEK_END);
    const long fs = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    Byte *buf = (fs >= 54 && fs < 0x7fffffffL) ? (Byte *)malloc((size_t)fs) : NULL;
    if (!buf || fread(buf, 1, (size_t)fs, fp) != (size_t)fs) {
        b->error = BMP_READ_FAILED;
        fclose(fp);
        free(buf);
        return b;
    }
    fclose(fp);
    b->fileSize = (int)fs;
    memcpy(&b->offsetRGB, buf + 10, 4);
    memcpy(&b->width, buf + 18, 4);
    memcpy(&b->height, buf + 22, 4);
    memcpy(&b->bitDepth, buf + 28, 2);
    b->numPixels = b->width * b->height;
    b->error = split_planes(b, buf);
    free(buf);
    return b;
}

/* Strict parsing of an uncompressed 24-bit BMP -- what the reference's README lists under "current
 * issues" (README.md:16-24): the pixel-data offset is honoured, rows are 4-byte aligned,
 * bottom-up files (positive height) are flipped, and the B,G,R bytes go to the planes they name.
 * The accessors then return true red / green / blue, top-down. */
BmpImage bmp_OpenBitmapStrict(const char *file)
{
    BmpImage b = (BmpImage)calloc(1, sizeof(struct _bitmap));
    if (!b) return NULL;
    strncpy(b->filename, file, NAME_MAX_LEN - 1);
    FILE *fp = fopen(file, "rb");
    if (!fp) {
        b->error = BMP_FILE_DOESNT_EXIST;
        return b;
    }
    fseek(fp, 0, SEEK_END);
    const long fs = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    Byte *buf = (fs >= 54 && fs < 0x7fffffffL) ? (Byte *)malloc((size_t)fs) : NULL;
    if (!buf || fread(buf, 1, (size_t)fs, fp) != (size_t)fs) {
        b->error = BMP_READ_FAILED;
        fclose(fp);
        free(buf);
        return b;
    }
    fclose(fp);
    int compression = 0, hgt = 0;
    b->fileSize = (int)fs;
    memcpy(&b->offsetRGB, buf + 10, 4);
    memcpy(&b->width, buf + 18, 4);
    memcpy(&hgt, buf + 22, 4);
    memcpy(&b->bitDepth, buf + 28, 2);
    memcpy(&compression, buf + 30, 4);
    const int top_down = hgt < 0;
    b->height = top_down ? -hgt : hgt;
    b->numPixels = b->width * b->height;
    const long pitch = ((long)b->width * 3 + 3) & ~3L;
    if (buf[0] != 'B' || buf[1] != 'M' || b->bitDepth != 24 || compression != 0 || b->width <= 0 || b->height <= 0 ||
        b->offsetRGB < 54 || (long)b->offsetRGB + pitch * b->height > fs) {
        b->error = BMP_READ_FAILED;
        free(buf);
        return b;
    }
    const long np = (long)b->width * b->height;
    b->red = (Byte *)malloc((size_t)np);
    b->green = (Byte *)malloc((size_t)np);
    b->blue = (Byte *)malloc((size_t)np);
    if (!b->red || !b->green || !b->blue) {
        b->error = BMP_FAILED_ALLOCATE_BUFFER;
        free(buf);
        return b;
    }
    for (int y = 0; y < b->height; y++) {
        const Byte *px = buf + b->offsetRGB + pitch * (top_down ? y : b->height - 1 - y);
        Byte *r = b->red + (long)y * b->width, *g = b->green + (long)y * b->width, *bl = b->blue + (long)y * b->width;
        for (int x = 0; x < b->width; x++, px += 3) {
            bl[x] = px[0];
            g[x] = px[1];
            r[x] = px[2];
        }
    }
    b->error = BMP_SUCCESS;
    free(buf);
    return b;
}

Byte *bmp_GetRed(BmpImage b) { return b->red; }
Byte *bmp_GetGreen(BmpImage b) { return b->green; }
Byte *bmp_GetBlue(BmpImage b) { return b->blue; }
int bmp_GetWidth(BmpImage b) { return b->width; }
int bmp_GetHeight(BmpImage b) { return b->height; }
int bmp_GetNumPixels(BmpImage b) { return b->numPixels; }
int bmp_GetFileSize(BmpImage b) { return b->fileSize; }
int bmp_GetErrorCode(BmpImage b) { return b ? b->error : BMP_FAILED_ALLOCATE_BUFFER; }

void bmp_ShowBmpInfo(BmpImage b)
{
    printf("[BMP Info]:\n");
    printf("Filename: %s\n", b->filename);
    printf("File size: %d kb\n", b->fileSize / 1000);
    printf("Height: %d\n", b->height);
    printf("Width: %d\n", b->width);
    printf("Bit depth: %d\n", b->bitDepth);
}

void bmp_GetLastError(BmpImage b)
{
    switch (b->error) {
    case BMP_FILE_DOESNT_EXIST: printf("%s doesn't exist\n", b->filename); break;
    case BMP_READ_FAILED: printf("Reading BMP image failed\n"); break;
    case BMP_FAILED_ALLOCATE_BUFFER: printf("Failed to allocate a buffer.\n"); break;
    default: printf("No error.\n"); break;
    }
}

void bmp_DestroyBitmap(BmpImage b)
{
    if (!b) return;
    free(b->red);
    free(b->green);
    free(b->blue);
    free(b);
}
