/*
 * 24-bit BMP reader -- same interface as the reference's src/headers/bitmap.h:14-79, and the
 * same behaviour, quirks included, because the encode path consumes exactly
 * what this loader returns: three planar byte arrays, rows taken from the END of the file,
 * file bytes 0,1,2 of a pixel labelled red, green, blue, no row padding, data offset ignored.
 */
#ifndef BITMAP_H
#define BITMAP_H

#define BMP_SUCCESS 0
#define BMP_FILE_DOESNT_EXIST 1
#define BMP_READ_FAILED 2
#define BMP_FAILED_ALLOCATE_BUFFER 3

typedef struct _bitmap *BmpImage;
typedef unsigned char Byte;

BmpImage bmp_OpenBitmap(const char *file);
Byte *bmp_GetRed(BmpImage b);
Byte *bmp_GetGreen(BmpImage b);
Byte *bmp_GetBlue(BmpImage b);
int bmp_GetWidth(BmpImage b);
int bmp_GetHeight(BmpImage b);
int bmp_GetNumPixels(BmpImage b);
void bmp_ShowBmpInfo(BmpImage b);
int bmp_GetFileSize(BmpImage b);
void bmp_GetLastError(BmpImage b);
void bmp_DestroyBitmap(BmpImage b);

/* not in the reference: the stored error code (the reference only prints it) */
int bmp_GetErrorCode(BmpImage b);
/* not in the reference: strict parsing of a 24-bit BMP (pixel-data offset, 4-byte row pitch,
 * bottom-up rows, B,G,R order): the accessors return true red/green/blue, top-down */
BmpImage bmp_OpenBitmapStrict(const char *file);

#endif
