/*
 * Driver in the shape of the reference's src/jpg_driver.c: with no arguments it runs the same
 * known-answer walk (the Wikipedia 8x8 example through dct_block, quantise_lum and
 * zig_zag_block, src/jpg_driver.c:54-150) -- every stage on the GPU -- and prints it in the
 * same format; with arguments it encodes a file:
 *     jpg <in.bmp> <out.jpg> [quality=50] [sample_ratio=0] [--staged | --robust]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "headers/block.h"
#include "headers/dct.h"
#include "headers/jpg_encode.h"
#include "headers/quantise.h"
#include "headers/zig_zag.h"

JpgData encode_bmp_to_jpeg_staged(const char *input, const char *output, int quality, int sample_ratio);

/* the example block of https://en.wikipedia.org/wiki/JPEG#Discrete_cosine_transform, level-shifted */
static const int wiki[8][8] = {{-76, -73, -67, -62, -58, -67, -64, -55}, {-65, -69, -73, -38, -19, -43, -59, -56},
                               {-66, -69, -60, -15, 16, -24, -62, -55},  {-65, -70, -57, -6, 26, -22, -58, -59},
                               {-61, -67, -60, -24, -2, -40, -60, -58},  {-49, -63, -68, -58, -51, -60, -70, -53},
                               {-43, -57, -64, -69, -73, -67, -63, -45}, {-41, -49, -59, -60, -63, -52, -50, -34}};

static void test_dct(void)
{
    Block b = new_block();
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++) set_value_block(b, x, y, wiki[y][x]);
    show_block(b);
    dct_block(b);
    show_block(b);
    quantise_lum(b);
    show_block(b);
    int a[64] = {0};
    zig_zag_block(b, a);
    printf("Testing zig-zag ordering: \n");
    for (int i = 0; i < 64; i++) printf("%d\n", a[i]);
    destroy_block(b);
}

int main(int argc, char **argv)
{
    if (argc < 3) {
        test_dct();
        return 0;
    }
    const int quality = argc > 3 ? atoi(argv[3]) : 50;
    const int ratio = argc > 4 ? atoi(argv[4]) : NO_CHROMA_SUBSAMPLING;
    if (argc > 5 && !strcmp(argv[5], "--staged")) {
        JpgData j = encode_bmp_to_jpeg_staged(argv[1], argv[2], quality, ratio);
        if (!j || !j->Y) return 1;
        printf("%dx%d blocks Y/Cb/Cr %d/%d/%d\n", j->width, j->height, j->num_blocks_Y, j->num_blocks_Cb, j->num_blocks_Cr);
        return 0;
    }
    if (argc > 5 && !strcmp(argv[5], "--robust")) return encode_bmp_to_jpeg_robust(argv[1], argv[2], quality, ratio) == 0 ? 0 : 1;
    return encode_bmp_to_jpeg_status(argv[1], argv[2], quality, ratio) == 0 ? 0 : 1;
}
