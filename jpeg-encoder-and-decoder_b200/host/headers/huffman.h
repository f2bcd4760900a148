#ifndef HUFFMAN_H
#define HUFFMAN_H
#include "jpg_encode.h"
/* src/headers/huffman.h.  Fills the four freq[] histograms like the reference's loops
 * (src/huffman.c:29-42); instead of the non-terminating construct_huffman_table it then writes
 * the JPEG file to j_data->output_filename with the Annex-K tables of src/headers/tables.h. */
void huffman_encode(JpgData j_data);
void calculate_freq_block_DC(HuffmanData *huffman_data, int *image_data);
void calculate_freq_block_AC(HuffmanData *huffman_data, int *image_data);
#endif
