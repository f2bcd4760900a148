#ifndef DCT_H
#define DCT_H
#include "block.h"
#include "jpg_encode.h"
/* src/headers/dct.h */
void dct(JpgData j_data);
void dct_block(Block b);
#endif
