#ifndef ZIG_ZAG_H
#define ZIG_ZAG_H
#include "block.h"
#include "jpg_encode.h"
/* src/headers/zig_zag.h (the reference's per-block printf is not reproduced) */
void zig_zag(JpgData j_data);
void zig_zag_block(Block b, int *zz);
#endif
