#ifndef DOWN_SAMPLE_H
#define DOWN_SAMPLE_H
#include "jpg_encode.h"
/* src/headers/downsample.h: 4:2:2 halves the chroma blocks (the reference only prints) */
void chroma_subsample(JpgData j_data);
#endif
