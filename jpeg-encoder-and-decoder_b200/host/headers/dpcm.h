#ifndef DPCM_H
#define DPCM_H
#include "jpg_encode.h"
/* src/headers/dpcm.h: in place, d[i] = dc[i] - d[i-1] -- the recurrence the reference computes */
void dpcm(JpgData jpg_data);
#endif
