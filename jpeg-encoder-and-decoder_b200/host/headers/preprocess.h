#ifndef PREPROCESS_H
#define PREPROCESS_H
#include "jpg_encode.h"
/* src/headers/preprocess.h: load the BMP, set dimensions, fill the level-shifted Y/Cb/Cr blocks */
void preprocess_jpeg(JpgData j_data);
#endif
