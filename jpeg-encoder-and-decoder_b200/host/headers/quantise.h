#ifndef QUANTISE_H
#define QUANTISE_H
#include "block.h"
#include "jpg_encode.h"
#define TABLE_SIZE 8
/* src/headers/quantise.h.  quantise_lum/chr use the CURRENT tables: the q50 base tables until
 * quantise(j) has scaled them, exactly like the reference's globals (src/quantise.c:8-25,34-35);
 * unlike the reference, scaling always starts from the pristine tables. */
void quantise_lum(Block b);
void quantise_chr(Block b);
void quantise(JpgData j_data);
#endif
