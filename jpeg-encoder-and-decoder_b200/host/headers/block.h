/*
 * 8x8 block container -- same interface as the reference's src/headers/block.h:10-40
 * (opaque handle, values addressed (x, y) with x horizontal), so callers of the block-level
 * API (src/jpg_driver.c:54-150) link unchanged.
 */
#ifndef BLOCK_H
#define BLOCK_H

typedef struct _block *Block;

Block new_block();                                    /* uninitialised, like the reference   */
double get_value_block(Block b, int x, int y);        /* values[y*8 + x]                     */
void set_value_block(Block b, int x, int y, double v);
Block copy_block(Block b);
void show_block(Block b);                             /* 8 rows of "%8.2f "                  */
void destroy_block(Block b);

/* not in the reference: contiguous access for the GPU shims */
double *block_values(Block b);

#endif
