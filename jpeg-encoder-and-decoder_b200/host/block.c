/* 8x8 FP64 block container: host-side handle only; the arithmetic on blocks runs on the GPU
 * (stages.c).  Interface of the reference's src/block.c. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "headers/block.h"

struct _block {
    double values[64];
};

Block new_block() { return (Block)malloc(sizeof(struct _block)); }

double get_value_block(Block b, int x, int y) { return b->values[(y << 3) + x]; }

void set_value_block(Block b, int x, int y, double v) { b->values[(y << 3) + x] = v; }

Block copy_block(Block b)
{
    Block c = new_block();
    if (c) memcpy(c->values, b->values, sizeof c->values);
    return c;
}

void show_block(Block b)
{
    for (int y = 0; y < 8; y++) {
        printf("\n");
        for (int x = 0; x < 8; x++) printf("%8.2f ", b->values[(y << 3) + x]);
    }
    printf("\n");
}

void destroy_block(Block b) { free(b); }

double *block_values(Block b) { return b->values; }
