#include <stdio.h>
#include <stdlib.h>

#include "host_ctx.h"

static jpgb200_ctx *g_ctx = NULL;
static int g_tried = 0;

static void release(void)
{
    if (g_ctx) jpgb200_destroy(g_ctx);
    g_ctx = NULL;
}

jpgb200_ctx *host_ctx(void)
{
    if (!g_ctx && !g_tried) {
        g_tried = 1;
        const char *d = getenv("JPGB200_DEVICE");
        const int rc = jpgb200_create(&g_ctx, d ? atoi(d) : 0);
        if (rc != JPGB200_OK) {
            fprintf(stderr, "jpgb200: %s\n", jpgb200_strerror(rc));
            g_ctx = NULL;
        } else {
            /* JPGB200_OPTIMIZED_HUFFMAN=1: per-image optimal tables -- what the reference's
             * construct_huffman_table (src/huffman.c:76-180) sets out to build -- instead of the
             * standard tables of src/headers/tables.h */
            const char *o = getenv("JPGB200_OPTIMIZED_HUFFMAN");
            if (o && atoi(o) != 0) jpgb200_set_optimized_huffman(g_ctx, 1);
            atexit(release);
        }
    }
    return g_ctx;
}

const char *host_ctx_error(void) { return g_ctx ? jpgb200_last_error(g_ctx) : "no context"; }
