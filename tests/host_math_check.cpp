// tests/host_math_check.cpp -- TEST HARNESS ONLY (never shipped, never a fallback).
// Compiles csrc/jpgb200_math.cuh with g++ so the FP32 block arithmetic of kernel K1 can
// be compared with the FP64 oracle on the CPU, before any GPU time is spent.
#include <cstdint>
#include <cstring>
#include "../jpeg-encoder-and-decoder_b200/csrc/jpgb200_math.cuh"
#include "../jpeg-encoder-and-decoder_b200/csrc/jpgb200_tables.h"
using namespace jpgb200;

static inline int sample(const uint8_t *p, int w, int bw, int bx, int by, int x, int y, const uint8_t *oob)
{
    long row = (long)by * 8 + y - (bx == bw - 1 ? 1 : 0);
    if (row < 0) return oob ? oob[x] : 0;
    return p[row * w + (long)bx * 8 + x];
}

extern "C" int hm_coefficients(const uint8_t *r, const uint8_t *g, const uint8_t *b, int w, int h, int quality,
                               int ratio, const uint8_t *oob, int16_t *zz)
{
    int lum[64], chr[64];
    float ml[64], mc[64];
    for (int i = 0; i < 64; i++) {
        lum[i] = scale_quant_entry(kBaseLum[i], quality);
        chr[i] = scale_quant_entry(kBaseChr[i], quality);
    }
    make_quant_multipliers(lum, ml);
    make_quant_multipliers(chr, mc);
    int bw = w / 8, bh = h / 8;
    long ny = (long)bw * bh, nc = ratio ? ny / 2 : ny;
    int16_t *zy = zz, *zcb = zz + ny * 64, *zcr = zz + (ny + nc) * 64;
    for (int by = 0; by < bh; by++) {
        for (int bx = 0; bx < bw; bx++) {
            float Y[64], Cb[64], Cr[64];
            uint32_t o[32];
            for (int y = 0; y < 8; y++)
                for (int x = 0; x < 8; x++) {
                    float rv = (float)sample(r, w, bw, bx, by, x, y, oob);
                    float gv = (float)sample(g, w, bw, bx, by, x, y, oob ? oob + 8 : 0);
                    float bv = (float)sample(b, w, bw, bx, by, x, y, oob ? oob + 16 : 0);
                    Y[y * 8 + x] = colour_y(rv, gv, bv);
                    Cb[y * 8 + x] = colour_cb(rv, gv, bv);
                    Cr[y * 8 + x] = colour_cr(rv, gv, bv);
                }
            long bi = (long)by * bw + bx;
            transform_block(Y, ml, o);
            memcpy(zy + bi * 64, o, 128);
            if (!ratio) {
                transform_block(Cb, mc, o);
                memcpy(zcb + bi * 64, o, 128);
                transform_block(Cr, mc, o);
                memcpy(zcr + bi * 64, o, 128);
            }
        }
        if (ratio) {
            for (int cx = 0; cx < bw / 2; cx++) {
                float Cb[64], Cr[64];
                uint32_t o[32];
                for (int y = 0; y < 8; y++)
                    for (int x = 0; x < 8; x++) {
                        int bx = 2 * cx + (x >> 2), xx = (x & 3) * 2;
                        float rs = (float)(sample(r, w, bw, bx, by, xx, y, oob) + sample(r, w, bw, bx, by, xx + 1, y, oob));
                        float gs = (float)(sample(g, w, bw, bx, by, xx, y, oob ? oob + 8 : 0) + sample(g, w, bw, bx, by, xx + 1, y, oob ? oob + 8 : 0));
                        float bs = (float)(sample(b, w, bw, bx, by, xx, y, oob ? oob + 16 : 0) + sample(b, w, bw, bx, by, xx + 1, y, oob ? oob + 16 : 0));
                        Cb[y * 8 + x] = colour_cb2(rs, gs, bs);
                        Cr[y * 8 + x] = colour_cr2(rs, gs, bs);
                    }
                long ci = (long)by * (bw / 2) + cx;
                transform_block(Cb, mc, o);
                memcpy(zcb + ci * 64, o, 128);
                transform_block(Cr, mc, o);
                memcpy(zcr + ci * 64, o, 128);
            }
        }
    }
    return 0;
}

extern "C" void hm_synth_plane(uint8_t *dst, int w, int h, uint64_t seed, uint32_t frame, int channel, int kind)
{
    uint64_t key = synth_key(seed, frame, channel);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) dst[(long)y * w + x] = synth_px(key, kind, x, y);
}
