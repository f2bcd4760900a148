/*
 * Per-block arithmetic of the fused transform kernel (K1), written once as
 * __host__ __device__ so that tests/ can compile the very same code with g++ and
 * compare it with the oracle before any GPU time is spent (tests/host_math_check.cpp).
 * The host build is a test harness only; the product has no CPU path.
 *
 * What it replaces in the reference (citations relative to /root/reference/):
 *   colour + level shift : src/preprocess.c:160-162,186-188   (FP64 there, FP32 here)
 *   FDCT                 : src/dct.c:36-59  (naive O(N^4) FP64; here a separable scaled
 *                          8-point butterfly after Arai/Agui/Nakajima, 5 mul + 29 add)
 *   quantise             : src/quantise.c:52-72 (divide + round(); here one FFMA against a
 *                          reciprocal table that also carries the butterfly's output scale
 *                          and the reference's TRANSPOSED table orientation)
 *   zig-zag              : src/zig_zag.c:48-59 (static register permutation here)
 * Tolerance contract (BASELINE.json north_star): quantised coefficients within +-1 LSB of
 * the reference on <= 0.1 % of coefficients.
 */
#ifndef JPGB200_MATH_CUH
#define JPGB200_MATH_CUH

#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define JPGB_HD __host__ __device__ __forceinline__
#else
#define JPGB_HD static inline
#endif

namespace jpgb200 {

/* output scale of the 8-point butterfly relative to the orthonormal JPEG DCT:
 * out_aan[k] = true[k] * 8^(1/2)... folded in pairs: 2-D factor = 8*s[u]*s[v] */
static const double kAanScale[8] = {1.0,         1.387039845, 1.306562965, 1.175875602,
                                    1.0,         0.785694958, 0.541196100, 0.275899379};

JPGB_HD float jfma(float a, float b, float c)
{
#ifdef __CUDA_ARCH__
    return __fmaf_rn(a, b, c);
#else
    return fmaf(a, b, c);
#endif
}

/* ---- colour conversion + level shift ------------------------------------------
 * Reference formulas INCLUDING its Cb sign slip (src/preprocess.c:161):
 *   Y  = 0.299 r + 0.587 g + 0.114 b - 128
 *   Cb = (128 - (0.168736 r - 0.331264 g + 0.5 b)) - 128
 *   Cr = (128 + (0.5 r - 0.418688 g - 0.081312 b)) - 128
 * r,g,b are the planes as bmp_GetRed/Green/Blue deliver them. */
JPGB_HD float colour_y(float r, float g, float b) { return jfma(0.114f, b, jfma(0.587f, g, jfma(0.299f, r, -128.0f))); }
JPGB_HD float colour_cb(float r, float g, float b) { return jfma(-0.5f, b, jfma(0.331264f, g, -0.168736f * r)); }
JPGB_HD float colour_cr(float r, float g, float b) { return jfma(-0.081312f, b, jfma(-0.418688f, g, 0.5f * r)); }
/* 4:2:2: mean of two horizontally adjacent chroma samples == formula applied to the
 * (exact, integer) channel sums with halved weights */
JPGB_HD float colour_cb2(float rs, float gs, float bs) { return jfma(-0.25f, bs, jfma(0.165632f, gs, -0.084368f * rs)); }
JPGB_HD float colour_cr2(float rs, float gs, float bs) { return jfma(-0.040656f, bs, jfma(-0.209344f, gs, 0.25f * rs)); }

/* ---- scaled 8-point forward DCT (in place on 8 registers) ---------------------- */
JPGB_HD void fdct8(float &d0, float &d1, float &d2, float &d3, float &d4, float &d5, float &d6, float &d7)
{
    const float c4 = 0.707106781f, c6 = 0.382683433f, a2 = 0.541196100f, a4 = 1.306562965f;
    float t0 = d0 + d7, t7 = d0 - d7;
    float t1 = d1 + d6, t6 = d1 - d6;
    float t2 = d2 + d5, t5 = d2 - d5;
    float t3 = d3 + d4, t4 = d3 - d4;
    /* even part */
    float e0 = t0 + t3, e3 = t0 - t3;
    float e1 = t1 + t2, e2 = t1 - t2;
    d0 = e0 + e1;
    d4 = e0 - e1;
    float s = e2 + e3;
    d2 = jfma(c4, s, e3);
    d6 = jfma(-c4, s, e3);
    /* odd part */
    float o0 = t4 + t5, o1 = t5 + t6, o2 = t6 + t7;
    float z5 = (o0 - o2) * c6;
    float z2 = jfma(a2, o0, z5);
    float z4 = jfma(a4, o2, z5);
    float z11 = jfma(c4, o1, t7);
    float z13 = jfma(-c4, o1, t7);
    d5 = z13 + z2;
    d3 = z13 - z2;
    d1 = z11 + z4;
    d7 = z11 - z4;
}

/* 2-D transform of f[y*8+x] in place; result f[v*8+u] (v vertical frequency),
 * scaled by 8*s[u]*s[v]. */
JPGB_HD void fdct8x8(float *f)
{
#pragma unroll
    for (int y = 0; y < 8; y++)
        fdct8(f[y * 8 + 0], f[y * 8 + 1], f[y * 8 + 2], f[y * 8 + 3], f[y * 8 + 4], f[y * 8 + 5], f[y * 8 + 6], f[y * 8 + 7]);
#pragma unroll
    for (int x = 0; x < 8; x++)
        fdct8(f[0 * 8 + x], f[1 * 8 + x], f[2 * 8 + x], f[3 * 8 + x], f[4 * 8 + x], f[5 * 8 + x], f[6 * 8 + x], f[7 * 8 + x]);
}

#ifdef __CUDACC__
/* ---- packed FP32x2 forms (sm_100a FADD2 / FMUL2 / FFMA2: two IEEE-RN FP32 operations per
 * issue slot, immediates broadcast to both halves).  Every operation below is the scalar
 * function's operation applied to two independent instances, in the same order, so the results
 * are bit-identical to fdct8x8 / colour_* above -- only the instruction count halves. */
typedef float2 f32x2;
__device__ __forceinline__ f32x2 p_bc(float x) { return make_float2(x, x); }
__device__ __forceinline__ f32x2 p_add(f32x2 a, f32x2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ f32x2 p_sub(f32x2 a, f32x2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ f32x2 p_mul(f32x2 a, f32x2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f32x2 p_fma(f32x2 a, f32x2 b, f32x2 c) { return __ffma2_rn(a, b, c); }

__device__ __forceinline__ f32x2 colour_y_p(f32x2 r, f32x2 g, f32x2 b)
{
    return p_fma(p_bc(0.114f), b, p_fma(p_bc(0.587f), g, p_fma(p_bc(0.299f), r, p_bc(-128.0f))));
}
__device__ __forceinline__ f32x2 colour_cb_p(f32x2 r, f32x2 g, f32x2 b)
{
    return p_fma(p_bc(-0.5f), b, p_fma(p_bc(0.331264f), g, p_mul(p_bc(-0.168736f), r)));
}
/* the textbook Cb (JFIF): -0.168736 r - 0.331264 g + 0.5 b, for the standards-conforming front door */
__device__ __forceinline__ f32x2 colour_cb_std_p(f32x2 r, f32x2 g, f32x2 b)
{
    return p_fma(p_bc(0.5f), b, p_fma(p_bc(-0.331264f), g, p_mul(p_bc(-0.168736f), r)));
}
__device__ __forceinline__ f32x2 colour_cb2_std_p(f32x2 rs, f32x2 gs, f32x2 bs)
{
    return p_fma(p_bc(0.25f), bs, p_fma(p_bc(-0.165632f), gs, p_mul(p_bc(-0.084368f), rs)));
}
__device__ __forceinline__ f32x2 colour_cr_p(f32x2 r, f32x2 g, f32x2 b)
{
    return p_fma(p_bc(-0.081312f), b, p_fma(p_bc(-0.418688f), g, p_mul(p_bc(0.5f), r)));
}
__device__ __forceinline__ f32x2 colour_cb2_p(f32x2 rs, f32x2 gs, f32x2 bs)
{
    return p_fma(p_bc(-0.25f), bs, p_fma(p_bc(0.165632f), gs, p_mul(p_bc(-0.084368f), rs)));
}
__device__ __forceinline__ f32x2 colour_cr2_p(f32x2 rs, f32x2 gs, f32x2 bs)
{
    return p_fma(p_bc(-0.040656f), bs, p_fma(p_bc(-0.209344f), gs, p_mul(p_bc(0.25f), rs)));
}

/* 4:2:0: mean of a 2x2 group == formula applied to the (exact, integer) sums of four with
 * quarter weights */
__device__ __forceinline__ f32x2 colour_cb4_p(f32x2 rs, f32x2 gs, f32x2 bs)
{
    return p_fma(p_bc(-0.125f), bs, p_fma(p_bc(0.082816f), gs, p_mul(p_bc(-0.042184f), rs)));
}
__device__ __forceinline__ f32x2 colour_cb4_std_p(f32x2 rs, f32x2 gs, f32x2 bs)
{
    return p_fma(p_bc(0.125f), bs, p_fma(p_bc(-0.082816f), gs, p_mul(p_bc(-0.042184f), rs)));
}
__device__ __forceinline__ f32x2 colour_cr4_p(f32x2 rs, f32x2 gs, f32x2 bs)
{
    return p_fma(p_bc(-0.020328f), bs, p_fma(p_bc(-0.104672f), gs, p_mul(p_bc(0.125f), rs)));
}

/* fdct8 after its first butterfly stage (t0..t7 given) */
__device__ __forceinline__ void fdct8_tail_p(f32x2 t0, f32x2 t1, f32x2 t2, f32x2 t3, f32x2 t4, f32x2 t5, f32x2 t6, f32x2 t7,
                                             f32x2 &d0, f32x2 &d1, f32x2 &d2, f32x2 &d3, f32x2 &d4, f32x2 &d5, f32x2 &d6,
                                             f32x2 &d7)
{
    const float c4 = 0.707106781f, c6 = 0.382683433f, a2 = 0.541196100f, a4 = 1.306562965f;
    f32x2 e0 = p_add(t0, t3), e3 = p_sub(t0, t3);
    f32x2 e1 = p_add(t1, t2), e2 = p_sub(t1, t2);
    d0 = p_add(e0, e1);
    d4 = p_sub(e0, e1);
    f32x2 s = p_add(e2, e3);
    d2 = p_fma(p_bc(c4), s, e3);
    d6 = p_fma(p_bc(-c4), s, e3);
    f32x2 o0 = p_add(t4, t5), o1 = p_add(t5, t6), o2 = p_add(t6, t7);
    f32x2 z5 = p_mul(p_sub(o0, o2), p_bc(c6));
    f32x2 z2 = p_fma(p_bc(a2), o0, z5);
    f32x2 z4 = p_fma(p_bc(a4), o2, z5);
    f32x2 z11 = p_fma(p_bc(c4), o1, t7);
    f32x2 z13 = p_fma(p_bc(-c4), o1, t7);
    d5 = p_add(z13, z2);
    d3 = p_sub(z13, z2);
    d1 = p_add(z11, z4);
    d7 = p_sub(z11, z4);
}
__device__ __forceinline__ void fdct8_p(f32x2 &d0, f32x2 &d1, f32x2 &d2, f32x2 &d3, f32x2 &d4, f32x2 &d5, f32x2 &d6,
                                        f32x2 &d7)
{
    f32x2 t0 = p_add(d0, d7), t7 = p_sub(d0, d7);
    f32x2 t1 = p_add(d1, d6), t6 = p_sub(d1, d6);
    f32x2 t2 = p_add(d2, d5), t5 = p_sub(d2, d5);
    f32x2 t3 = p_add(d3, d4), t4 = p_sub(d3, d4);
    fdct8_tail_p(t0, t1, t2, t3, t4, t5, t6, t7, d0, d1, d2, d3, d4, d5, d6, d7);
}

/* 2-D transform of one block held as row pairs P[k][x] = (f[k][x], f[7-k][x]), k = 0..3.
 * Pass A transforms the eight rows two at a time.  The column pass's first butterfly stage
 * combines exactly the two rows of a pair (t_k = F_k + F_{7-k}, t_{7-k} = F_k - F_{7-k}), so
 * it is done with scalar adds that write their results already paired over the horizontal
 * frequency -- no register transposition anywhere.  Result Q[v][h] = (G[v][2h], G[v][2h+1]),
 * the natural order f[v*8+u] of fdct8x8, bit-identical to it. */
__device__ __forceinline__ void fdct8x8_p(f32x2 (&P)[4][8], f32x2 (&Q)[8][4])
{
#pragma unroll
    for (int k = 0; k < 4; k++) fdct8_p(P[k][0], P[k][1], P[k][2], P[k][3], P[k][4], P[k][5], P[k][6], P[k][7]);
#pragma unroll
    for (int h = 0; h < 4; h++) {
        f32x2 t[8];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const f32x2 a = P[k][2 * h], b = P[k][2 * h + 1];
            t[k] = make_float2(a.x + a.y, b.x + b.y);
            t[7 - k] = make_float2(a.x - a.y, b.x - b.y);
        }
        fdct8_tail_p(t[0], t[1], t[2], t[3], t[4], t[5], t[6], t[7], Q[0][h], Q[1][h], Q[2][h], Q[3][h], Q[4][h], Q[5][h],
                     Q[6][h], Q[7][h]);
    }
}

/* bytes -> exact FP32 without a conversion instruction: PRMT builds the bit pattern of
 * 2^23 + byte, one FADD2 removes the 2^23 from two values */
#define JPGB_BYTE_MAGIC 0x4B000000u
__device__ __forceinline__ f32x2 bytes_to_f32x2(uint32_t wa, uint32_t wb, int byte_index)
{
    /* (byte `byte_index` of wa, same byte of wb) */
    /* magic first: the selector stays an immediate and the magic word one loop-invariant register */
    const uint32_t sel = 0x3104u + (uint32_t)byte_index;
    f32x2 v;
    v.x = __uint_as_float(__byte_perm(JPGB_BYTE_MAGIC, wa, sel));
    v.y = __uint_as_float(__byte_perm(JPGB_BYTE_MAGIC, wb, sel));
    return p_add(v, p_bc(-8388608.0f));
}
/* same for 16-bit halves (pair sums of two bytes, <= 510): half 0 or 1 of wa / wb */
__device__ __forceinline__ f32x2 halves_to_f32x2(uint32_t wa, uint32_t wb, int half_index)
{
    const uint32_t sel = half_index ? 0x3176u : 0x3154u;
    f32x2 v;
    v.x = __uint_as_float(__byte_perm(JPGB_BYTE_MAGIC, wa, sel));
    v.y = __uint_as_float(__byte_perm(JPGB_BYTE_MAGIC, wb, sel));
    return p_add(v, p_bc(-8388608.0f));
}
#endif /* __CUDACC__ */

/* ---- quantise: one FFMA per coefficient ---------------------------------------
 * q = coef * m + 1.5*2^23 rounds (to nearest, ties to even) at unit granularity; the
 * low 16 bits of the float's bit pattern are then the two's-complement integer.
 * The reference rounds half away from zero (round(), src/quantise.c:58); the two rules
 * differ only on exact ties of the FP64 quotient, which its own FP64 noise already makes
 * arbitrary -- covered by the +-1 LSB tolerance and counted in tests. */
#define JPGB_ROUND_MAGIC 12582912.0f

JPGB_HD uint32_t fbits(float x)
{
#ifdef __CUDA_ARCH__
    return __float_as_uint(x);
#else
    union { float f; uint32_t u; } c;
    c.f = x;
    return c.u;
#endif
}

JPGB_HD uint32_t pack16(float lo, float hi)
{
#ifdef __CUDA_ARCH__
    return __byte_perm(__float_as_uint(lo), __float_as_uint(hi), 0x5410);
#else
    return (fbits(lo) & 0xffffu) | (fbits(hi) << 16);
#endif
}
#define JPGB_PACK(a, b) pack16((a), (b))

/* f[v*8+u]: quantised-with-magic floats in natural order -> 32 words holding the 64
 * int16 zig-zag coefficients (src/zig_zag.c:6-15, inverted at compile time) */
JPGB_HD void zigzag_pack(const float *f, uint32_t *o)
{
    o[0] = JPGB_PACK(f[0], f[1]);
    o[1] = JPGB_PACK(f[8], f[16]);
    o[2] = JPGB_PACK(f[9], f[2]);
    o[3] = JPGB_PACK(f[3], f[10]);
    o[4] = JPGB_PACK(f[17], f[24]);
    o[5] = JPGB_PACK(f[32], f[25]);
    o[6] = JPGB_PACK(f[18], f[11]);
    o[7] = JPGB_PACK(f[4], f[5]);
    o[8] = JPGB_PACK(f[12], f[19]);
    o[9] = JPGB_PACK(f[26], f[33]);
    o[10] = JPGB_PACK(f[40], f[48]);
    o[11] = JPGB_PACK(f[41], f[34]);
    o[12] = JPGB_PACK(f[27], f[20]);
    o[13] = JPGB_PACK(f[13], f[6]);
    o[14] = JPGB_PACK(f[7], f[14]);
    o[15] = JPGB_PACK(f[21], f[28]);
    o[16] = JPGB_PACK(f[35], f[42]);
    o[17] = JPGB_PACK(f[49], f[56]);
    o[18] = JPGB_PACK(f[57], f[50]);
    o[19] = JPGB_PACK(f[43], f[36]);
    o[20] = JPGB_PACK(f[29], f[22]);
    o[21] = JPGB_PACK(f[15], f[23]);
    o[22] = JPGB_PACK(f[30], f[37]);
    o[23] = JPGB_PACK(f[44], f[51]);
    o[24] = JPGB_PACK(f[58], f[59]);
    o[25] = JPGB_PACK(f[52], f[45]);
    o[26] = JPGB_PACK(f[38], f[31]);
    o[27] = JPGB_PACK(f[39], f[46]);
    o[28] = JPGB_PACK(f[53], f[60]);
    o[29] = JPGB_PACK(f[61], f[54]);
    o[30] = JPGB_PACK(f[47], f[55]);
    o[31] = JPGB_PACK(f[62], f[63]);
}

/* level-shifted samples f[y*8+x] -> 32 packed zig-zag words; mul[v*8+u] from
 * make_quant_multipliers() */
JPGB_HD void transform_block(float *f, const float *mul, uint32_t *o)
{
    fdct8x8(f);
#pragma unroll
    for (int i = 0; i < 64; i++) f[i] = jfma(f[i], mul[i], JPGB_ROUND_MAGIC);
    zigzag_pack(f, o);
}

/* Host: scaled tables exactly as the reference's first scale_table call leaves them
 * (src/quantise.c:74-86): s = q<50 ? 5000/q : 200-2q ; T' = (s*T+50)/100, integer. */
static inline int scale_quant_entry(int t, int quality)
{
    int s = (quality < 50) ? 5000 / quality : 200 - 2 * quality;
    return (s * t + 50) / 100;
}

/* Host: multiplier for natural position (v,u).  The reference divides the coefficient
 * with horizontal frequency u and vertical frequency v by table[u][v] (src/quantise.c:58:
 * transposed orientation), so mul[v][u] = 1 / (T'[u][v] * 8 * s[u] * s[v]). */
static inline void make_quant_multipliers(const int *scaled_table /*[8][8]*/, float *mul /*[v*8+u]*/)
{
    for (int v = 0; v < 8; v++)
        for (int u = 0; u < 8; u++)
            mul[v * 8 + u] = (float)(1.0 / ((double)scaled_table[u * 8 + v] * 8.0 * kAanScale[u] * kAanScale[v]));
}

/* ---- synthetic inputs; bit-identical to oracle/jpgb200_oracle.c -- */
JPGB_HD uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
JPGB_HD uint64_t synth_key(uint64_t seed, uint32_t frame, int channel)
{
    return splitmix64(seed ^ splitmix64(((uint64_t)frame << 8) | (uint64_t)(channel & 0xff)));
}
JPGB_HD uint8_t synth_px(uint64_t key, int kind, int x, int y)
{
    uint64_t hp = splitmix64(key ^ (((uint64_t)(uint32_t)y << 32) | (uint32_t)x));
    if (kind == 1) return (uint8_t)(hp & 0xff);
    uint64_t hc = splitmix64(key ^ 0xC3115EEDULL ^ (((uint64_t)(uint32_t)(y >> 4) << 32) | (uint32_t)(x >> 4)));
    int base = (int)(hc & 0xff);
    int noise = (int)((hp >> 8) % 41u) - 20;
    int v = (base * 4) / 5 + 25 + noise;
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

}  // namespace jpgb200
#endif
