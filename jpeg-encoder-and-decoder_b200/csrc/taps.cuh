/*
 * Parity taps and staged-mode kernels: small kernels that expose intermediate results in
 * the reference's own shapes so tests can compare them with the compiled reference.
 *   K2  k2_*            dpcm()                          src/dpcm.c:6-21   (REF recurrence or SPEC diff)
 *   KH  kh_histograms   calculate_freq_block_DC / _AC   src/huffman.c:182-222
 *   KF  kf_*            dct_block / quantise_* / zig_zag_block on FP64 blocks (the driver's
 *                       block-level API, src/jpg_driver.c:135-144)
 *   KS  ks_synth        deterministic synthetic planes
 */
#ifndef JPGB200_TAPS_CUH
#define JPGB200_TAPS_CUH

#include "entropy.cuh"
#include "jpgb200_math.cuh"

namespace jpgb200 {

/* ---- K2: DC sequence ---------------------------------------------------------------
 * REF mode reproduces the in-place forward loop of src/dpcm.c:10-12, i.e. the recurrence
 * d[i] = dc[i] - d[i-1]  <=>  d[i] = s_i * sum_{k<=i} s_k dc[k], s_k = (-1)^k : an
 * inclusive scan of sign-alternated DCs.  Three launches: tile sums, single-CTA scan of the
 * tile sums, tile-local scan + carry.  SPEC mode is the adjacent difference. */
#define K2_THREADS 256
#define K2_ITEMS 4
#define K2_TILE (K2_THREADS * K2_ITEMS)

__device__ __forceinline__ int warp_inclusive_scan_i32(int v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

__global__ void __launch_bounds__(K2_THREADS) k2_tile_sums(const int16_t *zz, long long n, int *tile_sum)
{
    __shared__ int sW[K2_THREADS / 32];
    const int tid = threadIdx.x;
    const long long i0 = (long long)blockIdx.x * K2_TILE + (long long)tid * K2_ITEMS;
    int s = 0;
#pragma unroll
    for (int k = 0; k < K2_ITEMS; k++) {
        const long long i = i0 + k;
        if (i < n) {
            const int dc = zz[i * 64];
            s += (i & 1) ? -dc : dc;
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if ((tid & 31) == 0) sW[tid >> 5] = s;
    __syncthreads();
    if (tid == 0) {
        int t = 0;
        for (int w = 0; w < K2_THREADS / 32; w++) t += sW[w];
        tile_sum[blockIdx.x] = t;
    }
}

/* in-place exclusive scan of int32 by one CTA */
__global__ void __launch_bounds__(1024) k2_scan_tiles(int *tile_sum, long long n)
{
    __shared__ int sW[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int carry = 0;
    for (long long base = 0; base < n; base += 1024) {
        const long long i = base + tid;
        const int v = (i < n) ? tile_sum[i] : 0;
        const int inc = warp_inclusive_scan_i32(v, lane);
        if (lane == 31) sW[wid] = inc;
        __syncthreads();
        int wbase = 0, total = 0;
        for (int w = 0; w < 32; w++) {
            const int s = sW[w];
            if (w < wid) wbase += s;
            total += s;
        }
        if (i < n) tile_sum[i] = carry + wbase + inc - v;
        carry += total;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(K2_THREADS) k2_apply(const int16_t *zz, long long n, const int *tile_base,
                                                       int mode, int first_pred, int *d)
{
    __shared__ int sW[K2_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const long long i0 = (long long)blockIdx.x * K2_TILE + (long long)tid * K2_ITEMS;
    if (mode == 1) { /* SPEC: adjacent difference (T.81 F.1.2.1) */
#pragma unroll
        for (int k = 0; k < K2_ITEMS; k++) {
            const long long i = i0 + k;
            if (i < n) d[i] = (int)zz[i * 64] - (i == 0 ? first_pred : (int)zz[(i - 1) * 64]);
        }
        return;
    }
    int v[K2_ITEMS], s = 0;
#pragma unroll
    for (int k = 0; k < K2_ITEMS; k++) {
        const long long i = i0 + k;
        int dc = (i < n) ? (int)zz[i * 64] : 0;
        s += (i & 1) ? -dc : dc;
        v[k] = s; /* thread-local inclusive sums */
    }
    const int inc = warp_inclusive_scan_i32(s, lane);
    if (lane == 31) sW[wid] = inc;
    __syncthreads();
    int wbase = 0;
    for (int w = 0; w < wid; w++) wbase += sW[w];
    const int base = tile_base[blockIdx.x] + wbase + inc - s;
#pragma unroll
    for (int k = 0; k < K2_ITEMS; k++) {
        const long long i = i0 + k;
        if (i < n) {
            const int t = base + v[k];
            d[i] = (i & 1) ? -t : t;
        }
    }
}

/* ---- KH: symbol histograms ---------------------------------------------------------
 * freq[4][257]: lum_DC, lum_AC, chrom_DC, chrom_AC.  One thread per block of ONE
 * component; d = that component's DC sequence from K2.  sym_mode 0 reproduces the
 * reference's `run | size` index (src/huffman.c:218), 1 is (run<<4)|size. */
__global__ void __launch_bounds__(256) kh_histograms(const int16_t *zz, const int *d, long long n, int is_chroma,
                                                     int sym_mode, int *freq)
{
    __shared__ int sH[2][257];
    const int tid = threadIdx.x;
    for (int i = tid; i < 2 * 257; i += 256) (&sH[0][0])[i] = 0;
    __syncthreads();
    const long long b = (long long)blockIdx.x * 256 + tid;
    if (b < n) {
        const int16_t *z = zz + b * 64;
        int cls = 32 - __clz(abs(d[b]));
        atomicAdd(&sH[0][cls], 1);
        int last = 0;
        for (int i = 63; i > 0; i--)
            if (z[i] != 0) {
                last = i;
                break;
            }
        int run = 0;
        for (int i = 1; i < 64; i++) {
            if (i == last + 1) {
                atomicAdd(&sH[1][0x00], 1);
                break;
            }
            const int v = z[i];
            if (v == 0) {
                if (++run == 16) {
                    atomicAdd(&sH[1][0xF0], 1);
                    run = 0;
                }
            } else {
                const int c = 32 - __clz(abs(v));
                atomicAdd(&sH[1][sym_mode ? ((run << 4) | c) : (run | c)], 1);
                run = 0;
            }
        }
    }
    __syncthreads();
    int *fdc = freq + (is_chroma ? 2 : 0) * 257, *fac = freq + (is_chroma ? 3 : 1) * 257;
    for (int i = tid; i < 257; i += 256) {
        if (sH[0][i]) atomicAdd(&fdc[i], sH[0][i]);
        if (sH[1][i]) atomicAdd(&fac[i], sH[1][i]);
    }
}

/* ---- KF: staged FP64 block kernels (the driver's block-level API) -------------------- */
/* dct_block: one CTA of 64 threads per block, thread (u,v) -> values[v*8+u]; same formula
 * and summation order as src/dct.c:43-56 */
__global__ void __launch_bounds__(64) kf_dct_blocks(double *blocks, long long n)
{
    __shared__ double t[64];
    double *b = blocks + (long long)blockIdx.x * 64;
    const int u = threadIdx.x >> 3, v = threadIdx.x & 7;
    t[threadIdx.x] = b[threadIdx.x];
    __syncthreads();
    const double pi = 3.14159265358979323846;
    double s = 0.0;
    for (int x = 0; x < 8; x++)
        for (int y = 0; y < 8; y++)
            s += t[y * 8 + x] * cos(((2 * x + 1) * u * pi) / 16) * cos(((2 * y + 1) * v * pi) / 16);
    const double au = (u == 0) ? 1 / sqrt(2.0) : 1.0, av = (v == 0) ? 1 / sqrt(2.0) : 1.0;
    b[v * 8 + u] = 0.25 * au * av * s;
}

/* quantise_lum / quantise_chr: values[j*8+i] = round(values[j*8+i] / table[i][j])
 * (src/quantise.c:52-72; round() = half away from zero) */
__global__ void __launch_bounds__(64) kf_quantise_blocks(double *blocks, long long n, const int *table)
{
    double *b = blocks + (long long)blockIdx.x * 64;
    const int j = threadIdx.x >> 3, i = threadIdx.x & 7;
    b[j * 8 + i] = round(b[j * 8 + i] / (double)table[i * 8 + j]);
}

/* zig_zag_block: zz[scan_order[i][j]] = (int) values[i*8+j]  (src/zig_zag.c:54-57) */
__global__ void __launch_bounds__(64) kf_zigzag_blocks(const double *blocks, long long n, const int *scan_order,
                                                       int *zz)
{
    const double *b = blocks + (long long)blockIdx.x * 64;
    zz[(long long)blockIdx.x * 64 + scan_order[threadIdx.x]] = (int)b[threadIdx.x];
}

/* preprocess_jpeg's block fill (src/preprocess.c:155-169,186-188) in FP64 with the reference's
 * operation order (no FMA contraction), including blockToCoords' last-column shift.
 * One thread per sample; blocks[comp][block][y*8+x]. */
__global__ void __launch_bounds__(256) kf_preprocess_blocks(const uint8_t *r, const uint8_t *g, const uint8_t *b, int w,
                                                            int h, const uint8_t *oob, double *Y, double *Cb, double *Cr)
{
    const long long n = (long long)w * h;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int bw = w >> 3;
    const long long blk = i >> 6;
    const int s = (int)(i & 63), x = s & 7, y = s >> 3;
    const int bx = (int)(blk % bw), by = (int)(blk / bw);
    long long row = (long long)by * 8 + y - (bx == bw - 1 ? 1 : 0);
    double rv, gv, bv;
    if (row < 0) {
        rv = oob[x];
        gv = oob[8 + x];
        bv = oob[16 + x];
    } else {
        const long long o = row * w + (long long)bx * 8 + x;
        rv = r[o];
        gv = g[o];
        bv = b[o];
    }
    const double yv = __dadd_rn(__dadd_rn(__dmul_rn(0.299, rv), __dmul_rn(0.587, gv)), __dmul_rn(0.114, bv));
    const double cb = 128.0 - __dadd_rn(__dadd_rn(__dmul_rn(0.168736, rv), -__dmul_rn(0.331264, gv)), __dmul_rn(0.5, bv));
    const double cr = 128.0 + __dadd_rn(__dadd_rn(__dmul_rn(0.5, rv), -__dmul_rn(0.418688, gv)), -__dmul_rn(0.081312, bv));
    Y[i] = yv - 128.0;
    Cb[i] = cb - 128.0;
    Cr[i] = cr - 128.0;
}

/* 4:2:2: mean of horizontally adjacent samples of two neighbouring
 * full-resolution chroma blocks -> one half-resolution block */
__global__ void __launch_bounds__(64) kf_subsample422_blocks(const double *in, double *out, int bw)
{
    const long long ob = blockIdx.x; /* output block: row-major over (bw/2) x bh */
    const int cbw = bw >> 1;
    const long long by = ob / cbw, cx = ob % cbw;
    const int x = threadIdx.x & 7, y = threadIdx.x >> 3;
    const double *src = in + ((by * bw + 2 * cx + (x >> 2)) * 64) + y * 8 + (x & 3) * 2;
    out[ob * 64 + y * 8 + x] = 0.5 * (src[0] + src[1]);
}

/* dpcm() on a contiguous int32 DC array (staged API): REF recurrence via the same 3-kernel
 * scan as K2 */
__global__ void __launch_bounds__(K2_THREADS) k2s_tile_sums(const int *dc, long long n, int *tile_sum)
{
    __shared__ int sW[K2_THREADS / 32];
    const int tid = threadIdx.x;
    const long long i0 = (long long)blockIdx.x * K2_TILE + (long long)tid * K2_ITEMS;
    int s = 0;
#pragma unroll
    for (int k = 0; k < K2_ITEMS; k++) {
        const long long i = i0 + k;
        if (i < n) s += (i & 1) ? -dc[i] : dc[i];
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if ((tid & 31) == 0) sW[tid >> 5] = s;
    __syncthreads();
    if (tid == 0) {
        int t = 0;
        for (int w = 0; w < K2_THREADS / 32; w++) t += sW[w];
        tile_sum[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(K2_THREADS) k2s_apply(const int *dc, long long n, const int *tile_base, int *d)
{
    __shared__ int sW[K2_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const long long i0 = (long long)blockIdx.x * K2_TILE + (long long)tid * K2_ITEMS;
    int v[K2_ITEMS], s = 0;
#pragma unroll
    for (int k = 0; k < K2_ITEMS; k++) {
        const long long i = i0 + k;
        const int x = (i < n) ? dc[i] : 0;
        s += (i & 1) ? -x : x;
        v[k] = s;
    }
    const int inc = warp_inclusive_scan_i32(s, lane);
    if (lane == 31) sW[wid] = inc;
    __syncthreads();
    int wbase = 0;
    for (int w = 0; w < wid; w++) wbase += sW[w];
    const int base = tile_base[blockIdx.x] + wbase + inc - s;
#pragma unroll
    for (int k = 0; k < K2_ITEMS; k++) {
        const long long i = i0 + k;
        if (i < n) {
            const int t = base + v[k];
            d[i] = (i & 1) ? -t : t;
        }
    }
}

/* calculate_freq_block_DC/AC on int32 zig-zag blocks whose [0] already holds the dpcm output */
__global__ void __launch_bounds__(256) khs_histograms(const int *zz, long long n, int sym_mode, int *fdc, int *fac)
{
    __shared__ int sH[2][257];
    const int tid = threadIdx.x;
    for (int i = tid; i < 2 * 257; i += 256) (&sH[0][0])[i] = 0;
    __syncthreads();
    const long long b = (long long)blockIdx.x * 256 + tid;
    if (b < n) {
        const int *z = zz + b * 64;
        int cls = 32 - __clz(abs(z[0]));
        if (cls > 256) cls = 256;
        atomicAdd(&sH[0][cls], 1);
        int last = 0;
        for (int i = 63; i > 0; i--)
            if (z[i] != 0) {
                last = i;
                break;
            }
        int run = 0;
        for (int i = 1; i < 64; i++) {
            if (i == last + 1) {
                atomicAdd(&sH[1][0x00], 1);
                break;
            }
            const int v = z[i];
            if (v == 0) {
                if (++run == 16) {
                    atomicAdd(&sH[1][0xF0], 1);
                    run = 0;
                }
            } else {
                const int c = 32 - __clz(abs(v));
                atomicAdd(&sH[1][(sym_mode ? ((run << 4) | c) : (run | c)) & 0xff], 1);
                run = 0;
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < 257; i += 256) {
        if (sH[0][i]) atomicAdd(&fdc[i], sH[0][i]);
        if (sH[1][i]) atomicAdd(&fac[i], sH[1][i]);
    }
}

/* ---- KS: synthetic planes ------------------------------------------------------------- */
/* rows y0 .. y0 + h - 1 of the synthetic image into planes [3][h][w] (y0 = 0: the whole image) */
__global__ void __launch_bounds__(256) ks_synth(uint8_t *planes, int w, int h, unsigned long long seed,
                                                unsigned frame, int kind, int y0)
{
    const int c = blockIdx.z;
    const unsigned long long key = synth_key(seed, frame, c);
    uint8_t *dst = planes + (long long)c * w * h;
    const long long npx4 = (long long)w * h / 4;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < npx4; q += stride) {
        const long long i = q * 4;
        const int y = y0 + (int)(i / w), x = (int)(i % w); /* w % 8 == 0: the 4 pixels share a row */
        uint32_t v = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) v |= (uint32_t)synth_px(key, kind, x + k, y) << (8 * k);
        reinterpret_cast<uint32_t *>(dst)[q] = v;
    }
}

/* Range check of caller-supplied coefficients (jpgb200_encode_coefficients, the staged huffman_encode): the
 * entropy kernels assume what K1 guarantees -- AC categories <= 10 and DC-difference categories <= 11 (T.81 tables
 * F.1 / F.2) -- and index their code tables with them.  One thread per block; *bad becomes 1 on any violation. */
__global__ void __launch_bounds__(256) k_validate_coefs(const int16_t *zz, long long n0, long long n1, long long n2, int *bad)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x, nb = n0 + n1 + n2;
    if (i >= nb) return;
    const uint4 *row = reinterpret_cast<const uint4 *>(zz + i * 64);
    bool ok = true;
    int dc = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const uint4 v = row[j];
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int lo = (int)(short)(w[k] & 0xffffu), hi = (int)(short)(w[k] >> 16);
            if (j == 0 && k == 0)
                dc = lo;
            else
                ok = ok && lo >= -1023 && lo <= 1023;
            ok = ok && hi >= -1023 && hi <= 1023;
        }
    }
    const bool first = (i == 0 || i == n0 || i == n0 + n1);
    const int prev = first ? 0 : (int)zz[(i - 1) * 64];
    const int diff = dc - prev;
    ok = ok && diff >= -2047 && diff <= 2047;
    if (!ok) *bad = 1;
}

/* interleaved RGB24 -> planes (for encode_rgb_to_jpeg_ex) */
__global__ void __launch_bounds__(256) ks_deinterleave(const uint8_t *rgb, uint8_t *r, uint8_t *g, uint8_t *b,
                                                       long long npx)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += stride) {
        r[i] = rgb[3 * i];
        g[i] = rgb[3 * i + 1];
        b[i] = rgb[3 * i + 2];
    }
}

/* three tight planes w x h -> three planes pw x ph, edge samples replicated to the right and below */
__global__ void __launch_bounds__(256) ks_pad_planes(const uint8_t *src, uint8_t *dst, int w, int h, int pw, int ph)
{
    const long long per = (long long)pw * ph, total = 3 * per;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i / per);
        const long long rem = i - c * per;
        const int y = (int)(rem / pw), x = (int)(rem - (long long)y * pw);
        const int sy = y < h ? y : h - 1, sx = x < w ? x : w - 1;
        dst[i] = src[(long long)c * w * h + (long long)sy * w + sx];
    }
}

}  // namespace jpgb200
#endif
