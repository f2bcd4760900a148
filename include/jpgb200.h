/*
 * include/jpgb200.h -- C ABI of libjpgb200.so, the B200 (sm_100a) JPEG encode hot path.
 *
 * This is the drop-in boundary for the encode path of matthewT53/JPEG-Encoder-and-Decoder:
 * plain pointers and sizes, no CUDA or torch types.  The host-side mirror of the
 * reference's own C API (encode_bmp_to_jpeg, the per-stage JpgData functions, bmp_*) lives
 * in jpeg-encoder-and-decoder_b200/host/ and calls only the functions declared here.
 * Each entry point cites the reference interface it replaces (relative to
 * /root/reference/).  There is no CPU fallback: every function returns
 * JPGB200_ERR_NO_DEVICE when no CUDA device is usable.
 *
 * Conventions
 *   - r,g,b are the three u8 planes exactly as bmp_GetRed/Green/Blue deliver them
 *     (src/headers/bitmap.h:34-44; top-down raster, no row padding).
 *   - sample_ratio: 0 = NO_CHROMA_SUBSAMPLING (4:4:4), 1 = HORIZONTAL_SUBSAMPLING (4:2:2),
 *     2 = HORIZONTAL_VERTICAL_SUBSAMPLING (4:2:0) (src/headers/jpg_encode.h:13-15).  The
 *     reference's 4:2:2 and 4:2:0 are printf stubs (src/downsample.c:24-32): both are defined
 *     here as box averages of the level-shifted chroma.  For 4:2:0 the luma blocks of the
 *     coefficient layout are in MCU order (Y00 Y01 Y10 Y11 per 16x16 MCU); the MCU-row partition
 *     (jpgb200_band_*) and the single-launch entropy kernel do not take it.
 *   - width % 8 == 0 (4:4:4) or % 16 == 0 (4:2:2, 4:2:0), height % 8 == 0 (% 16 for 4:2:0): the
 *     reference's padding path is broken (src/preprocess.c:83-98,131-152), so other sizes are
 *     rejected -- jpgb200_encode_planar_any takes any size.
 *   - quality 24..97: outside it the reference's scale_table yields zero divisors or
 *     entries above 255 (src/quantise.c:81-83).
 *   - "device" pointers are CUDA device addresses on the context's GPU, passed as plain
 *     pointer values.
 */
#ifndef JPGB200_H
#define JPGB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JPGB200_OK 0
#define JPGB200_ERR_NO_DEVICE (-1)
#define JPGB200_ERR_BAD_DIMS (-2)
#define JPGB200_ERR_BAD_QUALITY (-3)
#define JPGB200_ERR_CUDA (-4)
#define JPGB200_ERR_OUT_TOO_SMALL (-5)
#define JPGB200_ERR_BAD_ARG (-6)
#define JPGB200_ERR_NOMEM (-7)
#define JPGB200_ERR_NCCL (-8)

#define JPGB200_DC_REF_RECURRENCE 0 /* d[i]=dc[i]-d[i-1], what dpcm() computes (src/dpcm.c:10-12) */
#define JPGB200_DC_SPEC_DIFF 1      /* d[i]=dc[i]-dc[i-1], what a decodable stream needs        */
#define JPGB200_SYM_REF 0           /* AC histogram index run|size (src/huffman.c:218)          */
#define JPGB200_SYM_SPEC 1          /* (run<<4)|size                                            */

#define JPGB200_NUM_KERNELS 6 /* K1 transform, K35 single-launch entropy, KE encode chunks, KF count 0xFF, KS scatter, K1E fused transform + chunk encode */

typedef struct jpgb200_ctx jpgb200_ctx;

/* ---- lifetime ---------------------------------------------------------------------- */
/* One context per GPU (or per host thread); owns a stream and all device workspaces.
 * Replaces create_jpeg_data (src/jpg_encode.c:54-58), whose JpegData is never freed. */
int jpgb200_create(jpgb200_ctx **ctx, int device);
void jpgb200_destroy(jpgb200_ctx *ctx);
const char *jpgb200_strerror(int code);
/* text of the last CUDA error seen by this context ("" if none) */
const char *jpgb200_last_error(const jpgb200_ctx *ctx);
/* Run on a caller-owned CUDA stream (cudaStream_t passed as void*); NULL = context's own. */
int jpgb200_set_stream(jpgb200_ctx *ctx, void *cuda_stream);
int jpgb200_sync(jpgb200_ctx *ctx);
/* 1 when the library was built with the A/B forms (jpeg-encoder-and-decoder_b200/build.py with JPGB200_AB=1):
 * the three older K1 kernels (forms 2..4 below) and the single-launch entropy kernel K35.  They produce the
 * same coefficients and bytes as the product kernels and exist for cross-checks and A/B measurements; the
 * default build leaves them out, and selecting them then returns JPGB200_ERR_BAD_ARG. */
int jpgb200_has_ab_forms(void);
/* Entropy stage selection: 1 (default) = the staged pipeline KE (chunk encode) / KF (0xFF count) /
 * KS (stuffed scatter); 0 = the single-launch kernel K35 (count -> chained scan -> pack -> stuff),
 * byte-identical and slower, kept as a cross-check; K35 takes 4:4:4 and 4:2:2 only and disables the
 * batch overlap and the fused transform below. */
int jpgb200_set_staged_entropy(jpgb200_ctx *ctx, int on);
/* KE's form.  The chunk encoder (what src/huffman.c:182-235 does per block, with the bit writer the reference
 * lacks) keeps every block's code string in a small private region of shared memory and lives on occupancy:
 *   1 = tight: 5-word regions at 4:4:4 (8 at 4:2:2): seven (five) CTAs per SM; a block with a longer string is
 *       coded a second time -- right for ordinary images, slow when most blocks are that busy
 *   2 = roomy: 10-word regions, six (four) CTAs per SM
 *   0 = automatic (default): either form counts the blocks whose string outgrows the tight region, the host reads
 *       that counter every fourth launch without ever waiting for it, and switches to the roomy form above 5 % of
 *       the blocks and back below 2 %.
 * Identical bytes in every mode.  jpgb200_entropy_form returns the form the next launch would use (1 or 2). */
int jpgb200_set_entropy_form(jpgb200_ctx *ctx, int mode);
int jpgb200_entropy_form(jpgb200_ctx *ctx);
/* Transform/entropy fusion: 1 = K1E, one persistent kernel from pixels to chunk streams
 * (csrc/k1_fused.cuh; the chain dct -> quantise -> zig_zag -> dpcm -> huffman_encode of
 * src/jpg_encode.c:36-50 without the coefficients ever reaching HBM: 3.3 B/pixel of DRAM traffic
 * instead of 15), followed by KF/KS; 0 (default: measured faster) = K1 and KE as separate
 * kernels with the int16 coefficients in HBM between them.  Identical bytes.  K1E needs width % 16 == 0, 4:4:4 or 4:2:2 and the tables.h Huffman tables; other
 * cases (4:2:0, per-image optimal tables, the parity taps, the MCU-row bands) take the separate kernels. */
int jpgb200_set_fused_transform(jpgb200_ctx *ctx, int on);
/* jpgb200_encode_batch_device on large frames: 1 (default) = three streams -- K1 of frame f+1 (high
 * priority), KE of frame f and the tails of frame f-1 are in flight together, two coefficient buffers
 * alternate; 0 = K1 and KE one after the other with only the tails overlapped.  Identical bytes. */
int jpgb200_set_pipelined_batch(jpgb200_ctx *ctx, int on);
/* jpgb200_encode_batch_device: 1 (default) = groups of frames share kernel launches (K1's tile index and the entropy
 * kernels' blockIdx.y run over the frames of a group: four launches per group).  Frames of up to 4096x2304 pixels are
 * grouped to about one 8K frame's worth of pixels per launch; larger frames by n_frames / 3, at most eight (a launch
 * pays ~10 us for filling and draining the GPU whatever its size: 260 k against 238 k Mpixel/s on 8K frames).  Three
 * or more groups run through the three-stream pipeline.  0 = four launches per frame.  Identical bytes. */
int jpgb200_set_frame_groups(jpgb200_ctx *ctx, int on);
/* 1 = K1E also stores its coefficients (jpgb200_last_coefficients reads them): the parity tap of
 * the fused kernel.  Default 0. */
int jpgb200_set_fused_tap(jpgb200_ctx *ctx, int on);
/* K1 selection (identical results; the alternatives are kept for A/B runs and as cross-checks):
 *   0 (default) producer/consumer form: convert, transform and store warps hand tiles over through
 *               mbarriers, TMA tensor stores (csrc/k1_pc.cuh); needs width % 16 == 0
 *   1           one CTA per tile, direct loads (the only form for width % 16 == 8)
 *   2           persistent CTAs, TMA bulk loads, phases separated by CTA barriers
 *   3           software-pipelined CTA (csrc/k1_pipe.cuh)
 *   4           autonomous warps, per-warp cp.async ring (csrc/k1_warp.cuh)
 *   6           form 0 fed by 2-D TMA tile loads (cp.async.bulk.tensor.2d, one per plane and tile, issued by the
 *               store warp: k1_transform_pct in csrc/k1_pc.cuh); measured 9 % slower than form 0; needs width >= 256
 *   5           form 0 named explicitly
 *   7           form 0 with FOUR transform warps per CTA taking the tile's components in turn and issuing their own
 *               TMA stores (k1_transform_pc4, 4:4:4 only; measured 3 % slower than form 0)
 *   8           twelve warps per CTA: eight convert warps with one pixel row each, four transform warps, registers
 *               re-divided with setmaxnreg (k1_transform_pc12, 4:4:4 only; measured 10 % slower than form 0)
 * Forms 2..4, 7 and 8 exist in A/B builds only (jpgb200_has_ab_forms). */
int jpgb200_set_plain_transform(jpgb200_ctx *ctx, int on);
/* Arithmetic of the transform: 0 (default) = bug-compatible with the reference (its Cb sign slip,
 * src/preprocess.c:161, and the last block column read one pixel row higher, :199-211); 1 =
 * standards-conforming (textbook Cb, no shift) -- the "robust front door" of SURVEY.md 8(f) rank 4:
 * its streams decode to the source colours in any JPEG decoder. */
int jpgb200_set_standard_mode(jpgb200_ctx *ctx, int on);
/* Huffman tables: 0 (default) = the standard Annex-K tables the reference carries in
 * src/headers/tables.h:37-72; 1 = per-image optimal tables built from the image's own symbol
 * histograms by T.81 Annex K.2 -- what construct_huffman_table (src/huffman.c:76-180) sets out to
 * do and never finishes.  Costs one host synchronisation per image
 * (histograms to the host, tables and DHT segments back); streams are 10-15 % smaller and decode
 * to the same pixels. */
int jpgb200_set_optimized_huffman(jpgb200_ctx *ctx, int on);
/* BITS (nr[t][1..16]) and HUFFVAL of the tables used by the last encode; t = 0 DC lum, 1 DC chr,
 * 2 AC lum, 3 AC chr */
int jpgb200_huffman_specs(const jpgb200_ctx *ctx, uint8_t nr[4][17], uint8_t vals[4][256], int nvals[4]);
/* The 8 bytes the reference reads in front of each plane buffer for block W/8-1
 * (src/preprocess.c:159,206-210): oob[0..7]=r, [8..15]=g, [16..23]=b.
 * Default: zeros. */
int jpgb200_set_oob_bytes(jpgb200_ctx *ctx, const uint8_t oob[24]);

/* ---- the encode path --------------------------------------------------------------- */
/* Whole pipeline of encode_bmp_to_jpeg after the BMP load (src/jpg_encode.c:32-50):
 * host planes in, complete JFIF byte stream out (host).  H2D + kernels + D2H. */
int jpgb200_encode_planar(jpgb200_ctx *ctx, const uint8_t *r, const uint8_t *g, const uint8_t *b, int width,
                          int height, int quality, int sample_ratio, uint8_t *out, size_t cap, size_t *len);
/* Any width and height (1..65535): the reference's padding path is broken
 * (src/preprocess.c:83-98,131-152: determine_resolutions adds w%8 instead of 8-w%8); here the
 * planes are extended to whole MCUs by edge replication on the device and SOF0 carries the true
 * size.  Works in either arithmetic mode. */
int jpgb200_encode_planar_any(jpgb200_ctx *ctx, const uint8_t *r, const uint8_t *g, const uint8_t *b, int width,
                              int height, int quality, int sample_ratio, uint8_t *out, size_t cap, size_t *len);
/* Same from interleaved RGB24: the body encode_rgb_to_jpeg (src/headers/jpg_encode.h:99,
 * declared but never defined, and lacking width/height) would need. */
int jpgb200_encode_rgb(jpgb200_ctx *ctx, const uint8_t *rgb, int width, int height, int quality, int sample_ratio,
                       uint8_t *out, size_t cap, size_t *len);
/* Device-resident: planes and output already on the GPU.  d_len (device, 8 bytes) receives
 * the stream length; nothing is synchronised.  d_planes = r plane, then g, then b. */
int jpgb200_encode_device(jpgb200_ctx *ctx, const uint8_t *d_r, const uint8_t *d_g, const uint8_t *d_b, int width,
                          int height, int quality, int sample_ratio, uint8_t *d_out, size_t cap,
                          unsigned long long *d_len);
/* n frames back to back: frame f reads planes at d_planes + f*3*w*h (r,g,b contiguous),
 * writes its stream at d_out + f*out_stride and its length to d_lens[f].  Inside the call the
 * tail kernels (0xFF count, stuffed scatter) of frame f run on a second stream beside the chunk
 * encode of frame f+1 (two sets of entropy workspaces alternate); the context's stream has
 * waited for all of it when the call's work is done, so the caller sees plain stream order.
 * jpgb200_set_batch_overlap(ctx, 0), or JPGB200_OVERLAP=0 in the environment, turns the overlap
 * off (per-kernel event timings are only meaningful without it). */
int jpgb200_set_batch_overlap(jpgb200_ctx *ctx, int on);
int jpgb200_encode_batch_device(jpgb200_ctx *ctx, const uint8_t *d_planes, int n_frames, int width, int height,
                                int quality, int sample_ratio, uint8_t *d_out, size_t out_stride,
                                unsigned long long *d_lens);
/* Host-buffer batch: frames[f] -> host planes (r,g,b contiguous, 3*w*h bytes, ideally
 * pinned); outs[f] receives the stream, lens[f] its length.  Copies are double-buffered
 * against the kernels on two streams. */
int jpgb200_encode_batch_host(jpgb200_ctx *ctx, const uint8_t *const *frames, int n_frames, int width, int height,
                              int quality, int sample_ratio, uint8_t *const *outs, size_t cap, size_t *lens);

/* ---- MCU-row partition of one large image across GPUs ---------------------
 * The reference has no counterpart (it cannot even load such an image: int fileSize,
 * src/bitmap.c:27).  Rank k holds a band of `band_height` rows (multiple of 8) on its GPU and
 * calls, in order, with four small exchanges in between (INTEGRATION.md shows them as NCCL
 * allgathers; jpeg-encoder-and-decoder_b200/mgpu.py is the reference orchestration):
 *   band_transform: K1.  d_r/g/b point at the band's first row; when has_halo != 0 the image
 *                row above it must be valid memory in front of the band (the 1-row halo of the
 *                reference's last-block-column quirk).  -> last_dc[3]          [exchange 0]
 *   band_bits    : K3+K4 with pred_dc = previous band's last_dc (zeros for band 0): the DC
 *                predictor changes the length of the band's first blocks.  -> band_bits [exchange 1]
 *   band_pack    : K5 at bit_phase = (sum of earlier bands' bits) & 7; is_last for the final
 *                band.  -> owned_bytes (stream bytes whose last bit lies in this band) and
 *                tail_byte (its trailing partial byte, which the NEXT band owns)  [exchange 2]
 *   band_finish  : K6.  carry_byte = previous band's tail_byte; stuffs the owned bytes into
 *                d_segment.  -> segment_len                                    [exchange 3]
 * Segments are byte-aligned: header, then the segments in band order, then EOI. */
int jpgb200_band_transform(jpgb200_ctx *ctx, const uint8_t *d_r, const uint8_t *d_g, const uint8_t *d_b, int width,
                           int band_height, int has_halo, int quality, int sample_ratio, int last_dc[3]);
int jpgb200_band_bits(jpgb200_ctx *ctx, const int pred_dc[3], unsigned long long *band_bits);
int jpgb200_band_pack(jpgb200_ctx *ctx, int bit_phase, int is_last, unsigned long long *owned_bytes,
                      uint8_t *tail_byte);
int jpgb200_band_finish(jpgb200_ctx *ctx, int carry_byte, uint8_t *d_segment, size_t cap,
                        unsigned long long *segment_len);
/* JFIF header bytes for (w,h,q,ratio): SOI..SOS, 623 bytes. */
int jpgb200_jfif_header(int width, int height, int quality, int sample_ratio, uint8_t *out, size_t cap, size_t *len);

/* ---- one very large image across the GPUs of a box (BASELINE.json configs[4]) ---------------
 * The reference cannot hold such an image at all (int fileSize, src/bitmap.c:27); what is replaced is
 * encode_bmp_to_jpeg's single pass (src/jpg_encode.c:19-52) by one band of whole MCU rows per GPU.  One
 * process (or host thread) per GPU, each with its own context; `nccl_comm` is an ncclComm_t over the
 * participating ranks (the caller's, or one made with the three helpers below; NCCL is resolved at run
 * time with dlopen, this library does not link against it).  Rank 0 owns the output buffer, the other
 * ranks map it through CUDA IPC and their KS kernels store their segments straight into it over NVLink.
 * Per encode: two ncclAllGathers of 32 / 24 bytes per rank on the context's stream (DC predictors + bit
 * counts, then stuffed lengths + boundary bytes; csrc/mgpu.cuh), one 8-byte allgather as the completion
 * barrier, no host synchronisation until jpgb200_mgpu_result. */
typedef struct jpgb200_mgpu jpgb200_mgpu;
int jpgb200_nccl_unique_id(uint8_t id[128]);                                                   /* ncclGetUniqueId */
int jpgb200_nccl_init(void **comm, int world, int rank, const uint8_t id[128], int device);   /* ncclCommInitRank */
int jpgb200_nccl_destroy(void *comm);
int jpgb200_mgpu_create(jpgb200_mgpu **mg, jpgb200_ctx *ctx, void *nccl_comm, int rank, int world, size_t out_cap);
int jpgb200_mgpu_destroy(jpgb200_mgpu *mg);
/* d_r/d_g/d_b: the band's first row on this GPU; the row above it must be readable when rank > 0 (the
 * reference reads the last block column one pixel row higher).  Enqueues everything. */
int jpgb200_mgpu_encode(jpgb200_mgpu *mg, const uint8_t *d_r, const uint8_t *d_g, const uint8_t *d_b, int width, int band_height,
                        int image_height, int quality, int sample_ratio);
/* waits for this rank's stream; on rank 0 *total_len = bytes of the stitched JFIF stream */
int jpgb200_mgpu_result(jpgb200_mgpu *mg, unsigned long long *total_len);
uint8_t *jpgb200_mgpu_output(jpgb200_mgpu *mg); /* rank 0: device pointer of the stitched stream; NULL elsewhere */

/* ---- parity taps (test and staged-mode surface) ------------------------------------- */
/* Quantised zig-zag coefficients after zig_zag() and before dpcm() (src/jpg_encode.c:44),
 * int16 [Y blocks | Cb blocks | Cr blocks][64]; blocks per component = (w/8)*(h/8)
 * (chroma halved for 4:2:2). */
int jpgb200_coefficients(jpgb200_ctx *ctx, const uint8_t *r, const uint8_t *g, const uint8_t *b, int width,
                         int height, int quality, int sample_ratio, int16_t *zz_out);
/* The coefficients the last encode left in HBM (same layout): every encode through the separate
 * kernels, or through K1E with jpgb200_set_fused_tap(ctx, 1). */
int jpgb200_last_coefficients(jpgb200_ctx *ctx, int16_t *zz_out);
/* dpcm() on the coefficients of the last jpgb200_coefficients/encode call
 * (src/dpcm.c:6-21): d_out = int32 [Y | Cb | Cr] DC sequences in the requested mode. */
int jpgb200_dc_sequence(jpgb200_ctx *ctx, int mode, int32_t *d_out);
/* The histogram loops of huffman_encode (src/huffman.c:29-42) on the same coefficients:
 * freq[4][257] = lum_DC, lum_AC, chrom_DC, chrom_AC, freq[*][256] = 1 (src/huffman.c:54). */
int jpgb200_histograms(jpgb200_ctx *ctx, int dc_mode, int sym_mode, int32_t *freq);
/* Global bit offset of every block in scan order (n_blocks+1 values, last = total bits),
 * from the last encode/coefficients call. */
int jpgb200_block_bit_offsets(jpgb200_ctx *ctx, unsigned long long *offsets);
long long jpgb200_num_blocks(int width, int height, int sample_ratio, long long n[3]);

/* Block-level stage API used directly by the reference's driver (src/jpg_driver.c:135-144):
 * n FP64 blocks of 64 values, values[y*8+x] (src/block.c:25-33), processed in place. */
int jpgb200_stage_dct(jpgb200_ctx *ctx, double *blocks, long long n);                        /* dct_block */
int jpgb200_stage_quantise(jpgb200_ctx *ctx, double *blocks, long long n, const int table[64]); /* quantise_lum|chr */
int jpgb200_stage_zigzag(jpgb200_ctx *ctx, const double *blocks, long long n, int *zz);       /* zig_zag_block */
/* The remaining per-stage functions of the reference, GPU-backed, so that a caller of its
 * stage-level API (void f(JpgData), src/jpg_encode.c:32-50) gets reference-shaped results:
 *   preprocess_jpeg's block fill + level shift (src/preprocess.c:155-169,176-196), FP64, same
 *   operation order, including the last-block-column shift; blocks are [N][64] doubles */
int jpgb200_stage_preprocess(jpgb200_ctx *ctx, const uint8_t *r, const uint8_t *g, const uint8_t *b, int width,
                             int height, double *Y, double *Cb, double *Cr);
/*   subsample_422 (a stub in src/downsample.c:24; semantics SURVEY.md A.3): in = full-resolution
 *   chroma blocks [bh][bw][64], out = [bh][bw/2][64] */
int jpgb200_stage_subsample422(jpgb200_ctx *ctx, const double *in, double *out, int width_blocks, int height_blocks);
/*   dpcm() (src/dpcm.c:6-21) on one component's DC sequence, in place: d[i] = dc[i] - d[i-1] */
int jpgb200_stage_dpcm(jpgb200_ctx *ctx, int *dc, long long n);
/*   calculate_freq_block_DC/AC (src/huffman.c:182-222) over n int[64] blocks, accumulated into
 *   freq_dc[257] / freq_ac[257]; sym_mode JPGB200_SYM_REF reproduces run|size */
int jpgb200_stage_histograms(jpgb200_ctx *ctx, const int *zz, long long n, int sym_mode, int *freq_dc, int *freq_ac);
/* Entropy-code caller-supplied quantised zig-zag coefficients (layout of jpgb200_coefficients)
 * into a complete JFIF stream: what huffman_encode (src/huffman.c:23) was meant to end in. */
int jpgb200_encode_coefficients(jpgb200_ctx *ctx, const int16_t *zz, int width, int height, int quality,
                                int sample_ratio, uint8_t *out, size_t cap, size_t *len);
/* scaled tables as the reference's first scale_table call leaves them (src/quantise.c:74-86) */
int jpgb200_quant_tables(int quality, int lum[64], int chr[64]);

/* ---- device helpers for hosts without a CUDA runtime of their own ---------------------- */
int jpgb200_device_count(void);
int jpgb200_malloc(jpgb200_ctx *ctx, void **dptr, size_t bytes);
int jpgb200_free(jpgb200_ctx *ctx, void *dptr);
int jpgb200_host_alloc(void **hptr, size_t bytes); /* pinned */
int jpgb200_host_alloc_ex(void **hptr, size_t bytes, int flags); /* flags: 1 = write-combined (host-write-only buffers) */
/* NUMA node of the device's PCIe function (-1 when the platform does not report one, e.g. a VM) and the
 * number of NUMA nodes visible to this process: placement hints for the caller's pinned buffers. */
int jpgb200_device_numa(int device, int *gpu_node, int *visible_nodes);
/* The host<->device copies of jpgb200_encode_batch_host without its kernels (one cudaMemcpyAsync per frame
 * and direction, both directions at once): the PCIe ceiling of an end-to-end encode on this host. */
int jpgb200_copy_probe(jpgb200_ctx *ctx, const uint8_t *const *frames, int n_frames, size_t in_bytes, uint8_t *const *outs,
                       size_t out_bytes);
int jpgb200_host_free(void *hptr);
int jpgb200_memcpy_h2d(jpgb200_ctx *ctx, void *dptr, const void *hptr, size_t bytes);
int jpgb200_memcpy_d2h(jpgb200_ctx *ctx, void *hptr, const void *dptr, size_t bytes);
/* deterministic synthetic planes [3][h][w] on the device (kind 0 natural, 1 noise) */
int jpgb200_synth_device(jpgb200_ctx *ctx, uint8_t *d_planes, int width, int height, uint64_t seed, uint32_t frame,
                         int kind);
/* rows row0 .. row0 + rows - 1 of the same image into planes [3][rows][w] (bands of the MCU-row partition) */
int jpgb200_synth_device_rows(jpgb200_ctx *ctx, uint8_t *d_planes, int width, int rows, int row0, uint64_t seed,
                              uint32_t frame, int kind);

/* ---- measurement --------------------------------------------------------------------- */
/* With profiling on (1), every kernel launch of the encode path is bracketed by CUDA events on
 * the launching stream; with 2 only K1's launches are (two event records per frame instead of
 * eight: for a timed region that wants the roofline kernel's live duration and little else), with
 * 3 only every eighth K1 launch and, in other frames, every eighth KE launch
 * (jpgb200_kernel_times then reports the sampled launches: their count and their summed duration); sums are read back with jpgb200_kernel_times (ms and launch counts
 * per kernel, JPGB200_NUM_KERNELS entries) and cleared by jpgb200_reset_profile. */
int jpgb200_set_profiling(jpgb200_ctx *ctx, int on);
int jpgb200_kernel_times(jpgb200_ctx *ctx, double *ms, long long *launches);
int jpgb200_reset_profile(jpgb200_ctx *ctx);
/* kernels launched by this context since creation (bench.py's gpu_launches) */
long long jpgb200_launch_count(const jpgb200_ctx *ctx);
const char *jpgb200_kernel_name(int k);

#ifdef __cplusplus
}
#endif
#endif
