/*
 * vcs_b200.h -- C ABI of the B200-native VCS-h264 interframe hot path.
 *
 * The reference (miatang13/VCS-h264) is pure Python and has no FFI; the interface each
 * entry point replaces is therefore a Python method.  Citations are file:line relative to
 * the reference checkout.  A maintainer binds these with ctypes; the
 * in-tree binding is vcs_h264_b200/_capi.py.
 *
 * Conventions
 *   - every function returns 0 (VCS_OK) or a negative VCS_E_* code; vcs_last_error(ctx)
 *     gives the text of the last failure on that context;
 *   - one context per host thread and device; contexts are not thread-safe;
 *   - frames are BGR-interleaved uint8, C-contiguous, H x W x 3 (what cv2 hands the
 *     reference); a "clip" is T such frames back to back;
 *   - *_dev functions take DEVICE pointers, enqueue on the context's stream and do not
 *     synchronise; *_host functions take HOST pointers, copy in, run, copy out and return
 *     after the results are in the caller's buffers;
 *   - device image buffers must be 4-byte aligned and device coefficient buffers 16-byte aligned (anything
 *     from cudaMalloc or a torch tensor is); misaligned pointers are refused with VCS_E_INVALID;
 *   - there is no CPU fallback: without a CUDA device vcs_create fails.
 */
#ifndef VCS_B200_H
#define VCS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VCS_OK 0
#define VCS_E_INVALID (-1)     /* bad argument */
#define VCS_E_CUDA (-2)        /* CUDA runtime/driver error, see vcs_last_error */
#define VCS_E_NOMEM (-3)
#define VCS_E_UNSUPPORTED (-4) /* valid request this build has no kernel for */

/* matching cost */
#define VCS_METRIC_WRAP8 0 /* sum((ref - cur) mod 256): what motion.py:146 computes on uint8 */
#define VCS_METRIC_SAD 1   /* sum(|ref - cur|): generalised mode, no literal reference oracle */

/* per-macroblock flag bits written by the search */
#define VCS_MB_STATIC 1 /* static test passed (motion.py:113), mv = (0,0), no search */
#define VCS_MB_NOCAND 2 /* candidate set empty, mv = (-x,-y) (motion.py:102,156-161) */

/* coefficient output formats of the DCT/quant stage */
#define VCS_COEF_F64 0      /* float64 planes, D/Q un-rounded: DCTcompressor.py:71 */
#define VCS_COEF_F64_RINT 1 /* float64 planes, np.round(D/Q): DCTCompression/dct.py:179 */
#define VCS_COEF_I16_RINT 2 /* the same integers as int16 planes (compact wire format) */
#define VCS_COEF_I8_RINT 3  /* the same integers as int8 planes: |index| <= 1024/min(Q), so this is lossless
                               exactly when min(Q) >= 9 (e.g. QF <= 50); refused with VCS_E_INVALID otherwise */

/* which ME kernel a call may use */
#define VCS_ME_AUTO 0    /* tiled full-search kernel when step==1 and bs is 8 or 16 */
#define VCS_ME_GENERIC 1 /* one CTA per macroblock, any bs/step */
#define VCS_ME_TILED 2   /* fail with VCS_E_UNSUPPORTED rather than fall back */

typedef struct vcs_ctx vcs_ctx;

/*
 * Candidate set and cost of MotionProcessor._find_match (InterframeCompression/motion.py:100-154)
 * in interval form:
 *     rows i = max(y+lo,0), +step, ... while i <= min(y+hi, H-bs-slack)   (cols likewise with W)
 * scan order rows outer / cols inner ascending, first strict minimum wins (motion.py:133-152).
 * The reference's own loop is lo=-R, hi=R-bs-1, slack=1, R=2*bs (motion.py:18), step=round(bs/3)
 * (motion.py:132); vcs_me_reference_params() fills that in.  BASELINE.json's symmetric +/-R
 * full search is lo=-R, hi=R, slack=0, step=1.
 */
typedef struct vcs_me_params {
    int32_t H, W;       /* frame size; partial macroblocks are dropped (motion.py:82-87) */
    int32_t bs;         /* macroblock size */
    int32_t lo, hi;     /* inclusive offset interval, both axes */
    int32_t step;       /* candidate stride, anchored at the clamped window start */
    int32_t slack;      /* upper clamp is dim - bs - slack */
    int32_t metric;     /* VCS_METRIC_* */
    int64_t static_thr; /* SIMILARITY_THRESHOLD of motion.py:8,113; < 0 disables the test */
    int32_t kernel;     /* VCS_ME_* */
    int32_t reserved;
} vcs_me_params;

/* ---- library / context ------------------------------------------------------------------ */
int vcs_version(void);
int vcs_create(int device, vcs_ctx **out);
int vcs_destroy(vcs_ctx *ctx);
const char *vcs_last_error(const vcs_ctx *ctx);
/* kernels are enqueued on this cudaStream_t (0 = CUDA's legacy default stream); a new context
 * starts on a non-blocking stream of its own, vcs_use_own_stream goes back to it */
int vcs_set_stream(vcs_ctx *ctx, void *cuda_stream);
int vcs_use_own_stream(vcs_ctx *ctx);
int vcs_synchronize(vcs_ctx *ctx);
int vcs_device_info(vcs_ctx *ctx, int *sm_count, int *cc_major, int *cc_minor,
                    size_t *smem_per_block_optin);
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
int64_t vcs_launch_count(const vcs_ctx *ctx);

/* ---- host-side constants (pure functions, no device) ---------------------------------- */
/* MotionProcessor.__init__ + _find_match defaults: motion.py:15-18,123-132 */
int vcs_me_reference_params(int H, int W, int bs, vcs_me_params *out);
/* symmetric +/-R step-1 full search (BASELINE.json configs 2/3/5) */
int vcs_me_fullsearch_params(int H, int W, int bs, int R, int metric, int64_t static_thr,
                             vcs_me_params *out);
/* number of macroblocks of _split_frame_into_mblocks (motion.py:74-98) */
int vcs_num_blocks(int H, int W, int bs);
/* Q = [QY', QC', QC'] of DCTcompressor.py:11-38 / dct.py:139-166 for quality factor qf;
 * Q is double[3][64] (Y, Cr, Cb). VCS_E_INVALID for qf >= 100. */
int vcs_q_tables(double qf, double *Q);
/* DCTCompressor._dctMatrix (DCTcompressor.py:124-133), double[8][8] */
int vcs_dct_matrix(double *C);
/* DCTCompressor.Q of this context (default: qf = 50, DCTcompressor.py:29) */
int vcs_set_q(vcs_ctx *ctx, const double *Q);

/* ---- motion estimation: MotionProcessor.process_motion_prediction (motion.py:20-36) ---- */
/* One (cur, ref) pair.  mv: int16[N][2] = [dx,dy]; cost: uint32[N] best search cost (static
 * blocks: the one-sided static sum; no candidate: 0xFFFFFFFF); flags: uint8[N] VCS_MB_*.
 * cost / flags may be NULL. */
int vcs_me_search_dev(vcs_ctx *ctx, const vcs_me_params *p, const uint8_t *cur,
                      const uint8_t *ref, int16_t *mv, uint32_t *cost, uint8_t *flags);
int vcs_me_search_host(vcs_ctx *ctx, const vcs_me_params *p, const uint8_t *cur,
                       const uint8_t *ref, int16_t *mv, uint32_t *cost, uint8_t *flags);
/* Whole clip under the reference's GOP rule (encoder.py:25,51-52): frame t is an I-frame iff
 * t % gop_len == 0, every other frame is a P-frame predicted from ORIGINAL frame
 * (t / gop_len) * gop_len.  Outputs are indexed by P-frame ordinal p (the p-th non-I frame):
 * mv int16[nP][N][2] etc.; vcs_num_p_frames gives nP. */
int vcs_num_p_frames(int T, int gop_len);
int vcs_me_search_clip_dev(vcs_ctx *ctx, const vcs_me_params *p, const uint8_t *frames, int T,
                           int gop_len, int16_t *mv, uint32_t *cost, uint8_t *flags);

/* ---- motion compensation / residual ------------------------------------------------------ */
/* MotionProcessor.reconstruct_from_motion_vectors (motion.py:42-69): pred = 0 outside MBs */
int vcs_mc_dev(vcs_ctx *ctx, int H, int W, int bs, const uint8_t *ref, const int16_t *mv,
               uint8_t *pred);
int vcs_mc_host(vcs_ctx *ctx, int H, int W, int bs, const uint8_t *ref, const int16_t *mv,
                uint8_t *pred);
/* MotionProcessor.get_residuals (motion.py:38-40): out = a - b mod 256;
 * Decoder._fully_reconstruct (decoder.py:57): out = a + b mod 256 */
int vcs_sub_wrap_dev(vcs_ctx *ctx, const uint8_t *a, const uint8_t *b, size_t n, uint8_t *out);
int vcs_add_wrap_dev(vcs_ctx *ctx, const uint8_t *a, const uint8_t *b, size_t n, uint8_t *out);
int vcs_sub_wrap_host(vcs_ctx *ctx, const uint8_t *a, const uint8_t *b, size_t n, uint8_t *out);
int vcs_add_wrap_host(vcs_ctx *ctx, const uint8_t *a, const uint8_t *b, size_t n, uint8_t *out);

/* ---- 8x8 DCT / quantise / dequantise / IDCT ---------------------------------------------- */
/*
 * DCTCompressor._dct2 / _idct2 (DCTcompressor.py:111-121) on n bare 8x8 float64 blocks, row-major, host
 * buffers: out = C.X.C^T (inverse = 0) or C^T.X.C (inverse != 0), every element a sequential-k FMA chain
 * like np.matmul.  Backs the private helpers of the drop-in class; the hot path is vcs_compress_* below.
 */
int vcs_dct2_blocks_host(vcs_ctx *ctx, int nblocks, int inverse, const double *in, double *out);

/* DCTCompressor.compress (DCTcompressor.py:49-74; rounded: dct.py:169-186).  H, W must be
 * multiples of 8 (the reference would bilinear-resize; VCS_E_INVALID here).  coef: 3 planes
 * H x W of float64 (VCS_COEF_F64*), int16 (VCS_COEF_I16_RINT) or int8 (VCS_COEF_I8_RINT), order Y, Cr, Cb. */
int vcs_compress_dev(vcs_ctx *ctx, int H, int W, const uint8_t *bgr, int coef_mode, void *coef);
int vcs_compress_host(vcs_ctx *ctx, int H, int W, const uint8_t *bgr, int coef_mode, void *coef);
/* DCTCompressor.decompress (DCTcompressor.py:76-93): planes -> BGR uint8.  If pred != NULL the
 * result is pred + decoded mod 256 (Decoder._fully_reconstruct, decoder.py:52-60). */
int vcs_decompress_dev(vcs_ctx *ctx, int H, int W, int coef_mode, const void *coef,
                       const uint8_t *pred, uint8_t *bgr);
int vcs_decompress_host(vcs_ctx *ctx, int H, int W, int coef_mode, const void *coef,
                        const uint8_t *pred, uint8_t *bgr);

/* ---- fused P-frame path: Encoder._process_P_frame (encoder.py:49-70) + the decoder-side
 *      reconstruction loop (decoder.py:52-69) ---------------------------------------------- */
/* Given MVs: MC gather, residual, BGR->YCrCb, DCT, /Q (+round), and optionally xQ, IDCT,
 * truncating store, YCrCb->BGR, + pred.  coef and recon may each be NULL. Clip form: outputs
 * indexed by P-frame ordinal (coef planes [nP][3][H][W], recon [nP][H][W][3]). */
int vcs_residual_dct_clip_dev(vcs_ctx *ctx, int H, int W, int bs, const uint8_t *frames, int T,
                              int gop_len, const int16_t *mv, int coef_mode, void *coef,
                              uint8_t *recon);
/* ME + the above for a whole clip, device resident (bench.py `value`). */
int vcs_encode_clip_dev(vcs_ctx *ctx, const vcs_me_params *p, const uint8_t *frames, int T,
                        int gop_len, int coef_mode, int16_t *mv, uint32_t *cost, uint8_t *flags,
                        void *coef, uint8_t *recon);
/* The same from HOST buffers: frames are copied in, mv / cost / flags / coef / recon copied
 * out (any output may be NULL = not wanted, stays on the device) -- bench.py `e2e`.
 * Copies and kernels are pipelined GOP-chunk by GOP-chunk over two streams; device and pinned
 * staging memory is grown on first use and kept. */
int vcs_encode_clip_host(vcs_ctx *ctx, const vcs_me_params *p, const uint8_t *frames, int T,
                         int gop_len, int coef_mode, int16_t *mv, uint32_t *cost, uint8_t *flags,
                         void *coef, uint8_t *recon);

/*
 * Packed form of the int8 indices (the wire form of Frame.r, frame.py:1-8; the reference keeps dense float64 planes
 * and never wrote the zero-run stage its proposal promised).  Exact (lossless) whenever int8 indices are (QF <= 50):
 *   bitmap    uint64 [nP][3][H/8][W/8]  occupancy of every 8x8 block, bit 8*i+j = row i, column j
 *   nibbles   uint8 stream: one 4-bit code per non-zero index of a block in bit order, low nibble first, every block
 *             padded to a whole byte; code = v & 15 for v in [-8, 7], code 0 = escape
 *   escapes   int8 stream: the value of every escaped index, in the same order
 *   row_count uint32 [nP][3][H/8][2]    per block row: bytes of its nibble stream, number of its escapes (a prefix
 *             sum locates any block row in both streams)
 * Blocks in (P-frame, channel Y/Cr/Cb, block row, block column) order; both streams run through the whole clip.
 *
 * vcs_encode_clip_host_packed = vcs_encode_clip_host with VCS_COEF_I8_RINT whose coefficients come back packed; the
 * dense planes never cross the bus.  nibbles_capacity = 3*H*W*nP/2 + 3*(H/8)*(W/8)*nP and escapes_capacity =
 * 3*H*W*nP are always enough; smaller buffers are refused with VCS_E_INVALID when they overflow.  lengths[2] (out) =
 * bytes of the nibble stream, number of escapes.
 */
int vcs_encode_clip_host_packed(vcs_ctx *ctx, const vcs_me_params *p, const uint8_t *frames, int T, int gop_len,
                                int16_t *mv, uint32_t *cost, uint8_t *flags, uint64_t *bitmap, uint32_t *row_count,
                                uint8_t *nibbles, size_t nibbles_capacity, int8_t *escapes, size_t escapes_capacity,
                                uint64_t *lengths, uint8_t *recon);
/* dense int8 planes [nP][3][H][W] -> packed (all DEVICE pointers, worst-case capacities as above; lengths_host[2] is
 * a host pointer; synchronises) */
int vcs_pack_coef_dev(vcs_ctx *ctx, int H, int W, int nP, const int8_t *coef, uint64_t *bitmap, uint32_t *row_count,
                      uint8_t *nibbles, int8_t *escapes, uint64_t *lengths_host);
/* the exact inverse (all DEVICE pointers, enqueued on the context's stream); streams shorter than their bitmaps claim
 * are never read past their end: the affected blocks decode to zero and the context reports VCS_E_INVALID */
int vcs_unpack_coef_dev(vcs_ctx *ctx, int H, int W, int nP, const uint64_t *bitmap, const uint32_t *row_count,
                        const uint8_t *nibbles, uint64_t nnibble_bytes, const int8_t *escapes, uint64_t nescapes,
                        int8_t *coef);
/* vcs_decode_clip_host from the packed form (host buffers) */
int vcs_decode_clip_host_packed(vcs_ctx *ctx, int H, int W, int bs, const uint8_t *ref_frames, int T, int gop_len,
                                const int16_t *mv, const uint64_t *bitmap, const uint32_t *row_count,
                                const uint8_t *nibbles, uint64_t nnibble_bytes, const int8_t *escapes, uint64_t nescapes,
                                uint8_t *recon);

/* ---- decoder side: Decoder._reconstruct_P_frame over a clip (decoder.py:52-69) --------------------- */
/* ref_frames: the ORIGINAL I-frames only, uint8 [nG][H][W][3] (Decoder.ref_frames); mv / coef indexed by
 * P-frame ordinal as the encoder wrote them; recon: uint8 [nP][H][W][3] = MC(ref, mv) + decompress(coef). */
int vcs_decode_clip_dev(vcs_ctx *ctx, int H, int W, int bs, const uint8_t *ref_frames, int T,
                        int gop_len, const int16_t *mv, int coef_mode, const void *coef,
                        uint8_t *recon);
int vcs_decode_clip_host(vcs_ctx *ctx, int H, int W, int bs, const uint8_t *ref_frames, int T,
                         int gop_len, const int16_t *mv, int coef_mode, const void *coef,
                         uint8_t *recon);
/* number of non-zero coefficients among n elements of a DEVICE coefficient buffer: the numerator of the
 * sparsity print of DCTCompression/dct.py:188-191 (sparsity = 1 - count / n) */
int vcs_count_nonzero_dev(vcs_ctx *ctx, int coef_mode, const void *coef, size_t n,
                          unsigned long long *count_host);

/* ---- fp32 tier of the DCT stage (BASELINE.json north_star: "fp32 DCT reconstruction ... any coefficient that flips at
 * a quantisation rounding boundary is counted and reported"; SURVEY appendix A, tier T2) ---------------------------- */
/* bits = 64 (default): the transforms of DCTcompressor.py:111-121 in float64 with the reference's operation order,
 * bit-identical.  bits = 32: the same kernel in float (FFMA), no exact-quotient fallback: results are close, not equal,
 * and are judged with vcs_flip_counters_dev.  Applies to every later DCT-stage launch of this context (compress,
 * decompress, encode/decode clip); the packed sink and the float64 coefficient planes stay with the exact tier. */
int vcs_set_dct_precision(vcs_ctx *ctx, int bits);
/* Compares two DEVICE results of the same shape: ncoef rounded indices (coef_mode VCS_COEF_I8_RINT or _I16_RINT) and
 * npx uint8 pixels (either count may be 0).  out3_host[0] = indices that differ (rounding-boundary flips),
 * [1] = pixels that differ (truncation-boundary flips, decoder.py:52-60), [2] = sum of squared pixel differences. */
int vcs_flip_counters_dev(vcs_ctx *ctx, int coef_mode, const void *coef_a, const void *coef_b, size_t ncoef,
                          const uint8_t *px_a, const uint8_t *px_b, size_t npx, unsigned long long *out3_host);

/* ---- intra mode decision: IntraframeCompression/intraframe.py + intramodes.py --------- */
/* luma4x4 (intraframe.py:24-151): 9 predictors per 4x4 block, first strict SAD minimum, neighbours from the
 * ORIGINAL plane.  Y: uint8 H x W (multiples of 4); res / pred: int32 H x W; modes: uint8 (H/4) x (W/4). */
int vcs_intra_luma4x4_dev(vcs_ctx *ctx, int H, int W, const uint8_t *Y, int32_t *res, int32_t *pred,
                          uint8_t *modes);
/* luma16x16 (intraframe.py:153-225): vertical / horizontal / dc per 16x16 block. */
int vcs_intra_luma16x16_dev(vcs_ctx *ctx, int H, int W, const uint8_t *Y, int32_t *res, int32_t *pred,
                            uint8_t *modes);
/* chroma8x8 (intraframe.py:228-317): joint mode for Cr and Cb per 8x8 block; Cb's upper neighbour is the
 * residual row above (intraframe.py:266), so residuals / predictions are int32. */
int vcs_intra_chroma8x8_dev(vcs_ctx *ctx, int H, int W, const uint8_t *Cr, const uint8_t *Cb,
                            int32_t *crres, int32_t *crpred, int32_t *cbres, int32_t *cbpred,
                            uint8_t *modes);
/* the three above from host buffers: which = 0 luma4x4, 1 luma16x16 (p0 = Y; p1, res1, pred1 unused),
 * 2 chroma8x8 (p0 = Cr, p1 = Cb; res0/pred0 = Cr outputs, res1/pred1 = Cb outputs) */
int vcs_intra_host(vcs_ctx *ctx, int which, int H, int W, const uint8_t *p0, const uint8_t *p1,
                   int32_t *res0, int32_t *pred0, int32_t *res1, int32_t *pred1, uint8_t *modes);

/* ---- 4:2:0 chroma subsampling demo: ChromaSubsampling/chroma.py ------------------------- */
/* chroma.py:9-21: BGR -> YCrCb, 2x2 box filter of Cr and Cb (cv2.boxFilter: anchor (1,1), reflect-101 border,
 * ceil(sum/4)), every second sample.  Y is [H][W]; cr and cb are [ceil(H/2)][ceil(W/2)].  Any H, W >= 1. */
int vcs_chroma420_dev(vcs_ctx *ctx, int H, int W, const uint8_t *bgr, uint8_t *Y, uint8_t *cr, uint8_t *cb);
/* chroma.py:27-41: the demo's float64 reconstruction to BGR [H][W][3] (as the script behaves under NumPy 2:
 * `Cr - 128` wraps in uint8). */
int vcs_chroma420_to_bgr_dev(vcs_ctx *ctx, int H, int W, const uint8_t *Y, const uint8_t *cr, const uint8_t *cb,
                             uint8_t *bgr);
/* both from host buffers; bgr_out may be NULL (subsample only) */
int vcs_chroma420_host(vcs_ctx *ctx, int H, int W, const uint8_t *bgr, uint8_t *Y, uint8_t *cr, uint8_t *cb,
                       uint8_t *bgr_out);
int vcs_chroma420_to_bgr_host(vcs_ctx *ctx, int H, int W, const uint8_t *Y, const uint8_t *cr, const uint8_t *cb,
                              uint8_t *bgr);

/* ---- measurement support ------------------------------------------------------------------ */
/* Register-only issue-rate microbenchmarks that define the INT32-pipe roofline on the box the
 * bench runs on.  which: 0 VABSDIFF4.U8.ACC, 1 IADD3, 2 LOP3, 3 IMAD, 4 IDP.4A,
 * 5 the wrap8 triple (IADD+LOP3+IDP.4A per word), 6 VABSDIFF4 interleaved with LDS.
 * Returns warp-instructions per second over the whole GPU in *warp_instr_per_s (for 5: words
 * per second / 32) and the measured SM clock in *sm_mhz. */
int vcs_microbench(vcs_ctx *ctx, int which, int iters, double *warp_instr_per_s, double *sm_mhz);
/* With timing enabled every *_clip_dev / *_clip_host call brackets its ME launch and its
 * residual/DCT launch with CUDA events on the launching stream.  vcs_kernel_times synchronises
 * the device, returns the summed milliseconds of both kernels over the calls made since the last
 * query (ncalls = number of bracketed launches pairs) and resets the accumulation. */
int vcs_enable_kernel_timing(vcs_ctx *ctx, int on);
int vcs_kernel_times(vcs_ctx *ctx, double *me_ms_total, double *dct_ms_total, int *ncalls);

#ifdef __cplusplus
}
#endif
#endif /* VCS_B200_H */
