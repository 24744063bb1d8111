/*
 * oracle/vcs_oracle.c -- CPU restatement of the VCS-h264 interframe hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under vcs_h264_b200/ may link, import or call this
 * file.  It is used by tests/, by __graft_entry__.smoke() and by bench.py's cpu_baseline /
 * --impl reference legs, as the checker and as the timed CPU port -- never as the product.
 *
 * The reference is pure Python (NumPy + OpenCV); it cannot be compiled, so there is no
 * oracle/_ref build.  Parity is pinned instead by golden vectors generated in the build
 * container by importing the unmodified reference from /root/reference
 * (tests/golden/make_golden.py -> tests/golden/golden.npz); tests/test_oracle_golden.py checks
 * every function below against them.  The arithmetic that lives outside the reference repo
 * (OpenCV 4.13 cvtColor fixed point, cv2.subtract saturation, NumPy/OpenBLAS 8x8 float64
 * matmul, NumPy float64->uint8 cast) is restated here and pinned by the same vectors.
 *
 * Citations are file:line relative to /root/reference/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#ifdef _OPENMP
#include <omp.h>
#endif

#define VCS_METRIC_WRAP8 0
#define VCS_METRIC_SAD 1

/* ------------------------------------------------------------------------------------ */
/* Block costs.  A block row is 3*bs contiguous bytes of a BGR-interleaved frame.        */
/* ------------------------------------------------------------------------------------ */

/* InterframeCompression/motion.py:146  np.sum(np.abs(ref_block - block)) on uint8 arrays:
 * the subtraction wraps mod 256 and np.abs is a no-op, the sum is taken in uint64.       */
static uint64_t cost_wrap8_scalar(const uint8_t *r, const uint8_t *c, int pitch, int bs) {
    uint64_t s = 0;
    for (int v = 0; v < bs; ++v)
        for (int u = 0; u < 3 * bs; ++u)
            s += (uint8_t)(r[v * pitch + u] - c[v * pitch + u]);
    return s;
}

/* Generalised metric (no literal oracle in the reference): true SAD.          */
static uint64_t cost_sad_scalar(const uint8_t *r, const uint8_t *c, int pitch, int bs) {
    uint64_t s = 0;
    for (int v = 0; v < bs; ++v)
        for (int u = 0; u < 3 * bs; ++u) {
            int d = (int)r[v * pitch + u] - (int)c[v * pitch + u];
            s += (uint64_t)(d < 0 ? -d : d);
        }
    return s;
}

/* InterframeCompression/motion.py:111  np.sum(np.abs(cv2.subtract(ref_block, block))):
 * cv2.subtract saturates at 0, so this is the one-sided sum of max(ref - cur, 0).        */
static uint64_t cost_onesided_scalar(const uint8_t *r, const uint8_t *c, int pitch, int bs) {
    uint64_t s = 0;
    for (int v = 0; v < bs; ++v)
        for (int u = 0; u < 3 * bs; ++u) {
            int d = (int)r[v * pitch + u] - (int)c[v * pitch + u];
            s += (uint64_t)(d > 0 ? d : 0);
        }
    return s;
}

#if defined(__SSE2__)
/* SSE2 versions of the three costs (same integers, used so the timed CPU port is a fair
 * baseline).  mode: 0 wrap8, 1 sad, 2 one-sided.  Checked against the scalar versions by
 * vcs_oracle_selfcheck_simd().                                                            */
static uint64_t cost_simd(const uint8_t *r, const uint8_t *c, int pitch, int bs, int mode) {
    const int n = 3 * bs;
    const __m128i zero = _mm_setzero_si128();
    __m128i acc = zero;
    uint64_t tail = 0;
    for (int v = 0; v < bs; ++v) {
        const uint8_t *rr = r + v * pitch, *cc = c + v * pitch;
        int u = 0;
        for (; u + 16 <= n; u += 16) {
            __m128i a = _mm_loadu_si128((const __m128i *)(rr + u));
            __m128i b = _mm_loadu_si128((const __m128i *)(cc + u));
            __m128i s;
            if (mode == 0) s = _mm_sad_epu8(_mm_sub_epi8(a, b), zero);
            else if (mode == 1) s = _mm_sad_epu8(a, b);
            else s = _mm_sad_epu8(_mm_subs_epu8(a, b), zero);
            acc = _mm_add_epi64(acc, s);
        }
        for (; u < n; ++u) {
            int d = (int)rr[u] - (int)cc[u];
            if (mode == 0) tail += (uint8_t)d;
            else if (mode == 1) tail += (uint64_t)(d < 0 ? -d : d);
            else tail += (uint64_t)(d > 0 ? d : 0);
        }
    }
    uint64_t lanes[2];
    _mm_storeu_si128((__m128i *)lanes, acc);
    return lanes[0] + lanes[1] + tail;
}
#endif

static inline uint64_t block_cost(const uint8_t *r, const uint8_t *c, int pitch, int bs,
                                  int mode, int use_simd) {
#if defined(__SSE2__)
    if (use_simd) return cost_simd(r, c, pitch, bs, mode);
#endif
    (void)use_simd;
    if (mode == 0) return cost_wrap8_scalar(r, c, pitch, bs);
    if (mode == 1) return cost_sad_scalar(r, c, pitch, bs);
    return cost_onesided_scalar(r, c, pitch, bs);
}

/* 0 when the SIMD and scalar costs agree on pseudo-random blocks, else the failing mode+1 */
int vcs_oracle_selfcheck_simd(void) {
#if defined(__SSE2__)
    enum { P = 3 * 40 };
    static uint8_t a[40 * P], b[40 * P];
    uint32_t s = 12345u;
    for (int rep = 0; rep < 50; ++rep) {
        for (int i = 0; i < 40 * P; ++i) {
            s = s * 1664525u + 1013904223u; a[i] = (uint8_t)(s >> 24);
            s = s * 1664525u + 1013904223u; b[i] = (uint8_t)(s >> 24);
        }
        for (int bs = 1; bs <= 32; ++bs) {
            if (cost_simd(a, b, P, bs, 0) != cost_wrap8_scalar(a, b, P, bs)) return 1;
            if (cost_simd(a, b, P, bs, 1) != cost_sad_scalar(a, b, P, bs)) return 2;
            if (cost_simd(a, b, P, bs, 2) != cost_onesided_scalar(a, b, P, bs)) return 3;
        }
    }
#endif
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* Motion estimation.                                                                    */
/* ------------------------------------------------------------------------------------ */

/* Number of macroblocks: partial rows/columns are dropped
 * (InterframeCompression/motion.py:82-87).                                              */
int vcs_oracle_num_blocks(int H, int W, int bs) { return (H / bs) * (W / bs); }

/*
 * One macroblock of MotionProcessor._find_match + _get_motion_vector
 * (InterframeCompression/motion.py:100-161), with the candidate set written in the
 * interval form shared with the CUDA path:
 *
 *   rows  i = max(y+lo,0), +step, ... while i <= min(y+hi, H-bs-slack)
 *   cols  j = max(x+lo,0), +step, ... while j <= min(x+hi, W-bs-slack)
 *
 * The reference's own loop (motion.py:125-140: range(max(y-R,0), min(y+R,H), step), skip
 * when i+bs >= i_max) is exactly lo=-R, hi=R-bs-1, slack=1 with R = 2*bs (motion.py:18)
 * and step = round(bs/3) (motion.py:132).  The symmetric +/-R full search of
 * BASELINE.json configs 2/3/5 is lo=-R, hi=R, slack=0, step=1 (restated).
 * Scan order rows outer / cols inner ascending, strict '<' (motion.py:133,138,149).
 * With no candidate, best_coord stays [0,0] (motion.py:102) so mv = (-x,-y).
 */
static void me_one_block(const uint8_t *cur, const uint8_t *ref, int H, int W, int bs, int x,
                         int y, int lo, int hi, int step, int slack, int metric,
                         long long static_thr, int use_simd, int32_t *mv, uint32_t *cost,
                         uint8_t *flag) {
    const int pitch = 3 * W;
    const uint8_t *cblk = cur + (size_t)y * pitch + 3 * x;
    if (static_thr >= 0) {
        /* motion.py:109-116 static test */
        uint64_t S = block_cost(ref + (size_t)y * pitch + 3 * x, cblk, pitch, bs, 2, use_simd);
        if (S <= (uint64_t)static_thr) {
            mv[0] = 0; mv[1] = 0; *cost = (uint32_t)S; *flag = 1;
            return;
        }
    }
    int i0 = y + lo > 0 ? y + lo : 0, j0 = x + lo > 0 ? x + lo : 0;
    int i1 = y + hi < H - bs - slack ? y + hi : H - bs - slack;
    int j1 = x + hi < W - bs - slack ? x + hi : W - bs - slack;
    uint64_t best = 9999999999ull; /* motion.py:120 */
    int bx = 0, by = 0, found = 0;
    for (int i = i0; i <= i1; i += step)
        for (int j = j0; j <= j1; j += step) {
            uint64_t c = block_cost(ref + (size_t)i * pitch + 3 * j, cblk, pitch, bs, metric,
                                    use_simd);
            if (c < best) { best = c; bx = j; by = i; found = 1; }
        }
    mv[0] = bx - x; /* motion.py:156-161: [dx, dy] */
    mv[1] = by - y;
    *cost = found ? (uint32_t)best : 0xFFFFFFFFu;
    *flag = found ? 0 : 2;
}

/*
 * MotionProcessor.process_motion_prediction (InterframeCompression/motion.py:20-36) over
 * all macroblocks in raster order (motion.py:74-98).  mv: int32[N][2] = [dx,dy];
 * cost: best search cost (static blocks: the one-sided static sum; no candidate: 2^32-1);
 * flags: bit0 static, bit1 no candidate.  nthreads<=0: all cores.
 */
int vcs_oracle_me(const uint8_t *cur, const uint8_t *ref, int H, int W, int bs, int lo, int hi,
                  int step, int slack, int metric, long long static_thr, int use_simd,
                  int nthreads, int32_t *mv, uint32_t *cost, uint8_t *flags) {
    if (bs <= 0 || step <= 0 || H < bs || W < bs) return -1;
    const int nbx = W / bs, nby = H / bs;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 8) num_threads(nthreads)
#endif
    for (int k = 0; k < nbx * nby; ++k) {
        int x = (k % nbx) * bs, y = (k / nbx) * bs;
        me_one_block(cur, ref, H, W, bs, x, y, lo, hi, step, slack, metric, static_thr,
                     use_simd, mv + 2 * k, cost + k, flags + k);
    }
    return 0;
}

/* MotionProcessor.reconstruct_from_motion_vectors (motion.py:42-69): zero image, then one
 * bs x bs copy per macroblock from ref at (x+dx, y+dy); pixels not covered stay 0.       */
int vcs_oracle_mc(const uint8_t *ref, int H, int W, int bs, const int32_t *mv, uint8_t *pred) {
    const int pitch = 3 * W, nbx = W / bs, nby = H / bs;
    memset(pred, 0, (size_t)H * pitch);
    for (int k = 0; k < nbx * nby; ++k) {
        int x = (k % nbx) * bs, y = (k / nbx) * bs;
        int sx = x + mv[2 * k], sy = y + mv[2 * k + 1];
        if (sx < 0 || sy < 0 || sx + bs > W || sy + bs > H) return -2;
        for (int v = 0; v < bs; ++v)
            memcpy(pred + (size_t)(y + v) * pitch + 3 * x, ref + (size_t)(sy + v) * pitch + 3 * sx,
                   (size_t)3 * bs);
    }
    return 0;
}

/* MotionProcessor.get_residuals (motion.py:38-40): uint8 wrap-around subtract.           */
void vcs_oracle_residual(const uint8_t *cur, const uint8_t *pred, size_t n, uint8_t *resid) {
    for (size_t i = 0; i < n; ++i) resid[i] = (uint8_t)(cur[i] - pred[i]);
}

/* Decoder._fully_reconstruct (InterframeCompression/decoder.py:57): uint8 wrap add.      */
void vcs_oracle_add_wrap(const uint8_t *a, const uint8_t *b, size_t n, uint8_t *out) {
    for (size_t i = 0; i < n; ++i) out[i] = (uint8_t)(a[i] + b[i]);
}

/* ------------------------------------------------------------------------------------ */
/* Colour conversion: OpenCV 4.13 8-bit fixed point (14 fractional bits), restated.      */
/* Call sites: InterframeCompression/DCTcompressor.py:55,92; DCTCompression/dct.py:25,208 */
/* ------------------------------------------------------------------------------------ */
static inline int clip_u8(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

/* cv2.COLOR_BGR2YCR_CB; output channel order (Y, Cr, Cb) as cv2.split returns it.        */
void vcs_oracle_bgr2ycrcb(const uint8_t *bgr, size_t npix, uint8_t *ycrcb) {
    const int half = 1 << 13;
    for (size_t p = 0; p < npix; ++p) {
        int B = bgr[3 * p], G = bgr[3 * p + 1], R = bgr[3 * p + 2];
        int Y = (1868 * B + 9617 * G + 4899 * R + half) >> 14;
        int Cr = ((R - Y) * 11682 + (128 << 14) + half) >> 14;
        int Cb = ((B - Y) * 9241 + (128 << 14) + half) >> 14;
        ycrcb[3 * p] = (uint8_t)Y;
        ycrcb[3 * p + 1] = (uint8_t)clip_u8(Cr);
        ycrcb[3 * p + 2] = (uint8_t)clip_u8(Cb);
    }
}

/* cv2.COLOR_YCR_CB2BGR */
void vcs_oracle_ycrcb2bgr(const uint8_t *ycrcb, size_t npix, uint8_t *bgr) {
    const int half = 1 << 13;
    for (size_t p = 0; p < npix; ++p) {
        int Y = ycrcb[3 * p], Cr = ycrcb[3 * p + 1] - 128, Cb = ycrcb[3 * p + 2] - 128;
        int B = Y + ((Cb * 29049 + half) >> 14);
        int G = Y + ((Cb * -5636 + Cr * -11698 + half) >> 14);
        int R = Y + ((Cr * 22987 + half) >> 14);
        bgr[3 * p] = (uint8_t)clip_u8(B);
        bgr[3 * p + 1] = (uint8_t)clip_u8(G);
        bgr[3 * p + 2] = (uint8_t)clip_u8(R);
    }
}

/* ------------------------------------------------------------------------------------ */
/* 8x8 DCT / quantiser.                                                                  */
/* ------------------------------------------------------------------------------------ */

/* DCTCompressor._dctMatrix (InterframeCompression/DCTcompressor.py:124-133,
 * DCTCompression/dct.py:86-95) with libm cos/sqrt like Python's math module.            */
void vcs_oracle_dct_matrix(double *C /* [8][8] row-major */) {
    const int N = 8;
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            if (i == 0) C[i * N + j] = 1.0 / sqrt((double)N);
            else C[i * N + j] = sqrt(2.0 / N) * cos((double)((2 * j + 1) * i) * M_PI / (double)(2 * N));
        }
}

/* NB: entry [1][5] is 48 in the reference (DCTcompressor.py:12, dct.py:140), not Annex K's 58. */
static const int QY_TAB[64] = {16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  48,  60,  55,
                               14, 13, 16, 24, 40,  57,  69,  56,  14, 17, 22, 29, 51,  87,  80,  62,
                               18, 22, 37, 56, 68,  109, 103, 77,  24, 35, 55, 64, 81,  104, 113, 92,
                               49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
static const int QC_TAB[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99,
                               24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
                               99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                               99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};

/* JPEG Annex-K tables scaled by the quality factor
 * (InterframeCompression/DCTcompressor.py:11-38, DCTCompression/dct.py:139-166):
 * scale = 50/QF for 1<QF<50 else (100-QF)/50; Q = clip(np.round(T*scale),1,255), np.round is
 * half-to-even; channels (Y,Cr,Cb) use (QY,QC,QC).  Returns -1 for QF>=100 (the reference
 * prints a message and then fails on the undefined `scale`).                              */
int vcs_oracle_qtables(double qf, double *Q /* [3][64] */) {
    double scale;
    if (qf < 50 && qf > 1) scale = 50 / qf;
    else if (qf < 100) scale = (100 - qf) / 50;
    else return -1;
    for (int k = 0; k < 64; ++k) {
        double y = nearbyint((double)QY_TAB[k] * scale), c = nearbyint((double)QC_TAB[k] * scale);
        y = y < 1 ? 1 : (y > 255 ? 255 : y);
        c = c < 1 ? 1 : (c > 255 ? 255 : c);
        Q[k] = y; Q[64 + k] = c; Q[128 + k] = c;
    }
    return 0;
}

/* np.matmul of two 8x8 float64 matrices as OpenBLAS computes it: every element is a
 * sequential-k chain s = fma(a[i][k], b[k][j], s) from s = 0 (pinned by
 * tests/golden dct vectors).                                                              */
static void matmul8(const double *A, const double *B, double *Cout) {
    for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 8; ++j) {
            double s = 0.0;
            for (int k = 0; k < 8; ++k) s = fma(A[i * 8 + k], B[k * 8 + j], s);
            Cout[i * 8 + j] = s;
        }
}

/* DCTCompressor._dct2 (DCTcompressor.py:111-115): (C . X) . C^T                           */
void vcs_oracle_dct2(const double *X, double *D) {
    double C[64], Ct[64], T[64];
    vcs_oracle_dct_matrix(C);
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) Ct[i * 8 + j] = C[j * 8 + i];
    matmul8(C, X, T);
    matmul8(T, Ct, D);
}

/* DCTCompressor._idct2 (DCTcompressor.py:117-121): (C^T . X) . C                          */
void vcs_oracle_idct2(const double *X, double *P) {
    double C[64], Ct[64], T[64];
    vcs_oracle_dct_matrix(C);
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) Ct[i * 8 + j] = C[j * 8 + i];
    matmul8(Ct, X, T);
    matmul8(T, C, P);
}

/*
 * DCTCompressor.compress (InterframeCompression/DCTcompressor.py:49-74) for H, W multiples
 * of 8 (cv2.resize is then the identity): BGR->YCrCb, int16 - 128, per channel per 8x8
 * block D = dct2(block); out = D / Q[ch]  (np.true_divide, no rounding, :71).
 * round_mode 1 = DCTCompression/dct.py:179  np.round(np.divide(d, Q)) (half-to-even).
 * planes: 3 x H x W float64, same geometry as the image.
 */
static int compress_impl(const uint8_t *bgr, int H, int W, const double *Q, int round_mode,
                         int nthreads, double *planes, int8_t *idx8);

int vcs_oracle_compress(const uint8_t *bgr, int H, int W, const double *Q, int round_mode,
                        int nthreads, double *planes) {
    return compress_impl(bgr, H, W, Q, round_mode, nthreads, planes, NULL);
}

/* idx8 != NULL: the rounded indices are stored as int8 instead of float64 (the compact form the CUDA path
 * ships at QF <= 50, where |index| <= 1024 / min(Q) <= 127); same arithmetic, narrower store. */
static int compress_impl(const uint8_t *bgr, int H, int W, const double *Q, int round_mode,
                         int nthreads, double *planes, int8_t *idx8) {
    if (H % 8 || W % 8) return -1;
    double C[64], Ct[64];
    vcs_oracle_dct_matrix(C);
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) Ct[i * 8 + j] = C[j * 8 + i];
    const size_t npix = (size_t)H * W;
    uint8_t *ycc = (uint8_t *)malloc(3 * npix);
    if (!ycc) return -3;
    vcs_oracle_bgr2ycrcb(bgr, npix, ycc);
    const int nbx = W / 8, nby = H / 8;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(nthreads)
#endif
    for (int t = 0; t < 3 * nbx * nby; ++t) {
        int ch = t / (nbx * nby), b = t % (nbx * nby);
        int by = (b / nbx) * 8, bx = (b % nbx) * 8;
        double X[64], T[64], D[64];
        for (int v = 0; v < 8; ++v)
            for (int u = 0; u < 8; ++u)
                X[v * 8 + u] = (double)((int)ycc[3 * ((size_t)(by + v) * W + bx + u) + ch] - 128);
        matmul8(C, X, T);
        matmul8(T, Ct, D);
        for (int v = 0; v < 8; ++v)
            for (int u = 0; u < 8; ++u) {
                double q = D[v * 8 + u] / Q[ch * 64 + v * 8 + u];
                if (round_mode) q = nearbyint(q);
                if (idx8) idx8[(size_t)ch * npix + (size_t)(by + v) * W + bx + u] = (int8_t)q;
                else planes[(size_t)ch * npix + (size_t)(by + v) * W + bx + u] = q;
            }
    }
    free(ycc);
    return 0;
}

/*
 * DCTCompressor.decompress (InterframeCompression/DCTcompressor.py:76-93; twin
 * DCTCompression/dct.py:195-208): per block E = blk * Q[ch]; P = idct2(E); store into a
 * uint8 array = C cast (truncate toward zero, keep the low 8 bits); + 128 wraps;
 * dstack (Y,Cr,Cb); YCrCb->BGR.
 */
int vcs_oracle_decompress(const double *planes, int H, int W, const double *Q, int nthreads,
                          uint8_t *bgr) {
    if (H % 8 || W % 8) return -1;
    double C[64], Ct[64];
    vcs_oracle_dct_matrix(C);
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) Ct[i * 8 + j] = C[j * 8 + i];
    const size_t npix = (size_t)H * W;
    uint8_t *ycc = (uint8_t *)malloc(3 * npix);
    if (!ycc) return -3;
    const int nbx = W / 8, nby = H / 8;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(nthreads)
#endif
    for (int t = 0; t < 3 * nbx * nby; ++t) {
        int ch = t / (nbx * nby), b = t % (nbx * nby);
        int by = (b / nbx) * 8, bx = (b % nbx) * 8;
        double E[64], T[64], P[64];
        for (int v = 0; v < 8; ++v)
            for (int u = 0; u < 8; ++u)
                E[v * 8 + u] = planes[(size_t)ch * npix + (size_t)(by + v) * W + bx + u] *
                               Q[ch * 64 + v * 8 + u];
        matmul8(Ct, E, T);
        matmul8(T, C, P);
        for (int v = 0; v < 8; ++v)
            for (int u = 0; u < 8; ++u) {
                uint8_t p8 = (uint8_t)(int64_t)P[v * 8 + u];
                ycc[3 * ((size_t)(by + v) * W + bx + u) + ch] = (uint8_t)(p8 + 128);
            }
    }
    vcs_oracle_ycrcb2bgr(ycc, npix, bgr);
    free(ycc);
    return 0;
}

/*
 * Encoder._process_P_frame + Decoder._reconstruct_P_frame
 * (InterframeCompression/encoder.py:49-70, decoder.py:52-69) for one P-frame, with the ME
 * block size and the 8x8 DCT decoupled.  Any output pointer may be NULL.
 * This is the unit bench.py times as the CPU port.
 */
int vcs_oracle_encode_p(const uint8_t *cur, const uint8_t *ref, int H, int W, int bs, int lo,
                        int hi, int step, int slack, int metric, long long static_thr,
                        const double *Q, int round_mode, int use_simd, int nthreads,
                        int32_t *mv, uint32_t *cost, uint8_t *flags, double *planes,
                        uint8_t *recon) {
    const size_t n = (size_t)H * W * 3;
    const int N = vcs_oracle_num_blocks(H, W, bs);
    int rc = 0;
    int32_t *mv_l = mv ? mv : (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)N);
    uint32_t *cost_l = cost ? cost : (uint32_t *)malloc(sizeof(uint32_t) * (size_t)N);
    uint8_t *flags_l = flags ? flags : (uint8_t *)malloc((size_t)N);
    double *planes_l = planes ? planes : (double *)malloc(sizeof(double) * n);
    uint8_t *pred = (uint8_t *)malloc(n), *resid = (uint8_t *)malloc(n), *dec = (uint8_t *)malloc(n);
    if (!mv_l || !cost_l || !flags_l || !planes_l || !pred || !resid || !dec) { rc = -3; goto done; }
    rc = vcs_oracle_me(cur, ref, H, W, bs, lo, hi, step, slack, metric, static_thr, use_simd,
                       nthreads, mv_l, cost_l, flags_l);
    if (rc) goto done;
    rc = vcs_oracle_mc(ref, H, W, bs, mv_l, pred);
    if (rc) goto done;
    vcs_oracle_residual(cur, pred, n, resid);
    rc = vcs_oracle_compress(resid, H, W, Q, round_mode, nthreads, planes_l);
    if (rc) goto done;
    if (recon) {
        rc = vcs_oracle_decompress(planes_l, H, W, Q, nthreads, dec);
        if (rc) goto done;
        vcs_oracle_add_wrap(pred, dec, n, recon);
    }
done:
    if (!mv) free(mv_l);
    if (!cost) free(cost_l);
    if (!flags) free(flags_l);
    if (!planes) free(planes_l);
    free(pred); free(resid); free(dec);
    return rc;
}

/*
 * The forward half only (encoder.py:49-70), indices as int8: ME, MC, residual, DCT, rounded quantiser.  This is
 * exactly what bench.py's end-to-end GPU leg returns (mv, cost, flags, int8 indices), so it is the unit its CPU
 * arm times.  scratch: caller-provided 2 * H*W*3 bytes (pred, resid) so that nothing is allocated per frame.
 */
int vcs_oracle_encode_p_i8(const uint8_t *cur, const uint8_t *ref, int H, int W, int bs, int lo,
                           int hi, int step, int slack, int metric, long long static_thr,
                           const double *Q, int use_simd, int nthreads,
                           int32_t *mv, uint32_t *cost, uint8_t *flags, int8_t *idx8, uint8_t *scratch) {
    const size_t n = (size_t)H * W * 3;
    if (!mv || !cost || !flags || !idx8 || !scratch) return -1;
    uint8_t *pred = scratch, *resid = scratch + n;
    int rc = vcs_oracle_me(cur, ref, H, W, bs, lo, hi, step, slack, metric, static_thr, use_simd,
                           nthreads, mv, cost, flags);
    if (rc) return rc;
    rc = vcs_oracle_mc(ref, H, W, bs, mv, pred);
    if (rc) return rc;
    vcs_oracle_residual(cur, pred, n, resid);
    return compress_impl(resid, H, W, Q, 1, nthreads, NULL, idx8);
}

int vcs_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ==================================================================================== */
/* Intra mode decision: IntraframeCompression/intraframe.py + intramodes.py */
/* ==================================================================================== */
/* All predictor values are integers (the reference computes them with // in uint8 or float64
 * arithmetic), so the restatement works in int.  The reference's type quirks are reproduced:
 *  - 3*x//4 wraps mod 256 before the division when x is a uint8 pixel (a real neighbour slice),
 *    not when it is the float fallback 128 or a replicated value (intramodes.py:42,135);
 *  - dc4x4 adds u+l element-wise in uint8 (wrap) only when BOTH are real neighbours (intramodes.py:21);
 *  - neighbours are ORIGINAL pixels (intraframe.py:58-77), availability follows the aliased
 *    `available` list: first row: left only; first column: up (+up-right for 4x4); last 4x4
 *    column: no up-right (replicated u[3]); otherwise everything (intraframe.py:38-55);
 *  - chroma: the Cb UP neighbour is the RESIDUAL row above (Cbres, intraframe.py:266), which makes
 *    block rows of one column depend on each other; residuals/predictions can leave [0,255];
 *  - first strict minimum over the modes, initial best = bs*bs*255 (x2 for chroma) with an
 *    all-zero prediction and mode 0 if nothing beats it (intraframe.py:79-81).                  */

static inline int fdiv(int a, int b) { /* Python floor division, b > 0 */
    int q = a / b;
    return (a % b != 0 && a < 0) ? q - 1 : q;
}

/* one 4x4 predictor set; u8flags: bit0 u is uint8, bit1 ur is uint8, bit2 l is uint8 (for the wraps) */
static void pred4x4(int mode, int ul, const int *u, const int *ur, const int *l, int u_u8, int ur_u8, int l_u8,
                    int P[16]) {
#define Q4(x) fdiv((x), 4)
#define H2(x) fdiv((x), 2)
    const int t3ur = ur_u8 ? ((3 * ur[3]) & 255) / 4 : fdiv(3 * ur[3], 4);
    const int t3l = l_u8 ? ((3 * l[3]) & 255) / 4 : fdiv(3 * l[3], 4);
    switch (mode) {
    case 0: for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) P[i * 4 + j] = u[j]; break;
    case 1: for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) P[i * 4 + j] = l[i]; break; /* pred[:,c] = l (intramodes.py:16) */
    case 2: {
        int s = 0;
        for (int k = 0; k < 4; ++k) s += (u_u8 && l_u8) ? ((u[k] + l[k]) & 255) : (u[k] + l[k]);
        int avg = fdiv(s, 8);
        for (int k = 0; k < 16; ++k) P[k] = avg;
        break;
    }
    case 3: /* downleft4x4 (intramodes.py:26-44) */
        P[0] = Q4(u[0]) + H2(u[1]) + Q4(u[2]);
        P[1] = Q4(u[1]) + H2(u[2]) + Q4(u[3]); P[4] = P[1];
        P[2] = Q4(u[2]) + H2(u[3]) + Q4(ur[0]); P[5] = P[2]; P[8] = P[2];
        P[3] = Q4(u[3]) + H2(ur[0]) + Q4(ur[1]); P[6] = P[3]; P[9] = P[3]; P[12] = P[3];
        P[7] = Q4(ur[0]) + H2(ur[1]) + Q4(ur[2]); P[10] = P[7]; P[13] = P[7];
        P[11] = Q4(ur[1]) + H2(ur[2]) + Q4(ur[3]); P[14] = P[11];
        P[15] = Q4(ur[2]) + t3ur;
        break;
    case 4: /* downright4x4 (intramodes.py:46-64) */
        P[3] = Q4(u[1]) + H2(u[2]) + Q4(u[3]);
        P[2] = Q4(u[0]) + H2(u[1]) + Q4(u[2]); P[7] = P[2];
        P[1] = Q4(ul) + H2(u[0]) + Q4(u[1]); P[6] = P[1]; P[11] = P[1];
        P[0] = Q4(ul) + H2(u[0]) + Q4(l[0]); P[5] = P[0]; P[10] = P[0]; P[15] = P[0];
        P[4] = Q4(u[0]) + H2(l[0]) + Q4(l[1]); P[9] = P[4]; P[14] = P[4];
        P[8] = Q4(l[0]) + H2(l[1]) + Q4(l[2]); P[13] = P[8];
        P[12] = Q4(l[1]) + H2(l[2]) + Q4(l[3]);
        break;
    case 5: /* verticalright4x4 (intramodes.py:66-84) */
        P[0] = H2(ul) + H2(u[0]); P[9] = P[0];
        P[1] = H2(u[0]) + H2(u[1]); P[10] = P[1];
        P[2] = H2(u[1]) + H2(u[2]); P[11] = P[2];
        P[3] = H2(u[2]) + H2(u[3]);
        P[4] = Q4(u[0]) + H2(ul) + Q4(l[0]); P[13] = P[4];
        P[5] = Q4(ul) + H2(u[0]) + Q4(u[1]); P[14] = P[5];
        P[6] = Q4(u[0]) + H2(u[1]) + Q4(u[2]); P[15] = P[6];
        P[7] = Q4(u[1]) + H2(u[2]) + Q4(u[3]);
        P[8] = Q4(ul) + H2(l[0]) + Q4(l[1]);
        P[12] = Q4(l[0]) + H2(l[1]) + Q4(l[2]);
        break;
    case 6: /* horizontaldown4x4 (intramodes.py:86-104) */
        P[0] = H2(ul) + H2(l[0]); P[6] = P[0];
        P[1] = Q4(u[0]) + H2(ul) + Q4(l[0]); P[7] = P[1];
        P[2] = Q4(ul) + H2(u[0]) + Q4(u[1]);
        P[3] = Q4(u[0]) + H2(u[1]) + Q4(u[2]);
        P[4] = H2(l[0]) + H2(l[1]); P[10] = P[4];
        P[5] = Q4(ul) + H2(l[1]) + Q4(l[2]); P[11] = P[5];
        P[8] = H2(l[1]) + H2(l[2]); P[14] = P[8];
        P[9] = Q4(l[0]) + H2(l[1]) + Q4(l[2]); P[15] = P[9];
        P[12] = H2(l[2]) + H2(l[3]);
        P[13] = Q4(l[1]) + H2(l[2]) + Q4(l[3]);
        break;
    case 7: /* verticalleft4x4 (intramodes.py:106-124) */
        P[0] = H2(u[0]) + H2(u[1]);
        P[1] = H2(u[1]) + H2(u[2]); P[8] = P[1];
        P[2] = H2(u[2]) + H2(u[3]); P[9] = P[2];
        P[3] = H2(u[3]) + H2(ur[0]); P[10] = P[3];
        P[11] = H2(ur[0]) + H2(ur[1]);
        P[4] = Q4(u[0]) + H2(u[1]) + Q4(u[2]);
        P[5] = Q4(u[1]) + H2(u[2]) + Q4(u[3]); P[12] = P[5];
        P[6] = Q4(u[2]) + H2(u[3]) + Q4(ur[0]); P[13] = P[6];
        P[7] = Q4(u[3]) + H2(ur[0]) + Q4(ur[1]); P[14] = P[7];
        P[15] = Q4(ur[0]) + H2(ur[1]) + Q4(ur[2]);
        break;
    default: /* 8: horizontalup4x4 (intramodes.py:126-143) */
        P[0] = H2(l[0]) + H2(l[1]);
        P[1] = Q4(l[0]) + H2(l[1]) + Q4(l[2]);
        P[2] = H2(l[1]) + H2(l[2]); P[4] = P[2];
        P[3] = Q4(l[1]) + H2(l[2]) + Q4(l[3]); P[5] = P[3];
        P[6] = H2(l[2]) + H2(l[3]); P[8] = P[6];
        P[7] = Q4(l[2]) + t3l; P[9] = P[7];
        P[12] = l[3]; P[10] = l[3]; P[11] = l[3]; P[13] = l[3]; P[14] = l[3]; P[15] = l[3];
        break;
    }
#undef Q4
#undef H2
}

/* luma4x4 (IntraframeCompression/intraframe.py:24-151).  Y: H x W uint8 (multiples of 4).
 * res, pred: H x W int32; modes: (H/4) x (W/4) uint8. */
int vcs_oracle_luma4x4(const uint8_t *Y, int H, int W, int32_t *res, int32_t *pred, uint8_t *modes) {
    if (H % 4 || W % 4 || H <= 0 || W <= 0) return -1;
    const int mr = H / 4, mc = W / 4;
    for (int im = 0; im < mr; ++im)
        for (int jm = 0; jm < mc; ++jm) {
            const int i = im * 4, j = jm * 4;
            int s_ul = 0, s_u = 0, s_ur = 0, s_l = 0;
            if (im == 0 && jm == 0) { }
            else if (im == 0) s_l = 1;
            else if (jm == 0) { s_u = 1; s_ur = 1; }
            else if (jm + 1 == mc) { s_ul = 1; s_u = 1; s_l = 1; }
            else { s_ul = s_u = s_ur = s_l = 1; }
            if (mc == 1 && im > 0) s_ur = 0;  /* single column: available[...][jMac+1] raises IndexError in the
                                                 reference; a one-block-wide plane is outside its domain */
            int ul = s_ul ? Y[(i - 1) * W + j - 1] : 128, u[4], ur[4], l[4];
            for (int k = 0; k < 4; ++k) {
                u[k] = s_u ? Y[(i - 1) * W + j + k] : 128;
                ur[k] = s_ur ? Y[(i - 1) * W + j + 4 + k] : (s_u ? Y[(i - 1) * W + j + 3] : 128);
                l[k] = s_l ? Y[(i + k) * W + j - 1] : 128;
            }
            int best = 16 * 255, bmode = 0, bp[16] = {0};
            for (int m = 0; m < 9; ++m) {
                int P[16];
                pred4x4(m, ul, u, ur, l, s_u, s_ur, s_l, P);
                int d = 0;
                for (int a = 0; a < 4; ++a)
                    for (int b = 0; b < 4; ++b) d += abs(P[a * 4 + b] - (int)Y[(i + a) * W + j + b]);
                if (d < best) { best = d; bmode = m; memcpy(bp, P, sizeof(bp)); }
            }
            for (int a = 0; a < 4; ++a)
                for (int b = 0; b < 4; ++b) {
                    pred[(i + a) * W + j + b] = bp[a * 4 + b];
                    res[(i + a) * W + j + b] = (int)Y[(i + a) * W + j + b] - bp[a * 4 + b];
                }
            modes[im * mc + jm] = (uint8_t)bmode;
        }
    return 0;
}

/* luma16x16 (intraframe.py:153-225): vertical / horizontal / dc16x16 (intramodes.py:145-161). */
int vcs_oracle_luma16x16(const uint8_t *Y, int H, int W, int32_t *res, int32_t *pred, uint8_t *modes) {
    if (H % 16 || W % 16 || H <= 0 || W <= 0) return -1;
    const int mr = H / 16, mc = W / 16;
    for (int im = 0; im < mr; ++im)
        for (int jm = 0; jm < mc; ++jm) {
            const int i = im * 16, j = jm * 16;
            const int s_u = im > 0, s_l = jm > 0;   /* ul is fetched but never used by the three modes */
            int u[16], l[16];
            long su = 0, sl = 0;
            for (int k = 0; k < 16; ++k) {
                u[k] = s_u ? Y[(i - 1) * W + j + k] : 128;
                l[k] = s_l ? Y[(i + k) * W + j - 1] : 128;
                su += u[k]; sl += l[k];
            }
            const int dc = fdiv((int)(su + sl), 32);
            long best = 16 * 16 * 255; int bmode = 0, any = 0;
            for (int m = 0; m < 3; ++m) {
                long d = 0;
                for (int a = 0; a < 16; ++a)
                    for (int b = 0; b < 16; ++b) {
                        int pv = m == 0 ? u[b] : (m == 1 ? l[a] : dc);   /* horizontal16x16: pred[:,c] = l */
                        d += labs((long)pv - (long)Y[(i + a) * W + j + b]);
                    }
                if (d < best) { best = d; bmode = m; any = 1; }
            }
            for (int a = 0; a < 16; ++a)
                for (int b = 0; b < 16; ++b) {
                    int pv = !any ? 0 : (bmode == 0 ? u[b] : (bmode == 1 ? l[a] : dc));
                    pred[(i + a) * W + j + b] = pv;
                    res[(i + a) * W + j + b] = (int)Y[(i + a) * W + j + b] - pv;
                }
            modes[im * mc + jm] = (uint8_t)bmode;
        }
    return 0;
}

/* chroma8x8 (intraframe.py:228-317): joint mode for Cr and Cb; Cb's up neighbour is the RESIDUAL row
 * above (Cbres, intraframe.py:266). */
int vcs_oracle_chroma8x8(const uint8_t *Cr, const uint8_t *Cb, int H, int W, int32_t *crres, int32_t *crpred,
                         int32_t *cbres, int32_t *cbpred, uint8_t *modes) {
    if (H % 8 || W % 8 || H <= 0 || W <= 0) return -1;
    const int mr = H / 8, mc = W / 8;
    for (int im = 0; im < mr; ++im)
        for (int jm = 0; jm < mc; ++jm) {
            const int i = im * 8, j = jm * 8;
            const int s_u = im > 0, s_l = jm > 0;
            int ur[8], ub[8], lr[8], lb[8];
            long sur = 0, sub = 0, slr = 0, slb = 0;
            for (int k = 0; k < 8; ++k) {
                ur[k] = s_u ? Cr[(i - 1) * W + j + k] : 128;
                ub[k] = s_u ? cbres[(i - 1) * W + j + k] : 128;
                lr[k] = s_l ? Cr[(i + k) * W + j - 1] : 128;
                lb[k] = s_l ? Cb[(i + k) * W + j - 1] : 128;
                sur += ur[k]; sub += ub[k]; slr += lr[k]; slb += lb[k];
            }
            const int dcr = fdiv((int)(sur + slr), 16), dcb = fdiv((int)(sub + slb), 16);
            long best = 2 * 8 * 8 * 255; int bmode = 0, any = 0;
            for (int m = 0; m < 3; ++m) {
                long d = 0;
                for (int a = 0; a < 8; ++a)
                    for (int b = 0; b < 8; ++b) {
                        int pr = m == 0 ? ur[b] : (m == 1 ? lr[a] : dcr);
                        int pb = m == 0 ? ub[b] : (m == 1 ? lb[a] : dcb);
                        d += labs((long)pr - (long)Cr[(i + a) * W + j + b]) + labs((long)pb - (long)Cb[(i + a) * W + j + b]);
                    }
                if (d < best) { best = d; bmode = m; any = 1; }
            }
            for (int a = 0; a < 8; ++a)
                for (int b = 0; b < 8; ++b) {
                    int pr = !any ? 0 : (bmode == 0 ? ur[b] : (bmode == 1 ? lr[a] : dcr));
                    int pb = !any ? 0 : (bmode == 0 ? ub[b] : (bmode == 1 ? lb[a] : dcb));
                    crpred[(i + a) * W + j + b] = pr; crres[(i + a) * W + j + b] = (int)Cr[(i + a) * W + j + b] - pr;
                    cbpred[(i + a) * W + j + b] = pb; cbres[(i + a) * W + j + b] = (int)Cb[(i + a) * W + j + b] - pb;
                }
            modes[im * mc + jm] = (uint8_t)bmode;
        }
    return 0;
}

/* ---- 4:2:0 chroma subsampling demo (ChromaSubsampling/chroma.py) -----------------
 * chroma.py:9        imgYYC = cv2.cvtColor(img, COLOR_BGR2YCR_CB)
 * chroma.py:16-17    cr/cb = cv2.boxFilter(plane, ddepth=-1, ksize=(2,2))
 *                    OpenCV 4.13: anchor = ksize/2 = (1,1), BORDER_REFLECT_101, normalised; for uint8 the
 *                    2x2 mean comes out as ceil(sum/4) = (sum+3)>>2 (pinned for every sum 0..1020 by
 *                    tests/golden/make_golden_chroma.py), so
 *                    out[y][x] = (p[y-1][x-1] + p[y-1][x] + p[y][x-1] + p[y][x] + 3) >> 2, index -1 -> 1
 * chroma.py:20-21    samples = filtered[::2, ::2]  -> ceil(H/2) x ceil(W/2)
 * Y [H][W], crS/cbS [ceil(H/2)][ceil(W/2)]. */
static inline int refl101(int i, int n) { return n == 1 ? 0 : (i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i)); }

int vcs_oracle_chroma420(const uint8_t *bgr, int H, int W, uint8_t *Y, uint8_t *crS, uint8_t *cbS) {
    if (H <= 0 || W <= 0) return -1;
    const size_t npix = (size_t)H * W;
    uint8_t *ycc = (uint8_t *)malloc(npix * 3);
    if (!ycc) return -2;
    vcs_oracle_bgr2ycrcb(bgr, npix, ycc);
    for (size_t k = 0; k < npix; ++k) Y[k] = ycc[3 * k];
    const int h2 = (H + 1) / 2, w2 = (W + 1) / 2;
    for (int i = 0; i < h2; ++i)
        for (int j = 0; j < w2; ++j) {
            const int y0 = refl101(2 * i - 1, H), y1 = 2 * i, x0 = refl101(2 * j - 1, W), x1 = 2 * j;
            for (int c = 1; c <= 2; ++c) {
                const int s = ycc[((size_t)y0 * W + x0) * 3 + c] + ycc[((size_t)y0 * W + x1) * 3 + c] +
                              ycc[((size_t)y1 * W + x0) * 3 + c] + ycc[((size_t)y1 * W + x1) * 3 + c];
                (c == 1 ? crS : cbS)[(size_t)i * w2 + j] = (uint8_t)((s + 3) >> 2);
            }
        }
    free(ycc);
    return 0;
}

/* chroma.py:27-41: per pixel, with NumPy-2 scalar semantics as run in this container: Y, Cr, Cb are np.uint8
 * scalars, so `Cr - 128` WRAPS mod 256 (uint8 - weak Python int), the products with Python floats are float64,
 *   r = Y + 1.4022*crw;  g = (Y - 0.34414*cbw) - 0.71414*crw;  b = Y + 1.772*cbw
 * each clamped to [0,255] and truncated by the store into the uint8 image (chroma.py:41 `[b, g, r]`). */
int vcs_oracle_chroma420_to_bgr(const uint8_t *Y, const uint8_t *crS, const uint8_t *cbS, int H, int W, uint8_t *bgr) {
    if (H <= 0 || W <= 0) return -1;
    const int w2 = (W + 1) / 2;
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) {
            const double y = Y[(size_t)i * W + j];
            const double crw = (uint8_t)(crS[(size_t)(i / 2) * w2 + j / 2] - 128);
            const double cbw = (uint8_t)(cbS[(size_t)(i / 2) * w2 + j / 2] - 128);
            double r = y + 1.4022 * crw;
            double g = (y - 0.34414 * cbw) - 0.71414 * crw;
            double b = y + 1.77200 * cbw;
            r = r < 0 ? 0 : (r > 255 ? 255 : r);
            g = g < 0 ? 0 : (g > 255 ? 255 : g);
            b = b < 0 ? 0 : (b > 255 ? 255 : b);
            uint8_t *o = bgr + ((size_t)i * W + j) * 3;
            o[0] = (uint8_t)b; o[1] = (uint8_t)g; o[2] = (uint8_t)r;
        }
    return 0;
}
