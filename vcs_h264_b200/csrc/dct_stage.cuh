// dct_stage.cuh -- motion-compensated residual + 8x8 DCT / quantise / dequantise / IDCT /
// reconstruction, float64 "exact" arithmetic.
//
// What it replaces (reference file:line):
//   MotionProcessor.reconstruct_from_motion_vectors   InterframeCompression/motion.py:42-69
//   MotionProcessor.get_residuals                     motion.py:38-40
//   DCTCompressor.compress / _dct2 / _dctMatrix        DCTcompressor.py:49-74,111-133
//   rounded quantiser                                  DCTCompression/dct.py:169-186
//   DCTCompressor.decompress / _idct2                  DCTcompressor.py:76-93,117-121
//   Decoder._fully_reconstruct                         decoder.py:52-60
//
// Arithmetic contract (bit-exact with NumPy/OpenBLAS): every element of an
// 8x8 product is the chain s = fma(a_ik, b_kj, s), k = 0..7 ascending from s = 0, in IEEE
// double; the quantiser is an IEEE divide; rounding is rint (half-to-even); the store into a
// uint8 image is the C cast (uint8)(int64)x.  The file is compiled with -fmad=false and uses
// explicit __fma_rn so no other contraction can occur.
//
// Work decomposition: WARP-private tiles, no CTA barriers.  A warp owns a tile of 8 rows x 32 pixels
// (4 blocks x 3 channels) and walks through the stages with __syncwarp() only, so the warps of an SM
// are always in different stages and hide each other's latencies:
//   A  lane = one 8-pixel group: cur (3 x LDG.64) and the motion-compensated prediction (aligned
//      words + funnel shift), residual, BGR->YCrCb-128 packed as int8 into shared memory;
//   then, one channel at a time (DCT_NCH = 1; the loop is not unrolled, which keeps the kernel at <= 72
//   registers = 7 CTAs of 4 warps per SM and a third of the code):
//   B  lane = one pixel column: 8 inputs in registers, 8 DFMA chains, results in place (doubles, row stride 34);
//   C  lane = (block, row): row pass (all 8 chains first), quantise, coefficients straight to global
//      memory (8 int8 / int16 = one STG.64 / STG.128), E = q*Q back in place;
//   D/E the inverse column and row passes the same way, truncating store into packed bytes;
//   F  lane = one 8-pixel group: YCrCb->BGR, + prediction, 3 x STG.64.
// The kernel moves 12 / 15 / 33 B/px (int8 / int16 / float64 coefficients) but is bound by instruction issue
// and the FP64 pipe first: 96 DFMA per pixel plus one constant fetch per two of them.
//
// Quantiser: the reference computes rint(RN(D/Q)).  For the rounded modes the kernel takes
// q0 = D * RN(1/Q) (within 2^-40 of the true quotient for |q| < 2^11) and only when q0 lies within
// 2^-22 of a half-integer re-does that lane's row with the IEEE divide, so the result is bit-identical
// while the common path costs a DMUL.  The un-rounded mode (the reference's inter path,
// DCTcompressor.py:71) always uses the IEEE divide.
#pragma once
#include "common.cuh"

namespace vcs {

__constant__ double c_dct[64];  // _dctMatrix(), row-major, computed on the host with libm
__constant__ float c_dctf[64];  // the same values rounded to float: the fp32 tier (T2) only

// The transform arithmetic is written once over a real type R.  double = the exact tier (bit-identical with the reference,
// the default everywhere); float = the fp32 tier of SURVEY appendix A (T2): same kernel, FFMA instead of DFMA, no
// exact-quotient fallback -- its results are judged by tolerance and by counted flips (vcs_flip_counters_dev), never by equality.
template <typename R> struct DctReal;
template <> struct DctReal<double> {
    typedef double2 R2;
    static constexpr bool exact = true;
    static __device__ __forceinline__ double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
    static __device__ __forceinline__ double rnd(double x) { return rint(x); }
    static __device__ __forceinline__ double2 mk2(double a, double b) { return make_double2(a, b); }
    static __device__ __forceinline__ double cst(int k) { return c_dct[k]; }
};
template <> struct DctReal<float> {
    typedef float2 R2;
    static constexpr bool exact = false;
    static __device__ __forceinline__ float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
    static __device__ __forceinline__ float rnd(float x) { return rintf(x); }
    static __device__ __forceinline__ float2 mk2(float a, float b) { return make_float2(a, b); }
    static __device__ __forceinline__ float cst(int k) { return c_dctf[k]; }
};

constexpr int DCT_TILE_W = 32;                        // pixels per warp tile row (4 blocks)
constexpr int DCT_WARPS = 4;                          // warps per CTA, each with a private tile
constexpr int DCT_NCH = 1;                            // channels per pass group (1 or 3)
constexpr int DCT_THREADS = 32 * DCT_WARPS;
constexpr int DCT_RS = DCT_TILE_W + 2;                // row stride of the double tiles: even, so a lane's 8 row values are 4 aligned
                                                      // LDS.128, and (4i + 16 blk + 2k) mod 32 keeps column and row accesses conflict-free
constexpr int DCT_QS = 10;                            // row stride of the Q tables [ch][i][QS]: 20 i mod 32 words, conflict-free LDS.128
constexpr int DCT_QN = 3 * 8 * DCT_QS;                // doubles per table
constexpr int DCT_X_DOUBLES = DCT_NCH * 8 * DCT_RS;   // per-warp transform tile [NCH][8][RS]
constexpr int DCT_PLANE = 8 * DCT_TILE_W * 3;         // bytes of one 8 x 32 x 3 byte tile
template <typename R> struct DctSizes {
    static constexpr size_t WARP_BYTES = (size_t)DCT_X_DOUBLES * sizeof(R) + 3 * DCT_PLANE;   // transform tile + 3 byte tiles
    static constexpr size_t SMEM_BYTES = 2 * DCT_QN * sizeof(R) + DCT_WARPS * WARP_BYTES;
};
constexpr size_t DCT_SMEM_BYTES = DctSizes<double>::SMEM_BYTES;

struct DctArgs {
    int H, W;
    // forward stage input: image = cur (- pred gathered from ref by mv when mv != nullptr)
    FrameAddr fa;
    int has_fa;              // 0: `img` below is the only image, no prediction
    const uint8_t *img;      // plain image input (compress API) or nullptr
    const int16_t *mv;       // [nP][N][2] or nullptr
    int bs, nbx, nby;        // macroblock grid of the MVs
    const double *Q;         // [3][64]
    int forward;             // run cur -> coefficients
    int inverse;             // run coefficients -> pixels
    int coef_mode;           // VCS_COEF_*
    void *coef;              // [nP][3][H][W] output (forward) or input (inverse-only); may be null
    const uint8_t *pred_in;  // inverse-only: optional pred image to add (decoder.py:57)
    uint8_t *recon;          // [nP][H][W][3] or nullptr
    int *err;                // mapped host flag: set when a motion vector points outside the frame
    // packed sink (pack.cuh), int8 forward-only variant: per 8x8 block its occupancy bitmap (8 bytes, byte i = row i)
    // and escape count, per block row {bytes of its nibble stream, escapes} accumulated with one 64-bit atomic per
    // tile (zero the counts before the launch); all three null when unused
    uint8_t *bitmap;                  // [nP][3][H/8][W/8][8]
    uint8_t *blk_esc;                 // [nP][3][H/8][W/8]
    unsigned long long *row_count;    // [nP][3][H/8], low word = nibble bytes, high word = escapes
};

// 24 bytes (8 BGR pixels) starting at an arbitrary byte address, as 6 words
__device__ __forceinline__ void load24(const uint8_t *p, uint32_t w[6]) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    if ((a & 7) == 0) {
        const uint2 *q = reinterpret_cast<const uint2 *>(p);
        const uint2 v0 = __ldg(q), v1 = __ldg(q + 1), v2 = __ldg(q + 2);
        w[0] = v0.x; w[1] = v0.y; w[2] = v1.x; w[3] = v1.y; w[4] = v2.x; w[5] = v2.y;
        return;
    }
    const uint32_t *q = reinterpret_cast<const uint32_t *>(a & ~uintptr_t(3));
    const int sh = (int)(a & 3) * 8;
    uint32_t r[7];
#pragma unroll
    for (int k = 0; k < 6; ++k) r[k] = __ldg(q + k);
    r[6] = sh ? __ldg(q + 6) : 0u;
#pragma unroll
    for (int k = 0; k < 6; ++k) w[k] = sh ? __funnelshift_r(r[k], r[k + 1], sh) : r[k];
}

// byte k of a word array, zero-extended: one PRMT (the shift-and-mask form costs two ALU instructions)
__device__ __forceinline__ uint32_t byte_of(const uint32_t *w, int k) { return __byte_perm(w[k >> 2], 0u, 0x4440u | (uint32_t)(k & 3)); }
// word with its byte `pos` replaced by the low byte of v.  Position 0 STARTS a word with v as it is: the callers fill
// positions 0..3 in order, so whatever v carries above its low byte is overwritten by the next three calls.
__device__ __forceinline__ uint32_t put_byte(uint32_t word, uint32_t v, int pos) {
    return pos == 0 ? v : pos == 1 ? __byte_perm(word, v, 0x3240u)
         : pos == 2 ? __byte_perm(word, v, 0x3410u) : __byte_perm(word, v, 0x4210u);
}
// the same for 16-bit halves (positions 0, 1)
__device__ __forceinline__ uint32_t put_half(uint32_t word, uint32_t v, int pos) {
    return pos == 0 ? v : __byte_perm(word, v, 0x5410u);
}

// Compile-time variants: CM = coefficient format (VCS_COEF_*), PATH = which halves run.  The quantiser sits in the
// innermost loop, so run-time mode tests there cost more issue slots than the arithmetic they guard.
enum { DCT_FWD = 0, DCT_FWD_INV = 1, DCT_FWD_INV_NOCOEF = 2, DCT_INV = 3 };

#ifndef VCS_DCT_MINB_FWD
#define VCS_DCT_MINB_FWD 7
#endif
#ifndef VCS_DCT_MINB
#define VCS_DCT_MINB 7
#endif
// (A full-width specialisation that drops the partial-tile predicates was tried: without the branches ptxas merges the
// stages into one block, hoists loads across them and spills -- 104 bytes of stack at 80 registers.  Not kept.)
template <int CM, int PATH, typename R = double>
__global__ void __launch_bounds__(DCT_THREADS, DCT_NCH == 1 ? (PATH == DCT_FWD ? VCS_DCT_MINB_FWD : VCS_DCT_MINB) : 4)
dct_stage_kernel(const __grid_constant__ DctArgs a, int nP) {
    typedef DctReal<R> RT;
    typedef typename RT::R2 R2;
    constexpr bool forward = PATH != DCT_INV, do_inverse = PATH != DCT_FWD, has_coef = PATH != DCT_FWD_INV_NOCOEF;
    constexpr int coef_mode = CM;
    extern __shared__ __align__(16) unsigned char dct_smem[];
    R *s_q = reinterpret_cast<R *>(dct_smem);                    // Q [3][8][QS]
    R *s_rq = s_q + DCT_QN;                                      // RN(1/Q)
    // lane and the warp's shared-memory offset are pinned in registers (the empty asm makes them opaque): ptxas otherwise
    // re-derives them from %tid at every use, 19 S2R + 50 integer instructions per tile
    int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    uint32_t wofs = (uint32_t)(2 * DCT_QN * sizeof(R)) + (uint32_t)warp * (uint32_t)DctSizes<R>::WARP_BYTES;
    asm volatile("" : "+r"(lane));
    asm volatile("" : "+r"(wofs));
    unsigned char *wbase = dct_smem + wofs;
    R *s_x = reinterpret_cast<R *>(wbase);                                        // [NCH][8][RS]
    uint8_t *s_pred = wbase + DCT_X_DOUBLES * sizeof(R);                          // [8][32*3] prediction (BGR)
    int8_t *s_in8 = reinterpret_cast<int8_t *>(s_pred + DCT_PLANE);               // [3][8][32] YCrCb-128
    uint8_t *s_out = reinterpret_cast<uint8_t *>(s_in8 + DCT_PLANE);              // [3][8][32] decoded YCrCb

    const int W = a.W, H = a.H, bs = a.bs, nbx = a.nbx, nby = a.nby;
    size_t npix = (size_t)H * W;
    asm volatile("" : "+l"(npix));   // pinned like lane below: the 64-bit product was rebuilt at every use
    const int N = nbx * nby;

    // [ch][i][QS]: in the row passes a lane (block, row i) reads its 8 divisors as 4 aligned 16-byte words
    for (int k = threadIdx.x; k < 192; k += DCT_THREADS) {
        const double q = a.Q[k];
        const int ch = k >> 6, i = (k >> 3) & 7, j = k & 7;
        s_q[(ch * 8 + i) * DCT_QS + j] = (R)q;
        s_rq[(ch * 8 + i) * DCT_QS + j] = (R)(1.0 / q);
    }
    __syncthreads();   // the only CTA-wide barrier

    const int ntx = (W + DCT_TILE_W - 1) / DCT_TILE_W, nty = H / 8;
    const unsigned tiles_per_frame = (unsigned)(ntx * nty);
    const unsigned nitems = tiles_per_frame * (unsigned)nP;       // < 2^31 (checked on the host)
    const unsigned nwarps = gridDim.x * DCT_WARPS;
    const int bs_shift = (bs & (bs - 1)) == 0 ? 31 - __clz(bs) : -1;

    // lane roles
    const int g_r = lane >> 2, g_c = (lane & 3) * 8;        // stages A/F: 8-pixel group (row, first column)
    const int rp_i = lane & 7, rp_blk = lane >> 3;          // row passes: row i of block blk
    int rbase0 = rp_i * DCT_RS + rp_blk * 8;                // + ch * 8 * RS
    asm volatile("" : "+r"(rbase0));

    // The warp walks items first, first + nwarps, ...: (p, ty, tx) advance by constant steps with two carries, and the
    // frame pointers are refreshed only when p changes -- the divisions this replaces were 100 instructions per tile.
    const unsigned first = blockIdx.x * DCT_WARPS + warp;
    const unsigned step_p = nwarps / tiles_per_frame, step_r = nwarps - step_p * tiles_per_frame;
    const int step_ty = (int)(step_r / (unsigned)ntx), step_tx = (int)(step_r - (unsigned)step_ty * (unsigned)ntx);
    int p = (int)(first / tiles_per_frame), ty, tx;
    {
        const unsigned rem = first - (unsigned)p * tiles_per_frame;
        ty = (int)(rem / (unsigned)ntx);
        tx = (int)(rem - (unsigned)ty * (unsigned)ntx);
    }
    // Row 0 of the DCT matrix is the single value 1/sqrt(8) (vcs_dct_matrix, like _dctMatrix()): it lives in a register, which
    // saves the constant fetches of one chain in 8 in every pass.
    R a0 = RT::cst(0);
    if constexpr (RT::exact) asm volatile("" : "+d"(a0)); else asm volatile("" : "+f"(a0));
    int p_have = -1;
    const uint8_t *cur = nullptr, *ref = nullptr;
    for (unsigned item = first; item < nitems; item += nwarps, p += (int)step_p, ty += step_ty, tx += step_tx) {
        if (tx >= ntx) { tx -= ntx; ++ty; }
        if (ty >= nty) { ty -= nty; ++p; }
        const int x0 = tx * DCT_TILE_W, y0 = ty * 8;
        const int tw = min(DCT_TILE_W, W - x0);   // multiple of 8
        const bool g_on = g_c < tw, col_on = lane < tw, row_on = rp_blk * 8 < tw;
        if (p != p_have) {      // warp-uniform
            p_have = p;
            if (a.has_fa) {
                const unsigned pg = (unsigned)(p + a.fa.p_off), gop = pg / (unsigned)a.fa.ppg, in_gop = pg - gop * (unsigned)a.fa.ppg;
                cur = a.fa.cur_base + (long long)gop * a.fa.cur_gop_stride + (long long)in_gop * a.fa.cur_frame_stride;
                ref = a.fa.ref_base + (long long)gop * a.fa.ref_gop_stride;
            } else if (a.img) {
                cur = a.img + (size_t)p * npix * 3;
            }
        }
        const int16_t *mv = a.mv ? a.mv + (size_t)p * N * 2 : nullptr;

        // ---- A: gather, residual, colour ------------------------------------------------------------
        if (g_on) {
            const int x = x0 + g_c, y = y0 + g_r;
            uint32_t pw[6] = {0, 0, 0, 0, 0, 0};
            // prediction: motion-compensated gather from ref (motion.py:42-69) or a given pred image
            if (mv && ref) {
                if (bs % 8 == 0) {   // the group lies inside one macroblock
                    const int mbx = bs_shift >= 0 ? x >> bs_shift : x / bs, mby = bs_shift >= 0 ? y >> bs_shift : y / bs;
                    if (mbx < nbx && mby < nby) {   // uncovered border stays 0 (motion.py:45-46)
                        const int16_t *m = mv + 2 * (mby * nbx + mbx);
                        const int sx = x + m[0], sy = y + m[1];
                        // a vector that leaves the frame (corrupt or foreign input; the reference would raise,
                        // motion.py:62-65) is never followed: zero prediction, and the host is told
                        if (sx >= 0 && sy >= 0 && sx + 8 <= W && sy < H) load24(ref + ((size_t)sy * W + sx) * 3, pw);
                        else if (a.err) *(volatile int *)a.err = 1;
                    }
                } else {             // per-pixel gather (block sizes that are not multiples of 8)
                    for (int u = 0; u < 8; ++u) {
                        const int mbx = (x + u) / bs, mby = y / bs;
                        uint32_t b3 = 0;
                        if (mbx < nbx && mby < nby) {
                            const int16_t *m = mv + 2 * (mby * nbx + mbx);
                            const int sx = x + u + m[0], sy = y + m[1];
                            if (sx >= 0 && sy >= 0 && sx < W && sy < H) {
                                const uint8_t *rp = ref + ((size_t)sy * W + sx) * 3;
                                b3 = (uint32_t)__ldg(rp) | ((uint32_t)__ldg(rp + 1) << 8) | ((uint32_t)__ldg(rp + 2) << 16);
                            } else if (a.err) *(volatile int *)a.err = 1;
                        }
                        for (int e = 0; e < 3; ++e) s_pred[(g_r * DCT_TILE_W + g_c + u) * 3 + e] = (uint8_t)(b3 >> (8 * e));
                    }
                    const uint32_t *sp = reinterpret_cast<const uint32_t *>(s_pred + (g_r * DCT_TILE_W + g_c) * 3);
#pragma unroll
                    for (int k = 0; k < 6; ++k) pw[k] = sp[k];
                }
            } else if (a.pred_in) {
                load24(a.pred_in + (size_t)p * npix * 3 + ((size_t)y * W + x) * 3, pw);
            }
            if (forward) {
                uint32_t cw[6];
                load24(cur + ((size_t)y * W + x) * 3, cw);
                // residual wraps mod 256 per byte (motion.py:39) and is then treated as a BGR image
                uint32_t rw[6];
#pragma unroll
                for (int k = 0; k < 6; ++k) rw[k] = __vsub4(cw[k], pw[k]);
                uint32_t yv[2] = {0, 0}, crv[2] = {0, 0}, cbv[2] = {0, 0};
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int pb = byte_of(rw, 3 * u), pg = byte_of(rw, 3 * u + 1), pr = byte_of(rw, 3 * u + 2);
                    int Y, Cr, Cb;
                    bgr2ycrcb(pb, pg, pr, Y, Cr, Cb);
                    yv[u >> 2] = put_byte(yv[u >> 2], (uint32_t)Y, u & 3);
                    crv[u >> 2] = put_byte(crv[u >> 2], (uint32_t)Cr, u & 3);
                    cbv[u >> 2] = put_byte(cbv[u >> 2], (uint32_t)Cb, u & 3);
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {   // - 128 as int8 = flip bit 7 of every byte
                    yv[h] ^= 0x80808080u; crv[h] ^= 0x80808080u; cbv[h] ^= 0x80808080u;
                }
                *reinterpret_cast<uint2 *>(s_in8 + (0 * 8 + g_r) * DCT_TILE_W + g_c) = make_uint2(yv[0], yv[1]);
                *reinterpret_cast<uint2 *>(s_in8 + (1 * 8 + g_r) * DCT_TILE_W + g_c) = make_uint2(crv[0], crv[1]);
                *reinterpret_cast<uint2 *>(s_in8 + (2 * 8 + g_r) * DCT_TILE_W + g_c) = make_uint2(cbv[0], cbv[1]);
            }
            uint32_t *sp = reinterpret_cast<uint32_t *>(s_pred + (g_r * DCT_TILE_W + g_c) * 3);
#pragma unroll
            for (int k = 0; k < 6; ++k) sp[k] = pw[k];
        }
        __syncwarp();

        // this lane's coefficient row of channel 0, advanced by one plane per channel (a running pointer: rebuilding the
        // 64-bit index for every channel was 68 instructions per tile); the packed sink's outputs likewise
        static_assert(DCT_NCH == 1, "the running pointers below advance one channel per pass group");
        constexpr int ELEM = CM == 3 ? 1 : CM == 2 ? 2 : 8;
        unsigned char *cptr = reinterpret_cast<unsigned char *>(a.coef) +     // never dereferenced when there is no coefficient buffer
            ((size_t)p * 3 * npix + (size_t)((unsigned)(y0 + rp_i) * (unsigned)W + (unsigned)(x0 + rp_blk * 8))) * ELEM;
        const unsigned brows = (unsigned)(H / 8), bcols = (unsigned)(W / 8);
        size_t bidx = ((size_t)p * 3 * brows + (unsigned)ty) * bcols + (unsigned)(tx * 4 + rp_blk);   // block index, channel 0
        // Passes B-E run for NCH channels at a time (outer loop not unrolled): NCH = 3 fetches every DCT-matrix
        // constant once per 3 chains but needs ~120 registers and 3x the code; NCH = 1 halves both.
#pragma unroll 1
        for (int ch0 = 0; ch0 < 3; ch0 += DCT_NCH) {
        if (forward) {
            // ---- B: column pass  T = C . X  (T[i][j] = sum_k C[i][k] X[k][j]) -------------------------------
            if (col_on) {
                R xk[DCT_NCH][8];
#pragma unroll
                for (int c = 0; c < DCT_NCH; ++c)
#pragma unroll
                    for (int k = 0; k < 8; ++k) xk[c][k] = (R)(int)s_in8[((ch0 + c) * 8 + k) * DCT_TILE_W + lane];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    R s[DCT_NCH];
#pragma unroll
                    for (int c = 0; c < DCT_NCH; ++c) s[c] = (R)0;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const R cc = i == 0 ? a0 : RT::cst(i * 8 + k);
#pragma unroll
                        for (int c = 0; c < DCT_NCH; ++c) s[c] = RT::fma(cc, xk[c][k], s[c]);
                    }
#pragma unroll
                    for (int c = 0; c < DCT_NCH; ++c) s_x[(c * 8 + i) * DCT_RS + lane] = s[c];
                }
            }
            __syncwarp();
            // ---- C: row pass  D = T . C^T  (D[i][j] = sum_k T[i][k] C[j][k]); quantise; store ------------
            uint32_t nnz_lane[DCT_NCH];
#pragma unroll
            for (int c = 0; c < DCT_NCH; ++c) nnz_lane[c] = 0;
            if (row_on) {
                R tk[DCT_NCH][8];
#pragma unroll
                for (int c = 0; c < DCT_NCH; ++c)
#pragma unroll
                    for (int k = 0; k < 8; k += 2) {
                        const R2 t = *reinterpret_cast<const R2 *>(s_x + c * 8 * DCT_RS + rbase0 + k);
                        tk[c][k] = t.x; tk[c][k + 1] = t.y;
                    }
                // All 8 chains first (one basic block: the DFMAs of different j interleave), then the quantiser.
                // The rounded modes take q0 = D * RN(1/Q), within 2^-40 of the true quotient; a lane whose q0 comes within
                // 2^-22 of a half-integer is re-done with the IEEE quotient in a rarely taken second pass.  The test is on
                // the high word of |q0 - rint(q0)| (two integer instructions per index): >= 0x3FDFFFFF means
                // |.| >= 0.5 - 2^-22, far wider than the 2^-40 that could flip a rounding; second passes are exact.
                constexpr uint32_t NEAR_HALF_HI = 0x3FDFFFFFu;
                R sj[DCT_NCH][8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
#pragma unroll
                    for (int c = 0; c < DCT_NCH; ++c) sj[c][j] = (R)0;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const R cc = j == 0 ? a0 : RT::cst(j * 8 + k);
#pragma unroll
                        for (int c = 0; c < DCT_NCH; ++c) sj[c][j] = RT::fma(tk[c][k], cc, sj[c][j]);
                    }
                }
#pragma unroll
                for (int c = 0; c < DCT_NCH; ++c) {
                    const int ch = ch0 + c;
                    bool exact = coef_mode == 0 || !RT::exact;   // un-rounded mode: np.true_divide (DCTcompressor.py:71); fp32 tier: no second pass
                    uint32_t pk[4];
#pragma unroll 1
                    for (int attempt = 0; attempt < 2; ++attempt) {
                        uint32_t far = 0;               // max over the row of the high word of |q0 - rint(q0)|
                        R dprev = (R)0, eprev = (R)0;
                        R2 Q2 = RT::mk2((R)0, (R)0), rq2 = RT::mk2((R)0, (R)0);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if ((j & 1) == 0) {
                                Q2 = *reinterpret_cast<const R2 *>(s_q + (ch * 8 + rp_i) * DCT_QS + j);
                                if (coef_mode != 0) rq2 = *reinterpret_cast<const R2 *>(s_rq + (ch * 8 + rp_i) * DCT_QS + j);
                            }
                            const R Q = (j & 1) ? Q2.y : Q2.x;
                            R v;
                            if (coef_mode == 0) {
                                v = sj[c][j] / Q;
                            } else if (!RT::exact) {
                                v = RT::rnd(sj[c][j] * ((j & 1) ? rq2.y : rq2.x));   // fp32 tier: one multiply, flips are counted, not prevented
                            } else if (!exact) {
                                const R q0 = sj[c][j] * ((j & 1) ? rq2.y : rq2.x);
                                v = RT::rnd(q0);                           // np.round (dct.py:179) of RN(s/Q)
                                if constexpr (RT::exact) far = max(far, (uint32_t)__double2hiint(q0 - v) & 0x7fffffffu);
                            } else {
                                v = RT::rnd(sj[c][j] / Q);                 // the exact quotient decides
                            }
                            if (has_coef) {
                                if (coef_mode == 2) {
                                    pk[j >> 1] = put_half(pk[j >> 1], (uint32_t)(int)v, j & 1);
                                } else if (coef_mode == 3) {   // int8: lossless when 1024 / min(Q) <= 127 (checked on the host)
                                    pk[j >> 2] = put_byte(pk[j >> 2], (uint32_t)(int)v, j & 3);
                                } else if (j & 1) {
                                    *reinterpret_cast<double2 *>(reinterpret_cast<double *>(cptr) + j - 1) =
                                        make_double2((double)dprev, (double)v);
                                } else {
                                    dprev = v;
                                }
                            }
                            if (do_inverse) {                                           // E = blk * Q (DCTcompressor.py:86)
                                if (j & 1) *reinterpret_cast<R2 *>(s_x + c * 8 * DCT_RS + rbase0 + j - 1) = RT::mk2(eprev, v * Q);
                                else eprev = v * Q;
                            }
                        }
                        if (has_coef && coef_mode == 2)
                            *reinterpret_cast<uint4 *>(cptr) =
                                make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        else if (has_coef && coef_mode == 3)
                            *reinterpret_cast<uint2 *>(cptr) =
                                make_uint2(pk[0], pk[1]);
                        if (exact || far < NEAR_HALF_HI) break;   // per lane; the second attempt redoes this lane's row exactly
                        exact = true;
                    }
                    if (PATH == DCT_FWD && coef_mode == 3 && a.bitmap) {
                        // occupancy of this lane's 8 indices = byte rp_i of the block's bitmap; 32 lanes = 32 consecutive bytes
                        const uint32_t m8 = nz_nibble(pk[0]) | (nz_nibble(pk[1]) << 4);
                        a.bitmap[bidx * 8 + rp_i] = (uint8_t)m8;
                        // low half: non-zero indices, high half: those outside [-8, 7] (escapes of the nibble code)
                        nnz_lane[c] = __popc(m8) | ((__popc(esc_nibble(pk[0])) + __popc(esc_nibble(pk[1]))) << 16);
                    }
                }
            }
            if (PATH == DCT_FWD && coef_mode == 3 && a.bitmap) {       // warp-uniform: every lane takes part in the reduction
#pragma unroll
                for (int c = 0; c < DCT_NCH; ++c) {
                    uint32_t n = nnz_lane[c];
#pragma unroll
                    for (int o = 1; o < 8; o <<= 1) n += __shfl_xor_sync(0xffffffffu, n, o);     // the block's 8 rows
                    const uint32_t esc = n >> 16;
                    if (row_on && rp_i == 0)
                        a.blk_esc[bidx] = (uint8_t)esc;
                    unsigned long long t = (((n & 0xffffu) + 1) >> 1) | ((unsigned long long)esc << 32);   // nibble bytes | escapes
#pragma unroll
                    for (int o = 8; o < 32; o <<= 1) t += shfl_xor_u64(t, o);                    // the tile's 4 blocks
                    if (lane == 0 && t) atomicAdd(a.row_count + ((size_t)p * 3 + ch0 + c) * (H / 8) + ty, t);
                }
            }
        } else if (do_inverse && row_on) {
            // inverse-only: E = coefficient planes * Q
#pragma unroll
            for (int c = 0; c < DCT_NCH; ++c) {
                const int ch = ch0 + c;
                R qv[8];
                if (coef_mode == 2) {
                    const uint4 v = *reinterpret_cast<const uint4 *>(cptr);
                    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int j = 0; j < 8; ++j) qv[j] = (R)(int)(int16_t)((w[j >> 1] >> (16 * (j & 1))) & 0xffff);
                } else if (coef_mode == 3) {
                    const uint2 v = *reinterpret_cast<const uint2 *>(cptr);
#pragma unroll
                    for (int j = 0; j < 8; ++j) qv[j] = (R)(int)(int8_t)(((j < 4 ? v.x : v.y) >> (8 * (j & 3))) & 0xff);
                } else {
                    const double *o = reinterpret_cast<const double *>(cptr);
#pragma unroll
                    for (int j = 0; j < 8; j += 2) {
                        const double2 v = *reinterpret_cast<const double2 *>(o + j);
                        qv[j] = (R)v.x; qv[j + 1] = (R)v.y;
                    }
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) s_x[c * 8 * DCT_RS + rbase0 + j] = qv[j] * s_q[(ch * 8 + rp_i) * DCT_QS + j];
            }
        }
        if (do_inverse) {
            __syncwarp();
            // ---- D: inverse column pass  T' = C^T . E  (T'[i][j] = sum_k C[k][i] E[k][j]) ---------------------
            if (col_on) {
                R ek[DCT_NCH][8];
#pragma unroll
                for (int c = 0; c < DCT_NCH; ++c)
#pragma unroll
                    for (int k = 0; k < 8; ++k) ek[c][k] = s_x[(c * 8 + k) * DCT_RS + lane];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    R s[DCT_NCH];
#pragma unroll
                    for (int c = 0; c < DCT_NCH; ++c) s[c] = (R)0;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const R cc = k == 0 ? a0 : RT::cst(k * 8 + i);
#pragma unroll
                        for (int c = 0; c < DCT_NCH; ++c) s[c] = RT::fma(cc, ek[c][k], s[c]);
                    }
#pragma unroll
                    for (int c = 0; c < DCT_NCH; ++c) s_x[(c * 8 + i) * DCT_RS + lane] = s[c];
                }
            }
            __syncwarp();
            // ---- E: inverse row pass  P = T' . C ; truncating uint8 store ; +128 ---------------------------------
            if (row_on) {
                R tk[DCT_NCH][8];
#pragma unroll
                for (int c = 0; c < DCT_NCH; ++c)
#pragma unroll
                    for (int k = 0; k < 8; k += 2) {
                        const R2 t = *reinterpret_cast<const R2 *>(s_x + c * 8 * DCT_RS + rbase0 + k);
                        tk[c][k] = t.x; tk[c][k + 1] = t.y;
                    }
                uint32_t w[DCT_NCH][2];
#pragma unroll
                for (int c = 0; c < DCT_NCH; ++c) w[c][0] = w[c][1] = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    R s[DCT_NCH];
#pragma unroll
                    for (int c = 0; c < DCT_NCH; ++c) s[c] = (R)0;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const R cc = k == 0 ? a0 : RT::cst(k * 8 + j);
#pragma unroll
                        for (int c = 0; c < DCT_NCH; ++c) s[c] = RT::fma(tk[c][k], cc, s[c]);
                    }
                    // float64 -> uint8 store (DCTcompressor.py:81,88): truncate toward zero, low 8 bits; then +128
#pragma unroll
                    for (int c = 0; c < DCT_NCH; ++c) w[c][j >> 2] = put_byte(w[c][j >> 2], (uint32_t)(long long)s[c], j & 3);
                }
#pragma unroll
                for (int c = 0; c < DCT_NCH; ++c) { w[c][0] ^= 0x80808080u; w[c][1] ^= 0x80808080u; }   // + 128 (mod 256)
#pragma unroll
                for (int c = 0; c < DCT_NCH; ++c)
                    *reinterpret_cast<uint2 *>(s_out + ((ch0 + c) * 8 + rp_i) * DCT_TILE_W + rp_blk * 8) = make_uint2(w[c][0], w[c][1]);
            }
        }
        cptr += npix * ELEM;
        bidx += (size_t)brows * bcols;
        }   // channel groups
        if (do_inverse) {
            __syncwarp();
            // ---- F: YCrCb -> BGR, + pred (wrap) -----------------------------------------------------------------------
            if (g_on) {
                const uint2 yv = *reinterpret_cast<const uint2 *>(s_out + (0 * 8 + g_r) * DCT_TILE_W + g_c);
                const uint2 crv = *reinterpret_cast<const uint2 *>(s_out + (1 * 8 + g_r) * DCT_TILE_W + g_c);
                const uint2 cbv = *reinterpret_cast<const uint2 *>(s_out + (2 * 8 + g_r) * DCT_TILE_W + g_c);
                const uint32_t *sp = reinterpret_cast<const uint32_t *>(s_pred + (g_r * DCT_TILE_W + g_c) * 3);
                uint32_t ow[6] = {0, 0, 0, 0, 0, 0};
                const uint32_t yw[2] = {yv.x, yv.y}, crw[2] = {crv.x, crv.y}, cbw[2] = {cbv.x, cbv.y};
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int Y = byte_of(yw, u), Cr = byte_of(crw, u), Cb = byte_of(cbw, u);
                    int pb, pg, pr;
                    ycrcb2bgr(Y, Cr, Cb, pb, pg, pr);
                    ow[(3 * u) >> 2] = put_byte(ow[(3 * u) >> 2], (uint32_t)pb, (3 * u) & 3);
                    ow[(3 * u + 1) >> 2] = put_byte(ow[(3 * u + 1) >> 2], (uint32_t)pg, (3 * u + 1) & 3);
                    ow[(3 * u + 2) >> 2] = put_byte(ow[(3 * u + 2) >> 2], (uint32_t)pr, (3 * u + 2) & 3);
                }
                uint32_t ov[6];   // pred + decoded, uint8 wrap (decoder.py:57)
#pragma unroll
                for (int k = 0; k < 6; ++k) ov[k] = __vadd4(ow[k], sp[k]);
                uint8_t *ob = a.recon + (size_t)p * npix * 3 + ((size_t)(y0 + g_r) * W + x0 + g_c) * 3;
                if ((reinterpret_cast<uintptr_t>(ob) & 7) == 0) {
                    uint2 *op = reinterpret_cast<uint2 *>(ob);
                    op[0] = make_uint2(ov[0], ov[1]); op[1] = make_uint2(ov[2], ov[3]); op[2] = make_uint2(ov[4], ov[5]);
                } else {
#pragma unroll
                    for (int k = 0; k < 24; ++k) ob[k] = (uint8_t)(ov[k >> 2] >> (8 * (k & 3)));
                }
            }
        }
        __syncwarp();   // the next item reuses this warp's tiles
    }
}

// MotionProcessor.reconstruct_from_motion_vectors (motion.py:42-69) as its own kernel, for the
// drop-in method; the fused path above never materialises pred.  One thread = 4 pixels = 12 bytes = 3 aligned
// words of the output row (W % 4 == 0 fast path: a 4-pixel group never straddles a macroblock when bs % 4 == 0).
__global__ void mc_kernel(const uint8_t *__restrict__ ref, const int16_t *__restrict__ mv, int H,
                          int W, int bs, int nbx, int nby, uint8_t *__restrict__ pred, int *err) {
    const size_t npix = (size_t)H * W;
    if ((W & 3) == 0 && (bs & 3) == 0 && (reinterpret_cast<uintptr_t>(pred) & 3) == 0) {
        const size_t ngrp = npix / 4;
        for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < ngrp; k += (size_t)gridDim.x * blockDim.x) {
            const int y = (int)(k / (W / 4)), x = (int)(k - (size_t)y * (W / 4)) * 4;
            const int mbx = x / bs, mby = y / bs;
            uint32_t w[3] = {0, 0, 0};
            if (mbx < nbx && mby < nby) {
                const int16_t *m = mv + 2 * (mby * nbx + mbx);
                const int sx = x + m[0], sy = y + m[1];
                if (sx >= 0 && sy >= 0 && sx + 4 <= W && sy < H) {
                    const uint8_t *rp = ref + ((size_t)sy * W + sx) * 3;
                    const uintptr_t a = reinterpret_cast<uintptr_t>(rp);
                    const uint32_t *q = reinterpret_cast<const uint32_t *>(a & ~uintptr_t(3));
                    const int sh = (int)(a & 3) * 8;
                    const uint32_t r0 = __ldg(q), r1 = __ldg(q + 1), r2 = __ldg(q + 2), r3 = sh ? __ldg(q + 3) : 0u;
                    w[0] = __funnelshift_r(r0, r1, sh); w[1] = __funnelshift_r(r1, r2, sh); w[2] = __funnelshift_r(r2, r3, sh);
                } else if (err) *(volatile int *)err = 1;
            }
            uint32_t *o = reinterpret_cast<uint32_t *>(pred + 12 * k);
            o[0] = w[0]; o[1] = w[1]; o[2] = w[2];
        }
        return;
    }
    for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < npix;
         k += (size_t)gridDim.x * blockDim.x) {
        const int y = (int)(k / W), x = (int)(k - (size_t)y * W);
        const int mbx = x / bs, mby = y / bs;
        uint8_t b = 0, g = 0, r = 0;
        if (mbx < nbx && mby < nby) {
            const int16_t *m = mv + 2 * (mby * nbx + mbx);
            const int sx = x + m[0], sy = y + m[1];
            if (sx >= 0 && sy >= 0 && sx < W && sy < H) {
                const uint8_t *rp = ref + ((size_t)sy * W + sx) * 3;
                b = rp[0]; g = rp[1]; r = rp[2];
            } else if (err) *(volatile int *)err = 1;
        }
        pred[3 * k] = b; pred[3 * k + 1] = g; pred[3 * k + 2] = r;
    }
}

// Non-zero count of quantised coefficients: the numerator of DCTCompression/dct.py:188-191's sparsity
// print (1 - nnz/size).  ELEM: 2 = int16 planes, 8 = float64 planes.
template <int ELEM>
__global__ void count_nonzero_kernel(const void *__restrict__ coef, size_t n, unsigned long long *__restrict__ out) {
    unsigned long long c = 0;
    for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        if (ELEM == 1) c += reinterpret_cast<const int8_t *>(coef)[k] != 0;
        else if (ELEM == 2) c += reinterpret_cast<const int16_t *>(coef)[k] != 0;
        else c += reinterpret_cast<const double *>(coef)[k] != 0.0;
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) c += shfl_xor_u64(c, m);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

// Flip counters of the fp32 tier: how many elements of two equally shaped arrays differ
// (rounded indices that crossed a rounding boundary, pixels that crossed a truncation boundary) and, for pixels, the sum
// of squared differences (PSNR of one against the other).
template <typename E>
__global__ void flip_count_kernel(const E *__restrict__ x, const E *__restrict__ y, size_t n, unsigned long long *__restrict__ flips,
                                  unsigned long long *__restrict__ sse) {
    unsigned long long f = 0, q = 0;
    for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        const int d = (int)x[k] - (int)y[k];
        f += d != 0;
        q += (unsigned long long)(d * d);
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) { f += shfl_xor_u64(f, m); q += shfl_xor_u64(q, m); }
    if ((threadIdx.x & 31) == 0) {
        if (f) atomicAdd(flips, f);
        if (sse && q) atomicAdd(sse, q);
    }
}

// get_residuals (motion.py:38-40) / _fully_reconstruct (decoder.py:57): byte-wise wrap, 16 bytes per thread
// when the three pointers are 16-byte aligned (__vsub4 / __vadd4 on each word), byte-wise otherwise and for the tail.
template <int ADD>
__global__ void wrap_kernel(const uint8_t *__restrict__ a, const uint8_t *__restrict__ b, size_t n,
                            uint8_t *__restrict__ out) {
    size_t done = 0;
    if (((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
        const size_t nv = n / 16;
        const uint4 *av = reinterpret_cast<const uint4 *>(a), *bv = reinterpret_cast<const uint4 *>(b);
        uint4 *ov = reinterpret_cast<uint4 *>(out);
        for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < nv; k += (size_t)gridDim.x * blockDim.x) {
            const uint4 x = __ldg(av + k), y = __ldg(bv + k);
            ov[k] = ADD ? make_uint4(__vadd4(x.x, y.x), __vadd4(x.y, y.y), __vadd4(x.z, y.z), __vadd4(x.w, y.w))
                        : make_uint4(__vsub4(x.x, y.x), __vsub4(x.y, y.y), __vsub4(x.z, y.z), __vsub4(x.w, y.w));
        }
        done = nv * 16;
    }
    for (size_t k = done + blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n;
         k += (size_t)gridDim.x * blockDim.x)
        out[k] = ADD ? (uint8_t)(a[k] + b[k]) : (uint8_t)(a[k] - b[k]);
}

// DCTCompressor._dct2 / _idct2 (DCTcompressor.py:111-121) on bare 8x8 float64 blocks: C.X.C^T or C^T.X.C as two
// products of sequential-k FMA chains (what np.matmul does for these shapes).  One thread per
// output element; the private helpers of the class surface call this, the hot path uses dct_stage_kernel.
__global__ void dct2_blocks_kernel(const double *__restrict__ in, int nblocks, int inverse, double *__restrict__ out) {
    __shared__ double sx[4][64], st[4][64];
    const int t = threadIdx.x & 63, lb = threadIdx.x >> 6, i = t >> 3, j = t & 7;
    for (int b0 = blockIdx.x * 4; b0 < nblocks; b0 += gridDim.x * 4) {
        const int b = b0 + lb;
        if (b < nblocks) sx[lb][t] = in[(size_t)b * 64 + t];
        __syncthreads();
        if (b < nblocks) {
            double s = 0.0;   // T = C.X (forward) or C^T.X (inverse)
#pragma unroll
            for (int k = 0; k < 8; ++k) s = __fma_rn(inverse ? c_dct[k * 8 + i] : c_dct[i * 8 + k], sx[lb][k * 8 + j], s);
            st[lb][t] = s;
        }
        __syncthreads();
        if (b < nblocks) {
            double s = 0.0;   // T.C^T (forward) or T.C (inverse)
#pragma unroll
            for (int k = 0; k < 8; ++k) s = __fma_rn(st[lb][i * 8 + k], inverse ? c_dct[k * 8 + j] : c_dct[j * 8 + k], s);
            out[(size_t)b * 64 + t] = s;
        }
        __syncthreads();
    }
}

}  // namespace vcs
