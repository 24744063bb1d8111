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
// Work decomposition: one CTA per tile of 8 rows x 128 pixels (16 blocks x 3 channels), 384
// threads = (block-channel, lane-in-block).  Column pass and row pass of each transform are
// thread-local 8-term chains on shared-memory tiles; all global traffic is staged through
// shared memory so it is coalesced.  The kernel is HBM-bound: per pixel it reads cur 3 B +
// ref 3 B and writes coefficients (24 B f64 | 6 B int16) and 3 B of reconstruction.
#pragma once
#include "common.cuh"

namespace vcs {

__constant__ double c_dct[64];  // _dctMatrix(), row-major, computed on the host with libm

constexpr int DCT_TILE_W = 128;                 // pixels per tile row
constexpr int DCT_THREADS = 3 * (DCT_TILE_W / 8) * 8;  // 384
constexpr int DCT_RS = DCT_TILE_W + 1;          // padded row stride (doubles): conflict-free
constexpr size_t DCT_SMEM_BYTES = (size_t)(2 * 3 * 8 * DCT_RS + 192) * sizeof(double) +
                                  2 * 8 * DCT_TILE_W * 3;

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
};

__device__ __forceinline__ double quantise(double d, double q, int coef_mode) {
    double v = d / q;  // np.true_divide (DCTcompressor.py:71)
    return coef_mode == 0 ? v : rint(v);  // np.round (dct.py:179)
}

__global__ void __launch_bounds__(DCT_THREADS)
dct_stage_kernel(DctArgs a) {
    extern __shared__ __align__(16) unsigned char dct_smem[];
    double *s_a = reinterpret_cast<double *>(dct_smem);
    double *s_b = s_a + 3 * 8 * DCT_RS;
    double *s_q = s_b + 3 * 8 * DCT_RS;
    uint8_t *s_pred = reinterpret_cast<uint8_t *>(s_q + 192);
    uint8_t *s_ycc = s_pred + 8 * DCT_TILE_W * 3;

    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * DCT_TILE_W, y0 = blockIdx.y * 8, p = blockIdx.z;
    const int tw = min(DCT_TILE_W, a.W - x0);  // multiple of 8
    const size_t npix = (size_t)a.H * a.W;
    const int N = a.nbx * a.nby;

    for (int k = tid; k < 192; k += DCT_THREADS) s_q[k] = a.Q[k];

    const uint8_t *cur = nullptr, *ref = nullptr;
    if (a.has_fa) {
        cur = cur_frame(a.fa, p);
        ref = ref_frame(a.fa, p);
    } else if (a.img) {
        cur = a.img + (size_t)p * npix * 3;
    }
    const int16_t *mv = a.mv ? a.mv + (size_t)p * N * 2 : nullptr;

    // ---- 1. gather: pred (MC), residual, BGR->YCrCb, -128 ---------------------------------
    if (a.forward) {
        for (int k = tid; k < 8 * DCT_TILE_W; k += DCT_THREADS) {
            const int r = k / DCT_TILE_W, c = k - r * DCT_TILE_W;
            if (c >= tw) continue;
            const int x = x0 + c, y = y0 + r;
            const uint8_t *cp = cur + ((size_t)y * a.W + x) * 3;
            int pb = 0, pg = 0, pr = 0;
            if (mv) {
                const int mbx = x / a.bs, mby = y / a.bs;
                if (mbx < a.nbx && mby < a.nby) {  // uncovered border stays 0 (motion.py:45-46)
                    const int16_t *m = mv + 2 * (mby * a.nbx + mbx);
                    const uint8_t *rp = ref + ((size_t)(y + m[1]) * a.W + (x + m[0])) * 3;
                    pb = __ldg(rp); pg = __ldg(rp + 1); pr = __ldg(rp + 2);
                }
            }
            s_pred[3 * k] = (uint8_t)pb; s_pred[3 * k + 1] = (uint8_t)pg; s_pred[3 * k + 2] = (uint8_t)pr;
            // residual wraps mod 256 (motion.py:39) and is then treated as a BGR image
            const int B = (uint8_t)(__ldg(cp) - pb), G = (uint8_t)(__ldg(cp + 1) - pg),
                      R = (uint8_t)(__ldg(cp + 2) - pr);
            int Y, Cr, Cb;
            bgr2ycrcb(B, G, R, Y, Cr, Cb);
            s_a[(0 * 8 + r) * DCT_RS + c] = (double)(Y - 128);
            s_a[(1 * 8 + r) * DCT_RS + c] = (double)(Cr - 128);
            s_a[(2 * 8 + r) * DCT_RS + c] = (double)(Cb - 128);
        }
    } else {
        // inverse-only: coefficient planes in, optional pred image
        for (int k = tid; k < 3 * 8 * DCT_TILE_W; k += DCT_THREADS) {
            const int ch = k / (8 * DCT_TILE_W), rem = k - ch * 8 * DCT_TILE_W;
            const int r = rem / DCT_TILE_W, c = rem - r * DCT_TILE_W;
            if (c >= tw) continue;
            const size_t gi = ((size_t)p * 3 + ch) * npix + (size_t)(y0 + r) * a.W + x0 + c;
            double v = a.coef_mode == 2 ? (double)((const int16_t *)a.coef)[gi]
                                        : ((const double *)a.coef)[gi];
            s_a[(ch * 8 + r) * DCT_RS + c] = v;
        }
        for (int k = tid; k < 8 * DCT_TILE_W * 3; k += DCT_THREADS) {
            const int r = k / (DCT_TILE_W * 3), cb = k - r * DCT_TILE_W * 3;
            uint8_t v = 0;
            if (a.pred_in && cb < tw * 3)
                v = a.pred_in[(size_t)p * npix * 3 + ((size_t)(y0 + r) * a.W + x0) * 3 + cb];
            s_pred[k] = v;
        }
    }
    __syncthreads();

    // thread roles for the transform passes
    const int colpass_ch = tid / DCT_TILE_W, colpass_c = tid - colpass_ch * DCT_TILE_W;  // (ch, column)
    // row pass: lane = i + 8 * (blk & 3) keeps the 8-byte shared accesses conflict-free
    const int rp_ch = tid / 128, rp_rem = tid - rp_ch * 128;
    const int rp_i = rp_rem & 7, rp_blk = ((rp_rem >> 5) << 2) | ((rp_rem >> 3) & 3);
    const bool col_on = colpass_c < tw, row_on = rp_blk * 8 < tw;

    if (a.forward) {
        // ---- 2. column pass  T = C . X  (T[i][j] = sum_k C[i][k] X[k][j]) ------------------
        if (col_on) {
            double xk[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) xk[k] = s_a[(colpass_ch * 8 + k) * DCT_RS + colpass_c];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < 8; ++k) s = __fma_rn(c_dct[i * 8 + k], xk[k], s);
                s_b[(colpass_ch * 8 + i) * DCT_RS + colpass_c] = s;
            }
        }
        __syncthreads();
        // ---- 3. row pass  D = T . C^T  (D[i][j] = sum_k T[i][k] C[j][k]);  / Q -------------
        if (row_on) {
            double tk[8];
            const int base = (rp_ch * 8 + rp_i) * DCT_RS + rp_blk * 8;
#pragma unroll
            for (int k = 0; k < 8; ++k) tk[k] = s_b[base + k];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < 8; ++k) s = __fma_rn(tk[k], c_dct[j * 8 + k], s);
                s_a[base + j] = quantise(s, s_q[rp_ch * 64 + rp_i * 8 + j], a.coef_mode);
            }
        }
        __syncthreads();
        // ---- 4. coalesced coefficient store ------------------------------------------------
        if (a.coef) {
            for (int k = tid; k < 3 * 8 * DCT_TILE_W; k += DCT_THREADS) {
                const int ch = k / (8 * DCT_TILE_W), rem = k - ch * 8 * DCT_TILE_W;
                const int r = rem / DCT_TILE_W, c = rem - r * DCT_TILE_W;
                if (c >= tw) continue;
                const size_t gi = ((size_t)p * 3 + ch) * npix + (size_t)(y0 + r) * a.W + x0 + c;
                const double v = s_a[(ch * 8 + r) * DCT_RS + c];
                if (a.coef_mode == 2) ((int16_t *)a.coef)[gi] = (int16_t)(int)v;
                else ((double *)a.coef)[gi] = v;
            }
        }
    }
    if (!a.inverse || !a.recon) return;

    // ---- 5. dequantise + column pass  T' = C^T . E  (T'[i][j] = sum_k C[k][i] E[k][j]) ------
    if (col_on) {
        double ek[8];
        const int j = colpass_c & 7;
#pragma unroll
        for (int k = 0; k < 8; ++k)  // np.multiply(block, Q) (DCTcompressor.py:86)
            ek[k] = s_a[(colpass_ch * 8 + k) * DCT_RS + colpass_c] * s_q[colpass_ch * 64 + k * 8 + j];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k) s = __fma_rn(c_dct[k * 8 + i], ek[k], s);
            s_b[(colpass_ch * 8 + i) * DCT_RS + colpass_c] = s;
        }
    }
    __syncthreads();
    // ---- 6. row pass  P = T' . C ; truncating uint8 store ; +128 -----------------------------
    if (row_on) {
        double tk[8];
        const int base = (rp_ch * 8 + rp_i) * DCT_RS + rp_blk * 8;
#pragma unroll
        for (int k = 0; k < 8; ++k) tk[k] = s_b[base + k];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k) s = __fma_rn(tk[k], c_dct[k * 8 + j], s);
            // float64 -> uint8 store (DCTcompressor.py:81,88): truncate toward zero, low 8 bits
            const uint8_t p8 = (uint8_t)(long long)s;
            s_ycc[(rp_i * DCT_TILE_W + rp_blk * 8 + j) * 3 + rp_ch] = (uint8_t)(p8 + 128);
        }
    }
    __syncthreads();
    // ---- 7. YCrCb -> BGR, + pred (wrap), coalesced store -------------------------------------
    uint8_t *recon = a.recon + (size_t)p * npix * 3;
    for (int k = tid; k < 8 * DCT_TILE_W; k += DCT_THREADS) {
        const int r = k / DCT_TILE_W, c = k - r * DCT_TILE_W;
        if (c >= tw) continue;
        int B, G, R;
        ycrcb2bgr(s_ycc[3 * k], s_ycc[3 * k + 1], s_ycc[3 * k + 2], B, G, R);
        uint8_t *op = recon + ((size_t)(y0 + r) * a.W + x0 + c) * 3;
        op[0] = (uint8_t)(B + s_pred[3 * k]);
        op[1] = (uint8_t)(G + s_pred[3 * k + 1]);
        op[2] = (uint8_t)(R + s_pred[3 * k + 2]);
    }
}

// MotionProcessor.reconstruct_from_motion_vectors (motion.py:42-69) as its own kernel, for the
// drop-in method; the fused path above never materialises pred.
__global__ void mc_kernel(const uint8_t *__restrict__ ref, const int16_t *__restrict__ mv, int H,
                          int W, int bs, int nbx, int nby, uint8_t *__restrict__ pred) {
    const size_t npix = (size_t)H * W;
    for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < npix;
         k += (size_t)gridDim.x * blockDim.x) {
        const int y = (int)(k / W), x = (int)(k - (size_t)y * W);
        const int mbx = x / bs, mby = y / bs;
        uint8_t b = 0, g = 0, r = 0;
        if (mbx < nbx && mby < nby) {
            const int16_t *m = mv + 2 * (mby * nbx + mbx);
            const uint8_t *rp = ref + ((size_t)(y + m[1]) * W + (x + m[0])) * 3;
            b = rp[0]; g = rp[1]; r = rp[2];
        }
        pred[3 * k] = b; pred[3 * k + 1] = g; pred[3 * k + 2] = r;
    }
}

// get_residuals (motion.py:38-40) / _fully_reconstruct (decoder.py:57): byte-wise wrap
template <int ADD>
__global__ void wrap_kernel(const uint8_t *__restrict__ a, const uint8_t *__restrict__ b, size_t n,
                            uint8_t *__restrict__ out) {
    for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n;
         k += (size_t)gridDim.x * blockDim.x)
        out[k] = ADD ? (uint8_t)(a[k] + b[k]) : (uint8_t)(a[k] - b[k]);
}

}  // namespace vcs
