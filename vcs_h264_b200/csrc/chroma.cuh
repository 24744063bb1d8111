// chroma.cuh -- 4:2:0 chroma subsampling of ChromaSubsampling/chroma.py.
//
//   chroma.py:9      imgYYC = cv2.cvtColor(img, COLOR_BGR2YCR_CB)            (fixed point, common.cuh)
//   chroma.py:16-17  cr/cb  = cv2.boxFilter(plane, ddepth=-1, ksize=(2,2))   anchor (1,1), BORDER_REFLECT_101,
//                    uint8 result = ceil(sum/4) (OpenCV 4.13, pinned in tests/golden/golden_chroma_meta.json)
//   chroma.py:20-21  samples = filtered[::2, ::2]                            ceil(H/2) x ceil(W/2)
//   chroma.py:27-41  the demo's float reconstruction (NumPy-2 uint8 scalar wrap of `Cr - 128`, float64,
//                    clamp, truncating store)
// Both kernels are HBM-bound streaming passes: 3 B/px in, 1.5 B/px out (subsample); 1.5 in, 3 out (rebuild).
#pragma once
#include "common.cuh"

namespace vcs {

// One thread per chroma sample (i, j): it owns image rows {2i-1, 2i} x columns {2j-1, 2j} -- the box
// filter's window -- converts those pixels once, writes their Y, and the two chroma means.  The grid has one
// extra row / column of threads so that the last image row / column (outside every window when the size is
// even) still gets its Y written.
__global__ void chroma420_kernel(const uint8_t *__restrict__ bgr, int H, int W, uint8_t *__restrict__ Yp,
                                 uint8_t *__restrict__ crS, uint8_t *__restrict__ cbS) {
    const int h2 = (H + 1) / 2, w2 = (W + 1) / 2;
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i > H / 2 || j > W / 2) return;
    const bool sample = i < h2 && j < w2;
    int scr = 0, scb = 0;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            int y = 2 * i - 1 + a, x = 2 * j - 1 + b;
            const bool own = y >= 0 && y < H && x >= 0 && x < W;     // this thread writes Y of its own pixels
            if (y < 0) y = H > 1 ? 1 : 0;                            // BORDER_REFLECT_101: -1 -> 1
            if (x < 0) x = W > 1 ? 1 : 0;
            if (y >= H || x >= W) continue;                          // only the extra row / column: no sample there
            const uint8_t *p = bgr + ((size_t)y * W + x) * 3;
            int Yv, Cr, Cb;
            bgr2ycrcb(__ldg(p), __ldg(p + 1), __ldg(p + 2), Yv, Cr, Cb);
            if (own) Yp[(size_t)y * W + x] = (uint8_t)Yv;
            scr += Cr;
            scb += Cb;
        }
    }
    if (sample) {
        crS[(size_t)i * w2 + j] = (uint8_t)((scr + 3) >> 2);
        cbS[(size_t)i * w2 + j] = (uint8_t)((scb + 3) >> 2);
    }
}

// chroma.py:27-41, one thread per pixel.  -fmad=false keeps every product and sum separately rounded.
__global__ void chroma420_to_bgr_kernel(const uint8_t *__restrict__ Yp, const uint8_t *__restrict__ crS,
                                        const uint8_t *__restrict__ cbS, int H, int W, uint8_t *__restrict__ bgr) {
    const int w2 = (W + 1) / 2;
    const size_t npix = (size_t)H * W;
    for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < npix; k += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(k / W), j = (int)(k - (size_t)i * W);
        const double y = (double)Yp[k];
        const double crw = (double)(uint8_t)(crS[(size_t)(i >> 1) * w2 + (j >> 1)] - 128);   // uint8 wrap (NumPy 2)
        const double cbw = (double)(uint8_t)(cbS[(size_t)(i >> 1) * w2 + (j >> 1)] - 128);
        double r = y + 1.4022 * crw;
        double g = (y - 0.34414 * cbw) - 0.71414 * crw;
        double b = y + 1.77200 * cbw;
        r = fmin(fmax(r, 0.0), 255.0);
        g = fmin(fmax(g, 0.0), 255.0);
        b = fmin(fmax(b, 0.0), 255.0);
        bgr[3 * k] = (uint8_t)(int)b;
        bgr[3 * k + 1] = (uint8_t)(int)g;
        bgr[3 * k + 2] = (uint8_t)(int)r;
    }
}

}  // namespace vcs
