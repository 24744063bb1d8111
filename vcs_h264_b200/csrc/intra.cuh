// intra.cuh -- intra mode decision (BASELINE.json config 4):
// luma4x4 / luma16x16 / chroma8x8 of IntraframeCompression/intraframe.py:24-317 with the 15
// predictors of IntraframeCompression/intramodes.py:7-179.
//
// Neighbours are ORIGINAL pixels (intraframe.py:58-77), so luma blocks are independent: one thread
// per 4x4 block, one warp per 16x16 block.  chroma8x8 takes Cb's upper neighbour from the RESIDUAL
// row above (Cbres, intraframe.py:266): blocks of one column form a chain, columns are independent,
// so one warp walks down one block column.  HBM-bound integer work; all reference type quirks are
// reproduced (see oracle/vcs_oracle.c "Intra mode decision" for the list, checked against the
// unmodified reference on the whole 736x736 test image).
#pragma once
#include "common.cuh"

namespace vcs {

__device__ __forceinline__ int fdiv_i(int a, int b) {  // Python floor division, b > 0
    int q = a / b;
    return (a % b != 0 && a < 0) ? q - 1 : q;
}

// predictor m of a 4x4 block into P[16] (row-major)
__device__ __forceinline__ void pred4x4(int m, int ul, const int *u, const int *ur, const int *l, bool u_u8,
                                        bool ur_u8, bool l_u8, int *P) {
#define Q4(x) ((x) >> 2)   // all operands here are >= 0
#define H2(x) ((x) >> 1)
    const int t3ur = ur_u8 ? ((3 * ur[3]) & 255) >> 2 : (3 * ur[3]) >> 2;   // uint8 wrap of 3*x (intramodes.py:42)
    const int t3l = l_u8 ? ((3 * l[3]) & 255) >> 2 : (3 * l[3]) >> 2;       // intramodes.py:135
    switch (m) {
    case 0:
#pragma unroll
        for (int k = 0; k < 16; ++k) P[k] = u[k & 3];
        break;
    case 1:
#pragma unroll
        for (int k = 0; k < 16; ++k) P[k] = l[k >> 2];
        break;
    case 2: {
        int s = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) s += (u_u8 && l_u8) ? ((u[k] + l[k]) & 255) : (u[k] + l[k]);   // intramodes.py:21
        const int avg = s >> 3;
#pragma unroll
        for (int k = 0; k < 16; ++k) P[k] = avg;
        break;
    }
    case 3:
        P[0] = Q4(u[0]) + H2(u[1]) + Q4(u[2]);
        P[1] = Q4(u[1]) + H2(u[2]) + Q4(u[3]); P[4] = P[1];
        P[2] = Q4(u[2]) + H2(u[3]) + Q4(ur[0]); P[5] = P[2]; P[8] = P[2];
        P[3] = Q4(u[3]) + H2(ur[0]) + Q4(ur[1]); P[6] = P[3]; P[9] = P[3]; P[12] = P[3];
        P[7] = Q4(ur[0]) + H2(ur[1]) + Q4(ur[2]); P[10] = P[7]; P[13] = P[7];
        P[11] = Q4(ur[1]) + H2(ur[2]) + Q4(ur[3]); P[14] = P[11];
        P[15] = Q4(ur[2]) + t3ur;
        break;
    case 4:
        P[3] = Q4(u[1]) + H2(u[2]) + Q4(u[3]);
        P[2] = Q4(u[0]) + H2(u[1]) + Q4(u[2]); P[7] = P[2];
        P[1] = Q4(ul) + H2(u[0]) + Q4(u[1]); P[6] = P[1]; P[11] = P[1];
        P[0] = Q4(ul) + H2(u[0]) + Q4(l[0]); P[5] = P[0]; P[10] = P[0]; P[15] = P[0];
        P[4] = Q4(u[0]) + H2(l[0]) + Q4(l[1]); P[9] = P[4]; P[14] = P[4];
        P[8] = Q4(l[0]) + H2(l[1]) + Q4(l[2]); P[13] = P[8];
        P[12] = Q4(l[1]) + H2(l[2]) + Q4(l[3]);
        break;
    case 5:
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
    case 6:
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
    case 7:
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
    default:
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

// luma4x4 (intraframe.py:24-151): one thread per 4x4 block.
__global__ void intra_luma4x4_kernel(const uint8_t *__restrict__ Y, int H, int W, int32_t *__restrict__ res,
                                     int32_t *__restrict__ pred, uint8_t *__restrict__ modes) {
    const int mc = W / 4, mr = H / 4;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= mc * mr) return;
    const int im = b / mc, jm = b - im * mc, i = im * 4, j = jm * 4;
    bool s_ul = false, s_u = false, s_ur = false, s_l = false;   // availability (intraframe.py:38-55)
    if (im == 0 && jm == 0) { }
    else if (im == 0) s_l = true;
    else if (jm == 0) { s_u = true; s_ur = mc > 1; }
    else if (jm + 1 == mc) { s_ul = true; s_u = true; s_l = true; }
    else { s_ul = s_u = s_ur = s_l = true; }
    const int ul = s_ul ? Y[(size_t)(i - 1) * W + j - 1] : 128;
    int u[4], ur[4], l[4], y[16];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        u[k] = s_u ? Y[(size_t)(i - 1) * W + j + k] : 128;
        l[k] = s_l ? Y[(size_t)(i + k) * W + j - 1] : 128;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) ur[k] = s_ur ? Y[(size_t)(i - 1) * W + j + 4 + k] : (s_u ? u[3] : 128);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const uint32_t w = *reinterpret_cast<const uint32_t *>(Y + (size_t)(i + a) * W + j);
#pragma unroll
        for (int c = 0; c < 4; ++c) y[a * 4 + c] = (w >> (8 * c)) & 0xff;
    }
    int best = 16 * 255, bmode = 0, bp[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) bp[k] = 0;
#pragma unroll 1
    for (int m = 0; m < 9; ++m) {
        int P[16];
        pred4x4(m, ul, u, ur, l, s_u, s_ur, s_l, P);
        int d = 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) d += abs(P[k] - y[k]);
        if (d < best) {   // first strict minimum (intraframe.py:84-144)
            best = d; bmode = m;
#pragma unroll
            for (int k = 0; k < 16; ++k) bp[k] = P[k];
        }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const size_t o = (size_t)(i + a) * W + j;
        *reinterpret_cast<int4 *>(pred + o) = make_int4(bp[a * 4], bp[a * 4 + 1], bp[a * 4 + 2], bp[a * 4 + 3]);
        *reinterpret_cast<int4 *>(res + o) = make_int4(y[a * 4] - bp[a * 4], y[a * 4 + 1] - bp[a * 4 + 1],
                                                       y[a * 4 + 2] - bp[a * 4 + 2], y[a * 4 + 3] - bp[a * 4 + 3]);
    }
    modes[b] = (uint8_t)bmode;
}

// luma16x16 (intraframe.py:153-225): one warp per block, lane = (row, half row of 8 pixels).
__global__ void intra_luma16x16_kernel(const uint8_t *__restrict__ Y, int H, int W, int32_t *__restrict__ res,
                                       int32_t *__restrict__ pred, uint8_t *__restrict__ modes) {
    const int mc = W / 16, mr = H / 16;
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (b >= mc * mr) return;
    const int im = b / mc, jm = b - im * mc, i = im * 16, j = jm * 16;
    const bool s_u = im > 0, s_l = jm > 0;
    const int r = lane >> 1, c0 = (lane & 1) * 8;
    int y[8], u[8];
    const uint2 yw = *reinterpret_cast<const uint2 *>(Y + (size_t)(i + r) * W + j + c0);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        y[k] = ((k < 4 ? yw.x : yw.y) >> (8 * (k & 3))) & 0xff;
        u[k] = s_u ? Y[(size_t)(i - 1) * W + j + c0 + k] : 128;
    }
    const int lrow = s_l ? Y[(size_t)(i + r) * W + j - 1] : 128;
    // dc = (sum(u) + sum(l)) // 32 (intramodes.py:157-161): u over 16 columns (lanes 0,1), l over 16 rows
    int su = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) su += u[k];
    int part = (lane < 2 ? su : 0) + ((lane & 1) == 0 ? lrow : 0);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    const int dc = part >> 5;
    int d0 = 0, d1 = 0, d2 = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { d0 += abs(u[k] - y[k]); d1 += abs(lrow - y[k]); d2 += abs(dc - y[k]); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        d0 += __shfl_xor_sync(0xffffffffu, d0, o);
        d1 += __shfl_xor_sync(0xffffffffu, d1, o);
        d2 += __shfl_xor_sync(0xffffffffu, d2, o);
    }
    int best = 16 * 16 * 255, bmode = 0; bool any = false;
    if (d0 < best) { best = d0; bmode = 0; any = true; }
    if (d1 < best) { best = d1; bmode = 1; any = true; }
    if (d2 < best) { best = d2; bmode = 2; any = true; }
    const size_t o = (size_t)(i + r) * W + j + c0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int pv = !any ? 0 : (bmode == 0 ? u[k] : (bmode == 1 ? lrow : dc));
        pred[o + k] = pv;
        res[o + k] = y[k] - pv;
    }
    if (lane == 0) modes[b] = (uint8_t)bmode;
}

// chroma8x8 (intraframe.py:228-317): one warp per block COLUMN, walking down; lane = (row, 2 pixels).
__global__ void intra_chroma8x8_kernel(const uint8_t *__restrict__ Cr, const uint8_t *__restrict__ Cb, int H, int W,
                                       int32_t *__restrict__ crres, int32_t *__restrict__ crpred,
                                       int32_t *__restrict__ cbres, int32_t *__restrict__ cbpred,
                                       uint8_t *__restrict__ modes) {
    const int mc = W / 8, mr = H / 8;
    const int jm = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (jm >= mc) return;
    const int j = jm * 8, r = lane >> 2, c0 = (lane & 3) * 2;
    const bool s_l = jm > 0;
    int prev_b0 = 0, prev_b1 = 0;   // this lane's Cb residuals of the block above (rows r, cols c0, c0+1)
    // Only the Cb up-neighbours chain from block to block; every load is independent of the chain, so the pixels of
    // block im + 1 are fetched while block im is decided (the walk is latency bound: one warp per column).
    struct Px { int yr0, yr1, yb0, yb1, ur0, ur1, lr, lb; };
    auto fetch = [&](int im) {
        Px q;
        const int i = im * 8;
        const size_t o = (size_t)(i + r) * W + j + c0;
        q.yr0 = Cr[o]; q.yr1 = Cr[o + 1]; q.yb0 = Cb[o]; q.yb1 = Cb[o + 1];
        q.ur0 = im > 0 ? Cr[(size_t)(i - 1) * W + j + c0] : 128;
        q.ur1 = im > 0 ? Cr[(size_t)(i - 1) * W + j + c0 + 1] : 128;
        q.lr = s_l ? Cr[(size_t)(i + r) * W + j - 1] : 128;
        q.lb = s_l ? Cb[(size_t)(i + r) * W + j - 1] : 128;
        return q;
    };
    Px nxt = fetch(0);
    for (int im = 0; im < mr; ++im) {
        const int i = im * 8;
        const bool s_u = im > 0;
        const size_t o = (size_t)(i + r) * W + j + c0;
        const Px cur = nxt;
        if (im + 1 < mr) nxt = fetch(im + 1);
        const int yr0 = cur.yr0, yr1 = cur.yr1, yb0 = cur.yb0, yb1 = cur.yb1;
        // up neighbours of this lane's two columns: Cr from the image, Cb from the RESIDUAL row above (:266)
        const int ur0 = cur.ur0, ur1 = cur.ur1;
        const int src = 28 + (lane & 3);   // lanes holding row 7 of the block above
        const int pb0 = __shfl_sync(0xffffffffu, prev_b0, src), pb1 = __shfl_sync(0xffffffffu, prev_b1, src);
        const int ub0 = s_u ? pb0 : 128, ub1 = s_u ? pb1 : 128;
        const int lr = cur.lr, lb = cur.lb;
        // dc = (sum(u) + sum(l)) // 16, floor division (Cb sums can be negative)
        int sr = (r == 0 ? ur0 + ur1 : 0) + ((lane & 3) == 0 ? lr : 0);
        int sb = (r == 0 ? ub0 + ub1 : 0) + ((lane & 3) == 0 ? lb : 0);
#pragma unroll
        for (int k = 16; k > 0; k >>= 1) {
            sr += __shfl_xor_sync(0xffffffffu, sr, k);
            sb += __shfl_xor_sync(0xffffffffu, sb, k);
        }
        const int dcr = fdiv_i(sr, 16), dcb = fdiv_i(sb, 16);
        // |.| <= 510 per term (Cb's up neighbour is a residual in [-255, 255]), 128 terms per mode: int32 is ample
        int d0 = abs(ur0 - yr0) + abs(ur1 - yr1) + abs(ub0 - yb0) + abs(ub1 - yb1);
        int d1 = abs(lr - yr0) + abs(lr - yr1) + abs(lb - yb0) + abs(lb - yb1);
        int d2 = abs(dcr - yr0) + abs(dcr - yr1) + abs(dcb - yb0) + abs(dcb - yb1);
#pragma unroll
        for (int k = 16; k > 0; k >>= 1) {
            d0 += __shfl_xor_sync(0xffffffffu, d0, k);
            d1 += __shfl_xor_sync(0xffffffffu, d1, k);
            d2 += __shfl_xor_sync(0xffffffffu, d2, k);
        }
        int best = 2 * 8 * 8 * 255, bmode = 0; bool any = false;
        if (d0 < best) { best = d0; bmode = 0; any = true; }
        if (d1 < best) { best = d1; bmode = 1; any = true; }
        if (d2 < best) { best = d2; bmode = 2; any = true; }
        const int pr0 = !any ? 0 : (bmode == 0 ? ur0 : (bmode == 1 ? lr : dcr)), pr1 = !any ? 0 : (bmode == 0 ? ur1 : (bmode == 1 ? lr : dcr));
        const int q0 = !any ? 0 : (bmode == 0 ? ub0 : (bmode == 1 ? lb : dcb)), q1 = !any ? 0 : (bmode == 0 ? ub1 : (bmode == 1 ? lb : dcb));
        crpred[o] = pr0; crpred[o + 1] = pr1; crres[o] = yr0 - pr0; crres[o + 1] = yr1 - pr1;
        cbpred[o] = q0; cbpred[o + 1] = q1;
        prev_b0 = yb0 - q0; prev_b1 = yb1 - q1;
        cbres[o] = prev_b0; cbres[o + 1] = prev_b1;
        if (lane == 0) modes[im * mc + jm] = (uint8_t)bmode;
    }
}

}  // namespace vcs
