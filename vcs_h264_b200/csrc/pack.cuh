// pack.cuh -- compaction of the quantised int8 indices before they leave the GPU (the packed wire form
// of the reference's Frame.r planes; the reference itself keeps dense float64 planes, frame.py:1-8, and has no
// bitstream -- proposal section 5.1 promised a zero-run stage that was never written).
//
// The reference's wrapped residual (motion.py:39) leaves the indices dense (43 % non-zero at QF 50 on the bench clip),
// so the format is the cheapest one that is still exact: per 8x8 block a 64-bit occupancy bitmap (bit 8*i+j = row i,
// column j) followed -- in one byte stream for the whole clip -- by the non-zero int8 values of the block in bit order.
// Blocks are ordered (P-frame, channel Y/Cr/Cb, block row, block column); `row_count[p][ch][by]` holds the number of
// values of one block row (W/8 blocks), so any block row can be located by a prefix sum and decoded independently.
//   dense : 3*H*W bytes per P-frame           packed : 3*H*W/8 (bitmaps) + nnz (values) + 12*H/8 (row counts)
//
// Kernels (HBM-bound streaming passes, one warp per block row, lanes = blocks, every load coalesced):
//   pack_count_kernel   dense int8 planes -> bitmaps + row counts (stand-alone packing; in the encoder the DCT stage
//                       emits both while it still holds the indices in registers, dct_stage.cuh)
//   pack_scan_kernel    exclusive prefix of the row counts of one segment (single CTA) + running clip total
//   pack_write_kernel   dense planes + bitmaps + row offsets -> value stream
//   unpack_kernel       the exact inverse (decoder side)
#pragma once
#include "common.cuh"

namespace vcs {

// bitmap of the 8x8 block whose top-left byte is p (row pitch W); p is 8-byte aligned
__device__ __forceinline__ unsigned long long block_bitmap(const int8_t *p, int W, uint2 rows[8]) {
    unsigned long long bm = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        rows[i] = __ldg(reinterpret_cast<const uint2 *>(p + (size_t)i * W));
        bm |= (unsigned long long)(nz_nibble(rows[i].x) | (nz_nibble(rows[i].y) << 4)) << (8 * i);
    }
    return bm;
}

constexpr int PACK_WARPS = 8;

// coef: [nrows_total/ (H/8)...] dense planes viewed as nrows = nP*3*(H/8) block rows of 8 x W bytes each
__global__ void __launch_bounds__(32 * PACK_WARPS)
pack_count_kernel(const int8_t *__restrict__ coef, int W, int nrows, unsigned long long *__restrict__ bitmap,
                  uint32_t *__restrict__ row_count) {
    const int lane = threadIdx.x & 31, nbx = W / 8;
    for (int row = blockIdx.x * PACK_WARPS + (threadIdx.x >> 5); row < nrows; row += gridDim.x * PACK_WARPS) {
        const int8_t *base = coef + (size_t)row * 8 * W;
        uint32_t cnt = 0;
        for (int bx = lane; bx < nbx; bx += 32) {
            uint2 rows[8];
            const unsigned long long bm = block_bitmap(base + 8 * bx, W, rows);
            bitmap[(size_t)row * nbx + bx] = bm;
            cnt += __popcll(bm);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (lane == 0) row_count[row] = cnt;
    }
}

// row_off[k] = *total + sum_{j<k} row_count[j] for the n rows of a segment; *total += sum.  One CTA.
__global__ void __launch_bounds__(1024)
pack_scan_kernel(const uint32_t *__restrict__ row_count, int n, unsigned long long *__restrict__ row_off,
                 unsigned long long *total, unsigned long long *seg_end) {
    __shared__ unsigned long long warp_sum[32];
    __shared__ unsigned long long carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = *total;
    __syncthreads();
    for (int k0 = 0; k0 < n; k0 += 1024) {
        const int k = k0 + threadIdx.x;
        const unsigned long long v = k < n ? row_count[k] : 0;
        unsigned long long s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = shfl_up_u64(s, o);
            if (lane >= o) s += t;
        }
        if (lane == 31) warp_sum[warp] = s;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = warp_sum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long t = shfl_up_u64(w, o);
                if (lane >= o) w += t;
            }
            warp_sum[lane] = w;          // inclusive over warps
        }
        __syncthreads();
        const unsigned long long before = carry + (warp ? warp_sum[warp - 1] : 0) + (s - v);
        if (k < n) row_off[k] = before;
        __syncthreads();
        if (threadIdx.x == 0) carry += warp_sum[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) { *total = carry; if (seg_end) *seg_end = carry; }
}

// One warp per (block row, batch of 32 blocks).  The batch's values start at row_off[row] + the popcount of the row's
// earlier bitmaps (read back, <= 7 coalesced loads per lane); lanes = blocks, a warp-wide prefix places each block.
// A lane's values are a byte string at an arbitrary byte offset: written straight to global memory every store
// instruction would touch 32 different sectors, so the warp first compacts its 32 blocks into shared memory and then
// streams the contiguous run (<= 2 KB) out with one sector per store instruction.
__global__ void __launch_bounds__(32 * PACK_WARPS)
pack_write_kernel(const int8_t *__restrict__ coef, int W, int nrows, const unsigned long long *__restrict__ bitmap,
                  const unsigned long long *__restrict__ row_off, int8_t *__restrict__ values) {
    __shared__ __align__(16) uint8_t stage[PACK_WARPS][32 * 64 + 16];
    const int lane = threadIdx.x & 31, nbx = W / 8, nbatch = (nbx + 31) / 32;
    uint8_t *sw = stage[threadIdx.x >> 5];
    const long long nitems = (long long)nrows * nbatch;
    for (long long item = (long long)blockIdx.x * PACK_WARPS + (threadIdx.x >> 5); item < nitems;
         item += (long long)gridDim.x * PACK_WARPS) {
        const int row = (int)(item / nbatch), bx0 = (int)(item - (long long)row * nbatch) * 32;
        const unsigned long long *bmrow = bitmap + (size_t)row * nbx;
        uint32_t before = 0;
        for (int b = lane; b < bx0; b += 32) before += __popcll(__ldg(bmrow + b));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
        const int bx = bx0 + lane;
        const unsigned long long bm = bx < nbx ? __ldg(bmrow + bx) : 0ull;
        const uint32_t cnt = __popcll(bm);
        uint32_t incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        if (cnt) {
            const int8_t *blk = coef + (size_t)row * 8 * W + 8 * bx;
            uint8_t *dst = sw + (incl - cnt);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if ((bm >> (8 * i)) & 0xffull) {
                    const uint2 r = __ldg(reinterpret_cast<const uint2 *>(blk + (size_t)i * W));
                    const uint32_t w[2] = {r.x, r.y};
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint32_t b = (w[j >> 2] >> (8 * (j & 3))) & 0xffu;
                        if (b) *dst++ = (uint8_t)b;
                    }
                }
            }
        }
        __syncwarp();
        int8_t *g = values + row_off[row] + before;
        for (uint32_t k = lane; k < total; k += 32) g[k] = (int8_t)sw[k];
        __syncwarp();
    }
}

// inverse: bitmaps + row offsets + value stream -> dense int8 planes (every byte of the planes is written)
__global__ void __launch_bounds__(32 * PACK_WARPS)
unpack_kernel(const unsigned long long *__restrict__ bitmap, const unsigned long long *__restrict__ row_off,
              const int8_t *__restrict__ values, unsigned long long nvalues, int W, int nrows,
              int8_t *__restrict__ coef, int *err) {
    const int lane = threadIdx.x & 31, nbx = W / 8;
    for (int row = blockIdx.x * PACK_WARPS + (threadIdx.x >> 5); row < nrows; row += gridDim.x * PACK_WARPS) {
        int8_t *base = coef + (size_t)row * 8 * W;
        unsigned long long off = row_off[row];
        for (int bx0 = 0; bx0 < nbx; bx0 += 32) {
            const int bx = bx0 + lane;
            const unsigned long long bm = bx < nbx ? __ldg(bitmap + (size_t)row * nbx + bx) : 0ull;
            const uint32_t cnt = __popcll(bm);
            uint32_t incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            unsigned long long src = off + (incl - cnt);
            if (bx < nbx) {
                const bool ok = src + cnt <= nvalues;       // a damaged stream is never read past its end
                if (!ok && err) *(volatile int *)err = 2;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    uint32_t w[2] = {0, 0};
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (ok && ((bm >> (8 * i + j)) & 1)) w[j >> 2] |= (uint32_t)(uint8_t)__ldg(values + src++) << (8 * (j & 3));
                    *reinterpret_cast<uint2 *>(base + (size_t)i * W + 8 * bx) = make_uint2(w[0], w[1]);
                }
            }
            off += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
}

}  // namespace vcs
