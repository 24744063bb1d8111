// me_generic.cuh -- block-matching search, one CTA per macroblock, any bs / step / interval.
//
// This is the kernel behind the reference's own settings (bs 8 -> step 3, <= 64 candidates
// per macroblock; InterframeCompression/motion.py:100-154).  It favours generality over
// throughput: candidates are spread over the CTA's threads, ref bytes come through L1, the
// macroblock itself is staged in shared memory as packed words.  The step-1 full search
// of BASELINE.json's configs 2/3/5 runs on me_tiled.cuh instead.
#pragma once
#include "common.cuh"

namespace vcs {

constexpr int ME_GENERIC_THREADS = 128;

// Word w of a row of `rowbytes` bytes starting at `row` (any alignment); bytes past the end
// of the row read as 0 so both operands of the cost agree there.
__device__ __forceinline__ uint32_t load_row_word(const uint8_t *row, int w, int rowbytes) {
    uint32_t v = 0;
    const int b0 = 4 * w;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (b0 + k < rowbytes) v |= (uint32_t)__ldg(row + b0 + k) << (8 * k);
    return v;
}

template <int METRIC>
__global__ void __launch_bounds__(ME_GENERIC_THREADS)
me_generic_kernel(FrameAddr fa, MeGeom g, int16_t *__restrict__ mv_out,
                  uint32_t *__restrict__ cost_out, uint8_t *__restrict__ flag_out) {
    extern __shared__ uint32_t s_cur[];  // [bs][rw] packed macroblock rows
    __shared__ unsigned long long s_red[ME_GENERIC_THREADS / 32];
    __shared__ int s_static;

    const int mb = blockIdx.x, p = blockIdx.y;
    const int N = g.nbx * g.nby;
    const int x = (mb % g.nbx) * g.bs, y = (mb / g.nbx) * g.bs;
    const int pitch = 3 * g.W, rowbytes = 3 * g.bs, rw = (rowbytes + 3) / 4;
    const uint8_t *cur = cur_frame(fa, p), *ref = ref_frame(fa, p);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t o = (size_t)p * N + mb;

    for (int k = tid; k < g.bs * rw; k += ME_GENERIC_THREADS) {
        int v = k / rw, w = k - v * rw;
        s_cur[k] = load_row_word(cur + (size_t)(y + v) * pitch + 3 * x, w, rowbytes);
    }
    __syncthreads();

    // ---- static test (motion.py:109-116): S = sum max(ref - cur, 0) at the MB's own position.
    // max(d,0) = (|d| + d) / 2, so S = (SAD + sum(ref) - sum(cur)) / 2 exactly.
    if (g.static_thr >= 0) {
        uint32_t sad = 0, sr = 0, sc = 0;
        for (int k = tid; k < g.bs * rw; k += ME_GENERIC_THREADS) {
            int v = k / rw, w = k - v * rw;
            uint32_t r = load_row_word(ref + (size_t)(y + v) * pitch + 3 * x, w, rowbytes);
            uint32_t c = s_cur[k];
            sad = sad4_acc(r, c, sad);
            sr = bytesum_acc(r, sr);
            sc = bytesum_acc(c, sc);
        }
        long long part = (long long)sad + (long long)sr - (long long)sc;  // 2 * partial S
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) {
            long long q = (long long)shfl_xor_u64((unsigned long long)part, m);
            part += q;
        }
        if (lane == 0) s_red[warp] = (unsigned long long)part;
        __syncthreads();
        if (tid == 0) {
            long long S2 = 0;
            for (int k = 0; k < ME_GENERIC_THREADS / 32; ++k) S2 += (long long)s_red[k];
            long long S = S2 / 2;
            s_static = S <= g.static_thr;
            if (s_static) {
                mv_out[2 * o] = 0;
                mv_out[2 * o + 1] = 0;
                if (cost_out) cost_out[o] = (uint32_t)S;
                if (flag_out) flag_out[o] = 1;
            }
        }
        __syncthreads();
        if (s_static) return;
    }

    // ---- candidate scan (motion.py:117-152)
    const int i0 = max(y + g.lo, 0), j0 = max(x + g.lo, 0);
    const int i1 = min(y + g.hi, g.H - g.bs - g.slack), j1 = min(x + g.hi, g.W - g.bs - g.slack);
    const int ny = i1 >= i0 ? (i1 - i0) / g.step + 1 : 0;
    const int nx = j1 >= j0 ? (j1 - j0) / g.step + 1 : 0;
    const int total = ny * nx;
    unsigned long long best = ~0ull;
    for (int c = tid; c < total; c += ME_GENERIC_THREADS) {
        int iy = c / nx, ix = c - iy * nx;
        const uint8_t *rp = ref + (size_t)(i0 + iy * g.step) * pitch + 3 * (j0 + ix * g.step);
        uint32_t acc = 0;
        for (int v = 0; v < g.bs; ++v) {
            const uint8_t *row = rp + (size_t)v * pitch;
            for (int w = 0; w < rw; ++w)
                acc = cost4_acc<METRIC>(load_row_word(row, w, rowbytes), s_cur[v * rw + w], acc);
        }
        // scan index c grows in (row outer, col inner) order, so min over (cost, c) is the
        // reference's "first strict minimum" (motion.py:149)
        unsigned long long key = ((unsigned long long)acc << 32) | (uint32_t)c;
        best = key < best ? key : best;
    }
    best = warp_min_u64(best);
    __syncthreads();
    if (lane == 0) s_red[warp] = best;
    __syncthreads();
    if (tid == 0) {
        for (int k = 1; k < ME_GENERIC_THREADS / 32; ++k) best = s_red[k] < best ? s_red[k] : best;
        if (total == 0) {
            mv_out[2 * o] = (int16_t)(-x);  // best_coord stays [0,0] (motion.py:102)
            mv_out[2 * o + 1] = (int16_t)(-y);
            if (cost_out) cost_out[o] = 0xFFFFFFFFu;
            if (flag_out) flag_out[o] = 2;
        } else {
            int c = (int)(uint32_t)best;
            int iy = c / nx, ix = c - iy * nx;
            mv_out[2 * o] = (int16_t)(j0 + ix * g.step - x);
            mv_out[2 * o + 1] = (int16_t)(i0 + iy * g.step - y);
            if (cost_out) cost_out[o] = (uint32_t)(best >> 32);
            if (flag_out) flag_out[o] = 0;
        }
    }
}

}  // namespace vcs
