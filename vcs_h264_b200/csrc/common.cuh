// common.cuh -- shared device helpers of the VCS-h264 B200 hot path (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vcs {

// How P-frame ordinal p maps to its (cur, ref) frames.  The reference predicts every P-frame
// from the ORIGINAL I-frame of its GOP (InterframeCompression/encoder.py:42,51-52), so all
// P-frames of a clip are independent and one launch covers them all.
//   cur(p) = cur_base + (p / ppg) * cur_gop_stride + (p % ppg) * cur_frame_stride
//   ref(p) = ref_base + (p / ppg) * ref_gop_stride
struct FrameAddr {
    const uint8_t *cur_base;
    const uint8_t *ref_base;
    long long cur_gop_stride, cur_frame_stride, ref_gop_stride;
    int ppg;        // P-frames per GOP
    int p_off = 0;  // P-ordinal of launch-local p = 0 relative to cur_base / ref_base (launches that start inside a GOP)
};

__device__ __forceinline__ const uint8_t *cur_frame(const FrameAddr &a, int p) {
    p += a.p_off;
    return a.cur_base + (long long)(p / a.ppg) * a.cur_gop_stride +
           (long long)(p % a.ppg) * a.cur_frame_stride;
}
__device__ __forceinline__ const uint8_t *ref_frame(const FrameAddr &a, int p) {
    return a.ref_base + (long long)((p + a.p_off) / a.ppg) * a.ref_gop_stride;
}

// Search geometry shared by both ME kernels (include/vcs_b200.h: vcs_me_params).
struct MeGeom {
    int H, W, bs, lo, hi, step, slack, nbx, nby;
    long long static_thr;
};

// ---- packed-byte cost primitives ----------------------------------------------------------
// One 32-bit word = 4 bytes of a BGR-interleaved row.

// acc + sum_i |a_i - b_i| : a single VABSDIFF4.U8.ACC on sm_100a.
__device__ __forceinline__ uint32_t sad4_acc(uint32_t a, uint32_t b, uint32_t acc) {
    uint32_t d;
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(acc));
    return d;
}

// acc + sum_i ((a_i - b_i) mod 256): the reference's cost (motion.py:146 subtracts uint8
// arrays, np.abs is then a no-op).  Per-byte borrow-isolated subtract, then IDP.4A sums the
// bytes.  (a|H) - (b&~H) never borrows across bytes; bit 7 is fixed up by the xor.
__device__ __forceinline__ uint32_t wrap4_acc(uint32_t a, uint32_t b, uint32_t acc) {
    const uint32_t Hm = 0x80808080u;
    uint32_t t = (a | Hm) - (b & ~Hm);
    uint32_t z = t ^ ((a ^ ~b) & Hm);
    return __dp4a(z, 0x01010101u, acc);
}

// acc + sum of the 4 bytes
__device__ __forceinline__ uint32_t bytesum_acc(uint32_t a, uint32_t acc) {
    return __dp4a(a, 0x01010101u, acc);
}

// occupancy nibble of one word: bit k = (byte k != 0)
__device__ __forceinline__ uint32_t nz_nibble(uint32_t x) {
    const uint32_t nz = (x | ((x & 0x7f7f7f7fu) + 0x7f7f7f7fu)) & 0x80808080u;   // bit 7 of every non-zero byte
    return (((nz >> 7) * 0x01020408u) >> 24) & 0xfu;                              // gather bits 0,8,16,24 -> 0..3
}

// bit k = byte k of x is outside [-8, 7] (an escape; zero bytes are inside)
__device__ __forceinline__ uint32_t esc_nibble(uint32_t x) {
    const uint32_t t = ((x & 0x7f7f7f7fu) + 0x08080808u) ^ (x & 0x80808080u);   // bytewise x + 8 (mod 256)
    return nz_nibble(t & 0xf0f0f0f0u);
}

template <int METRIC>
__device__ __forceinline__ uint32_t cost4_acc(uint32_t r, uint32_t c, uint32_t acc) {
    if (METRIC == 0) return wrap4_acc(r, c, acc);
    return sad4_acc(r, c, acc);
}

__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int m) {
    uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
    lo = __shfl_xor_sync(0xffffffffu, lo, m);
    hi = __shfl_xor_sync(0xffffffffu, hi, m);
    return ((unsigned long long)hi << 32) | lo;
}

__device__ __forceinline__ unsigned long long shfl_up_u64(unsigned long long v, int d) {
    uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
    lo = __shfl_up_sync(0xffffffffu, lo, d);
    hi = __shfl_up_sync(0xffffffffu, hi, d);
    return ((unsigned long long)hi << 32) | lo;
}

__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        unsigned long long o = shfl_xor_u64(v, m);
        v = o < v ? o : v;
    }
    return v;
}

// ---- OpenCV 8-bit fixed-point colour conversion (14 fractional bits) ------------------------
// Call sites in the reference: DCTcompressor.py:55,92.  Verified on all 2^24 inputs against
// cv2 4.13 through the oracle (tests/golden/golden_meta.json).
__device__ __forceinline__ int clip_u8(int v) { return min(max(v, 0), 255); }

__device__ __forceinline__ void bgr2ycrcb(int B, int G, int R, int &Y, int &Cr, int &Cb) {
    const int half = 1 << 13;
    // ranges over all 2^24 inputs: Y 0..255, Cr 0..256, Cb 1..255 -- only Cr can need the saturation
    Y = (1868 * B + 9617 * G + 4899 * R + half) >> 14;
    Cr = min(((R - Y) * 11682 + (128 << 14) + half) >> 14, 255);
    Cb = ((B - Y) * 9241 + (128 << 14) + half) >> 14;
}

__device__ __forceinline__ void ycrcb2bgr(int Y, int Cr, int Cb, int &B, int &G, int &R) {
    const int half = 1 << 13;
    Cr -= 128;
    Cb -= 128;
    B = clip_u8(Y + ((Cb * 29049 + half) >> 14));
    G = clip_u8(Y + ((Cb * -5636 + Cr * -11698 + half) >> 14));
    R = clip_u8(Y + ((Cr * 22987 + half) >> 14));
}

}  // namespace vcs
