// me_tiled.cuh -- step-1 full-search block matching, the hot kernel of the path (sm_100a).
//
// Replaces the candidate loop of MotionProcessor._find_match
// (InterframeCompression/motion.py:117-152) for step-1 searches (BASELINE.json configs 2/3/5):
// same candidate set (interval form of include/vcs_b200.h), same cost, same scan-order
// tie-break, same static early-out (motion.py:109-116).
//
// Design (INT32-ALU bound; DESIGN.md "me_tiled"):
//   * persistent CTAs, one per SM; a work unit is (tile of MX x MY macroblocks, ND x ND chunk of
//     the offset range).  The reference window of a unit is fetched by ONE TMA tile load
//     (cp.async.bulk.tensor, zero fill outside the frame) into ONE raw stage: it is free again as soon as the
//     window has been re-laid, so the next unit's loads are issued right then and land while this
//     unit is being searched; the tile's macroblocks come by a second TMA load.
//   * frames are BGR-interleaved, so a dx shift is 3 bytes and 3 of 4 candidates are not word
//     aligned.  The raw window is re-laid once per unit into 4 byte-phase copies, transposed to
//     [phase][word column][row]; after that every candidate reads ALIGNED 128-bit words and every
//     VABSDIFF4 lane does algorithmic work (no pad byte, no per-use funnel shift).
//   * thread = (macroblock, dx): at ND = 33 warp m holds dx 0..31 of macroblock m and one more warp dx = 32 of
//     all macroblocks.  A thread keeps ND accumulators (all dy of the chunk) and, per word
//     column, the BS macroblock words in registers; one LDS.128 delivers 4 window rows that feed
//     up to 4*min(ND,BS) VABSDIFF4.U8.ACC -- ~33 ALU ops per shared-memory load at ND=33.
//     Row padding RP = 16 (mod 32) words and phase stride PS = 4 (mod 32) words make the 16-byte
//     bank group of a lane equal (3*dx + const) mod 8, so 8 consecutive dx never conflict.
//   * the minimum is a packed 64-bit key (cost << 32 | dy_index << 16 | dx_index): min over keys
//     is the first strict minimum in rows-outer/cols-inner scan order; threads reduce over their
//     ND dy values in registers and across the macroblock with a shared-memory atomicMin.
//   * wrap8 cost (the reference's metric): per word t = (r|H) - (c&~H); z = t ^ (~r&H) ^ (c&H);
//     acc += 64 * bytesum(z)  (IADD + LOP3 + IDP.4A instead of one VABSDIFF4; the factor 64 is the key scale).
#pragma once
#include <cuda.h>

#include "common.cuh"

// wrap8 loop: the subtracts of the offsets d with bit (d mod VCS_WRAP_ALU_DEN) set in VCS_WRAP_ALU_MASK are forced onto
// the ALU pipe (IADD3), the rest go to the FMA pipe (IMAD.IADD).  Round-2 A/B of 18 splits on one B200 (tools/ab_me.sh +
// tools/ab_run.sh, two runs each, 45 P-frames): 0xa/5 9.02 ms, 0x9/5 9.05, 0x29/8 9.06, 0x6/5 9.10, ... 0x1/2 9.24 -- the
// instruction mix is the same for equal ratios, only the ptxas schedule differs.
#ifndef VCS_CUR_PAD
#define VCS_CUR_PAD 4
#endif
#ifndef VCS_WRAP_ALU_MASK
#define VCS_WRAP_ALU_MASK 0xa
#define VCS_WRAP_ALU_DEN 5
#endif

namespace vcs {

// ---- compile-time configuration of one kernel variant -------------------------------------
template <int BS_, int ND_>
struct TiledCfg {
    static constexpr int BS = BS_;            // macroblock size (8 or 16)
    static constexpr int ND = ND_;            // offsets per chunk and axis (9, 17 or 33)
    // macroblocks per tile: the more offsets a thread set covers, the fewer macroblocks fill the CTA.  Small ranges get
    // larger tiles (the window overhead (tile + ND)^2 / tile^2 shrinks and the CTA keeps 8-12 warps instead of 4-8).
    static constexpr int MX = ND_ == 33 ? 5 : 7;              // x
    static constexpr int MY = ND_ == 33 ? 3 : (ND_ == 17 ? 3 : 4);   // y
    static constexpr int NMB = MX * MY;
    static constexpr int ITEMS = NMB * ND;    // (macroblock, dx) pairs
    static constexpr int THREADS = (ITEMS + 127) / 128 * 128;
    static constexpr int WPR = 3 * BS / 4;    // words per macroblock row
    static constexpr int WPX = ND - 1 + BS * MX;  // window width, pixels
    static constexpr int WR = ND - 1 + BS * MY;   // window rows
    static constexpr int RP = (WR - 16 + 31) / 32 * 32 + 16;  // padded rows, = 16 (mod 32)
    // raw stage: words per row; multiple of 4 with RAWW/4 odd (conflict-free LDS.128 by row)
    // TMA needs a 16-byte aligned global start, so up to 15 bytes precede the window in a raw row
    static constexpr int RAW_NEED = (15 + 3 * WPX + 3 + 3) / 4 + 1;
    static constexpr int RAWW4 = (RAW_NEED + 3) / 4;
    static constexpr int RAWW = (RAWW4 % 2 ? RAWW4 : RAWW4 + 1) * 4;
    static constexpr int NC = ((15 + 3 * (BS * (MX - 1) + ND - 1)) >> 2) + WPR;  // word columns of T
    static constexpr int PS = (NC * RP + 31) / 32 * 32 + 4;  // phase stride, = 4 (mod 32)
    static constexpr int CURW4 = (MX * WPR + 3 + 3) / 4;
    static constexpr int CURW = (CURW4 % 2 ? CURW4 : CURW4 + 1) * 4;  // words per raw MB-tile row: 16 B multiple, CURW/4 odd
    static constexpr int CURR = MY * BS;      // rows of it
    // shared memory carve-up (bytes)
    static constexpr size_t OFF_T = 0;
    static constexpr size_t SZ_T = (size_t)4 * PS * 4;
    static constexpr size_t OFF_RAW = (OFF_T + SZ_T + 127) / 128 * 128;
    static constexpr size_t SZ_RAW = (size_t)(RAWW * WR + 4) * 4;     // +4: funnel read past the end
    static constexpr size_t SZ_RAW_AL = (SZ_RAW + 127) / 128 * 128;
    // ONE raw stage: it is free again as soon as the window has been re-laid, i.e. for the whole search of the unit,
    // which is when the next unit's TMA loads land in it
    static constexpr size_t OFF_CRAW = OFF_RAW + SZ_RAW_AL;
    static constexpr size_t SZ_CRAW = (size_t)CURW * CURR * 4;
    static constexpr size_t SZ_CRAW_AL = (SZ_CRAW + 127) / 128 * 128;
    static constexpr size_t OFF_CURT = OFF_CRAW + SZ_CRAW_AL;
    // macroblock words [mb][w][v], macroblocks CURS words apart: +4 (one 16-byte bank group) so that the lanes of the warp
    // that holds dx = ND-1 of ALL macroblocks read different banks (a stride of WPR*BS = 0 mod 32 words made its 8
    // LDS.128 per word column 15-way conflicts; the macroblock warps' loads are broadcasts either way).  Timing is
    // unchanged (200.3 against 200.4 us per P-frame): that warp's loads were never on the critical path.
    static constexpr int CURS = WPR * BS + VCS_CUR_PAD;
    static constexpr size_t SZ_CURT = (size_t)NMB * CURS * 4;
    static constexpr size_t OFF_MISC = (OFF_CURT + 2 * SZ_CURT + 127) / 128 * 128;  // 2: wrap8 L/H
    static constexpr size_t SZ_MISC = 8 * NMB + 8 * NMB + 4 * NMB + 4 * NMB + 64;
    static constexpr size_t SMEM = OFF_MISC + SZ_MISC + 128;
};

struct TiledArgs {
    int H, W, nbx, nby;
    int lo, hi, slack;
    long long static_thr;
    int tiles_x, tiles_y;      // tiles per frame
    int ncy, ncx;              // chunks of the offset range per axis
    int zchunk;                // index (cy*ncx+cx) of the chunk holding offset (0,0), or 0
    int npairs, ppg, p_off;
    uint32_t zero;             // always 0; opaque to the compiler (pipe balancing, see the wrap8 loop)
    int16_t *mv;
    uint32_t *cost;
    uint8_t *flags;
};

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *tm, uint64_t *bar, int c0, int c1,
                                            int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *tm, uint64_t *bar, int c0, int c1,
                                            int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ uint4 lds128(const uint32_t *p) { return *reinterpret_cast<const uint4 *>(p); }

__device__ __forceinline__ int floordiv16(int v) { return v >> 4; }  // arithmetic shift = floor for negatives

// ---- the kernel ---------------------------------------------------------------------------------
template <class C, int METRIC>
__global__ void __launch_bounds__(C::THREADS, 1)
me_tiled_kernel(const __grid_constant__ CUtensorMap tm_ref, const __grid_constant__ CUtensorMap tm_cur,
                const TiledArgs a) {
    constexpr int BS = C::BS, ND = C::ND, MX = C::MX, NMB = C::NMB, WPR = C::WPR, RP = C::RP, PS = C::PS;
    constexpr uint32_t Hm = 0x80808080u;
    const uint32_t zero = a.zero;
    extern __shared__ __align__(128) unsigned char smem[];
    uint32_t *sT = reinterpret_cast<uint32_t *>(smem + C::OFF_T);
    uint32_t *sCurT = reinterpret_cast<uint32_t *>(smem + C::OFF_CURT);           // SAD: c ; wrap8: c & ~H
    uint32_t *sCurH = sCurT + NMB * C::CURS;                                       // wrap8: c & H
    unsigned long long *sBest = reinterpret_cast<unsigned long long *>(smem + C::OFF_MISC);
    unsigned long long *sDyMask = sBest + NMB;                                     // valid dy bits per mb
    uint32_t *sStatic = reinterpret_cast<uint32_t *>(sDyMask + NMB);               // 0 / 1 per mb
    uint32_t *sStaticS = sStatic + NMB;                                            // static sum S
    uint64_t *sBar = reinterpret_cast<uint64_t *>(sStaticS + NMB);                 // 2 mbarriers (8B aligned)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nch = a.ncy * a.ncx;
    const int tiles_per_pair = a.tiles_x * a.tiles_y;
    const int ntiles = tiles_per_pair * a.npairs;        // < 2^31 (checked on the host)
    // tiles of this CTA: blockIdx.x, +gridDim.x, ...
    const int my_tiles = ntiles > (int)blockIdx.x ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int n_units = my_tiles * nch;
    if (n_units == 0) return;

    // geometry of work unit s of this CTA.  32-bit arithmetic only: this runs in every thread for every unit, and the
    // 64-bit divisions it used to make were a third of the instructions outside the inner loop.
    auto unit_geom = [&](int s, int &p, int &tx, int &ty, int &cy, int &cx, bool &first, bool &last) {
        const unsigned q = nch == 1 ? (unsigned)s : (unsigned)s / (unsigned)nch;
        const int c = nch == 1 ? 0 : s - (int)q * nch;
        const unsigned tile = blockIdx.x + q * gridDim.x;
        int cc = c + a.zchunk;                         // the chunk holding offset (0,0) goes first
        if (cc >= nch) cc -= nch;
        first = c == 0;
        last = c == nch - 1;
        p = (int)(tile / (unsigned)tiles_per_pair);
        const unsigned t = tile - (unsigned)p * (unsigned)tiles_per_pair;
        ty = (int)(t / (unsigned)a.tiles_x);
        tx = (int)t - ty * a.tiles_x;
        cy = nch == 1 ? 0 : cc / a.ncx;
        cx = cc - cy * a.ncx;
    };
    auto issue = [&](int s) {   // one elected thread: arm the barrier, start both TMA loads
        int p, tx, ty, cy, cx; bool f, l;
        unit_geom(s, p, tx, ty, cy, cx, f, l);
        const int xw0 = tx * MX * BS + a.lo + cx * ND, yw0 = ty * C::MY * BS + a.lo + cy * ND;
        mbar_expect_tx(&sBar[0], (uint32_t)(C::RAWW * C::WR * 4 + C::CURW * C::CURR * 4));
        // the innermost TMA coordinate must land on a 16-byte boundary: align down, keep the remainder
        const int pg = p + a.p_off;   // launches may start inside a GOP
        tma_load_3d(smem + C::OFF_RAW, &tm_ref, &sBar[0], 4 * floordiv16(3 * xw0), yw0, pg / a.ppg);
        tma_load_4d(smem + C::OFF_CRAW, &tm_cur, &sBar[0], 4 * floordiv16(tx * MX * BS * 3), ty * C::MY * BS, pg % a.ppg,
                    pg / a.ppg);
    };

    if (tid == 0) {
        mbar_init(&sBar[0], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) issue(0);

    // per-thread role in the search: (macroblock, dx index inside the chunk).  A quarter-warp that straddles two
    // macroblocks breaks the 3*dx bank rotation of its LDS.128 (2-way conflict), so when ND = 32k + 1 warp m takes
    // dx 0..31 of macroblock m (every LDS of the macroblock words is then a pure broadcast, and a static or
    // out-of-frame macroblock idles whole warps) and the leftover dx = ND-1 of all macroblocks share the last warp(s).
    constexpr bool WARP_PER_MB = ND == 33 && 32 * NMB + NMB <= C::THREADS;
    int mb, dxw;
    if (WARP_PER_MB) {
        if (tid < 32 * NMB) { mb = tid >> 5; dxw = tid & 31; }
        else { mb = tid - 32 * NMB; dxw = 32; }
    } else {
        mb = tid / ND; dxw = tid - mb * ND;
    }
    const int mbx = mb % MX, mby = mb / MX;
    const bool has_item = WARP_PER_MB ? tid < 33 * NMB : tid < C::ITEMS;

    for (int s = 0; s < n_units; ++s) {
        int p, tx, ty, cy, cx; bool first, last;
        unit_geom(s, p, tx, ty, cy, cx, first, last);
        const int xw0 = tx * MX * BS + a.lo + cx * ND;
        const int ao = 3 * xw0 - 16 * floordiv16(3 * xw0);   // byte offset of the window inside a raw row
        const int cao = (tx * MX * BS * 3 - 16 * floordiv16(tx * MX * BS * 3)) >> 2;   // word offset of the MB tile
        mbar_wait(&sBar[0], (uint32_t)(s & 1));

        // ---- re-lay the raw window: 4 byte phases, transposed [phase][word col][row] ----------
        {
            const uint32_t *raw = reinterpret_cast<const uint32_t *>(smem + C::OFF_RAW);
            constexpr int RG = (C::WR + 31) / 32, JQ = (C::NC + 3) / 4;   // row groups x column quads
            for (int task = warp; task < RG * JQ; task += C::THREADS / 32) {
                const int rg = task / JQ, jq = task - rg * JQ;
                const int r = rg * 32 + lane;
                if (r < C::WR) {
                    const uint32_t *src = raw + r * C::RAWW + 4 * jq;
                    const uint4 q = lds128(src);
                    const uint32_t nx = src[4];
                    const uint32_t wv[5] = {q.x, q.y, q.z, q.w, nx};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int j = 4 * jq + e;
                        if (j < C::NC) {
                            uint32_t *dst = sT + j * RP + r;
                            dst[0] = wv[e];
                            dst[PS] = __funnelshift_r(wv[e], wv[e + 1], 8);
                            dst[2 * PS] = __funnelshift_r(wv[e], wv[e + 1], 16);
                            dst[3 * PS] = __funnelshift_r(wv[e], wv[e + 1], 24);
                        }
                    }
                }
            }
            // macroblocks: raw [row][word] -> [mb][w][v]
            const uint32_t *craw = reinterpret_cast<const uint32_t *>(smem + C::OFF_CRAW);
            // lanes run along v (rows): conflict-free stores; a row pitch of CURW words with CURW/4 odd
            // keeps the strided reads at <= 2-way
            for (int k = tid; k < NMB * WPR * BS; k += C::THREADS) {
                const int v = k % BS, w = (k / BS) % WPR, m = k / (BS * WPR);
                const uint32_t c = craw[((m / MX) * BS + v) * C::CURW + cao + (m % MX) * WPR + w];
                const int kd = m * C::CURS + w * BS + v;
                if (METRIC == 0) { sCurT[kd] = c & ~Hm; sCurH[kd] = c & Hm; }
                else sCurT[kd] = c;
            }
            if (tid < NMB) {
                if (first) { sBest[tid] = ~0ull; sStatic[tid] = 0; }
                // valid dy of this chunk for macroblock row (tid / MX): offsets inside [lo,hi] whose row i = y + off lies
                // inside the frame, 0 <= i <= H - BS - slack: one interval of d, so a mask in closed form
                const int y = (ty * C::MY + tid / MX) * BS, off0 = a.lo + cy * ND;
                const int d_lo = max(0, -(y + off0));
                const int d_hi = min(ND - 1, min(a.hi - off0, a.H - BS - a.slack - y - off0));
                sDyMask[tid] = d_hi >= d_lo ? ((~0ull >> (63 - d_hi)) & (~0ull << d_lo)) : 0ull;
            }
        }
        __syncthreads();   // T, curT ready; the raw stage is free again: the next unit's loads overlap this unit's search
        if (tid == 0 && s + 1 < n_units) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(s + 1);
        }

        // ---- static test (motion.py:109-116), on the chunk that holds offset (0,0) ------------
        if (first && a.static_thr >= 0) {
            const int d0 = -(a.lo + cy * ND), x0w = -(a.lo + cx * ND);   // window-relative position of (0,0)
            for (int m = warp; m < NMB; m += C::THREADS / 32) {
                const int sb = ao + 3 * (BS * (m % MX) + x0w);
                const uint32_t *col = sT + (sb & 3) * PS + (sb >> 2) * RP + BS * (m / MX) + d0;
                uint32_t sad = 0, sr = 0, sc = 0;
                for (int k = lane; k < WPR * BS; k += 32) {
                    const int w = k / BS, v = k - w * BS;
                    const uint32_t r = col[w * RP + v];
                    uint32_t c = sCurT[m * C::CURS + w * BS + v];
                    if (METRIC == 0) c |= sCurH[m * C::CURS + w * BS + v];
                    sad = sad4_acc(r, c, sad);
                    sr = bytesum_acc(r, sr);
                    sc = bytesum_acc(c, sc);
                }
                int part = (int)sad + (int)sr - (int)sc;   // 2 * sum max(ref - cur, 0)
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                if (lane == 0) {
                    const long long S = part / 2;
                    sStatic[m] = S <= a.static_thr;
                    sStaticS[m] = (uint32_t)S;
                }
            }
            __syncthreads();
        }

        // ---- the search ---------------------------------------------------------------------------
        const int gx = tx * MX + mbx, gy = ty * C::MY + mby;     // macroblock coordinates in the frame
        bool active = has_item && gx < a.nbx && gy < a.nby && !sStatic[has_item ? mb : 0];
        if (active) {
            const int offx = a.lo + cx * ND + dxw, j = gx * BS + offx;
            active = offx <= a.hi && j >= 0 && j <= a.W - BS - a.slack;
        }
        const unsigned long long dymask = has_item ? sDyMask[mb] : 0;
        if (active && dymask) {
            const int sb = ao + 3 * (BS * mbx + dxw);
            const uint32_t *col = sT + (sb & 3) * PS + (sb >> 2) * RP + BS * mby;
            const uint32_t *cl = sCurT + mb * C::CURS;
            const uint32_t *ch = sCurH + mb * C::CURS;
            // wrap8 accumulates 64 * byte-sum (IDP.4A with every multiplier byte = 64), so that the argmin key
            // cost * 64 + d below needs no shift; SAD (VABSDIFF4.ACC adds plain bytes) scales at the end.
            constexpr uint32_t WRAP_MUL = 0x40404040u, KEY_MUL = METRIC == 0 ? 1u : 64u;
            uint32_t acc[ND];
#pragma unroll
            for (int d = 0; d < ND; ++d) acc[d] = 0;
#pragma unroll 1
            for (int w = 0; w < WPR; ++w) {
                uint32_t c[BS], chh[METRIC == 0 ? BS : 1];
#pragma unroll
                for (int q = 0; q < BS / 4; ++q) {
                    const uint4 t = lds128(cl + w * BS + 4 * q);
                    c[4 * q] = t.x; c[4 * q + 1] = t.y; c[4 * q + 2] = t.z; c[4 * q + 3] = t.w;
                    if (METRIC == 0) {
                        const uint4 u = lds128(ch + w * BS + 4 * q);
                        chh[4 * q] = u.x; chh[4 * q + 1] = u.y; chh[4 * q + 2] = u.z; chh[4 * q + 3] = u.w;
                    }
                }
                const uint32_t *cw = col + w * RP;
                constexpr int ROWS = ND + BS - 1;
#pragma unroll
                for (int rq = 0; rq < (ROWS + 3) / 4; ++rq) {
                    const uint4 t = lds128(cw + 4 * rq);
                    const uint32_t rr[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int r = 4 * rq + e;        // window row relative to the macroblock row
                        if (r < ROWS) {
                            if (METRIC == 0) {
                                // r1, r2 are pinned with asm so ptxas keeps ONE 3-input xor per use
                                // (it otherwise re-derives r2 inside every use: 2 LOP3 on the ALU pipe)
                                uint32_t r1, r2;
                                asm("lop3.b32 %0, %1, %2, %2, 0xfc;" : "=r"(r1) : "r"(rr[e]), "r"(Hm));   // r | H
                                asm("lop3.b32 %0, %1, %2, %2, 0x0c;" : "=r"(r2) : "r"(rr[e]), "r"(Hm));   // ~r & H
#pragma unroll
                                for (int d = 0; d < ND; ++d) {
                                    const int v = r - d;
                                    if (v >= 0 && v < BS) {
                                        // the subtract can issue on either pipe (IADD3 = ALU, IMAD.IADD =
                                        // FMA); ptxas sends them all to FMA, where IDP.4A also lives.  A third
                                        // (opaque zero) addend forces IADD3 for 2 of 5 so both pipes fill up.
                                        uint32_t z;
                                        if ((VCS_WRAP_ALU_MASK >> (d % VCS_WRAP_ALU_DEN)) & 1)
                                            asm("{\n.reg .u32 t;\nsub.u32 t, %1, %2;\nadd.u32 t, t, %5;\n"
                                                "lop3.b32 %0, t, %3, %4, 0x96;\n}"
                                                : "=r"(z) : "r"(r1), "r"(c[v]), "r"(r2), "r"(chh[v]), "r"(zero));
                                        else
                                            asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(z) : "r"(r1 - c[v]), "r"(r2), "r"(chh[v]));
                                        acc[d] = __dp4a(z, WRAP_MUL, acc[d]);
                                    }
                                }
                            } else {
#pragma unroll
                                for (int d = 0; d < ND; ++d) {
                                    const int v = r - d;
                                    if (v >= 0 && v < BS) acc[d] = sad4_acc(rr[e], c[v], acc[d]);
                                }
                            }
                        }
                    }
                }
            }
            // first strict minimum in scan order.  Within the thread dx is fixed, so (cost, dy) decides:
            // a 32-bit key cost*64 + d (cost < 2^18 for bs <= 16, d < 64) reduced with integer min.  wrap8 sums already
            // carry the factor 64, so a key is one fused add+min (VIADDMNMX); SAD sums are scaled here (LEA).
            static_assert(ND <= 64 && 255 * 3 * BS * BS < (1 << 26), "32-bit key layout");
            uint32_t k32 = 0xFFFFFFFFu;
            if (dymask == (~0ull >> (64 - ND))) {        // interior rows: every dy is valid (the common case)
#pragma unroll
                for (int d = 0; d < ND; ++d) k32 = min(k32, acc[d] * KEY_MUL + (uint32_t)d);
            } else {
#pragma unroll
                for (int d = 0; d < ND; ++d) {
                    const uint32_t k = ((dymask >> d) & 1) ? acc[d] * KEY_MUL + (uint32_t)d : 0xFFFFFFFFu;
                    k32 = min(k32, k);
                }
            }
            unsigned long long best = ~0ull;
            if (k32 != 0xFFFFFFFFu)
                best = ((unsigned long long)(k32 >> 6) << 32) | ((unsigned long long)(cy * ND + (int)(k32 & 63u)) << 16) |
                       (unsigned long long)(cx * ND + dxw);
            atomicMin(&sBest[mb], best);
        }
        __syncthreads();   // search of this unit done: T may be overwritten, sBest is complete

        if (last && tid < NMB) {
            const int ox = tx * MX + tid % MX, oy = ty * C::MY + tid / MX;
            if (ox < a.nbx && oy < a.nby) {
                const size_t o = (size_t)p * a.nbx * a.nby + (size_t)oy * a.nbx + ox;
                const unsigned long long k = sBest[tid];
                int mvx, mvy; uint32_t cst; uint8_t fl;
                if (sStatic[tid]) { mvx = 0; mvy = 0; cst = sStaticS[tid]; fl = 1; }
                else if (k == ~0ull) { mvx = -ox * BS; mvy = -oy * BS; cst = 0xFFFFFFFFu; fl = 2; }
                else {
                    mvx = a.lo + (int)(k & 0xffff); mvy = a.lo + (int)((k >> 16) & 0xffff);
                    cst = (uint32_t)(k >> 32); fl = 0;
                }
                a.mv[2 * o] = (int16_t)mvx;
                a.mv[2 * o + 1] = (int16_t)mvy;
                if (a.cost) a.cost[o] = cst;
                if (a.flags) a.flags[o] = fl;
            }
        }
    }
}

// ---- host side -------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

struct MeTiledState {
    PFN_encodeTiled encode = nullptr;
    // 0: persistent CTAs, one wave (one launch per clip).  k > 0: ceil(tiles / k) CTAs of ~k tiles each, so that the
    // CTA scheduler can start the next launch (or a higher-priority kernel) on SMs as they free up: the pipelined
    // host path, whose launches would otherwise each waste their last partial wave.
    int tiles_per_cta = 0;
    bool attr_set[2][3][2] = {{{false}}};
    int occupancy[2][3][2] = {{{0}}};
};

inline void me_tiled_destroy(MeTiledState &) {}

template <class C, int METRIC>
int me_tiled_run(MeTiledState &st, cudaStream_t stream, const MeGeom &g, const FrameAddr &fa, int npairs,
                 int16_t *mv, uint32_t *cost, uint8_t *flags, int sm_count, size_t smem_optin, int bi, int ni,
                 char *err, size_t errlen) {
    if (C::SMEM > smem_optin) { snprintf(err, errlen, "tiled ME needs %zu B shared memory", C::SMEM); return -4; }
    auto kern = me_tiled_kernel<C, METRIC>;
    if (!st.attr_set[bi][ni][METRIC]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
        if (e != cudaSuccess) { snprintf(err, errlen, "cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -2; }
        int occ = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, C::THREADS, C::SMEM);
        if (e != cudaSuccess || occ < 1) { snprintf(err, errlen, "tiled ME kernel does not fit an SM"); return -2; }
        st.occupancy[bi][ni][METRIC] = occ;
        st.attr_set[bi][ni][METRIC] = true;
    }
    const int pitch = 3 * g.W;
    const long long fs = (long long)pitch * g.H;
    const int nG = (fa.p_off + npairs + fa.ppg - 1) / fa.ppg;
    // reference frames: words x rows x gop
    CUtensorMap tm_ref, tm_cur;
    {
        cuuint64_t dims[3] = {(cuuint64_t)pitch / 4, (cuuint64_t)g.H, (cuuint64_t)nG};
        cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)(fa.ref_gop_stride ? fa.ref_gop_stride : fs)};
        cuuint32_t box[3] = {(cuuint32_t)C::RAWW, (cuuint32_t)C::WR, 1};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = st.encode(&tm_ref, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, (void *)fa.ref_base, dims, strides, box,
                               es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { snprintf(err, errlen, "cuTensorMapEncodeTiled(ref) -> %d", (int)r); return -2; }
    }
    {
        cuuint64_t dims[4] = {(cuuint64_t)pitch / 4, (cuuint64_t)g.H, (cuuint64_t)fa.ppg, (cuuint64_t)nG};
        cuuint64_t strides[3] = {(cuuint64_t)pitch, (cuuint64_t)(fa.cur_frame_stride ? fa.cur_frame_stride : fs),
                                 (cuuint64_t)(fa.cur_gop_stride ? fa.cur_gop_stride : fs * fa.ppg)};
        cuuint32_t box[4] = {(cuuint32_t)C::CURW, (cuuint32_t)C::CURR, 1, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = st.encode(&tm_cur, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, (void *)fa.cur_base, dims, strides, box,
                               es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { snprintf(err, errlen, "cuTensorMapEncodeTiled(cur) -> %d", (int)r); return -2; }
    }
    TiledArgs a;
    a.H = g.H; a.W = g.W; a.nbx = g.nbx; a.nby = g.nby; a.lo = g.lo; a.hi = g.hi; a.slack = g.slack;
    a.static_thr = g.static_thr;
    a.tiles_x = (g.nbx + C::MX - 1) / C::MX; a.tiles_y = (g.nby + C::MY - 1) / C::MY;
    const int nd_total = g.hi - g.lo + 1;
    a.ncy = a.ncx = (nd_total + C::ND - 1) / C::ND;
    a.zchunk = 0;
    if (g.lo <= 0 && g.hi >= 0) { const int cz = (-g.lo) / C::ND; a.zchunk = cz * a.ncx + cz; }
    a.npairs = npairs; a.ppg = fa.ppg; a.p_off = fa.p_off; a.zero = 0; a.mv = mv; a.cost = cost; a.flags = flags;
    const long long ntiles = (long long)a.tiles_x * a.tiles_y * npairs;
    if (ntiles * a.ncy * a.ncx >= (1ll << 31)) { snprintf(err, errlen, "too many work units in one launch (%lld tiles)", ntiles); return -1; }
    long long grid = (long long)sm_count * st.occupancy[bi][ni][METRIC];
    if (st.tiles_per_cta > 0) {
        const long long g2 = (ntiles + st.tiles_per_cta - 1) / st.tiles_per_cta;
        if (g2 > grid) grid = g2;
    }
    if (grid > ntiles) grid = ntiles;
    kern<<<(unsigned)grid, C::THREADS, C::SMEM, stream>>>(tm_ref, tm_cur, a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { snprintf(err, errlen, "me_tiled launch: %s", cudaGetErrorString(e)); return -2; }
    return 0;
}

// Runs the tiled kernel when it covers the request (*used = launches made), else leaves *used = 0.
inline int me_tiled_launch(MeTiledState &st, cudaStream_t stream, int metric, const MeGeom &g, const FrameAddr &fa,
                           int npairs, int16_t *mv, uint32_t *cost, uint8_t *flags, int sm_count,
                           size_t smem_optin, int *used, char *err, size_t errlen) {
    *used = 0;
    if (g.step != 1 || (g.bs != 8 && g.bs != 16)) return 0;
    if (g.W % 16) return 0;                                   // TMA: row pitch must be a multiple of 16 B
    const long long fs = 3ll * g.W * g.H;
    if (((uintptr_t)fa.ref_base | (uintptr_t)fa.cur_base) & 15) return 0;
    if ((fa.ref_gop_stride | fa.cur_gop_stride | fa.cur_frame_stride | fs) & 15) return 0;
    const int nd_total = g.hi - g.lo + 1;
    if (nd_total > 0xffff) return 0;
    if (g.static_thr >= 0 && !(g.lo <= 0 && g.hi >= 0)) return 0;   // static test reads offset (0,0) from the window
    if (!st.encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
        if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
            snprintf(err, errlen, "cuTensorMapEncodeTiled not available from the driver");
            return -2;
        }
        st.encode = (PFN_encodeTiled)fn;
    }
    const int ni = nd_total <= 9 ? 0 : (nd_total <= 17 ? 1 : 2);
    const int bi = g.bs == 16 ? 1 : 0;
    int rc;
#define VCS_RUN(BSV, NDV)                                                                                       \
    rc = metric == 0 ? me_tiled_run<TiledCfg<BSV, NDV>, 0>(st, stream, g, fa, npairs, mv, cost, flags, sm_count, \
                                                            smem_optin, bi, ni, err, errlen)                    \
                     : me_tiled_run<TiledCfg<BSV, NDV>, 1>(st, stream, g, fa, npairs, mv, cost, flags, sm_count, \
                                                            smem_optin, bi, ni, err, errlen)
    if (bi == 1) {
        if (ni == 0) { VCS_RUN(16, 9); } else if (ni == 1) { VCS_RUN(16, 17); } else { VCS_RUN(16, 33); }
    } else {
        if (ni == 0) { VCS_RUN(8, 9); } else if (ni == 1) { VCS_RUN(8, 17); } else { VCS_RUN(8, 33); }
    }
#undef VCS_RUN
    if (rc) return rc;
    *used = 1;
    return 0;
}

}  // namespace vcs
