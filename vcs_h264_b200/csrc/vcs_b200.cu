// vcs_b200.cu -- C ABI (include/vcs_b200.h) of the B200-native VCS-h264 interframe hot path.
//
// Host side only: argument checking, device/pinned scratch, stream plumbing and kernel
// launches.  All arithmetic lives in the kernels (me_generic.cuh, me_tiled.cuh,
// dct_stage.cuh).  There is no CPU fallback anywhere in this library.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <chrono>
#include <vector>

#include "../../include/vcs_b200.h"
#include "common.cuh"
#include "chroma.cuh"
#include "dct_stage.cuh"
#include "intra.cuh"
#include "me_generic.cuh"
#include "me_tiled.cuh"
#include "microbench.cuh"
#include "pack.cuh"

using namespace vcs;

namespace {

constexpr int NUM_DEV_SLOTS = 19;
enum DevSlot { S_FRAMES = 0, S_MV, S_COST, S_FLAGS, S_COEF, S_RECON, S_AUX0, S_AUX1, S_AUX2, S_MB, S_CYC,
               S_PK_BITMAP, S_PK_ROWCNT, S_PK_ROWOFF, S_PK_VALUES, S_PK_TOTAL, S_PK_BLKESC, S_PK_ESC };

struct EvTriple {
    cudaEvent_t e0, e1, e2;
    bool has_me, has_dct;
};

}  // namespace

struct vcs_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr, s_aux = nullptr;
    cudaStream_t s_search = nullptr, s_dct = nullptr;   // host pipeline: low-priority searches, high-priority DCT/pack
    std::vector<cudaEvent_t> me_events;
    char err[512] = {0};
    double h_Q[192];
    double *d_Q = nullptr;
    int64_t launches = 0;
    int sm_count = 0, cc_major = 0, cc_minor = 0;
    int dct_occupancy[2][4][4] = {{{0}}};   // [fp32 tier][coef_mode][path], filled on first use
    int dct_fp32 = 0;                  // vcs_set_dct_precision: 0 = float64 exact (default), 1 = fp32 tier
    size_t smem_optin = 0;
    void *dev[NUM_DEV_SLOTS] = {nullptr};
    size_t dev_cap[NUM_DEV_SLOTS] = {0};
    std::vector<cudaEvent_t> chunk_events;
    bool timing = false;
    std::vector<EvTriple> ev_pool;
    size_t ev_used = 0;
    MeTiledState tiled;
    int *h_errflag = nullptr;          // pinned, mapped: kernels store 1 when a motion vector leaves the frame
    unsigned long long *h_segend = nullptr;   // pinned: running value-stream length after each pipeline segment
    size_t h_segend_cap = 0;
    std::vector<cudaEvent_t> seg_events;
    uint64_t q_version = 0;
};

namespace {

int fail(vcs_ctx *ctx, int code, const char *fmt, ...) {
    if (ctx) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(ctx->err, sizeof(ctx->err), fmt, ap);
        va_end(ap);
    }
    return code;
}

// Every entry point runs with the context's device current and puts the caller's device back on return
// (torch and other contexts in the same process keep theirs).
struct DevGuard {
    int prev = -1, dev;
    explicit DevGuard(int d) : dev(d) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) cudaSetDevice(dev);
    }
    ~DevGuard() { if (prev >= 0 && prev != dev) cudaSetDevice(prev); }
};

// A kernel that met a motion vector pointing outside the frame wrote zero prediction there and raised the flag;
// the next call on the context (or vcs_synchronize) reports it once.
int pending_device_error(vcs_ctx *ctx) {
    if (ctx->h_errflag && *(volatile int *)ctx->h_errflag) {
        const int what = *(volatile int *)ctx->h_errflag;
        *(volatile int *)ctx->h_errflag = 0;
        if (what == 2)
            return fail(ctx, VCS_E_INVALID, "an earlier launch met a packed coefficient stream shorter than its bitmaps "
                                            "claim (the affected blocks were decoded as zero)");
        return fail(ctx, VCS_E_INVALID, "an earlier launch met a motion vector pointing outside the frame "
                                        "(the prediction was zero-filled there)");
    }
    return VCS_OK;
}

#define VCS_ENTER(ctx)                                       \
    if (!(ctx)) return VCS_E_INVALID;                        \
    DevGuard dev_guard__((ctx)->device);                     \
    do { int rc__ = pending_device_error(ctx); if (rc__) return rc__; } while (0)

#define CK(ctx, call)                                                                          \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess)                                                                \
            return fail(ctx, VCS_E_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call,          \
                        cudaGetErrorString(e__));                                              \
    } while (0)

int dev_buf(vcs_ctx *ctx, int slot, size_t bytes, void **out) {
    if (bytes > ctx->dev_cap[slot]) {
        if (ctx->dev[slot]) {
            CK(ctx, cudaDeviceSynchronize());
            CK(ctx, cudaFree(ctx->dev[slot]));
            ctx->dev[slot] = nullptr;
            ctx->dev_cap[slot] = 0;
        }
        size_t cap = (bytes + 255) & ~size_t(255);
        cudaError_t e = cudaMalloc(&ctx->dev[slot], cap);
        if (e != cudaSuccess)
            return fail(ctx, VCS_E_NOMEM, "cudaMalloc(%zu) -> %s", cap, cudaGetErrorString(e));
        ctx->dev_cap[slot] = cap;
    }
    *out = ctx->dev[slot];
    return VCS_OK;
}

int check_me_params(vcs_ctx *ctx, const vcs_me_params *p) {
    if (!p) return fail(ctx, VCS_E_INVALID, "params is NULL");
    if (p->bs <= 0 || p->step <= 0 || p->H < p->bs || p->W < p->bs || p->slack < 0)
        return fail(ctx, VCS_E_INVALID, "bad ME geometry H=%d W=%d bs=%d step=%d", p->H, p->W,
                    p->bs, p->step);
    if (p->metric != VCS_METRIC_WRAP8 && p->metric != VCS_METRIC_SAD)
        return fail(ctx, VCS_E_INVALID, "unknown metric %d", p->metric);
    if (p->hi < p->lo) return fail(ctx, VCS_E_INVALID, "empty offset interval [%d,%d]", p->lo, p->hi);
    if (p->H > 32767 || p->W > 32767)
        return fail(ctx, VCS_E_INVALID, "frame larger than int16 motion vectors allow");
    if ((long long)255 * 3 * p->bs * p->bs >= (1ll << 32))
        return fail(ctx, VCS_E_INVALID, "block too large for 32-bit costs");
    return VCS_OK;
}

MeGeom make_geom(const vcs_me_params *p) {
    MeGeom g;
    g.H = p->H; g.W = p->W; g.bs = p->bs; g.lo = p->lo; g.hi = p->hi; g.step = p->step;
    g.slack = p->slack; g.nbx = p->W / p->bs; g.nby = p->H / p->bs; g.static_thr = p->static_thr;
    return g;
}

FrameAddr pair_addr(const uint8_t *cur, const uint8_t *ref) {
    FrameAddr fa;
    fa.cur_base = cur; fa.ref_base = ref;
    fa.cur_gop_stride = 0; fa.cur_frame_stride = 0; fa.ref_gop_stride = 0; fa.ppg = 1;
    return fa;
}

// Clip under the reference's GOP rule (encoder.py:25,51-52).  Only whole GOPs plus a trailing
// partial GOP are addressed: P-ordinal p -> gop p/(g-1), frame 1 + p%(g-1) inside it.
FrameAddr clip_addr(const uint8_t *frames, int H, int W, int gop_len) {
    const long long fs = (long long)H * W * 3;
    FrameAddr fa;
    fa.cur_base = frames + fs; fa.ref_base = frames;
    fa.cur_gop_stride = fs * gop_len; fa.cur_frame_stride = fs; fa.ref_gop_stride = fs * gop_len;
    fa.ppg = gop_len - 1;
    return fa;
}

int launch_me(vcs_ctx *ctx, cudaStream_t st, const vcs_me_params *p, const FrameAddr &fa, int npairs,
              int16_t *mv, uint32_t *cost, uint8_t *flags) {
    int rc = check_me_params(ctx, p);
    if (rc) return rc;
    if (npairs <= 0) return VCS_OK;
    if (!mv) return fail(ctx, VCS_E_INVALID, "mv output is NULL");
    const MeGeom g = make_geom(p);
    const int N = g.nbx * g.nby;
    if (p->kernel != VCS_ME_GENERIC) {
        int used = 0;
        rc = me_tiled_launch(ctx->tiled, st, p->metric, g, fa, npairs, mv, cost, flags, ctx->sm_count,
                             ctx->smem_optin, &used, ctx->err, sizeof(ctx->err));
        if (rc) return rc;
        if (used) { ctx->launches += used; return VCS_OK; }
        if (p->kernel == VCS_ME_TILED)
            return fail(ctx, VCS_E_UNSUPPORTED, "tiled ME kernel does not cover bs=%d step=%d lo=%d hi=%d",
                        p->bs, p->step, p->lo, p->hi);
    }
    const int rw = (3 * g.bs + 3) / 4;
    const size_t smem = (size_t)g.bs * rw * 4;
    if (smem > 48 * 1024) return fail(ctx, VCS_E_UNSUPPORTED, "bs=%d too large for the generic kernel", g.bs);
    dim3 grid(N, npairs);
    if (p->metric == VCS_METRIC_WRAP8)
        me_generic_kernel<0><<<grid, ME_GENERIC_THREADS, smem, st>>>(fa, g, mv, cost, flags);
    else
        me_generic_kernel<1><<<grid, ME_GENERIC_THREADS, smem, st>>>(fa, g, mv, cost, flags);
    CK(ctx, cudaGetLastError());
    ctx->launches += 1;
    return VCS_OK;
}

int launch_dct(vcs_ctx *ctx, cudaStream_t st, DctArgs &a, int nP) {
    if (a.H <= 0 || a.W <= 0 || a.H % 8 || a.W % 8)
        return fail(ctx, VCS_E_INVALID,
                    "H=%d W=%d must be multiples of 8 (the reference resizes, DCTcompressor.py:52)", a.H, a.W);
    if (a.coef_mode < 0 || a.coef_mode > 3) return fail(ctx, VCS_E_INVALID, "coef_mode %d", a.coef_mode);
    if (a.coef_mode == VCS_COEF_I8_RINT) {
        // |D| <= 8 * 128 = 1024 for inputs in [-128,127], so |index| <= 1024 / min(Q): int8 is lossless iff that fits
        double qmin = ctx->h_Q[0];
        for (int k = 1; k < 192; ++k) qmin = ctx->h_Q[k] < qmin ? ctx->h_Q[k] : qmin;
        if (1024.0 / qmin > 127.0)
            return fail(ctx, VCS_E_INVALID, "int8 indices are not lossless for this Q (min %g < 9): use VCS_COEF_I16_RINT", qmin);
    }
    if (nP <= 0) return VCS_OK;
    // vector accesses: coefficient rows go out as 8 x int8 / int16 / 2 x float64 (16-byte stores for the wide
    // formats), pixel groups are read as aligned 32-bit words
    if ((uintptr_t)a.coef & 15) return fail(ctx, VCS_E_INVALID, "coefficient buffer must be 16-byte aligned");
    if (((uintptr_t)a.img | (uintptr_t)a.pred_in | (uintptr_t)a.recon |
         (a.has_fa ? ((uintptr_t)a.fa.cur_base | (uintptr_t)a.fa.ref_base) : 0)) & 3)
        return fail(ctx, VCS_E_INVALID, "image buffers must be 4-byte aligned");
    const bool inv = a.inverse && a.recon;
    if (!a.forward && !inv) return fail(ctx, VCS_E_INVALID, "nothing to do: no forward pass and no reconstruction");
    if (a.forward && !inv && !a.coef) return fail(ctx, VCS_E_INVALID, "forward pass without an output");
    if (!a.forward && !a.coef) return fail(ctx, VCS_E_INVALID, "inverse pass without coefficients");
    const int path = !a.forward ? DCT_INV : (!inv ? DCT_FWD : (a.coef ? DCT_FWD_INV : DCT_FWD_INV_NOCOEF));
    typedef void (*dct_fn)(const DctArgs, int);
#define VCS_DCT_ROW(CM, R) {dct_stage_kernel<CM, 0, R>, dct_stage_kernel<CM, 1, R>, dct_stage_kernel<CM, 2, R>, dct_stage_kernel<CM, 3, R>}
    static const dct_fn table[2][4][4] = {{VCS_DCT_ROW(0, double), VCS_DCT_ROW(1, double), VCS_DCT_ROW(2, double), VCS_DCT_ROW(3, double)},
                                          {VCS_DCT_ROW(0, float), VCS_DCT_ROW(1, float), VCS_DCT_ROW(2, float), VCS_DCT_ROW(3, float)}};
#undef VCS_DCT_ROW
    const int f32 = ctx->dct_fp32 ? 1 : 0;
    if (f32 && a.bitmap) return fail(ctx, VCS_E_INVALID, "the packed sink belongs to the exact tier");
    const size_t smem_bytes = f32 ? DctSizes<float>::SMEM_BYTES : DctSizes<double>::SMEM_BYTES;
    const dct_fn kern = table[f32][a.coef_mode][path];
    int &occ = ctx->dct_occupancy[f32][a.coef_mode][path];
    if (occ == 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, DCT_THREADS, smem_bytes) != cudaSuccess || occ < 1) {
            occ = 0;
            return fail(ctx, VCS_E_CUDA, "dct_stage_kernel does not fit an SM");
        }
    }
    a.Q = ctx->d_Q;
    a.err = ctx->h_errflag;
    // persistent warps: 4 CTAs of 4 warps per SM, each warp walks 8x32-pixel tiles
    const long long nitems = (long long)((a.W + DCT_TILE_W - 1) / DCT_TILE_W) * (a.H / 8) * nP;
    if (nitems >= (1ll << 31)) return fail(ctx, VCS_E_INVALID, "too many 8x32 tiles in one launch (%lld)", nitems);
    long long grid = (long long)ctx->sm_count * occ;   // exactly one resident wave
    if (grid * DCT_WARPS > nitems) grid = (nitems + DCT_WARPS - 1) / DCT_WARPS;
    kern<<<(unsigned)grid, DCT_THREADS, smem_bytes, st>>>(a, nP);
    CK(ctx, cudaGetLastError());
    ctx->launches += 1;
    return VCS_OK;
}

size_t coef_elem(int coef_mode) { return coef_mode == VCS_COEF_I8_RINT ? 1 : (coef_mode == VCS_COEF_I16_RINT ? 2 : 8); }

EvTriple *next_events(vcs_ctx *ctx) {
    if (!ctx->timing) return nullptr;
    if (ctx->ev_used == ctx->ev_pool.size()) {
        EvTriple t;
        if (cudaEventCreate(&t.e0) || cudaEventCreate(&t.e1) || cudaEventCreate(&t.e2)) return nullptr;
        t.has_me = t.has_dct = false;
        ctx->ev_pool.push_back(t);
    }
    EvTriple *t = &ctx->ev_pool[ctx->ev_used++];
    t->has_me = t->has_dct = false;
    return t;
}

// ME + residual/DCT/recon for nP P-frames addressed by fa, on stream st.
// ME (+ residual/DCT/quant [+ recon]) of nP P-frames.  st2/ev_me: when given, the DCT stage runs on st2 after ev_me,
// which is recorded on st behind the search (the pipelined host path keeps the next search off the DCT's critical path).
int encode_dev(vcs_ctx *ctx, cudaStream_t st, const vcs_me_params *p, const FrameAddr &fa, int nP,
               int coef_mode, int16_t *mv, uint32_t *cost, uint8_t *flags, void *coef, uint8_t *recon,
               unsigned long long *bitmap = nullptr, uint8_t *blk_esc = nullptr, uint2 *row_count = nullptr,
               cudaStream_t st2 = nullptr, cudaEvent_t ev_me = nullptr) {
    EvTriple *ev = next_events(ctx);
    if (ev) CK(ctx, cudaEventRecord(ev->e0, st));
    int rc = launch_me(ctx, st, p, fa, nP, mv, cost, flags);
    if (rc) return rc;
    if (ev) { CK(ctx, cudaEventRecord(ev->e1, st)); ev->has_me = true; }
    if (st2) {
        CK(ctx, cudaEventRecord(ev_me, st));
        CK(ctx, cudaStreamWaitEvent(st2, ev_me, 0));
        st = st2;
    }
    if (coef || recon) {
        DctArgs a{};
        a.H = p->H; a.W = p->W; a.fa = fa; a.has_fa = 1; a.mv = mv; a.bs = p->bs;
        a.nbx = p->W / p->bs; a.nby = p->H / p->bs; a.forward = 1; a.inverse = recon != nullptr;
        a.coef_mode = coef_mode; a.coef = coef; a.recon = recon;
        if (bitmap && blk_esc && row_count && coef && !recon && coef_mode == VCS_COEF_I8_RINT) {   // forward-only variant of the kernel
            // the DCT stage emits the occupancy bitmaps, escape counts and (atomically) the per-block-row counts
            a.bitmap = reinterpret_cast<uint8_t *>(bitmap); a.blk_esc = blk_esc;
            a.row_count = reinterpret_cast<unsigned long long *>(row_count);
            CK(ctx, cudaMemsetAsync(row_count, 0, sizeof(uint2) * (size_t)nP * 3 * (p->H / 8), st));
        }
        rc = launch_dct(ctx, st, a, nP);
        if (rc) return rc;
        if (ev && !st2) { CK(ctx, cudaEventRecord(ev->e2, st)); ev->has_dct = true; }
    }
    return VCS_OK;
}

// ---- packed coefficient streams (pack.cuh) ---------------------------------------------------------------
struct PackedHost {              // host destinations of vcs_encode_clip_host_packed
    unsigned long long *bitmap;  // [nP][3][H/8][W/8]
    uint32_t *row_count;         // [nP][3][H/8][2]
    uint8_t *nibbles;            // nibble stream, cap_n bytes
    size_t cap_n;
    int8_t *escapes;             // escape stream, cap_e bytes
    size_t cap_e;
    unsigned long long *lengths; // out: [2] = bytes of the nibble stream, number of escapes
};

struct PackedDev {               // device side of one clip (or of one stand-alone pack call)
    unsigned long long *bitmap;  // [rows][W/8]
    uint8_t *blk_esc;            // [rows][W/8]
    uint2 *row_count;            // [rows]
    unsigned long long *nib_off, *esc_off;   // [rows]
    uint8_t *nibbles;
    int8_t *escapes;
    unsigned long long *totals;  // [2] running lengths
};

static int pack_grid(vcs_ctx *ctx, long long nwarps) {
    long long g = (nwarps + PACK_WARPS - 1) / PACK_WARPS;
    const long long cap = (long long)ctx->sm_count * 8;
    return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

// dense int8 planes of nP frames -> the packed form on the device, rows [row0, row0 + nP*3*H/8) of d.  d.totals (device,
// 2 x 8 bytes) are the running stream lengths: this call's rows are appended after them.  d_segend (device, 2 x 8
// bytes, may be null) receives the new totals.  have_bitmaps: the DCT stage already wrote bitmaps, escape counts and
// row counts for these rows.
static int launch_pack(vcs_ctx *ctx, cudaStream_t st, int H, int W, int nP, const int8_t *coef, const PackedDev &d,
                       size_t row0, unsigned long long *d_segend, bool have_bitmaps) {
    const int nrows = nP * 3 * (H / 8), nbx = W / 8;
    if (nrows <= 0) return VCS_OK;
    const int nbatch = (nbx + 31) / 32;
    if (!have_bitmaps) {     // stand-alone packing of given planes
        pack_count_kernel<<<pack_grid(ctx, nrows), 32 * PACK_WARPS, 0, st>>>(coef, W, nrows, d.bitmap + row0 * nbx,
                                                                            d.blk_esc + row0 * nbx, d.row_count + row0);
        ctx->launches += 1;
    }
    pack_scan_kernel<<<1, 1024, 0, st>>>(d.row_count + row0, nrows, d.nib_off + row0, d.esc_off + row0, d.totals, d_segend);
    pack_write_kernel<<<pack_grid(ctx, (long long)nrows * nbatch), 32 * PACK_WARPS, 0, st>>>(
        coef, W, nrows, d.bitmap + row0 * nbx, d.blk_esc + row0 * nbx, d.nib_off + row0, d.esc_off + row0, d.nibbles, d.escapes);
    CK(ctx, cudaGetLastError());
    ctx->launches += 2;
    return VCS_OK;
}

// scratch of the packed form for `rows` block rows of W/8 blocks; `dense` = bytes of the dense planes they code
static int packed_scratch(vcs_ctx *ctx, size_t rows, int W, size_t dense, int nsegs, PackedDev &d) {
    int rc;
    const size_t nbx = (size_t)W / 8;
    if ((rc = dev_buf(ctx, S_PK_BITMAP, rows * nbx * 8 + 8, (void **)&d.bitmap))) return rc;
    if ((rc = dev_buf(ctx, S_PK_BLKESC, rows * nbx + 8, (void **)&d.blk_esc))) return rc;
    if ((rc = dev_buf(ctx, S_PK_ROWCNT, rows * 8 + 8, (void **)&d.row_count))) return rc;
    unsigned long long *off;
    if ((rc = dev_buf(ctx, S_PK_ROWOFF, rows * 16 + 16, (void **)&off))) return rc;
    d.nib_off = off; d.esc_off = off + rows;
    if ((rc = dev_buf(ctx, S_PK_VALUES, dense / 2 + rows * nbx + 64, (void **)&d.nibbles))) return rc;   // <= 32 bytes per block
    if ((rc = dev_buf(ctx, S_PK_ESC, dense + 64, (void **)&d.escapes))) return rc;
    // [0..1]: running totals; [2 + 2c .. 3 + 2c]: their values after segment c (a private slot per segment: the
    // download stream may read it while the compute stream is already extending the streams for the next segment)
    if ((rc = dev_buf(ctx, S_PK_TOTAL, 16 * (size_t)(nsegs + 1), (void **)&d.totals))) return rc;
    return VCS_OK;
}

}  // namespace

template <int W>
int run_mb(vcs_ctx *ctx, int iters, double *rate, double *mhz) {
    const int blocks = ctx->sm_count * 8;
    uint32_t *d_out; long long *d_cyc; int rc;
    if ((rc = dev_buf(ctx, S_MB, (size_t)blocks * MB_THREADS * 4, (void **)&d_out))) return rc;
    if ((rc = dev_buf(ctx, S_CYC, 16, (void **)&d_cyc))) return rc;
    cudaEvent_t e0, e1;
    CK(ctx, cudaEventCreate(&e0));
    CK(ctx, cudaEventCreate(&e1));
    cudaStream_t st = ctx->stream;
    microbench_kernel<W><<<blocks, MB_THREADS, 0, st>>>(d_out, iters / 4 + 1, 12345u, d_cyc);  // warm-up
    CK(ctx, cudaEventRecord(e0, st));
    microbench_kernel<W><<<blocks, MB_THREADS, 0, st>>>(d_out, iters, 12345u, d_cyc);
    CK(ctx, cudaEventRecord(e1, st));
    CK(ctx, cudaStreamSynchronize(st));
    CK(ctx, cudaGetLastError());
    ctx->launches += 2;
    float ms;
    CK(ctx, cudaEventElapsedTime(&ms, e0, e1));
    long long cyc[2] = {0, 0};   // block 0: SM clocks and globaltimer nanoseconds over its own loop
    CK(ctx, cudaMemcpy(cyc, d_cyc, 16, cudaMemcpyDeviceToHost));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const double warps = (double)blocks * MB_THREADS / 32.0;
    double ops = warps * (double)iters * mb_ops_per_iter(W);
    if (W == 5) ops /= 3.0;  // report words/s/32 for the wrap8 triple
    if (W == 6) ops = warps * (double)iters * MB_ACC * MB_UNROLL;  // count the VABSDIFF4s only
    if (rate) *rate = ops / (ms * 1e-3);
    if (mhz) *mhz = cyc[1] > 0 ? (double)cyc[0] / ((double)cyc[1] * 1e-3) : 0.0;
    return VCS_OK;
}


// ------------------------------------------------------------------------------------------
extern "C" {

int vcs_version(void) { return 100; }

const char *vcs_last_error(const vcs_ctx *ctx) { return ctx ? ctx->err : "no context"; }

int vcs_dct_matrix(double *C) {
    // DCTCompressor._dctMatrix (DCTcompressor.py:124-133): host libm, like Python's math module
    if (!C) return VCS_E_INVALID;
    const int N = 8;
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j)
            C[i * N + j] = i == 0 ? 1.0 / sqrt((double)N)
                                  : sqrt(2.0 / N) * cos((double)((2 * j + 1) * i) * M_PI / (double)(2 * N));
    return VCS_OK;
}

int vcs_q_tables(double qf, double *Q) {
    // DCTcompressor.py:11-38.  Luma entry [1][5] is 48 as in the reference, not Annex K's 58.
    static const int QY[64] = {16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  48,  60,  55,
                               14, 13, 16, 24, 40,  57,  69,  56,  14, 17, 22, 29, 51,  87,  80,  62,
                               18, 22, 37, 56, 68,  109, 103, 77,  24, 35, 55, 64, 81,  104, 113, 92,
                               49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
    static const int QC[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99,
                               24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
                               99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                               99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};
    if (!Q) return VCS_E_INVALID;
    double scale;
    if (qf < 50 && qf > 1) scale = 50 / qf;
    else if (qf < 100) scale = (100 - qf) / 50;
    else return VCS_E_INVALID;
    for (int k = 0; k < 64; ++k) {
        double y = nearbyint(QY[k] * scale), c = nearbyint(QC[k] * scale);  // np.round: half-even
        y = fmin(fmax(y, 1.0), 255.0);
        c = fmin(fmax(c, 1.0), 255.0);
        Q[k] = y; Q[64 + k] = c; Q[128 + k] = c;
    }
    return VCS_OK;
}

int vcs_num_blocks(int H, int W, int bs) { return bs > 0 ? (H / bs) * (W / bs) : 0; }

int vcs_num_p_frames(int T, int gop_len) {
    if (T <= 0 || gop_len <= 0) return 0;
    return T - (T + gop_len - 1) / gop_len;
}

int vcs_me_reference_params(int H, int W, int bs, vcs_me_params *out) {
    if (!out || bs <= 0) return VCS_E_INVALID;
    memset(out, 0, sizeof(*out));
    const int R = 2 * bs;  // motion.py:18
    out->H = H; out->W = W; out->bs = bs;
    out->lo = -R; out->hi = R - bs - 1; out->slack = 1;  // motion.py:125-140
    out->step = (int)nearbyint(bs / 3.0);                // Python round(bs/3), motion.py:132
    if (out->step < 1) return VCS_E_INVALID;             // range() step 0 raises in the reference
    out->metric = VCS_METRIC_WRAP8;                      // motion.py:146
    out->static_thr = 2000;                              // motion.py:8
    out->kernel = VCS_ME_AUTO;
    return VCS_OK;
}

int vcs_me_fullsearch_params(int H, int W, int bs, int R, int metric, int64_t static_thr,
                             vcs_me_params *out) {
    if (!out || bs <= 0 || R < 0) return VCS_E_INVALID;
    memset(out, 0, sizeof(*out));
    out->H = H; out->W = W; out->bs = bs; out->lo = -R; out->hi = R; out->step = 1; out->slack = 0;
    out->metric = metric; out->static_thr = static_thr; out->kernel = VCS_ME_AUTO;
    return VCS_OK;
}

int vcs_create(int device, vcs_ctx **out) {
    if (!out) return VCS_E_INVALID;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0 || device < 0 || device >= n) return VCS_E_CUDA;  // no CPU fallback
    DevGuard guard(device);             // the caller's current device is restored on return
    vcs_ctx *ctx = new vcs_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        delete ctx;
        return VCS_E_CUDA;
    }
    ctx->sm_count = prop.multiProcessorCount;
    ctx->cc_major = prop.major; ctx->cc_minor = prop.minor;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    double C[64];
    float Cf[64];
    vcs_dct_matrix(C);
    for (int k = 0; k < 64; ++k) Cf[k] = (float)C[k];
    static_assert(DCT_SMEM_BYTES <= 48 * 1024, "dct_stage_kernel relies on the default shared-memory limit");
    vcs_q_tables(50.0, ctx->h_Q);  // DCTcompressor.py:29 QF = 50
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->s_aux, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithPriority(&ctx->s_search, cudaStreamNonBlocking, prio_lo) != cudaSuccess ||
        cudaStreamCreateWithPriority(&ctx->s_dct, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaMalloc(&ctx->d_Q, sizeof(ctx->h_Q)) != cudaSuccess ||
        cudaHostAlloc((void **)&ctx->h_errflag, sizeof(int), cudaHostAllocMapped) != cudaSuccess ||
        cudaMemcpyToSymbol(c_dct, C, sizeof(C)) != cudaSuccess ||
        cudaMemcpyToSymbol(c_dctf, Cf, sizeof(Cf)) != cudaSuccess ||
        cudaMemcpy(ctx->d_Q, ctx->h_Q, sizeof(ctx->h_Q), cudaMemcpyHostToDevice) != cudaSuccess) {
        vcs_destroy(ctx);
        return VCS_E_CUDA;
    }
    *ctx->h_errflag = 0;
    ctx->stream = ctx->own_stream;
    *out = ctx;
    return VCS_OK;
}

int vcs_destroy(vcs_ctx *ctx) {
    if (!ctx) return VCS_OK;
    DevGuard guard(ctx->device);
    cudaDeviceSynchronize();
    for (int s = 0; s < NUM_DEV_SLOTS; ++s)
        if (ctx->dev[s]) cudaFree(ctx->dev[s]);
    if (ctx->d_Q) cudaFree(ctx->d_Q);
    if (ctx->h_errflag) cudaFreeHost(ctx->h_errflag);
    for (auto &t : ctx->ev_pool) { cudaEventDestroy(t.e0); cudaEventDestroy(t.e1); cudaEventDestroy(t.e2); }
    for (auto &e : ctx->chunk_events) cudaEventDestroy(e);
    for (auto &e : ctx->seg_events) cudaEventDestroy(e);
    if (ctx->h_segend) cudaFreeHost(ctx->h_segend);
    me_tiled_destroy(ctx->tiled);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    if (ctx->s_h2d) cudaStreamDestroy(ctx->s_h2d);
    if (ctx->s_d2h) cudaStreamDestroy(ctx->s_d2h);
    if (ctx->s_aux) cudaStreamDestroy(ctx->s_aux);
    if (ctx->s_search) cudaStreamDestroy(ctx->s_search);
    if (ctx->s_dct) cudaStreamDestroy(ctx->s_dct);
    for (auto &e : ctx->me_events) cudaEventDestroy(e);
    delete ctx;
    return VCS_OK;
}

int vcs_set_stream(vcs_ctx *ctx, void *cuda_stream) {
    if (!ctx) return VCS_E_INVALID;
    ctx->stream = (cudaStream_t)cuda_stream;   // 0 is CUDA's legacy default stream, taken literally
    return VCS_OK;
}

int vcs_use_own_stream(vcs_ctx *ctx) {
    if (!ctx) return VCS_E_INVALID;
    ctx->stream = ctx->own_stream;
    return VCS_OK;
}

int vcs_synchronize(vcs_ctx *ctx) {
    if (!ctx) return VCS_E_INVALID;
    DevGuard guard(ctx->device);
    // first wait, then look at the device's error flag: a kernel that is still running may raise it again
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return pending_device_error(ctx);
}

int vcs_device_info(vcs_ctx *ctx, int *sm_count, int *cc_major, int *cc_minor, size_t *smem_optin) {
    if (!ctx) return VCS_E_INVALID;
    if (sm_count) *sm_count = ctx->sm_count;
    if (cc_major) *cc_major = ctx->cc_major;
    if (cc_minor) *cc_minor = ctx->cc_minor;
    if (smem_optin) *smem_optin = ctx->smem_optin;
    return VCS_OK;
}

int64_t vcs_launch_count(const vcs_ctx *ctx) { return ctx ? ctx->launches : 0; }

int vcs_set_q(vcs_ctx *ctx, const double *Q) {
    if (!ctx || !Q) return VCS_E_INVALID;
    for (int k = 0; k < 192; ++k)
        if (!(Q[k] != 0.0)) return fail(ctx, VCS_E_INVALID, "Q[%d] is zero or NaN", k);
    if (memcmp(ctx->h_Q, Q, sizeof(ctx->h_Q)) == 0) return VCS_OK;   // unchanged: nothing to upload
    ctx->q_version += 1;
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(ctx->h_Q, Q, sizeof(ctx->h_Q));
    CK(ctx, cudaMemcpy(ctx->d_Q, ctx->h_Q, sizeof(ctx->h_Q), cudaMemcpyHostToDevice));
    return VCS_OK;
}

// ---- motion estimation -------------------------------------------------------------------
int vcs_me_search_dev(vcs_ctx *ctx, const vcs_me_params *p, const uint8_t *cur, const uint8_t *ref,
                      int16_t *mv, uint32_t *cost, uint8_t *flags) {
    VCS_ENTER(ctx);
    if (!cur || !ref) return fail(ctx, VCS_E_INVALID, "frame pointer is NULL");
    return launch_me(ctx, ctx->stream, p, pair_addr(cur, ref), 1, mv, cost, flags);
}

int vcs_me_search_clip_dev(vcs_ctx *ctx, const vcs_me_params *p, const uint8_t *frames, int T,
                           int gop_len, int16_t *mv, uint32_t *cost, uint8_t *flags) {
    VCS_ENTER(ctx);
    if (!frames || !p || gop_len < 2 || T < 1) return fail(ctx, VCS_E_INVALID, "bad clip arguments");
    return launch_me(ctx, ctx->stream, p, clip_addr(frames, p->H, p->W, gop_len),
                     vcs_num_p_frames(T, gop_len), mv, cost, flags);
}

int vcs_me_search_host(vcs_ctx *ctx, const vcs_me_params *p, const uint8_t *cur, const uint8_t *ref,
                       int16_t *mv, uint32_t *cost, uint8_t *flags) {
    VCS_ENTER(ctx);
    int rc = check_me_params(ctx, p);
    if (rc) return rc;
    if (!cur || !ref || !mv) return fail(ctx, VCS_E_INVALID, "NULL argument");
    const size_t fs = (size_t)p->H * p->W * 3;
    const int N = vcs_num_blocks(p->H, p->W, p->bs);
    uint8_t *d_fr; int16_t *d_mv; uint32_t *d_cost; uint8_t *d_fl;
    if ((rc = dev_buf(ctx, S_FRAMES, 2 * fs, (void **)&d_fr))) return rc;
    if ((rc = dev_buf(ctx, S_MV, (size_t)N * 4, (void **)&d_mv))) return rc;
    if ((rc = dev_buf(ctx, S_COST, (size_t)N * 4, (void **)&d_cost))) return rc;
    if ((rc = dev_buf(ctx, S_FLAGS, (size_t)N, (void **)&d_fl))) return rc;
    cudaStream_t st = ctx->stream;
    CK(ctx, cudaMemcpyAsync(d_fr, ref, fs, cudaMemcpyHostToDevice, st));
    CK(ctx, cudaMemcpyAsync(d_fr + fs, cur, fs, cudaMemcpyHostToDevice, st));
    if ((rc = launch_me(ctx, st, p, pair_addr(d_fr + fs, d_fr), 1, d_mv, d_cost, d_fl))) return rc;
    CK(ctx, cudaMemcpyAsync(mv, d_mv, (size_t)N * 4, cudaMemcpyDeviceToHost, st));
    if (cost) CK(ctx, cudaMemcpyAsync(cost, d_cost, (size_t)N * 4, cudaMemcpyDeviceToHost, st));
    if (flags) CK(ctx, cudaMemcpyAsync(flags, d_fl, (size_t)N, cudaMemcpyDeviceToHost, st));
    CK(ctx, cudaStreamSynchronize(st));
    return VCS_OK;
}

// ---- MC / wrap arithmetic ------------------------------------------------------------------
int vcs_mc_dev(vcs_ctx *ctx, int H, int W, int bs, const uint8_t *ref, const int16_t *mv, uint8_t *pred) {
    VCS_ENTER(ctx);
    if (!ref || !mv || !pred || bs <= 0 || H < bs || W < bs) return fail(ctx, VCS_E_INVALID, "bad MC arguments");
    const size_t npix = (size_t)H * W;
    int blocks = (int)((npix + 255) / 256);
    if (blocks > ctx->sm_count * 16) blocks = ctx->sm_count * 16;
    mc_kernel<<<blocks, 256, 0, ctx->stream>>>(ref, mv, H, W, bs, W / bs, H / bs, pred, ctx->h_errflag);
    CK(ctx, cudaGetLastError());
    ctx->launches += 1;
    return VCS_OK;
}

int vcs_mc_host(vcs_ctx *ctx, int H, int W, int bs, const uint8_t *ref, const int16_t *mv, uint8_t *pred) {
    VCS_ENTER(ctx);
    if (!ref || !mv || !pred || bs <= 0 || H < bs || W < bs) return fail(ctx, VCS_E_INVALID, "bad MC arguments");
    const size_t fs = (size_t)H * W * 3;
    const int N = vcs_num_blocks(H, W, bs);
    // the reference would raise IndexError / broadcast errors on an out-of-frame vector
    for (int k = 0; k < N; ++k) {
        int x = (k % (W / bs)) * bs + mv[2 * k], y = (k / (W / bs)) * bs + mv[2 * k + 1];
        if (x < 0 || y < 0 || x + bs > W || y + bs > H)
            return fail(ctx, VCS_E_INVALID, "motion vector %d points outside the frame", k);
    }
    uint8_t *d_fr; int16_t *d_mv; int rc;
    if ((rc = dev_buf(ctx, S_FRAMES, 2 * fs, (void **)&d_fr))) return rc;
    if ((rc = dev_buf(ctx, S_MV, (size_t)N * 4, (void **)&d_mv))) return rc;
    cudaStream_t st = ctx->stream;
    CK(ctx, cudaMemcpyAsync(d_fr, ref, fs, cudaMemcpyHostToDevice, st));
    CK(ctx, cudaMemcpyAsync(d_mv, mv, (size_t)N * 4, cudaMemcpyHostToDevice, st));
    if ((rc = vcs_mc_dev(ctx, H, W, bs, d_fr, d_mv, d_fr + fs))) return rc;
    CK(ctx, cudaMemcpyAsync(pred, d_fr + fs, fs, cudaMemcpyDeviceToHost, st));
    CK(ctx, cudaStreamSynchronize(st));
    return VCS_OK;
}

static int wrap_dev(vcs_ctx *ctx, int add, const uint8_t *a, const uint8_t *b, size_t n, uint8_t *out) {
    VCS_ENTER(ctx);
    if (!a || !b || !out) return fail(ctx, VCS_E_INVALID, "NULL argument");
    if (n == 0) return VCS_OK;
    int blocks = (int)((n + 255) / 256);
    if (blocks > ctx->sm_count * 16) blocks = ctx->sm_count * 16;
    if (add) wrap_kernel<1><<<blocks, 256, 0, ctx->stream>>>(a, b, n, out);
    else wrap_kernel<0><<<blocks, 256, 0, ctx->stream>>>(a, b, n, out);
    CK(ctx, cudaGetLastError());
    ctx->launches += 1;
    return VCS_OK;
}

static int wrap_host(vcs_ctx *ctx, int add, const uint8_t *a, const uint8_t *b, size_t n, uint8_t *out) {
    VCS_ENTER(ctx);
    if (!a || !b || !out) return fail(ctx, VCS_E_INVALID, "NULL argument");
    if (n == 0) return VCS_OK;
    uint8_t *d; int rc;
    if ((rc = dev_buf(ctx, S_FRAMES, 3 * n, (void **)&d))) return rc;
    cudaStream_t st = ctx->stream;
    CK(ctx, cudaMemcpyAsync(d, a, n, cudaMemcpyHostToDevice, st));
    CK(ctx, cudaMemcpyAsync(d + n, b, n, cudaMemcpyHostToDevice, st));
    if ((rc = wrap_dev(ctx, add, d, d + n, n, d + 2 * n))) return rc;
    CK(ctx, cudaMemcpyAsync(out, d + 2 * n, n, cudaMemcpyDeviceToHost, st));
    CK(ctx, cudaStreamSynchronize(st));
    return VCS_OK;
}

int vcs_sub_wrap_dev(vcs_ctx *c, const uint8_t *a, const uint8_t *b, size_t n, uint8_t *o) { return wrap_dev(c, 0, a, b, n, o); }
int vcs_add_wrap_dev(vcs_ctx *c, const uint8_t *a, const uint8_t *b, size_t n, uint8_t *o) { return wrap_dev(c, 1, a, b, n, o); }
int vcs_sub_wrap_host(vcs_ctx *c, const uint8_t *a, const uint8_t *b, size_t n, uint8_t *o) { return wrap_host(c, 0, a, b, n, o); }
int vcs_add_wrap_host(vcs_ctx *c, const uint8_t *a, const uint8_t *b, size_t n, uint8_t *o) { return wrap_host(c, 1, a, b, n, o); }

// ---- DCT stage -------------------------------------------------------------------------------
int vcs_compress_dev(vcs_ctx *ctx, int H, int W, const uint8_t *bgr, int coef_mode, void *coef) {
    VCS_ENTER(ctx);
    if (!bgr || !coef) return fail(ctx, VCS_E_INVALID, "NULL argument");
    DctArgs a{};
    a.H = H; a.W = W; a.img = bgr; a.bs = 8; a.forward = 1; a.coef_mode = coef_mode; a.coef = coef;
    return launch_dct(ctx, ctx->stream, a, 1);
}

int vcs_compress_host(vcs_ctx *ctx, int H, int W, const uint8_t *bgr, int coef_mode, void *coef) {
    VCS_ENTER(ctx);
    if (!bgr || !coef) return fail(ctx, VCS_E_INVALID, "NULL argument");
    if (H <= 0 || W <= 0 || H % 8 || W % 8 || coef_mode < 0 || coef_mode > 3)
        return fail(ctx, VCS_E_INVALID, "H=%d W=%d must be multiples of 8; coef_mode=%d", H, W, coef_mode);
    const size_t npix = (size_t)H * W;
    uint8_t *d_img; void *d_coef; int rc;
    if ((rc = dev_buf(ctx, S_FRAMES, npix * 3, (void **)&d_img))) return rc;
    if ((rc = dev_buf(ctx, S_COEF, npix * 3 * coef_elem(coef_mode), &d_coef))) return rc;
    cudaStream_t st = ctx->stream;
    CK(ctx, cudaMemcpyAsync(d_img, bgr, npix * 3, cudaMemcpyHostToDevice, st));
    if ((rc = vcs_compress_dev(ctx, H, W, d_img, coef_mode, d_coef))) return rc;
    CK(ctx, cudaMemcpyAsync(coef, d_coef, npix * 3 * coef_elem(coef_mode), cudaMemcpyDeviceToHost, st));
    CK(ctx, cudaStreamSynchronize(st));
    return VCS_OK;
}

int vcs_decompress_dev(vcs_ctx *ctx, int H, int W, int coef_mode, const void *coef, const uint8_t *pred,
                       uint8_t *bgr) {
    VCS_ENTER(ctx);
    if (!coef || !bgr) return fail(ctx, VCS_E_INVALID, "NULL argument");
    DctArgs a{};
    a.H = H; a.W = W; a.bs = 8; a.forward = 0; a.inverse = 1; a.coef_mode = coef_mode;
    a.coef = const_cast<void *>(coef); a.pred_in = pred; a.recon = bgr;
    return launch_dct(ctx, ctx->stream, a, 1);
}

int vcs_decompress_host(vcs_ctx *ctx, int H, int W, int coef_mode, const void *coef, const uint8_t *pred,
                        uint8_t *bgr) {
    VCS_ENTER(ctx);
    if (!coef || !bgr) return fail(ctx, VCS_E_INVALID, "NULL argument");
    if (H <= 0 || W <= 0 || H % 8 || W % 8 || coef_mode < 0 || coef_mode > 3)
        return fail(ctx, VCS_E_INVALID, "H=%d W=%d must be multiples of 8; coef_mode=%d", H, W, coef_mode);
    const size_t npix = (size_t)H * W;
    uint8_t *d_img; void *d_coef; int rc;
    if ((rc = dev_buf(ctx, S_FRAMES, npix * 6, (void **)&d_img))) return rc;
    if ((rc = dev_buf(ctx, S_COEF, npix * 3 * coef_elem(coef_mode), &d_coef))) return rc;
    cudaStream_t st = ctx->stream;
    CK(ctx, cudaMemcpyAsync(d_coef, coef, npix * 3 * coef_elem(coef_mode), cudaMemcpyHostToDevice, st));
    if (pred) CK(ctx, cudaMemcpyAsync(d_img + npix * 3, pred, npix * 3, cudaMemcpyHostToDevice, st));
    if ((rc = vcs_decompress_dev(ctx, H, W, coef_mode, d_coef, pred ? d_img + npix * 3 : nullptr, d_img)))
        return rc;
    CK(ctx, cudaMemcpyAsync(bgr, d_img, npix * 3, cudaMemcpyDeviceToHost, st));
    CK(ctx, cudaStreamSynchronize(st));
    return VCS_OK;
}

// DCTCompressor._dct2 / _idct2 (DCTcompressor.py:111-121) on n bare 8x8 float64 blocks (host buffers)
int vcs_dct2_blocks_host(vcs_ctx *ctx, int nblocks, int inverse, const double *in, double *out) {
    VCS_ENTER(ctx);
    if (!in || !out || nblocks < 0) return fail(ctx, VCS_E_INVALID, "bad dct2 arguments");
    if (nblocks == 0) return VCS_OK;
    const size_t bytes = (size_t)nblocks * 64 * sizeof(double);
    double *d; int rc;
    if ((rc = dev_buf(ctx, S_COEF, 2 * bytes, (void **)&d))) return rc;
    cudaStream_t st = ctx->stream;
    CK(ctx, cudaMemcpyAsync(d, in, bytes, cudaMemcpyHostToDevice, st));
    int blocks = (nblocks + 3) / 4;
    if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
    dct2_blocks_kernel<<<blocks, 256, 0, st>>>(d, nblocks, inverse != 0, d + (size_t)nblocks * 64);
    CK(ctx, cudaGetLastError());
    ctx->launches += 1;
    CK(ctx, cudaMemcpyAsync(out, d + (size_t)nblocks * 64, bytes, cudaMemcpyDeviceToHost, st));
    CK(ctx, cudaStreamSynchronize(st));
    return VCS_OK;
}

// ---- fused clip paths ------------------------------------------------------------------------
int vcs_residual_dct_clip_dev(vcs_ctx *ctx, int H, int W, int bs, const uint8_t *frames, int T, int gop_len,
                              const int16_t *mv, int coef_mode, void *coef, uint8_t *recon) {
    VCS_ENTER(ctx);
    if (!frames || !mv || gop_len < 2 || T < 1 || bs <= 0 || H < bs || W < bs)
        return fail(ctx, VCS_E_INVALID, "bad clip arguments");
    DctArgs a{};
    a.H = H; a.W = W; a.fa = clip_addr(frames, H, W, gop_len); a.has_fa = 1; a.mv = mv; a.bs = bs;
    a.nbx = W / bs; a.nby = H / bs; a.forward = 1; a.inverse = recon != nullptr;
    a.coef_mode = coef_mode; a.coef = coef; a.recon = recon;
    return launch_dct(ctx, ctx->stream, a, vcs_num_p_frames(T, gop_len));
}

// Decoder._reconstruct_P_frame over a clip (decoder.py:52-69): MC from the ORIGINAL I-frames + dequantise +
// IDCT + truncating store + YCrCb->BGR + wrap add, one launch.
int vcs_decode_clip_dev(vcs_ctx *ctx, int H, int W, int bs, const uint8_t *ref_frames, int T, int gop_len,
                        const int16_t *mv, int coef_mode, const void *coef, uint8_t *recon) {
    VCS_ENTER(ctx);
    if (!ref_frames || !mv || !coef || !recon || gop_len < 2 || T < 1 || bs <= 0 || H < bs || W < bs)
        return fail(ctx, VCS_E_INVALID, "bad decode arguments");
    const long long fs = (long long)H * W * 3;
    DctArgs a{};
    a.H = H; a.W = W; a.has_fa = 1; a.mv = mv; a.bs = bs; a.nbx = W / bs; a.nby = H / bs;
    a.fa.cur_base = ref_frames; a.fa.ref_base = ref_frames;      // only the I-frames are given: [nG][H][W][3]
    a.fa.cur_gop_stride = fs; a.fa.cur_frame_stride = 0; a.fa.ref_gop_stride = fs; a.fa.ppg = gop_len - 1;
    a.forward = 0; a.inverse = 1; a.coef_mode = coef_mode; a.coef = const_cast<void *>(coef); a.recon = recon;
    return launch_dct(ctx, ctx->stream, a, vcs_num_p_frames(T, gop_len));
}

int vcs_decode_clip_host(vcs_ctx *ctx, int H, int W, int bs, const uint8_t *ref_frames, int T, int gop_len,
                         const int16_t *mv, int coef_mode, const void *coef, uint8_t *recon) {
    VCS_ENTER(ctx);
    if (!ref_frames || !mv || !coef || !recon || gop_len < 2 || T < 1 || bs <= 0 || H < bs || W < bs)
        return fail(ctx, VCS_E_INVALID, "bad decode arguments");
    if (H % 8 || W % 8 || coef_mode < 0 || coef_mode > 3)
        return fail(ctx, VCS_E_INVALID, "H=%d W=%d must be multiples of 8; coef_mode=%d", H, W, coef_mode);
    const size_t fs = (size_t)H * W * 3, npix = (size_t)H * W, ce = coef_elem(coef_mode);
    const int N = vcs_num_blocks(H, W, bs), nP = vcs_num_p_frames(T, gop_len), nG = (T + gop_len - 1) / gop_len;
    // the reference would fail on a vector pointing outside the frame (motion.py:62-65)
    for (long long k = 0; k < (long long)nP * N; ++k) {
        const int b = (int)(k % N);
        const int x = (b % (W / bs)) * bs + mv[2 * k], y = (b / (W / bs)) * bs + mv[2 * k + 1];
        if (x < 0 || y < 0 || x + bs > W || y + bs > H)
            return fail(ctx, VCS_E_INVALID, "motion vector %lld points outside the frame", k);
    }
    uint8_t *d_ref, *d_rec; int16_t *d_mv; void *d_coef; int rc;
    if ((rc = dev_buf(ctx, S_FRAMES, fs * nG, (void **)&d_ref))) return rc;
    if ((rc = dev_buf(ctx, S_MV, (size_t)nP * N * 4 + 4, (void **)&d_mv))) return rc;
    if ((rc = dev_buf(ctx, S_COEF, (size_t)nP * npix * 3 * ce + 8, &d_coef))) return rc;
    if ((rc = dev_buf(ctx, S_RECON, (size_t)nP * fs + 4, (void **)&d_rec))) return rc;
    cudaStream_t st = ctx->stream;
    CK(ctx, cudaMemcpyAsync(d_ref, ref_frames, fs * nG, cudaMemcpyHostToDevice, st));
    CK(ctx, cudaMemcpyAsync(d_mv, mv, (size_t)nP * N * 4, cudaMemcpyHostToDevice, st));
    CK(ctx, cudaMemcpyAsync(d_coef, coef, (size_t)nP * npix * 3 * ce, cudaMemcpyHostToDevice, st));
    if ((rc = vcs_decode_clip_dev(ctx, H, W, bs, d_ref, T, gop_len, d_mv, coef_mode, d_coef, d_rec))) return rc;
    CK(ctx, cudaMemcpyAsync(recon, d_rec, (size_t)nP * fs, cudaMemcpyDeviceToHost, st));
    CK(ctx, cudaStreamSynchronize(st));
    return VCS_OK;
}

// numerator of dct.py:188-191's sparsity: number of non-zero coefficients in n elements (device pointer)
int vcs_count_nonzero_dev(vcs_ctx *ctx, int coef_mode, const void *coef, size_t n, unsigned long long *count_host) {
    VCS_ENTER(ctx);
    if (!coef || !count_host || coef_mode < 0 || coef_mode > 3) return fail(ctx, VCS_E_INVALID, "bad arguments");
    unsigned long long *d_cnt; int rc;
    if ((rc = dev_buf(ctx, S_CYC, 8, (void **)&d_cnt))) return rc;
    cudaStream_t st = ctx->stream;
    CK(ctx, cudaMemsetAsync(d_cnt, 0, 8, st));
    if (n) {
        int blocks = (int)((n + 1023) / 1024);
        if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
        if (coef_mode == VCS_COEF_I8_RINT) count_nonzero_kernel<1><<<blocks, 256, 0, st>>>(coef, n, d_cnt);
        else if (coef_mode == VCS_COEF_I16_RINT) count_nonzero_kernel<2><<<blocks, 256, 0, st>>>(coef, n, d_cnt);
        else count_nonzero_kernel<8><<<blocks, 256, 0, st>>>(coef, n, d_cnt);
        CK(ctx, cudaGetLastError());
        ctx->launches += 1;
    }
    CK(ctx, cudaMemcpyAsync(count_host, d_cnt, 8, cudaMemcpyDeviceToHost, st));
    CK(ctx, cudaStreamSynchronize(st));
    return VCS_OK;
}

int vcs_set_dct_precision(vcs_ctx *ctx, int bits) {
    VCS_ENTER(ctx);
    if (bits != 32 && bits != 64) return fail(ctx, VCS_E_INVALID, "DCT precision is 64 (exact tier) or 32 (fp32 tier), not %d", bits);
    ctx->dct_fp32 = bits == 32;
    return VCS_OK;
}

int vcs_flip_counters_dev(vcs_ctx *ctx, int coef_mode, const void *coef_a, const void *coef_b, size_t ncoef,
                          const uint8_t *px_a, const uint8_t *px_b, size_t npx, unsigned long long *out3_host) {
    VCS_ENTER(ctx);
    if (!out3_host || (coef_mode != VCS_COEF_I8_RINT && coef_mode != VCS_COEF_I16_RINT) || (ncoef && (!coef_a || !coef_b)) ||
        (npx && (!px_a || !px_b)))
        return fail(ctx, VCS_E_INVALID, "bad arguments (flip counters compare int8 or int16 index planes and uint8 frames)");
    unsigned long long *d_cnt; int rc;
    if ((rc = dev_buf(ctx, S_CYC, 24, (void **)&d_cnt))) return rc;
    cudaStream_t st = ctx->stream;
    CK(ctx, cudaMemsetAsync(d_cnt, 0, 24, st));
    const int cap = ctx->sm_count * 8;
    if (ncoef) {
        int blocks = (int)((ncoef + 1023) / 1024);
        if (coef_mode == VCS_COEF_I8_RINT) flip_count_kernel<int8_t><<<blocks < cap ? blocks : cap, 256, 0, st>>>((const int8_t *)coef_a, (const int8_t *)coef_b, ncoef, d_cnt, nullptr);
        else flip_count_kernel<int16_t><<<blocks < cap ? blocks : cap, 256, 0, st>>>((const int16_t *)coef_a, (const int16_t *)coef_b, ncoef, d_cnt, nullptr);
        ctx->launches += 1;
    }
    if (npx) {
        int blocks = (int)((npx + 1023) / 1024);
        flip_count_kernel<uint8_t><<<blocks < cap ? blocks : cap, 256, 0, st>>>(px_a, px_b, npx, d_cnt + 1, d_cnt + 2);
        ctx->launches += 1;
    }
    CK(ctx, cudaGetLastError());
    CK(ctx, cudaMemcpyAsync(out3_host, d_cnt, 24, cudaMemcpyDeviceToHost, st));
    CK(ctx, cudaStreamSynchronize(st));
    return VCS_OK;
}

int vcs_encode_clip_dev(vcs_ctx *ctx, const vcs_me_params *p, const uint8_t *frames, int T, int gop_len,
                        int coef_mode, int16_t *mv, uint32_t *cost, uint8_t *flags, void *coef,
                        uint8_t *recon) {
    VCS_ENTER(ctx);
    if (!frames || !p || gop_len < 2 || T < 1) return fail(ctx, VCS_E_INVALID, "bad clip arguments");
    return encode_dev(ctx, ctx->stream, p, clip_addr(frames, p->H, p->W, gop_len),
                      vcs_num_p_frames(T, gop_len), coef_mode, mv, cost, flags, coef, recon);
}

static int encode_clip_host_impl(vcs_ctx *ctx, const vcs_me_params *p, const uint8_t *frames, int T, int gop_len,
                                 int coef_mode, int16_t *mv, uint32_t *cost, uint8_t *flags, void *coef,
                                 uint8_t *recon, const PackedHost *pk) {
    VCS_ENTER(ctx);
    int rc = check_me_params(ctx, p);
    if (rc) return rc;
    if (!frames || gop_len < 2 || T < 1) return fail(ctx, VCS_E_INVALID, "bad clip arguments");
    if ((coef || recon) && (p->H % 8 || p->W % 8))
        return fail(ctx, VCS_E_INVALID, "H=%d W=%d must be multiples of 8 for the DCT stage", p->H, p->W);
    const size_t fs = (size_t)p->H * p->W * 3, npix = (size_t)p->H * p->W;
    const int N = vcs_num_blocks(p->H, p->W, p->bs);
    const int nP = vcs_num_p_frames(T, gop_len);
    const size_t ce = coef_elem(coef_mode);
    uint8_t *d_fr; int16_t *d_mv; uint32_t *d_cost; uint8_t *d_fl; void *d_coef = nullptr; uint8_t *d_rec = nullptr;
    if ((rc = dev_buf(ctx, S_FRAMES, fs * T, (void **)&d_fr))) return rc;
    if ((rc = dev_buf(ctx, S_MV, (size_t)nP * N * 4 + 4, (void **)&d_mv))) return rc;
    if ((rc = dev_buf(ctx, S_COST, (size_t)nP * N * 4 + 4, (void **)&d_cost))) return rc;
    if ((rc = dev_buf(ctx, S_FLAGS, (size_t)nP * N + 4, (void **)&d_fl))) return rc;
    if ((coef || pk) && (rc = dev_buf(ctx, S_COEF, (size_t)nP * npix * 3 * ce + 8, &d_coef))) return rc;
    if (recon && (rc = dev_buf(ctx, S_RECON, (size_t)nP * fs + 4, (void **)&d_rec))) return rc;
    // packed sink: the dense int8 planes stay on the device; bitmaps, row counts and the two streams travel
    const int rows_per_p = 3 * (p->H / 8), nbx8 = p->W / 8;
    PackedDev pd;
    memset(&pd, 0, sizeof(pd));

    // Pipeline: copy-in on s_h2d, kernels on the compute stream, copy-out on s_d2h, chained with events; PCIe
    // is full duplex so the three overlap.  Segments are ranges of P-frames (they may start and end inside a GOP):
    // only the first upload and the last download are exposed, so the schedule starts and ends with one P-frame
    // and ramps in between (a segment's upload has to fit under the previous segment's kernels).  A search
    // launch wastes its last partial wave of tiles (every CTA owns an SM for ~55 us per tile at 1080p), so each
    // size is nudged by +-1 towards a whole number of waves.  VCS_PIPELINE_P="1,2,4,8" overrides the sizes (the
    // last value repeats).
    const int ppg = gop_len - 1;
    std::vector<int> sizes;
    {
        std::vector<int> sched;
        if (const char *e = getenv("VCS_PIPELINE_P")) {
            for (const char *q = e; *q;) {
                char *end; long v = strtol(q, &end, 10);
                if (end == q) break;
                if (v > 0) sched.push_back((int)v);
                q = *end ? end + 1 : end;
            }
        }
        if (!sched.empty()) {
            for (size_t i = 0, left = (size_t)nP; left > 0; ++i) {
                size_t n = (size_t)sched[i < sched.size() ? i : sched.size() - 1];
                if (n > left) n = left;
                sizes.push_back((int)n);
                left -= n;
            }
        } else {
            const double waves1 = (double)(((p->W / p->bs) + 4) / 5) * (((p->H / p->bs) + 2) / 3) / ctx->sm_count;
            auto waste = [&](int n) { const double w = n * waves1; return (ceil(w - 1e-9) - w) / w; };
            auto nudge = [&](int n, int cap) {          // n-1, n or n+1, whichever wastes least of its last wave
                int best = n;
                for (int c = n - 1; c <= n + 1; c += 2)
                    if (c >= 1 && c <= cap && waste(c) < waste(best) - 1e-9) best = c;
                return best;
            };
            const int head[5] = {1, 2, 3, 4, 6}, tail[3] = {4, 2, 1};
            int left = nP, tail_sum = 0;
            std::vector<int> back;
            if (nP >= 12) for (int k = 2; k >= 0; --k) { back.push_back(tail[k]); tail_sum += tail[k]; }   // 1, 2, 4
            left -= tail_sum;
            for (int k = 0; left > 0; ++k) {
                int n = k < 5 ? head[k] : 8;
                // the ramp is paced by the upload (a segment's frames must have arrived when the previous segment's
                // kernels end: 1,2,4 stalls the third segment for 0.16 ms at 1080p where 1,2,3 does not), so only the
                // full-size segments are nudged
                if (n > left) n = left; else if (k >= 5) { n = nudge(n, left); }
                sizes.push_back(n);
                left -= n;
            }
            for (size_t k = back.size(); k-- > 0;) sizes.push_back(back[k]);
        }
    }
    const int nsegs = (int)sizes.size();
    if (pk) {
        if ((size_t)nsegs > ctx->h_segend_cap) {
            if (ctx->h_segend) cudaFreeHost(ctx->h_segend);
            ctx->h_segend = nullptr; ctx->h_segend_cap = 0;
            CK(ctx, cudaHostAlloc((void **)&ctx->h_segend, 2 * sizeof(unsigned long long) * (size_t)nsegs, cudaHostAllocDefault));
            ctx->h_segend_cap = (size_t)nsegs;
        }
        while ((int)ctx->seg_events.size() < nsegs) {
            cudaEvent_t e;
            CK(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ctx->seg_events.push_back(e);
        }
        if ((rc = packed_scratch(ctx, (size_t)nP * rows_per_p, p->W, (size_t)nP * npix * 3, nsegs, pd))) return rc;
        CK(ctx, cudaMemsetAsync(pd.totals, 0, 16, ctx->stream));   // ordered before the pipeline's streams (see below)
        pk->lengths[0] = pk->lengths[1] = 0;
    }
    while ((int)ctx->chunk_events.size() < 2 * nsegs) {
        cudaEvent_t e;
        CK(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->chunk_events.push_back(e);
    }
    // VCS_TRACE=1: a timeline of the pipeline on stderr (timing-enabled events; diagnostics only)
    const bool trace = getenv("VCS_TRACE") != nullptr;
    std::vector<cudaEvent_t> tev;          // [0]: start of the call; then 4 per segment: upload done, compute may start, compute done, download done
    auto tmark = [&](cudaStream_t st, int seg = -1, int what = 0) {
        if (!trace) return;
        cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st);
        if (seg < 0) { tev.push_back(e); return; }
        const size_t k = 1 + 4 * (size_t)seg + (size_t)what;
        if (tev.size() <= k) tev.resize(k + 1, nullptr);
        tev[k] = e;
    };
    std::vector<double> host_ms;          // when the host had queued each segment's kernels, on the same time base
    std::chrono::steady_clock::time_point host_t0;
    if (trace) {
        cudaStreamSynchronize(ctx->stream); tmark(ctx->stream); cudaStreamWaitEvent(ctx->s_h2d, tev[0], 0); cudaStreamWaitEvent(ctx->s_d2h, tev[0], 0);
        host_t0 = std::chrono::steady_clock::now();
    }
    // (Running consecutive searches on two streams so that one fills the other's tail was tried and is slower:
    // the persistent search CTAs of the next chunk then keep the DCT stage of this chunk off the SMs.)
    // Optional two-stream schedule (VCS_TILES_PER_CTA=k > 0): searches back to back on a low-priority stream with k
    // tiles per CTA instead of persistent CTAs, each segment's DCT stage and packing on a high-priority stream behind
    // its search, so that the next search fills the SMs the current one's last wave leaves idle.  Measured slower than
    // the default single stream with persistent CTAs (k = 1/2/4/8: 10.98/11.21/11.59/11.91 ms against 10.95 ms dense):
    // a DCT stage that becomes ready has to wait for running search CTAs to retire, and CTAs that own few tiles lose
    // the cross-tile TMA prefetch.
    int tiles_per_cta = 0;
    if (const char *e = getenv("VCS_TILES_PER_CTA")) tiles_per_cta = atoi(e);
    const bool two_streams = tiles_per_cta > 0;
    cudaStream_t sc = two_streams ? ctx->s_search : ctx->stream;
    cudaStream_t sd = two_streams ? ctx->s_dct : ctx->stream;
    // VCS_PACK_ASIDE=0 keeps the packing kernels on the compute stream (11.14 ms per bench clip against 10.83 ms on the side stream)
    const bool pack_aside = pk && !two_streams && !(getenv("VCS_PACK_ASIDE") && atoi(getenv("VCS_PACK_ASIDE")) == 0);
    // (The DCT stage itself on a side stream, overlapping the next search's last wave with the search after that waiting for
    // it, measured slower: 10.95 ms against 10.55 ms dense, 10.94 against 10.81 packed.)
    while ((two_streams || pack_aside) && (int)ctx->me_events.size() < nsegs) {
        cudaEvent_t e;
        CK(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->me_events.push_back(e);
    }
    struct TilesGuard {          // the device-resident entry points keep persistent CTAs
        MeTiledState &st; int saved;
        TilesGuard(MeTiledState &s, int k) : st(s), saved(s.tiles_per_cta) { st.tiles_per_cta = k; }
        ~TilesGuard() { st.tiles_per_cta = saved; }
    } tiles_guard(ctx->tiled, two_streams ? tiles_per_cta : 0);
    if (two_streams) {           // work queued on the context's stream before this call comes first
        CK(ctx, cudaEventRecord(ctx->chunk_events[0], ctx->stream));
        CK(ctx, cudaStreamWaitEvent(sc, ctx->chunk_events[0], 0));
        CK(ctx, cudaStreamWaitEvent(sd, ctx->chunk_events[0], 0));
    }
    // Pass 1 queues everything that does not depend on what the host knows: uploads, kernels and -- without the packed
    // sink -- the downloads.  With the packed sink the size of a segment's value stream is only known once its scan
    // kernel has run, so each segment just sends that length (8 bytes, own stream: it must not queue behind the
    // uploads or the other downloads) and pass 2 queues a segment's downloads as soon as its length has landed.
    auto download = [&](int p0, int np) -> int {
        if (mv) CK(ctx, cudaMemcpyAsync(mv + (size_t)p0 * N * 2, d_mv + (size_t)p0 * N * 2, (size_t)np * N * 4,
                                        cudaMemcpyDeviceToHost, ctx->s_d2h));
        if (cost) CK(ctx, cudaMemcpyAsync(cost + (size_t)p0 * N, d_cost + (size_t)p0 * N, (size_t)np * N * 4,
                                          cudaMemcpyDeviceToHost, ctx->s_d2h));
        if (flags) CK(ctx, cudaMemcpyAsync(flags + (size_t)p0 * N, d_fl + (size_t)p0 * N, (size_t)np * N,
                                           cudaMemcpyDeviceToHost, ctx->s_d2h));
        if (coef) CK(ctx, cudaMemcpyAsync((char *)coef + (size_t)p0 * npix * 3 * ce,
                                          (char *)d_coef + (size_t)p0 * npix * 3 * ce,
                                          (size_t)np * npix * 3 * ce, cudaMemcpyDeviceToHost, ctx->s_d2h));
        if (recon) CK(ctx, cudaMemcpyAsync(recon + (size_t)p0 * fs, d_rec + (size_t)p0 * fs, (size_t)np * fs,
                                           cudaMemcpyDeviceToHost, ctx->s_d2h));
        return VCS_OK;
    };
    auto pipeline = [&]() -> int {
    int uploaded = 0, p0 = 0;
    // frames [0, need(c)) must be resident for segment c; frames after the last P-frame are never needed on the device
    std::vector<int> need(nsegs);
    for (int c = 0, q = 0; c < nsegs; q += sizes[c], ++c) {
        const int plast = q + sizes[c] - 1;
        need[c] = (plast / ppg) * gop_len + 1 + plast % ppg + 1;
    }
    std::vector<char> has_upload(nsegs, 0);
    auto upload = [&](int c) -> int {           // queue segment c's frames; chunk_events[2c] = they have arrived
        if (need[c] > uploaded) {
            CK(ctx, cudaMemcpyAsync(d_fr + fs * uploaded, frames + fs * uploaded, fs * (need[c] - uploaded),
                                    cudaMemcpyHostToDevice, ctx->s_h2d));
            CK(ctx, cudaEventRecord(ctx->chunk_events[2 * c], ctx->s_h2d));
            uploaded = need[c];
            has_upload[c] = 1;
        }
        tmark(ctx->s_h2d, c, 0);
        return VCS_OK;
    };
    // (Waiting for segment c+1's frames between the search and the DCT stage of segment c, so that the cross-stream wait
    // hides behind a running search, was tried against the ~40 us bubble between the first segments: no change.)
    for (int c = 0; c < nsegs; ++c) {
        const int np = sizes[c];
        if ((rc = upload(c))) return rc;
        if (has_upload[c]) CK(ctx, cudaStreamWaitEvent(sc, ctx->chunk_events[2 * c], 0));
        tmark(sc, c, 1);
        cudaStream_t sdone = sd;   // the stream whose progress makes segment c ready for download
        {
            const int g0 = p0 / ppg;
            FrameAddr fa = clip_addr(d_fr + fs * (size_t)g0 * gop_len, p->H, p->W, gop_len);
            fa.p_off = p0 - g0 * ppg;
            rc = encode_dev(ctx, sc, p, fa, np, coef_mode, d_mv + (size_t)p0 * N * 2, d_cost + (size_t)p0 * N,
                            d_fl + (size_t)p0 * N,
                            d_coef ? (void *)((char *)d_coef + (size_t)p0 * npix * 3 * ce) : nullptr,
                            d_rec ? d_rec + (size_t)p0 * fs : nullptr,
                            pk ? pd.bitmap + (size_t)p0 * rows_per_p * nbx8 : nullptr,
                            pk ? pd.blk_esc + (size_t)p0 * rows_per_p * nbx8 : nullptr,
                            pk ? pd.row_count + (size_t)p0 * rows_per_p : nullptr,
                            two_streams ? sd : nullptr, two_streams ? ctx->me_events[c] : nullptr);
            if (rc) return rc;
            // Packing is not on the compute chain (only this segment's download needs it): on its own stream it runs in
            // the SMs the NEXT segment's search leaves idle in its last partial wave instead of delaying that search.
            if (pk && pack_aside) {
                CK(ctx, cudaEventRecord(ctx->me_events[c], sd));
                CK(ctx, cudaStreamWaitEvent(ctx->s_dct, ctx->me_events[c], 0));
                sdone = ctx->s_dct;
            }
            if (pk && (rc = launch_pack(ctx, sdone, p->H, p->W, np, (const int8_t *)d_coef + (size_t)p0 * npix * 3, pd,
                                        (size_t)p0 * rows_per_p, pd.totals + 2 + 2 * c, d_rec == nullptr)))
                return rc;
        }
        CK(ctx, cudaEventRecord(ctx->chunk_events[2 * c + 1], sdone));
        if (trace) host_ms.push_back(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count());
        tmark(sd, c, 2);
        if (pk) {
            CK(ctx, cudaStreamWaitEvent(ctx->s_aux, ctx->chunk_events[2 * c + 1], 0));
            CK(ctx, cudaMemcpyAsync(&ctx->h_segend[2 * c], pd.totals + 2 + 2 * c, 16, cudaMemcpyDeviceToHost, ctx->s_aux));
            CK(ctx, cudaEventRecord(ctx->seg_events[c], ctx->s_aux));
        } else {
            CK(ctx, cudaStreamWaitEvent(ctx->s_d2h, ctx->chunk_events[2 * c + 1], 0));
            if (np > 0 && (rc = download(p0, np))) return rc;
            tmark(ctx->s_d2h, c, 3);
        }
        p0 += np;
    }
    if (!pk) return VCS_OK;
    unsigned long long have_n = 0, have_e = 0;      // bytes of the two streams already queued for download
    p0 = 0;
    for (int c = 0; c < nsegs; ++c) {
        const int np = sizes[c];
        CK(ctx, cudaEventSynchronize(ctx->seg_events[c]));       // segment c is complete and its stream lengths are here
        const unsigned long long end_n = ctx->h_segend[2 * c], end_e = ctx->h_segend[2 * c + 1];
        if (end_n > pk->cap_n || end_e > pk->cap_e)
            return fail(ctx, VCS_E_INVALID, "packed output buffers too small (nibbles %llu > %zu or escapes %llu > %zu bytes)",
                        end_n, pk->cap_n, end_e, pk->cap_e);
        if (np > 0) {
            if (end_n > have_n)
                CK(ctx, cudaMemcpyAsync(pk->nibbles + have_n, pd.nibbles + have_n, end_n - have_n, cudaMemcpyDeviceToHost, ctx->s_d2h));
            if (end_e > have_e)
                CK(ctx, cudaMemcpyAsync(pk->escapes + have_e, pd.escapes + have_e, end_e - have_e, cudaMemcpyDeviceToHost, ctx->s_d2h));
            CK(ctx, cudaMemcpyAsync(pk->bitmap + (size_t)p0 * rows_per_p * nbx8, pd.bitmap + (size_t)p0 * rows_per_p * nbx8,
                                    (size_t)np * rows_per_p * nbx8 * 8, cudaMemcpyDeviceToHost, ctx->s_d2h));
            CK(ctx, cudaMemcpyAsync(pk->row_count + (size_t)p0 * rows_per_p * 2, pd.row_count + (size_t)p0 * rows_per_p,
                                    (size_t)np * rows_per_p * 8, cudaMemcpyDeviceToHost, ctx->s_d2h));
            if ((rc = download(p0, np))) return rc;
        }
        tmark(ctx->s_d2h, c, 3);
        have_n = end_n; have_e = end_e;
        pk->lengths[0] = have_n; pk->lengths[1] = have_e;
        p0 += np;
    }
    return VCS_OK;
    };
    rc = pipeline();
    // success or not, nothing may still be reading or writing the caller's buffers when this returns
    cudaError_t e1 = cudaStreamSynchronize(ctx->s_h2d), e2 = cudaStreamSynchronize(sc), e3 = cudaStreamSynchronize(ctx->s_d2h),
                e4 = cudaStreamSynchronize(ctx->s_aux), e5 = cudaStreamSynchronize(sd);
    if (pack_aside) { const cudaError_t e6 = cudaStreamSynchronize(ctx->s_dct); if (e5 == cudaSuccess) e5 = e6; }
    if (trace && !tev.empty()) {
        auto ms = [&](size_t k) { float t = 0; cudaEventElapsedTime(&t, tev[0], tev[k]); return t; };
        fprintf(stderr, "[vcs trace] %d segments, %s sink; per segment: P-frames | upload done, compute start, compute done, download done | host had queued the kernels (ms)\n",
                nsegs, pk ? "packed" : "dense");
        for (int c = 0; c < nsegs; ++c) {
            const size_t b = 1 + 4 * (size_t)c;
            if (b + 3 < tev.size() && tev[b] && tev[b + 1] && tev[b + 2] && tev[b + 3])
                fprintf(stderr, "[vcs trace] seg %2d np %2d | %7.3f %7.3f %7.3f %7.3f | %7.3f\n", c, sizes[c], ms(b), ms(b + 1), ms(b + 2), ms(b + 3),
                        c < (int)host_ms.size() ? host_ms[c] : -1.0);
        }
        for (auto e : tev) if (e) cudaEventDestroy(e);
    }
    if (rc) return rc;
    CK(ctx, e1); CK(ctx, e2); CK(ctx, e3); CK(ctx, e4); CK(ctx, e5);
    return pending_device_error(ctx);
}

int vcs_encode_clip_host(vcs_ctx *ctx, const vcs_me_params *p, const uint8_t *frames, int T, int gop_len,
                         int coef_mode, int16_t *mv, uint32_t *cost, uint8_t *flags, void *coef,
                         uint8_t *recon) {
    return encode_clip_host_impl(ctx, p, frames, T, gop_len, coef_mode, mv, cost, flags, coef, recon, nullptr);
}

int vcs_encode_clip_host_packed(vcs_ctx *ctx, const vcs_me_params *p, const uint8_t *frames, int T, int gop_len,
                                int16_t *mv, uint32_t *cost, uint8_t *flags, uint64_t *bitmap, uint32_t *row_count,
                                uint8_t *nibbles, size_t nibbles_capacity, int8_t *escapes, size_t escapes_capacity,
                                uint64_t *lengths, uint8_t *recon) {
    if (!ctx) return VCS_E_INVALID;
    if (!bitmap || !row_count || !nibbles || !escapes || !lengths) return fail(ctx, VCS_E_INVALID, "NULL packed output");
    PackedHost pk{(unsigned long long *)bitmap, row_count, nibbles, nibbles_capacity, escapes, escapes_capacity,
                  (unsigned long long *)lengths};
    return encode_clip_host_impl(ctx, p, frames, T, gop_len, VCS_COEF_I8_RINT, mv, cost, flags, nullptr, recon, &pk);
}

// dense int8 index planes [nP][3][H][W] (device) -> packed form (device): bitmap [nP][3][H/8][W/8], row_count
// [nP][3][H/8][2], nibbles (3*H*W*nP/2 + one byte per 2 blocks is always enough), escapes (3*H*W*nP is always enough);
// lengths_host[2] = bytes of the nibble stream, number of escapes (synchronises).
int vcs_pack_coef_dev(vcs_ctx *ctx, int H, int W, int nP, const int8_t *coef, uint64_t *bitmap, uint32_t *row_count,
                      uint8_t *nibbles, int8_t *escapes, uint64_t *lengths_host) {
    VCS_ENTER(ctx);
    if (!coef || !bitmap || !row_count || !nibbles || !escapes || !lengths_host || H <= 0 || W <= 0 || H % 8 || W % 8 || nP < 0)
        return fail(ctx, VCS_E_INVALID, "bad pack arguments");
    if (((uintptr_t)coef | (uintptr_t)bitmap | (uintptr_t)row_count) & 7)
        return fail(ctx, VCS_E_INVALID, "coef, bitmap and row_count must be 8-byte aligned");
    const size_t nrows = (size_t)nP * 3 * (H / 8), nbx = (size_t)W / 8;
    PackedDev d; int rc;
    memset(&d, 0, sizeof(d));
    d.bitmap = (unsigned long long *)bitmap; d.row_count = (uint2 *)row_count; d.nibbles = nibbles; d.escapes = escapes;
    unsigned long long *off;
    if ((rc = dev_buf(ctx, S_PK_BLKESC, nrows * nbx + 8, (void **)&d.blk_esc))) return rc;
    if ((rc = dev_buf(ctx, S_PK_ROWOFF, nrows * 16 + 16, (void **)&off))) return rc;
    d.nib_off = off; d.esc_off = off + nrows;
    if ((rc = dev_buf(ctx, S_PK_TOTAL, 16, (void **)&d.totals))) return rc;
    cudaStream_t st = ctx->stream;
    CK(ctx, cudaMemsetAsync(d.totals, 0, 16, st));
    if ((rc = launch_pack(ctx, st, H, W, nP, coef, d, 0, nullptr, false))) return rc;
    unsigned long long tot[2] = {0, 0};
    CK(ctx, cudaMemcpyAsync(tot, d.totals, 16, cudaMemcpyDeviceToHost, st));
    CK(ctx, cudaStreamSynchronize(st));
    lengths_host[0] = tot[0]; lengths_host[1] = tot[1];
    return VCS_OK;
}

// the exact inverse, all device pointers; enqueued on the context's stream
int vcs_unpack_coef_dev(vcs_ctx *ctx, int H, int W, int nP, const uint64_t *bitmap, const uint32_t *row_count,
                        const uint8_t *nibbles, uint64_t nnibble_bytes, const int8_t *escapes, uint64_t nescapes,
                        int8_t *coef) {
    VCS_ENTER(ctx);
    if (!coef || !bitmap || !row_count || (!nibbles && nnibble_bytes) || (!escapes && nescapes) || H <= 0 || W <= 0 ||
        H % 8 || W % 8 || nP < 0)
        return fail(ctx, VCS_E_INVALID, "bad unpack arguments");
    if (((uintptr_t)coef | (uintptr_t)bitmap | (uintptr_t)row_count) & 7)
        return fail(ctx, VCS_E_INVALID, "coef, bitmap and row_count must be 8-byte aligned");
    const int nrows = nP * 3 * (H / 8);
    if (nrows == 0) return VCS_OK;
    unsigned long long *off, *d_total; int rc;
    if ((rc = dev_buf(ctx, S_PK_ROWOFF, (size_t)nrows * 16 + 16, (void **)&off))) return rc;
    if ((rc = dev_buf(ctx, S_PK_TOTAL, 16, (void **)&d_total))) return rc;
    cudaStream_t st = ctx->stream;
    CK(ctx, cudaMemsetAsync(d_total, 0, 16, st));
    pack_scan_kernel<<<1, 1024, 0, st>>>((const uint2 *)row_count, nrows, off, off + nrows, d_total, nullptr);
    unpack_kernel<<<pack_grid(ctx, nrows), 32 * PACK_WARPS, 0, st>>>((const unsigned long long *)bitmap, off, off + nrows, nibbles,
                                                                    nnibble_bytes, escapes, nescapes, W, nrows, coef, ctx->h_errflag);
    CK(ctx, cudaGetLastError());
    ctx->launches += 2;
    return VCS_OK;
}

// Decoder side from the packed form, host buffers: upload, unpack, then vcs_decode_clip_dev's kernel.
int vcs_decode_clip_host_packed(vcs_ctx *ctx, int H, int W, int bs, const uint8_t *ref_frames, int T, int gop_len,
                                const int16_t *mv, const uint64_t *bitmap, const uint32_t *row_count,
                                const uint8_t *nibbles, uint64_t nnibble_bytes, const int8_t *escapes, uint64_t nescapes,
                                uint8_t *recon) {
    VCS_ENTER(ctx);
    if (!ref_frames || !mv || !bitmap || !row_count || !recon || gop_len < 2 || T < 1 || bs <= 0 || H < bs || W < bs)
        return fail(ctx, VCS_E_INVALID, "bad decode arguments");
    if (H % 8 || W % 8) return fail(ctx, VCS_E_INVALID, "H=%d W=%d must be multiples of 8", H, W);
    const size_t fs = (size_t)H * W * 3, npix = (size_t)H * W;
    const int N = vcs_num_blocks(H, W, bs), nP = vcs_num_p_frames(T, gop_len), nG = (T + gop_len - 1) / gop_len;
    for (long long k = 0; k < (long long)nP * N; ++k) {
        const int b = (int)(k % N);
        const int x = (b % (W / bs)) * bs + mv[2 * k], y = (b / (W / bs)) * bs + mv[2 * k + 1];
        if (x < 0 || y < 0 || x + bs > W || y + bs > H)
            return fail(ctx, VCS_E_INVALID, "motion vector %lld points outside the frame", k);
    }
    const size_t nrows = (size_t)nP * 3 * (H / 8), nbx8 = W / 8;
    unsigned long long tot_n = 0, tot_e = 0;
    for (size_t k = 0; k < nrows; ++k) { tot_n += row_count[2 * k]; tot_e += row_count[2 * k + 1]; }
    if (tot_n != nnibble_bytes || tot_e != nescapes)
        return fail(ctx, VCS_E_INVALID, "row counts (%llu, %llu) do not add up to the stream lengths (%llu, %llu)", tot_n, tot_e,
                    (unsigned long long)nnibble_bytes, (unsigned long long)nescapes);
    uint8_t *d_ref, *d_rec, *d_nib; int16_t *d_mv; int8_t *d_coef, *d_esc; unsigned long long *d_bitmap; uint32_t *d_rowcnt; int rc;
    if ((rc = dev_buf(ctx, S_FRAMES, fs * nG, (void **)&d_ref))) return rc;
    if ((rc = dev_buf(ctx, S_MV, (size_t)nP * N * 4 + 4, (void **)&d_mv))) return rc;
    if ((rc = dev_buf(ctx, S_COEF, (size_t)nP * npix * 3 + 8, (void **)&d_coef))) return rc;
    if ((rc = dev_buf(ctx, S_RECON, (size_t)nP * fs + 4, (void **)&d_rec))) return rc;
    if ((rc = dev_buf(ctx, S_PK_BITMAP, nrows * nbx8 * 8 + 8, (void **)&d_bitmap))) return rc;
    if ((rc = dev_buf(ctx, S_PK_ROWCNT, nrows * 8 + 8, (void **)&d_rowcnt))) return rc;
    if ((rc = dev_buf(ctx, S_PK_VALUES, (size_t)nnibble_bytes + 64, (void **)&d_nib))) return rc;
    if ((rc = dev_buf(ctx, S_PK_ESC, (size_t)nescapes + 64, (void **)&d_esc))) return rc;
    cudaStream_t st = ctx->stream;
    CK(ctx, cudaMemcpyAsync(d_ref, ref_frames, fs * nG, cudaMemcpyHostToDevice, st));
    CK(ctx, cudaMemcpyAsync(d_mv, mv, (size_t)nP * N * 4, cudaMemcpyHostToDevice, st));
    CK(ctx, cudaMemcpyAsync(d_bitmap, bitmap, nrows * nbx8 * 8, cudaMemcpyHostToDevice, st));
    CK(ctx, cudaMemcpyAsync(d_rowcnt, row_count, nrows * 8, cudaMemcpyHostToDevice, st));
    if (nnibble_bytes) CK(ctx, cudaMemcpyAsync(d_nib, nibbles, (size_t)nnibble_bytes, cudaMemcpyHostToDevice, st));
    if (nescapes) CK(ctx, cudaMemcpyAsync(d_esc, escapes, (size_t)nescapes, cudaMemcpyHostToDevice, st));
    if ((rc = vcs_unpack_coef_dev(ctx, H, W, nP, (const uint64_t *)d_bitmap, d_rowcnt, d_nib, nnibble_bytes, d_esc, nescapes, d_coef)))
        return rc;
    if ((rc = vcs_decode_clip_dev(ctx, H, W, bs, d_ref, T, gop_len, d_mv, VCS_COEF_I8_RINT, d_coef, d_rec))) return rc;
    CK(ctx, cudaMemcpyAsync(recon, d_rec, (size_t)nP * fs, cudaMemcpyDeviceToHost, st));
    CK(ctx, cudaStreamSynchronize(st));
    return pending_device_error(ctx);
}

// ---- intra mode decision (IntraframeCompression/intraframe.py:24-317) ---------------------------
static int intra_check(vcs_ctx *ctx, int H, int W, int m, const void *a, const void *b) {
    if (!a || !b) return fail(ctx, VCS_E_INVALID, "NULL argument");
    if (H <= 0 || W <= 0 || H % m || W % m)
        return fail(ctx, VCS_E_INVALID, "plane sides must be positive multiples of %d", m);
    return VCS_OK;
}

int vcs_intra_luma4x4_dev(vcs_ctx *ctx, int H, int W, const uint8_t *Y, int32_t *res, int32_t *pred, uint8_t *modes) {
    VCS_ENTER(ctx);
    int rc = intra_check(ctx, H, W, 4, Y, modes);
    if (rc) return rc;
    if (!res || !pred || (W / 4 < 2 && H / 4 > 1))
        return fail(ctx, VCS_E_INVALID, "luma4x4 needs at least two block columns (the reference indexes column j+1)");
    if (((uintptr_t)Y & 3) || (((uintptr_t)res | (uintptr_t)pred) & 15))
        return fail(ctx, VCS_E_INVALID, "luma4x4: Y must be 4-byte aligned, res/pred 16-byte aligned");
    const int nb = (H / 4) * (W / 4);
    intra_luma4x4_kernel<<<(nb + 127) / 128, 128, 0, ctx->stream>>>(Y, H, W, res, pred, modes);
    CK(ctx, cudaGetLastError());
    ctx->launches += 1;
    return VCS_OK;
}

int vcs_intra_luma16x16_dev(vcs_ctx *ctx, int H, int W, const uint8_t *Y, int32_t *res, int32_t *pred, uint8_t *modes) {
    VCS_ENTER(ctx);
    int rc = intra_check(ctx, H, W, 16, Y, modes);
    if (rc) return rc;
    if (!res || !pred) return fail(ctx, VCS_E_INVALID, "NULL argument");
    if (((uintptr_t)Y & 7) || (((uintptr_t)res | (uintptr_t)pred) & 15))
        return fail(ctx, VCS_E_INVALID, "luma16x16: Y must be 8-byte aligned, res/pred 16-byte aligned");
    const int nb = (H / 16) * (W / 16);
    intra_luma16x16_kernel<<<(nb * 32 + 255) / 256, 256, 0, ctx->stream>>>(Y, H, W, res, pred, modes);
    CK(ctx, cudaGetLastError());
    ctx->launches += 1;
    return VCS_OK;
}

int vcs_intra_chroma8x8_dev(vcs_ctx *ctx, int H, int W, const uint8_t *Cr, const uint8_t *Cb, int32_t *crres,
                            int32_t *crpred, int32_t *cbres, int32_t *cbpred, uint8_t *modes) {
    VCS_ENTER(ctx);
    int rc = intra_check(ctx, H, W, 8, Cr, Cb);
    if (rc) return rc;
    if (!crres || !crpred || !cbres || !cbpred || !modes) return fail(ctx, VCS_E_INVALID, "NULL argument");
    const int ncol = W / 8;
    intra_chroma8x8_kernel<<<(ncol * 32 + 127) / 128, 128, 0, ctx->stream>>>(Cr, Cb, H, W, crres, crpred, cbres,
                                                                             cbpred, modes);
    CK(ctx, cudaGetLastError());
    ctx->launches += 1;
    return VCS_OK;
}

// which: 0 luma4x4, 1 luma16x16 (planes: Y), 2 chroma8x8 (planes: Cr then Cb).  Host buffers.
int vcs_intra_host(vcs_ctx *ctx, int which, int H, int W, const uint8_t *p0, const uint8_t *p1, int32_t *res0,
                   int32_t *pred0, int32_t *res1, int32_t *pred1, uint8_t *modes) {
    VCS_ENTER(ctx);
    if (which < 0 || which > 2) return fail(ctx, VCS_E_INVALID, "which must be 0, 1 or 2");
    const int m = which == 0 ? 4 : (which == 1 ? 16 : 8);
    int rc = intra_check(ctx, H, W, m, p0, modes);
    if (rc) return rc;
    if (!res0 || !pred0 || (which == 2 && (!p1 || !res1 || !pred1))) return fail(ctx, VCS_E_INVALID, "NULL argument");
    const size_t n = (size_t)H * W, nm = (size_t)(H / m) * (W / m);
    uint8_t *d_in, *d_modes; int32_t *d_out;
    if ((rc = dev_buf(ctx, S_FRAMES, 2 * n, (void **)&d_in))) return rc;
    if ((rc = dev_buf(ctx, S_COEF, 4 * n * 4, (void **)&d_out))) return rc;
    if ((rc = dev_buf(ctx, S_FLAGS, nm, (void **)&d_modes))) return rc;
    cudaStream_t st = ctx->stream;
    CK(ctx, cudaMemcpyAsync(d_in, p0, n, cudaMemcpyHostToDevice, st));
    if (which == 2) CK(ctx, cudaMemcpyAsync(d_in + n, p1, n, cudaMemcpyHostToDevice, st));
    if (which == 0) rc = vcs_intra_luma4x4_dev(ctx, H, W, d_in, d_out, d_out + n, d_modes);
    else if (which == 1) rc = vcs_intra_luma16x16_dev(ctx, H, W, d_in, d_out, d_out + n, d_modes);
    else rc = vcs_intra_chroma8x8_dev(ctx, H, W, d_in, d_in + n, d_out, d_out + n, d_out + 2 * n, d_out + 3 * n, d_modes);
    if (rc) return rc;
    CK(ctx, cudaMemcpyAsync(res0, d_out, n * 4, cudaMemcpyDeviceToHost, st));
    CK(ctx, cudaMemcpyAsync(pred0, d_out + n, n * 4, cudaMemcpyDeviceToHost, st));
    if (which == 2) {
        CK(ctx, cudaMemcpyAsync(res1, d_out + 2 * n, n * 4, cudaMemcpyDeviceToHost, st));
        CK(ctx, cudaMemcpyAsync(pred1, d_out + 3 * n, n * 4, cudaMemcpyDeviceToHost, st));
    }
    CK(ctx, cudaMemcpyAsync(modes, d_modes, nm, cudaMemcpyDeviceToHost, st));
    CK(ctx, cudaStreamSynchronize(st));
    return VCS_OK;
}

// ---- measurement support ---------------------------------------------------------------------
int vcs_enable_kernel_timing(vcs_ctx *ctx, int on) {
    VCS_ENTER(ctx);
    ctx->timing = on != 0;
    ctx->ev_used = 0;
    return VCS_OK;
}

int vcs_kernel_times(vcs_ctx *ctx, double *me_ms_total, double *dct_ms_total, int *ncalls) {
    VCS_ENTER(ctx);
    CK(ctx, cudaDeviceSynchronize());
    double me = 0, dct = 0;
    for (size_t k = 0; k < ctx->ev_used; ++k) {
        EvTriple &t = ctx->ev_pool[k];
        float ms;
        if (t.has_me) { CK(ctx, cudaEventElapsedTime(&ms, t.e0, t.e1)); me += ms; }
        if (t.has_dct) { CK(ctx, cudaEventElapsedTime(&ms, t.e1, t.e2)); dct += ms; }
    }
    if (me_ms_total) *me_ms_total = me;
    if (dct_ms_total) *dct_ms_total = dct;
    if (ncalls) *ncalls = (int)ctx->ev_used;
    ctx->ev_used = 0;
    return VCS_OK;
}

// ---- 4:2:0 chroma subsampling demo (ChromaSubsampling/chroma.py) ------------------------------------
int vcs_chroma420_dev(vcs_ctx *ctx, int H, int W, const uint8_t *bgr, uint8_t *Y, uint8_t *cr, uint8_t *cb) {
    VCS_ENTER(ctx);
    if (H <= 0 || W <= 0 || !bgr || !Y || !cr || !cb) return fail(ctx, VCS_E_INVALID, "bad chroma420 arguments");
    const dim3 block(32, 8), grid((W / 2 + 1 + 31) / 32, (H / 2 + 1 + 7) / 8);
    chroma420_kernel<<<grid, block, 0, ctx->stream>>>(bgr, H, W, Y, cr, cb);
    CK(ctx, cudaGetLastError());
    ctx->launches += 1;
    return VCS_OK;
}

int vcs_chroma420_to_bgr_dev(vcs_ctx *ctx, int H, int W, const uint8_t *Y, const uint8_t *cr, const uint8_t *cb,
                             uint8_t *bgr) {
    VCS_ENTER(ctx);
    if (H <= 0 || W <= 0 || !bgr || !Y || !cr || !cb) return fail(ctx, VCS_E_INVALID, "bad chroma420 arguments");
    const size_t npix = (size_t)H * W;
    size_t blocks = (npix + 255) / 256;
    const size_t cap = (size_t)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    chroma420_to_bgr_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(Y, cr, cb, H, W, bgr);
    CK(ctx, cudaGetLastError());
    ctx->launches += 1;
    return VCS_OK;
}

int vcs_chroma420_host(vcs_ctx *ctx, int H, int W, const uint8_t *bgr, uint8_t *Y, uint8_t *cr, uint8_t *cb,
                       uint8_t *bgr_out) {
    VCS_ENTER(ctx);
    if (H <= 0 || W <= 0 || !bgr || !Y || !cr || !cb) return fail(ctx, VCS_E_INVALID, "bad chroma420 arguments");
    const size_t npix = (size_t)H * W, ns = (size_t)((H + 1) / 2) * ((W + 1) / 2);
    uint8_t *d_img, *d_pl, *d_out = nullptr;
    int rc;
    if ((rc = dev_buf(ctx, S_FRAMES, npix * 3, (void **)&d_img))) return rc;
    if ((rc = dev_buf(ctx, S_AUX0, npix + 2 * ns, (void **)&d_pl))) return rc;
    if (bgr_out && (rc = dev_buf(ctx, S_RECON, npix * 3, (void **)&d_out))) return rc;
    cudaStream_t st = ctx->stream;
    CK(ctx, cudaMemcpyAsync(d_img, bgr, npix * 3, cudaMemcpyHostToDevice, st));
    if ((rc = vcs_chroma420_dev(ctx, H, W, d_img, d_pl, d_pl + npix, d_pl + npix + ns))) return rc;
    if (bgr_out && (rc = vcs_chroma420_to_bgr_dev(ctx, H, W, d_pl, d_pl + npix, d_pl + npix + ns, d_out))) return rc;
    CK(ctx, cudaMemcpyAsync(Y, d_pl, npix, cudaMemcpyDeviceToHost, st));
    CK(ctx, cudaMemcpyAsync(cr, d_pl + npix, ns, cudaMemcpyDeviceToHost, st));
    CK(ctx, cudaMemcpyAsync(cb, d_pl + npix + ns, ns, cudaMemcpyDeviceToHost, st));
    if (bgr_out) CK(ctx, cudaMemcpyAsync(bgr_out, d_out, npix * 3, cudaMemcpyDeviceToHost, st));
    CK(ctx, cudaStreamSynchronize(st));
    return VCS_OK;
}

int vcs_chroma420_to_bgr_host(vcs_ctx *ctx, int H, int W, const uint8_t *Y, const uint8_t *cr, const uint8_t *cb,
                              uint8_t *bgr) {
    VCS_ENTER(ctx);
    if (H <= 0 || W <= 0 || !bgr || !Y || !cr || !cb) return fail(ctx, VCS_E_INVALID, "bad chroma420 arguments");
    const size_t npix = (size_t)H * W, ns = (size_t)((H + 1) / 2) * ((W + 1) / 2);
    uint8_t *d_pl, *d_out;
    int rc;
    if ((rc = dev_buf(ctx, S_AUX0, npix + 2 * ns, (void **)&d_pl))) return rc;
    if ((rc = dev_buf(ctx, S_RECON, npix * 3, (void **)&d_out))) return rc;
    cudaStream_t st = ctx->stream;
    CK(ctx, cudaMemcpyAsync(d_pl, Y, npix, cudaMemcpyHostToDevice, st));
    CK(ctx, cudaMemcpyAsync(d_pl + npix, cr, ns, cudaMemcpyHostToDevice, st));
    CK(ctx, cudaMemcpyAsync(d_pl + npix + ns, cb, ns, cudaMemcpyHostToDevice, st));
    if ((rc = vcs_chroma420_to_bgr_dev(ctx, H, W, d_pl, d_pl + npix, d_pl + npix + ns, d_out))) return rc;
    CK(ctx, cudaMemcpyAsync(bgr, d_out, npix * 3, cudaMemcpyDeviceToHost, st));
    CK(ctx, cudaStreamSynchronize(st));
    return VCS_OK;
}

int vcs_microbench(vcs_ctx *ctx, int which, int iters, double *warp_instr_per_s, double *sm_mhz) {
    VCS_ENTER(ctx);
    if (iters <= 0) iters = 2000;
    switch (which) {
        case 0: return run_mb<0>(ctx, iters, warp_instr_per_s, sm_mhz);
        case 1: return run_mb<1>(ctx, iters, warp_instr_per_s, sm_mhz);
        case 2: return run_mb<2>(ctx, iters, warp_instr_per_s, sm_mhz);
        case 3: return run_mb<3>(ctx, iters, warp_instr_per_s, sm_mhz);
        case 4: return run_mb<4>(ctx, iters, warp_instr_per_s, sm_mhz);
        case 5: return run_mb<5>(ctx, iters, warp_instr_per_s, sm_mhz);
        case 6: return run_mb<6>(ctx, iters, warp_instr_per_s, sm_mhz);
        default: return fail(ctx, VCS_E_INVALID, "unknown microbenchmark %d", which);
    }
}

}  // extern "C"
