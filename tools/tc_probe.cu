// tc_probe.cu -- can the 5th-gen tensor core be the BYTE-SUM engine of the wrapped-cost search? (sm_100a)
//
// The wrapped cost of the reference (motion.py:146) needs, per 4 bytes, a borrow-isolated subtract, a 3-input
// xor and a sum of the 4 result bytes.  The byte sum is IDP.4A today (1 of 3 issue slots).  This probe moves
// it to tcgen05.mma.kind::i8: every thread writes its result words z to TMEM (tcgen05.st.32x32b, A operand,
// row = thread, 4 K-bytes per 32-bit column), B is a 0/1 selector matrix in shared memory that routes word j
// of a row to accumulator column j, D accumulates in TMEM.
//
//   part 1 (mode 0..7)  : correctness of one MMA against a host byte sum, for the B layouts / D column
//                         offsets the kernel would need (which LBO/SBO reading is right, is an unaligned
//                         D base legal, N = 8).
//   part 2 (mode 10..15): throughput of the producer/consumer pipeline (16 worker warps compute z, one
//                         thread issues MMAs) against the IDP.4A loop it would replace.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/tc_probe tools/tc_probe.cu
//   for m in 0 1 2 3 4 5 6 7 10 11 12 13 14 15; do timeout 60 tools/tc_probe $m; done
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); return 2; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bounded wait: returns false on timeout (the probe must never hang the box)
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity, int max_tries = 1 << 20) {
    for (int t = 0; t < max_tries; ++t) {
        uint32_t ok;
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}
// non-suspending poll (mbarrier.test_wait): separates the barrier's own latency from try_wait's sleep/wake-up
__device__ __forceinline__ bool mbar_poll(uint64_t *bar, uint32_t parity, int max_tries = 1 << 22) {
    for (int t = 0; t < max_tries; ++t) {
        uint32_t ok;
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t *v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t *v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                   "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
                 : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t *v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]; int8 kinds; one thread issues
__device__ __forceinline__ void tc_mma_i8_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_alloc(uint32_t *slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// shared-memory matrix descriptor, no swizzle (cute::UMMA::SmemDescriptor): address, LBO, SBO in 16-byte units
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;   // version = 1 (Blackwell)
    return d;                 // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}
// instruction descriptor (cute::UMMA::InstrDescriptor): S32 accumulate, unsigned 8-bit A and B, A K-major
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int b_mn_major) {
    return (2u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__host__ __device__ inline uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// ------------------------------------------------------------------------------------------------------------
// part 1: one CTA of 128 threads, two MMAs (K = 32 each) into one N-wide accumulator block
//   mode 0: B N-major (byte k*16+n), N=16        mode 1: B K-major, LBO=128 (k half), SBO=256 (n group)
//   mode 2: B K-major with LBO/SBO swapped       mode 3/4/5: mode 0 with the D base at column +8 / +4 / +1
//   mode 6: N=8 (K-major, one n group)           mode 7: mode 0, N=32 (selector into columns 16..31)
__global__ void __launch_bounds__(128, 1) k_check(int mode, uint32_t *out, uint32_t *status) {
    __shared__ __align__(128) uint8_t sB[2][1024];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int N = mode == 6 ? 8 : (mode == 7 ? 32 : 16);
    const int dofs = mode == 3 ? 8 : (mode == 4 ? 4 : (mode == 5 ? 1 : 0));
    // selector: MMA q routes word j (K bytes 4j..4j+3) to column 8q + j (+16 in mode 7)
    for (int i = tid; i < 2 * 1024; i += 128) (&sB[0][0])[i] = 0;
    __syncthreads();
    for (int i = tid; i < 2 * 32; i += 128) {
        const int q = i / 32, k = i % 32;
        int n = (mode == 6 ? 0 : 8 * q) + k / 4 + (mode == 7 ? 16 : 0);
        if (mode == 6 && q == 1) n = 7 - k / 4;        // second MMA: reversed routing
        uint32_t off;
        if (mode == 1 || mode == 6) off = (n % 8) * 16 + (n / 8) * 256 + (k % 16) + (k / 16) * 128;
        else if (mode == 2) off = (n % 8) * 16 + (n / 8) * 128 + (k % 16) + (k / 16) * 256;
        else off = k * N + n;                          // N-major: N contiguous bytes per k (N=16: 16 B rows; N=32: see below)
        if (mode == 7) off = (n / 16) * 512 + k * 16 + (n % 16);   // two 16-wide n groups, SBO = 512
        sB[q][off] = 1;
    }
    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tc_alloc(&tmem_slot, 64);
    // make the generic-proxy writes of B visible to the async proxy (the MMA reads smem through it)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_slot;
    const uint32_t lane_base = tbase + ((uint32_t)(32 * warp) << 16);
    // columns: D at [0, 40) (N <= 32 plus offset), A at [48, 64)
    uint32_t zeros[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) zeros[j] = 0;
    tc_st16(lane_base + 0, zeros);
    tc_st16(lane_base + 16, zeros);
    tc_st8(lane_base + 32, zeros);
    uint32_t a[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = hash32(tid * 16 + j + 1);
    tc_st16(lane_base + 48, a);
    tc_wait_st();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
        const uint32_t idesc = make_idesc(128, N, (mode == 1 || mode == 2 || mode == 6) ? 0 : 1);
        for (int q = 0; q < 2; ++q) {
            uint64_t desc;
            const uint32_t sa = smem_u32(&sB[q][0]);
            if (mode == 1 || mode == 6) desc = make_desc(sa, 128, 256);
            else if (mode == 2) desc = make_desc(sa, 256, 128);
            else if (mode == 7) desc = make_desc(sa, 128, 512);
            else desc = make_desc(sa, 128, 128);
            tc_mma_i8_ts(tbase + dofs, tbase + 48 + 8 * q, desc, idesc, 1);
        }
        tc_commit(&bar);
    }
    const bool ok = mbar_wait(&bar, 0);
    tc_fence_after();
    uint32_t d[16], d2[16], d3[16];
    tc_ld16(lane_base + 0, d);
    tc_ld16(lane_base + 16, d2);
    tc_ld16(lane_base + 24, d3);
    tc_wait_ld();
#pragma unroll
    for (int j = 0; j < 16; ++j) { out[tid * 40 + j] = d[j]; out[tid * 40 + 16 + j] = d2[j]; }
#pragma unroll
    for (int j = 8; j < 16; ++j) out[tid * 40 + 24 + j] = d3[j];
    if (!ok) atomicAdd(status, 1);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tc_dealloc(tbase, 64);
}

// ------------------------------------------------------------------------------------------------------------
// part 2: throughput.  16 worker warps, each iteration ("row step") = one window-row word used against the 16
// macroblock rows: 16 x (subtract, 3-input xor) [+ 16 x IDP.4A in the baseline], like the search kernel's loop.
//   mode 10: baseline, IDP.4A byte sums in registers (today's loop)
//   mode 11: subtract + xor only (no byte sum at all): the ALU floor of the tensor variant
//   mode 12: + tcgen05.st.x16 of the 16 results + wait::st every step (no MMA, no barriers)
//   mode 13: full pipeline: st -> (4 warps of a lane-quarter group arrive) -> MMA thread issues 2 MMAs -> commit
//            frees the buffer; NBUF A buffers per group
//   mode 14: as 13 but wait::st / arrive deferred by one step (the st of step i overlaps the ALU work of i+1)
//   mode 15: as 14 with steps of 2 rows (32 results, 2 st.x16, 4 MMAs per arrive)
constexpr int WORKERS = 16;
struct Pipe {
    uint64_t full[4][4], empty[4][4];
};

template <int MODE>
__global__ void __launch_bounds__(32 * (WORKERS + 1), 1) k_pipe(int iters, uint32_t seed, uint32_t *out, long long *cyc, uint32_t *status) {
    __shared__ __align__(128) uint8_t sB[2][512];
    __shared__ Pipe pipe;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int ROWS = MODE == 15 ? 2 : 1;            // window rows per step
    constexpr int NBUF = ROWS == 2 ? 3 : 4;             // A buffers per group: 16 + 16*ROWS*NBUF <= 128 columns
    for (int i = tid; i < 2 * 512; i += blockDim.x) (&sB[0][0])[i] = 0;
    __syncthreads();
    for (int i = tid; i < 2 * 32; i += blockDim.x) { const int q = i / 32, k = i % 32; sB[q][k * 16 + 8 * q + k / 4] = 1; }
    if (tid == 0) {
        for (int g = 0; g < 4; ++g)
            for (int b = 0; b < NBUF; ++b) { mbar_init(&pipe.full[g][b], 4); mbar_init(&pipe.empty[g][b], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tc_alloc(&tmem_slot, 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_slot;
    // group g = warp / 4 owns columns [128 g, 128 g + 128): D at +0 (16 columns), A buffers at +16 + 16*ROWS*b
    const int g = (warp / 4) & 3, quarter = warp & 3;
    const uint32_t gcol = 128u * g;
    const uint32_t lane_base = tbase + ((uint32_t)(32 * quarter) << 16) + gcol;
    long long t0 = 0;
    bool ok = true;
    if (warp < WORKERS) {
        uint32_t c[16], ch[16], acc[16], z[16 * ROWS];
#pragma unroll
        for (int j = 0; j < 16; ++j) { c[j] = hash32(seed + tid * 16 + j) & 0x7f7f7f7fu; ch[j] = hash32(seed * 3 + tid * 16 + j) & 0x80808080u; acc[j] = 0; }
        uint32_t r = hash32(seed + tid);
        {
            uint32_t zeros[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) zeros[j] = 0;
            tc_st16(lane_base, zeros);
            tc_wait_st();
        }
        __syncwarp();
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const int b = it % NBUF;
            const uint32_t par = (uint32_t)(it / NBUF) & 1;
#pragma unroll
            for (int rr = 0; rr < ROWS; ++rr) {
                r = r * 1664525u + 1013904223u;                      // the next window word (stands in for the LDS)
                uint32_t r1, r2;
                asm volatile("lop3.b32 %0, %1, %2, %2, 0xfc;" : "=r"(r1) : "r"(r), "r"(0x80808080u));
                r2 = r1 - r;                                         // ~r & H on the FMA pipe
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    uint32_t zz;
                    asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(zz) : "r"(r1 - c[j]), "r"(r2), "r"(ch[j]));
                    if (MODE == 10) acc[j] = __dp4a(zz, 0x01010101u, acc[j]);
                    else if (MODE == 11) ch[j] = zz;                 // the result feeds the next step's xor: 2 instructions per word, all live
                    else z[16 * rr + j] = zz;
                }
            }
            if (MODE == 12) {
#pragma unroll
                for (int rr = 0; rr < ROWS; ++rr) tc_st16(lane_base + 16 + 16 * rr, z + 16 * rr);
                tc_wait_st();
            }
            if (MODE >= 13) {
                if (MODE >= 14 && it > 0) {                          // finish the previous step's stores
                    tc_wait_st();
                    tc_fence_before();
                    if (lane == 0) mbar_arrive(&pipe.full[g][(it - 1) % NBUF]);
                }
                if (it >= NBUF) ok = ok && mbar_wait(&pipe.empty[g][b], par ^ 1);
                tc_fence_after();
#pragma unroll
                for (int rr = 0; rr < ROWS; ++rr) tc_st16(lane_base + 16 + 16 * ROWS * b + 16 * rr, z + 16 * rr);
                if (MODE == 13) {
                    tc_wait_st();
                    tc_fence_before();
                    if (lane == 0) mbar_arrive(&pipe.full[g][b]);
                }
            }
        }
        if (MODE >= 14) {
            tc_wait_st();
            tc_fence_before();
            if (lane == 0) mbar_arrive(&pipe.full[g][(iters - 1) % NBUF]);
        }
        uint32_t s = 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) s += acc[j] + (MODE == 11 ? ch[j] : 0u);
        out[blockIdx.x * blockDim.x + tid] = s + r;
    } else if (MODE >= 13) {
        // the MMA warp: one thread walks (step, group) in order
        if (lane == 0) {
            const uint32_t idesc = make_idesc(128, 16, 1);
            const uint64_t d0 = make_desc(smem_u32(&sB[0][0]), 128, 128), d1 = make_desc(smem_u32(&sB[1][0]), 128, 128);
            for (int it = 0; it < iters && ok; ++it) {
                const int b = it % NBUF;
                const uint32_t par = (uint32_t)(it / NBUF) & 1;
                for (int gg = 0; gg < 4; ++gg) {
                    ok = ok && mbar_wait(&pipe.full[gg][b], par);
                    tc_fence_after();
                    const uint32_t col = tbase + 128u * gg;
#pragma unroll
                    for (int rr = 0; rr < ROWS; ++rr) {
                        tc_mma_i8_ts(col, col + 16 + 16 * ROWS * b + 16 * rr, d0, idesc, 1);
                        tc_mma_i8_ts(col, col + 16 + 16 * ROWS * b + 16 * rr + 8, d1, idesc, 1);
                    }
                    tc_commit(&pipe.empty[gg][b]);
                }
            }
        }
        __syncwarp();
    }
    const long long t1 = clock64();
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
    if (!ok) atomicAdd(status, 1);
    // drain: every MMA must have completed before TMEM goes away
    tc_fence_before();
    __syncthreads();
    if (MODE >= 13 && warp == WORKERS && lane == 0) {
        // last commits: wait until the final buffer of every group is free again
        for (int gg = 0; gg < 4; ++gg) {
            const int it = iters - 1;
            mbar_wait(&pipe.empty[gg][it % NBUF], (uint32_t)(it / NBUF) & 1);
        }
    }
    __syncthreads();
    tc_fence_after();
    if (MODE >= 13 && warp < WORKERS && (warp & 3) == quarter) {
        uint32_t d[16];
        tc_ld16(lane_base, d);
        tc_wait_ld();
        uint32_t s = 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) s += d[j];
        out[blockIdx.x * blockDim.x + tid] += s;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tc_dealloc(tbase, 512);
}


// ------------------------------------------------------------------------------------------------------------
// part 3 (mode 20..): raw issue rate of small tcgen05.mma.kind::i8 -- NI issuing warps (one elected lane each, own
// 128-column group: D at +0, A at +64), each issues `count` MMAs back to back, then one commit and a wait.
// Answers: is a 128 x N x 32 MMA issue-bound (time independent of N) and do several issuers overlap?
template <int N, int NI>
__global__ void __launch_bounds__(128, 1) k_issue(int count, uint32_t *out, long long *cyc, uint32_t *status) {
    __shared__ __align__(128) uint8_t sB[32 * 64];
    __shared__ __align__(8) uint64_t bar[4];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 32 * 64; i += 128) sB[i] = 0;
    __syncthreads();
    if (tid < 32) {                                // word j (K bytes 4j..4j+3) -> column j
        if (N >= 16) sB[tid * 16 + tid / 4] = 1;   // N-major, first 16-wide n group
        else sB[(tid / 4) * 16 + (tid % 16) + (tid / 16) * 128] = 1;   // K-major, one 8-row n group
    }
    if (tid == 0) {
        for (int g = 0; g < 4; ++g) mbar_init(&bar[g], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tc_alloc(&tmem_slot, 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_slot;
    {   // every warp zeroes its lane quarter of all four groups' D and fills A
        uint32_t z[16], a[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) { z[j] = 0; a[j] = hash32(tid * 16 + j + 1); }
        for (int g = 0; g < 4; ++g) {
            const uint32_t lb = tbase + ((uint32_t)(32 * warp) << 16) + 128u * g;
            for (int c = 0; c < 64; c += 16) tc_st16(lb + c, z);
            tc_st16(lb + 64, a);
        }
        tc_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    long long t0 = clock64(), t1 = t0;
    bool ok = true;
    if (warp < NI && lane == 0) {
        const uint32_t idesc = make_idesc(128, N, N >= 16 ? 1 : 0);
        // N = 8 uses the K-major form (one n group): same bytes work because only n < 8 rows are read: (n%8)*16 + k%16 + (k/16)*128
        const uint64_t desc = N >= 16 ? make_desc(smem_u32(sB), 128, 512) : make_desc(smem_u32(sB), 128, 256);
        const uint32_t col = tbase + 128u * warp;
        for (int i = 0; i < count; ++i) tc_mma_i8_ts(col, col + 64 + 8 * (i & 1), desc, idesc, 1);
        tc_commit(&bar[warp]);
        t1 = clock64();          // issue done
        ok = mbar_wait(&bar[warp], 0);
    }
    const long long t2 = clock64();
    if (warp < NI && lane == 0) { cyc[2 * warp] = t1 - t0; cyc[2 * warp + 1] = t2 - t0; }
    if (!ok) atomicAdd(status, 1);
    __syncthreads();
    tc_fence_after();
    uint32_t d[16];
    tc_ld16(tbase + ((uint32_t)(32 * warp) << 16), d);
    tc_wait_ld();
    out[tid] = d[0] + d[7];
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tc_dealloc(tbase, 512);
}

template <int N, int NI>
static int run_issue(int count) {
    uint32_t *d_out, *d_status; long long *d_cyc;
    CK(cudaMalloc(&d_out, 128 * 4)); CK(cudaMalloc(&d_cyc, 64)); CK(cudaMalloc(&d_status, 4));
    CK(cudaMemset(d_status, 0, 4));
    k_issue<N, NI><<<1, 128>>>(count / 4 + 1, d_out, d_cyc, d_status);
    k_issue<N, NI><<<1, 128>>>(count, d_out, d_cyc, d_status);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    long long cyc[8]; uint32_t st, o[128];
    CK(cudaMemcpy(cyc, d_cyc, 64, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&st, d_status, 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(o, d_out, 512, cudaMemcpyDeviceToHost));
    const uint32_t a0 = hash32(1), a8 = hash32(9);   // row 0: word 0 of buffer 0 and of buffer 1 (A + 8 columns)
    auto bsum = [](uint32_t a) { return (a & 255) + ((a >> 8) & 255) + ((a >> 16) & 255) + (a >> 24); };
    const uint32_t want0 = (uint32_t)((count + 1) / 2) * bsum(a0) + (uint32_t)(count / 2) * bsum(a8);
    printf("issue N=%d issuers=%d count=%d: issue %.1f clk/MMA, complete %.1f clk/MMA per issuer (%.1f aggregate), timeouts %u, D[0][0] %s\n",
           N, NI, count, (double)cyc[0] / count, (double)cyc[1] / count, (double)cyc[1] / count / NI, st,
           o[0] - (uint32_t)((count + 1) / 2) * bsum(hash32(8)) - (uint32_t)(count / 2) * bsum(hash32(16)) == want0 ? "ok" : "MISMATCH");
    return 0;
}


// ------------------------------------------------------------------------------------------------------------
// part 4 (mode 16..19): the protocol the search kernel would use.  No dedicated MMA warp: the 4 warps of a lane-quarter
// group meet at a named barrier once their stores are visible, and ONE of them (rotating) issues the group's MMAs and
// the commit that frees the A buffer; the others only arrive and go on.  ROWS window rows per step, NBUF A buffers.
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

template <int ROWS, int NBUF, bool DEFER, bool POLL>
__global__ void __launch_bounds__(512, 1) k_pipe2(int iters, uint32_t seed, uint32_t *out, long long *cyc, uint32_t *status) {
    __shared__ __align__(128) uint8_t sB[2][512];
    __shared__ __align__(8) uint64_t empty[4][4];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 2 * 512; i += blockDim.x) (&sB[0][0])[i] = 0;
    __syncthreads();
    for (int i = tid; i < 2 * 32; i += blockDim.x) { const int q = i / 32, k = i % 32; sB[q][k * 16 + 8 * q + k / 4] = 1; }
    if (tid == 0) {
        for (int g = 0; g < 4; ++g)
            for (int b = 0; b < 4; ++b) mbar_init(&empty[g][b], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tc_alloc(&tmem_slot, 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_slot;
    const int g = warp >> 2, quarter = warp & 3;
    const uint32_t gcol = tbase + 128u * g;                    // D at +0 (16 columns), A buffers at +64 + 16*ROWS*b
    const uint32_t lane_base = gcol + ((uint32_t)(32 * quarter) << 16);
    const uint32_t idesc = make_idesc(128, 16, 1);
    const uint64_t d0 = make_desc(smem_u32(&sB[0][0]), 128, 128), d1 = make_desc(smem_u32(&sB[1][0]), 128, 128);
    bool ok = true;
    uint32_t c[16], ch[16], z[16 * ROWS];
#pragma unroll
    for (int j = 0; j < 16; ++j) { c[j] = hash32(seed + tid * 16 + j) & 0x7f7f7f7fu; ch[j] = hash32(seed * 3 + tid * 16 + j) & 0x80808080u; }
    uint32_t r = hash32(seed + tid);
    {
        uint32_t zeros[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) zeros[j] = 0;
        tc_st16(lane_base, zeros);
        tc_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    auto publish = [&](int it) {      // stores of step `it` are done: meet the group; the step's issuer feeds the tensor core
        const int b = it % NBUF;
        tc_wait_st();
        tc_fence_before();
        const int id = 1 + NBUF * g + b;
        if (quarter == (it & 3)) {
            bar_sync(id, 128);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t a = gcol + 64 + 16 * ROWS * b;
#pragma unroll
                for (int rr = 0; rr < ROWS; ++rr) {
                    tc_mma_i8_ts(gcol, a + 16 * rr, d0, idesc, 1);
                    tc_mma_i8_ts(gcol, a + 16 * rr + 8, d1, idesc, 1);
                }
                tc_commit(&empty[g][b]);
            }
            __syncwarp();
        } else {
            bar_arrive(id, 128);
        }
    };
    long long waited = 0, published = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const int b = it % NBUF;
#pragma unroll
        for (int rr = 0; rr < ROWS; ++rr) {
            r = r * 1664525u + 1013904223u;
            uint32_t r1, r2;
            asm volatile("lop3.b32 %0, %1, %2, %2, 0xfc;" : "=r"(r1) : "r"(r), "r"(0x80808080u));
            r2 = r1 - r;
#pragma unroll
            for (int j = 0; j < 16; ++j)
                asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(z[16 * rr + j]) : "r"(r1 - c[j]), "r"(r2), "r"(ch[j]));
        }
        if (DEFER && it > 0) publish(it - 1);
        if (it >= NBUF) {
            const long long w0 = clock64();
            ok = ok && (POLL ? mbar_poll(&empty[g][b], (uint32_t)(it / NBUF - 1) & 1) : mbar_wait(&empty[g][b], (uint32_t)(it / NBUF - 1) & 1));
            waited += clock64() - w0;
        }
        tc_fence_after();
#pragma unroll
        for (int rr = 0; rr < ROWS; ++rr) tc_st16(lane_base + 64 + 16 * ROWS * b + 16 * rr, z + 16 * rr);
        if (!DEFER) { const long long w0 = clock64(); publish(it); published += clock64() - w0; }
    }
    if (DEFER) publish(iters - 1);
    const long long t1 = clock64();
    if (tid == 0) { cyc[blockIdx.x] = t1 - t0; if (blockIdx.x == 0) { cyc[gridDim.x] = waited; cyc[gridDim.x + 1] = published; } }
    // drain: the last NBUF commits
    for (int k = 0; k < NBUF && k < iters; ++k) {
        const int it = iters - 1 - k;
        ok = ok && mbar_wait(&empty[g][it % NBUF], (uint32_t)(it / NBUF) & 1);
    }
    if (!ok) atomicAdd(status, 1);
    tc_fence_after();
    uint32_t d[16];
    tc_ld16(lane_base, d);
    tc_wait_ld();
    uint32_t sum = r;
#pragma unroll
    for (int j = 0; j < 16; ++j) sum += d[j];
    out[blockIdx.x * blockDim.x + tid] = sum;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tc_dealloc(tbase, 512);
}

template <int ROWS, int NBUF, bool DEFER, bool POLL>
static int run_pipe2(int iters, int mode) {
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    uint32_t *d_out, *d_status; long long *d_cyc;
    CK(cudaMalloc(&d_out, (size_t)sms * 512 * 4)); CK(cudaMalloc(&d_cyc, (sms + 2) * 8)); CK(cudaMalloc(&d_status, 4));
    CK(cudaMemset(d_status, 0, 4));
    k_pipe2<ROWS, NBUF, DEFER, POLL><<<sms, 512>>>(iters / 8 + 1, 1234u, d_out, d_cyc, d_status);
    k_pipe2<ROWS, NBUF, DEFER, POLL><<<sms, 512>>>(iters, 1234u, d_out, d_cyc, d_status);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    long long cyc, extra[2]; uint32_t st;
    CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(extra, d_cyc + sms, 16, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&st, d_status, 4, cudaMemcpyDeviceToHost));
    printf("pipe2 mode %d (rows/step %d, buffers %d, deferred publish %d, poll %d): %lld clk for %d steps -> %.3f clk/word/SMSP; warp 0 per step: %.0f clk waiting for a free buffer, %.0f clk in publish; timeouts %u\n",
           mode, ROWS, NBUF, (int)DEFER, (int)POLL, cyc, iters, (double)cyc / (4.0 * iters * ROWS * 16), (double)extra[0] / iters, (double)extra[1] / iters, st);
    return 0;
}

static int run_check(int mode) {
    uint32_t *d_out, *d_status;
    CK(cudaMalloc(&d_out, 128 * 40 * 4));
    CK(cudaMalloc(&d_status, 4));
    CK(cudaMemset(d_out, 0xff, 128 * 40 * 4));
    CK(cudaMemset(d_status, 0, 4));
    k_check<<<1, 128>>>(mode, d_out, d_status);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    static uint32_t h[128 * 40];
    uint32_t st;
    CK(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&st, d_status, 4, cudaMemcpyDeviceToHost));
    const int dofs = mode == 3 ? 8 : (mode == 4 ? 4 : (mode == 5 ? 1 : 0));
    int bad = 0, shown = 0;
    for (int t = 0; t < 128; ++t) {
        uint32_t want[40] = {0};
        for (int j = 0; j < 16; ++j) {
            const uint32_t a = hash32(t * 16 + j + 1);
            const uint32_t bs = (a & 255) + ((a >> 8) & 255) + ((a >> 16) & 255) + (a >> 24);
            int col = j;
            if (mode == 6) col = j < 8 ? j : 7 - (j - 8);
            if (mode == 7) col = 16 + j;
            want[dofs + col] += bs;
        }
        for (int j = 0; j < 40; ++j)
            if (h[t * 40 + j] != want[j]) {
                ++bad;
                if (shown < 6) { printf("  row %d col %d: got %u want %u\n", t, j, h[t * 40 + j], want[j]); ++shown; }
            }
    }
    printf("check mode %d: %s (%d mismatching cells, barrier timeouts %u)\n", mode, bad == 0 && st == 0 ? "PASS" : "FAIL", bad, st);
    return bad != 0;
}

template <int MODE>
static int run_pipe(int iters) {
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int threads = 32 * (WORKERS + 1);
    uint32_t *d_out, *d_status; long long *d_cyc;
    CK(cudaMalloc(&d_out, (size_t)sms * threads * 4));
    CK(cudaMalloc(&d_cyc, sms * 8));
    CK(cudaMalloc(&d_status, 4));
    CK(cudaMemset(d_status, 0, 4));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_pipe<MODE><<<sms, threads>>>(iters / 8 + 1, 1234u, d_out, d_cyc, d_status);
    CK(cudaEventRecord(e0));
    k_pipe<MODE><<<sms, threads>>>(iters, 1234u, d_out, d_cyc, d_status);
    CK(cudaEventRecord(e1));
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    long long cyc; uint32_t st;
    CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&st, d_status, 4, cudaMemcpyDeviceToHost));
    const int rows = MODE == 15 ? 2 : 1;
    const double words = (double)iters * rows * 16;           // per thread
    // per SMSP: 4 worker warps; clocks per word per warp-slot = cyc / (4 warps * words)
    printf("pipe mode %d: %.3f ms, %lld clk for %d steps -> %.2f clk per row step per warp-quad slot, %.3f clk/word/SMSP, timeouts %u\n",
           MODE, ms, cyc, iters, (double)cyc / (iters * rows) / 4.0, (double)cyc / (4.0 * words), st);
    return 0;
}

int main(int argc, char **argv) {
    const int mode = argc > 1 ? atoi(argv[1]) : 0;
    const int iters = argc > 2 ? atoi(argv[2]) : 20000;
    if (mode < 10) return run_check(mode);
    if (mode >= 20 && mode < 30) {
        const int count = argc > 2 ? atoi(argv[2]) : 4096;
        switch (mode) {
            case 20: return run_issue<16, 1>(count);
            case 21: return run_issue<16, 2>(count);
            case 22: return run_issue<16, 4>(count);
            case 23: return run_issue<8, 1>(count);
            case 24: return run_issue<8, 4>(count);
            case 25: return run_issue<64, 1>(count);
            case 26: return run_issue<64, 4>(count);
            case 27: return run_issue<32, 1>(count);
        }
        return 1;
    }
    switch (mode) {
        case 16: return run_pipe2<1, 3, false, false>(iters, mode);
        case 17: return run_pipe2<2, 2, false, false>(iters, mode);
        case 18: return run_pipe2<1, 3, true, false>(iters, mode);
        case 19: return run_pipe2<2, 2, true, false>(iters, mode);
        case 36: return run_pipe2<1, 3, false, true>(iters, mode);
        case 37: return run_pipe2<2, 2, false, true>(iters, mode);
        case 10: return run_pipe<10>(iters);
        case 11: return run_pipe<11>(iters);
        case 12: return run_pipe<12>(iters);
        case 13: return run_pipe<13>(iters);
        case 14: return run_pipe<14>(iters);
        case 15: return run_pipe<15>(iters);
    }
    return 1;
}
