// microbench.cuh -- register-only issue-rate probes that define the INT32-pipe roofline of the
// search kernel on the GPU the bench is running on ("peak_pxops = 4 x measured
// VABSDIFF4.U8.ACC warp-instruction issue rate x 32").
#pragma once
#include "common.cuh"

namespace vcs {

constexpr int MB_ACC = 16;      // independent dependency chains per thread
constexpr int MB_UNROLL = 8;    // chain steps per loop iteration
constexpr int MB_THREADS = 256;

// ops issued per thread per loop iteration, per probe
__host__ __device__ constexpr int mb_ops_per_iter(int which) {
    return which == 5 ? 3 * MB_ACC * MB_UNROLL : (which == 6 ? MB_ACC * MB_UNROLL + MB_UNROLL
                                                             : MB_ACC * MB_UNROLL);
}

template <int WHICH>
__global__ void __launch_bounds__(MB_THREADS)
microbench_kernel(uint32_t *out, int iters, uint32_t seed, long long *cycles) {
    __shared__ uint32_t s_buf[MB_THREADS * 2];
    uint32_t acc[MB_ACC];
#pragma unroll
    for (int j = 0; j < MB_ACC; ++j) acc[j] = seed * (threadIdx.x + 1) + 0x9E3779B9u * j;
    uint32_t a = seed ^ (threadIdx.x * 2654435761u), b = ~a * 40503u;
    s_buf[threadIdx.x] = a;
    s_buf[threadIdx.x + MB_THREADS] = b;
    __syncthreads();
    long long t0, t1;
    unsigned long long g0, g1;   // nanoseconds: SM clock = clock64 span / globaltimer span of the SAME block
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0) :: "memory");
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t0) :: "memory");
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < MB_UNROLL; ++u) {
            if (WHICH == 6) {  // one conflict-free LDS.32 feeding MB_ACC VABSDIFF4s
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(a)
                             : "r"((uint32_t)__cvta_generic_to_shared(
                                   &s_buf[(threadIdx.x + ((it + u) & 1) * MB_THREADS)])));
            }
#pragma unroll
            for (int j = 0; j < MB_ACC; ++j) {
                if (WHICH == 0 || WHICH == 6) {
                    asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;"
                                 : "+r"(acc[j]) : "r"(a), "r"(b));
                } else if (WHICH == 1) {
                    asm volatile("add.u32 %0, %0, %1;" : "+r"(acc[j]) : "r"(acc[(j + 5) % MB_ACC]));
                } else if (WHICH == 2) {
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;"
                                 : "+r"(acc[j]) : "r"(acc[(j + 5) % MB_ACC]), "r"(a));
                } else if (WHICH == 3) {
                    asm volatile("mad.lo.u32 %0, %1, %2, %0;"
                                 : "+r"(acc[j]) : "r"(acc[(j + 5) % MB_ACC]), "r"(a));
                } else if (WHICH == 4) {
                    asm volatile("dp4a.u32.u32 %0, %1, %2, %0;"
                                 : "+r"(acc[j]) : "r"(acc[(j + 5) % MB_ACC]), "r"(a));
                } else if (WHICH == 5) {
                    // the wrap8 inner step of me_tiled: t = r1 - cL; z = t ^ r2 ^ cH; acc += bytes(z)
                    uint32_t t, z;
                    asm volatile("sub.u32 %0, %1, %2;" : "=r"(t) : "r"(a), "r"(acc[(j + 5) % MB_ACC]));
                    asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(z) : "r"(t), "r"(b), "r"(a));
                    asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(acc[j]) : "r"(z), "r"(0x01010101u));
                }
            }
        }
    }
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t1) :: "memory");
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1) :: "memory");
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < MB_ACC; ++j) s ^= acc[j];
    out[blockIdx.x * MB_THREADS + threadIdx.x] = s + a;
    if (blockIdx.x == 0 && threadIdx.x == 0) { cycles[0] = t1 - t0; cycles[1] = (long long)(g1 - g0); }
}

}  // namespace vcs
