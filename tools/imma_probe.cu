// imma_probe.cu -- can IMMA.16832.U8.U8 (mma.sync m16n8k32, B = word-selecting ones) replace the
// 4 IDP.4A byte sums of the wrap8 search loop?  Checks the fragment mapping and measures issue rates.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/imma_probe tools/imma_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define MMA(c, a0, a1, a2, a3, b0, b1)                                                              \
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "   \
                 "{%8,%9}, {%0,%1,%2,%3};"                                                          \
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])                                   \
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1))

__global__ void map_kernel(const uint32_t *in, int *out) {   // in[lane*4+i] = a_i of the lane
    const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
    const uint32_t b0 = g == t ? 0x01010101u : 0, b1 = g == 4 + t ? 0x01010101u : 0;
    int c[4] = {0, 0, 0, 0};
    MMA(c, in[lane * 4], in[lane * 4 + 1], in[lane * 4 + 2], in[lane * 4 + 3], b0, b1);
    for (int i = 0; i < 4; ++i) out[lane * 4 + i] = c[i];
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) rate_kernel(uint32_t *out, int iters, uint32_t seed, uint32_t zero) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const uint32_t b0 = g == t ? 0x01010101u : 0, b1 = g == 4 + t ? 0x01010101u : 0;
    constexpr int NT = 8;       // accumulator tiles (32 accumulators)
    int c[NT][4];
    uint32_t acc[NT * 4];
    for (int j = 0; j < NT; ++j) for (int i = 0; i < 4; ++i) { c[j][i] = 0; acc[j * 4 + i] = 0; }
    uint32_t r1v[4], r2v[4];
    for (int i = 0; i < 4; ++i) { r1v[i] = (seed * (threadIdx.x + 1 + 977 * i)) | 0x80808080u; r2v[i] = ~(r1v[i] * 31u) & 0x80808080u; }
    uint32_t cl[8], ch[8];
    for (int i = 0; i < 8; ++i) { cl[i] = (seed * (i + 3) * 2654435761u) & 0x7f7f7f7fu; ch[i] = (seed >> i) & 0x80808080u; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                uint32_t z[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int v = (i + 4 * (j >> 2)) & 7;
                    const uint32_t r1 = r1v[j & 3], r2 = r2v[j & 3];
                    if (MODE == 2) { z[i] = cl[v]; continue; }           // IMMA only
                    if ((j * 4 + i) % 5 < 2)
                        asm volatile("{\n.reg .u32 t;\nsub.u32 t, %1, %2;\nadd.u32 t, t, %5;\nlop3.b32 %0, t, %3, %4, 0x96;\n}"
                                     : "=r"(z[i]) : "r"(r1), "r"(cl[v]), "r"(r2), "r"(ch[v]), "r"(zero));
                    else
                        asm volatile("{\n.reg .u32 t;\nsub.u32 t, %1, %2;\nlop3.b32 %0, t, %3, %4, 0x96;\n}"
                                     : "=r"(z[i]) : "r"(r1), "r"(cl[v]), "r"(r2), "r"(ch[v]));
                }
                if (MODE == 0) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[j * 4 + i] = __dp4a(z[i], 0x01010101u, acc[j * 4 + i]);
                } else {
                    MMA(c[j], z[0], z[1], z[2], z[3], b0, b1);
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) { r1v[i] += 0x01010101u; r1v[i] |= 0x80808080u; }
        }
    }
    uint32_t s = 0;
    for (int j = 0; j < NT; ++j) for (int i = 0; i < 4; ++i) s += c[j][i] + acc[j * 4 + i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
double run(int iters, int sms, uint32_t *d_out) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    rate_kernel<MODE><<<sms, 512>>>(d_out, iters / 4, 77u, 0);
    cudaEventRecord(e0);
    rate_kernel<MODE><<<sms, 512>>>(d_out, iters, 77u, 0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double words = (double)sms * 16 /*warps*/ * iters * 2 * 8 * 4;   // warp-level word steps
    return words / (ms * 1e-3);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    uint32_t h_in[128]; int h_out[128];
    for (int l = 0; l < 32; ++l) for (int i = 0; i < 4; ++i) h_in[l * 4 + i] = (uint32_t)(l * 4 + i + 1);   // bytesum = value (< 256)
    uint32_t *d_in; int *d_o; uint32_t *d_out;
    cudaMalloc(&d_in, 512); cudaMalloc(&d_o, 512); cudaMalloc(&d_out, (size_t)p.multiProcessorCount * 512 * 4);
    cudaMemcpy(d_in, h_in, 512, cudaMemcpyHostToDevice);
    map_kernel<<<1, 32>>>(d_in, d_o);
    cudaMemcpy(h_out, d_o, 512, cudaMemcpyDeviceToHost);
    // expectation: lane (g,t) c0,c1 = words 2t,2t+1 of row g; c2,c3 = same of row g+8;
    // word n of row g   = a0 (n<4) / a2 (n>=4) of lane 4g + n%4;  row g+8 -> a1 / a3
    int bad = 0;
    for (int l = 0; l < 32; ++l) for (int i = 0; i < 4; ++i) {
        const int g = l >> 2, t = l & 3, n = 2 * t + (i & 1), hi = i >> 1;
        const int src_lane = 4 * g + (n & 3), src_reg = (n < 4 ? 0 : 2) + hi;
        const int want = src_lane * 4 + src_reg + 1;
        if (h_out[l * 4 + i] != want) { if (bad < 8) printf("lane %d c%d = %d want %d\n", l, i, h_out[l * 4 + i], want); ++bad; }
    }
    printf("mapping %s\n", bad ? "MISMATCH" : "ok");
    const int iters = 20000, sms = p.multiProcessorCount;
    const double clk = clk_khz * 1e3;
    double r0 = run<0>(iters, sms, d_out), r1 = run<1>(iters, sms, d_out), r2 = run<2>(iters, sms, d_out);
    printf("clock %.0f MHz (attr)  SMs %d\n", clk / 1e6, sms);
    printf("dp4a triple : %.3e warp-words/s = %.3f clk/word/SMSP\n", r0, clk * sms * 4 / r0);
    printf("imma (2+1/4): %.3e warp-words/s = %.3f clk/word/SMSP\n", r1, clk * sms * 4 / r1);
    printf("imma only   : %.3e warp-words/s = %.3f clk/IMMA/SMSP\n", r2, clk * sms * 4 / r2 * 4);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(e));
    return bad != 0;
}
