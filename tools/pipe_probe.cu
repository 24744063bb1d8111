// pipe_probe.cu -- how well do the ALU pipe (LOP3/IADD3) and the FMA pipe (IMAD/IDP.4A) overlap on sm_100a?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_probe tools/pipe_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int NA = 12;   // independent chains per pipe

template <int MODE, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k(uint32_t *out, int iters, uint32_t seed, uint32_t zero) {
    uint32_t x[NA], y[NA], w[NA];
    const uint32_t one = zero + 1;
    uint32_t x2[NA], y2[NA], w2[NA];
    for (int j = 0; j < NA; ++j) { x2[j] = seed + j; y2[j] = seed * j; w2[j] = seed ^ j; }
    for (int j = 0; j < NA; ++j) { x[j] = seed * (threadIdx.x + j + 1); y[j] = ~x[j] * 31u; w[j] = x[j] ^ 99u; }
    const uint32_t a = seed ^ 0x5bd1e995u, b = seed * 77u + zero;
    uint32_t c8[8];
    for (int j = 0; j < 8; ++j) c8[j] = seed * (j + 11) + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int j = 0; j < NA; ++j) {
                if (MODE == 0) {          // LOP3 + IDP.4A, independent
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[j]) : "r"(a), "r"(b));
                    asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(y[j]) : "r"(a), "r"(b));
                } else if (MODE == 1) {   // LOP3 + IMAD
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[j]) : "r"(a), "r"(b));
                    asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(y[j]) : "r"(a), "r"(b));
                } else if (MODE == 2) {   // LOP3 only (2 per step)
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[j]) : "r"(a), "r"(b));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[j]) : "r"(a), "r"(b));
                } else if (MODE == 3) {   // IDP only
                    asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(x[j]) : "r"(a), "r"(b));
                    asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(y[j]) : "r"(a), "r"(b));
                } else if (MODE == 4) {   // LOP3 feeding IDP.4A of the same step (dependent pair)
                    uint32_t z;
                    asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(z) : "r"(x[j]), "r"(a), "r"(b));
                    asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(y[j]) : "r"(z), "r"(0x01010101u));
                } else if (MODE == 5) {   // LOP3 + VABSDIFF4 (both ALU): sanity, expect 0.5/clk
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[j]) : "r"(a), "r"(b));
                    asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(y[j]) : "r"(a), "r"(b));
                } else if (MODE >= 8 && MODE <= 11) {   // the wrap8 step: sub, 3-input xor, byte sum
                    // 8: subtract on FMA (IMAD.IADD); 9: on ALU (IADD3, opaque zero); 10: alternating; 11: 2 of 5 on ALU
                    const bool alu = MODE == 9 || (MODE == 10 && (j & 1)) || (MODE == 11 && (j % 5) < 2);
                    // x[j] evolves (t -> xor -> next t), so nothing is loop invariant
                    if (alu)
                        asm volatile("{\n.reg .u32 t;\nsub.u32 t, %0, %1;\nadd.u32 t, t, %4;\nlop3.b32 %0, t, %2, %3, 0x96;\n}"
                                     : "+r"(x[j]) : "r"(c8[(j + u) & 7]), "r"(a), "r"(b), "r"(zero));
                    else
                        asm volatile("{\n.reg .u32 t;\nsub.u32 t, %0, %1;\nlop3.b32 %0, t, %2, %3, 0x96;\n}"
                                     : "+r"(x[j]) : "r"(c8[(j + u) & 7]), "r"(a), "r"(b));
                    asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(y[j]) : "r"(x[j]), "r"(0x01010101u));
                } else if (MODE == 12) {  // FMA-pipe add (IMAD, opaque multiplier 1) + LOP3
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[j]) : "r"(one), "r"(a));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[j]) : "r"(a), "r"(b));
                } else if (MODE == 13) {  // 3-input add (IADD3, ALU) + IDP.4A
                    asm volatile("{\n.reg .u32 t;\nadd.u32 t, %0, %1;\nadd.u32 %0, t, %2;\n}" : "+r"(x[j]) : "r"(a), "r"(b));
                    asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(y[j]) : "r"(x[(j + 1) % NA]), "r"(0x01010101u));
                } else if (MODE == 14) {  // IADD3 + LOP3 (both ALU)
                    asm volatile("{\n.reg .u32 t;\nadd.u32 t, %0, %1;\nadd.u32 %0, t, %2;\n}" : "+r"(x[j]) : "r"(a), "r"(b));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[j]) : "r"(a), "r"(b));
                } else if (MODE == 15) {  // IADD3 + LOP3 + IDP.4A, independent (2 ALU : 1 FMA)
                    asm volatile("{\n.reg .u32 t;\nadd.u32 t, %0, %1;\nadd.u32 %0, t, %2;\n}" : "+r"(x[j]) : "r"(a), "r"(b));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[j]) : "r"(a), "r"(b));
                    asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(w[j]) : "r"(a), "r"(0x01010101u));
                } else if (MODE == 16) {  // IMAD + LOP3 + IDP.4A, independent (1 ALU : 2 FMA)
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[j]) : "r"(one), "r"(a));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[j]) : "r"(a), "r"(b));
                    asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(w[j]) : "r"(a), "r"(0x01010101u));
                } else if (MODE == 17) {  // 2 LOP3 + 2 IDP.4A with immediate-free operands: 4 chains per j
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[j]) : "r"(a), "r"(b));
                    asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(y[j]) : "r"(a), "r"(0x01010101u));
                } else if (MODE == 18 || MODE == 19) {  // 3 ALU + 3 FMA per j, independent; 18 strictly alternating, 19 grouped
#define I_A asm volatile("{\n.reg .u32 t;\nadd.u32 t, %0, %1;\nadd.u32 %0, t, %2;\n}" : "+r"(x[j]) : "r"(a), "r"(b))
#define I_X1 asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[j]) : "r"(a), "r"(b))
#define I_X2 asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(w[j]) : "r"(a), "r"(b))
#define I_F asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x2[j]) : "r"(one), "r"(a))
#define I_D1 asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(y2[j]) : "r"(a), "r"(0x01010101u))
#define I_D2 asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(w2[j]) : "r"(b), "r"(0x01010101u))
                    if (MODE == 18) { I_A; I_D1; I_X1; I_F; I_X2; I_D2; }
                    else { I_A; I_X1; I_X2; I_F; I_D1; I_D2; }
                } else if (MODE == 6) {   // LOP3 + FFMA (FP32 pipes)
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[j]) : "r"(a), "r"(b));
                    float f = __uint_as_float(y[j]);
                    asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(f) : "f"(__uint_as_float(a)), "f"(__uint_as_float(b)));
                    y[j] = __float_as_uint(f);
                }
            }
        }
    }
    uint32_t s = 0;
    for (int j = 0; j < NA; ++j) s ^= x[j] + y[j] + (MODE == 15 || MODE == 16 ? w[j] : 0) + (MODE >= 18 ? w[j] + x2[j] + y2[j] + w2[j] : 0);
    out[blockIdx.x * THREADS + threadIdx.x] = s;
}

template <int MODE, int THREADS>
void run(const char *name, int sms, uint32_t *d_out, double clk) {
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE, THREADS><<<sms, THREADS>>>(d_out, iters / 4, 77u, 0);
    cudaEventRecord(e0);
    k<MODE, THREADS><<<sms, THREADS>>>(d_out, iters, 77u, 0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double instr = (double)sms * (THREADS / 32) * iters * 4.0 * NA * ((MODE >= 8 && MODE <= 11) || MODE == 15 || MODE == 16 ? 3 : (MODE >= 18 ? 6 : 2));
    printf("%-28s %4d thr: %.3f warp-instr/clk/SMSP\n", name, THREADS, instr / (ms * 1e-3) / (clk * sms * 4));
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double clk = clk_khz * 1e3;
    const int sms = p.multiProcessorCount;
    uint32_t *d_out; cudaMalloc(&d_out, (size_t)sms * 1024 * 4);
#define R(M, N) run<M, 512>(N, sms, d_out, clk); run<M, 1024>(N, sms, d_out, clk)
    R(0, "LOP3 + IDP.4A independent");
    R(1, "LOP3 + IMAD independent");
    R(2, "LOP3 only");
    R(3, "IDP.4A only");
    R(4, "LOP3 -> IDP.4A dependent");
    R(5, "LOP3 + VABSDIFF4");
    R(6, "LOP3 + FFMA");
    R(12, "IMAD(add) + LOP3");
    R(13, "IADD3 + IDP.4A(imm)");
    R(14, "IADD3 + LOP3");
    R(15, "IADD3 + LOP3 + IDP.4A");
    R(16, "IMAD + LOP3 + IDP.4A");
    R(17, "LOP3 + IDP.4A(imm) indep");
    R(18, "3 ALU + 3 FMA alternating");
    R(19, "3 ALU + 3 FMA grouped");
    R(8, "wrap8 step, sub on FMA");
    R(9, "wrap8 step, sub on ALU");
    R(10, "wrap8 step, sub 1:1");
    R(11, "wrap8 step, sub 2:3");
    printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
