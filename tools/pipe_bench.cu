// pipe_bench.cu -- issue rates of the integer instructions the coders are made of (development tool).
// Each kernel runs ITER x 8 independent chains per thread of one instruction kind (or a mix); prints warp
// instructions per clock per SM at 8, 16 and 32 warps per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 4096
template <int KIND>
__global__ void k(uint32_t *out, uint32_t a, uint32_t b)
{
    uint32_t x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 8 + i + a;
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (KIND == 0) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));            // SHF
            if (KIND == 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));                // IMAD
            if (KIND == 2) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(a));                            // IMAD.HI
            if (KIND == 3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(a), "r"(b));            // LOP3
            if (KIND == 4) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));                  // PRMT
            if (KIND == 5) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(a));                               // IADD3 / IMAD.IADD
            if (KIND == 6) {                                                                                         // SHF + IMAD alternating
                if (i & 1) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
                else asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
            }
            if (KIND == 7) {                                                                                         // LOP3 + IMAD.HI alternating
                if (i & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(a), "r"(b));
                else asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(a));
            }
            if (KIND == 8) asm volatile("shl.b32 %0, %0, %1;" : "+r"(x[i]) : "r"(b));                               // SHL variable
            if (KIND == 9) {                                                                                         // 3 ALU : 1 IMAD
                if ((i & 3) == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
                else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(a), "r"(b));
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= x[i];
    if (s == 0x12345) out[0] = s;
}

template <int KIND> void run(const char *name, uint32_t *d, int clock_khz)
{
    printf("%-28s", name);
    for (int warps : {8, 16, 32}) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        k<KIND><<<148, warps * 32>>>(d, 3, 5);
        cudaEventRecord(e0);
        k<KIND><<<148, warps * 32>>>(d, 3, 5);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double clocks = ms * 1e-3 * clock_khz * 1e3;
        double winst = (double)ITER * 8 * warps;     // warp instructions per SM
        printf("  %2dw: %.2f/clk/SM", warps, winst / clocks);
    }
    printf("\n");
}

int main()
{
    uint32_t *d; cudaMalloc(&d, 64);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = p.clockRate;
    printf("%s, %d SMs, %d kHz (nominal; rates assume this clock)\n", p.name, p.multiProcessorCount, khz);
    run<0>("SHF (funnel, variable)", d, khz);
    run<8>("SHL (variable)", d, khz);
    run<1>("IMAD", d, khz);
    run<2>("IMAD.HI", d, khz);
    run<3>("LOP3", d, khz);
    run<4>("PRMT", d, khz);
    run<5>("ADD", d, khz);
    run<6>("SHF + IMAD 1:1", d, khz);
    run<7>("LOP3 + IMAD.HI 1:1", d, khz);
    run<9>("LOP3 + IMAD 3:1", d, khz);
    return 0;
}
