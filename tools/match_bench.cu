// microbenchmark: __match_any_sync against an 8-ballot construction of the same mask (8-bit keys)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t match8(uint32_t s)
{
    uint32_t m = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        uint32_t b = __ballot_sync(0xffffffffu, (s >> k) & 1);
        m &= ((s >> k) & 1) ? b : ~b;
    }
    return m;
}
template <int MODE> __global__ void k(const uint8_t *in, uint32_t *out, int iters)
{
    uint32_t acc = 0, s = in[threadIdx.x + blockIdx.x * blockDim.x];
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        uint32_t m = MODE == 0 ? __match_any_sync(0xffffffffu, s) : match8(s);
        acc += __popc(m);
        s = (s * 5 + acc) & 0xff;
    }
    long long t1 = clock64();
    out[threadIdx.x + blockIdx.x * blockDim.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[gridDim.x * blockDim.x] = (uint32_t)(t1 - t0);
}
int main()
{
    const int iters = 4096;
    uint8_t *in; uint32_t *out;
    cudaMalloc(&in, 1 << 20); cudaMemset(in, 7, 1 << 20); cudaMalloc(&out, (1 << 22) + 4);
    for (int warps : {1, 4, 8, 16, 28}) {
        for (int mode = 0; mode < 2; mode++) {
            int threads = 32 * (warps > 16 ? warps / 2 : warps), blocks = 148 * (warps > 16 ? 2 : 1);
            cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
            for (int rep = 0; rep < 2; rep++) {
                cudaEventRecord(a);
                if (mode == 0) k<0><<<blocks, threads>>>(in, out, iters); else k<1><<<blocks, threads>>>(in, out, iters);
                cudaEventRecord(b); cudaEventSynchronize(b);
            }
            float ms; cudaEventElapsedTime(&ms, a, b);
            uint32_t cyc; cudaMemcpy(&cyc, out + blocks * threads, 4, cudaMemcpyDeviceToHost);
            printf("warps/SM %2d %s: %.1f cycles per op per warp (chain), %.2f ops/cycle/SM\n", warps, mode ? "8 ballots " : "match.any ",
                   (double)cyc / iters, (double)warps * iters / cyc);
        }
    }
    return 0;
}
