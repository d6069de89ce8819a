// Micro-benchmark of per-block byte histogram variants (development tool, not part of the library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/hist_bench tools/hist_bench.cu && /tmp/hist_bench
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define FULL 0xffffffffu

// variant A: warp-aggregated with match.any; one warp per block, 1 KiB of counters per warp
__global__ void __launch_bounds__(512) hist_match(const uint8_t *__restrict__ src, uint32_t block_size, uint32_t nblocks,
                                                  uint32_t *__restrict__ counts)
{
    extern __shared__ uint32_t sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    uint32_t *cnt = sm + warp * 256;
    for (uint32_t b = blockIdx.x * wpc + warp; b < nblocks; b += gridDim.x * wpc) {
        for (int k = 0; k < 8; k++) cnt[k * 32 + lane] = 0;
        __syncwarp();
        const uint4 *v = reinterpret_cast<const uint4 *>(src + (size_t)b * block_size);
        const uint32_t nvec = block_size >> 4;
        for (uint32_t i = lane; i < nvec; i += 32) {
            uint4 x = __ldg(v + i);
            uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int q = 0; q < 4; q++) {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    uint32_t byte = (w[q] >> (8 * k)) & 0xff;
                    uint32_t m = __match_any_sync(FULL, byte);
                    if ((m & ((1u << lane) - 1)) == 0) cnt[byte] += __popc(m);
                    __syncwarp();
                }
            }
        }
        __syncwarp();
        for (int k = 0; k < 8; k++) counts[(size_t)b * 256 + k * 32 + lane] = cnt[k * 32 + lane];
        __syncwarp();
    }
}

// variant B: lane-private 16-bit counters (16 KiB per warp), 4 bytes per step
__device__ __forceinline__ void hist_word16(uint16_t *cnt, uint32_t w)
{
    uint32_t b0 = w & 0xff, b1 = (w >> 8) & 0xff, b2 = (w >> 16) & 0xff, b3 = w >> 24;
    uint32_t c0 = cnt[b0 << 5], c1 = cnt[b1 << 5], c2 = cnt[b2 << 5], c3 = cnt[b3 << 5];
    uint32_t i1 = (b1 == b0), i2 = (b2 == b0) + (b2 == b1), i3 = (b3 == b0) + (b3 == b1) + (b3 == b2);
    cnt[b0 << 5] = (uint16_t)(c0 + 1);
    cnt[b1 << 5] = (uint16_t)(c1 + 1 + i1);
    cnt[b2 << 5] = (uint16_t)(c2 + 1 + i2);
    cnt[b3 << 5] = (uint16_t)(c3 + 1 + i3);
}
__global__ void __launch_bounds__(512) hist_priv16(const uint8_t *__restrict__ src, uint32_t block_size, uint32_t nblocks,
                                                   uint32_t *__restrict__ counts)
{
    extern __shared__ uint32_t sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    uint16_t *cnt_all = reinterpret_cast<uint16_t *>(sm) + warp * 8192;
    uint16_t *cnt = cnt_all + lane;
    for (uint32_t b = blockIdx.x * wpc + warp; b < nblocks; b += gridDim.x * wpc) {
        for (int k = 0; k < 256; k++) cnt[k << 5] = 0;
        const uint4 *v = reinterpret_cast<const uint4 *>(src + (size_t)b * block_size);
        const uint32_t nvec = block_size >> 4;
        for (uint32_t i = lane; i < nvec; i += 32) {
            uint4 x = __ldg(v + i);
            hist_word16(cnt, x.x); hist_word16(cnt, x.y); hist_word16(cnt, x.z); hist_word16(cnt, x.w);
        }
        __syncwarp();
        for (int k = 0; k < 8; k++) {
            uint32_t bin = k * 32 + lane, s = 0;
            for (int l = 0; l < 32; l++) s += cnt_all[(bin << 5) + ((l + lane) & 31)];
            counts[(size_t)b * 256 + bin] = s;
        }
        __syncwarp();
    }
}

// variant C: lane-private 32-bit counters shared by a CTA of 2 warps (the library's current kernel shape)
__device__ __forceinline__ void hist_word32(uint32_t *cnt, uint32_t w)
{
    uint32_t b0 = w & 0xff, b1 = (w >> 8) & 0xff, b2 = (w >> 16) & 0xff, b3 = w >> 24;
    uint32_t c0 = cnt[b0 << 5], c1 = cnt[b1 << 5], c2 = cnt[b2 << 5], c3 = cnt[b3 << 5];
    uint32_t i1 = (b1 == b0), i2 = (b2 == b0) + (b2 == b1), i3 = (b3 == b0) + (b3 == b1) + (b3 == b2);
    cnt[b0 << 5] = c0 + 1; cnt[b1 << 5] = c1 + 1 + i1; cnt[b2 << 5] = c2 + 1 + i2; cnt[b3 << 5] = c3 + 1 + i3;
}
__global__ void __launch_bounds__(64) hist_priv32(const uint8_t *__restrict__ src, uint32_t block_size, uint32_t nblocks,
                                                  uint32_t *__restrict__ counts)
{
    extern __shared__ uint32_t sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t *cnt = sm + warp * 8192 + lane;
    for (uint32_t b = blockIdx.x; b < nblocks; b += gridDim.x) {
        for (int i = 0; i < 256; i++) cnt[i << 5] = 0;
        const uint4 *v = reinterpret_cast<const uint4 *>(src + (size_t)b * block_size);
        const uint32_t nvec = block_size >> 4;
        for (uint32_t i = tid; i < nvec; i += 64) {
            uint4 x = __ldg(v + i);
            hist_word32(cnt, x.x); hist_word32(cnt, x.y); hist_word32(cnt, x.z); hist_word32(cnt, x.w);
        }
        __syncthreads();
        uint32_t s[4] = {0, 0, 0, 0};
        for (int w = 0; w < 2; w++)
            for (int l = 0; l < 32; l++)
                for (int q = 0; q < 4; q++) s[q] += sm[w * 8192 + ((tid * 4 + q) << 5) + ((l + tid) & 31)];
        for (int q = 0; q < 4; q++) counts[(size_t)b * 256 + tid * 4 + q] = s[q];
        __syncthreads();
    }
}

static uint64_t splitmix64(uint64_t x)
{
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

int main()
{
    const size_t n = 256u << 20;
    const uint32_t bs = 65536, nb = n / bs;
    std::vector<uint8_t> h(n);
    for (int dist = 0; dist < 2; dist++) {
        for (size_t i = 0; i < n; i += 8) {
            uint64_t z = splitmix64(i);
            for (int k = 0; k < 8; k++) {
                uint32_t r = (z >> (8 * k)) & 0xff;
                h[i + k] = dist == 0 ? (uint8_t)(r & 0x3f) : (uint8_t)(__builtin_ctz(r | 0x100));   // 64 symbols / geometric
            }
        }
        uint8_t *d; uint32_t *c, *cref;
        cudaMalloc(&d, n); cudaMalloc(&c, nb * 1024); cudaMalloc(&cref, nb * 1024);
        cudaMemcpy(d, h.data(), n, cudaMemcpyHostToDevice);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaFuncSetAttribute(hist_priv32, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
        cudaFuncSetAttribute(hist_priv16, cudaFuncAttributeMaxDynamicSharedMemorySize, 14 * 16384);
        std::vector<uint32_t> ref(nb * 256), got(nb * 256);
        for (int variant = 0; variant < 4; variant++) {
            float best = 1e9;
            for (int rep = 0; rep < 5; rep++) {
                cudaEventRecord(e0);
                if (variant == 0) hist_priv32<<<148 * 3, 64, 65536>>>(d, bs, nb, cref);
                if (variant == 1) hist_match<<<148, 512, 16 * 1024>>>(d, bs, nb, c);
                if (variant == 2) hist_match<<<148 * 2, 512, 16 * 1024>>>(d, bs, nb, c);
                if (variant == 3) hist_priv16<<<148, 448, 14 * 16384>>>(d, bs, nb, c);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            cudaError_t err = cudaGetLastError();
            cudaMemcpy(ref.data(), cref, nb * 1024, cudaMemcpyDeviceToHost);
            cudaMemcpy(got.data(), variant == 0 ? cref : c, nb * 1024, cudaMemcpyDeviceToHost);
            bool ok = memcmp(ref.data(), got.data(), nb * 1024) == 0;
            printf("dist %d variant %d: %.3f ms  %.0f GB/s  %s %s\n", dist, variant, best, n / best / 1e6, ok ? "ok" : "MISMATCH",
                   err == cudaSuccess ? "" : cudaGetErrorString(err));
        }
        cudaFree(d); cudaFree(c); cudaFree(cref);
    }
    return 0;
}
