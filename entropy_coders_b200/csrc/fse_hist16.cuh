// fse_hist16.cuh -- Histogram::new per block (src/histogram.rs:18-66), one warp per block.
//
// Every lane owns a private column of 256 16-bit counters (16 KiB per warp, 14 warps per SM): no
// atomics (shared-memory atomics cost 2 cycles per lane here) and no cross-lane hazards.  Four bytes
// are counted per step; the increments of equal bytes inside the step are resolved in registers
// (later stores win, and they carry the larger count).  Measured 1.37 TB/s on 256 MiB against
// 0.64-0.97 TB/s for 32-bit columns (6 warps per SM) and 0.16-0.7 TB/s for match.any aggregation
// (tools/hist_bench.cu).  A lane sees block_size/32 + 30 bytes, so blocks up to 1 MiB cannot overflow.
// Lanes 2k and 2k+1 share a bank in this layout (two wavefronts per access: ncu 1.89 / 2.00).  Round 2 measured the
// conflict-free alternative (bins 2j / 2j+1 as the halves of the lane's own word at j * 128 + lane * 4: one wavefront
// per access, two more address instructions per byte): SLOWER, 4.94 vs 4.39 ms on 8 GiB geometric and 0.646 vs 0.519 ms
// on 1 GiB few-symbol: the kernel is bound by instruction issue and latency before shared-memory wavefronts.
#pragma once
#include "fse_device.cuh"

namespace fsed {

constexpr int HIST16_WARPS = 14;
constexpr int HIST16_SMEM = HIST16_WARPS * 256 * 32 * 2;
constexpr uint32_t HIST16_MAX_BLOCK = 1u << 20;

__device__ __forceinline__ void hist16_word(uint16_t *cnt, uint32_t w)
{
    // byte offsets of the four counters inside the lane's column (bin * 64): extracted once, also used as the keys
    // of the duplicate test
    const uint32_t a0 = (w << 6) & 0x3fc0u, a1 = (w >> 2) & 0x3fc0u, a2 = (w >> 10) & 0x3fc0u, a3 = (w >> 18) & 0x3fc0u;
    uint8_t *col = reinterpret_cast<uint8_t *>(cnt);
    uint16_t *p0 = reinterpret_cast<uint16_t *>(col + a0), *p1 = reinterpret_cast<uint16_t *>(col + a1);
    uint16_t *p2 = reinterpret_cast<uint16_t *>(col + a2), *p3 = reinterpret_cast<uint16_t *>(col + a3);
    uint32_t c0 = *p0, c1 = *p1, c2 = *p2, c3 = *p3;
    uint32_t i1 = (a1 == a0), i2 = (a2 == a0) + (a2 == a1), i3 = (a3 == a0) + (a3 == a1) + (a3 == a2);
    *p0 = (uint16_t)(c0 + 1);
    *p1 = (uint16_t)(c1 + 1 + i1);
    *p2 = (uint16_t)(c2 + 1 + i2);
    *p3 = (uint16_t)(c3 + 1 + i3);
}

__global__ void __launch_bounds__(HIST16_WARPS * 32)
k_hist_blocks16(const uint8_t *__restrict__ src, size_t n, uint32_t block_size, uint32_t nblocks,
                uint32_t *__restrict__ counts, uint32_t *__restrict__ table_len)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    uint16_t *cnt_all = reinterpret_cast<uint16_t *>(smem_raw) + warp * 8192;
    uint16_t *cnt = cnt_all + lane;                       // counter(bin) = cnt[bin << 5]
    for (uint32_t b = blockIdx.x * wpc + warp; b < nblocks; b += gridDim.x * wpc) {
        {
            uint4 *z = reinterpret_cast<uint4 *>(cnt_all);
            for (int i = lane; i < 1024; i += 32) z[i] = make_uint4(0, 0, 0, 0);
        }
        __syncwarp();
        const size_t off = (size_t)b * block_size;
        const uint32_t bn = (uint32_t)min((size_t)block_size, n - off);
        const uint8_t *p = src + off;
        uint32_t head = (uint32_t)((16 - ((uintptr_t)p & 15)) & 15);
        if (head > bn) head = bn;
        const uint32_t nvec = (bn - head) >> 4;
        const uint32_t tail0 = head + (nvec << 4);
        if ((uint32_t)lane < head) cnt[(uint32_t)p[lane] << 5] += 1;          // head, tail < 16 bytes
        if (tail0 + lane < bn) cnt[(uint32_t)p[tail0 + lane] << 5] += 1;
        const uint4 *v = reinterpret_cast<const uint4 *>(p + head);
        uint32_t i = lane;
        for (; i + 96 < nvec; i += 128) {                                      // four loads in flight
            uint4 x0 = __ldg(v + i), x1 = __ldg(v + i + 32), x2 = __ldg(v + i + 64), x3 = __ldg(v + i + 96);
            hist16_word(cnt, x0.x); hist16_word(cnt, x0.y); hist16_word(cnt, x0.z); hist16_word(cnt, x0.w);
            hist16_word(cnt, x1.x); hist16_word(cnt, x1.y); hist16_word(cnt, x1.z); hist16_word(cnt, x1.w);
            hist16_word(cnt, x2.x); hist16_word(cnt, x2.y); hist16_word(cnt, x2.z); hist16_word(cnt, x2.w);
            hist16_word(cnt, x3.x); hist16_word(cnt, x3.y); hist16_word(cnt, x3.z); hist16_word(cnt, x3.w);
        }
        for (; i < nvec; i += 32) {
            uint4 x = __ldg(v + i);
            hist16_word(cnt, x.x); hist16_word(cnt, x.y); hist16_word(cnt, x.z); hist16_word(cnt, x.w);
        }
        __syncwarp();
        // merge the 32 columns: lane sums bins k*32+lane, columns rotated so that a step hits 32 banks
        int hi = -1;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            uint32_t bin = k * 32 + lane, s = 0;
#pragma unroll 8
            for (int l = 0; l < 32; l++) s += cnt_all[(bin << 5) + ((l + lane) & 31)];
            counts[(size_t)b * 256 + bin] = s;
            if (s) hi = (int)bin;
        }
        if (table_len) {                                   // histogram.rs:52-59
#pragma unroll
            for (int d = 16; d; d >>= 1) hi = max(hi, __shfl_xor_sync(FULL, hi, d));
            if (lane == 0) table_len[b] = (uint32_t)(hi < 0 ? 0 : hi) + 1;
        }
        __syncwarp();
    }
}

}  // namespace fsed
