// fse_decode64w.cuh -- the per-warp layout of the decoders that keep full DecodeTransform entries (fse.rs:260-265: one
// 32-bit load per state gives new_state, symbol and num_bits; k_decode128_blocks, table_log 13) in a 4*size + 2 KiB
// footprint: the spread lives in the last quarter of the table region and is consumed in place by the table build (entry c
// only overwrites spread cells <= c), the build scratch is aliased with the payload ring.  The 64-state kernel that used
// it lost to the compact layout (0.70 vs 0.49 ms on c2) and was removed in round 2.
#pragma once
#include "fse_decode64c.cuh"

namespace fsed {

struct Dec64wLayout {
    uint32_t tab;      // uint32[size]
    uint32_t sym;      // uint8[size], aliased with tab's last quarter
    uint32_t scratch;  // build: norm i32[256], then the counters u32[256] in the same array; decode: ring u32[258]
    uint32_t total;
};
__host__ __device__ inline Dec64wLayout dec64w_layout(uint32_t tlmax)
{
    Dec64wLayout l;
    uint32_t size = 1u << tlmax;
    l.tab = 0;
    l.sym = size * 3;
    l.scratch = size * 4;
    l.total = size * 4 + 1040;
    return l;
}

}  // namespace fsed
