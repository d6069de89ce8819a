// fse_zstd_norm.cuh -- a second normaliser: libzstd's FSE_normalizeCount / FSE_normalizeM2 (lib/compress/fse_compress.c,
// zstd 1.5.x), SURVEY 8f f3.  The crate writes zstd's NCount header format (src/histogram.rs:342) but normalises slightly
// differently (`to_distribute != 0 &&` guard at :144, low-probability symbols always -1, table_log 5..15); tables
// normalised HERE are the ones libzstd's entropy stage would build for the same counts.  One thread per table: every
// branch depends on running totals, and the work is a few hundred operations.  Not on the reference's path: checked
// against the oracle's restatement and hand-derived vectors only (no libzstd with FSE symbols exists in this image).
#pragma once
#include "fse_device.cuh"

namespace fsed {

__device__ int zstd_normalize_m2(int32_t *norm, uint32_t table_log, const unsigned long long *count, uint64_t total, uint32_t max_symbol,
                                 int32_t low_prob_count)
{
    const int32_t NOT_YET_ASSIGNED = -2;
    uint32_t s, distributed = 0, to_distribute;
    const uint64_t low_threshold = total >> table_log;
    uint64_t low_one = (total * 3) >> (table_log + 1);
    for (s = 0; s <= max_symbol; s++) {
        const uint64_t c = count[s];
        if (c == 0) { norm[s] = 0; continue; }
        if (c <= low_threshold) { norm[s] = low_prob_count; distributed++; total -= c; continue; }
        if (c <= low_one) { norm[s] = 1; distributed++; total -= c; continue; }
        norm[s] = NOT_YET_ASSIGNED;
    }
    to_distribute = (1u << table_log) - distributed;
    if (to_distribute == 0) return 0;
    if ((total / to_distribute) > low_one) {
        low_one = (total * 3) / ((uint64_t)to_distribute * 2);
        for (s = 0; s <= max_symbol; s++)
            if (norm[s] == NOT_YET_ASSIGNED && count[s] <= low_one) { norm[s] = 1; distributed++; total -= count[s]; }
        to_distribute = (1u << table_log) - distributed;
    }
    if (distributed == max_symbol + 1) {                      // every symbol is poor: the remainder goes to the largest count
        uint32_t max_v = 0;
        uint64_t max_c = 0;
        for (s = 0; s <= max_symbol; s++)
            if (count[s] > max_c) { max_v = s; max_c = count[s]; }
        norm[max_v] += (int32_t)to_distribute;
        return 0;
    }
    if (total == 0) {                                         // round robin over the symbols that hold a point
        for (s = 0; to_distribute > 0; s = (s + 1) % (max_symbol + 1))
            if (norm[s] > 0) { to_distribute--; norm[s]++; }
        return 0;
    }
    const uint64_t v_step_log = 62 - (uint64_t)table_log;
    const uint64_t mid = (1ull << (v_step_log - 1)) - 1;
    const uint64_t r_step = (((1ull << v_step_log) * to_distribute) + mid) / total;
    uint64_t tmp_total = mid;
    for (s = 0; s <= max_symbol; s++) {
        if (norm[s] == NOT_YET_ASSIGNED) {
            const uint64_t end = tmp_total + count[s] * r_step;
            const uint32_t weight = (uint32_t)(end >> v_step_log) - (uint32_t)(tmp_total >> v_step_log);
            if (weight < 1) return ST_PANIC;
            norm[s] = (int32_t)weight;
            tmp_total = end;
        }
    }
    return 0;
}

// status: 0, 3 (one symbol holds every count: zstd returns 0 = "rle", norm is all zero), ST_TABLE_LOG (> 12), ST_PANIC (GENERIC)
__global__ void k_normalize_zstd(const unsigned long long *__restrict__ counts64, uint32_t ntables, uint32_t req_log2, int use_low_prob_count,
                                 int32_t *norm_out, uint32_t *log2_out, uint32_t *table_len_out, int32_t *status)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntables) return;
    const unsigned long long *count = counts64 + (size_t)t * 256;
    int32_t *norm = norm_out + (size_t)t * 256;
    uint64_t total = 0;
    int hi = -1;
    for (int i = 0; i < 256; i++) {
        norm[i] = 0;
        total += count[i];
        if (count[i]) hi = i;
    }
    uint32_t table_log = req_log2 ? req_log2 : 11u;          // FSE_DEFAULT_TABLELOG
    log2_out[t] = table_log;
    table_len_out[t] = (uint32_t)(hi < 0 ? 0 : hi) + 1;
    if (total == 0) { status[t] = ST_PANIC; return; }
    const uint32_t max_symbol = (uint32_t)hi;
    if (table_log < 5) { status[t] = ST_PANIC; return; }     // FSE_MIN_TABLELOG
    if (table_log > 12) { status[t] = ST_TABLE_LOG; return; } // FSE_MAX_TABLELOG
    const uint32_t min_src = ilog2u64(total) + 1, min_sym = (max_symbol ? ilog2u(max_symbol) : 0u) + 2;
    if (table_log < min(min_src, min_sym)) { status[t] = ST_PANIC; return; }   // FSE_minTableLog
    const int32_t low_prob_count = use_low_prob_count ? -1 : 1;
    const uint64_t scale = 62 - (uint64_t)table_log;
    const uint64_t step = (1ull << 62) / total;
    const uint64_t v_step = 1ull << (scale - 20);
    long long still = 1ll << table_log;
    uint32_t largest = 0;
    int32_t largest_p = 0;
    const uint64_t low_threshold = total >> table_log;
    for (uint32_t s = 0; s <= max_symbol; s++) {
        const uint64_t c = count[s];
        if (c == total) {
            for (uint32_t k = 0; k < s; k++) norm[k] = 0;
            status[t] = 3;
            return;
        }
        if (c == 0) continue;
        if (c <= low_threshold) { norm[s] = low_prob_count; still--; }
        else {
            int32_t proba = (int32_t)((c * step) >> scale);
            if (proba < 8) {
                const uint64_t rest_to_beat = v_step * RTB_TABLE[proba];
                proba += ((c * step) - ((uint64_t)proba << scale) > rest_to_beat) ? 1 : 0;
            }
            if (proba > largest_p) { largest_p = proba; largest = s; }
            norm[s] = proba;
            still -= proba;
        }
    }
    int rc = 0;
    if (-still >= (long long)(norm[largest] >> 1)) rc = zstd_normalize_m2(norm, table_log, count, total, max_symbol, low_prob_count);
    else norm[largest] += (int32_t)still;
    status[t] = rc;
}

}  // namespace fsed
