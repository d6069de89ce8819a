// fse_device.cuh -- warp-level building blocks of the B200 FSE (tANS) path.
//
// One warp owns one table / one block; everything here is warp-synchronous (no __syncthreads),
// works out of a per-warp slice of shared memory, and never uses shared-memory atomics (measured
// at 2 cycles per lane on this architecture).  Reference citations are relative to the crate
// root of Cognoscan/entropy_coders.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fsed {

constexpr uint32_t FULL = 0xffffffffu;
constexpr int TL_MIN = 5, TL_MAX = 15, TL_DEFAULT = 11;  // src/lib.rs:9-12
constexpr int ROW_STRIDE = 17;                             // words per lane row (odd => bank staggered)
constexpr int ROWS_WORDS = 32 * ROW_STRIDE;
constexpr int HDR_RESERVE = 512;                           // bytes reserved for a header in block scratch

// status codes mirrored from include/fse_b200.h
constexpr int ST_OK = 0, ST_CAPACITY = -2, ST_TABLE_LOG = -3, ST_TOO_MANY = -4, ST_IO = -5, ST_NO_MARKER = -6,
              ST_LENGTH = -7, ST_PANIC = -8, ST_UNSUPPORTED = -9;

__device__ __forceinline__ uint32_t ilog2u(uint32_t v) { return 31u - (uint32_t)__clz((int)v); }
__device__ __forceinline__ uint32_t ilog2u64(uint64_t v) { return 63u - (uint32_t)__clzll((long long)v); }
__device__ __forceinline__ uint32_t lt_mask(int lane) { return (1u << lane) - 1u; }

template <typename T>
__device__ __forceinline__ T warp_incl_add(T v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T t = __shfl_up_sync(FULL, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// Inclusive prefix sum of a value < 32 per lane from five ballots: five independent VOTE + POPC
// pairs instead of a chain of five dependent shuffles (the scan sits on the decoder's critical path).
__device__ __forceinline__ uint32_t warp_incl_add5(uint32_t v, int lane)
{
    const uint32_t le = 0xffffffffu >> (31 - lane);
    uint32_t b0 = __ballot_sync(FULL, v & 1u), b1 = __ballot_sync(FULL, v & 2u), b2 = __ballot_sync(FULL, v & 4u);
    uint32_t b3 = __ballot_sync(FULL, v & 8u), b4 = __ballot_sync(FULL, v & 16u);
    return __popc(b0 & le) + 2u * __popc(b1 & le) + 4u * __popc(b2 & le) + 8u * __popc(b3 & le) + 16u * __popc(b4 & le);
}

__device__ __forceinline__ int warp_incl_max(int v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(FULL, v, d);
        if (lane >= d) v = max(v, t);
    }
    return v;
}

// histogram.rs:100
__constant__ uint32_t RTB_TABLE[8] = {0, 473195, 504333, 520860, 550000, 700000, 750000, 830000};

// fse.rs:68-70
__device__ __forceinline__ uint32_t table_step(uint32_t size) { return size * 5u / 8u + 3u; }

// ------------------------------------------------------------------------------------------
// Lane-private LSB-first bit string in a shared-memory row (the device analogue of
// BitStackWriter's accumulator, src/bitstream/writer.rs:163-180: OR in `val << bits`, bits leave
// little endian).  Up to 16 full words + a remainder word.
// ------------------------------------------------------------------------------------------
struct BitRow {
    uint32_t *row;
    uint32_t lo, hi, pos, nw;
    __device__ __forceinline__ void init(uint32_t *r, uint32_t carry_word, uint32_t carry_bits)
    {
        row = r; lo = carry_word; hi = 0; pos = carry_bits; nw = 0;
    }
    // v < 2^nb, nb <= 16
    __device__ __forceinline__ void put(uint32_t v, uint32_t nb)
    {
        lo |= v << pos;
        hi = __funnelshift_l(v, 0u, pos);  // v >> (32 - pos); 0 when pos == 0
        pos += nb;
        if (pos >= 32) { row[nw] = lo; nw++; lo = hi; pos -= 32; }
    }
    __device__ __forceinline__ uint32_t finish()
    {
        row[nw] = lo;
        return nw * 32 + pos;
    }
};

// ------------------------------------------------------------------------------------------
// Concatenate the 32 lane strings (lane order == stream order) into 32-bit words at `dst`
// (word aligned; lane 0's string already starts with the bits carried from the previous call, so
// the concatenation starts on a word boundary).  Full words are stored by the lane whose string
// holds the word's last bit; the trailing partial word is returned as the new carry.
// Boundary words shared by several lanes are combined with one segmented scan instead of atomics:
// lane L maps the partial word content c entering it to  k ? v : (c | v).
// Returns the number of full words stored.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t warp_place(const uint32_t *row, uint32_t tot, uint32_t *dst, uint32_t dst_cap_words,
                                               int lane, uint32_t &carry_word, uint32_t &carry_bits, bool &overflow)
{
    uint32_t incl = warp_incl_add(tot, lane);
    uint32_t T = __shfl_sync(FULL, incl, 31);
    uint32_t p = incl - tot, e = incl;
    uint32_t s = p & 31, W = p >> 5, We = e >> 5;
    uint32_t nfull = We - W;  // words whose last bit lies in my string
    uint32_t r0 = tot ? row[0] : 0u;
    uint32_t x0 = r0 << s;
    uint32_t k = nfull ? 1u : 0u, v;
    if (k) {  // my bits in word We (the word my string ends in)
        uint32_t a = (nfull * 32 < tot) ? row[nfull] : 0u;
        uint32_t b = row[nfull - 1];
        v = __funnelshift_l(b, a, s);  // (a << s) | (b >> (32 - s))
    } else {
        v = x0;
    }
    // inclusive segmented scan of (k, v); a earlier, b later:  b.k ? b : (a.k, a.v | b.v)
    uint32_t sk = k, sv = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t ok = __shfl_up_sync(FULL, sk, d);
        uint32_t ov = __shfl_up_sync(FULL, sv, d);
        if (lane >= d && !sk) { sk = ok; sv |= ov; }
    }
    uint32_t cin = __shfl_up_sync(FULL, sv, 1);
    if (lane == 0) cin = 0;
    uint32_t new_carry = __shfl_sync(FULL, sv, 31);

    if (W + nfull > dst_cap_words) { overflow = true; nfull = 0; }
    if (nfull) {
        dst[W] = cin | x0;
        uint32_t prev = r0;
        uint32_t j = 1;
        for (; j + 4 <= nfull; j += 4) {                       // four words per trip: the loop overhead was 2/3 of it
            uint32_t c0 = row[j], c1 = row[j + 1], c2 = row[j + 2], c3 = row[j + 3];
            dst[W + j] = __funnelshift_l(prev, c0, s);
            dst[W + j + 1] = __funnelshift_l(c0, c1, s);
            dst[W + j + 2] = __funnelshift_l(c1, c2, s);
            dst[W + j + 3] = __funnelshift_l(c2, c3, s);
            prev = c3;
        }
        for (; j < nfull; j++) {
            uint32_t cur = row[j];
            dst[W + j] = __funnelshift_l(prev, cur, s);
            prev = cur;
        }
    }
    overflow = __any_sync(FULL, overflow);
    carry_word = new_carry;
    carry_bits = T & 31;
    return T >> 5;
}

// ------------------------------------------------------------------------------------------
// Histogram::optimal_log2, src/histogram.rs:264-277.  Caller guarantees size > 4, table_len > 1.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t optimal_log2(uint64_t size, uint32_t table_len)
{
    uint32_t min_bits_src = ilog2u64(size) + 1;
    uint32_t min_bits_symbols = ilog2u(table_len - 1) + 2;
    uint32_t min_bits = min(min_bits_src, min_bits_symbols);
    uint32_t max_bits = ilog2u64(size - 1) - 2;
    uint32_t v = min((uint32_t)TL_DEFAULT, max_bits);
    v = max(v, min_bits);
    return min(max(v, (uint32_t)TL_MIN), (uint32_t)TL_MAX);
}

// Histogram::normalize_slow, src/histogram.rs:157-261.  Serial (one lane), rare.
template <typename CT>
__device__ int normalize_slow_serial(const CT *counts, uint64_t size, uint32_t table_len, uint32_t log2, int32_t *table)
{
    const int32_t UNASSIGNED = -2;
    uint64_t low_threshold = size >> log2;
    uint64_t low_one = (size * 3) >> (log2 + 1);
    uint64_t to_distribute = 1ull << log2;
    uint64_t total = size;
    for (uint32_t i = 0; i < 256; i++) table[i] = 0;
    for (uint32_t i = 0; i < table_len; i++) {
        uint64_t t = counts[i];
        if (t == 0) continue;
        if (t <= low_threshold) { table[i] = -1; to_distribute -= 1; total -= t; }
        else if (t <= low_one) { table[i] = 1; to_distribute -= 1; total -= t; }
        else table[i] = UNASSIGNED;
    }
    if (to_distribute == 0) return 1;
    if ((total / to_distribute) > low_one) {
        uint64_t low = (total * 3) / (to_distribute * 2);
        for (uint32_t i = 0; i < table_len; i++)
            if (table[i] == UNASSIGNED && (uint64_t)counts[i] <= low) { table[i] = 1; to_distribute -= 1; total -= counts[i]; }
    }
    if (((1ull << log2) - to_distribute) == (uint64_t)table_len) {
        uint64_t v_max = 0; uint32_t i_max = 0;
        for (uint32_t i = 0; i < 256; i++) {
            uint64_t c = (i < table_len) ? (uint64_t)counts[i] : 0;  // counts beyond table_len are zero
            if (c > v_max) { v_max = c; i_max = i; }
        }
        table[i_max] += (int32_t)to_distribute;
        return 1;
    } else if (total == 0) {
        while (to_distribute != 0) {
            bool progressed = false;
            for (uint32_t i = 0; i < table_len; i++) {
                if (table[i] > 0) {
                    table[i] += 1; to_distribute -= 1; progressed = true;
                    if (to_distribute == 0) break;
                }
            }
            if (!progressed) return ST_PANIC;
        }
    } else {
        uint64_t v_step_log = 62 - (uint64_t)log2;
        uint64_t mid = (1ull << (v_step_log - 1)) - 1;
        uint64_t r_step = (((1ull << v_step_log) * to_distribute) + mid) / total;
        uint64_t tmp_total = mid;
        for (uint32_t i = 0; i < table_len; i++) {
            if (table[i] == UNASSIGNED) {
                uint64_t end = tmp_total + (uint64_t)counts[i] * r_step;
                uint64_t weight = (end >> v_step_log) - (tmp_total >> v_step_log);
                if (weight < 1) return ST_PANIC;
                table[i] = (int32_t)weight;
                tmp_total = end;
            }
        }
    }
    return 1;
}

// ------------------------------------------------------------------------------------------
// Histogram::normalize, src/histogram.rs:95-155, one warp, lane L owns symbols 8L..8L+7.
// counts: 256 entries (shared or global), norm: int32[256] in shared memory (distinct from counts).
// req_log2 == 0 -> optimal_log2.  Returns 0, 1 (slow path taken) or ST_PANIC; log2 / table_len out.
// ------------------------------------------------------------------------------------------
template <typename CT>
__device__ int warp_normalize(const CT *counts, uint64_t size, uint32_t req_log2, int32_t *norm, int lane,
                              uint32_t &log2_out, uint32_t &table_len_out)
{
    uint64_t c[8];
    int hi_local = -1;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        c[k] = counts[lane * 8 + k];
        if (c[k] != 0) hi_local = lane * 8 + k;
    }
    int hi = hi_local;
#pragma unroll
    for (int d = 16; d; d >>= 1) hi = max(hi, __shfl_xor_sync(FULL, hi, d));
    uint32_t table_len = (uint32_t)(hi < 0 ? 0 : hi) + 1;  // histogram.rs:52-59
    table_len_out = table_len;
    if (table_len <= 1 || size == 0) return ST_PANIC;       // ilog2(0): histogram.rs:98,267
    if (req_log2 == 0) {
        if (size <= 4) return ST_PANIC;                     // histogram.rs:271 underflow
        req_log2 = optimal_log2(size, table_len);
    }
    uint32_t log2 = min(max(req_log2, (uint32_t)TL_MIN), (uint32_t)TL_MAX);
    log2 = max(log2, ilog2u(table_len - 1) + 2);           // histogram.rs:96-98
    log2_out = log2;

    uint64_t scale = 62 - (uint64_t)log2;
    uint64_t step = (1ull << 62) / size;
    uint64_t v_step = 1ull << (scale - 20);
    uint64_t low_threshold = size >> log2;
    int32_t pr[8];
    int32_t sum = 0;
    uint32_t best = 255;  // (prob << 8) | (255 - index): first strict maximum wins (histogram.rs:135-138)
    int single = -1;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        uint64_t t = c[k];
        int32_t p = 0;
        if (t == size) single = lane * 8 + k;               // histogram.rs:113-120
        if (t != 0) {
            if (t <= low_threshold) { p = -1; sum += 1; }
            else {
                uint64_t prob = (t * step) >> scale;
                if (prob < 8) {
                    uint64_t rest_to_beat = v_step * (uint64_t)RTB_TABLE[prob];
                    prob += ((t * step - (prob << scale)) > rest_to_beat) ? 1 : 0;
                }
                p = (int32_t)prob;
                sum += p;
                uint32_t key = ((uint32_t)p << 8) | (uint32_t)(255 - (lane * 8 + k));
                best = max(best, key);
            }
        }
        pr[k] = p;
    }
    uint32_t any_single = __ballot_sync(FULL, single >= 0);
    if (any_single) {
#pragma unroll
        for (int k = 0; k < 8; k++) norm[lane * 8 + k] = (single == lane * 8 + k) ? (int32_t)(1u << log2) : 0;
        __syncwarp();
        return 0;
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        sum += __shfl_xor_sync(FULL, sum, d);
        best = max(best, __shfl_xor_sync(FULL, best, d));
    }
    int32_t to_distribute = (int32_t)(1u << log2) - sum;
    int32_t largest_prob = (int32_t)(best >> 8);
    uint32_t largest = 255 - (best & 255);
    if (to_distribute != 0 && -to_distribute >= (largest_prob >> 1)) {  // histogram.rs:144-145
        int rc = 0;
        if (lane == 0) rc = normalize_slow_serial(counts, size, table_len, log2, norm);
        rc = __shfl_sync(FULL, rc, 0);
        __syncwarp();
        return rc;
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
        int32_t p = pr[k];
        if ((uint32_t)(lane * 8 + k) == largest) p += to_distribute;     // histogram.rs:147
        norm[lane * 8 + k] = p;
    }
    __syncwarp();
    return 0;
}

// ------------------------------------------------------------------------------------------
// NormHistogram::write, src/histogram.rs:376-431, one warp.  Field widths depend only on the
// running sum of |count| (threshold = 1 << ilog2(remaining), num_bits = ilog2(remaining) + 1 after
// the adjust loop :424-427), zero-run codes on the index of the previous non-zero symbol, so each
// lane serialises its 8 symbols into a private bit string and warp_place concatenates them.
// dst: word aligned, >= HDR_RESERVE bytes.  Returns header bits; bytes = ceil(bits / 8).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t warp_ncount_write(const int32_t *norm, uint32_t log2, uint32_t table_len,
                                                      uint32_t *rows, uint32_t *dst, int lane)
{
    int32_t x[8];
    uint32_t asum = 0;
    int last_nz = -1;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        int i = lane * 8 + k;
        x[k] = (i < (int)table_len) ? norm[i] : 0;
        asum += (uint32_t)abs(x[k]);
        if (x[k] != 0) last_nz = i;
    }
    uint32_t before = warp_incl_add(asum, lane) - asum;  // sum of |count| over symbols before mine
    int pm = warp_incl_max(last_nz, lane);
    int prev_nz = __shfl_up_sync(FULL, pm, 1);           // last non-zero symbol index before my 8
    if (lane == 0) prev_nz = -1;

    BitRow br;
    br.init(rows + lane * ROW_STRIDE, 0, 0);
    if (lane == 0) br.put(log2 - TL_MIN, 4);             // :380-381
    int32_t remaining = (int32_t)((1u << log2) + 1u - before);
#pragma unroll
    for (int k = 0; k < 8; k++) {
        int i = lane * 8 + k;
        if (i < (int)table_len) {
            int32_t s = x[k];
            bool prev_zero = (i > 0) && (prev_nz != i - 1);
            if (s != 0 || !prev_zero) {                  // zeros inside a run emit nothing (:392-395)
                if (s != 0 && prev_zero) {               // run of zeros ended: repeat codes (:399-408)
                    uint32_t z = (uint32_t)(i - 1 - prev_nz) - 1;
                    while (z >= 24) { br.put(0xFFFF, 16); z -= 24; }
                    while (z >= 3) { br.put(3, 2); z -= 3; }
                    br.put(z, 2);
                }
                int32_t threshold = 1 << ilog2u((uint32_t)remaining);
                uint32_t num_bits = ilog2u((uint32_t)remaining) + 1;
                int32_t mx = (2 * threshold - 1) - remaining;   // :410
                int32_t count = s + 1;
                if (count >= threshold) count += mx;
                uint32_t nb = num_bits - (count < mx ? 1u : 0u);
                br.put((uint32_t)count & ((1u << nb) - 1u), nb);
            }
            remaining -= abs(s);
            if (s != 0) prev_nz = i;
        }
    }
    uint32_t tot = br.finish();
    __syncwarp();
    uint32_t cw = 0, cb = 0;
    bool ovf = false;
    uint32_t nwords = warp_place(rows + lane * ROW_STRIDE, tot, dst, HDR_RESERVE / 4, lane, cw, cb, ovf);
    if (lane == 0 && cb) dst[nwords] = cw;
    __syncwarp();
    return nwords * 32 + cb;
}

// ------------------------------------------------------------------------------------------
// Forward LSB-first bit reader over global bytes with a total_bits bound
// (BitStreamReader, src/bitstream/stream_reader.rs:56-114), used by one lane for the header.
// ------------------------------------------------------------------------------------------
struct FwdBits {
    const uint8_t *p;
    uint32_t nbytes, pos;  // pos in bits
    __device__ __forceinline__ bool peek(uint32_t n, uint32_t &out) const
    {
        if (pos + n > nbytes * 8) return false;  // UnexpectedEof, :85-87
        uint32_t byte = pos >> 3;
        uint32_t v = 0;
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (byte + k < nbytes) v |= (uint32_t)p[byte + k] << (8 * k);
        out = (v >> (pos & 7)) & ((1u << n) - 1u);  // n <= 16
        return true;
    }
};

// The same reader over a word-aligned shared-memory copy of the header (warp_stage_header): one funnel shift
// per peek instead of four byte loads from global memory.  w[] must be readable one word past the data.
struct SmemBits {
    const uint32_t *w;
    uint32_t nbytes, pos, bias;  // pos in bits from the first header byte; bias = bit offset of that byte in w[0]
    __device__ __forceinline__ bool peek(uint32_t n, uint32_t &out) const
    {
        if (pos + n > nbytes * 8) return false;
        const uint32_t q = pos + bias, i = q >> 5;
        out = __funnelshift_r(w[i], w[i + 1], q & 31) & ((1u << n) - 1u);
        return true;
    }
};

constexpr uint32_t HDR_STAGE_BYTES = 508;   // NormHistogram::write_bound is at most 483 bytes (histogram.rs:330-337)
constexpr uint32_t HDR_STAGE_WORDS = 129;

// Copies the first min(nbytes, HDR_STAGE_BYTES) bytes at src (any alignment) into hw[0 .. HDR_STAGE_WORDS) with
// aligned 32-bit loads; returns the number of bytes staged.  Reads stay inside the 4-byte words that hold the data.
__device__ __forceinline__ uint32_t warp_stage_header(const uint8_t *src, uint32_t nbytes, uint32_t *hw, int lane, uint32_t &bias_bits)
{
    const uint32_t bias = (uint32_t)((uintptr_t)src & 3);
    const uint32_t *origin = reinterpret_cast<const uint32_t *>(src - bias);
    const uint32_t staged = min(nbytes, HDR_STAGE_BYTES);
    const uint32_t nwords = (bias + staged + 3) >> 2;          // <= 128
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t w = lane + 32 * k;
        hw[w] = w < nwords ? __ldg(origin + w) : 0u;
    }
    if (lane == 0) hw[128] = 0;
    bias_bits = bias * 8;
    __syncwarp();
    return staged;
}

// NormHistogram::read, src/histogram.rs:436-505.  Serial on the calling lane.  norm must be zeroed.
template <typename Reader>
__device__ __forceinline__ int ncount_read_impl(Reader r, uint32_t nbytes, int32_t *norm, uint32_t &log2_out,
                                                uint32_t &table_len_out, uint32_t &consumed_out)
{
    if (nbytes == 0) return ST_PANIC;  // stream_reader.rs:17
    uint32_t v;
    if (!r.peek(4, v)) return ST_IO;
    r.pos += 4;
    uint32_t log2 = v + TL_MIN;
    if (log2 > TL_MAX) return ST_TABLE_LOG;
    log2_out = log2;
    uint32_t symbol = 0;
    uint32_t threshold = 1u << log2;
    uint32_t remaining = threshold + 1;
    uint32_t rbc = log2 + 1;
    bool previous0 = false;
    while (remaining > 1 && symbol < 256) {
        if (previous0) {
            for (;;) {
                uint32_t pk = 0;
                if (!r.peek(16, pk)) pk = 0;
                if (pk != 0xFFFF) break;
                r.pos += 16;
                symbol += 24;
            }
            for (;;) {
                uint32_t pk = 0;
                if (!r.peek(2, pk)) pk = 0;
                if (pk != 3) break;
                r.pos += 2;
                symbol += 3;
            }
            if (!r.peek(2, v)) return ST_IO;
            r.pos += 2;
            symbol += v;
        }
        if (symbol >= 256) break;
        uint32_t mx = (2 * threshold - 1) - remaining;
        uint32_t raw;
        if (!r.peek(rbc, raw)) {
            if (!r.peek(rbc - 1, raw)) return ST_IO;
        }
        uint32_t value;
        if ((raw & (threshold - 1)) < mx) {
            if (r.pos + rbc - 1 > nbytes * 8) return ST_IO;
            r.pos += rbc - 1;
            value = raw & (threshold - 1);
        } else {
            if (r.pos + rbc > nbytes * 8) return ST_IO;
            r.pos += rbc;
            value = raw & (2 * threshold - 1);
            if (value >= threshold) value -= mx;
        }
        int32_t val = (int32_t)value - 1;
        remaining -= (uint32_t)abs(val);
        norm[symbol] = val;
        symbol += 1;
        previous0 = (val == 0);
        while (remaining < threshold) { rbc -= 1; threshold >>= 1; }
    }
    if (remaining != 1) return ST_TOO_MANY;
    table_len_out = symbol;
    consumed_out = (r.pos + 7) >> 3;
    return 0;
}
__device__ int ncount_read_serial(const uint8_t *src, uint32_t nbytes, int32_t *norm, uint32_t &log2_out,
                                  uint32_t &table_len_out, uint32_t &consumed_out)
{
    return ncount_read_impl(FwdBits{src, nbytes, 0}, nbytes, norm, log2_out, table_len_out, consumed_out);
}
// Header parse for the decode kernels: the warp stages the header in shared memory (hw: HDR_STAGE_WORDS words),
// lane 0 parses it there.  A stream that runs past the staged bytes (no valid header does) is re-parsed from
// global memory so that the error codes stay those of the reference reader.  norm must be zeroed; all lanes call.
__device__ __forceinline__ int warp_ncount_read(const uint8_t *src, uint32_t nbytes, uint32_t *hw, int32_t *norm, int lane,
                                                uint32_t &log2_out, uint32_t &table_len_out, uint32_t &consumed_out)
{
    uint32_t bias_bits;
    const uint32_t staged = warp_stage_header(src, nbytes, hw, lane, bias_bits);
    int rc = 0;
    uint32_t log2 = 0, table_len = 0, consumed = 0;
    if (lane == 0) rc = ncount_read_impl(SmemBits{hw, staged, 0, bias_bits}, staged, norm, log2, table_len, consumed);
    rc = __shfl_sync(FULL, rc, 0);
    if (rc == ST_IO && staged < nbytes) {
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 8; k++) norm[k * 32 + lane] = 0;
        __syncwarp();
        if (lane == 0) rc = ncount_read_serial(src, nbytes, norm, log2, table_len, consumed);
        rc = __shfl_sync(FULL, rc, 0);
    }
    log2_out = __shfl_sync(FULL, log2, 0);
    table_len_out = __shfl_sync(FULL, table_len, 0);
    consumed_out = __shfl_sync(FULL, consumed, 0);
    __syncwarp();
    return rc;
}

// ------------------------------------------------------------------------------------------
// Symbol spread shared by EncodeTable::update (src/fse.rs:119-151) and DecodeTable::update
// (src/fse.rs:294-324), one warp.  Closed form: the k-th accepted cell of the walk
// position = (position + step) & mask is (j*step)&mask for the k-th j whose cell is
// <= high_threshold; it receives the k-th entry of the symbols expanded by positive counts;
// low-probability symbols (-1) take the cells size-1, size-2, ... in symbol order.
// The fill runs over ranks, 32 per step: the first rank of every present symbol is flagged in bit 15 of
// the rank -> cell map, so a ballot and a running popcount give each rank the index of its symbol in
// the list of present symbols (independent of how many symbols there are or how small they are).
//   spread  : uint8[size] out (cell -> symbol)
//   cum     : uint32[256] out: exclusive prefix of |norm| (EncodeTable's cumul / symbol_tt total);
//             its storage holds the list of present symbols until the end of the call
//   posmap  : uint16[size] scratch (rank -> cell, bit 15 = a symbol starts here)
// norm is read into registers before anything is written, so `cum` may be norm's own storage.  CTR_OUT (decoders): `cum`
// receives the running counters DecodeTable::update starts from (fse.rs:327-328: the count, 1 for -1) instead of the
// prefix, which lets norm, the symbol list and the counters share one 1 KiB array.
// ------------------------------------------------------------------------------------------
template <bool CTR_OUT = false>
__device__ __forceinline__ void warp_spread(const int32_t *norm, uint32_t log2, uint32_t table_len, uint8_t *spread,
                                            uint32_t *cum, uint16_t *posmap, int lane)
{
    const uint32_t size = 1u << log2, mask = size - 1, step = table_step(size);
    uint8_t *psym = reinterpret_cast<uint8_t *>(cum);
    int32_t x[8];
    uint32_t apre[8], ppre[8];
    uint32_t asum = 0, psum = 0, lsum = 0, nsum = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        int i = lane * 8 + k;
        x[k] = (i < (int)table_len) ? norm[i] : 0;
        apre[k] = asum; ppre[k] = psum;
        asum += (uint32_t)abs(x[k]);
        psum += (uint32_t)max(x[k], 0);
        lsum += (x[k] < 0) ? 1u : 0u;
        nsum += (x[k] > 0) ? 1u : 0u;
    }
    uint32_t packed = warp_incl_add((asum << 16) | psum, lane);  // sums <= 2^15 each
    uint32_t abase = (packed >> 16) - asum, pbase = (packed & 0xffff) - psum;
    uint32_t packed2 = warp_incl_add((lsum << 16) | nsum, lane); // <= 256 each
    uint32_t lbase = (packed2 >> 16) - lsum, nbase = (packed2 & 0xffff) - nsum;
    const uint32_t L = __shfl_sync(FULL, packed2, 31) >> 16;
    const uint32_t high_threshold = size - 1 - L, P = size - L;   // P cells go to positive counts
    uint32_t lrank = lbase;
#pragma unroll
    for (int k = 0; k < 8; k++)
        if (x[k] < 0) { spread[size - 1 - lrank] = (uint8_t)(lane * 8 + k); lrank++; }  // fse.rs:122-125
    if (L) {  // rank -> cell map of the accepted positions
        uint32_t base = 0;
        for (uint32_t j0 = 0; j0 < size; j0 += 32) {
            uint32_t pos = ((j0 + lane) * step) & mask;
            bool acc = pos <= high_threshold;
            uint32_t b = __ballot_sync(FULL, acc);
            if (acc) posmap[base + __popc(b & lt_mask(lane))] = (uint16_t)pos;
            base += __popc(b);
        }
    } else {
        for (uint32_t j = lane; j < size; j += 32) posmap[j] = (uint16_t)((j * step) & mask);
    }
    __syncwarp();
    uint32_t nrank = nbase;
#pragma unroll
    for (int k = 0; k < 8; k++)
        if (x[k] > 0) {
            psym[nrank++] = (uint8_t)(lane * 8 + k);
            posmap[pbase + ppre[k]] |= 0x8000u;
        }
    __syncwarp();
    uint32_t run = 0;
    for (uint32_t j0 = 0; j0 < P; j0 += 32) {
        const uint32_t r = j0 + lane;
        const uint32_t v = r < P ? (uint32_t)posmap[r] : 0u;
        const uint32_t b = __ballot_sync(FULL, (v & 0x8000u) != 0);
        const uint32_t kk = run + __popc(b & (lt_mask(lane) | (1u << lane)));   // symbols started up to this rank
        run += __popc(b);
        if (r < P) spread[v & 0x7fffu] = psym[kk - 1];
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 8; k++) cum[lane * 8 + k] = CTR_OUT ? (x[k] < 0 ? 1u : (uint32_t)x[k]) : abase + apre[k];
    __syncwarp();
}

// EncodeTable::update, src/fse.rs:153-188: next-state table and symbol transforms.
//   table[cumul[s] + r] = size + cell, r = rank of the cell among the cells of s in increasing cell order.
//   cum is consumed as the running counter (fse.rs:160-161).
__device__ __forceinline__ void warp_build_encode(const int32_t *norm, uint32_t log2, uint32_t table_len,
                                                  const uint8_t *spread, uint32_t *cum, uint16_t *table, uint2 *tt,
                                                  int lane)
{
    const uint32_t size = 1u << log2;
#pragma unroll
    for (int k = 0; k < 8; k++) {  // fse.rs:165-188, before cum is consumed
        int i = lane * 8 + k;
        uint2 t = make_uint2(0u, 0u);
        if (i < (int)table_len) {
            int32_t x = norm[i];
            int32_t total = (int32_t)cum[i];
            if (x == 0) t.x = ((log2 + 1) << 16) - size;
            else if (x == -1 || x == 1) { t.x = (log2 << 16) - size; t.y = (uint32_t)(total - 1); }
            else {
                uint32_t mbo = log2 - ilog2u((uint32_t)(x - 1));
                t.x = (mbo << 16) - ((uint32_t)x << mbo);
                t.y = (uint32_t)(total - x);
            }
        }
        tt[i] = t;
    }
    __syncwarp();
    for (uint32_t c0 = 0; c0 < size; c0 += 32) {
        uint32_t cell = c0 + lane;
        uint32_t s = spread[cell];
        uint32_t m = __match_any_sync(FULL, s);
        uint32_t r = __popc(m & lt_mask(lane));
        uint32_t base = cum[s];
        __syncwarp();
        if (r == 0) cum[s] = base + __popc(m);
        table[base + r] = (uint16_t)(size + cell);
        __syncwarp();
    }
}

// DecodeTable::update, src/fse.rs:294-337: entry = new_state | symbol << 16 | num_bits << 24
// (the little-endian image of DecodeTransform {u16 new_state, u8 symbol, u8 num_bits}, fse.rs:260-265).
//   ctr: uint32[256] scratch (symbol_next, fse.rs:295-310; u32 so the dead wrap of Q5 cannot happen)
template <bool INIT = true>     // false: ctr already holds the counters (warp_spread<true>)
__device__ __forceinline__ void warp_build_decode(const int32_t *norm, uint32_t log2, uint32_t table_len,
                                                  const uint8_t *spread, uint32_t *ctr, uint32_t *table, int lane)
{
    const uint32_t size = 1u << log2;
    if (INIT) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            int i = lane * 8 + k;
            int32_t x = (i < (int)table_len) ? norm[i] : 0;
            ctr[i] = (x < 0) ? 1u : (uint32_t)x;
        }
        __syncwarp();
    }
    for (uint32_t c0 = 0; c0 < size; c0 += 32) {
        uint32_t cell = c0 + lane;
        uint32_t s = spread[cell];
        uint32_t m = __match_any_sync(FULL, s);
        uint32_t r = __popc(m & lt_mask(lane));
        uint32_t base = ctr[s];
        __syncwarp();
        if (r == 0) ctr[s] = base + __popc(m);
        uint32_t next = base + r;
        uint32_t nb = log2 - ilog2u(next | (next == 0));  // next == 0 only for malformed tables
        uint32_t ns = ((next << nb) - size) & 0xffffu;
        table[cell] = ns | (s << 16) | (nb << 24);
        __syncwarp();
    }
}

// splitmix64-indexed generators of SURVEY.md 8(d)
__device__ __forceinline__ uint64_t splitmix64(uint64_t x)
{
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

}  // namespace fsed
