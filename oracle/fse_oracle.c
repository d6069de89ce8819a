/*
 * fse_oracle.c -- CPU oracle for the FSE (tANS) hot path of Cognoscan/entropy_coders.
 *
 * TEST INFRASTRUCTURE ONLY (see fse_oracle.h).  Plain C restatement of the
 * reference crate's arithmetic; every function cites the reference file:line it
 * follows (paths relative to /root/reference).  PARITY UNPINNED against a
 * reference *execution*: the crate is Rust, no Rust toolchain exists in this
 * image, and the crate's own tests hold no encoded-byte golden vectors.  It is
 * pinned against the crate's known-answer/property tests, the hand-derivable
 * KATs in tests/golden/, the independent mechanics model oracle/pymodel.py, and --
 * for the NCount header, the spread, both tables and the two-state stream format
 * of fse_compress2 / fse_decompress2 -- a real libzstd (1.5.5, the FSE the crate
 * ports) in both directions: it decodes the FSE streams libzstd writes and libzstd
 * decodes the streams it writes, and its tables of up to 512 entries decode the
 * sequence streams of ordinary zstd frames (tests/test_zstd_interop.py).
 */
#include "fse_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ helpers */

static inline uint32_t ilog2_u64(uint64_t v) { return 63u - (uint32_t)__builtin_clzll(v); }
static inline uint32_t ilog2_u32(uint32_t v) { return 31u - (uint32_t)__builtin_clz(v); }

/* ------------------------------------------------------------ histogram.rs */

/* histogram.rs:18-66.  The four privatised tables of :20-50 sum to plain counts. */
void fse_or_histogram(const uint8_t *data, size_t n, fse_or_hist *out)
{
    memset(out, 0, sizeof(*out));
    uint32_t t[4][256];
    memset(t, 0, sizeof(t));
    size_t i = 0;
    for (; i + 4 <= n; i += 4) {
        t[0][data[i]]++;
        t[1][data[i + 1]]++;
        t[2][data[i + 2]]++;
        t[3][data[i + 3]]++;
    }
    for (size_t j = 0; i + j < n; j++) t[j][data[i + j]]++;
    for (int s = 0; s < 256; s++)
        out->table[s] = (uint64_t)t[0][s] + t[1][s] + t[2][s] + t[3][s];
    uint32_t table_len = 0; /* :52-59 */
    for (int s = 255; s >= 0; s--) {
        if (out->table[s] != 0) { table_len = (uint32_t)s; break; }
    }
    out->table_len = table_len + 1;
    out->size = n;
}

/* histogram.rs:79-81 / :321-323 -- counts the ZERO entries (quirk Q3). */
uint32_t fse_or_symbol_count(const int32_t *table256)
{
    uint32_t c = 0;
    for (int i = 0; i < 256; i++) c += (table256[i] == 0);
    return c;
}

/* histogram.rs:264-277 */
int fse_or_optimal_log2(const fse_or_hist *h, uint32_t *log2_out)
{
    /* ilog2(0) panics (:266,:267); ilog2(size-1)-2 underflows u32 for size<=4 (:271,
     * debug panic).  Report both as the reference panicking. */
    if (h->size == 0 || h->table_len <= 1 || h->size <= 4) return FSE_OR_ERR_PANIC;
    uint32_t min_bits_src = ilog2_u64(h->size) + 1;
    uint32_t min_bits_symbols = ilog2_u32(h->table_len - 1) + 2;
    uint32_t min_bits = min_bits_src < min_bits_symbols ? min_bits_src : min_bits_symbols;
    uint32_t max_bits = ilog2_u64(h->size - 1) - 2;
    uint32_t v = FSE_OR_TABLE_LOG_DEFAULT;
    if (max_bits < v) v = max_bits;
    if (min_bits > v) v = min_bits;
    if (v < FSE_OR_TABLE_LOG_MIN) v = FSE_OR_TABLE_LOG_MIN;
    if (v > FSE_OR_TABLE_LOG_MAX) v = FSE_OR_TABLE_LOG_MAX;
    *log2_out = v;
    return FSE_OR_OK;
}

/* histogram.rs:157-261 */
static int normalize_slow(const fse_or_hist *h, uint32_t log2, fse_or_norm *out)
{
    const int32_t UNASSIGNED = -2;
    const uint32_t tl = h->table_len;
    uint64_t low_threshold = h->size >> log2;
    uint64_t low_one = (h->size * 3) >> (log2 + 1);
    uint64_t to_distribute = 1ull << log2;
    uint64_t total = h->size;
    int32_t *table = out->table;
    memset(table, 0, sizeof(out->table));
    out->log2 = log2;
    out->table_len = tl;

    for (uint32_t i = 0; i < tl; i++) { /* :167-181 */
        uint64_t t = h->table[i];
        if (t == 0) continue;
        if (t <= low_threshold) { table[i] = -1; to_distribute -= 1; total -= t; }
        else if (t <= low_one)  { table[i] = 1;  to_distribute -= 1; total -= t; }
        else table[i] = UNASSIGNED;
    }
    if (to_distribute == 0) return 1; /* :183-189 */

    if ((total / to_distribute) > low_one) { /* :192-201 */
        uint64_t low = (total * 3) / (to_distribute * 2);
        for (uint32_t i = 0; i < tl; i++) {
            if (table[i] == UNASSIGNED && h->table[i] <= low) {
                table[i] = 1; to_distribute -= 1; total -= h->table[i];
            }
        }
    }

    if (((1ull << log2) - to_distribute) == (uint64_t)tl) { /* :203-220 */
        uint64_t v_max = 0; int i_max = 0;
        for (int i = 0; i < 256; i++)
            if (h->table[i] > v_max) { v_max = h->table[i]; i_max = i; }
        table[i_max] += (int32_t)to_distribute;
        return 1;
    } else if (total == 0) { /* :221-235 */
        while (to_distribute != 0) {
            int progressed = 0;
            for (uint32_t i = 0; i < tl; i++) {
                if (table[i] > 0) {
                    table[i] += 1; to_distribute -= 1; progressed = 1;
                    if (to_distribute == 0) break;
                }
            }
            if (!progressed) return FSE_OR_ERR_PANIC; /* the reference would spin forever */
        }
    } else { /* :236-254 */
        uint64_t v_step_log = 62 - (uint64_t)log2;
        uint64_t mid = (1ull << (v_step_log - 1)) - 1;
        uint64_t r_step = (((1ull << v_step_log) * to_distribute) + mid) / total;
        uint64_t tmp_total = mid;
        for (uint32_t i = 0; i < tl; i++) {
            if (table[i] == UNASSIGNED) {
                uint64_t end = tmp_total + h->table[i] * r_step;
                uint64_t weight = (end >> v_step_log) - (tmp_total >> v_step_log);
                if (weight < 1) return FSE_OR_ERR_PANIC; /* :247-249 */
                table[i] = (int32_t)weight;
                tmp_total = end;
            }
        }
    }
    return 1;
}

/* histogram.rs:95-155 */
int fse_or_normalize(const fse_or_hist *h, uint32_t log2, fse_or_norm *out)
{
    static const uint32_t RTB[8] = {0, 473195, 504333, 520860, 550000, 700000, 750000, 830000};
    if (h->table_len <= 1 || h->size == 0) return FSE_OR_ERR_PANIC; /* ilog2(0) :98, div by 0 :103 */
    if (log2 < FSE_OR_TABLE_LOG_MIN) log2 = FSE_OR_TABLE_LOG_MIN;
    if (log2 > FSE_OR_TABLE_LOG_MAX) log2 = FSE_OR_TABLE_LOG_MAX;
    uint32_t need = ilog2_u32(h->table_len - 1) + 2;
    if (need > log2) log2 = need;

    uint64_t scale = 62 - (uint64_t)log2;
    uint64_t step = (1ull << 62) / h->size;
    uint64_t v_step = 1ull << (scale - 20);
    uint64_t low_threshold = h->size >> log2;
    int32_t to_distribute = 1 << log2;
    uint32_t largest = 0;
    int32_t largest_prob = 0;

    memset(out->table, 0, sizeof(out->table));
    out->log2 = log2;
    out->table_len = h->table_len;

    for (uint32_t i = 0; i < h->table_len; i++) {
        uint64_t t = h->table[i];
        if (t == h->size) { out->table[i] = to_distribute; return FSE_OR_OK; } /* :113-120 */
        if (t == 0) continue;
        if (t <= low_threshold) { out->table[i] = -1; to_distribute -= 1; continue; }
        uint64_t prob = (t * step) >> scale;
        if (prob < 8) {
            uint64_t rest_to_beat = v_step * (uint64_t)RTB[prob];
            prob += ((t * step - (prob << scale)) > rest_to_beat);
        }
        int32_t p = (int32_t)prob;
        if (p > largest_prob) { largest_prob = p; largest = i; }
        out->table[i] = p;
        to_distribute -= p;
    }
    if (to_distribute != 0 && -to_distribute >= (largest_prob >> 1)) /* :144-145 */
        return normalize_slow(h, log2, out);
    out->table[largest] += to_distribute; /* :147 */
    return FSE_OR_OK;
}

/* histogram.rs:299-303 */
int fse_or_norm_new(const uint8_t *data, size_t n, fse_or_norm *out)
{
    fse_or_hist h;
    uint32_t log2;
    fse_or_histogram(data, n, &h);
    int rc = fse_or_optimal_log2(&h, &log2);
    if (rc < 0) return rc;
    return fse_or_normalize(&h, log2, out);
}


/* ---------------------------------------------------------------- zstd-interoperable normalisation
 * NOT the reference crate's arithmetic: this restates libzstd's published FSE_normalizeCount / FSE_normalizeM2
 * (lib/compress/fse_compress.c, zstd 1.5.x), whose NCount header format the crate shares (histogram.rs:342).  The crate's
 * normalize differs from it in three places: the `to_distribute != 0 &&` guard (histogram.rs:144), low-probability symbols
 * are always -1 (zstd: -1 only with useLowProbCount), and the table_log range (zstd: 5..12 and >= FSE_minTableLog).
 * No libzstd build in this image exports FSE_normalizeCount, but its output is visible in the frames libzstd writes: the
 * function is pinned against libzstd 1.5.5 on the weight histograms of real Huffman tree descriptions (useLowProbCount = 0)
 * and on the sequence-code histograms of real frames (useLowProbCount = 1, table_log 7 - 9), tests/test_zstd_interop.py,
 * next to the hand-derived vectors of tests/test_zstd_normalize.py.  Returns 0, 3 (one symbol holds every count: zstd's "rle special case", norm untouched = 0)
 * or a negative status. */
static int zstd_normalize_m2(int32_t *norm, uint32_t table_log, const uint64_t *count, uint64_t total, uint32_t max_symbol,
                             int32_t low_prob_count)
{
    const int32_t NOT_YET_ASSIGNED = -2;
    uint32_t s, distributed = 0, to_distribute;
    const uint64_t low_threshold = total >> table_log;
    uint64_t low_one = (total * 3) >> (table_log + 1);
    for (s = 0; s <= max_symbol; s++) {
        if (count[s] == 0) { norm[s] = 0; continue; }
        if (count[s] <= low_threshold) { norm[s] = low_prob_count; distributed++; total -= count[s]; continue; }
        if (count[s] <= low_one) { norm[s] = 1; distributed++; total -= count[s]; continue; }
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
    if (distributed == max_symbol + 1) {
        uint32_t max_v = 0;
        uint64_t max_c = 0;
        for (s = 0; s <= max_symbol; s++)
            if (count[s] > max_c) { max_v = s; max_c = count[s]; }
        norm[max_v] += (int32_t)to_distribute;
        return 0;
    }
    if (total == 0) {
        for (s = 0; to_distribute > 0; s = (s + 1) % (max_symbol + 1))
            if (norm[s] > 0) { to_distribute--; norm[s]++; }
        return 0;
    }
    {
        const uint64_t v_step_log = 62 - (uint64_t)table_log;
        const uint64_t mid = (1ull << (v_step_log - 1)) - 1;
        const uint64_t r_step = (((1ull << v_step_log) * to_distribute) + mid) / total;
        uint64_t tmp_total = mid;
        for (s = 0; s <= max_symbol; s++) {
            if (norm[s] == NOT_YET_ASSIGNED) {
                const uint64_t end = tmp_total + count[s] * r_step;
                const uint32_t weight = (uint32_t)(end >> v_step_log) - (uint32_t)(tmp_total >> v_step_log);
                if (weight < 1) return FSE_OR_ERR_PANIC;
                norm[s] = (int32_t)weight;
                tmp_total = end;
            }
        }
    }
    return 0;
}

int fse_or_normalize_zstd(const fse_or_hist *h, uint32_t table_log, int use_low_prob_count, fse_or_norm *out)
{
    static const uint32_t RTB[8] = {0, 473195, 504333, 520860, 550000, 700000, 750000, 830000};
    memset(out, 0, sizeof(*out));
    if (h->size == 0 || h->table_len == 0) return FSE_OR_ERR_PANIC;
    const uint32_t max_symbol = h->table_len - 1;
    if (table_log == 0) table_log = 11;                                          /* FSE_DEFAULT_TABLELOG */
    if (table_log < 5) return FSE_OR_ERR_PANIC;                                  /* FSE_MIN_TABLELOG: ERROR(GENERIC) */
    if (table_log > 12) return FSE_OR_ERR_TABLE_LOG;                             /* FSE_MAX_TABLELOG: ERROR(tableLog_tooLarge) */
    {
        uint32_t min_src = ilog2_u64(h->size) + 1, min_sym = (max_symbol ? ilog2_u32(max_symbol) : 0) + 2;
        if (table_log < (min_src < min_sym ? min_src : min_sym)) return FSE_OR_ERR_PANIC;   /* FSE_minTableLog */
    }
    out->log2 = table_log;
    out->table_len = h->table_len;
    const int32_t low_prob_count = use_low_prob_count ? -1 : 1;
    const uint64_t scale = 62 - (uint64_t)table_log;
    const uint64_t step = (1ull << 62) / h->size;
    const uint64_t v_step = 1ull << (scale - 20);
    int64_t still = 1ll << table_log;
    uint32_t largest = 0;
    int32_t largest_p = 0;
    const uint64_t low_threshold = h->size >> table_log;
    for (uint32_t s = 0; s <= max_symbol; s++) {
        const uint64_t c = h->table[s];
        if (c == h->size) return 3;                                              /* rle special case */
        if (c == 0) { out->table[s] = 0; continue; }
        if (c <= low_threshold) { out->table[s] = low_prob_count; still--; }
        else {
            int32_t proba = (int32_t)((c * step) >> scale);
            if (proba < 8) {
                const uint64_t rest_to_beat = v_step * RTB[proba];
                proba += (c * step) - ((uint64_t)proba << scale) > rest_to_beat;
            }
            if (proba > largest_p) { largest_p = proba; largest = s; }
            out->table[s] = proba;
            still -= proba;
        }
    }
    if (-still >= (out->table[largest] >> 1)) {
        int rc = zstd_normalize_m2(out->table, table_log, h->table, h->size, max_symbol, low_prob_count);
        if (rc < 0) return rc;
    } else out->table[largest] += (int32_t)still;
    return 0;
}

/* histogram.rs:330-337 */
size_t fse_or_write_bound(const fse_or_norm *nh)
{
    size_t m = (((size_t)nh->table_len * nh->log2) >> 3) + 3;
    return nh->table_len > 1 ? m : 512;
}

/* ---------------------------------------------------------------- bit I/O */

/* writer.rs:16-40: a writer starts byte aligned at the Vec's current length. */
void fse_or_bitw_init(fse_or_bitw *w, uint8_t *buf, size_t cap, size_t start)
{
    w->buf = buf; w->cap = cap; w->start = start; w->bitpos = 0; w->overflow = 0;
}

/* writer.rs:140-198: val is masked to `bits` (<=16) and ORed in LSB first
 * (:177-178); bytes leave little endian (:50,:116-118) => bit k of the stream is
 * bit (k%8) of byte start + k/8. */
void fse_or_bitw_put(fse_or_bitw *w, uint64_t val, unsigned bits)
{
    if (bits == 0) return;
    val &= (bits >= 64) ? ~0ull : ((1ull << bits) - 1);
    size_t pos = w->bitpos;
    size_t byte = w->start + (pos >> 3);
    unsigned sh = (unsigned)(pos & 7);
    size_t last = w->start + ((pos + bits - 1) >> 3);
    if (last >= w->cap) { w->overflow = 1; w->bitpos += bits; return; }
    /* bits <= 32 in every caller: at most 5 bytes touched */
    __uint128_t v = (__uint128_t)val << sh;
    if (sh == 0) w->buf[byte] = 0;
    for (size_t b = byte; b <= last; b++) {
        if (b != byte) w->buf[b] = 0;
        w->buf[b] |= (uint8_t)(v & 0xff);
        v >>= 8;
    }
    w->bitpos += bits;
}

/* writer.rs:201-222: returns bits written since new(); Vec len = ceil(total/8). */
size_t fse_or_bitw_finish(fse_or_bitw *w, size_t *new_len)
{
    if (new_len) *new_len = w->start + ((w->bitpos + 7) >> 3);
    return w->bitpos;
}

/* stack_reader.rs:17-92: None on empty input (:18-20) or when the last byte is 0
 * (:77-83); the highest set bit of the last byte is the marker; bits below it
 * are the stack. */
int fse_or_bitstack_init(fse_or_bitstack *r, const uint8_t *buf, size_t n)
{
    if (n == 0) return FSE_OR_ERR_NONE;
    uint8_t last = buf[n - 1];
    if (last == 0) return FSE_OR_ERR_NONE;
    r->buf = buf;
    r->bits = (n - 1) * 8 + ilog2_u32(last);
    return FSE_OR_OK;
}

static inline uint32_t get_bits(const uint8_t *buf, size_t bitpos, unsigned bits)
{
    /* value of stream bits [bitpos, bitpos+bits), bits <= 32; never reads past the
     * byte holding bit (bitpos+bits-1) */
    if (bits == 0) return 0;
    size_t first = bitpos >> 3, last = (bitpos + bits - 1) >> 3;
    uint64_t v = 0;
    for (size_t b = last + 1; b-- > first;) v = (v << 8) | buf[b];
    v >>= (bitpos & 7);
    return (uint32_t)(v & ((bits >= 32) ? 0xffffffffull : ((1ull << bits) - 1)));
}

/* stack_reader.rs:176-184: the n most recently written unread bits, original
 * significance.  None if n > bits held. */
int fse_or_bitstack_peek(const fse_or_bitstack *r, unsigned bits, uint32_t *out)
{
    if (bits > r->bits) return FSE_OR_ERR_NONE;
    *out = get_bits(r->buf, r->bits - bits, bits);
    return FSE_OR_OK;
}

/* stack_reader.rs:193-197, :211-215 */
int fse_or_bitstack_read(fse_or_bitstack *r, unsigned bits, uint32_t *out)
{
    int rc = fse_or_bitstack_peek(r, bits, out);
    if (rc < 0) return rc;
    r->bits -= bits;
    return FSE_OR_OK;
}

/* stream_reader.rs:16-49 (asserts -> PANIC) */
int fse_or_bitstream_init(fse_or_bitstream *r, const uint8_t *buf, size_t n, size_t total_bits)
{
    if (n == 0) return FSE_OR_ERR_PANIC;
    if (((total_bits + 7) / 8) != n) return FSE_OR_ERR_PANIC;
    r->buf = buf; r->len = n; r->total_bits = total_bits; r->bits_read = 0;
    return FSE_OR_OK;
}

/* stream_reader.rs:82-114.  The last0/last1 cached words of :23-41,:93-98 are the
 * zero-extended little-endian words at the same offsets, so the value is simply
 * stream bits [bits_read, bits_read+bits). */
int fse_or_bitstream_peek(const fse_or_bitstream *r, unsigned bits, uint32_t *out)
{
    if (r->bits_read + bits > r->total_bits) return FSE_OR_ERR_IO;
    *out = get_bits(r->buf, r->bits_read, bits);
    return FSE_OR_OK;
}

/* stream_reader.rs:67-75 */
int fse_or_bitstream_advance(fse_or_bitstream *r, unsigned bits)
{
    if (r->bits_read + bits > r->total_bits) return FSE_OR_ERR_IO;
    r->bits_read += bits;
    return FSE_OR_OK;
}

/* stream_reader.rs:56-60 */
int fse_or_bitstream_read(fse_or_bitstream *r, unsigned bits, uint32_t *out)
{
    int rc = fse_or_bitstream_peek(r, bits, out);
    if (rc < 0) return rc;
    return fse_or_bitstream_advance(r, bits);
}

/* ------------------------------------------------- NCount header write/read */

/* histogram.rs:376-431 */
long fse_or_ncount_write(const fse_or_norm *nh, uint8_t *dst, size_t cap, size_t *bits_out)
{
    fse_or_bitw w;
    fse_or_bitw_init(&w, dst, cap, 0);
    fse_or_bitw_put(&w, nh->log2 - FSE_OR_TABLE_LOG_MIN, 4); /* :380-381 */

    int32_t threshold = 1 << nh->log2;
    int32_t remaining = threshold + 1;
    size_t zero_count = 0;
    unsigned num_bits = nh->log2 + 1;
    for (uint32_t i = 0; i < nh->table_len; i++) {
        int32_t s = nh->table[i];
        if (remaining <= 1) break;
        if (zero_count != 0) {
            if (s == 0) { zero_count += 1; continue; }
            zero_count -= 1; /* :399-408 */
            while (zero_count >= 24) { fse_or_bitw_put(&w, 0xFFFF, 16); zero_count -= 24; }
            while (zero_count >= 3) { fse_or_bitw_put(&w, 0x3, 2); zero_count -= 3; }
            fse_or_bitw_put(&w, zero_count, 2);
        }
        int32_t max = (2 * threshold - 1) - remaining;
        remaining -= (s < 0 ? -s : s);
        int32_t count = s + 1;
        if (count >= threshold) count += max;
        unsigned bits_to_write = num_bits - (count < max);
        fse_or_bitw_put(&w, (uint64_t)(uint32_t)count, bits_to_write);
        zero_count = (count == 1);
        if (remaining < 1) return FSE_OR_ERR_PANIC; /* :419-421 */
        while (remaining < threshold) { num_bits -= 1; threshold >>= 1; }
    }
    size_t len;
    size_t bits = fse_or_bitw_finish(&w, &len);
    if (w.overflow) return FSE_OR_ERR_CAPACITY;
    if (bits_out) *bits_out = bits;
    return (long)len;
}

/* histogram.rs:436-505 */
int fse_or_ncount_read(const uint8_t *src, size_t n, fse_or_norm *out, size_t *consumed)
{
    fse_or_bitstream r;
    int rc = fse_or_bitstream_init(&r, src, n, n * 8);
    if (rc < 0) return rc;
    uint32_t v;
    if ((rc = fse_or_bitstream_read(&r, 4, &v)) < 0) return rc;
    uint32_t log2 = v + FSE_OR_TABLE_LOG_MIN;
    if (log2 > FSE_OR_TABLE_LOG_MAX) return FSE_OR_ERR_TABLE_LOG;
    memset(out->table, 0, sizeof(out->table));
    out->log2 = log2;
    out->table_len = 256;
    size_t symbol = 0;
    size_t threshold = (size_t)1 << log2;
    size_t remaining = threshold + 1;
    unsigned read_bit_count = log2 + 1;
    int previous0 = 0;

    while (remaining > 1 && symbol < 256) {
        if (previous0) { /* :455-465 */
            for (;;) {
                uint32_t p;
                if (fse_or_bitstream_peek(&r, 16, &p) < 0) p = 0;
                if (p != 0xFFFF) break;
                if ((rc = fse_or_bitstream_advance(&r, 16)) < 0) return rc;
                symbol += 24;
            }
            for (;;) {
                uint32_t p;
                if (fse_or_bitstream_peek(&r, 2, &p) < 0) p = 0;
                if (p != 3) break;
                if ((rc = fse_or_bitstream_advance(&r, 2)) < 0) return rc;
                symbol += 3;
            }
            if ((rc = fse_or_bitstream_read(&r, 2, &v)) < 0) return rc;
            symbol += v;
        }
        if (symbol >= 256) break;

        size_t max = (2 * threshold - 1) - remaining;
        uint32_t raw;
        if (fse_or_bitstream_peek(&r, read_bit_count, &raw) < 0) {
            if ((rc = fse_or_bitstream_peek(&r, read_bit_count - 1, &raw)) < 0) return rc;
        }
        size_t value;
        if ((raw & (threshold - 1)) < max) {
            if ((rc = fse_or_bitstream_advance(&r, read_bit_count - 1)) < 0) return rc;
            value = raw & (threshold - 1);
        } else {
            if ((rc = fse_or_bitstream_advance(&r, read_bit_count)) < 0) return rc;
            value = raw & (2 * threshold - 1);
            if (value >= threshold) value -= max;
        }
        int32_t val = (int32_t)value - 1;
        remaining -= (size_t)(val < 0 ? -val : val);
        out->table[symbol] = val;
        symbol += 1;
        previous0 = (val == 0);
        while (remaining < threshold) { read_bit_count -= 1; threshold >>= 1; }
    }
    if (remaining != 1) return FSE_OR_ERR_TOO_MANY; /* :498-500 */
    out->table_len = (uint32_t)symbol;
    if (consumed) *consumed = (r.bits_read + 7) / 8; /* finish_byte, stream_reader.rs:132-135 */
    return FSE_OR_OK;
}

/* histogram.rs:508-536 */
int fse_or_norm_try_from(const int32_t *table256, fse_or_norm *out)
{
    uint64_t sum = 0;
    for (int i = 0; i < 256; i++) sum += (uint64_t)(table256[i] < 0 ? -(int64_t)table256[i] : table256[i]);
    if (sum == 0) return FSE_OR_ERR_PANIC; /* ilog2(0) */
    uint32_t log2 = ilog2_u64(sum);
    if ((1ull << log2) != sum) return FSE_OR_ERR_NONE; /* Err(()) */
    uint32_t table_len = 0;
    for (int i = 255; i >= 0; i--) if (table256[i] != 0) { table_len = (uint32_t)i; break; }
    memcpy(out->table, table256, sizeof(out->table));
    out->log2 = log2;
    out->table_len = table_len + 1;
    return FSE_OR_OK;
}

/* ------------------------------------------------------------------ fse.rs */

/* fse.rs:68-70 */
size_t fse_or_table_step(size_t size) { return size * 5 / 8 + 3; }

/* fse.rs:191-193 */
size_t fse_or_compress_bound(size_t size) { return 512 + size + (size >> 7) + 4 + 8; }

/* fse.rs:101-189 */
int fse_or_enc_table_build(const fse_or_norm *nh, fse_or_enc_table *t)
{
    if (nh->log2 < FSE_OR_TABLE_LOG_MIN || nh->log2 > FSE_OR_TABLE_LOG_MAX) return FSE_OR_ERR_PANIC;
    t->table_log = nh->log2;
    size_t size = (size_t)1 << nh->log2;
    uint32_t cumul[256];
    memset(cumul, 0, sizeof(cumul));
    size_t high_threshold = size - 1;
    memset(t->symbols, 0, size);

    uint32_t acc = 0; /* :119-129 */
    for (uint32_t i = 0; i < nh->table_len; i++) {
        int32_t x = nh->table[i];
        cumul[i] = acc;
        if (x == -1) { acc += 1; t->symbols[high_threshold] = (uint8_t)i; high_threshold -= 1; }
        else acc += (uint32_t)x;
    }

    size_t position = 0, mask = size - 1, step = fse_or_table_step(size); /* :139-151 */
    for (uint32_t i = 0; i < nh->table_len; i++) {
        for (int32_t k = 0; k < nh->table[i]; k++) {
            t->symbols[position] = (uint8_t)i;
            position = (position + step) & mask;
            while (position > high_threshold) position = (position + step) & mask;
        }
    }
    if (position != 0) return FSE_OR_ERR_PANIC;

    for (size_t i = 0; i < size; i++) { /* :157-162 */
        uint8_t x = t->symbols[i];
        t->table[cumul[x]] = (uint16_t)(size + i);
        cumul[x] += 1;
    }

    memset(t->symbol_tt, 0, sizeof(t->symbol_tt)); /* :165-188 */
    int32_t total = 0;
    uint32_t tl = t->table_log;
    for (uint32_t i = 0; i < nh->table_len; i++) {
        int32_t x = nh->table[i];
        fse_or_symtt *tt = &t->symbol_tt[i];
        if (x == 0) {
            tt->bits = ((tl + 1) << 16) - (1u << tl);
        } else if (x == -1 || x == 1) {
            tt->bits = (tl << 16) - (1u << tl);
            tt->find_state = total - 1;
            total += 1;
        } else {
            uint32_t max_bits_out = tl - ilog2_u32((uint32_t)(x - 1));
            uint32_t min_state_plus = (uint32_t)x << max_bits_out;
            tt->bits = (max_bits_out << 16) - min_state_plus;
            tt->find_state = total - x;
            total += x;
        }
    }
    return FSE_OR_OK;
}

/* fse.rs:280-338 */
int fse_or_dec_table_build(const fse_or_norm *nh, fse_or_dec_table *t)
{
    if (nh->log2 < FSE_OR_TABLE_LOG_MIN || nh->log2 > FSE_OR_TABLE_LOG_MAX) return FSE_OR_ERR_PANIC;
    t->table_log = nh->log2;
    size_t size = (size_t)1 << nh->log2;
    t->fast_mode = 1;
    memset(t->table, 0, size * sizeof(t->table[0]));

    uint32_t symbol_next[256]; /* u16 in the reference; wraps only on a dead store (Q5) */
    memset(symbol_next, 0, sizeof(symbol_next));
    uint32_t large_limit = 1u << (nh->log2 - 1);
    size_t high_threshold = size - 1;
    for (uint32_t s = 0; s < nh->table_len; s++) { /* :298-310 */
        int32_t c = nh->table[s];
        if (c <= -1) {
            t->table[high_threshold].symbol = (uint8_t)s;
            high_threshold -= 1;
            symbol_next[s] = 1;
        } else {
            if ((uint32_t)c >= large_limit) t->fast_mode = 0;
            symbol_next[s] = (uint32_t)c;
        }
    }
    size_t position = 0, mask = size - 1, step = fse_or_table_step(size); /* :313-324 */
    for (uint32_t s = 0; s < nh->table_len; s++) {
        for (int32_t k = 0; k < nh->table[s]; k++) {
            t->table[position].symbol = (uint8_t)s;
            position = (position + step) & mask;
            while (position > high_threshold) position = (position + step) & mask;
        }
    }
    if (position != 0) return FSE_OR_ERR_PANIC;
    for (size_t i = 0; i < size; i++) { /* :329-337 */
        uint8_t sym = t->table[i].symbol;
        uint32_t next_state = symbol_next[sym]++;
        if (next_state == 0) return FSE_OR_ERR_PANIC; /* ilog2(0): malformed norm */
        uint32_t num_bits = nh->log2 - ilog2_u32(next_state);
        t->table[i].num_bits = (uint8_t)num_bits;
        t->table[i].new_state = (uint16_t)((next_state << num_bits) - (uint32_t)size);
    }
    return FSE_OR_OK;
}

/* fse.rs:210-218 */
static inline uint32_t enc_first(const fse_or_enc_table *t, uint8_t sym)
{
    fse_or_symtt tt = t->symbol_tt[sym];
    uint32_t bits_out = (tt.bits + (1u << 15)) >> 16;
    uint32_t value = (bits_out << 16) - tt.bits;
    size_t idx = (size_t)((int32_t)(value >> bits_out) + tt.find_state);
    return t->table[idx];
}

/* fse.rs:227-239 */
static inline uint32_t enc_step(const fse_or_enc_table *t, uint32_t value, uint8_t sym, fse_or_bitw *w)
{
    fse_or_symtt tt = t->symbol_tt[sym];
    uint32_t bits_out = (tt.bits + value) >> 16;
    fse_or_bitw_put(w, value, bits_out);
    size_t idx = (size_t)((int32_t)(value >> bits_out) + tt.find_state);
    return t->table[idx];
}

/* ------------------------------------------------------- lib.rs codecs (N) */

/* lib.rs:118-142 (N=1), :151-182 (N=2), generalised per SURVEY Appendix A.3/D:
 * state j owns indices == j (mod N); symbols are consumed in strictly decreasing
 * index order; final states are written N-1 .. 0; then the marker bit. */
long fse_or_encode_payload(const fse_or_enc_table *t, const uint8_t *src, size_t n,
                           unsigned n_states, uint8_t *dst, size_t cap, size_t *bits_out)
{
    if (n_states == 0 || n_states > 4096) return FSE_OR_ERR_PANIC;
    if (n < n_states) return FSE_OR_ERR_PANIC; /* unwrap on None: lib.rs:121,154,156 */
    uint32_t *st = (uint32_t *)malloc(sizeof(uint32_t) * n_states);
    if (!st) return FSE_OR_ERR_CAPACITY;
    fse_or_bitw w;
    fse_or_bitw_init(&w, dst, cap, 0);
    size_t i = n;
    for (unsigned k = 0; k < n_states; k++) { --i; st[i % n_states] = enc_first(t, src[i]); }
    while (i > 0) { --i; st[i % n_states] = enc_step(t, st[i % n_states], src[i], &w); }
    for (unsigned j = n_states; j-- > 0;) fse_or_bitw_put(&w, st[j], t->table_log); /* fse.rs:248-250 */
    fse_or_bitw_put(&w, 1, 1); /* lib.rs:141,181 */
    free(st);
    size_t len;
    size_t bits = fse_or_bitw_finish(&w, &len);
    if (w.overflow) return FSE_OR_ERR_CAPACITY;
    if (bits_out) *bits_out = bits;
    return (long)len;
}

long fse_or_compress_n(const uint8_t *src, size_t n, uint32_t table_log, unsigned n_states,
                       uint8_t *dst, size_t cap, size_t *header_bytes_out, size_t *payload_bits_out)
{
    fse_or_hist h;
    fse_or_norm nh;
    int rc;
    if (n == 0) return FSE_OR_ERR_PANIC;
    fse_or_histogram(src, n, &h);
    if (table_log == 0) {
        if ((rc = fse_or_optimal_log2(&h, &table_log)) < 0) return rc;
    }
    if ((rc = fse_or_normalize(&h, table_log, &nh)) < 0) return rc;
    long hb = fse_or_ncount_write(&nh, dst, cap, NULL);
    if (hb < 0) return hb;
    fse_or_enc_table *t = (fse_or_enc_table *)malloc(sizeof(*t));
    if (!t) return FSE_OR_ERR_CAPACITY;
    if ((rc = fse_or_enc_table_build(&nh, t)) < 0) { free(t); return rc; }
    long pb = fse_or_encode_payload(t, src, n, n_states, dst + hb, cap - (size_t)hb, payload_bits_out);
    free(t);
    if (pb < 0) return pb;
    if (header_bytes_out) *header_bytes_out = (size_t)hb;
    return hb + pb;
}

/* lib.rs:187-248 generalised: decode until the stack cannot supply num_bits. */
long fse_or_decode_payload_exhaust(const fse_or_dec_table *t, const uint8_t *src, size_t n,
                                   unsigned n_states, uint8_t *dst, size_t cap)
{
    fse_or_bitstack r;
    int rc = fse_or_bitstack_init(&r, src, n);
    if (rc < 0) return rc;
    if (n_states == 0 || n_states > 4096) return FSE_OR_ERR_PANIC;
    uint32_t *st = (uint32_t *)malloc(sizeof(uint32_t) * n_states);
    if (!st) return FSE_OR_ERR_CAPACITY;
    for (unsigned j = 0; j < n_states; j++) { /* fse.rs:349-352; unwrap: lib.rs:197,224-225 */
        if (fse_or_bitstack_read(&r, t->table_log, &st[j]) < 0) { free(st); return FSE_OR_ERR_PANIC; }
    }
    size_t out = 0;
    for (;;) {
        unsigned j = (unsigned)(out % n_states);
        fse_or_dec_entry e = t->table[st[j]];
        uint32_t low;
        if (fse_or_bitstack_read(&r, e.num_bits, &low) < 0) break; /* fse.rs:365 */
        if (out >= cap) { free(st); return FSE_OR_ERR_CAPACITY; }  /* Q1: would run until OOM */
        dst[out++] = e.symbol;
        st[j] = (uint32_t)e.new_state + low;
    }
    for (unsigned k = 0; k < n_states; k++) { /* lib.rs:208, :236-237, :242-243 */
        if (out >= cap) { free(st); return FSE_OR_ERR_CAPACITY; }
        unsigned j = (unsigned)(out % n_states);
        dst[out++] = t->table[st[j]].symbol;
    }
    free(st);
    return (long)out;
}

/* Length-driven decode of exactly n_out symbols (the GPU path; fixes Q1). */
int fse_or_decode_payload_len(const fse_or_dec_table *t, const uint8_t *src, size_t n,
                              unsigned n_states, uint8_t *dst, size_t n_out)
{
    fse_or_bitstack r;
    int rc = fse_or_bitstack_init(&r, src, n);
    if (rc < 0) return rc;
    if (n_states == 0 || n_states > 4096 || n_out < n_states) return FSE_OR_ERR_PANIC;
    uint32_t *st = (uint32_t *)malloc(sizeof(uint32_t) * n_states);
    if (!st) return FSE_OR_ERR_CAPACITY;
    for (unsigned j = 0; j < n_states; j++) {
        if (fse_or_bitstack_read(&r, t->table_log, &st[j]) < 0) { free(st); return FSE_OR_ERR_LENGTH; }
    }
    size_t body = n_out - n_states;
    for (size_t i = 0; i < body; i++) {
        unsigned j = (unsigned)(i % n_states);
        fse_or_dec_entry e = t->table[st[j]];
        uint32_t low;
        if (fse_or_bitstack_read(&r, e.num_bits, &low) < 0) { free(st); return FSE_OR_ERR_LENGTH; }
        dst[i] = e.symbol;
        st[j] = (uint32_t)e.new_state + low;
    }
    for (size_t i = body; i < n_out; i++) dst[i] = t->table[st[i % n_states]].symbol;
    free(st);
    return r.bits == 0 ? FSE_OR_OK : FSE_OR_ERR_LENGTH;
}

long fse_or_decompress_n_exhaust(const uint8_t *src, size_t n, unsigned n_states, uint8_t *dst, size_t cap)
{
    fse_or_norm nh;
    size_t consumed;
    int rc = fse_or_ncount_read(src, n, &nh, &consumed);
    if (rc == FSE_OR_ERR_PANIC) return rc;
    if (rc < 0) return FSE_OR_ERR_NONE; /* .ok()? lib.rs:191,219 */
    fse_or_dec_table *t = (fse_or_dec_table *)malloc(sizeof(*t));
    if (!t) return FSE_OR_ERR_CAPACITY;
    if ((rc = fse_or_dec_table_build(&nh, t)) < 0) { free(t); return rc; }
    long out = fse_or_decode_payload_exhaust(t, src + consumed, n - consumed, n_states, dst, cap);
    free(t);
    return out;
}

int fse_or_decompress_n_len(const uint8_t *src, size_t n, unsigned n_states, uint8_t *dst, size_t n_out)
{
    fse_or_norm nh;
    size_t consumed;
    int rc = fse_or_ncount_read(src, n, &nh, &consumed);
    if (rc < 0) return rc;
    fse_or_dec_table *t = (fse_or_dec_table *)malloc(sizeof(*t));
    if (!t) return FSE_OR_ERR_CAPACITY;
    if ((rc = fse_or_dec_table_build(&nh, t)) < 0) { free(t); return rc; }
    rc = fse_or_decode_payload_len(t, src + consumed, n - consumed, n_states, dst, n_out);
    free(t);
    return rc;
}

/* ------------------------------------------------------------------------
 * The reference's own loop structure, for CPU timing: lib.rs:146-183 and
 * lib.rs:215-248 with a 64-bit accumulator that is flushed in whole 32-bit
 * halves once per symbol pair (writer.rs:43-110, 64-bit host) and a stack reader
 * refilled by 32 bits whenever <= 32 bits are held (stack_reader.rs:97-172).
 * Byte-identical to the N=2 semantic path above (tested).
 * ---------------------------------------------------------------------- */

typedef struct { uint8_t *p; uint64_t acc; unsigned bits; } fastw;

static inline void fastw_put(fastw *w, uint32_t val, unsigned bits) /* writer.rs:140-149 */
{
    w->acc |= (uint64_t)(val & ((1u << bits) - 1)) << w->bits;
    w->bits += bits;
}
static inline void fastw_flush(fastw *w) /* writer.rs:86-91 (aligned fast path) */
{
    uint32_t lo = (uint32_t)w->acc;
    memcpy(w->p, &lo, 4);
    unsigned inc = (w->bits >> 5) != 0;
    w->acc >>= inc * 32;
    w->bits -= inc * 32;
    w->p += inc * 4;
}

long fse_or_ref_compress2(const uint8_t *src, size_t n, uint8_t *dst, size_t cap)
{
    fse_or_norm nh;
    int rc;
    if (n < 2) return FSE_OR_ERR_PANIC;
    if (cap < fse_or_compress_bound(n)) return FSE_OR_ERR_CAPACITY;
    if ((rc = fse_or_norm_new(src, n, &nh)) < 0) return rc;
    long hb = fse_or_ncount_write(&nh, dst, cap, NULL);
    if (hb < 0) return hb;
    fse_or_enc_table *t = (fse_or_enc_table *)malloc(sizeof(*t));
    if (!t) return FSE_OR_ERR_CAPACITY;
    if ((rc = fse_or_enc_table_build(&nh, t)) < 0) { free(t); return rc; }

    const uint16_t *tab = t->table;
    const fse_or_symtt *stt = t->symbol_tt;
    const unsigned tl = t->table_log;
    fastw w = { dst + hb, 0, 0 };
    size_t i = n; /* pairs are consumed from the top: lib.rs:153-165 */
    uint32_t s0, s1;
    if (n & 1) {
        s0 = enc_first(t, src[n - 1]);
        s1 = enc_first(t, src[n - 2]);
        {   /* encode0.encode(next[0]) */
            fse_or_symtt tt = stt[src[n - 3]];
            uint32_t bo = (tt.bits + s0) >> 16;
            fastw_flush(&w);
            fastw_put(&w, s0, bo);
            s0 = tab[(size_t)((int32_t)(s0 >> bo) + tt.find_state)];
        }
        i = n - 3;
    } else {
        s0 = enc_first(t, src[n - 2]);
        s1 = enc_first(t, src[n - 1]);
        i = n - 2;
    }
    while (i >= 2) { /* lib.rs:167-176 */
        fse_or_symtt t1 = stt[src[i - 1]];
        fse_or_symtt t0 = stt[src[i - 2]];
        uint32_t b1 = (t1.bits + s1) >> 16;
        fastw_put(&w, s1, b1);
        s1 = tab[(size_t)((int32_t)(s1 >> b1) + t1.find_state)];
        uint32_t b0 = (t0.bits + s0) >> 16;
        fastw_put(&w, s0, b0);
        s0 = tab[(size_t)((int32_t)(s0 >> b0) + t0.find_state)];
        fastw_flush(&w);
        i -= 2;
    }
    fastw_put(&w, s1, tl); fastw_flush(&w); /* lib.rs:178-181 */
    fastw_put(&w, s0, tl); fastw_flush(&w);
    fastw_put(&w, 1, 1);   fastw_flush(&w);
    size_t total_bits = (size_t)(w.p - (dst + hb)) * 8 + w.bits; /* writer.rs:201-222 */
    uint64_t rest = w.acc;
    memcpy(w.p, &rest, 8);
    free(t);
    return hb + (long)((total_bits + 7) / 8);
}

typedef struct { const uint8_t *base; const uint8_t *p; uint64_t buf; unsigned bits; } fastr;

static inline void fastr_reload(fastr *r) /* stack_reader.rs:143-167 */
{
    if (r->bits <= 32 && r->p > r->base) {
        size_t avail = (size_t)(r->p - r->base);
        if (avail >= 4) {
            uint32_t v;
            r->p -= 4;
            memcpy(&v, r->p, 4);
            r->buf = (r->buf << 32) | v;
            r->bits += 32;
        } else {
            uint32_t v = 0;
            r->p = r->base;
            memcpy(&v, r->p, avail);
            r->buf = (r->buf << (8 * avail)) | v;
            r->bits += 8 * (unsigned)avail;
        }
    }
}

long fse_or_ref_decompress2(const uint8_t *src, size_t n, uint8_t *dst, size_t cap)
{
    fse_or_norm nh;
    size_t consumed;
    int rc;
    if (n == 0) return FSE_OR_ERR_PANIC;
    rc = fse_or_ncount_read(src, n, &nh, &consumed);
    if (rc < 0) return FSE_OR_ERR_NONE;
    const uint8_t *pay = src + consumed;
    size_t pn = n - consumed;
    if (pn == 0 || pay[pn - 1] == 0) return FSE_OR_ERR_NONE;
    fse_or_dec_table *t = (fse_or_dec_table *)malloc(sizeof(*t));
    if (!t) return FSE_OR_ERR_CAPACITY;
    if ((rc = fse_or_dec_table_build(&nh, t)) < 0) { free(t); return rc; }
    const fse_or_dec_entry *tab = t->table;
    const unsigned tl = t->table_log;

    fastr r = { pay, pay + pn - 1, pay[pn - 1], ilog2_u32(pay[pn - 1]) };
    fastr_reload(&r); fastr_reload(&r);
    uint32_t s0, s1;
    if (r.bits < tl) { free(t); return FSE_OR_ERR_PANIC; }
    s0 = (uint32_t)(r.buf >> (r.bits - tl)) & ((1u << tl) - 1); r.bits -= tl; fastr_reload(&r);
    if (r.bits < tl) { free(t); return FSE_OR_ERR_PANIC; }
    s1 = (uint32_t)(r.buf >> (r.bits - tl)) & ((1u << tl) - 1); r.bits -= tl; fastr_reload(&r);

    size_t out = 0;
    for (;;) { /* lib.rs:228-241 */
        fse_or_dec_entry e0 = tab[s0];
        if (e0.num_bits > r.bits) { /* decode0 -> None: lib.rs:242-243 */
            if (out + 2 > cap) { free(t); return FSE_OR_ERR_CAPACITY; }
            dst[out++] = tab[s0].symbol; dst[out++] = tab[s1].symbol;
            break;
        }
        if (out + 3 > cap) { free(t); return FSE_OR_ERR_CAPACITY; } /* Q1: would run until OOM */
        uint32_t lo0 = (uint32_t)(r.buf >> (r.bits - e0.num_bits)) & ((1u << e0.num_bits) - 1);
        r.bits -= e0.num_bits;
        s0 = (uint32_t)e0.new_state + lo0;
        dst[out++] = e0.symbol;
        fse_or_dec_entry e1 = tab[s1];
        if (e1.num_bits > r.bits) { /* decode1 -> None: lib.rs:235-239 */
            dst[out++] = tab[s1].symbol; dst[out++] = tab[s0].symbol;
            break;
        }
        uint32_t lo1 = (uint32_t)(r.buf >> (r.bits - e1.num_bits)) & ((1u << e1.num_bits) - 1);
        r.bits -= e1.num_bits;
        s1 = (uint32_t)e1.new_state + lo1;
        dst[out++] = e1.symbol;
        fastr_reload(&r);
    }
    free(t);
    return (long)out;
}

/* ---------------------------------------------------------- block drivers */

typedef struct {
    const uint8_t *src; size_t n; const fse_or_block_params *p;
    uint8_t *dst; size_t stride; uint64_t *sizes; int32_t *status;
    size_t nblocks; unsigned tid, nthreads; int decode;
    const uint8_t *csrc; uint8_t *out;
} blk_job;

static void *blk_worker(void *arg)
{
    blk_job *j = (blk_job *)arg;
    const size_t bs = j->p->block_size;
    for (size_t b = j->tid; b < j->nblocks; b += j->nthreads) {
        size_t off = b * bs;
        size_t len = (off + bs <= j->n) ? bs : (j->n - off);
        if (!j->decode) {
            long rc;
            if (j->p->use_ref2)
                rc = fse_or_ref_compress2(j->src + off, len, j->dst + b * j->stride, j->stride);
            else
                rc = fse_or_compress_n(j->src + off, len, j->p->table_log, j->p->n_states,
                                       j->dst + b * j->stride, j->stride, NULL, NULL);
            j->sizes[b] = rc < 0 ? 0 : (uint64_t)rc;
            j->status[b] = rc < 0 ? (int32_t)rc : 0;
        } else {
            int rc;
            if (j->p->use_ref2) {
                long r = fse_or_ref_decompress2(j->csrc + b * j->stride, (size_t)j->sizes[b], j->out + off, len);
                rc = (r == (long)len) ? 0 : (r < 0 ? (int)r : FSE_OR_ERR_LENGTH);
            } else {
                rc = fse_or_decompress_n_len(j->csrc + b * j->stride, (size_t)j->sizes[b],
                                             j->p->n_states, j->out + off, len);
            }
            j->status[b] = rc;
        }
    }
    return NULL;
}

static int run_blocks(blk_job *proto)
{
    unsigned nt = proto->p->threads ? proto->p->threads : 1;
    if (nt > 256) nt = 256;
    if (nt == 1) { proto->tid = 0; proto->nthreads = 1; blk_worker(proto); }
    else {
        pthread_t th[256];
        blk_job jobs[256];
        for (unsigned t = 0; t < nt; t++) {
            jobs[t] = *proto; jobs[t].tid = t; jobs[t].nthreads = nt;
            if (pthread_create(&th[t], NULL, blk_worker, &jobs[t]) != 0) return FSE_OR_ERR_CAPACITY;
        }
        for (unsigned t = 0; t < nt; t++) pthread_join(th[t], NULL);
    }
    int worst = 0;
    for (size_t b = 0; b < proto->nblocks; b++) if (proto->status[b] < worst) worst = proto->status[b];
    return worst;
}

int fse_or_compress_blocks(const uint8_t *src, size_t n, const fse_or_block_params *p,
                           uint8_t *dst, size_t stride, uint64_t *sizes, int32_t *status)
{
    if (p->block_size == 0) return FSE_OR_ERR_PANIC;
    blk_job j;
    memset(&j, 0, sizeof(j));
    j.src = src; j.n = n; j.p = p; j.dst = dst; j.stride = stride; j.sizes = sizes; j.status = status;
    j.nblocks = (n + p->block_size - 1) / p->block_size; j.decode = 0;
    return run_blocks(&j);
}

int fse_or_decompress_blocks(const uint8_t *src, size_t stride, const uint64_t *sizes, size_t nblocks,
                             const fse_or_block_params *p, uint8_t *dst, size_t n, int32_t *status)
{
    if (p->block_size == 0) return FSE_OR_ERR_PANIC;
    blk_job j;
    memset(&j, 0, sizeof(j));
    j.csrc = src; j.stride = stride; j.sizes = (uint64_t *)sizes; j.nblocks = nblocks; j.p = p;
    j.out = dst; j.n = n; j.status = status; j.decode = 1;
    return run_blocks(&j);
}

/* ------------------------------------------------ synthetic generators §8(d) */

static inline uint64_t splitmix64(uint64_t x)
{
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static const uint8_t TEXT_RANKS[96] = {
    0x20,0x65,0x74,0x61,0x6f,0x69,0x6e,0x73,0x68,0x72,0x64,0x6c,0x63,0x75,0x6d,0x77,
    0x66,0x67,0x79,0x70,0x62,0x76,0x6b,0x6a,0x78,0x71,0x7a,0x45,0x54,0x41,0x4f,0x49,
    0x4e,0x53,0x48,0x52,0x44,0x4c,0x43,0x55,0x4d,0x57,0x46,0x47,0x59,0x50,0x42,0x56,
    0x4b,0x4a,0x58,0x51,0x5a,0x30,0x31,0x32,0x33,0x34,0x35,0x36,0x37,0x38,0x39,0x2e,
    0x2c,0x3b,0x3a,0x27,0x22,0x21,0x3f,0x2d,0x28,0x29,0x0a,0x09,0x2f,0x26,0x25,0x24,
    0x23,0x40,0x2a,0x2b,0x3c,0x3d,0x3e,0x5b,0x5d,0x5f,0x7b,0x7d,0x7c,0x7e,0x5e,0x60 };

size_t fse_or_gen_lut(int kind, uint8_t *lut)
{
    if (kind == FSE_OR_GEN_GEO) { /* lib.rs:255-270 with prob = 0.2 */
        size_t remaining = 4096, idx = 0;
        uint8_t s = 0;
        while (remaining > 0) {
            size_t n = (size_t)((double)remaining * 0.2);
            if (n < 1) n = 1;
            for (size_t k = 0; k < n; k++) lut[idx++] = s;
            s++;
            remaining -= n;
        }
        return 4096;
    }
    if (kind == FSE_OR_GEN_TEXT) { /* Zipf(s=1) over 96 printable symbols, 65536-entry LUT */
        uint64_t w[96], W = 0;
        for (int r = 0; r < 96; r++) { w[r] = (1ull << 20) / (uint64_t)(r + 1); W += w[r]; }
        uint64_t cnt[96], used = 0;
        for (int r = 0; r < 96; r++) { cnt[r] = (65536ull * w[r]) / W; used += cnt[r]; }
        cnt[0] += 65536 - used;
        size_t idx = 0;
        for (int r = 0; r < 96; r++) for (uint64_t k = 0; k < cnt[r]; k++) lut[idx++] = TEXT_RANKS[r];
        return 65536;
    }
    if (kind == FSE_OR_GEN_FEW) {
        static const unsigned c[4] = {3686, 205, 123, 82};
        size_t idx = 0;
        for (int s = 0; s < 4; s++) for (unsigned k = 0; k < c[s]; k++) lut[idx++] = (uint8_t)s;
        return 4096;
    }
    for (int i = 0; i < 256; i++) lut[i] = (uint8_t)i; /* uniform */
    return 256;
}

void fse_or_generate(int kind, uint64_t seed, uint64_t first_index, uint8_t *dst, size_t n)
{
    uint8_t local[65536];
    size_t m = fse_or_gen_lut(kind, local);
    uint64_t mask = (uint64_t)m - 1;
    for (size_t k = 0; k < n; k++) {
        uint64_t i = first_index + k;
        uint64_t r16 = (splitmix64(seed + (i >> 2)) >> (16 * (i & 3))) & 0xFFFF;
        dst[k] = local[r16 & mask];
    }
}
