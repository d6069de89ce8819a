/*
 * fse_oracle.h -- CPU oracle for the FSE (tANS) hot path of Cognoscan/entropy_coders.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference
 * crate's arithmetic, used as the checker for the CUDA path.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.  The product (libfse_b200.so) never links, loads or calls it.
 *
 * PARITY PINNING: the reference is a Rust crate and there is no Rust toolchain in
 * this image, so the reference itself cannot be executed here ("parity unpinned"
 * at the encoded-byte level against a reference *execution*).  The oracle is
 * pinned against (i) every known-answer and property test the reference holds for
 * this path (histogram.rs:589-670, bitstream/mod.rs:112-224, lib.rs:280-302,
 * fse.rs:461-506 -- re-expressed in tests/), (ii) the hand-derived vector KAT-A of
 * SURVEY.md Appendix C, and (iii) a second, independently written Python model
 * (oracle/pymodel.py) that mirrors the reference's *mechanics* (64-bit
 * accumulator, aligned flushes, reload cadence) rather than its semantics.
 *
 * Every function cites the reference file:line it follows
 * (paths relative to /root/reference).
 */
#ifndef FSE_ORACLE_H
#define FSE_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* src/lib.rs:9-12 */
#define FSE_OR_TABLE_LOG_MIN 5
#define FSE_OR_TABLE_LOG_MAX 15
#define FSE_OR_TABLE_LOG_DEFAULT 11

/* Status codes.  Negative = the reference would panic / return None / Err here. */
enum {
    FSE_OR_OK = 0,
    FSE_OR_ERR_PANIC = -1,        /* reference panics (assert / unwrap / ilog2(0)) */
    FSE_OR_ERR_NONE = -2,         /* reference returns None (missing marker, empty) */
    FSE_OR_ERR_TABLE_LOG = -3,    /* HistError::TableLogTooLarge  histogram.rs:439-441 */
    FSE_OR_ERR_TOO_MANY = -4,     /* HistError::TooManySymbols    histogram.rs:498-500 */
    FSE_OR_ERR_IO = -5,           /* HistError::Io (UnexpectedEof) stream_reader.rs:70-72 */
    FSE_OR_ERR_CAPACITY = -6,     /* caller buffer too small (oracle-only) */
    FSE_OR_ERR_LENGTH = -7        /* length-driven decode: bits left over / missing */
};

/* histogram.rs:10-14.  Counts are u64 here so that the same code serves the
 * multi-GPU "global table" extension; for size <= u32::MAX the
 * arithmetic below is identical to the reference's u32/u64 mix. */
typedef struct {
    uint64_t table[256];
    uint64_t size;
    uint32_t table_len;
} fse_or_hist;

/* histogram.rs:290-294 */
typedef struct {
    int32_t table[256];
    uint32_t log2;
    uint32_t table_len;
} fse_or_norm;

/* fse.rs:80-84 */
typedef struct {
    uint32_t bits;
    int32_t find_state;
} fse_or_symtt;

/* fse.rs:72-78 */
typedef struct {
    uint32_t table_log;
    uint16_t table[1 << FSE_OR_TABLE_LOG_MAX];
    fse_or_symtt symbol_tt[256];
    uint8_t symbols[1 << FSE_OR_TABLE_LOG_MAX]; /* the spread, cell -> symbol */
} fse_or_enc_table;

/* fse.rs:260-265 */
typedef struct {
    uint16_t new_state;
    uint8_t symbol;
    uint8_t num_bits;
} fse_or_dec_entry;

/* fse.rs:253-258 */
typedef struct {
    uint32_t table_log;
    int fast_mode; /* computed and never read by the reference (fse.rs:256,290,306) */
    fse_or_dec_entry table[1 << FSE_OR_TABLE_LOG_MAX];
} fse_or_dec_table;

/* ---- histogram.rs ---- */
void fse_or_histogram(const uint8_t *data, size_t n, fse_or_hist *out);       /* :18-66 */
uint32_t fse_or_symbol_count(const int32_t *table256);                         /* :79-81 / :321-323 (counts ZERO entries, Q3) */
int fse_or_optimal_log2(const fse_or_hist *h, uint32_t *log2_out);             /* :264-277 */
/* returns FSE_OR_OK, or 1 when normalize_slow was taken (:144-145), or <0 */
int fse_or_normalize(const fse_or_hist *h, uint32_t log2, fse_or_norm *out);   /* :95-261 */
int fse_or_norm_new(const uint8_t *data, size_t n, fse_or_norm *out);          /* :299-303 */
/* libzstd's FSE_normalizeCount (not the crate's arithmetic; pinned against libzstd's own output, see fse_oracle.c) */
int fse_or_normalize_zstd(const fse_or_hist *h, uint32_t table_log, int use_low_prob_count, fse_or_norm *out);
size_t fse_or_write_bound(const fse_or_norm *nh);                              /* :330-337 */
/* appends at dst[0..]; returns bytes written (>=0) or <0; *bits_out = header bits */
long fse_or_ncount_write(const fse_or_norm *nh, uint8_t *dst, size_t cap, size_t *bits_out); /* :376-431 */
/* parses src[0..n); *consumed = byte offset of the remainder (finish_byte, :504) */
int fse_or_ncount_read(const uint8_t *src, size_t n, fse_or_norm *out, size_t *consumed);    /* :436-505 */
int fse_or_norm_try_from(const int32_t *table256, fse_or_norm *out);           /* :508-536 */

/* ---- fse.rs ---- */
size_t fse_or_table_step(size_t size);                                         /* :68-70 */
int fse_or_enc_table_build(const fse_or_norm *nh, fse_or_enc_table *t);        /* :101-189 */
size_t fse_or_compress_bound(size_t size);                                     /* :191-193 */
int fse_or_dec_table_build(const fse_or_norm *nh, fse_or_dec_table *t);        /* :280-338 */

/* ---- bitstream (semantics of writer.rs / stack_reader.rs / stream_reader.rs,
 *      SURVEY Appendix A.1-A.2) ---- */
typedef struct {
    uint8_t *buf;     /* caller storage */
    size_t cap;       /* bytes */
    size_t start;     /* initial_len (byte aligned start, writer.rs:18-20) */
    size_t bitpos;    /* bits written since start */
    int overflow;
} fse_or_bitw;
void fse_or_bitw_init(fse_or_bitw *w, uint8_t *buf, size_t cap, size_t start);
void fse_or_bitw_put(fse_or_bitw *w, uint64_t val, unsigned bits);  /* write_bits_unmasked, writer.rs:195-198 */
size_t fse_or_bitw_finish(fse_or_bitw *w, size_t *new_len);         /* writer.rs:201-222 -> bits since start */

typedef struct {
    const uint8_t *buf;
    size_t bits;      /* unread bits below the marker */
} fse_or_bitstack;
int fse_or_bitstack_init(fse_or_bitstack *r, const uint8_t *buf, size_t n);  /* stack_reader.rs:17-92 */
int fse_or_bitstack_peek(const fse_or_bitstack *r, unsigned bits, uint32_t *out); /* :176-184 */
int fse_or_bitstack_read(fse_or_bitstack *r, unsigned bits, uint32_t *out);  /* :211-215 */

typedef struct {
    const uint8_t *buf;
    size_t len;
    size_t total_bits;
    size_t bits_read;
} fse_or_bitstream;
int fse_or_bitstream_init(fse_or_bitstream *r, const uint8_t *buf, size_t n, size_t total_bits); /* stream_reader.rs:16-49 */
int fse_or_bitstream_peek(const fse_or_bitstream *r, unsigned bits, uint32_t *out);  /* :82-114 */
int fse_or_bitstream_advance(fse_or_bitstream *r, unsigned bits);                     /* :67-75 */
int fse_or_bitstream_read(fse_or_bitstream *r, unsigned bits, uint32_t *out);         /* :56-60 */

/* ---- lib.rs codecs, generalised to N interleaved states.
 *      N=1 == fse_compress (lib.rs:112-143), N=2 == fse_compress2 (lib.rs:146-183). ---- */

/* Payload only (the fse.rs:394-421 header-less variant) with a caller table.
 * Returns payload bytes or <0.  *bits_out = payload bit count incl. marker. */
long fse_or_encode_payload(const fse_or_enc_table *t, const uint8_t *src, size_t n,
                           unsigned n_states, uint8_t *dst, size_t cap, size_t *bits_out);

/* header || payload.  table_log == 0 -> optimal_log2 (NormHistogram::new). */
long fse_or_compress_n(const uint8_t *src, size_t n, uint32_t table_log, unsigned n_states,
                       uint8_t *dst, size_t cap, size_t *header_bytes_out, size_t *payload_bits_out);

/* Reference decode semantics: runs until the bit stack cannot supply num_bits
 * (lib.rs:187-248).  Over-produces when states with num_bits == 0 exist (Q1). */
long fse_or_decode_payload_exhaust(const fse_or_dec_table *t, const uint8_t *src, size_t n,
                                   unsigned n_states, uint8_t *dst, size_t cap);
/* Length-driven decode of exactly n_out symbols (what the GPU path does). */
int fse_or_decode_payload_len(const fse_or_dec_table *t, const uint8_t *src, size_t n,
                              unsigned n_states, uint8_t *dst, size_t n_out);

long fse_or_decompress_n_exhaust(const uint8_t *src, size_t n, unsigned n_states, uint8_t *dst, size_t cap);
int fse_or_decompress_n_len(const uint8_t *src, size_t n, unsigned n_states, uint8_t *dst, size_t n_out);

/* ---- the reference's own loop structure, for CPU-baseline timing
 *      (2 states, 64-bit accumulator, one flush per symbol pair: lib.rs:167-176;
 *       refill every other symbol: lib.rs:228-241).  Byte-identical to
 *       fse_or_compress_n(..., 2, ...) / fse_or_decompress_n_exhaust(..., 2, ...). ---- */
long fse_or_ref_compress2(const uint8_t *src, size_t n, uint8_t *dst, size_t cap);
long fse_or_ref_decompress2(const uint8_t *src, size_t n, uint8_t *dst, size_t cap);

/* ---- block drivers (new: the reference has no blocks).  Each block is handed to
 *      the per-stream codec exactly as the reference would be handed that slice. ---- */
typedef struct {
    uint32_t block_size;
    uint32_t table_log;   /* 0 = optimal_log2 per block */
    uint32_t n_states;
    uint32_t threads;     /* pthreads over blocks; 0/1 = inline */
    uint32_t use_ref2;    /* 1: use fse_or_ref_compress2 loops (requires n_states==2, table_log==0) */
} fse_or_block_params;

/* dst must hold nblocks * fse_or_compress_bound(block_size) when scratch-strided;
 * this driver writes block b at dst + b*stride and records sizes[b]; returns 0 or <0. */
int fse_or_compress_blocks(const uint8_t *src, size_t n, const fse_or_block_params *p,
                           uint8_t *dst, size_t stride, uint64_t *sizes, int32_t *status);
int fse_or_decompress_blocks(const uint8_t *src, size_t stride, const uint64_t *sizes, size_t nblocks,
                             const fse_or_block_params *p, uint8_t *dst, size_t n, int32_t *status);

/* ---- synthetic generators of SURVEY section 8(d) (integer exact). ---- */
enum { FSE_OR_GEN_GEO = 0, FSE_OR_GEN_TEXT = 1, FSE_OR_GEN_FEW = 2, FSE_OR_GEN_UNIFORM = 3 };
void fse_or_generate(int kind, uint64_t seed, uint64_t first_index, uint8_t *dst, size_t n);
/* writes the LUT used by kind into lut (<= 65536 entries), returns its length */
size_t fse_or_gen_lut(int kind, uint8_t *lut);

#ifdef __cplusplus
}
#endif
#endif
