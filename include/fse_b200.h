/*
 * fse_b200.h -- C ABI of libfse_b200.so: a B200 (sm_100a) implementation of the FSE / tANS hot
 * path of the Rust crate Cognoscan/entropy_coders.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ or torch types.  Every entry
 * point names the reference interface it replaces (paths relative to the crate root).  The Rust
 * side binds these with `extern "C"` (see INTEGRATION.md and rust/); in this repository the same
 * symbols are bound from Python with ctypes (entropy_coders_b200/_capi.py).
 *
 * The library has no CPU fallback: every call needs a CUDA device and fails with
 * FSE_B200_ERR_CUDA otherwise.
 *
 * Conventions
 *  - "d_" pointers are device pointers on the context's device, "h_" pointers are host pointers.
 *  - All functions return FSE_B200_OK (0) or a negative fse_b200_status.  Nothing throws or aborts.
 *  - Entry points are synchronous (they return after the context's stream has drained) unless
 *    the name ends in _async.
 *  - A context is single-owner: one host thread at a time.  Distinct contexts are independent.
 *
 * Block streams
 *  The reference codes ONE slice into ONE stream (src/lib.rs:112-183).  This library cuts the
 *  input into independent blocks of `block_size` bytes (last block may be shorter); block b's bytes
 *  are exactly what the reference emits when handed that block as `src`:
 *      [ NCount header (src/histogram.rs:376-431) ][ payload (src/lib.rs:118-142 / :151-182) ]
 *  with `n_states` interleaved encoder states (1 = fse_compress, 2 = fse_compress2; wider N is the
 *  same composition of fse::Encoder the crate documents at src/fse.rs:16-17: state j owns the
 *  symbols whose index is j modulo N, symbols are consumed in decreasing index order, final states
 *  are written N-1..0, then the marker bit).  The compressed blocks are concatenated densely;
 *  offsets[b]..offsets[b+1] delimit block b.
 *  Blocks on which the reference panics (all bytes zero: src/histogram.rs:98; fewer symbols than
 *  states or <= 4 bytes with automatic table_log: src/lib.rs:121,154, src/histogram.rs:271) are
 *  stored with an escape byte that no valid header can start with (low nibble > 10 means
 *  table_log > 15, rejected at src/histogram.rs:439-441):  0x0F = raw bytes follow,
 *  0x0E = one byte follows, repeated.  status[b] reports 1 / 2 for those.
 *
 * Which kernels a call takes (same bytes whichever it is)
 *  n_states 128 / 64: one warp per block, four / two states per lane (the fast formats: ~570 / 1 020 GB/s encode / decode
 *  on one B200, 8 GiB in 128 KiB blocks).  n_states 1 and 2 -- the crate's own fse_compress / fse_compress2 block
 *  format -- are one or two serial state chains per block: every block is coded by ONE THREAD with the block's tables in
 *  shared memory (a 128 KiB block takes ~4 ms however many there are, up to ~5 000 blocks at a time: 125-148 GB/s each
 *  way from 8 192 blocks per call on, 33 GB/s at 1 024, 9 GB/s at 256): hand such data over in calls of many blocks (the
 *  host-buffer entry points take chunks of 8 192).
 */
#ifndef FSE_B200_H
#define FSE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FSE_B200_TABLE_LOG_MIN 5      /* src/lib.rs:9  */
#define FSE_B200_TABLE_LOG_MAX 15     /* src/lib.rs:10 */
#define FSE_B200_TABLE_LOG_DEFAULT 11 /* src/lib.rs:12 */

typedef enum {
    FSE_B200_OK = 0,
    FSE_B200_ERR_ARG = -1,          /* bad argument (null pointer, block_size 0, n_states not supported ...) */
    FSE_B200_ERR_CAPACITY = -2,     /* caller buffer too small */
    FSE_B200_ERR_TABLE_LOG = -3,    /* HistError::TableLogTooLarge  src/histogram.rs:439-441 */
    FSE_B200_ERR_TOO_MANY = -4,     /* HistError::TooManySymbols    src/histogram.rs:498-500 */
    FSE_B200_ERR_IO = -5,           /* HistError::Io (header ran out of bits) src/bitstream/stream_reader.rs:70-72 */
    FSE_B200_ERR_NO_MARKER = -6,    /* BitStackReader::new -> None  src/bitstream/stack_reader.rs:18-20,77-83 */
    FSE_B200_ERR_LENGTH = -7,       /* payload bits do not match the block length / fewer than N*table_log bits (src/lib.rs:197,224) */
    FSE_B200_ERR_PANIC = -8,        /* the reference would panic on this input (e.g. src/histogram.rs:248,420) */
    FSE_B200_ERR_UNSUPPORTED = -9,  /* table_log above the limit this launch was sized for */
    FSE_B200_ERR_CUDA = -10,        /* CUDA runtime error; see fse_b200_last_error() */
    FSE_B200_ERR_BLOCK = -11        /* at least one block failed; inspect status[] */
} fse_b200_status;

/* per-block status values >= 0 */
#define FSE_B200_BLOCK_FSE 0
#define FSE_B200_BLOCK_RAW 1
#define FSE_B200_BLOCK_RLE 2

#define FSE_B200_TABLE_PER_BLOCK 0
#define FSE_B200_TABLE_GLOBAL 1

typedef struct {
    uint32_t block_size;  /* bytes per block (> 0) */
    uint32_t table_log;   /* 0 = Histogram::optimal_log2 per table (src/histogram.rs:264-277); else the
                             value handed to Histogram::normalize (src/histogram.rs:95), 5..15 */
    uint32_t n_states;    /* interleaved states per block: 1, 2, 4, 8, 16, 32, 64 or 128 (64 / 128: table_log <= 13) */
    uint32_t table_mode;  /* FSE_B200_TABLE_PER_BLOCK or FSE_B200_TABLE_GLOBAL */
    uint32_t segment_size; /* 0: a block is one stream (above).  > 0 (per-block tables, n_states 128, table_log <= 11,
                             a divisor of block_size, >= 512): the block keeps ONE table and ONE header, and its bytes
                             are coded as block_size / segment_size independent streams ("segments") against that table,
                             each the header-less payload the crate tests at src/fse.rs:394-421 for its slice.  One
                             CTA codes a block: its warps share one bank-replicated copy of the table in shared memory.
                             The index then has one entry per SEGMENT (fse_b200_num_streams): stream s = block * (block_size /
                             segment_size) + k; the first stream of a block starts with the block's header (or its
                             escape byte, then the other streams of the block are empty); a tail segment shorter
                             than 128 bytes is stored raw (status 1).  Wherever an entry point says "nblocks",
                             "offsets[nblocks + 1]" or "status[nblocks]", read fse_b200_num_streams(n, p). */
    uint32_t flags;       /* FSE_B200_FLAG_* */
} fse_b200_params;

/* A block (or segment) whose coded form would be larger than 1 + its length is stored as 0x0F + raw bytes instead
 * (status 1).  Off by default: the reference has no such fallback (src/fse.rs:191-193 only bounds the expansion),
 * so the default output stays byte-identical to it. */
#define FSE_B200_FLAG_RAW_IF_EXPANDS 1u

/* src/fse.rs:80-84 SymbolTransform */
typedef struct { uint32_t bits; int32_t find_state; } fse_b200_symbol_transform;
/* src/fse.rs:260-265 DecodeTransform */
typedef struct { uint16_t new_state; uint8_t symbol; uint8_t num_bits; } fse_b200_decode_transform;

typedef struct fse_b200_ctx fse_b200_ctx;

/* ---- context ------------------------------------------------------------------------------- */
/* stream: a cudaStream_t, or NULL for a context-owned non-blocking stream (then inputs produced on other
 * streams must be complete, or ordered by an event, before a call; pass cudaStreamLegacy to run on the
 * legacy default stream).  All work of a context is issued on this stream.  The context owns device
 * workspaces that grow on demand and are reused across calls (the analogue of
 * EncodeTable::update / DecodeTable::update reusing their Vecs, src/fse.rs:101,280). */
int fse_b200_create(int device, void *stream, fse_b200_ctx **out);
void fse_b200_destroy(fse_b200_ctx *ctx);
const char *fse_b200_last_error(const fse_b200_ctx *ctx);
const char *fse_b200_version(void);
/* number of kernels launched by this context so far (bench.py's gpu_launches) */
uint64_t fse_b200_launch_count(const fse_b200_ctx *ctx);
int fse_b200_sync(fse_b200_ctx *ctx);

/* Per-kernel device timing of the fused pipelines (CUDA events on the context's stream around each
 * launch).  set_timing(1) starts a fresh measurement; get_timing synchronises and returns the summed
 * milliseconds and launch counts per kernel, indexed by the FSE_B200_K_* constants. */
#define FSE_B200_K_HIST 0     /* k_hist_blocks   */
#define FSE_B200_K_ENCODE 1   /* k_encode_blocks */
#define FSE_B200_K_SCAN 2     /* k_scan_sizes    */
#define FSE_B200_K_GATHER 3   /* k_gather        */
#define FSE_B200_K_DECODE 4   /* k_decode_blocks */
#define FSE_B200_NUM_KERNELS 5
int fse_b200_set_timing(fse_b200_ctx *ctx, int enable);
int fse_b200_get_timing(fse_b200_ctx *ctx, double *ms_total, uint64_t *count);

/* ---- sizing -------------------------------------------------------------------------------- */
/* EncodeTable::compress_bound, src/fse.rs:191-193 */
size_t fse_b200_compress_bound(size_t size);
/* worst-case bytes of the dense output of fse_b200_compress_blocks for n input bytes */
size_t fse_b200_compress_blocks_bound(size_t n, const fse_b200_params *p);
size_t fse_b200_num_blocks(size_t n, uint32_t block_size);
/* entries of the stream index: blocks, or segments when p->segment_size > 0 */
size_t fse_b200_num_streams(size_t n, const fse_b200_params *p);

/* ---- stage entry points (device pointers) --------------------------------------------------- */
/* Histogram::new per block, src/histogram.rs:18-66.
 * d_counts: uint32[nblocks*256]; d_table_len: uint32[nblocks] (highest present symbol + 1, :52-59). */
int fse_b200_histogram_blocks(fse_b200_ctx *ctx, const uint8_t *d_src, size_t n, uint32_t block_size,
                              uint32_t *d_counts, uint32_t *d_table_len);
/* Whole-buffer histogram accumulated in 64 bit (the multi-GPU global table sums these with an
 * all-reduce).  d_counts64: uint64[256], overwritten. */
int fse_b200_histogram_global(fse_b200_ctx *ctx, const uint8_t *d_src, size_t n, uint64_t *d_counts64);

/* Histogram::optimal_log2 + Histogram::normalize (+ normalize_slow), src/histogram.rs:95-277, for
 * `ntables` histograms.  d_counts64: uint64[ntables*256].  table_log 0 = optimal_log2.
 * Outputs: d_norm int32[ntables*256]; d_log2 uint32[ntables] (effective, may be raised: :96-98);
 * d_table_len uint32[ntables]; d_status int32[ntables] (0, 1 = normalize_slow taken, <0 error). */
int fse_b200_normalize(fse_b200_ctx *ctx, const uint64_t *d_counts64, size_t ntables, uint32_t table_log,
                       int32_t *d_norm, uint32_t *d_log2, uint32_t *d_table_len, int32_t *d_status);

/* A second normaliser: libzstd's FSE_normalizeCount + FSE_normalizeM2 (zstd 1.5.x
 * lib/compress/fse_compress.c), for tables that must equal the ones libzstd's entropy stage builds from the same counts.
 * The crate shares zstd's NCount header format (src/histogram.rs:342) but not every normalisation detail (the
 * `to_distribute != 0` guard at :144; low-probability symbols are always -1 there, here only with use_low_prob_count;
 * table_log 5..12, >= FSE_minTableLog, 0 = 11).  d_status: 0; 3 = one symbol holds every count (zstd's RLE case,
 * d_norm all zero); FSE_B200_ERR_TABLE_LOG (> 12); FSE_B200_ERR_PANIC (zstd's ERROR(GENERIC)).  Not part of the
 * reference's path; the outputs feed fse_b200_ncount_write / build_*_tables like those of fse_b200_normalize. */
int fse_b200_normalize_zstd(fse_b200_ctx *ctx, const uint64_t *d_counts64, size_t ntables, uint32_t table_log,
                            int use_low_prob_count, int32_t *d_norm, uint32_t *d_log2, uint32_t *d_table_len,
                            int32_t *d_status);

/* NormHistogram::write, src/histogram.rs:376-431.  d_out: ntables rows of `stride` bytes;
 * d_bytes / d_bits: uint32[ntables] (bytes appended / header bits, the latter is write()'s return). */
int fse_b200_ncount_write(fse_b200_ctx *ctx, const int32_t *d_norm, const uint32_t *d_log2,
                          const uint32_t *d_table_len, size_t ntables,
                          uint8_t *d_out, size_t stride, uint32_t *d_bytes, uint32_t *d_bits);
/* NormHistogram::read, src/histogram.rs:436-505.  d_in rows of `stride` bytes with d_len[t] valid.
 * d_consumed: byte offset of the remainder (finish_byte, src/bitstream/stream_reader.rs:132-135). */
int fse_b200_ncount_read(fse_b200_ctx *ctx, const uint8_t *d_in, size_t stride, const uint32_t *d_len,
                         size_t ntables, int32_t *d_norm, uint32_t *d_log2, uint32_t *d_table_len,
                         uint32_t *d_consumed, int32_t *d_status);

/* EncodeTable::new / update, src/fse.rs:88-189.  max_table_log sizes the rows: row t of
 * d_table has 1<<max_table_log uint16 entries (first 1<<log2[t] valid); d_symbol_tt: [ntables*256];
 * d_symbols (the spread, src/fse.rs:139-151): uint8 rows of 1<<max_table_log, may be NULL. */
int fse_b200_build_encode_tables(fse_b200_ctx *ctx, const int32_t *d_norm, const uint32_t *d_log2,
                                 const uint32_t *d_table_len, size_t ntables, uint32_t max_table_log,
                                 uint16_t *d_table, fse_b200_symbol_transform *d_symbol_tt,
                                 uint8_t *d_symbols, int32_t *d_status);
/* DecodeTable::new / update, src/fse.rs:269-338. */
int fse_b200_build_decode_tables(fse_b200_ctx *ctx, const int32_t *d_norm, const uint32_t *d_log2,
                                 const uint32_t *d_table_len, size_t ntables, uint32_t max_table_log,
                                 fse_b200_decode_transform *d_table, int32_t *d_status);

/* ---- bit I/O primitives (device pointers; one warp each: the packing / reading code of the coders on its own) ---- */
/* BitStackWriter, src/bitstream/writer.rs:140-222: the n fields (d_vals[i] masked to d_bits[i] <= 16 bits, as
 * write_bits_unmasked :195-198) are appended LSB first in index order, then a marker bit when `mark` (src/lib.rs:141,181).
 * d_out: 4-byte aligned, out_cap bytes; *h_nbits = bits written (finish(), :220-221); the stream occupies
 * ceil(bits / 8) bytes, zero padded. */
int fse_b200_bitstack_write(fse_b200_ctx *ctx, const uint32_t *d_vals, const uint8_t *d_bits, size_t n, int mark,
                            uint8_t *d_out, size_t out_cap, uint64_t *h_nbits);
/* BitStackReader, src/bitstream/stack_reader.rs:17-226: finds the marker (highest set bit of the last byte) and reads
 * the fields from the END (field n-1 first; d_vals[i] receives field i).  *h_status: FSE_B200_OK when every field was
 * supplied and nothing is left (finish() == true), FSE_B200_ERR_NO_MARKER (new -> None), FSE_B200_ERR_LENGTH (read -> None
 * or bits left over). */
int fse_b200_bitstack_read(fse_b200_ctx *ctx, const uint8_t *d_in, size_t nbytes, const uint8_t *d_bits, size_t n,
                           uint32_t *d_vals, int32_t *h_status);
/* BitStreamReader, src/bitstream/stream_reader.rs:16-135: forward reads under a total_bits bound.  *h_status: bits left
 * (finish(), >= 0), FSE_B200_ERR_IO on UnexpectedEof, FSE_B200_ERR_PANIC when the reference's constructor asserts (:17-21). */
int fse_b200_bitstream_read(fse_b200_ctx *ctx, const uint8_t *d_in, size_t nbytes, uint64_t total_bits, const uint8_t *d_bits,
                            size_t n, uint32_t *d_vals, int32_t *h_status);

/* ---- fused pipelines (device pointers) ------------------------------------------------------ */
/* fse_compress / fse_compress2 per block, src/lib.rs:112-183 (histogram + normalise + header +
 * encode table + encode), then an exclusive scan of the block sizes and a gather into d_dst.
 *   d_dst      : dense output, capacity dst_cap >= fse_b200_compress_blocks_bound(n, p)
 *   d_offsets  : uint64[nblocks+1] byte offsets into d_dst
 *   d_status   : int32[nblocks]
 *   h_total    : host, total compressed bytes (= offsets[nblocks])
 * In FSE_B200_TABLE_GLOBAL mode the blocks share the table installed with
 * fse_b200_set_global_table and carry no header (the header-less variant the crate tests at
 * src/fse.rs:394-421). */
int fse_b200_compress_blocks(fse_b200_ctx *ctx, const uint8_t *d_src, size_t n, const fse_b200_params *p,
                             uint8_t *d_dst, size_t dst_cap, uint64_t *d_offsets, int32_t *d_status,
                             uint64_t *h_total);
int fse_b200_compress_blocks_async(fse_b200_ctx *ctx, const uint8_t *d_src, size_t n, const fse_b200_params *p,
                                   uint8_t *d_dst, size_t dst_cap, uint64_t *d_offsets, int32_t *d_status);

/* fse_decompress / fse_decompress2 per block, src/lib.rs:187-248 (header parse + decode table +
 * decode), length driven: block b yields exactly min(block_size, n - b*block_size) bytes and must
 * consume its payload exactly (the reference stops on bit exhaustion and over-produces when the
 * final states need 0 bits -- SURVEY.md Q1). */
int fse_b200_decompress_blocks(fse_b200_ctx *ctx, const uint8_t *d_comp, size_t comp_bytes,
                               const uint64_t *d_offsets, size_t nblocks, const fse_b200_params *p,
                               uint8_t *d_dst, size_t n, int32_t *d_status);
int fse_b200_decompress_blocks_async(fse_b200_ctx *ctx, const uint8_t *d_comp, size_t comp_bytes,
                                     const uint64_t *d_offsets, size_t nblocks, const fse_b200_params *p,
                                     uint8_t *d_dst, size_t n, int32_t *d_status);

/* The reference's own termination rule (src/lib.rs:198, :228-241): no length is stored, decoding
 * stops when the bit stack cannot supply num_bits.  Block b may produce up to p->block_size bytes
 * into d_dst + b*block_size; d_out_len[b] receives the count.  A stream that would run past the
 * capacity (the reference never terminates on it) gets FSE_B200_ERR_CAPACITY.
 * p->table_log is the largest table_log accepted (0 = 11). */
int fse_b200_decompress_exhaust(fse_b200_ctx *ctx, const uint8_t *d_comp, size_t comp_bytes,
                                const uint64_t *d_offsets, size_t nblocks, const fse_b200_params *p,
                                uint8_t *d_dst, uint32_t *d_out_len, int32_t *d_status);

/* Global-table mode: normalise d_counts64 (uint64[256], e.g. the all-reduced sum of
 * fse_b200_histogram_global over all ranks) with `table_log` and keep the encode and decode tables
 * in the context.  h_header receives the NCount header (stored once for the whole job);
 * *h_header_bytes in: capacity, out: bytes. */
int fse_b200_set_global_table(fse_b200_ctx *ctx, const uint64_t *d_counts64, uint32_t table_log,
                              uint8_t *h_header, size_t *h_header_bytes, uint32_t *h_log2);
/* The table must give every byte value that occurs in the data a non-zero count (build it from the histogram of the
 * data, as fse_b200_frame_compress_host does): like the crate's Encoder, the encode kernels do not look for symbols the
 * table does not know, and a block that contains one is reported with status 0 but cannot be decoded. */
/* The check the encode kernels leave out: *h_unknown receives the number of bytes of d_src whose value has no entry in
 * the installed global table (0 = every block of d_src can be coded with it).  One histogram pass over d_src. */
int fse_b200_global_table_covers(fse_b200_ctx *ctx, const uint8_t *d_src, size_t n, uint64_t *h_unknown);
/* Same, from a stored header (decode side). */
int fse_b200_set_global_table_from_header(fse_b200_ctx *ctx, const uint8_t *h_header, size_t header_bytes,
                                          uint32_t *h_log2);

/* ---- host-buffer conveniences (what a crate user calls; copies are part of the call) --------- */
/* src/lib.rs:112 / :146 over blocks: h_src -> h_dst.  h_offsets uint64[nblocks+1], h_status int32[nblocks]
 * (either may be NULL).  *h_total out. */
int fse_b200_compress_host(fse_b200_ctx *ctx, const uint8_t *h_src, size_t n, const fse_b200_params *p,
                           uint8_t *h_dst, size_t dst_cap, uint64_t *h_offsets, int32_t *h_status,
                           uint64_t *h_total);
/* src/lib.rs:187 / :215 over blocks. */
int fse_b200_decompress_host(fse_b200_ctx *ctx, const uint8_t *h_comp, size_t comp_bytes,
                             const uint64_t *h_offsets, size_t nblocks, const fse_b200_params *p,
                             uint8_t *h_dst, size_t n, int32_t *h_status);

/* ---- self-describing frame (the container the reference does not have) ------- */
/* Layout (little endian):  "FSEB" | u16 version = 2 | u16 n_states | u32 block_size | u32 table_log |
 * u32 table_mode | u32 global_header_bytes | u64 n | u64 nstreams | u64 payload_bytes | u32 segment_size | u32 flags |
 * global NCount header (table_mode 1 only, padded to 8 bytes) | u64 offsets[nstreams + 1] | payload.
 * Everything a decoder needs travels with the data; the payload is the dense block streams above. */
#define FSE_B200_FRAME_MAGIC 0x42455346u /* "FSEB" */
size_t fse_b200_frame_bound(size_t n, const fse_b200_params *p);
/* Compress h_src into one frame.  In FSE_B200_TABLE_GLOBAL mode the table is built here from the
 * whole-buffer histogram (p->table_log).  *h_frame_bytes out. */
int fse_b200_frame_compress_host(fse_b200_ctx *ctx, const uint8_t *h_src, size_t n, const fse_b200_params *p,
                                 uint8_t *h_frame, size_t frame_cap, size_t *h_frame_bytes);
/* Parse a frame header: parameters and the uncompressed size (host only, no GPU work). */
int fse_b200_frame_info(const uint8_t *h_frame, size_t frame_bytes, fse_b200_params *p_out, size_t *n_out);
/* Decompress a frame into h_dst (capacity dst_cap >= n).  *h_n out. */
int fse_b200_frame_decompress_host(fse_b200_ctx *ctx, const uint8_t *h_frame, size_t frame_bytes, uint8_t *h_dst,
                                   size_t dst_cap, size_t *h_n);

/* Synthetic byte streams of SURVEY.md section 8(d), generated on the device (bench input).
 * kind: 0 geometric(0.2) (the crate's gen_sequence, src/lib.rs:255-278), 1 text-like, 2 few-symbol,
 * 3 uniform.  Byte i depends only on (seed, first_index + i). */
int fse_b200_generate(fse_b200_ctx *ctx, int kind, uint64_t seed, uint64_t first_index, uint8_t *d_dst, size_t n);

#ifdef __cplusplus
}
#endif
#endif /* FSE_B200_H */
