// entropy_coders.hpp -- C++ host-side mirror of the reference crate's public API for the FSE path
// (Cognoscan/entropy_coders src/lib.rs:7, :112-248), header only, on top of the C ABI of fse_b200.h.
//
// The reference is Rust; no Rust toolchain exists where this was built, so this is the compiled-language
// front that keeps the crate's names, argument meaning and error behaviour:
//   * where the crate panics, these throw entropy_coders::Panic;
//   * where it returns None / Err they return std::nullopt / throw entropy_coders::HistError;
//   * fse_compress* append to the caller's vector and return the payload BIT count (writer.rs:220-221).
// All arithmetic runs in libfse_b200.so on the GPU; there is no CPU fallback.
//
//   g++ -std=c++17 -I include -I /usr/local/cuda/include app.cpp -L entropy_coders_b200 -lfse_b200 -lcudart
#pragma once
#include <cuda_runtime.h>

#include <array>
#include <cstdint>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "fse_b200.h"

namespace entropy_coders {

constexpr uint32_t TABLE_LOG_MIN = 5, TABLE_LOG_MAX = 15, TABLE_LOG_DEFAULT = 11;  // src/lib.rs:9-12

struct Panic : std::runtime_error { using std::runtime_error::runtime_error; };
struct HistError : std::runtime_error {                    // src/histogram.rs:538-546
    enum Kind { TableLogTooLarge, TooManySymbols, Io } kind;
    HistError(Kind k, const char *what) : std::runtime_error(what), kind(k) {}
};

namespace detail {
inline fse_b200_ctx *ctx()
{
    static fse_b200_ctx *c = [] {
        fse_b200_ctx *p = nullptr;
        if (fse_b200_create(0, nullptr, &p) != FSE_B200_OK) throw std::runtime_error("fse_b200_create failed: CUDA device required");
        return p;
    }();
    return c;
}
inline void ck(int rc, const char *what)
{
    if (rc != FSE_B200_OK) throw std::runtime_error(std::string(what) + ": " + fse_b200_last_error(ctx()) + " (" + std::to_string(rc) + ")");
}
template <typename T> struct Dev {                          // a device array with host copies in and out
    T *p = nullptr; size_t n = 0;
    explicit Dev(size_t count) : n(count) { if (cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T)) != cudaSuccess) throw std::bad_alloc(); }
    Dev(const T *h, size_t count) : Dev(count) { if (count) cudaMemcpy(p, h, count * sizeof(T), cudaMemcpyHostToDevice); }
    ~Dev() { cudaFree(p); }
    Dev(const Dev &) = delete;
    std::vector<T> host() const { std::vector<T> v(n); if (n) cudaMemcpy(v.data(), p, n * sizeof(T), cudaMemcpyDeviceToHost); return v; }
    T at(size_t i) const { T v; cudaMemcpy(&v, p + i, sizeof(T), cudaMemcpyDeviceToHost); return v; }
};
inline uint32_t ilog2(uint64_t v) { if (!v) throw Panic("ilog2 of zero"); return 63u - (uint32_t)__builtin_clzll(v); }
}  // namespace detail

class NormHistogram;

class Histogram {                                          // src/histogram.rs:10-91
    std::array<uint32_t, 256> table_{};
    uint32_t size_ = 0;
    size_t table_len_ = 1;
public:
    explicit Histogram(const std::vector<uint8_t> &data)  // Histogram::new, :18-66
    {
        if (data.size() > 0xFFFFFFFFull) throw Panic("Data vector is too long");
        size_ = (uint32_t)data.size();
        if (data.empty()) return;
        detail::Dev<uint8_t> d(data.data(), data.size());
        detail::Dev<uint32_t> counts(256), tlen(1);
        detail::ck(fse_b200_histogram_blocks(detail::ctx(), d.p, data.size(), (uint32_t)data.size(), counts.p, tlen.p), "histogram");
        auto c = counts.host();
        std::copy(c.begin(), c.end(), table_.begin());
        table_len_ = tlen.at(0);
    }
    const std::array<uint32_t, 256> &table() const { return table_; }
    size_t table_len() const { return table_len_; }
    uint32_t size() const { return size_; }
    size_t symbol_count() const { size_t z = 0; for (auto x : table_) z += (x == 0); return z; }   // :79-81 counts zeros
    uint32_t optimal_log2() const                          // :264-277
    {
        uint32_t min_bits = std::min(detail::ilog2(size_) + 1, detail::ilog2(table_len_ - 1) + 2);
        uint32_t l = detail::ilog2(size_ - 1);
        if (l < 2) throw Panic("attempt to subtract with overflow");
        uint32_t v = std::max(std::min(TABLE_LOG_DEFAULT, l - 2), min_bits);
        return std::min(std::max(v, TABLE_LOG_MIN), TABLE_LOG_MAX);
    }
    NormHistogram normalize(uint32_t log2) const;          // :95-155 (+ :157-261)
    NormHistogram normalize_optimal() const;               // :281-284
};

class NormHistogram {                                      // src/histogram.rs:289-506
    std::array<int32_t, 256> table_{};
    uint32_t log2_ = 0;
    size_t table_len_ = 0;
    friend class Histogram;
public:
    NormHistogram() = default;
    NormHistogram(const std::array<int32_t, 256> &t, uint32_t log2, size_t table_len) : table_(t), log2_(log2), table_len_(table_len) {}
    static NormHistogram create(const std::vector<uint8_t> &data)   // NormHistogram::new, :299-303
    {
        Histogram h(data);
        return h.normalize(h.optimal_log2());
    }
    const std::array<int32_t, 256> &table() const { return table_; }
    uint32_t log2_sum() const { return log2_; }
    size_t table_len() const { return table_len_; }
    size_t symbol_count() const { size_t z = 0; for (auto x : table_) z += (x == 0); return z; }
    size_t write_bound() const { return table_len_ > 1 ? ((table_len_ * log2_) >> 3) + 3 : 512; }   // :330-337
    bool operator==(const NormHistogram &o) const { return table_ == o.table_ && log2_ == o.log2_ && table_len_ == o.table_len_; }

    size_t write(std::vector<uint8_t> &writer) const       // :376-431 -> bits written
    {
        detail::Dev<int32_t> norm(table_.data(), 256);
        uint32_t l2 = log2_, tl = (uint32_t)table_len_;
        detail::Dev<uint32_t> dl2(&l2, 1), dtl(&tl, 1), nbytes(1), nbits(1);
        detail::Dev<uint8_t> out(512);
        detail::ck(fse_b200_ncount_write(detail::ctx(), norm.p, dl2.p, dtl.p, 1, out.p, 512, nbytes.p, nbits.p), "ncount_write");
        auto bytes = out.host();
        writer.insert(writer.end(), bytes.begin(), bytes.begin() + nbytes.at(0));
        return nbits.at(0);
    }
    // :436-505 -> (histogram, offset of the remaining bytes)
    static std::pair<NormHistogram, size_t> read(const uint8_t *data, size_t len)
    {
        if (len == 0) throw Panic("No bytes provided to read from");       // stream_reader.rs:17
        detail::Dev<uint8_t> in(data, len);
        uint32_t l = (uint32_t)len;
        detail::Dev<uint32_t> dlen(&l, 1), dl2(1), dtl(1), cons(1);
        detail::Dev<int32_t> norm(256), st(1);
        detail::ck(fse_b200_ncount_read(detail::ctx(), in.p, len, dlen.p, 1, norm.p, dl2.p, dtl.p, cons.p, st.p), "ncount_read");
        int32_t rc = st.at(0);
        if (rc == FSE_B200_ERR_TABLE_LOG) throw HistError(HistError::TableLogTooLarge, "Table log2 size is higher than the accepted maximum");
        if (rc == FSE_B200_ERR_TOO_MANY) throw HistError(HistError::TooManySymbols, "Histogram counts are spread across more than 256 symbols");
        if (rc == FSE_B200_ERR_IO) throw HistError(HistError::Io, "Read error");
        if (rc < 0) throw Panic("NormHistogram::read");
        NormHistogram h;
        auto t = norm.host();
        std::copy(t.begin(), t.end(), h.table_.begin());
        h.log2_ = dl2.at(0);
        h.table_len_ = dtl.at(0);
        return {h, cons.at(0)};
    }
};

inline NormHistogram Histogram::normalize(uint32_t log2) const
{
    if (table_len_ <= 1 || size_ == 0) throw Panic("ilog2 of zero");     // :98 / :103
    std::array<uint64_t, 256> c64;
    for (int i = 0; i < 256; i++) c64[i] = table_[i];
    detail::Dev<uint64_t> counts(c64.data(), 256);
    detail::Dev<int32_t> norm(256), st(1);
    detail::Dev<uint32_t> dl2(1), dtl(1);
    log2 = std::min(std::max(log2, TABLE_LOG_MIN), TABLE_LOG_MAX);
    detail::ck(fse_b200_normalize(detail::ctx(), counts.p, 1, log2, norm.p, dl2.p, dtl.p, st.p), "normalize");
    if (st.at(0) < 0) throw Panic("Histogram::normalize");
    NormHistogram h;
    auto t = norm.host();
    std::copy(t.begin(), t.end(), h.table_.begin());
    h.log2_ = dl2.at(0);
    h.table_len_ = dtl.at(0);
    return h;
}
inline NormHistogram Histogram::normalize_optimal() const { return normalize(optimal_log2()); }


// ------------------------------------------------------------------------------------------------
// bitstream: the crate's bit I/O objects (src/bitstream/mod.rs:8-15) as plain host C++.  They are the
// host-side per-call objects of the API; the block pipelines run the same arithmetic per lane on the GPU
// (BitRowS / warp_place, the payload ring) and are checked against these through the oracle.  Byte-level
// behaviour follows the reference: bit k of a stream is bit (k mod 8) of byte (start + k div 8), fields are
// laid down LSB first in call order, the last byte is zero padded (writer.rs:163-180, :201-222).
// ------------------------------------------------------------------------------------------------
namespace bitstream {

struct UnexpectedEof : std::runtime_error { UnexpectedEof() : std::runtime_error("failed to fill whole buffer") {} };   // io::ErrorKind::UnexpectedEof

class BitStackWriter {                                     // src/bitstream/writer.rs:5-223
    std::vector<uint8_t> &out_;
    size_t start_;                                         // a writer starts at the vector's length: on a byte boundary (:18-20)
    uint64_t acc_ = 0;                                     // bits not yet in the vector, LSB first
    unsigned held_ = 0;
    size_t total_ = 0;
public:
    explicit BitStackWriter(std::vector<uint8_t> &writer) : out_(writer), start_(writer.size()) {}
    // whole bytes leave the accumulator (writer.rs:43-110 does it in aligned 32-bit stores; the bytes are the same)
    void flush()
    {
        while (held_ >= 8) { out_.push_back((uint8_t)acc_); acc_ >>= 8; held_ -= 8; }
    }
    // val must have no bits above `bits` (debug-asserted in the reference, :170-175); bits <= 16 (:142-146)
    void write_bits(size_t val, size_t bits)
    {
        if (bits > 16) throw Panic("write_bits: at most 16 bits per call");
        if (bits < 64 && (val >> bits) != 0) throw Panic("write_bits: value has bits above the field");
        acc_ |= (uint64_t)val << held_;
        held_ += (unsigned)bits;
        total_ += bits;
        flush();
    }
    void write_bits_unmasked(size_t val, size_t bits) { write_bits(val & (((size_t)1 << bits) - 1), bits); }   // :195-198
    // bits written since new (the reference's comment says bytes; it returns bits: :220-221, bitstream/mod.rs:38-47)
    size_t finish()
    {
        flush();
        if (held_) { out_.push_back((uint8_t)acc_); acc_ = 0; held_ = 0; }
        out_.resize(start_ + (total_ + 7) / 8);
        return total_;
    }
};

class BitStackReader {                                     // src/bitstream/stack_reader.rs:5-227
    const uint8_t *p_;
    size_t bits_;                                          // unread bits below the marker
    BitStackReader(const uint8_t *p, size_t bits) : p_(p), bits_(bits) {}
public:
    // None on an empty slice or when the last byte is zero: its highest set bit is the marker (:18-20, :77-83)
    static std::optional<BitStackReader> create(const uint8_t *data, size_t len)
    {
        if (len == 0 || data[len - 1] == 0) return std::nullopt;
        return BitStackReader(data, (len - 1) * 8 + (31 - __builtin_clz((unsigned)data[len - 1])));
    }
    // the n most recently written unread bits with their original significance (:176-184); None when fewer are left
    std::optional<size_t> peek(size_t n) const
    {
        if (n > 16) throw Panic("peek: at most 16 bits per call");
        if (n > bits_) return std::nullopt;
        size_t lo = bits_ - n, v = 0;
        for (size_t k = 0; k < n; k++) v |= (size_t)((p_[(lo + k) >> 3] >> ((lo + k) & 7)) & 1) << k;
        return v;
    }
    std::optional<size_t> read(size_t n)                   // :211-215
    {
        auto v = peek(n);
        if (v) bits_ -= n;
        return v;
    }
    void reload() {}                                       // the refill of the 64-bit window (:97-172) has no observable effect here
    size_t available() const { return bits_; }             // unread bits (the reference reports those in its window; both are 0 at the end)
    bool finish() const { return bits_ == 0; }             // :224-226
};

class BitStreamReader {                                    // src/bitstream/stream_reader.rs:5-136
    const uint8_t *p_;
    size_t len_, total_bits_, read_ = 0;
public:
    BitStreamReader(const uint8_t *data, size_t len, size_t total_bits) : p_(data), len_(len), total_bits_(total_bits)
    {
        if (len == 0) throw Panic("No bytes provided to read from");                      // :17
        if ((total_bits + 7) / 8 != len) throw Panic("Total number of bytes should be exactly enough to contain the total number of bits");   // :18-21
    }
    size_t peek(size_t n) const                            // :82-114
    {
        if (n > 32) throw Panic("peek: at most 32 bits per call");
        if (read_ + n > total_bits_) throw UnexpectedEof();
        size_t v = 0;
        for (size_t k = 0; k < n; k++) v |= (size_t)((p_[(read_ + k) >> 3] >> ((read_ + k) & 7)) & 1) << k;
        return v;
    }
    void advance_by(size_t n)                              // :67-75
    {
        if (read_ + n > total_bits_) throw UnexpectedEof();
        read_ += n;
    }
    size_t read(size_t n) { size_t v = peek(n); advance_by(n); return v; }   // :56-60
    size_t available() const { return total_bits_ - read_; }                 // :117-119
    // (remaining bytes, bits left, bit offset into the first of them) (:123-128)
    struct Rest { const uint8_t *data; size_t len, bits_left, bit_offset; };
    Rest finish() const { return {p_ + read_ / 8, len_ - read_ / 8, total_bits_ - read_, read_ % 8}; }
    // complete the current byte and return what is left (:132-135)
    std::pair<const uint8_t *, size_t> finish_byte() const { size_t b = (read_ + 7) / 8; return {p_ + b, len_ - b}; }
};

}  // namespace bitstream

namespace fse {
struct SymbolTransform { uint32_t bits; int32_t find_state; };          // src/fse.rs:80-84
struct DecodeTransform { uint16_t new_state; uint8_t symbol; uint8_t num_bits; };   // src/fse.rs:260-265

class EncodeTable {                                        // src/fse.rs:72-194
public:
    uint32_t table_log;
    std::vector<uint16_t> table;
    std::array<SymbolTransform, 256> symbol_tt;
    std::vector<uint8_t> symbols;
    explicit EncodeTable(const NormHistogram &hist) : table_log(hist.log2_sum())
    {
        if (table_log < TABLE_LOG_MIN || table_log > TABLE_LOG_MAX) throw Panic("FSE Table must be between 2^9 to 2^16");   // :103-106
        size_t size = (size_t)1 << table_log;
        detail::Dev<int32_t> norm(hist.table().data(), 256), st(1);
        uint32_t l2 = table_log, tl = (uint32_t)hist.table_len();
        detail::Dev<uint32_t> dl2(&l2, 1), dtl(&tl, 1);
        detail::Dev<uint16_t> t(size);
        detail::Dev<fse_b200_symbol_transform> tt(256);
        detail::Dev<uint8_t> sym(size);
        detail::ck(fse_b200_build_encode_tables(detail::ctx(), norm.p, dl2.p, dtl.p, 1, table_log, t.p, tt.p, sym.p, st.p), "build_encode_tables");
        if (st.at(0) < 0) throw Panic("EncodeTable::update");
        table = t.host();
        symbols = sym.host();
        auto h = tt.host();
        for (int i = 0; i < 256; i++) symbol_tt[i] = {h[i].bits, h[i].find_state};
    }
    static size_t compress_bound(size_t size) { return fse_b200_compress_bound(size); }   // :191-193
};

class DecodeTable {                                        // src/fse.rs:253-339
public:
    uint32_t table_log;
    std::vector<DecodeTransform> table;
    explicit DecodeTable(const NormHistogram &hist) : table_log(hist.log2_sum())
    {
        if (table_log < TABLE_LOG_MIN || table_log > TABLE_LOG_MAX) throw Panic("FSE Table must be between 2^9 to 2^16");
        size_t size = (size_t)1 << table_log;
        detail::Dev<int32_t> norm(hist.table().data(), 256), st(1);
        uint32_t l2 = table_log, tl = (uint32_t)hist.table_len();
        detail::Dev<uint32_t> dl2(&l2, 1), dtl(&tl, 1);
        detail::Dev<fse_b200_decode_transform> t(size);
        detail::ck(fse_b200_build_decode_tables(detail::ctx(), norm.p, dl2.p, dtl.p, 1, table_log, t.p, st.p), "build_decode_tables");
        if (st.at(0) < 0) throw Panic("DecodeTable::update");
        auto h = t.host();
        table.resize(size);
        for (size_t i = 0; i < size; i++) table[i] = {h[i].new_state, h[i].symbol, h[i].num_bits};
    }
};

// fse::Encoder / fse::Decoder (src/fse.rs:196-251, :341-386): the per-symbol state machines over tables built on the
// GPU (EncodeTable / DecodeTable above).  Host objects for callers that drive a stream symbol by symbol; "Encoders can
// be interleaved ... each one only mutably borrows the BitStackWriter when encoding a symbol" (src/fse.rs:16-17).
class Encoder {
    uint32_t value_ = 0;
    const EncodeTable *table_;
public:
    explicit Encoder(const EncodeTable &table) : table_(&table) {}                        // :203-205
    // the first symbol to encode (the last to be read) starts from the smallest state and costs no bits (:210-218)
    static Encoder new_first_symbol(const EncodeTable &table, uint8_t first_symbol)
    {
        Encoder e(table);
        const SymbolTransform &tt = table.symbol_tt[first_symbol];
        const uint32_t bits_out = (tt.bits + (1u << 15)) >> 16;
        const uint32_t v = (bits_out << 16) - tt.bits;
        e.value_ = table.table[(size_t)((int32_t)(v >> bits_out) + tt.find_state)];
        return e;
    }
    void encode(bitstream::BitStackWriter &writer, uint8_t sym)                            // :227-245
    {
        const SymbolTransform &tt = table_->symbol_tt[sym];
        const uint32_t bits_out = (tt.bits + value_) >> 16;
        writer.write_bits_unmasked(value_, bits_out);
        value_ = table_->table[(size_t)((int32_t)(value_ >> bits_out) + tt.find_state)];
    }
    void finish(bitstream::BitStackWriter &writer) const { writer.write_bits_unmasked(value_, table_->table_log); }   // :248-250
    uint32_t state() const { return value_; }
};

class Decoder {
    uint16_t state_;
    const DecodeTable *table_;
    Decoder(uint16_t s, const DecodeTable &t) : state_(s), table_(&t) {}
public:
    // None when the reader holds fewer than table_log bits (:349-352)
    static std::optional<Decoder> create(const DecodeTable &table, bitstream::BitStackReader &reader)
    {
        auto s = reader.read(table.table_log);
        if (!s) return std::nullopt;
        return Decoder((uint16_t)*s, table);
    }
    // None when the reader cannot supply num_bits: the end of the stream (:363-380)
    std::optional<uint8_t> decode_symbol(bitstream::BitStackReader &reader)
    {
        const DecodeTransform &e = table_->table[state_];
        auto low = reader.read(e.num_bits);
        if (!low) return std::nullopt;
        state_ = (uint16_t)(e.new_state + *low);
        return e.symbol;
    }
    uint8_t finish() const { return table_->table[state_].symbol; }                       // :383-385
};
}  // namespace fse

namespace detail {
inline size_t compress_one(const std::vector<uint8_t> &src, std::vector<uint8_t> &dst, uint32_t n_states, size_t *header_bytes)
{
    if (src.size() < n_states || src.empty()) throw Panic("called `Option::unwrap()` on a `None` value");   // lib.rs:121,154,156
    fse_b200_params p{(uint32_t)src.size(), 0, n_states, FSE_B200_TABLE_PER_BLOCK, 0, 0};
    size_t cap = fse_b200_compress_blocks_bound(src.size(), &p);
    size_t start = dst.size();
    dst.resize(start + cap);
    uint64_t off[2], total = 0;
    int32_t st = 0;
    int rc = fse_b200_compress_host(ctx(), src.data(), src.size(), &p, dst.data() + start, cap, off, &st, &total);
    if (rc != FSE_B200_OK && rc != FSE_B200_ERR_BLOCK) ck(rc, "compress_host");
    dst.resize(start + total);
    if (st != 0) throw Panic("the reference panics on this input");
    auto hr = NormHistogram::read(dst.data() + start, total);
    if (header_bytes) *header_bytes = hr.second;
    uint8_t last = dst.back();
    return (total - hr.second - 1) * 8 + (32 - __builtin_clz((unsigned)last));   // payload bits incl. the marker
}
inline std::optional<size_t> decompress_one(const std::vector<uint8_t> &src, std::vector<uint8_t> &dst, uint32_t n_states, size_t cap)
{
    if (src.empty()) throw Panic("No bytes provided to read from");           // stream_reader.rs:17 via lib.rs:191,219
    // No length is stored: unless the caller bounds it, the capacity grows geometrically until the stream fits (skewed
    // data expands far beyond 64 x); only the never-terminating stream of quirk Q1 reaches the 1 GiB limit and panics.
    const bool grow = cap == 0;
    if (cap == 0) cap = std::max<size_t>(4096, 64 * src.size());
    Dev<uint8_t> comp(src.data(), src.size());
    uint64_t off[2] = {0, src.size()};
    Dev<uint64_t> doff(off, 2);
    Dev<uint32_t> out_len(1);
    Dev<int32_t> st(1);
    std::unique_ptr<Dev<uint8_t>> outp;
    int32_t rc;
    for (;;) {
        outp.reset(new Dev<uint8_t>(cap));
        fse_b200_params p{(uint32_t)cap, 15, n_states, FSE_B200_TABLE_PER_BLOCK, 0, 0};
        ck(fse_b200_decompress_exhaust(ctx(), comp.p, src.size(), doff.p, 1, &p, outp->p, out_len.p, st.p), "decompress_exhaust");
        rc = st.at(0);
        if (rc == FSE_B200_ERR_CAPACITY && grow && cap < ((size_t)1 << 30)) { cap = std::min<size_t>(cap * 8, (size_t)1 << 30); continue; }
        break;
    }
    Dev<uint8_t> &out = *outp;
    if (rc == FSE_B200_ERR_TABLE_LOG || rc == FSE_B200_ERR_TOO_MANY || rc == FSE_B200_ERR_IO || rc == FSE_B200_ERR_NO_MARKER)
        return std::nullopt;                                                   // .ok()? / BitStackReader::new -> None
    if (rc == FSE_B200_ERR_LENGTH) throw Panic("called `Option::unwrap()` on a `None` value");   // lib.rs:197,224-225
    if (rc == FSE_B200_ERR_CAPACITY) throw Panic("decoder does not terminate within the capacity (reference quirk Q1)");
    if (rc < 0) throw Panic("fse_decompress");
    size_t n = out_len.at(0);
    size_t start = dst.size();
    dst.resize(start + n);
    if (n) cudaMemcpy(dst.data() + start, out.p, n, cudaMemcpyDeviceToHost);
    return n;
}
}  // namespace detail

// src/lib.rs:112-143
inline std::pair<NormHistogram, size_t> fse_compress(const std::vector<uint8_t> &src, std::vector<uint8_t> &dst)
{
    size_t start = dst.size(), hb = 0;
    size_t bits = detail::compress_one(src, dst, 1, &hb);
    return {NormHistogram::read(dst.data() + start, dst.size() - start).first, bits};
}
// src/lib.rs:146-183
inline size_t fse_compress2(const std::vector<uint8_t> &src, std::vector<uint8_t> &dst) { return detail::compress_one(src, dst, 2, nullptr); }
// src/lib.rs:187-211 / :215-248; max_len bounds the reference's non-terminating case (0 = 64 x input)
inline std::optional<size_t> fse_decompress(const std::vector<uint8_t> &src, std::vector<uint8_t> &dst, size_t max_len = 0) { return detail::decompress_one(src, dst, 1, max_len); }
inline std::optional<size_t> fse_decompress2(const std::vector<uint8_t> &src, std::vector<uint8_t> &dst, size_t max_len = 0) { return detail::decompress_one(src, dst, 2, max_len); }

}  // namespace entropy_coders
