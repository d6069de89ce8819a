// The reference crate's own tests (src/lib.rs:280-302, src/histogram.rs:553-670, src/fse.rs:461-506)
// re-expressed on the C++ mirror of its API (include/entropy_coders.hpp), with the oracle as the checker.
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <vector>

#include "entropy_coders.hpp"
extern "C" {
#include "fse_oracle.h"
}

using namespace entropy_coders;

#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #c); std::exit(1); } } while (0)

static std::vector<uint8_t> gen_sequence(size_t n, uint64_t seed)   // lib.rs:255-278 with the seeded generator of SURVEY 8(d)
{
    std::vector<uint8_t> v(n);
    fse_or_generate(FSE_OR_GEN_GEO, seed, 0, v.data(), n);
    return v;
}

static std::vector<uint8_t> oracle_stream(const std::vector<uint8_t> &src, unsigned n_states, size_t *pbits)
{
    std::vector<uint8_t> out(fse_or_compress_bound(src.size()) + 64);
    size_t hb = 0;
    long n = fse_or_compress_n(src.data(), src.size(), 0, n_states, out.data(), out.size(), &hb, pbits);
    CHECK(n > 0);
    out.resize((size_t)n);
    return out;
}

static void hist_verify(const std::vector<uint8_t> &data, uint32_t log2)   // histogram.rs:553-587
{
    Histogram hist(data);
    auto hist_table = hist.table();
    NormHistogram norm = hist.normalize(log2);
    int64_t sum = 0;
    for (auto x : norm.table()) sum += x < 0 ? -x : x;
    CHECK(sum == (int64_t)1 << norm.log2_sum());                             // :566-568
    for (int i = 0; i < 256; i++) CHECK((hist_table[i] == 0) == (norm.table()[i] == 0));   // :569-577
    std::vector<uint8_t> enc;
    enc.reserve(norm.write_bound());
    norm.write(enc);
    CHECK(enc.size() <= norm.write_bound());
    const char *test = "I am a test";
    enc.insert(enc.end(), test, test + 11);
    auto r = NormHistogram::read(enc.data(), enc.size());                     // :580-586
    CHECK(enc.size() - r.second == 11 && std::memcmp(enc.data() + r.second, test, 11) == 0);
    CHECK(r.first == norm);
    // and the oracle agrees on every count
    fse_or_hist oh;
    fse_or_norm on;
    fse_or_histogram(data.data(), data.size(), &oh);
    CHECK(fse_or_normalize(&oh, log2, &on) >= 0);
    CHECK(on.log2 == norm.log2_sum() && on.table_len == norm.table_len());
    for (int i = 0; i < 256; i++) CHECK(on.table[i] == norm.table()[i]);
}


// bitstream/mod.rs:26-110: encode a list of (value, bits) fields behind `offset` existing bytes, decode as a stack and as a stream
static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint32_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (uint32_t)(rng_state >> 16); }

static void bit_roundtrip(const std::vector<std::pair<size_t, size_t>> &fields, bool mark, size_t offset)
{
    std::vector<uint8_t> encoded(offset, 0xA5);
    bitstream::BitStackWriter enc(encoded);
    size_t total_bits = 0;
    for (auto &f : fields) { total_bits += f.second; enc.write_bits(f.first, f.second); }
    size_t written = mark ? (enc.write_bits(1, 1), enc.finish() - 1) : enc.finish();
    CHECK(written == total_bits);                                               // mod.rs:44-47
    CHECK(encoded.size() == (total_bits + (mark ? 1 : 0) + 7) / 8 + offset);     // mod.rs:52-59
    // the oracle's writer (same reference lines, restated in C) produces the same bytes
    std::vector<uint8_t> ob(offset + total_bits / 8 + 16, 0);
    fse_or_bitw w;
    fse_or_bitw_init(&w, ob.data(), ob.size(), offset);
    for (auto &f : fields) fse_or_bitw_put(&w, f.first, (unsigned)f.second);
    if (mark) fse_or_bitw_put(&w, 1, 1);
    size_t new_len = 0;
    size_t obits = fse_or_bitw_finish(&w, &new_len);
    CHECK(obits == total_bits + (mark ? 1 : 0) && new_len == encoded.size());
    CHECK(std::memcmp(ob.data() + offset, encoded.data() + offset, new_len - offset) == 0);
    if (mark) {                                                                 // mod.rs:68-91
        auto dec = bitstream::BitStackReader::create(encoded.data() + offset, encoded.size() - offset);
        CHECK(dec.has_value());
        for (size_t i = fields.size(); i-- > 0;) {
            auto v = dec->read(fields[i].second);
            CHECK(v && *v == fields[i].first);
        }
        CHECK(dec->available() == 0 && dec->finish());
        CHECK(!dec->read(1));
    } else if (total_bits) {                                                    // mod.rs:93-110
        bitstream::BitStreamReader dec(encoded.data() + offset, encoded.size() - offset, total_bits);
        for (auto &f : fields) CHECK(dec.read(f.second) == f.first);
        auto rest = dec.finish();
        CHECK(rest.len <= 1 && rest.bits_left == 0 && rest.bit_offset <= 8);
        bool eof = false;
        try { dec.read(1); } catch (const bitstream::UnexpectedEof &) { eof = true; }
        CHECK(eof);
    }
}

// lib.rs:112-143 / :146-183 written with the crate's own per-symbol objects
static std::vector<uint8_t> host_compress(const std::vector<uint8_t> &src, unsigned n_states, size_t *pbits)
{
    std::vector<uint8_t> dst;
    NormHistogram hist = NormHistogram::create(src);
    hist.write(dst);
    fse::EncodeTable table(hist);
    bitstream::BitStackWriter writer(dst);
    std::vector<fse::Encoder> enc;
    const size_t n = src.size();
    for (unsigned j = 0; j < n_states; j++) enc.push_back(fse::Encoder(table));
    // state j owns the indices congruent to j; the top n_states symbols initialise their states at no cost
    for (size_t i = n; i-- > n - n_states;) enc[i % n_states] = fse::Encoder::new_first_symbol(table, src[i]);
    for (size_t i = n - n_states; i-- > 0;) enc[i % n_states].encode(writer, src[i]);
    for (unsigned j = n_states; j-- > 0;) enc[j].finish(writer);
    writer.write_bits(1, 1);
    *pbits = writer.finish();
    return dst;
}
static std::vector<uint8_t> host_decompress(const std::vector<uint8_t> &src, unsigned n_states)
{
    auto hr = NormHistogram::read(src.data(), src.size());
    auto reader = bitstream::BitStackReader::create(src.data() + hr.second, src.size() - hr.second);
    CHECK(reader.has_value());
    fse::DecodeTable table(hr.first);
    std::vector<fse::Decoder> dec;
    for (unsigned j = 0; j < n_states; j++) { auto d = fse::Decoder::create(table, *reader); CHECK(d.has_value()); dec.push_back(*d); }
    std::vector<uint8_t> out;
    for (size_t i = 0;; i++) {
        auto s = dec[i % n_states].decode_symbol(*reader);
        if (!s) {                                                               // lib.rs:236-243: flush the final states in index order
            for (unsigned k = 0; k < n_states; k++) out.push_back(dec[(i + k) % n_states].finish());
            break;
        }
        out.push_back(*s);
    }
    return out;
}

int main()
{
    // lib.rs:280-290 `compress`
    {
        auto src = gen_sequence(1 << 16, 0xC0FFEE01);
        std::vector<uint8_t> dst, dec;
        auto r = fse_compress(src, dst);
        size_t pbits = 0;
        auto exp = oracle_stream(src, 1, &pbits);
        CHECK(dst == exp && r.second == pbits);
        auto n = fse_decompress(dst, dec);
        CHECK(n && *n == src.size() && dec == src);
    }
    // lib.rs:292-302 `compress2`, appending behind existing bytes
    {
        auto src = gen_sequence((1 << 16) - 1, 0xC0FFEE02);
        std::vector<uint8_t> dst = {1, 2, 3}, dec = {9};
        size_t bits = fse_compress2(src, dst);
        size_t pbits = 0;
        auto exp = oracle_stream(src, 2, &pbits);
        CHECK(bits == pbits && dst.size() == exp.size() + 3 && std::memcmp(dst.data() + 3, exp.data(), exp.size()) == 0);
        std::vector<uint8_t> stream(dst.begin() + 3, dst.end());
        auto n = fse_decompress2(stream, dec);
        CHECK(n && *n == src.size() && dec.size() == src.size() + 1 && std::memcmp(dec.data() + 1, src.data(), src.size()) == 0);
    }
    // histogram.rs:589-593 flat_256, :595-619 uniform_dist_256, :621-656 exp_dist
    {
        std::vector<uint8_t> flat(256);
        for (int i = 0; i < 256; i++) flat[i] = (uint8_t)i;
        NormHistogram nh = NormHistogram::create(flat);
        CHECK(nh.log2_sum() == 9 && nh.symbol_count() == 0);
        for (uint32_t log2 : {8u, 11u, 15u}) {
            std::vector<uint8_t> data;
            for (int x = 0; x < 256; x++) data.insert(data.end(), (size_t)1 << (log2 - 8), (uint8_t)x);
            Histogram h(data);
            for (auto c : h.table()) CHECK(c == 1u << (log2 - 8));          // :607-616
            hist_verify(data, log2);
            std::vector<uint8_t> e;
            size_t remaining = (size_t)1 << log2;
            uint8_t sym = 0;
            for (;;) {
                e.insert(e.end(), remaining >> 1, sym);
                remaining -= remaining >> 1;
                sym++;
                if (remaining == 1) { e.push_back(sym); break; }
            }
            Histogram he(e);
            for (uint32_t j = 0; j < 256; j++)                              // :640-653
                CHECK(he.table()[j] == (j < log2 ? ((1u << log2) >> (1 + j)) : (j == log2 ? 1u : 0u)));
            hist_verify(e, log2);
        }
    }
    // tables against the oracle (fse.rs:101-189, :280-338)
    {
        auto src = gen_sequence(5000, 7);
        NormHistogram nh = NormHistogram::create(src);
        fse::EncodeTable et(nh);
        fse::DecodeTable dt(nh);
        fse_or_norm on;
        CHECK(fse_or_norm_new(src.data(), src.size(), &on) >= 0);
        static fse_or_enc_table oe;
        static fse_or_dec_table od;
        CHECK(fse_or_enc_table_build(&on, &oe) == 0 && fse_or_dec_table_build(&on, &od) == 0);
        size_t size = (size_t)1 << nh.log2_sum();
        CHECK(et.table.size() == size && dt.table.size() == size);
        for (size_t i = 0; i < size; i++) {
            CHECK(et.table[i] == oe.table[i] && et.symbols[i] == oe.symbols[i]);
            CHECK(dt.table[i].new_state == od.table[i].new_state && dt.table[i].symbol == od.table[i].symbol && dt.table[i].num_bits == od.table[i].num_bits);
        }
        for (int i = 0; i < 256; i++) CHECK(et.symbol_tt[i].bits == oe.symbol_tt[i].bits && et.symbol_tt[i].find_state == oe.symbol_tt[i].find_state);
        CHECK(fse::EncodeTable::compress_bound(65536) == 66572);
    }
    // bitstream/mod.rs:112-224 stack_tests / stream_tests: growing field lists, 1-bit and 1..16-bit fields, Vec offsets 0..7
    {
        for (size_t offset = 0; offset < 8; offset++) {
            for (size_t len : {1u, 2u, 7u, 8u, 9u, 31u, 32u, 33u, 63u, 64u, 65u, 200u, 1000u}) {
                std::vector<std::pair<size_t, size_t>> ones, mixed;
                for (size_t i = 0; i < len; i++) {
                    ones.push_back({rnd() & 1u, 1});
                    size_t bits = 1 + rnd() % 16;
                    mixed.push_back({rnd() & ((1u << bits) - 1), bits});
                }
                bit_roundtrip(ones, true, offset);
                bit_roundtrip(ones, false, offset);
                bit_roundtrip(mixed, true, offset);
                bit_roundtrip(mixed, false, offset);
            }
        }
        CHECK(!bitstream::BitStackReader::create(nullptr, 0));                   // stack_reader.rs:18-20
        uint8_t zero[2] = {0xff, 0};
        CHECK(!bitstream::BitStackReader::create(zero, 2));                      // :77-83: the last byte must hold the marker
    }
    // fse::Encoder / fse::Decoder on GPU-built tables: the crate's 1- and 2-state drivers (lib.rs:112-248), and wider
    // compositions (fse.rs:16-17), produce the oracle's bytes and the bytes of the GPU block path
    {
        for (unsigned n_states : {1u, 2u, 4u}) {
            auto src = gen_sequence(20000 + n_states, 0xC0FFEE10 + n_states);
            size_t pbits = 0, obits = 0;
            auto mine = host_compress(src, n_states, &pbits);
            auto exp = oracle_stream(src, n_states, &obits);
            CHECK(mine == exp && pbits == obits);
            CHECK(host_decompress(mine, n_states) == src);
            if (n_states <= 2) {
                std::vector<uint8_t> gpu;
                if (n_states == 1) fse_compress(src, gpu); else fse_compress2(src, gpu);
                CHECK(gpu == mine);
            }
        }
    }
    // error behaviour: None on a bad header (lib.rs:219), HistError from read (histogram.rs:439-441), panic on empty input (lib.rs:154)
    {
        std::vector<uint8_t> bad = {0x0F, 0, 0, 0}, out;
        CHECK(!fse_decompress2(bad, out));
        bool threw = false;
        try { NormHistogram::read(bad.data(), bad.size()); } catch (const HistError &e) { threw = e.kind == HistError::TableLogTooLarge; }
        CHECK(threw);
        threw = false;
        try { std::vector<uint8_t> empty, d; fse_compress2(empty, d); } catch (const Panic &) { threw = true; }
        CHECK(threw);
    }
    std::printf("cpp crate-API tests ok\n");
    return 0;
}
