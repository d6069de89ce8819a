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
