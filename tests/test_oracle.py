"""CPU tests: the C oracle against (i) the committed known-answer vectors, (ii) the reference's own
unit tests re-expressed (file:line cited per test), (iii) the independent mechanics model."""
import ctypes as C
import json
import os
import random

import numpy as np
import pytest

import oracle_lib as O
import pymodel as M

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "kat.json")))


def kat_src(k):
    if k["src_hex"] is not None:
        return bytes.fromhex(k["src_hex"])
    return {"KAT-D": lambda: O.generate("geo", 0xC0FFEE01, 1000),
            "KAT-E": lambda: O.generate("text", 0xC0FFEE02, 777),
            "KAT-F": lambda: O.generate("few", 0xC0FFEE03, 513),
            "KAT-G": lambda: O.generate("uniform", 0xC0FFEE03, 2048)}[k["name"]]().tobytes()


@pytest.mark.parametrize("k", KAT["kats"], ids=[k["name"] for k in KAT["kats"]])
def test_kat(k):
    src = kat_src(k)
    assert len(src) == k["src_len"]
    h = O.histogram(src)
    tl = k["table_log_req"] or O.optimal_log2(h)[1]
    rc, nh = O.normalize(h, tl)
    assert rc >= 0
    assert nh.log2 == k["log2"] and nh.table_len == k["table_len"]
    assert list(nh.table)[: nh.table_len] == k["norm"]
    hdr, hbits = O.ncount_write(nh)
    assert hdr.hex() == k["header_hex"] and hbits == k["header_bits"]
    et, dt = O.enc_table(nh), O.dec_table(nh)
    size = 1 << nh.log2
    if "spread" in k:
        assert list(et.symbols)[:size] == k["spread"]
        assert list(et.table)[:size] == k["enc_table"]
        assert [[et.symbol_tt[i].bits, et.symbol_tt[i].find_state] for i in range(nh.table_len)] == k["symbol_tt"]
        assert [[dt.table[i].new_state, dt.table[i].symbol, dt.table[i].num_bits] for i in range(size)] == k["dec_table"]
    for ns, p in k["payload"].items():
        pay, bits = O.encode_payload(et, src, int(ns))
        assert pay.hex() == p["hex"] and bits == p["bits"]
        if k["table_log_req"] == 0:
            comp, hb, _ = O.compress_n(src, 0, int(ns))
            assert comp == hdr + pay and hb == len(hdr)
            assert O.decompress_n_len(comp, int(ns), len(src)) == src


def test_kat_a_hand_derived():
    """SURVEY.md Appendix C KAT-A, derived by hand from histogram.rs:95-155,376-431, fse.rs:101-250."""
    a = KAT["kats"][0]
    assert a["norm"] == [24, 8] and a["header_hex"] == "901f" and a["header_bits"] == 13
    assert a["symbol_tt"] == [[0xFFD0, -24], [0x2FFC0, 16]]
    assert a["payload"]["1"] == {"hex": "4c03", "bits": 10}
    assert a["payload"]["2"] == {"hex": "ac15", "bits": 13}


def test_generator_fixture():
    for i, kind in enumerate(["geo", "text", "few", "uniform"]):
        assert O.generate(kind, 0xC0FFEE00 + i, 32).tobytes().hex() == KAT["generator_first32"][kind]
    # position independence: shards generate their own slice
    full = O.generate("geo", 7, 4096)
    assert np.array_equal(O.generate("geo", 7, 1000, first_index=1234), full[1234:2234])
    # the G_geo LUT is the reference's gen_sequence LUT (lib.rs:255-270)
    lut = np.zeros(65536, np.uint8)
    n = O.lib().fse_or_gen_lut(0, lut.ctypes.data_as(C.c_void_p))
    assert n == 4096 and list(lut[:n]) == M.gen_sequence_lut(0.2)


# ---------------------------------------------------------------- histogram.rs tests

def hist_verify(data, log2):
    """histogram.rs:553-587"""
    h = O.histogram(data)
    rc, nh = O.normalize(h, log2)
    assert rc >= 0
    t = list(nh.table)
    assert sum(abs(x) for x in t) == 1 << nh.log2                       # :566-568
    assert all((h.table[i] == 0) == (t[i] == 0) for i in range(256))    # :569-577
    hdr, bits = O.ncount_write(nh)
    assert len(hdr) <= O.lib().fse_or_write_bound(C.byref(nh))
    rc, back, consumed = O.ncount_read(hdr + b"I am a test")            # :580-586
    assert rc == 0 and consumed == len(hdr)
    assert list(back.table) == t and back.log2 == nh.log2 and back.table_len == nh.table_len
    # mechanics model agrees, including its BitStreamReader tail words
    ph = M.Histogram(bytes(data)).normalize(log2)
    assert ph.table == t
    v = M.Vec()
    assert ph.write(v) == bits and v.bytes() == hdr
    pb, rest = M.NormHistogram.read(hdr + b"I am a test")
    assert rest == b"I am a test" and pb == ph
    return nh


def test_flat_256():
    """histogram.rs:589-593"""
    nh = hist_verify(bytes(range(256)), 9)
    assert nh.log2 == 9


@pytest.mark.parametrize("log2", range(8, 16))
def test_uniform_dist_256(log2):
    """histogram.rs:595-619"""
    data = np.repeat(np.arange(256, dtype=np.uint8), 1 << (log2 - 8))
    h = O.histogram(data)
    assert all(h.table[j] == 1 << (log2 - 8) for j in range(256))       # :607-616
    hist_verify(data.tobytes(), log2)


@pytest.mark.parametrize("log2", range(8, 16))
def test_exp_dist(log2):
    """histogram.rs:621-656"""
    data, remaining, sym = [], 1 << log2, 0
    while True:
        data += [sym] * (remaining >> 1)
        remaining -= remaining >> 1
        sym += 1
        if remaining == 1:
            data.append(sym)
            break
    h = O.histogram(bytes(data))
    for j in range(256):                                                 # :640-653
        exp = ((1 << log2) >> (1 + j)) if j < log2 else (1 if j == log2 else 0)
        assert h.table[j] == exp
    hist_verify(bytes(data), log2)


@pytest.mark.parametrize("log2", range(8, 14))
def test_rand_dist_uniform(log2):
    """histogram.rs:658-670 (seeded here; the reference uses thread_rng)"""
    rng = np.random.default_rng(1000 + log2)
    for _ in range(3):
        hist_verify(rng.integers(0, 256, 1 << (log2 + 2), dtype=np.uint8).tobytes(), log2)


def _mk_hist(counts):
    h = O.Hist()
    for i, c in enumerate(counts):
        h.table[i] = c
    h.size = sum(counts)
    h.table_len = max(i for i, c in enumerate(counts) if c) + 1
    return h


def test_normalize_slow_path_matches_model():
    """histogram.rs:157-261: force the slow path with near-flat histograms; compare with the model."""
    rng = random.Random(5)
    slow = 0
    for trial in range(400):
        nsym = rng.choice([200, 256, 130, 60])
        log2 = rng.choice([8, 9, 10]) if nsym > 128 else rng.choice([7, 8, 9])
        base = rng.choice([3, 8, 30, 250])
        counts = [max(0, base + rng.randint(-base // 2 - 1, base // 2 + 1)) for _ in range(nsym)]
        if trial % 3 == 0:
            counts[rng.randrange(nsym)] += base * rng.choice([5, 40])
        if sum(1 for c in counts if c) < 2:
            continue
        h = _mk_hist(counts + [0] * (256 - nsym))
        ph = M.Histogram(b"")
        ph.table, ph.size, ph.table_len = list(h.table), h.size, h.table_len
        try:
            pn = ph.normalize(log2)
        except ArithmeticError:
            rc, _ = O.normalize(h, log2)
            assert rc < 0
            continue
        rc, nh = O.normalize(h, log2)
        assert rc >= 0
        assert list(nh.table) == pn.table and nh.log2 == pn.log2
        assert (rc == 1) == pn.slow
        slow += pn.slow
        if rc == 1:
            assert all((h.table[i] == 0) == (nh.table[i] == 0) for i in range(256))
    assert slow >= 20, slow


# ---------------------------------------------------------------- bitstream tests

def _enc(test_vec, mark, offset, base):
    """bitstream/mod.rs:29-66 with the mechanics model's writer at a chosen Vec alignment"""
    v = M.Vec(bytes(offset), base=base)
    w = M.BitStackWriter(v)
    total = 0
    for val, bits in test_vec:
        total += bits
        w.write_bits(val, bits)
    if mark:
        w.write_bits(1, 1)
        written = w.finish() - 1
    else:
        written = w.finish()
    assert written == total                                              # :44-47
    assert v.len == (total + (1 if mark else 0) + 7) // 8 + offset       # :52-59
    return v.bytes(), total


def _semantic_bytes(test_vec, mark):
    """bit k of the stream = bit k%8 of byte k/8."""
    acc, n = 0, 0
    for val, bits in test_vec + ([(1, 1)] if mark else []):
        acc |= val << n
        n += bits
    return acc.to_bytes((n + 7) // 8, "little")


@pytest.mark.parametrize("offset", range(8))
def test_stack_offsets(offset):
    """bitstream/mod.rs:112-155: alignment sweep, 1-bit and 1..16-bit fields, read back in reverse."""
    rng = random.Random(offset)
    vecs = [[(i & 1, 1) for i in range(n)] for n in (1, 7, 31, 32, 33, 63, 64, 65, 200, 320)]
    for _ in range(6):
        vecs.append([(rng.getrandbits(b), b) for b in (rng.randint(1, 16) for _ in range(rng.randint(1, 100)))])
    for tv in vecs:
        for base in (0x1000, 0x1001, 0x1003):
            enc, total = _enc(tv, True, offset, base)
            body = enc[offset:]
            assert body == _semantic_bytes(tv, True)
            # C oracle stack reader
            buf = np.frombuffer(body, np.uint8).copy()
            r = O.lib()
            st = C.create_string_buffer(16)
            assert r.fse_or_bitstack_init(st, buf.ctypes.data_as(C.c_void_p), buf.size) == 0
            for val, bits in reversed(tv):                               # :68-84
                out = C.c_uint32()
                assert r.fse_or_bitstack_read(st, bits, C.byref(out)) == 0 and out.value == val
            out = C.c_uint32()
            assert r.fse_or_bitstack_read(st, 1, C.byref(out)) < 0       # fully consumed (:85-90)
            # mechanics reader at several slice alignments
            for rbase in (0x2000 + offset, 0x2001, 0x2002, 0x2003):
                d = M.BitStackReader(body, base=rbase)
                for val, bits in reversed(tv):
                    assert d.read(bits) == val
                assert d.available() == 0 and d.finish()


@pytest.mark.parametrize("offset", [0, 1, 5])
def test_stream_offsets(offset):
    """bitstream/mod.rs:167-214: forward reader; finish() leaves <= 1 byte and 0 bits (:106-109)."""
    rng = random.Random(100 + offset)
    vecs = [[(i & 1, 1) for i in range(n)] for n in (1, 8, 63, 64, 65, 128)]
    for _ in range(8):
        vecs.append([(rng.getrandbits(b), b) for b in (rng.randint(1, 16) for _ in range(rng.randint(1, 100)))])
    for tv in vecs:
        enc, total = _enc(tv, False, offset, 0x1000)
        body = enc[offset:]
        assert body == _semantic_bytes(tv, False)
        d = M.BitStreamReader(body, total)
        buf = np.frombuffer(body, np.uint8).copy()
        st = C.create_string_buffer(32)
        assert O.lib().fse_or_bitstream_init(st, buf.ctypes.data_as(C.c_void_p), buf.size, total) == 0
        for val, bits in tv:
            assert d.read(bits) == val
            out = C.c_uint32()
            assert O.lib().fse_or_bitstream_read(st, bits, C.byref(out)) == 0 and out.value == val
        rest, left, off = d.finish()
        assert len(rest) <= 1 and left == 0 and off <= 8
        out = C.c_uint32()
        assert O.lib().fse_or_bitstream_read(st, 1, C.byref(out)) == O.lib().fse_or_bitstream_peek(st, 1, C.byref(out)) < 0


def test_stack_reader_rejects_missing_marker():
    """stack_reader.rs:18-20, :77-83"""
    st = C.create_string_buffer(16)
    z = np.zeros(4, np.uint8)
    assert O.lib().fse_or_bitstack_init(st, z.ctypes.data_as(C.c_void_p), 0) < 0
    assert O.lib().fse_or_bitstack_init(st, z.ctypes.data_as(C.c_void_p), 4) < 0
    with pytest.raises(ValueError):
        M.BitStackReader(b"\x01\x02\x00")


# ---------------------------------------------------------------- codec round trips

@pytest.mark.parametrize("n", [1 << 16, (1 << 16) - 1, 1 << 15, 4099])
def test_compress_roundtrip_reference_shapes(n):
    """lib.rs:280-302 (64 KiB) and fse.rs:461-506 (32 KiB), gen_sequence(0.2); plus odd lengths."""
    src = O.generate("geo", 0xC0FFEE01, n).tobytes()
    for ns in (1, 2):
        comp, hb, pbits = O.compress_n(src, 0, ns)
        assert len(comp) == hb + (pbits + 7) // 8
        assert O.decompress_n_exhaust(comp, ns, n + 64) == src   # reference semantics (exhaustion)
        assert O.decompress_n_len(comp, ns, n) == src            # length-driven (GPU semantics)
    assert O.ref_compress2(src) == O.compress_n(src, 0, 2)[0]    # reference loop structure == semantics
    assert O.ref_decompress2(comp, n + 64) == src


def test_mechanics_model_matches_oracle_bytes():
    """fse_compress / fse_compress2 driven through the simulated 64-bit accumulator and Vec."""
    rng = random.Random(11)
    for trial in range(40):
        kind = ["geo", "text", "few", "uniform"][trial % 4]
        n = rng.choice([5, 6, 7, 8, 9, 31, 100, 257, 1000, 2049, 5000])
        src = O.generate(kind, trial, n).tobytes()
        try:
            c1 = O.compress_n(src, 0, 1)[0]
            c2 = O.compress_n(src, 0, 2)[0]
        except ValueError:
            with pytest.raises((ArithmeticError, AssertionError, IndexError)):
                M.fse_compress(src, M.Vec())
            continue
        for base in (0x1000, 0x1001, 0x1002, 0x1003):
            pre = bytes(rng.randrange(256) for _ in range(rng.randrange(0, 9)))
            v = M.Vec(pre, base=base)
            M.fse_compress(src, v)
            assert v.bytes() == pre + c1
            v = M.Vec(pre, base=base)
            M.fse_compress2(src, v)
            assert v.bytes() == pre + c2
        for base in (0x2000, 0x2001, 0x2002, 0x2003):
            d1 = M.fse_decompress(c1, limit=n + 64, base=base)
            d2 = M.fse_decompress2(c2, limit=n + 64, base=base)
            assert d1 == O.decompress_n_exhaust(c1, 1, n + 64)
            assert d2 == O.decompress_n_exhaust(c2, 2, n + 64)
            assert d1[:n] == src and d2[:n] == src


@pytest.mark.parametrize("ns", [1, 2, 3, 4, 7, 32, 64])
def test_n_state_composition_roundtrip(ns):
    """general N; every residue of n mod N; payload size accounting."""
    for kind in ("geo", "text", "few", "uniform"):
        for n in [ns, ns + 1, 2 * ns - 1, 2 * ns, 2 * ns + 1, 1000, 65536 - 7]:
            if n < 5:
                continue
            src = O.generate(kind, 99 + n, n).tobytes()
            if len(set(src)) < 2:
                continue
            comp, hb, pbits = O.compress_n(src, 0, ns)
            assert O.decompress_n_len(comp, ns, n) == src
            nh = O.ncount_read(comp)[1]
            assert pbits >= ns * nh.log2 + 1


def test_block_driver_threads_and_ref2():
    src = O.generate("text", 0xC0FFEE02, 10 * 4096 + 123)
    a = O.compress_blocks(src, 4096, n_states=2, threads=1)
    b = O.compress_blocks(src, 4096, n_states=2, threads=4, use_ref2=1)
    assert np.array_equal(a[1], b[1]) and not a[2].any() and not b[2].any()
    for i, s in enumerate(a[1]):
        assert np.array_equal(a[0][i, : int(s)], b[0][i, : int(s)])
    out, st = O.decompress_blocks(a[0], a[1], src.size, 4096, n_states=2, threads=3)
    assert not st.any() and np.array_equal(out, src)
    out, st = O.decompress_blocks(b[0], b[1], src.size, 4096, n_states=2, threads=2, use_ref2=1)
    assert not st.any() and np.array_equal(out, src)


# ---------------------------------------------------------------- quirks

def test_q1_single_symbol_block_never_terminates_in_reference():
    src = bytes([7]) * 100
    comp, hb, pbits = O.compress_n(src, 0, 1)
    assert pbits == O.ncount_read(comp)[1].log2 + 1          # tl state bits + marker only
    with pytest.raises(ValueError):                          # exhaustion decode hits the capacity guard
        O.decompress_n_exhaust(comp, 1, 4096)
    with pytest.raises(OverflowError):
        M.fse_decompress(comp, limit=4096)
    assert O.decompress_n_len(comp, 1, 100) == src           # length-driven decode is fine


def test_q2_degenerate_inputs_are_errors():
    for src in (b"", b"\x00", b"\x00" * 50, b"\x05", b"ab", b"abcd"):
        with pytest.raises(ValueError):
            O.compress_n(src, 0, 1)
    with pytest.raises(ValueError):
        O.compress_n(b"abcdefgh", 0, 9)                      # fewer symbols than states


def test_q3_symbol_count_counts_zero_entries():
    nh = O.normalize(O.histogram(bytes([0, 0, 0, 0, 0, 0, 1, 1])), 5)[1]
    assert O.lib().fse_or_symbol_count(nh.table) == 254


def test_q7_normalize_raises_requested_log2():
    nh = O.normalize(O.histogram(bytes(range(256)) * 4), 5)[1]
    assert nh.log2 == 9


def test_header_errors():
    """histogram.rs:439-441 TableLogTooLarge; :498-500 TooManySymbols; Io on truncation."""
    assert O.ncount_read(bytes([0x0F, 0, 0, 0]))[0] == -3
    hdr, _ = O.ncount_write(O.normalize(O.histogram(O.generate("text", 1, 5000)), 11)[1])
    assert O.ncount_read(hdr[: len(hdr) // 2])[0] in (-5, -4)
    with pytest.raises((ValueError, EOFError)):
        M.NormHistogram.read(hdr[: len(hdr) // 2])


def test_compress_bound_and_step():
    """fse.rs:68-70, :191-193"""
    assert O.lib().fse_or_compress_bound(65536) == 66572 and O.lib().fse_or_compress_bound(131072) == 132620
    assert O.lib().fse_or_table_step(2048) == 1283
    for kind in ("uniform", "geo"):
        src = O.generate(kind, 3, 65536)
        comp, _, _ = O.compress_n(src, 0, 2)
        assert len(comp) <= 66572
