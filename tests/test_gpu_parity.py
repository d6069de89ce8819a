"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle, the committed
known-answer vectors and size-independent properties.  Everything here is bit-exact."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "kat.json")))


@pytest.fixture(scope="module")
def ctx():
    import entropy_coders_b200 as E
    c = E.Context(0)
    yield c
    c.close()


def dev(ctx, arr):
    import torch
    return torch.from_numpy(np.ascontiguousarray(arr)).to(ctx.device)


def oracle_blocks(src, block_size, table_log, n_states):
    """list of per-block oracle streams (None where the reference panics)"""
    out = []
    for o in range(0, len(src), block_size):
        blk = src[o:o + block_size]
        try:
            out.append(O.compress_n(blk, table_log, n_states)[0])
        except ValueError:
            out.append(None)
    return out


def gpu_blocks(ctx, src, block_size, table_log, n_states):
    d, off, st, total = ctx.compress_blocks(dev(ctx, src), block_size, table_log, n_states)
    off = off.cpu().numpy().astype(np.int64)
    buf = d[:total].cpu().numpy().tobytes()
    assert off[0] == 0 and off[-1] == total
    return [buf[off[i]:off[i + 1]] for i in range(len(off) - 1)], st.cpu().numpy(), (d, off, total)


# ------------------------------------------------------------------------------------ stages

@pytest.mark.parametrize("kind", ["geo", "text", "few", "uniform"])
@pytest.mark.parametrize("block_size", [65536, 4096, 1000, 131072])
def test_histogram_blocks(ctx, kind, block_size):
    """Histogram::new, src/histogram.rs:18-66 (counts and table_len), ragged tail included"""
    n = 5 * block_size + 777
    src = O.generate(kind, 42, n)
    counts, tlen = ctx.histogram_blocks(dev(ctx, src), block_size)
    counts = counts.cpu().numpy().view(np.uint32)
    tlen = tlen.cpu().numpy()
    for b in range(counts.shape[0]):
        blk = src[b * block_size:(b + 1) * block_size]
        exp = np.bincount(blk, minlength=256)
        assert np.array_equal(counts[b], exp)
        assert tlen[b] == int(np.max(np.nonzero(exp)[0])) + 1


def test_histogram_unaligned_source(ctx):
    src = O.generate("geo", 7, 300000)
    for shift in (1, 3, 5, 15):
        counts, _ = ctx.histogram_blocks(dev(ctx, src)[shift:], 50001)
        counts = counts.cpu().numpy().view(np.uint32)
        s = src[shift:]
        for b in range(counts.shape[0]):
            assert np.array_equal(counts[b], np.bincount(s[b * 50001:(b + 1) * 50001], minlength=256))


def test_histogram_global(ctx):
    src = O.generate("text", 9, 3 * (1 << 20) + 12345)
    c = ctx.histogram_global(dev(ctx, src)).cpu().numpy()
    assert np.array_equal(c, np.bincount(src, minlength=256))


def _rand_hists(rng, count):
    hs = []
    for t in range(count):
        mode = t % 6
        nsym = int(rng.choice([2, 3, 17, 60, 130, 200, 256]))
        if mode == 0:      # near flat: exercises normalize_slow
            base = int(rng.choice([3, 8, 30, 250]))
            c = np.maximum(0, base + rng.integers(-base // 2 - 1, base // 2 + 2, nsym))
        elif mode == 1:    # geometric
            c = (rng.integers(1, 5000) * (0.5 + 0.5 * rng.random()) ** np.arange(nsym)).astype(np.int64)
        elif mode == 2:    # sparse with holes (zero runs in the header)
            c = rng.integers(0, 50, nsym) * (rng.random(nsym) < 0.3)
        elif mode == 3:    # one dominant symbol
            c = rng.integers(0, 4, nsym)
            c[rng.integers(nsym)] += 100000
        elif mode == 4:    # single symbol (not symbol 0)
            c = np.zeros(nsym, dtype=np.int64)
            c[nsym - 1] = rng.integers(5, 1000)
        else:
            c = rng.integers(0, 2000, nsym)
        full = np.zeros(256, dtype=np.int64)
        full[:nsym] = c
        if np.count_nonzero(full) == 0:
            full[1] = 7
        hs.append(full)
    return np.stack(hs)


@pytest.mark.parametrize("table_log", [0, 5, 9, 11, 12, 15])
def test_normalize_header_tables_vs_oracle(ctx, table_log):
    """Histogram::normalize (+slow path), NormHistogram::write/read, EncodeTable::update,
    DecodeTable::update against the oracle for a few hundred histograms"""
    import torch
    rng = np.random.default_rng(100 + table_log)
    H = _rand_hists(rng, 240)
    norm, log2, tlen, st = ctx.normalize(dev(ctx, H), table_log)
    normh, log2h, tlenh, sth = norm.cpu().numpy(), log2.cpu().numpy(), tlen.cpu().numpy(), st.cpu().numpy()
    keep, slow = [], 0
    for t in range(H.shape[0]):
        h = O.Hist()
        for i in range(256):
            h.table[i] = int(H[t, i])
        h.size = int(H[t].sum())
        h.table_len = int(np.max(np.nonzero(H[t])[0])) + 1
        if table_log == 0:
            rc, tl = O.optimal_log2(h)
            if rc < 0:
                assert sth[t] < 0
                continue
        else:
            tl = table_log
        rc, nh = O.normalize(h, tl)
        if rc < 0:                       # the reference panics here (oracle -1 <-> FSE_B200_ERR_PANIC)
            assert sth[t] == -8, (t, sth[t], rc)
            continue
        assert sth[t] == rc, (t, sth[t], rc)
        slow += rc
        assert log2h[t] == nh.log2 and tlenh[t] == nh.table_len
        assert np.array_equal(normh[t], np.array(nh.table, dtype=np.int32))
        keep.append((t, nh))
    assert len(keep) > 150
    if table_log == 9:
        assert slow > 0                  # normalize_slow (histogram.rs:157-261) was exercised
    idx = torch.tensor([t for t, _ in keep], device=ctx.device)
    n2, l2, t2 = norm[idx].contiguous(), log2[idx].contiguous(), tlen[idx].contiguous()
    # header write
    rows, nbytes, nbits = ctx.ncount_write(n2, l2, t2)
    rows, nbytes_h, nbits_h = rows.cpu().numpy(), nbytes.cpu().numpy(), nbits.cpu().numpy()
    for k, (t, nh) in enumerate(keep):
        hdr, bits = O.ncount_write(nh)
        assert nbits_h[k] == bits and nbytes_h[k] == len(hdr)
        assert rows[k, :len(hdr)].tobytes() == hdr
    # header read (with trailing bytes, like hist_verify histogram.rs:580-586)
    rows2 = rows.copy()
    for k in range(len(keep)):
        rows2[k, nbytes_h[k]:nbytes_h[k] + 11] = np.frombuffer(b"I am a test", np.uint8)
    rn, rl, rt, rc_, rs = ctx.ncount_read(dev(ctx, rows2), dev(ctx, (nbytes_h + 11).astype(np.int32)))
    assert not rs.cpu().numpy().any()
    assert np.array_equal(rn.cpu().numpy(), n2.cpu().numpy())
    assert np.array_equal(rl.cpu().numpy(), l2.cpu().numpy()) and np.array_equal(rt.cpu().numpy(), t2.cpu().numpy())
    assert np.array_equal(rc_.cpu().numpy(), nbytes_h)
    # tables
    maxl = int(l2.max().item())
    table, tt, sym, est = ctx.build_encode_tables(n2, l2, t2, maxl)
    dtab, dst_ = ctx.build_decode_tables(n2, l2, t2, maxl)
    assert not est.cpu().numpy().any() and not dst_.cpu().numpy().any()
    table, tt, sym, dtab = (table.cpu().numpy().view(np.uint16), tt.cpu().numpy(), sym.cpu().numpy(),
                            dtab.cpu().numpy().view(np.uint32))
    for k, (t, nh) in enumerate(keep[:80]):
        size = 1 << nh.log2
        et, dt = O.enc_table(nh), O.dec_table(nh)
        assert np.array_equal(sym[k, :size], np.frombuffer(bytes(et.symbols)[:size], np.uint8))
        assert np.array_equal(table[k, :size], np.array(et.table[:size], dtype=np.uint16))
        exp_tt = np.array([[et.symbol_tt[i].bits, et.symbol_tt[i].find_state & 0xFFFFFFFF] for i in range(256)], dtype=np.uint32)
        assert np.array_equal(tt[k].view(np.uint32), exp_tt)
        exp_d = np.array([dt.table[i].new_state | (dt.table[i].symbol << 16) | (dt.table[i].num_bits << 24)
                          for i in range(size)], dtype=np.uint32)
        assert np.array_equal(dtab[k, :size], exp_d)


def test_header_read_errors(ctx):
    """histogram.rs:439-441 TableLogTooLarge, :498-500 TooManySymbols, Io on truncation"""
    nh = O.normalize(O.histogram(O.generate("text", 1, 5000)), 11)[1]
    hdr, _ = O.ncount_write(nh)
    cases = [bytes([0x0F, 0, 0, 0]), hdr[:len(hdr) // 2], hdr[:1], b"\x00" * 40]
    rows = np.zeros((len(cases), 512), np.uint8)
    lens = np.zeros(len(cases), np.int32)
    for i, c in enumerate(cases):
        rows[i, :len(c)] = np.frombuffer(c, np.uint8)
        lens[i] = len(c)
    st = ctx.ncount_read(dev(ctx, rows), dev(ctx, lens))[4].cpu().numpy()
    for i, c in enumerate(cases):
        assert st[i] == O.ncount_read(c)[0], i


@pytest.mark.parametrize("n_states", [32, 64, 128])
def test_header_errors_inside_the_decode_kernels(ctx, n_states):
    """the decode kernels parse the header themselves (from a shared-memory copy for 64 / 128 states): a block whose
    header is malformed reports exactly the oracle's NormHistogram::read error, the other blocks still decode"""
    bs = 4096
    src = O.generate("text", 5, bs * 3)
    good = oracle_blocks(src, bs, 0, n_states)
    hdr, _ = O.ncount_write(O.normalize(O.histogram(src[:bs]), 11)[1])
    rng = np.random.default_rng(9)
    bad_blocks = [hdr[:len(hdr) // 2],                                   # truncated -> Io
                  bytes([0x0B]) + bytes(30),                             # table_log 16 -> TableLogTooLarge
                  bytes([0x06]) + bytes([0xFF]) * 700,                   # long garbage
                  bytes([0x0A]) + rng.integers(0, 256, 600, dtype=np.uint8).tobytes(),
                  bytes([0x00]) * 40]
    for bad in bad_blocks:
        exp = O.ncount_read(bad)[0]
        streams = [good[0], bad, good[2]]
        off = np.zeros(4, np.int64)
        off[1:] = np.cumsum([len(x) for x in streams])
        comp = np.frombuffer(b"".join(streams), np.uint8).copy()
        out, st = ctx.decompress_blocks(dev(ctx, comp), comp.size, dev(ctx, off), src.size, bs, 0, n_states)
        st = st.cpu().numpy()
        assert st[0] == 0 and st[2] == 0, st
        if exp < 0:
            assert st[1] == exp, (bad[:4], st[1], exp)
        else:
            assert st[1] < 0, (bad[:4], st[1])                           # a header parsed, the payload cannot fit
        out = out.cpu().numpy()
        assert np.array_equal(out[:bs], src[:bs]) and np.array_equal(out[2 * bs:], src[2 * bs:])


# ------------------------------------------------------------------------------------ encoded bytes

@pytest.mark.parametrize("k", KAT["kats"], ids=[k["name"] for k in KAT["kats"]])
def test_kat_bytes(ctx, k):
    """committed known-answer vectors (tests/golden/kat.json), every recorded state count"""
    from test_oracle import kat_src
    src = np.frombuffer(kat_src(k), np.uint8)
    for ns, p in k["payload"].items():
        if int(ns) not in (1, 2, 4, 32, 64, 128):
            continue
        blocks, st, _ = gpu_blocks(ctx, src, len(src), k["table_log_req"], int(ns))
        assert st[0] == 0
        assert blocks[0].hex() == k["header_hex"] + p["hex"]


@pytest.mark.parametrize("kind", ["geo", "text", "few", "uniform"])
@pytest.mark.parametrize("n_states", [1, 2, 4, 32, 64, 128])
def test_compress_blocks_bit_exact(ctx, kind, n_states):
    """every block's bytes equal the oracle's fse_compress(N)(block); decode round-trips"""
    block_size = 65536 if n_states >= 32 else 8192
    n = 9 * block_size + 4321
    src = O.generate(kind, 0xC0FFEE00 + n_states, n)
    blocks, st, (d, off, total) = gpu_blocks(ctx, src, block_size, 0, n_states)
    exp = oracle_blocks(src, block_size, 0, n_states)
    assert len(blocks) == len(exp)
    for b, (g, e) in enumerate(zip(blocks, exp)):
        assert st[b] == 0 and e is not None
        assert g == e, "block %d differs" % b
    out, dst_ = ctx.decompress_blocks(d, total, dev(ctx, off), n, block_size, 0, n_states)
    assert not dst_.cpu().numpy().any()
    assert np.array_equal(out.cpu().numpy(), src)


@pytest.mark.parametrize("n_states", [32, 64, 128])
@pytest.mark.parametrize("table_log", [9, 11, 12, 13])
@pytest.mark.parametrize("kind", ["few", "uniform", "text"])
def test_table_log_sweep(ctx, kind, table_log, n_states):
    """BASELINE config 3: explicit table_log 9/11/12 (Histogram::normalize(tl), histogram.rs:95)"""
    block_size, n = 65536, 6 * 65536
    src = O.generate(kind, 0xC0FFEE03, n)
    blocks, st, (d, off, total) = gpu_blocks(ctx, src, block_size, table_log, n_states)
    exp = oracle_blocks(src, block_size, table_log, n_states)
    for b, (g, e) in enumerate(zip(blocks, exp)):
        assert st[b] == 0 and g == e, "block %d differs" % b
    out, dst_ = ctx.decompress_blocks(d, total, dev(ctx, off), n, block_size, table_log, n_states)
    assert not dst_.cpu().numpy().any() and np.array_equal(out.cpu().numpy(), src)


@pytest.mark.parametrize("n_states", [32, 64, 128])
def test_block_128k(ctx, n_states):
    """BASELINE config 4 block size"""
    n = 5 * 131072 + 99
    src = O.generate("geo", 0xC0FFEE04, n)
    blocks, st, (d, off, total) = gpu_blocks(ctx, src, 131072, 0, n_states)
    exp = oracle_blocks(src, 131072, 0, n_states)
    for b, (g, e) in enumerate(zip(blocks, exp)):
        assert (st[b] == 0 and g == e) or (e is None and st[b] == 1)    # 99-byte tail < 128 states: raw
    out, _ = ctx.decompress_blocks(d, total, dev(ctx, off), n, 131072, 0, n_states)
    assert np.array_equal(out.cpu().numpy(), src)


@pytest.mark.parametrize("n_states", [64, 128])
def test_unaligned_source_and_destination_64(ctx, n_states):
    """the 64 / 128-state paths use 16 / 32-bit loads and stores when they can: odd base addresses take the byte path"""
    src = O.generate("text", 21, 3 * 8192 + 1)
    dsrc = dev(ctx, src)[1:]
    d, off, st, total = ctx.compress_blocks(dsrc, 8191, 0, n_states)
    offh = off.cpu().numpy()
    buf = d[:total].cpu().numpy().tobytes()
    exp = oracle_blocks(src[1:], 8191, 0, n_states)
    assert exp[-1] is None                                  # 3-byte tail: raw escape
    assert [buf[offh[i]:offh[i + 1]] for i in range(len(exp) - 1)] == exp[:-1]
    out, st2 = ctx.decompress_blocks(d, total, off, src.size - 1, 8191, 0, n_states)
    assert (st2.cpu().numpy() >= 0).all() and np.array_equal(out.cpu().numpy(), src[1:])


@pytest.mark.parametrize("n_states", [1, 2, 32, 64, 128])
def test_ragged_lengths(ctx, n_states):
    """every residue of the block length modulo N, lengths around multiples of the chunk (1024 / 2048 symbols)"""
    lens = list(range(max(n_states, 5), max(n_states, 5) + 70)) + [1023, 1024, 1025, 1056, 1057, 2047, 2048, 2049, 2111, 2112,
                                                                     2113, 4095, 4096, 4097, 4099, 4160, 4161, 6207, 6209]
    for ln in lens:
        src = O.generate("text", 1000 + ln, ln)
        blocks, st, (d, off, total) = gpu_blocks(ctx, src, ln, 0, n_states)
        try:
            e = O.compress_n(src, 0, n_states)[0]
        except ValueError:
            e = None
        if e is None:
            assert st[0] in (1, 2)
        else:
            assert st[0] == 0 and blocks[0] == e, ln
        out, dst_ = ctx.decompress_blocks(d, total, dev(ctx, off), ln, ln, 0, n_states)
        assert dst_.cpu().numpy()[0] >= 0 and np.array_equal(out.cpu().numpy(), src), ln


def test_degenerate_blocks(ctx):
    """blocks the reference panics on get escape codes; single-symbol blocks stay FSE"""
    bs = 256
    parts = [np.zeros(bs, np.uint8),                       # all zero: histogram.rs:98 -> RLE escape
             np.full(bs, 7, np.uint8),                     # single symbol: FSE stream, length-driven decode
             O.generate("geo", 5, bs),
             np.array([1, 2, 3], np.uint8)]                # 3-byte tail: histogram.rs:271 -> raw escape
    src = np.concatenate(parts)
    blocks, st, (d, off, total) = gpu_blocks(ctx, src, bs, 0, 2)
    assert list(st) == [2, 0, 0, 1]
    assert blocks[0] == bytes([0x0E, 0x00]) and blocks[3] == bytes([0x0F, 1, 2, 3])
    assert blocks[1] == O.compress_n(parts[1], 0, 2)[0] and blocks[2] == O.compress_n(parts[2], 0, 2)[0]
    out, dst_ = ctx.decompress_blocks(d, total, dev(ctx, off), len(src), bs, 0, 2)
    assert list(dst_.cpu().numpy()) == [2, 0, 0, 1] and np.array_equal(out.cpu().numpy(), src)
    # fewer symbols than states -> raw
    src = O.generate("text", 3, 64 + 20)
    blocks, st, (d, off, total) = gpu_blocks(ctx, src, 64, 0, 32)
    assert list(st) == [0, 1] and blocks[1] == bytes([0x0F]) + src[64:].tobytes()
    out, _ = ctx.decompress_blocks(d, total, dev(ctx, off), len(src), 64, 0, 32)
    assert np.array_equal(out.cpu().numpy(), src)


def test_empty_input(ctx):
    import torch
    src = torch.empty(0, dtype=torch.uint8, device=ctx.device)
    d, off, st, total = ctx.compress_blocks(src, 65536, 0, 32)
    assert total == 0 and off.cpu().numpy().tolist() == [0] and st.numel() == 0


def test_decode_oracle_streams_and_errors(ctx):
    """streams produced by the oracle decode on the GPU; corrupted streams report a status"""
    bs, nb = 4096, 6
    src = O.generate("geo", 77, bs * nb)
    streams = oracle_blocks(src, bs, 0, 32)
    off = np.zeros(nb + 1, np.int64)
    off[1:] = np.cumsum([len(s) for s in streams])
    comp = np.frombuffer(b"".join(streams), np.uint8).copy()
    out, st = ctx.decompress_blocks(dev(ctx, comp), comp.size, dev(ctx, off), src.size, bs, 0, 32)
    assert not st.cpu().numpy().any() and np.array_equal(out.cpu().numpy(), src)
    bad = comp.copy()
    bad[off[1] - 1] = 0                                    # block 0: marker byte zero -> stack_reader.rs:77-83
    bad[off[1]] = 0x0B                                     # block 1: table_log 16 -> histogram.rs:439-441
    bad[off[3] - 1] ^= 0x80 if bad[off[3] - 1] < 0x80 else 0xC0   # block 2: marker moved -> length mismatch
    out, st = ctx.decompress_blocks(dev(ctx, bad), bad.size, dev(ctx, off), src.size, bs, 0, 32)
    st = st.cpu().numpy()
    assert st[0] == -6 and st[1] == -3 and st[2] == -7 and not st[3:].any()
    out = out.cpu().numpy()
    assert np.array_equal(out[3 * bs:], src[3 * bs:])
    # wrong number of states: bit accounting cannot work out
    out, st = ctx.decompress_blocks(dev(ctx, comp), comp.size, dev(ctx, off), src.size, bs, 0, 16)
    assert (st.cpu().numpy() == -7).any() or not np.array_equal(out.cpu().numpy(), src)


def test_output_lands_at_arbitrary_byte_offsets(ctx):
    """the device analogue of the reference's alignment sweep (bitstream/mod.rs:151-155): block
    streams land at every byte alignment after compaction and decode from there"""
    src = O.generate("text", 5, 64 * 3000 + 17)
    blocks, st, (d, off, total) = gpu_blocks(ctx, src, 3000, 0, 32)
    assert len({int(o) % 16 for o in off}) == 16
    exp = oracle_blocks(src, 3000, 0, 32)
    assert exp[-1] is None and st[-1] == 1               # 17-byte tail < 32 states: raw escape
    assert blocks[:-1] == exp[:-1]
    out, _ = ctx.decompress_blocks(d, total, dev(ctx, off), src.size, 3000, 0, 32)
    assert np.array_equal(out.cpu().numpy(), src)


@pytest.mark.parametrize("n_states", [32, 64, 128])
def test_global_table_mode(ctx, n_states):
    """BASELINE config 5 on one GPU: one table from the whole-buffer histogram; blocks are
    header-less payloads (fse.rs:394-421); the header equals the oracle's for the u64 counts"""
    n, bs = 40 * 16384 + 5, 16384
    src = O.generate("geo", 0xC0FFEE05, n)
    dsrc = dev(ctx, src)
    counts = ctx.histogram_global(dsrc)
    header, log2 = ctx.set_global_table(counts, 11)
    h = O.histogram(src)
    rc, nh = O.normalize(h, 11)
    assert rc >= 0 and log2 == nh.log2 and header == O.ncount_write(nh)[0]
    d, off, st, total = ctx.compress_blocks(dsrc, bs, 11, n_states, table_mode=1)
    offh = off.cpu().numpy()
    buf = d[:total].cpu().numpy().tobytes()
    et = O.enc_table(nh)
    for b in range(len(offh) - 1):
        blk = src[b * bs:(b + 1) * bs]
        exp = blk.tobytes() if len(blk) < n_states else O.encode_payload(et, blk, n_states)[0]
        assert buf[offh[b]:offh[b + 1]] == exp, b
    # a fresh context decodes from the stored header alone
    import entropy_coders_b200 as E
    c2 = E.Context(0)
    assert c2.set_global_table_from_header(header) == log2
    out, dst_ = c2.decompress_blocks(d, total, off, n, bs, 11, n_states, table_mode=1)
    assert (dst_.cpu().numpy() >= 0).all() and np.array_equal(out.cpu().numpy(), src)
    c2.close()


def test_host_buffer_api(ctx):
    """fse_b200_compress_host / decompress_host: host in, host out"""
    src = O.generate("text", 11, 1 << 20)
    dst, off, st, total = ctx.compress_host(src, 65536, 0, 32)
    exp = oracle_blocks(src, 65536, 0, 32)
    assert dst[:total].tobytes() == b"".join(exp) and not st.any()
    out, st2 = ctx.decompress_host(dst, total, off, src.size, 65536, 0, 32)
    assert not st2.any() and np.array_equal(out, src)


def test_device_generators_match_oracle(ctx):
    for i, kind in enumerate(["geo", "text", "few", "uniform"]):
        g = ctx.generate(kind, 0xC0FFEE00 + i, 100003, first_index=12345).cpu().numpy()
        assert np.array_equal(g, O.generate(kind, 0xC0FFEE00 + i, 100003, first_index=12345))


# ------------------------------------------------------------------------------------ full size

def test_full_size_roundtrip_256mib(ctx):
    """BASELINE config 2 at full size: encode -> decode round trip, sizes consistent, and a sample of
    blocks bit-exact against the oracle"""
    import torch
    n, bs = 256 << 20, 65536
    src = ctx.generate("text", 0xC0FFEE02, n)
    d, off, st, total = ctx.compress_blocks(src, bs, 0, 32)
    assert not st.cpu().numpy().any()
    offh = off.cpu().numpy()
    assert offh[-1] == total and (np.diff(offh) > 0).all()
    assert 0.60 < total / n < 0.72
    out, dst_ = ctx.decompress_blocks(d, total, off, n, bs, 0, 32)
    assert not dst_.cpu().numpy().any()
    assert torch.equal(out, src)
    for b in (0, 1, 2047, 4095):
        blk = src[b * bs:(b + 1) * bs].cpu().numpy()
        assert np.array_equal(blk, O.generate("text", 0xC0FFEE02, bs, first_index=b * bs))
        assert d[offh[b]:offh[b + 1]].cpu().numpy().tobytes() == O.compress_n(blk, 0, 32)[0]


@pytest.mark.parametrize("n_states", [32, 128])
def test_blocks_of_4_mib(ctx, n_states):
    """blocks beyond the 16-bit histogram kernel's 1 MiB (32-bit counter columns, one CTA per block) and far beyond one
    chunk of the encoder: bytes equal the oracle's, round trip exact"""
    bs, n = 4 << 20, (9 << 20) + 777
    src = O.generate("text", 21, n)
    blocks, st, (d, off, total) = gpu_blocks(ctx, src, bs, 0, n_states)
    exp = oracle_blocks(src, bs, 0, n_states)
    assert not st.any() and all(g == e for g, e in zip(blocks, exp))
    out, dst_ = ctx.decompress_blocks(d, total, dev(ctx, off), n, bs, 0, n_states)
    assert not dst_.cpu().numpy().any() and np.array_equal(out.cpu().numpy(), src)


def test_more_than_4gib_on_one_gpu(ctx):
    """64-bit indexing end to end: 4.25 GiB of incompressible bytes (input offsets, compressed offsets and the dense
    output all pass 2^32), 128 states, 128 KiB blocks; round trip plus oracle bytes for blocks on both sides of 2^32"""
    import torch
    n, bs, ns = (17 << 28) + 12345, 131072, 128
    src = ctx.generate("uniform", 0xC0FFEE03, n)
    d, off, st, total = ctx.compress_blocks(src, bs, 11, ns)
    assert not st.cpu().numpy().any()
    offh = off.cpu().numpy()
    nb = len(offh) - 1
    assert nb == (n + bs - 1) // bs and offh[-1] == total and total > (1 << 32) and (np.diff(offh) > 0).all()
    for b in (0, nb // 2, int(np.searchsorted(offh, 1 << 32)) - 1, int(np.searchsorted(offh, 1 << 32)), nb - 2, nb - 1):
        blk = src[b * bs:(b + 1) * bs].cpu().numpy()
        assert d[int(offh[b]):int(offh[b + 1])].cpu().numpy().tobytes() == O.compress_n(blk, 11, ns)[0], b
    out, dst_ = ctx.decompress_blocks(d, total, off, n, bs, 11, ns)
    assert not dst_.cpu().numpy().any()
    assert torch.equal(out, src)
    del out, d, src
    torch.cuda.empty_cache()


# ------------------------------------------------------------------------------------ crate mirror

def test_crate_compress_roundtrip():
    """src/lib.rs:280-302 `compress` / `compress2` re-expressed on the crate-shaped API"""
    import entropy_coders_b200 as E
    src = O.generate("geo", 0xC0FFEE01, 1 << 16).tobytes()
    dst, dec = bytearray(), bytearray()
    hist, bits = E.fse_compress(src, dst)
    assert bytes(dst) == O.compress_n(src, 0, 1)[0] and bits == O.compress_n(src, 0, 1)[2]
    assert E.fse_decompress(bytes(dst), dec) == len(src) and bytes(dec) == src
    dst, dec = bytearray(b"prefix"), bytearray(b"xy")
    bits = E.fse_compress2(src, dst)
    assert bytes(dst[6:]) == O.compress_n(src, 0, 2)[0] and bits == O.compress_n(src, 0, 2)[2]
    assert E.fse_decompress2(bytes(dst[6:]), dec) == len(src) and bytes(dec[2:]) == src
    assert E.fse_decompress2(b"\x0f\x00\x00", bytearray()) is None      # header error -> None (lib.rs:219)
    with pytest.raises(E.crate.Panic):
        E.fse_compress2(b"", bytearray())                                # lib.rs:154 unwrap on None


def test_crate_hist_verify():
    """src/histogram.rs:553-587 `hist_verify`, :595-656 known answers, on the crate-shaped API"""
    import entropy_coders_b200 as E
    for log2 in (8, 11, 15):
        data = np.repeat(np.arange(256, dtype=np.uint8), 1 << (log2 - 8)).tobytes()
        hist = E.Histogram(data)
        assert all(x == 1 << (log2 - 8) for x in hist.table())           # :607-616
        nh = hist.normalize(log2)
        assert sum(abs(x) for x in nh.table()) == 1 << nh.log2_sum()      # :566-568
        assert all((h == 0) == (q == 0) for h, q in zip(hist.table(), nh.table()))
        enc = bytearray()
        nh.write(enc)
        assert len(enc) <= nh.write_bound()
        enc.extend(b"I am a test")
        dec, rem = E.NormHistogram.read(bytes(enc))                       # :580-586
        assert rem == b"I am a test" and dec == nh
    nh = E.NormHistogram.new(bytes(range(256)))                           # flat_256, :589-593
    assert nh.log2_sum() == 9 and nh.symbol_count() == 0
    et, dt = E.fse.EncodeTable(nh), E.fse.DecodeTable(nh)
    assert len(et.table) == 512 and len(dt.table) == 512
    with pytest.raises(E.HistError):
        E.NormHistogram.read(bytes([0x0F, 0, 0, 0]))


def test_host_buffer_api_pipelined(ctx):
    """inputs above 64 MiB take the chunked, copy/compute-overlapped host path: same bytes, same offsets"""
    import torch
    n = (70 << 20) + 1234
    src = ctx.generate("text", 31, n)
    hsrc = src.cpu().numpy()
    d, off, st, total = ctx.compress_blocks(src, 65536, 0, 64)
    hd, hoff, hst, htotal = ctx.compress_host(hsrc, 65536, 0, 64)
    assert htotal == total and np.array_equal(hoff.astype(np.int64), off.cpu().numpy())
    assert np.array_equal(hd[:htotal], d[:total].cpu().numpy()) and np.array_equal(hst, st.cpu().numpy())
    out, st2 = ctx.decompress_host(hd, htotal, hoff, n, 65536, 0, 64)
    assert (st2 >= 0).all() and np.array_equal(out, hsrc)
    # pinned torch buffers (what bench.py uses)
    psrc = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    psrc.copy_(src)
    pdst = torch.empty(htotal + 64, dtype=torch.uint8, pin_memory=True)
    _, poff, _, ptotal = ctx.compress_host(psrc, 65536, 0, 64, dst=pdst)
    assert ptotal == total and torch.equal(pdst[:total], d[:total].cpu())


@pytest.mark.parametrize("n_states", [1, 2])
def test_exhaust_decode_matches_reference_termination(ctx, n_states):
    """fse_decompress / fse_decompress2 stop when the bit stack cannot supply num_bits (lib.rs:198,228-241);
    with final states that need 0 bits they over-produce.  The exhaust-mode kernel must yield
    exactly what the oracle's exhaustion decoder yields, byte for byte, including the surplus."""
    import torch
    cases = [("geo", 5000), ("text", 4097), ("few", 3000), ("few", 777), ("uniform", 2048), ("geo", 33)]
    streams, expect = [], []
    for i, (kind, n) in enumerate(cases):
        src = O.generate(kind, 900 + i, n).tobytes()
        comp = O.compress_n(src, 0, n_states)[0]
        try:
            exp = O.decompress_n_exhaust(comp, n_states, 4 * n + 64)
        except ValueError:
            exp = None                                   # ran past the capacity: the reference never terminates
        streams.append(comp)
        expect.append((src, exp))
    cap = max(4 * n + 64 for _, n in cases)
    off = np.zeros(len(streams) + 1, np.int64)
    off[1:] = np.cumsum([len(s) for s in streams])
    comp = np.frombuffer(b"".join(streams), np.uint8).copy()
    out, out_len, st = ctx.decompress_exhaust(dev(ctx, comp), comp.size, dev(ctx, off), len(streams), cap, 15, n_states)
    out, out_len, st = out.cpu().numpy(), out_len.cpu().numpy(), st.cpu().numpy()
    surplus = 0
    for i, (src, exp) in enumerate(expect):
        if exp is None or len(exp) > cap:
            assert st[i] == -2, i
            continue
        assert st[i] == 0 and out_len[i] == len(exp), (i, st[i], out_len[i], len(exp))
        assert out[i, :len(exp)].tobytes() == exp
        assert exp[:len(src)] == src
        surplus += len(exp) - len(src)
    if n_states == 1:
        assert surplus >= 0


@pytest.mark.parametrize("block_size", [2 << 20, (2 << 20) + 48, (1 << 20) + 4099, 3 << 18, 1 << 17])
def test_histogram_of_few_or_large_blocks(ctx, block_size):
    """few blocks, or blocks above 1 MiB, are counted in pieces (one warp per piece, k_hist_sum_pieces); sizes that do not
    split evenly keep 32-bit counter columns (k_hist_blocks): counts and table_len either way, ragged tail included"""
    src = O.generate("geo", 3, (5 << 20) + 99)
    src[src > 200] = 7                                        # table_len below 256
    counts, tlen = ctx.histogram_blocks(dev(ctx, src), block_size)
    counts = counts.cpu().numpy().view(np.uint32)
    tlen = tlen.cpu().numpy()
    assert counts.shape[0] == -(-src.size // block_size)
    for b in range(counts.shape[0]):
        exp = np.bincount(src[b * block_size:(b + 1) * block_size], minlength=256)
        assert np.array_equal(counts[b], exp)
        assert tlen[b] == int(np.max(np.nonzero(exp)[0])) + 1


def test_argument_errors(ctx):
    import torch
    import entropy_coders_b200 as E
    src = dev(ctx, O.generate("geo", 1, 70000))
    with pytest.raises(E.FseError):
        ctx.compress_blocks(src, 65536, 0, 3)            # n_states must be a power of two
    with pytest.raises(E.FseError):
        ctx.compress_blocks(src, 65536, 14, 64)          # 64 states need table_log <= 13
    with pytest.raises(E.FseError):
        ctx.compress_blocks(src, 0, 0, 32)
    small = torch.empty(10, dtype=torch.uint8, device=ctx.device)
    off = torch.zeros(3, dtype=torch.int64, device=ctx.device)
    st = torch.zeros(2, dtype=torch.int32, device=ctx.device)
    with pytest.raises(E.FseError):
        ctx.compress_blocks(src, 65536, 0, 32, out=(small, off, st))     # dst_cap below the bound
    fresh = E.Context(0)
    with pytest.raises(E.FseError):
        fresh.compress_blocks(src, 65536, 11, 32, table_mode=1)         # global table not installed
    fresh.close()


def _structured(rng, n, style):
    if style == 0:      # long runs
        out = np.repeat(rng.integers(0, 256, n // 37 + 1, dtype=np.uint8), 37)[:n]
    elif style == 1:    # one dominant symbol (p ~ 0.99): most transitions cost 0 bits
        out = np.where(rng.random(n) < 0.99, 65, rng.integers(0, 256, n)).astype(np.uint8)
    elif style == 2:    # two symbols
        out = rng.integers(0, 2, n, dtype=np.uint8) * 200 + 3
    elif style == 3:    # sawtooth over the whole alphabet
        out = (np.arange(n) % 251).astype(np.uint8)
    elif style == 4:    # sparse alphabet with big holes (zero runs in the header)
        out = rng.choice(np.array([0, 1, 30, 31, 32, 100, 200, 255], dtype=np.uint8), n)
    else:               # mixture of segments
        out = np.concatenate([O.generate(k, int(rng.integers(1 << 30)), n // 3 + 1) for k in ("geo", "uniform", "few")])[:n]
    return np.ascontiguousarray(out)


def test_randomised_differential(ctx):
    """seeded sweep over data styles, block sizes, state counts and table_log: every block either is the
    oracle's bytes or is one the reference panics on (then it carries an escape); decode always round-trips"""
    rng = np.random.default_rng(20261018)
    checked = 0
    for trial in range(48):
        n_states = int(rng.choice([1, 2, 4, 8, 16, 32, 64, 128]))
        bs = int(rng.choice([130, 257, 1000, 4096, 5000, 20000, 65536]))
        if n_states <= 2:
            bs = min(bs, 5000)                              # the one-lane paths are for parity, not speed
        n = int(bs * rng.integers(1, 5) + rng.integers(0, bs))
        tl = int(rng.choice([0, 0, 5, 7, 9, 11, 12, 13]))
        src = _structured(rng, n, trial % 6)
        blocks, st, (d, off, total) = gpu_blocks(ctx, src, bs, tl, n_states)
        for b, g in enumerate(blocks):
            blk = src[b * bs:(b + 1) * bs]
            try:
                e = O.compress_n(blk, tl, n_states)[0]
            except ValueError:
                e = None
            if e is None:
                assert st[b] in (1, 2) and g[0] in (0x0E, 0x0F), (trial, b, st[b])
            else:
                assert st[b] == 0 and g == e, (trial, b, n_states, bs, tl)
                checked += 1
        out, dst_ = ctx.decompress_blocks(d, total, dev(ctx, off), n, bs, tl, n_states)
        assert (dst_.cpu().numpy() >= 0).all() and np.array_equal(out.cpu().numpy(), src), (trial, n_states, bs, tl)
    assert checked > 100


def test_frame_container(ctx):
    """a self-describing frame carries parameters, offsets and payload; both table modes"""
    import entropy_coders_b200 as E
    for kind, n, bs, mode, tl in [("text", 5 * 65536 + 321, 65536, 0, 0), ("geo", 300000, 16384, 1, 11), ("few", 1000, 300, 0, 9)]:
        src = O.generate(kind, 55, n)
        frame = ctx.frame_compress(src, bs, tl, 128, mode)
        info = ctx.frame_info(frame)
        assert info == {"block_size": bs, "table_log": tl, "n_states": 128, "table_mode": mode, "segment_size": 0, "flags": 0, "n": n}
        fresh = E.Context(0)                                  # nothing but the frame is needed to decode
        assert np.array_equal(fresh.frame_decompress(frame), src)
        fresh.close()
        if mode == 0:                                         # the payload is the dense block streams
            exp = b"".join(e if e is not None else bytes([0x0F]) + src[i * bs:(i + 1) * bs].tobytes()
                           for i, e in enumerate(oracle_blocks(src, bs, tl, 128)))
            assert frame[-len(exp):].tobytes() == exp
    bad = frame.copy()
    bad[0] ^= 0xFF
    with pytest.raises(E.FseError):
        ctx.frame_info(bad)
    with pytest.raises(E.FseError):
        ctx.frame_info(frame[:40])


# ------------------------------------------------------------------------------------ segmented per-block mode

def oracle_segments(src, block_size, seg, table_log):
    """Expected stream list of the segmented mode: per block ONE header (NormHistogram::new + write of the whole
    block, src/histogram.rs:299-303,376-431) and, per segment, the header-less 128-state payload of that slice coded
    with the block's table (the composition the crate tests at src/fse.rs:394-421).  -> (streams, statuses)"""
    streams, status = [], []
    for o in range(0, len(src), block_size):
        blk = src[o:o + block_size]
        nseg = (len(blk) + seg - 1) // seg
        h = O.histogram(blk)
        esc = None
        if h.table_len <= 1:
            esc = (bytes([0x0E, 0x00]), 2)
        elif len(blk) <= 4 and table_log == 0:
            esc = (bytes([0x0F]) + blk.tobytes(), 1)
        elif len(blk) < 128:
            esc = (bytes([0x0F]) + blk.tobytes(), 1)
        if esc is None:
            tl = table_log
            if tl == 0:
                rc, tl = O.optimal_log2(h)
                assert rc >= 0
            rc, nh = O.normalize(h, tl)
            assert rc >= 0
            header = O.ncount_write(nh)[0]
            et = O.enc_table(nh)
        for k in range(nseg):
            if esc is not None:
                streams.append(esc[0] if k == 0 else b"")
                status.append(esc[1])
                continue
            sl = blk[k * seg:(k + 1) * seg]
            if len(sl) < 128:
                body, st = sl.tobytes(), 1
            else:
                body, st = O.encode_payload(et, sl, 128)[0], 0
            streams.append((header if k == 0 else b"") + body)
            status.append(st)
    return streams, status


@pytest.mark.parametrize("kind,bs,seg,tl,n", [
    ("geo", 131072, 8192, 0, 5 * 131072),              # BASELINE config 4's shape
    ("text", 65536, 8192, 0, 3 * 65536 + 20000),        # config 2's shape, ragged last block (3 segments, one short)
    ("few", 65536, 4096, 11, 2 * 65536 + 4096 + 100),    # a 100-byte tail segment is stored raw
    ("uniform", 32768, 2048, 9, 4 * 32768 + 2048 * 3 + 777),
    ("geo", 16384, 512, 10, 16384 * 3 + 512 * 5 + 129),
    ("text", 8192, 1024, 0, 8192 * 37 + 1),              # more blocks than SMs hold at once; a 1-byte last block (escape)
])
def test_segmented_blocks_bit_exact(ctx, kind, bs, seg, tl, n):
    """segment_size > 0: one table and one header per block, its segments coded as independent 128-state streams by the
    warps of one CTA against bank-replicated tables.  Every stream equals the oracle's composition, byte for byte."""
    src = O.generate(kind, 0xC0FFEE04, n)
    d, off, st, total = ctx.compress_blocks(dev(ctx, src), bs, tl, 128, segment_size=seg)
    offh = off.cpu().numpy().astype(np.int64)
    buf = d[:total].cpu().numpy().tobytes()
    exp, est = oracle_segments(src, bs, seg, tl)
    assert len(offh) - 1 == len(exp) == ctx.num_streams(n, ctx.params(bs, tl, 128, 0, seg))
    sth = st.cpu().numpy()
    for s_, e in enumerate(exp):
        assert buf[offh[s_]:offh[s_ + 1]] == e, (s_, len(e), offh[s_ + 1] - offh[s_])
        assert sth[s_] == est[s_], (s_, sth[s_], est[s_])
    out, dst_ = ctx.decompress_blocks(d, total, off, n, bs, tl, 128, segment_size=seg)
    assert (dst_.cpu().numpy() == np.array(est)).all()
    assert np.array_equal(out.cpu().numpy(), src)


def test_segmented_degenerate_blocks(ctx):
    """blocks the reference panics on (all zero, tiny) inside a segmented stream: block-level escapes in the first stream"""
    bs, seg = 16384, 2048
    src = O.generate("text", 5, 6 * bs + 3)
    src[bs:2 * bs] = 0                                      # all-zero block -> 0x0E
    src[3 * bs:4 * bs] = 65                                 # one symbol: stays FSE (length-driven decode)
    d, off, st, total = ctx.compress_blocks(dev(ctx, src), bs, 0, 128, segment_size=seg)
    offh = off.cpu().numpy().astype(np.int64)
    buf = d[:total].cpu().numpy().tobytes()
    exp, est = oracle_segments(src, bs, seg, 0)
    for s_, e in enumerate(exp):
        assert buf[offh[s_]:offh[s_ + 1]] == e, s_
    assert list(st.cpu().numpy()) == est
    out, dst_ = ctx.decompress_blocks(d, total, off, src.size, bs, 0, 128, segment_size=seg)
    assert (dst_.cpu().numpy() >= 0).all() and np.array_equal(out.cpu().numpy(), src)


def test_segmented_decode_errors(ctx):
    """corrupt segmented streams give a negative status for the streams concerned, never a crash"""
    bs, seg = 65536, 8192
    src = O.generate("geo", 9, 4 * bs)
    d, off, st, total = ctx.compress_blocks(dev(ctx, src), bs, 0, 128, segment_size=seg)
    offh = off.cpu().numpy().astype(np.int64)
    bad = d.clone()
    bad[int(offh[8])] = 0x3B                                # block 1 header: table_log nibble 11 + 5 > 15
    out, dst_ = ctx.decompress_blocks(bad, total, off, src.size, bs, 0, 128, segment_size=seg)
    s = dst_.cpu().numpy()
    assert (s[8:16] < 0).all() and (s[:8] == 0).all() and (s[16:] == 0).all()
    o = out.cpu().numpy()
    assert np.array_equal(o[:bs], src[:bs]) and np.array_equal(o[2 * bs:], src[2 * bs:])
    bad = d.clone()
    bad[int(offh[20]) - 1] = 0                              # stream 19 loses its marker byte
    out, dst_ = ctx.decompress_blocks(bad, total, off, src.size, bs, 0, 128, segment_size=seg)
    s = dst_.cpu().numpy()
    assert s[19] < 0 and (np.delete(s, 19) == 0).all()


def test_segmented_host_and_frame(ctx):
    """the host-buffer entry points and the frame carry segment_size"""
    import entropy_coders_b200 as E
    bs, seg = 131072, 8192
    src = O.generate("geo", 77, 70 * bs + 12345)            # > 64 MiB? no: 9 MiB, one chunk; the pipelined path is in test_full_size
    dst, off, st, total = ctx.compress_host(src, bs, 0, 128, segment_size=seg)
    exp, est = oracle_segments(src, bs, seg, 0)
    assert total == sum(len(e) for e in exp) and dst[:total].tobytes() == b"".join(exp)
    out, st2 = ctx.decompress_host(dst, total, off, src.size, bs, 0, 128, segment_size=seg)
    assert np.array_equal(out, src) and (st2 >= 0).all()
    frame = ctx.frame_compress(src, bs, 0, 128, segment_size=seg)
    assert ctx.frame_info(frame)["segment_size"] == seg
    fresh = E.Context(0)
    assert np.array_equal(fresh.frame_decompress(frame), src)
    fresh.close()


def test_frame_forged_headers(ctx):
    """every size in a frame header is untrusted; forged values must be rejected, not wrapped"""
    import struct
    import entropy_coders_b200 as E
    src = O.generate("text", 3, 200000)
    frame = ctx.frame_compress(src, 65536, 0, 128)
    hdr = bytearray(frame[:56].tobytes())
    def forged(**kw):
        f = list(struct.unpack("<IHHIIIIQQQII", bytes(hdr)))
        names = ["magic", "version", "n_states", "block_size", "table_log", "table_mode", "ghb", "n", "nstreams", "payload", "seg", "flags"]
        for k, v in kw.items():
            f[names.index(k)] = v
        out = frame.copy()
        out[:56] = np.frombuffer(struct.pack("<IHHIIIIQQQII", *f), dtype=np.uint8)
        return out
    for kw in (dict(payload=2 ** 64 - 8), dict(payload=2 ** 63), dict(block_size=1, n=2 ** 61, nstreams=2 ** 61),
               dict(nstreams=2 ** 61 - 1), dict(n_states=3), dict(n_states=256), dict(table_log=16), dict(table_mode=7),
               dict(ghb=400), dict(seg=100), dict(flags=0x80), dict(n=2 ** 63), dict(version=1)):
        with pytest.raises(E.FseError):
            ctx.frame_info(forged(**kw))
        with pytest.raises(E.FseError):
            ctx.frame_decompress(forged(**kw))
    # a flipped payload byte: the frame parses, a block fails, and that is an exception, not silent garbage
    bad = frame.copy()
    pay0 = len(frame) - int(struct.unpack("<Q", bytes(hdr[40:48]))[0])
    bad[pay0] = 0x3F
    with pytest.raises(E.FseError):
        ctx.frame_decompress(bad)
    with pytest.raises(E.FseError):
        ctx.frame_decompress(frame[:len(frame) - 1000])      # truncated


# ------------------------------------------------------------------------------------ full-size runs at the bench's parameters

def _full_size(ctx, kind, seed, n, bs, tl, mode, sample_blocks, ratio_range, segment_size=0):
    """BASELINE.json shapes at the state count bench.py uses (128): round trip, index consistency, a checksum of the
    decoded bytes against the generator, and a sample of blocks byte-exact against the oracle"""
    import torch
    src = ctx.generate(kind, seed, n)
    header = None
    if mode == 1:
        header, log2 = ctx.set_global_table(ctx.histogram_global(src), tl)
        nh = O.ncount_read(header)[1]
        et = O.enc_table(nh)
    d, off, st, total = ctx.compress_blocks(src, bs, tl, 128, table_mode=mode, segment_size=segment_size)
    assert not st.cpu().numpy().any()
    offh = off.cpu().numpy()
    assert offh[0] == 0 and offh[-1] == total and (np.diff(offh) > 0).all()
    assert ratio_range[0] < total / n < ratio_range[1], total / n
    out, dst_ = ctx.decompress_blocks(d, total, off, n, bs, tl, 128, table_mode=mode, segment_size=segment_size)
    assert not dst_.cpu().numpy().any()
    assert torch.equal(out, src)
    for b in sample_blocks:
        blk = src[b * bs:(b + 1) * bs].cpu().numpy()
        assert np.array_equal(blk, O.generate(kind, seed, bs, first_index=b * bs))       # device generator == oracle generator
        got = d[offh[b]:offh[b + 1]].cpu().numpy().tobytes() if not segment_size else None
        if mode == 1:
            assert got == O.encode_payload(et, blk, 128)[0], b
        elif not segment_size:
            assert got == O.compress_n(blk, tl, 128)[0], b
    del src, d, out
    torch.cuda.empty_cache()


def test_full_size_c2_n128(ctx):
    """config 2: 256 MiB text-like, 64 KiB blocks (4 096), optimal_log2, 128 states"""
    _full_size(ctx, "text", 0xC0FFEE02, 256 << 20, 65536, 0, 0, (0, 1, 777, 2047, 4095), (0.60, 0.72))


@pytest.mark.parametrize("kind,tl,rng", [("few", 9, (0.05, 0.12)), ("few", 11, (0.05, 0.12)), ("few", 12, (0.05, 0.12)),
                                         ("uniform", 9, (1.0, 1.02)), ("uniform", 11, (1.0, 1.02)), ("uniform", 12, (1.0, 1.02))])
def test_full_size_c3_n128(ctx, kind, tl, rng):
    """config 3: 1 GiB few-symbol / uniform, 64 KiB blocks (16 384), table_log 9 / 11 / 12, 128 states"""
    _full_size(ctx, kind, 0xC0FFEE03, 1 << 30, 65536, tl, 0, (0, 5, 8191, 16383), rng)


def test_full_size_c4_n128(ctx):
    """config 4's shape: geometric bytes, 128 KiB blocks, 16 384 of them (2 GiB), 128 states"""
    _full_size(ctx, "geo", 0xC0FFEE04, 2 << 30, 131072, 0, 0, (0, 3, 9999, 16383), (0.44, 0.47))


def test_full_size_c5_global_n128(ctx):
    """config 5's shape on one GPU: one table from the whole-buffer histogram, 16 384 header-less 128 KiB blocks"""
    _full_size(ctx, "geo", 0xC0FFEE05, 2 << 30, 131072, 11, 1, (0, 4, 8000, 16383), (0.44, 0.47))


def test_full_size_c4_segmented(ctx):
    """the segmented per-block mode at config 4's shape (8 KiB segments, 16 per block)"""
    _full_size(ctx, "geo", 0xC0FFEE04, 1 << 30, 131072, 0, 0, (0, 8191), (0.45, 0.49), segment_size=8192)


@pytest.mark.parametrize("kind,tl,bs", [("text", 9, 30000), ("uniform", 10, 16384), ("few", 5, 4096), ("geo", 11, 131072), ("text", 0, 65536)])
def test_global_table_sweep(ctx, kind, tl, bs):
    """CTA-owned bank-replicated tables (global mode): table_log 5..11 (32 copies up to 10, 16 at 11), ragged blocks,
    every block byte-exact against the oracle's header-less 128-state stream"""
    n = 23 * bs + 1000
    src = O.generate(kind, 0xC0FFEE05, n)
    dsrc = dev(ctx, src)
    header, log2 = ctx.set_global_table(ctx.histogram_global(dsrc), tl)
    h = O.histogram(src)
    if tl == 0:
        rc, tl_eff = O.optimal_log2(h)
        assert rc >= 0
    else:
        tl_eff = tl
    rc, nh = O.normalize(h, tl_eff)
    assert rc >= 0 and log2 == nh.log2 and header == O.ncount_write(nh)[0]
    d, off, st, total = ctx.compress_blocks(dsrc, bs, tl, 128, table_mode=1)
    offh = off.cpu().numpy()
    buf = d[:total].cpu().numpy().tobytes()
    et = O.enc_table(nh)
    for b in range(len(offh) - 1):
        blk = src[b * bs:(b + 1) * bs]
        exp = blk.tobytes() if len(blk) < 128 else O.encode_payload(et, blk, 128)[0]
        assert buf[offh[b]:offh[b + 1]] == exp, b
    out, dst_ = ctx.decompress_blocks(d, total, off, n, bs, tl, 128, table_mode=1)
    assert (dst_.cpu().numpy() >= 0).all() and np.array_equal(out.cpu().numpy(), src)
    # unaligned source and destination
    d2, off2, st2, total2 = ctx.compress_blocks(dsrc[3:], bs, tl, 128, table_mode=1)
    out2, dst2 = ctx.decompress_blocks(d2, total2, off2, n - 3, bs, tl, 128, table_mode=1)
    assert (dst2.cpu().numpy() >= 0).all() and np.array_equal(out2.cpu().numpy(), src[3:])


@pytest.mark.parametrize("n_states", [2, 32, 64, 128])
def test_raw_if_expands(ctx, n_states):
    """SURVEY 8f (f2), opt-in: a block whose coded form is not smaller than 1 + its length is stored 0x0F + raw.  Without
    the flag the output stays the reference's (expanding) stream; compressible data is unaffected by the flag."""
    import entropy_coders_b200 as E
    bs = 65536
    uni = O.generate("uniform", 1, 6 * bs + 1000)
    d0, off0, st0, tot0 = ctx.compress_blocks(dev(ctx, uni), bs, 0, n_states)
    d1, off1, st1, tot1 = ctx.compress_blocks(dev(ctx, uni), bs, 0, n_states, flags=E.FLAG_RAW_IF_EXPANDS)
    assert tot0 > uni.size and not st0.cpu().numpy().any()           # the reference expands uniform bytes
    o1 = off1.cpu().numpy()
    assert (st1.cpu().numpy() == 1).all() and tot1 == uni.size + len(o1) - 1
    buf = d1[:tot1].cpu().numpy().tobytes()
    for b in range(len(o1) - 1):
        assert buf[o1[b]:o1[b + 1]] == bytes([0x0F]) + uni[b * bs:(b + 1) * bs].tobytes()
    out, dst_ = ctx.decompress_blocks(d1, tot1, off1, uni.size, bs, 0, n_states, flags=E.FLAG_RAW_IF_EXPANDS)
    assert np.array_equal(out.cpu().numpy(), uni)
    out, dst_ = ctx.decompress_blocks(d1, tot1, off1, uni.size, bs, 0, n_states)    # the decoder needs no flag
    assert np.array_equal(out.cpu().numpy(), uni)
    txt = O.generate("text", 2, 3 * bs)
    a = ctx.compress_blocks(dev(ctx, txt), bs, 0, n_states)
    b_ = ctx.compress_blocks(dev(ctx, txt), bs, 0, n_states, flags=E.FLAG_RAW_IF_EXPANDS)
    assert a[3] == b_[3] and a[0][:a[3]].cpu().numpy().tobytes() == b_[0][:b_[3]].cpu().numpy().tobytes()


def test_crate_decompress_of_highly_skewed_data():
    """a valid reference stream can expand more than 64 x (p ~ 0.997); the crate mirror grows its capacity
    instead of reporting 'does not terminate'"""
    import entropy_coders_b200 as E
    rng = np.random.default_rng(5)
    src = np.where(rng.random(400000) < 0.003, 1, 0).astype(np.uint8)
    src[-1] = 1
    comp = bytearray()
    E.fse_compress2(src.tobytes(), comp)
    assert len(src) > 64 * len(comp)
    out = bytearray()
    n = E.fse_decompress2(bytes(comp), out)
    assert n == len(src) and bytes(out) == src.tobytes()
    assert bytes(comp) == O.ref_compress2(src)


def test_global_table_with_unknown_symbols_is_memory_safe(ctx):
    """a byte the installed global table gives no probability (the table was built from other data).  Like the
    crate's Encoder the kernels do not look for it (documented in fse_b200.h); what must hold is that nothing faults,
    no other block is disturbed, and every block WITHOUT such a byte still round-trips."""
    bs = 16384
    a = O.generate("few", 21, 40 * bs)                       # symbols 0..3 only
    b = a.copy()
    b[5 * bs + 100: 5 * bs + 4000] = 77                        # block 5 and block 17 get bytes the table does not know
    b[17 * bs: 18 * bs] = 200
    header, log2 = ctx.set_global_table(ctx.histogram_global(dev(ctx, a)), 11)
    # the check the kernels leave out is an entry point of its own (fse_b200_global_table_covers)
    assert ctx.global_table_covers(dev(ctx, a)) == 0
    assert ctx.global_table_covers(dev(ctx, b)) == 3900 + bs
    d, off, st, total = ctx.compress_blocks(dev(ctx, b), bs, 11, 128, table_mode=1)
    out, st2 = ctx.decompress_blocks(d, total, off, b.size, bs, 11, 128, table_mode=1)
    o = out.cpu().numpy()
    for blk in range(40):
        if blk not in (5, 17):
            assert np.array_equal(o[blk * bs:(blk + 1) * bs], b[blk * bs:(blk + 1) * bs]), blk
    # and the context is still healthy
    d, off, st, total = ctx.compress_blocks(dev(ctx, a), bs, 11, 128, table_mode=1)
    out, st2 = ctx.decompress_blocks(d, total, off, a.size, bs, 11, 128, table_mode=1)
    assert np.array_equal(out.cpu().numpy(), a) and (st2.cpu().numpy() >= 0).all()


# ------------------------------------------------------------------------------------ thread-per-stream coders (fse_tps.cuh)

@pytest.mark.parametrize("n_states", [1, 2])
@pytest.mark.parametrize("block_size,table_log", [(1000, 0), (1001, 9), (777, 0), (2048, 12)])
def test_many_streams_in_the_reference_formats(ctx, n_states, block_size, table_log):
    """>= 4096 blocks per state of one or two states take the thread-per-stream kernels: every block's bytes equal the oracle's
    fse_compress / fse_compress2 (odd and even lengths: the state of a symbol is its index's parity), escapes included,
    and the decoder returns the input"""
    nb = 4500 * n_states                                       # the decoder switches at 4 096 blocks per state
    n = nb * block_size - 333                                  # ragged last block
    src = O.generate("text" if block_size & 1 else "geo", 0xC0FFEE10 + n_states + block_size, n)
    src[5 * block_size:6 * block_size] = 0                     # all-zero block: 0x0E escape (histogram.rs:98)
    src[9 * block_size:10 * block_size] = 65                   # one symbol: still FSE, zero payload bits per symbol
    blocks, st, (d, off, total) = gpu_blocks(ctx, src, block_size, table_log, n_states)
    scratch, sizes, status = O.compress_blocks(src, block_size, table_log, n_states, threads=8)
    assert len(blocks) == nb == len(sizes)
    for b in range(nb):
        if status[b] == 0:
            assert st[b] == 0 and blocks[b] == scratch[b, :int(sizes[b])].tobytes(), "block %d differs" % b
        else:
            assert st[b] in (1, 2), (b, st[b], status[b])     # what the reference panics on is stored with an escape
    out, dst_ = ctx.decompress_blocks(d, total, dev(ctx, off), n, block_size, table_log, n_states)
    dst_ = dst_.cpu().numpy()
    assert ((dst_ == 0) | (dst_ == 1) | (dst_ == 2)).all() and np.array_equal(out.cpu().numpy(), src)


@pytest.mark.parametrize("n_states", [1, 2])
def test_many_streams_decode_the_oracles_streams_and_survive_damage(ctx, n_states):
    """the thread-per-stream decoder on streams written by the oracle, then the seeded damage of
    test_corrupted_streams_never_fault: every block ends with a status, untouched blocks decode exactly"""
    bs, nb = 600, 4200 * n_states
    src = O.generate("text", 31 + n_states, bs * nb)
    scratch, sizes, status = O.compress_blocks(src, bs, 0, n_states, threads=8)
    assert not status.any()
    good = np.concatenate([scratch[b, :int(sizes[b])] for b in range(nb)])
    offh = np.zeros(nb + 1, dtype=np.int64)
    offh[1:] = np.cumsum(sizes)
    out, st = ctx.decompress_blocks(dev(ctx, good), good.size, dev(ctx, offh), src.size, bs, 0, n_states)
    assert not st.cpu().numpy().any() and np.array_equal(out.cpu().numpy(), src)
    rng = np.random.default_rng(77 + n_states)
    for trial in range(40):
        bad, boff = good.copy(), offh.copy()
        kind = trial % 4
        touched = set()
        for _ in range(60):
            b = int(rng.integers(0, nb))
            touched.add(b)
            if kind == 0:
                bad[int(rng.integers(boff[b], boff[b + 1]))] ^= np.uint8(1 << int(rng.integers(0, 8)))
            elif kind == 1:
                k = int(rng.integers(1, 30))
                bad[boff[b]:boff[b] + k] = rng.integers(0, 256, k, dtype=np.uint8)
            elif kind == 2:
                k = int(min(rng.integers(1, 40), boff[b + 1] - boff[b]))
                bad[boff[b + 1] - k:boff[b + 1]] = 0 if trial & 4 else rng.integers(0, 256, k, dtype=np.uint8)
            else:
                a = int(rng.integers(boff[b], boff[b + 1]))
                bad[a:min(a + int(rng.integers(1, 300)), boff[b + 1])] = np.uint8(rng.integers(0, 256))
        out, st2 = ctx.decompress_blocks(dev(ctx, bad), bad.size, dev(ctx, boff), src.size, bs, 0, n_states)
        st2, outh = st2.cpu().numpy(), out.cpu().numpy()
        assert ((st2 <= 2) & (st2 >= -11)).all()
        for b in range(0, nb, 7):
            if b not in touched:
                assert st2[b] == 0 and np.array_equal(outh[b * bs:(b + 1) * bs], src[b * bs:(b + 1) * bs]), (trial, b)
    out, st = ctx.decompress_blocks(dev(ctx, good), good.size, dev(ctx, offh), src.size, bs, 0, n_states)
    assert not st.cpu().numpy().any() and np.array_equal(out.cpu().numpy(), src)


def test_host_buffer_api_in_the_reference_formats_many_blocks(ctx):
    """the chunked host path with two states: its chunks hold >= 8 192 blocks so that they take the thread-per-stream
    kernels; same bytes and offsets as the device call, and the round trip"""
    n, bs = (96 << 20) + 777, 4096
    src = ctx.generate("geo", 32, n)
    hsrc = src.cpu().numpy()
    d, off, st, total = ctx.compress_blocks(src, bs, 0, 2)
    hd, hoff, hst, htotal = ctx.compress_host(hsrc, bs, 0, 2)
    assert htotal == total and np.array_equal(hoff.astype(np.int64), off.cpu().numpy())
    assert np.array_equal(hd[:htotal], d[:total].cpu().numpy()) and np.array_equal(hst, st.cpu().numpy())
    out, st2 = ctx.decompress_host(hd, htotal, hoff, n, bs, 0, 2)
    assert (st2 >= 0).all() and np.array_equal(out, hsrc)


def test_many_streams_across_encoder_waves(ctx):
    """two states, more blocks than one wave of the thread-per-stream encoder (16 384): bytes of every block vs the oracle"""
    bs, nb = 500, 20000
    src = O.generate("text", 0xC0FFEE20, bs * nb - 123)
    blocks, st, (d, off, total) = gpu_blocks(ctx, src, bs, 0, 2)
    scratch, sizes, status = O.compress_blocks(src, bs, 0, 2, threads=8)
    assert len(blocks) == nb and not status.any() and not st.any()
    for b in range(nb):
        assert blocks[b] == scratch[b, :int(sizes[b])].tobytes(), "block %d differs" % b
    out, dst_ = ctx.decompress_blocks(d, total, dev(ctx, off), src.size, bs, 0, 2)
    assert not dst_.cpu().numpy().any() and np.array_equal(out.cpu().numpy(), src)


@pytest.mark.parametrize("n_states,bs,tl,nb", [(2, 301, 0, 4200), (1, 256, 0, 4500), (2, 200, 12, 4100), (1, 333, 9, 4097)])
def test_many_streams_stay_inside_their_buffers(ctx, n_states, bs, tl, nb):
    """the thread-per-stream kernels (wide and compact shared-memory tables, unaligned block sizes) between guard bytes: the
    compressed stream, the output and the status array sit inside larger buffers whose margins must come back untouched,
    for good streams and for damaged ones (the decoder then runs past the end of its stack and decodes what it finds)"""
    import torch
    G = 4096
    n = nb * bs - 17
    src = O.generate("geo" if n_states == 2 else "text", 40 + n_states + tl, n)
    p = ctx.params(bs, tl, n_states, 0)
    assert ctx.num_streams(n, p) == nb
    cap = ctx.bound(n, p)

    def guarded(nbytes, dtype=torch.uint8, fill=0xA5):
        big = torch.full((nbytes + 2 * G,), fill, dtype=dtype, device=ctx.device)
        return big, big[G:G + nbytes]

    def margins_ok(big, nbytes, fill=0xA5):
        return bool((big[:G] == fill).all()) and bool((big[G + nbytes:] == fill).all())
    big_c, comp = guarded(cap)
    big_o, off = guarded(nb + 1, torch.int64, 0x5A5A)
    big_s, st = guarded(nb, torch.int32, 0x5A5A)
    ctx.compress_blocks_async(dev(ctx, src), p, comp, off, st)
    ctx.sync()
    assert margins_ok(big_c, cap) and margins_ok(big_o, nb + 1, 0x5A5A) and margins_ok(big_s, nb, 0x5A5A)
    assert (st.cpu().numpy() >= 0).all()
    total = int(off[nb].item())
    big_d, out = guarded(n)
    big_t, st2 = guarded(nb, torch.int32, 0x5A5A)
    ctx.decompress_blocks_async(comp, total, off, nb, p, out, n, st2)
    ctx.sync()
    assert margins_ok(big_d, n) and margins_ok(big_t, nb, 0x5A5A) and margins_ok(big_c, cap)
    assert np.array_equal(out.cpu().numpy(), src) and (st2.cpu().numpy() >= 0).all()
    rng = np.random.default_rng(5 + bs)
    bad = comp[:total].cpu().numpy().copy()
    offh = off.cpu().numpy().astype(np.int64)
    for _ in range(400):
        b = int(rng.integers(0, nb))
        lo, hi = int(offh[b]), int(offh[b + 1])
        k = int(rng.integers(1, max(2, hi - lo)))
        if rng.integers(0, 2):
            bad[hi - k:hi] = rng.integers(0, 256, k, dtype=np.uint8)
        else:
            bad[lo:lo + k] = rng.integers(0, 256, k, dtype=np.uint8)
    comp[:total] = dev(ctx, bad)
    out.fill_(0)
    ctx.decompress_blocks_async(comp, total, off, nb, p, out, n, st2)
    ctx.sync()
    st3 = st2.cpu().numpy()
    assert ((st3 <= 2) & (st3 >= -11)).all()
    assert margins_ok(big_d, n) and margins_ok(big_t, nb, 0x5A5A) and margins_ok(big_c, cap)


@pytest.mark.parametrize("seed", range(6))
def test_many_streams_randomised_differential(ctx, seed):
    """seeded shapes through the thread-per-stream kernels (block size, table_log, data kind, one or two states): every
    block against the oracle, and the round trip"""
    rng = np.random.default_rng(900 + seed)
    n_states = int(rng.choice([1, 2]))
    bs = int(rng.integers(130, 2600))
    tl = int(rng.choice([0, 5, 9, 11, 12]))
    kind = str(rng.choice(["geo", "text", "few", "uniform"]))
    nb = 4096 + int(rng.integers(1, 300))
    n = nb * bs - int(rng.integers(0, bs - 1))
    src = O.generate(kind, 1000 + seed, n)
    for _ in range(5):                                        # a few degenerate blocks
        b = int(rng.integers(0, nb - 1))
        src[b * bs:(b + 1) * bs] = int(rng.choice([0, 7, 255]))
    blocks, st, (d, off, total) = gpu_blocks(ctx, src, bs, tl, n_states)
    scratch, sizes, status = O.compress_blocks(src, bs, tl, n_states, threads=8)
    for b in range(nb):
        if status[b] == 0:
            assert st[b] == 0 and blocks[b] == scratch[b, :int(sizes[b])].tobytes(), (seed, b, bs, tl, kind, n_states)
        else:
            assert st[b] != 0
    ok = st >= 0
    out, dst_ = ctx.decompress_blocks(d, total, dev(ctx, off), n, bs, tl, n_states)
    outh, dst_ = out.cpu().numpy(), dst_.cpu().numpy()
    for b in range(nb):
        if ok[b]:
            assert dst_[b] >= 0 and np.array_equal(outh[b * bs:(b + 1) * bs], src[b * bs:(b + 1) * bs]), (seed, b)
