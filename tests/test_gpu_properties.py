"""Hypothesis-driven differential tests, CUDA path (through the C ABI) against the CPU oracle:
generated alphabets, skews, runs, block sizes, state counts and table_log; every block is either the oracle's bytes
or one the reference panics on (then it carries an escape), and decode always returns the input."""
import os

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import oracle_lib as O
from test_gpu_parity import dev, gpu_blocks
from test_oracle_properties import byte_strings

pytestmark = pytest.mark.gpu

SCALE = int(os.environ.get("FSE_HYP_SCALE", "1"))       # FSE_HYP_SCALE=20 for a long soak
COMMON = dict(deadline=None, derandomize=SCALE == 1, suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large,
                                                    HealthCheck.function_scoped_fixture])


@pytest.fixture(scope="module")
def ctx():
    import entropy_coders_b200 as E
    c = E.Context(0)
    yield c
    c.close()


def gpu_blocks_at(ctx, src, block_size, table_log, n_states, shift):
    """gpu_blocks with the source at byte offset `shift` of its allocation (unaligned vector-load paths)"""
    import torch
    buf = torch.empty(src.size + 16, dtype=torch.uint8, device=ctx.device)
    view = buf[shift:shift + src.size]
    view.copy_(torch.from_numpy(np.ascontiguousarray(src)))
    d, off, st, total = ctx.compress_blocks(view, block_size, table_log, n_states)
    off = off.cpu().numpy().astype(np.int64)
    raw = d[:total].cpu().numpy().tobytes()
    assert off[0] == 0 and off[-1] == total
    return [raw[off[i]:off[i + 1]] for i in range(len(off) - 1)], st.cpu().numpy(), (d, off, total)


@settings(max_examples=120 * SCALE, **COMMON)
@given(byte_strings(min_size=1, max_size=40000), st.sampled_from([1, 2, 4, 8, 16, 32, 64, 128]),
       st.sampled_from([130, 257, 1000, 4096, 5000, 20000]), st.sampled_from([0, 0, 5, 9, 11, 12]), st.integers(0, 15))
def test_blocks_equal_the_oracle_and_round_trip(ctx, data, n_states, bs, tl, shift):
    if n_states <= 2:
        data = data[:6000]                                   # the one-lane paths are for parity, not speed
    blocks, stat, (d, off, total) = gpu_blocks_at(ctx, data, bs, tl, n_states, shift)
    for b, g in enumerate(blocks):
        blk = data[b * bs:(b + 1) * bs]
        try:
            e = O.compress_n(blk, tl, n_states)[0]
        except ValueError:
            e = None
        if e is None:
            assert stat[b] in (1, 2) and g[0] in (0x0E, 0x0F), (b, stat[b])
        else:
            assert stat[b] == 0 and g == e, (b, n_states, bs, tl)
    out, dst_ = ctx.decompress_blocks(d, total, dev(ctx, off), data.size, bs, tl, n_states)
    assert (dst_.cpu().numpy() >= 0).all() and np.array_equal(out.cpu().numpy(), data)


@settings(max_examples=60 * SCALE, **COMMON)
@given(byte_strings(min_size=300, max_size=30000), st.sampled_from([32, 64, 128]), st.sampled_from([9, 11, 12]))
def test_stage_outputs_equal_the_oracle(ctx, data, n_states, tl):
    """histogram -> normalise -> header -> tables through the stage entry points, one table"""
    h = O.histogram(data)
    if h.table_len <= 1:
        return
    rc, n = O.normalize(h, tl)
    if rc < 0:
        return
    import torch
    counts, table_len = ctx.histogram_blocks(dev(ctx, data), data.size)
    assert np.array_equal(counts.cpu().numpy().view(np.uint32)[0], np.array(h.table[:256], dtype=np.uint32))
    assert int(table_len.cpu()[0]) == h.table_len
    norm, log2, tlen, status = ctx.normalize(counts.to(torch.int64) & 0xffffffff, tl)
    assert int(status.cpu()[0]) == rc and int(log2.cpu()[0]) == n.log2 and int(tlen.cpu()[0]) == n.table_len
    assert np.array_equal(norm.cpu().numpy()[0], np.array(n.table[:256], dtype=np.int32))
    hdr, bits = O.ncount_write(n)
    out, nbytes, nbits = ctx.ncount_write(norm, log2, tlen)
    assert int(nbits.cpu()[0]) == bits and out.cpu().numpy()[0, :int(nbytes.cpu()[0])].tobytes() == hdr
    size = 1 << n.log2
    table, tt, sym, est = ctx.build_encode_tables(norm, log2, tlen, n.log2)
    dtab, dst_ = ctx.build_decode_tables(norm, log2, tlen, n.log2)
    assert int(est.cpu()[0]) == 0 and int(dst_.cpu()[0]) == 0
    et, dt = O.enc_table(n), O.dec_table(n)
    assert np.array_equal(sym.cpu().numpy()[0, :size], np.frombuffer(bytes(et.symbols)[:size], np.uint8))
    assert np.array_equal(table.cpu().numpy().view(np.uint16)[0, :size], np.array(et.table[:size], dtype=np.uint16))
    exp_tt = np.array([[et.symbol_tt[i].bits, et.symbol_tt[i].find_state & 0xFFFFFFFF] for i in range(256)], dtype=np.uint32)
    assert np.array_equal(tt.cpu().numpy()[0].view(np.uint32), exp_tt)
    exp_d = np.array([dt.table[i].new_state | (dt.table[i].symbol << 16) | (dt.table[i].num_bits << 24) for i in range(size)],
                     dtype=np.uint32)
    assert np.array_equal(dtab.cpu().numpy().view(np.uint32)[0, :size], exp_d)


@settings(max_examples=40 * SCALE, **COMMON)
@given(byte_strings(min_size=1, max_size=60000), st.sampled_from([1, 2, 32, 64, 128]), st.sampled_from([300, 4096, 20000]),
       st.sampled_from([0, 9, 11]), st.sampled_from([0, 1]))
def test_frames_and_host_buffers_round_trip(ctx, data, n_states, bs, tl, mode):
    """fse_b200_frame_* and fse_b200_compress_host / decompress_host, per-block and global tables: what goes in comes
    out, and in per-block mode the host path's bytes are the device path's"""
    if n_states <= 2:
        data = data[:5000]
    import entropy_coders_b200 as E
    if mode == 1 and (O.histogram(data).table_len <= 1 or data.size < 8):
        return                                               # one global table needs two symbols (histogram.rs:98)
    try:
        frame = ctx.frame_compress(data, bs, tl, n_states, mode)
    except E.FseError as e:
        assert mode == 1, e                                  # per-block mode never fails: escapes cover every block
        return
    info = ctx.frame_info(frame)
    assert info["n"] == data.size and info["n_states"] == n_states and info["table_mode"] == mode
    assert np.array_equal(ctx.frame_decompress(frame), data)
    if mode == 0:
        dst, off, stat, total = ctx.compress_host(data, bs, tl, n_states)
        blocks, stat_d, _ = gpu_blocks(ctx, data, bs, tl, n_states)
        assert dst[:total].tobytes() == b"".join(blocks) and np.array_equal(stat[:len(blocks)], stat_d)
        out, st2 = ctx.decompress_host(dst, total, off, data.size, bs, tl, n_states)
        assert (st2[:len(blocks)] >= 0).all() and np.array_equal(out[:data.size], data)


@settings(max_examples=60 * SCALE, **COMMON)
@given(byte_strings(min_size=2, max_size=3000), st.sampled_from([1, 2]))
def test_reference_termination_rule(ctx, data, n_states):
    """fse_decompress / fse_decompress2 (lib.rs:187-248) stop on bit exhaustion: the exhaust-mode kernel yields what the
    oracle's literal restatement of that loop yields, surplus symbols included"""
    try:
        comp = O.compress_n(data, 0, n_states)[0]
    except ValueError:
        return
    cap = 4 * data.size + 64
    try:
        exp = O.decompress_n_exhaust(comp, n_states, cap)
    except ValueError:
        exp = None
    off = np.array([0, len(comp)], np.int64)
    out, out_len, st = ctx.decompress_exhaust(dev(ctx, np.frombuffer(comp, np.uint8).copy()), len(comp), dev(ctx, off), 1, cap, 15, n_states)
    if exp is None or len(exp) > cap:
        assert int(st.cpu()[0]) == -2
    else:
        assert int(st.cpu()[0]) == 0 and int(out_len.cpu()[0]) == len(exp)
        assert out.cpu().numpy()[0, :len(exp)].tobytes() == exp and exp[:data.size] == data.tobytes()


@pytest.mark.parametrize("n_states,tl", [(1, 0), (32, 0), (64, 0), (128, 0), (128, 9), (128, 13)])
def test_corrupted_streams_never_fault(ctx, n_states, tl):
    """seeded fuzz of the decoders: bit flips, byte smashes, truncations and offset damage.  Every block ends with a
    status (an error, or 0 with whatever bytes the damaged stream describes), the launch itself never fails, and an intact
    stream still decodes afterwards (compute-sanitizer is closed on this pool; this is the memory-safety net).  128 states:
    table_log 0 takes the compact-table decoder, 9 and 13 the wide-entry one."""
    rng = np.random.default_rng(1000 + n_states + tl)
    bs, nb = (2048 if n_states == 1 else 8192), 6
    src = O.generate("text", 77, bs * nb)
    d, off, st, total = ctx.compress_blocks(dev(ctx, src), bs, tl, n_states)
    good = d[:total].cpu().numpy().copy()
    offh = off.cpu().numpy().astype(np.int64)
    for trial in range(150):
        bad, boff = good.copy(), offh.copy()
        kind = trial % 5
        if kind == 0:                                        # single bit flips, anywhere
            for _ in range(int(rng.integers(1, 6))):
                bad[int(rng.integers(0, bad.size))] ^= np.uint8(1 << int(rng.integers(0, 8)))
        elif kind == 1:                                      # a header smashed with random bytes
            b = int(rng.integers(0, nb))
            k = int(rng.integers(1, 40))
            bad[boff[b]:boff[b] + k] = rng.integers(0, 256, k, dtype=np.uint8)
        elif kind == 2:                                      # payload tail zeroed (marker lost) or randomised
            b = int(rng.integers(0, nb))
            k = int(rng.integers(1, 64))
            bad[boff[b + 1] - k:boff[b + 1]] = 0 if trial % 2 else rng.integers(0, 256, k, dtype=np.uint8)
        elif kind == 3:                                      # block boundaries moved (still monotone)
            b = int(rng.integers(1, nb))
            boff[b] = int(np.clip(boff[b] + rng.integers(-300, 300), boff[b - 1], boff[b + 1]))
        else:                                                # long run of one value inside a block
            b = int(rng.integers(0, nb))
            a = int(rng.integers(boff[b], boff[b + 1]))
            bad[a:min(a + int(rng.integers(1, 2000)), boff[b + 1])] = np.uint8(rng.integers(0, 256))
        out, st2 = ctx.decompress_blocks(dev(ctx, bad), bad.size, dev(ctx, boff), src.size, bs, tl, n_states)
        st2 = st2.cpu().numpy()
        assert st2.shape[0] == nb and ((st2 <= 2) & (st2 >= -11)).all(), (trial, st2)
        if kind in (0, 1, 2, 4):                             # blocks that were not touched still decode exactly
            outh = out.cpu().numpy()
            for b in range(nb):
                if np.array_equal(bad[boff[b]:boff[b + 1]], good[offh[b]:offh[b + 1]]):
                    assert st2[b] == 0 and np.array_equal(outh[b * bs:(b + 1) * bs], src[b * bs:(b + 1) * bs]), (trial, b)
    out, st3 = ctx.decompress_blocks(dev(ctx, good), good.size, dev(ctx, offh), src.size, bs, tl, n_states)
    assert not st3.cpu().numpy().any() and np.array_equal(out.cpu().numpy(), src)


def test_corrupted_streams_exhaust_mode_never_fault(ctx):
    """the same fuzz through the reference-termination decoder (fse_b200_decompress_exhaust): bounded by the capacity"""
    rng = np.random.default_rng(4242)
    srcs = [O.generate(k, 5 + i, n).tobytes() for i, (k, n) in enumerate([("text", 3000), ("geo", 2000), ("few", 1500), ("uniform", 1024)])]
    comps = [O.compress_n(s, 0, 2)[0] for s in srcs]
    cap = 4 * 3000 + 64
    for trial in range(120):
        streams = [bytearray(c) for c in comps]
        for s in streams:
            for _ in range(int(rng.integers(0, 4))):
                s[int(rng.integers(0, len(s)))] ^= 1 << int(rng.integers(0, 8))
            if trial % 3 == 0:
                del s[int(rng.integers(1, len(s))):]
        off = np.zeros(len(streams) + 1, np.int64)
        off[1:] = np.cumsum([len(s) for s in streams])
        comp = np.frombuffer(b"".join(bytes(s) for s in streams), np.uint8).copy()
        out, out_len, st = ctx.decompress_exhaust(dev(ctx, comp), comp.size, dev(ctx, off), len(streams), cap, 15, 2)
        st, out_len = st.cpu().numpy(), out_len.cpu().numpy()
        assert ((st <= 0) & (st >= -11)).all() and (out_len <= cap).all(), (trial, st, out_len)
        for i, s in enumerate(streams):                      # the oracle's literal loop agrees on every damaged stream
            try:
                exp = O.decompress_n_exhaust(bytes(s), 2, cap)
            except ValueError:
                exp = None
            if exp is not None and st[i] == 0:
                assert out_len[i] == len(exp) and out.cpu().numpy()[i, :len(exp)].tobytes() == exp, (trial, i)


@settings(max_examples=60 * SCALE, **COMMON)
@given(byte_strings(min_size=1, max_size=60000), st.sampled_from([(4096, 512), (8192, 1024), (16384, 2048), (32768, 8192), (2048, 2048)]),
       st.sampled_from([0, 0, 5, 9, 11]), st.integers(0, 15))
def test_segmented_mode_equals_the_oracle_composition(ctx, data, shape, tl, shift):
    """segment_size > 0 (CTA-owned replicated tables, builder warp + coder warps): generated data, block / segment shapes,
    table_log and misaligned sources; every stream equals the oracle's composition and decode returns the input"""
    import torch
    from test_gpu_parity import oracle_segments
    bs, seg = shape
    buf = torch.empty(data.size + 16, dtype=torch.uint8, device=ctx.device)
    view = buf[shift:shift + data.size]
    view.copy_(torch.from_numpy(np.ascontiguousarray(data)))
    try:
        exp, est = oracle_segments(data, bs, seg, tl)
    except AssertionError:
        return                                               # a block the oracle's normalise rejects for this table_log
    d, off, st_, total = ctx.compress_blocks(view, bs, tl, 128, segment_size=seg)
    offh = off.cpu().numpy().astype(np.int64)
    raw = d[:total].cpu().numpy().tobytes()
    sth = st_.cpu().numpy()
    assert len(offh) - 1 == len(exp)
    for s_, e in enumerate(exp):
        if est[s_] in (0, 1, 2) and sth[s_] == est[s_]:
            assert raw[offh[s_]:offh[s_ + 1]] == e, (s_, bs, seg, tl)
        else:
            assert sth[s_] == est[s_], (s_, sth[s_], est[s_])
    out, dst_ = ctx.decompress_blocks(d, total, off, data.size, bs, tl, 128, segment_size=seg)
    assert (dst_.cpu().numpy() >= 0).all() and np.array_equal(out.cpu().numpy(), data)


@settings(max_examples=60 * SCALE, **COMMON)
@given(byte_strings(min_size=200, max_size=60000), st.sampled_from([130, 1000, 4096, 20000]), st.sampled_from([0, 5, 8, 10, 11]),
       st.integers(0, 15))
def test_global_table_mode_equals_the_oracle(ctx, data, bs, tl, shift):
    """one table for the whole input (CTA-owned bank-replicated tables for table_log <= 11, 32 / 16 copies): header and every
    header-less block equal the oracle's, any alignment, ragged tail"""
    import torch
    h = O.histogram(data)
    if h.table_len <= 1:
        return
    if tl == 0:
        rc, tle = O.optimal_log2(h)
        if rc < 0:
            return
    else:
        tle = tl
    rc, nh = O.normalize(h, tle)
    if rc < 0 or nh.log2 > 11:
        return
    buf = torch.empty(data.size + 16, dtype=torch.uint8, device=ctx.device)
    view = buf[shift:shift + data.size]
    view.copy_(torch.from_numpy(np.ascontiguousarray(data)))
    header, log2 = ctx.set_global_table(ctx.histogram_global(view), tl)
    assert log2 == nh.log2 and header == O.ncount_write(nh)[0]
    d, off, st_, total = ctx.compress_blocks(view, bs, tl, 128, table_mode=1)
    offh = off.cpu().numpy().astype(np.int64)
    raw = d[:total].cpu().numpy().tobytes()
    et = O.enc_table(nh)
    for b in range(len(offh) - 1):
        blk = data[b * bs:(b + 1) * bs]
        e = blk.tobytes() if len(blk) < 128 else O.encode_payload(et, blk, 128)[0]
        assert raw[offh[b]:offh[b + 1]] == e, (b, bs, tl)
    out, dst_ = ctx.decompress_blocks(d, total, off, data.size, bs, tl, 128, table_mode=1)
    assert (dst_.cpu().numpy() >= 0).all() and np.array_equal(out.cpu().numpy(), data)


@pytest.mark.parametrize("mode", ["global", "segmented"])
def test_corrupted_streams_never_fault_shared_tables(ctx, mode):
    """the seeded fuzz of test_corrupted_streams_never_fault through the CTA-owned-table decoders (global table; segmented
    per-block mode with its builder warp): every stream ends with a status, nothing faults, untouched streams decode, and
    an intact input still round-trips afterwards"""
    rng = np.random.default_rng(99 + len(mode))
    bs, nb = 8192, 6
    src = O.generate("text", 78, bs * nb)
    kw = dict(table_mode=1) if mode == "global" else dict(segment_size=1024)
    tl = 11 if mode == "global" else 0
    if mode == "global":
        ctx.set_global_table(ctx.histogram_global(dev(ctx, src)), tl)
    d, off, st, total = ctx.compress_blocks(dev(ctx, src), bs, tl, 128, **kw)
    good = d[:total].cpu().numpy().copy()
    offh = off.cpu().numpy().astype(np.int64)
    ns = len(offh) - 1
    unit = bs if mode == "global" else 1024
    for trial in range(150):
        bad, boff = good.copy(), offh.copy()
        kind = trial % 5
        if kind == 0:
            for _ in range(int(rng.integers(1, 6))):
                bad[int(rng.integers(0, bad.size))] ^= np.uint8(1 << int(rng.integers(0, 8)))
        elif kind == 1:
            b = int(rng.integers(0, ns))
            k = int(rng.integers(1, 40))
            bad[boff[b]:boff[b] + k] = rng.integers(0, 256, k, dtype=np.uint8)[:max(0, min(k, bad.size - boff[b]))]
        elif kind == 2:
            b = int(rng.integers(0, ns))
            k = int(min(rng.integers(1, 64), boff[b + 1] - boff[b]))
            if k:
                bad[boff[b + 1] - k:boff[b + 1]] = 0 if trial % 2 else rng.integers(0, 256, k, dtype=np.uint8)
        elif kind == 3:
            b = int(rng.integers(1, ns))
            boff[b] = int(np.clip(boff[b] + rng.integers(-300, 300), boff[b - 1], boff[b + 1]))
        else:
            b = int(rng.integers(0, ns))
            a = int(rng.integers(boff[b], boff[b + 1]))
            bad[a:min(a + int(rng.integers(1, 2000)), boff[b + 1])] = np.uint8(rng.integers(0, 256))
        out, st2 = ctx.decompress_blocks(dev(ctx, bad), bad.size, dev(ctx, boff), src.size, bs, tl, 128, **kw)
        st2 = st2.cpu().numpy()
        assert st2.shape[0] == ns and ((st2 <= 2) & (st2 >= -11)).all(), (trial, st2)
        if mode == "global" and kind in (0, 1, 2, 4):        # (in segmented mode a damaged header takes its whole block along)
            outh = out.cpu().numpy()
            for b in range(ns):
                if np.array_equal(bad[boff[b]:boff[b + 1]], good[offh[b]:offh[b + 1]]):
                    assert st2[b] == 0 and np.array_equal(outh[b * unit:(b + 1) * unit], src[b * unit:(b + 1) * unit]), (trial, b)
    out, st3 = ctx.decompress_blocks(dev(ctx, good), good.size, dev(ctx, offh), src.size, bs, tl, 128, **kw)
    assert not st3.cpu().numpy().any() and np.array_equal(out.cpu().numpy(), src)
