"""Interoperation with libzstd 1.5.x: the two-state FSE stream format of the reference (`fse_compress2` / `fse_decompress2`,
src/lib.rs:146-248; NCount header src/histogram.rs:376-505; tables src/fse.rs:101-339) is the format of zstd's
FSE_compress_usingCTable / FSE_decompress, which the crate ports.  These tests pin the oracle -- and on a GPU the product --
against a real libzstd in both directions (tests/zstd_interop.py), and `normalize_zstd` against libzstd's own
FSE_normalizeCount on real histograms.  This is not an execution of the crate (no Rust toolchain here): what it pins is
that the restated format is the one the crate says it implements."""
import numpy as np
import pytest

import oracle_lib as O
import zstd_interop as ZI

needs_zstd = pytest.mark.skipif(ZI.zstd() is None, reason="libzstd not available")


def _golden(key="streams"):
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "zstd_weight_streams.json")) as f:
        return json.load(f)[key]


def test_oracle_decodes_the_committed_libzstd_streams():
    """tests/golden/zstd_weight_streams.json (make_zstd_vectors.py): bytes libzstd 1.5.5 wrote, and what they decode to"""
    g = _golden()
    assert len(g) >= 6
    for v in g:
        blob = bytes.fromhex(v["blob"])
        rc, nh, consumed = O.ncount_read(blob)
        assert rc == 0 and nh.log2 == v["table_log"] and consumed == v["header_bytes"]
        assert list(nh.table[:nh.table_len]) == v["norm"]
        assert O.decompress_n_exhaust(blob, 2, 255).hex() == v["weights"]
        counts = np.bincount(np.frombuffer(bytes.fromhex(v["weights"]), dtype=np.uint8), minlength=256).tolist()
        rcz, nz = O.normalize_zstd(O.hist_from_counts(counts), v["table_log"], use_low_prob_count=False)
        assert rcz == 0 and list(nz.table[:len(v["norm"])]) == v["norm"]


@pytest.mark.gpu
def test_gpu_decodes_the_committed_libzstd_streams():
    import entropy_coders_b200 as E
    for v in _golden():
        out = bytearray()
        n = E.fse_decompress2(bytes.fromhex(v["blob"]), out)
        assert n == len(out) and bytes(out).hex() == v["weights"]

CASES = [(1, 4000, 90, 0.93, 20), (2, 6000, 120, 0.95, 0), (3, 3000, 60, 0.90, 64), (4, 12000, 200, 0.97, 10),
         (5, 5000, 150, 0.985, 30), (6, 2500, 40, 0.85, 97)]


def _zstd_weight_blob(case):
    seed, n, nsym, decay, base = case
    src = ZI.skewed_bytes(seed, n, nsym, decay, base)
    info = ZI.parse_first_block(ZI.zstd_compress_literals_only(np.frombuffer(src, dtype=np.uint8)))
    if info is None or info["tree"][0] >= 128 or info["sequences"][:1] != b"\x00" or info["regen"] != len(src):
        return None                                          # raw weights, or not literal-only: nothing to pin in this case
    h = info["tree"][0]
    return src, info, bytes(info["tree"][1:1 + h])


@needs_zstd
def test_oracle_decodes_the_fse_streams_libzstd_writes():
    used = 0
    for case in CASES:
        got = _zstd_weight_blob(case)
        if got is None:
            continue
        src, info, blob = got
        rc, nh, consumed = O.ncount_read(blob)                # NormHistogram::read on FSE_writeNCount's bytes
        assert rc == 0 and 5 <= nh.log2 <= 6 and sum(abs(x) for x in nh.table[:nh.table_len]) == 1 << nh.log2
        weights = O.decompress_n_exhaust(blob, 2, 255)        # fse_decompress2 on FSE_compress_usingCTable's stream
        assert ZI.huf_decode_literals(info, list(weights)) == src
        used += 1
    assert used >= 4


@needs_zstd
def test_normalize_zstd_equals_libzstd_on_real_histograms():
    """HUF_compressWeights: FSE_normalizeCount(norm, tableLog, count, n, maxSymbol, useLowProbCount = 0); the header of the
    weight stream carries its result, the decoded weights give its input"""
    used = 0
    for case in CASES:
        got = _zstd_weight_blob(case)
        if got is None:
            continue
        _, _, blob = got
        rc, nh, _ = O.ncount_read(blob)
        weights = O.decompress_n_exhaust(blob, 2, 255)
        counts = np.bincount(np.frombuffer(weights, dtype=np.uint8), minlength=256).tolist()
        rcz, nz = O.normalize_zstd(O.hist_from_counts(counts), nh.log2, use_low_prob_count=False)
        assert rcz == 0 and list(nz.table[:256]) == list(nh.table[:256]), case
        used += 1
    assert used >= 4


@needs_zstd
@pytest.mark.parametrize("table_log", [5, 6])
@pytest.mark.parametrize("seed", list(range(11, 31)))
def test_libzstd_decodes_the_fse_streams_the_oracle_writes(seed, table_log):
    lits = ZI.skewed_bytes(seed, 700 + 60 * seed % 300, 40 + seed, 0.88 + 0.01 * (seed % 5), 33)
    weights, _ = ZI.weights_for(lits)
    blob, _, _ = O.compress_n(weights, table_log, 2)          # fse_compress2 at table_log 5 / 6 (zstd accepts at most 6 here)
    assert ZI.zstd_decompress(ZI.craft_frame(lits, blob, weights), 2048) == lits


@needs_zstd
def test_the_streams_sent_to_libzstd_cover_low_probability_symbols():
    """the -1 counts (spread from the top of the table, fse.rs:122-125 / FSE_buildDTable's highThreshold) are part of
    what the previous test sends"""
    seen = 0
    for seed in range(11, 31):
        lits = ZI.skewed_bytes(seed, 700 + 60 * seed % 300, 40 + seed, 0.88 + 0.01 * (seed % 5), 33)
        weights, _ = ZI.weights_for(lits)
        rc, nh, _ = O.ncount_read(O.compress_n(weights, 5, 2)[0])
        seen += any(x == -1 for x in nh.table[:nh.table_len])
    assert seen >= 3


@needs_zstd
@pytest.mark.gpu
def test_gpu_decodes_the_fse_streams_libzstd_writes():
    import entropy_coders_b200 as E
    used = 0
    for case in CASES:
        got = _zstd_weight_blob(case)
        if got is None:
            continue
        src, info, blob = got
        out = bytearray()
        n = E.fse_decompress2(blob, out)                      # the crate-shaped front: fse_b200_decompress_exhaust, 2 states
        assert n is not None and n == len(out)
        assert ZI.huf_decode_literals(info, list(out)) == src
        assert bytes(out) == O.decompress_n_exhaust(blob, 2, 255)
        used += 1
    assert used >= 4


@needs_zstd
@pytest.mark.gpu
@pytest.mark.parametrize("table_log", [5, 6])
def test_libzstd_decodes_the_fse_streams_the_gpu_writes(table_log):
    import torch
    import entropy_coders_b200 as E
    ctx = E.Context(0)
    for seed in (21, 22, 23):
        lits = ZI.skewed_bytes(seed, 800, 50 + seed % 7, 0.9, 40)
        weights, _ = ZI.weights_for(lits)
        src = torch.frombuffer(bytearray(weights), dtype=torch.uint8).to(ctx.device)
        d, off, st, total = ctx.compress_blocks(src, len(weights), table_log, 2)      # one block = one fse_compress2 stream
        assert int(st.cpu()[0]) == 0
        blob = d[:total].cpu().numpy().tobytes()
        assert blob == O.compress_n(weights, table_log, 2)[0]
        assert ZI.zstd_decompress(ZI.craft_frame(lits, blob, weights), 2048) == lits
    ctx.close()


@needs_zstd
@pytest.mark.gpu
def test_gpu_normalize_zstd_equals_libzstd_on_real_histograms():
    import entropy_coders_b200 as E
    ctx = E.Context(0)
    used = 0
    for case in CASES:
        got = _zstd_weight_blob(case)
        if got is None:
            continue
        _, _, blob = got
        rc, nh, _ = O.ncount_read(blob)
        weights = O.decompress_n_exhaust(blob, 2, 255)
        import torch
        counts = torch.from_numpy(np.bincount(np.frombuffer(weights, dtype=np.uint8), minlength=256).astype(np.int64)).to(ctx.device)
        norm, log2, tlen, status = ctx.normalize_zstd(counts, nh.log2, use_low_prob_count=False)
        assert int(status.cpu()[0]) == 0 and int(log2.cpu()[0]) == nh.log2
        assert norm.cpu().numpy()[0].tolist() == list(nh.table[:256]), case
        used += 1
    assert used >= 4
    ctx.close()


# ------------------------------------------------------------------------------------ sequences sections: tables up to 512 entries

SEQ_CASES = [(5, 60000, 3), (6, 120000, 5), (7, 100000, 1), (8, 30000, 9), (9, 9000, 3)]


def _oracle_table_of_header(blob):
    rc, nh, used = O.ncount_read(blob)
    assert rc == 0
    t = O.dec_table(nh)
    return ([(t.table[i].new_state, t.table[i].symbol, t.table[i].num_bits) for i in range(1 << nh.log2)], nh.log2, used,
            list(nh.table[:nh.table_len]))


def _oracle_table_of_norm(counts, log2):
    nh = O.norm_from_table(list(counts) + [0] * (256 - len(counts)))
    assert nh.log2 == log2
    t = O.dec_table(nh)
    return [(t.table[i].new_state, t.table[i].symbol, t.table[i].num_bits) for i in range(1 << log2)]


def _walk(case, table_of_header, table_of_norm):
    seed, size, level = case
    fcs, blk = ZI.single_block(ZI.zstd_compress(ZI.wordy_bytes(seed, size), level))
    regen, lsz = ZI.literals_section(blk)
    return ZI.decode_sequences(blk[lsz:], regen, fcs, table_of_header, table_of_norm)


@needs_zstd
def test_oracle_tables_decode_the_sequence_streams_libzstd_writes():
    """NormHistogram::read + DecodeTable::update (spread with -1 symbols, new_state / num_bits) at table_log 7 .. 9 on the
    headers FSE_writeNCount wrote for FSE_buildCTable: the three interleaved state streams of thousands of sequences
    end on the last bit and the lengths add up to the block.  normalize_zstd reproduces every header from the decoded
    codes, with and without low-probability counts."""
    fse_tables = low_prob = 0
    for case in SEQ_CASES:
        nseq, info, codes = _walk(case, _oracle_table_of_header, _oracle_table_of_norm)
        assert nseq > 500
        for name, mode, log2, counts in info:
            if mode != "fse":
                continue
            cnt, total, low = ZI.sequence_code_histogram(codes[name], nseq)
            rc, nz = O.normalize_zstd(O.hist_from_counts(cnt), log2, use_low_prob_count=low)
            assert rc == 0 and list(nz.table[:len(counts)]) == counts, (case, name)
            fse_tables += 1
            low_prob += -1 in counts
    assert fse_tables >= 12 and low_prob >= 5


@needs_zstd
@pytest.mark.gpu
def test_gpu_tables_decode_the_sequence_streams_libzstd_writes():
    """the same walk with headers parsed and decode tables built by the GPU stage entry points
    (fse_b200_ncount_read, fse_b200_build_decode_tables), and fse_b200_normalize_zstd on the decoded codes"""
    import torch
    import entropy_coders_b200 as E
    ctx = E.Context(0)

    def table_of_norm_dev(norm, log2, tlen):
        table, st = ctx.build_decode_tables(norm, log2, tlen, 9)
        assert int(st.cpu()[0]) == 0
        t = table.cpu().numpy().view(np.uint32)[0]
        n = 1 << int(log2.cpu()[0])
        return [(int(e) & 0xffff, (int(e) >> 16) & 0xff, int(e) >> 24) for e in t[:n]]

    def table_of_header(blob):
        rows = torch.zeros((1, 640), dtype=torch.uint8, device=ctx.device)
        rows[0, :len(blob)] = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(ctx.device)
        lens = torch.tensor([len(blob)], dtype=torch.int32, device=ctx.device)
        norm, log2, tlen, cons, st = ctx.ncount_read(rows, lens)
        assert int(st.cpu()[0]) == 0
        k = int(tlen.cpu()[0])
        return table_of_norm_dev(norm, log2, tlen), int(log2.cpu()[0]), int(cons.cpu()[0]), norm.cpu().numpy()[0, :k].tolist()

    def table_of_norm(counts, log2):
        norm = torch.zeros((1, 256), dtype=torch.int32, device=ctx.device)
        norm[0, :len(counts)] = torch.tensor(counts, dtype=torch.int32)
        return table_of_norm_dev(norm, torch.tensor([log2], dtype=torch.int32, device=ctx.device),
                                 torch.tensor([len(counts)], dtype=torch.int32, device=ctx.device))
    checked = 0
    for case in SEQ_CASES[:3] + SEQ_CASES[4:]:
        nseq, info, codes = _walk(case, table_of_header, table_of_norm)
        for name, mode, log2, counts in info:
            if mode != "fse":
                continue
            cnt, total, low = ZI.sequence_code_histogram(codes[name], nseq)
            norm, l2, tlen, status = ctx.normalize_zstd(torch.tensor(cnt, dtype=torch.int64, device=ctx.device), log2, use_low_prob_count=low)
            assert int(status.cpu()[0]) == 0 and norm.cpu().numpy()[0, :len(counts)].tolist() == counts, (case, name)
            checked += 1
    assert checked >= 9
    ctx.close()


def test_oracle_tables_decode_the_committed_libzstd_sequence_frames():
    """the same walk on frames kept under tests/golden (no libzstd needed)"""
    for v in _golden("sequence_frames"):
        fcs, blk = ZI.single_block(bytes.fromhex(v["frame"]))
        regen, lsz = ZI.literals_section(blk)
        nseq, info, codes = ZI.decode_sequences(blk[lsz:], regen, fcs, _oracle_table_of_header, _oracle_table_of_norm)
        assert nseq > 300 and any(mode == "fse" for _, mode, _, _ in info)
        for name, mode, log2, counts in info:
            if mode == "fse":
                cnt, total, low = ZI.sequence_code_histogram(codes[name], nseq)
                rc, nz = O.normalize_zstd(O.hist_from_counts(cnt), log2, use_low_prob_count=low)
                assert rc == 0 and list(nz.table[:len(counts)]) == counts
