"""Test infrastructure: interoperation with libzstd (the implementation whose FSE the reference crate ports: same NCount
header, spread, tables, two interleaved states, stack-ordered bits, end mark).  libzstd exports no FSE entry point, so it is
used as a black box through the one place where a zstd frame carries a bare two-state FSE stream: the FSE-compressed
weights of a Huffman tree description (RFC 8878, 4.2.1.2).

  zstd -> us : ZSTD_compress2 of literal-only data; the frame's weight blob (NCount header || FSE stream) is decoded by the
               code under test; the Huffman code built from those weights must decode the frame's literal streams to the
               input, each stream ending on its last bit.  The histogram of the decoded weights also gives a real input /
               output pair of libzstd's FSE_normalizeCount (HUF_compressWeights calls it with useLowProbCount = 0).
  us -> zstd : a hand-built frame whose weight blob was written by the code under test; ZSTD_decompress must return the
               literals.
Only tests/ imports this module."""
import ctypes as C
import ctypes.util
import heapq

import numpy as np

MAGIC = bytes.fromhex("28b52ffd")
_Z = None


def zstd():
    """libzstd through ctypes, or None"""
    global _Z
    if _Z is None:
        name = ctypes.util.find_library("zstd") or "libzstd.so.1"
        try:
            z = C.CDLL(name)
        except OSError:
            _Z = False
            return None
        z.ZSTD_versionString.restype = C.c_char_p
        z.ZSTD_createCCtx.restype = C.c_void_p
        z.ZSTD_freeCCtx.argtypes = [C.c_void_p]
        z.ZSTD_CCtx_setParameter.argtypes = [C.c_void_p, C.c_int, C.c_int]
        z.ZSTD_CCtx_setParameter.restype = C.c_size_t
        z.ZSTD_compress2.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        z.ZSTD_compress2.restype = C.c_size_t
        z.ZSTD_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        z.ZSTD_decompress.restype = C.c_size_t
        z.ZSTD_isError.argtypes = [C.c_size_t]
        z.ZSTD_getErrorName.argtypes = [C.c_size_t]
        z.ZSTD_getErrorName.restype = C.c_char_p
        z.ZSTD_compressBound.argtypes = [C.c_size_t]
        z.ZSTD_compressBound.restype = C.c_size_t
        _Z = z
    return _Z or None


def zstd_compress_literals_only(src):
    """level 1, strategy fast, minMatch 7: random bytes of a few hundred symbols' worth of entropy have no 7-byte repeats, so
    the block is one Huffman-coded literals section and zero sequences"""
    z = zstd()
    cctx = z.ZSTD_createCCtx()
    for key, val in ((100, 1), (107, 1), (105, 7)):          # ZSTD_c_compressionLevel, ZSTD_c_strategy, ZSTD_c_minMatch
        assert not z.ZSTD_isError(z.ZSTD_CCtx_setParameter(cctx, key, val))
    src = np.ascontiguousarray(src, dtype=np.uint8)
    cap = z.ZSTD_compressBound(src.size)
    dst = (C.c_uint8 * cap)()
    n = z.ZSTD_compress2(cctx, dst, cap, src.ctypes.data, src.size)
    z.ZSTD_freeCCtx(cctx)
    assert not z.ZSTD_isError(n), z.ZSTD_getErrorName(n)
    return bytes(dst[:n])


def zstd_decompress(frame, cap):
    z = zstd()
    out = (C.c_uint8 * cap)()
    n = z.ZSTD_decompress(out, cap, frame, len(frame))
    if z.ZSTD_isError(n):
        raise ValueError(z.ZSTD_getErrorName(n).decode())
    return bytes(out[:n])


def parse_first_block(frame):
    """-> dict(regen, streams, tree (the Huffman tree description incl. its header byte .. end of literals), sequences)
    of the first block, which must be a compressed block with a Huffman-compressed literals section (RFC 8878 3.1.1)"""
    assert frame[:4] == MAGIC
    fhd = frame[4]
    pos = 5
    fcs_flag, single, dictid = fhd >> 6, (fhd >> 5) & 1, fhd & 3
    if not single:
        pos += 1                                             # window descriptor
    pos += (0, 1, 2, 4)[dictid]
    pos += (1 if single else 0, 2, 4, 8)[fcs_flag]
    bh = int.from_bytes(frame[pos:pos + 3], "little")
    pos += 3
    btype, bsize = (bh >> 1) & 3, bh >> 3
    if btype != 2:
        return None
    blk = frame[pos:pos + bsize]
    ltype, sf = blk[0] & 3, (blk[0] >> 2) & 3
    if ltype != 2:
        return None
    if sf in (0, 1):
        v = int.from_bytes(blk[:3], "little")
        regen, comp, hl, streams = (v >> 4) & 0x3ff, (v >> 14) & 0x3ff, 3, (1 if sf == 0 else 4)
    elif sf == 2:
        v = int.from_bytes(blk[:4], "little")
        regen, comp, hl, streams = (v >> 4) & 0x3fff, (v >> 18) & 0x3fff, 4, 4
    else:
        v = int.from_bytes(blk[:5], "little")
        regen, comp, hl, streams = (v >> 4) & 0x3ffff, (v >> 22) & 0x3ffff, 5, 4
    return dict(regen=regen, streams=streams, tree=blk[hl:hl + comp], sequences=blk[hl + comp:])


def huf_table(weights):
    """weights of symbols 0 .. n-2 -> (decode table [(symbol, nbits)] of 2^maxbits entries, maxbits, all n weights);
    the last weight is implied (RFC 8878 4.2.1.1), codes are assigned from the lowest weight up, symbols in natural order"""
    total = sum((1 << (x - 1)) for x in weights if x)
    assert total > 0
    maxbits = total.bit_length()
    left = (1 << maxbits) - total
    assert left > 0 and left & (left - 1) == 0, (total, left)
    wl = list(weights) + [left.bit_length()]
    table = [None] * (1 << maxbits)
    pos = 0
    for wgt in range(1, maxbits + 1):
        for s, x in enumerate(wl):
            if x == wgt:
                for k in range(1 << (wgt - 1)):
                    table[pos + k] = (s, maxbits + 1 - wgt)
                pos += 1 << (wgt - 1)
    assert pos == 1 << maxbits
    return table, maxbits, wl


def huf_decode_stream(data, table, maxbits, count):
    """one backward bit stream (RFC 8878 4.2.2): `count` symbols, and the stream must end on its last bit"""
    assert data and data[-1] != 0
    v = int.from_bytes(data, "little")
    nbits = v.bit_length() - 1                               # below the end mark
    mask = (1 << maxbits) - 1
    out = bytearray()
    for _ in range(count):
        idx = ((v >> (nbits - maxbits)) if nbits >= maxbits else (v << (maxbits - nbits))) & mask
        s, nb = table[idx]
        out.append(s)
        nbits -= nb
        assert nbits >= 0, "stream exhausted early"
    assert nbits == 0, "stream has %d bits left" % nbits
    return bytes(out)


def huf_decode_literals(info, weights):
    table, maxbits, _ = huf_table(weights)
    h = info["tree"][0]
    body = info["tree"][1 + h:]
    regen = info["regen"]
    if info["streams"] == 4:
        s1, s2, s3 = (int.from_bytes(body[i:i + 2], "little") for i in (0, 2, 4))
        segs = [body[6:6 + s1], body[6 + s1:6 + s1 + s2], body[6 + s1 + s2:6 + s1 + s2 + s3], body[6 + s1 + s2 + s3:]]
        per = (regen + 3) // 4
        counts = [per, per, per, regen - 3 * per]
    else:
        segs, counts = [body], [regen]
    return b"".join(huf_decode_stream(sg, table, maxbits, c) for sg, c in zip(segs, counts))


def huf_lengths(counts, limit=11):
    """Huffman code lengths of the present symbols, at most `limit` bits (counts are halved until they fit)"""
    while True:
        heap = [(c, i, (i,)) for i, c in enumerate(counts) if c]
        assert len(heap) >= 2
        heapq.heapify(heap)
        length = {i: 0 for _, i, _ in heap}
        tie = 1000
        while len(heap) > 1:
            a, b = heapq.heappop(heap), heapq.heappop(heap)
            for s in a[2] + b[2]:
                length[s] += 1
            heapq.heappush(heap, (a[0] + b[0], tie, a[2] + b[2]))
            tie += 1
        if max(length.values()) <= limit:
            return length
        counts = [(c + 1) // 2 if c else 0 for c in counts]


def weights_for(lits):
    """-> (weights of symbols 0 .. last-1 as bytes: what gets FSE-compressed, maxbits)"""
    length = huf_lengths(np.bincount(np.frombuffer(lits, dtype=np.uint8), minlength=256).tolist())
    maxbits = max(length.values())
    wts = [(maxbits + 1 - length[s]) if s in length else 0 for s in range(256)]
    last = max(length)
    table, mb, wl = huf_table(wts[:last])
    assert mb == maxbits and wl[last] == wts[last]
    return bytes(wts[:last]), maxbits


def craft_frame(lits, weight_blob, weights):
    """a single-segment frame with one compressed block: Huffman literals (one stream, 256 <= len(lits) < 1024) whose tree
    description carries `weight_blob` (NCount header || two-state FSE stream of `weights`), and no sequences"""
    table, maxbits, _ = huf_table(list(weights))
    first = {}
    for idx, (s, nb) in enumerate(table):
        first.setdefault(s, (idx, nb))
    v = 1
    for s in lits:
        idx, nb = first[s]
        v = (v << nb) | (idx >> (maxbits - nb))
    stream = v.to_bytes((v.bit_length() + 7) // 8, "little")
    assert len(weight_blob) < 128
    comp, regen = 1 + len(weight_blob) + len(stream), len(lits)
    assert 256 <= regen < 1024 and comp < 1024
    lit_header = (2 | (regen << 4) | (comp << 14)).to_bytes(3, "little")          # Compressed_Literals_Block, size format 0
    block = lit_header + bytes([len(weight_blob)]) + weight_blob + stream + b"\x00"   # + Number_of_Sequences = 0
    block_header = (1 | (2 << 1) | (len(block) << 3)).to_bytes(3, "little")        # last block, compressed
    return MAGIC + bytes([0x60]) + (regen - 256).to_bytes(2, "little") + block_header + block


def skewed_bytes(seed, n, nsym, decay, base=0):
    rng = np.random.default_rng(seed)
    p = decay ** np.arange(nsym)
    return (rng.choice(nsym, size=n, p=p / p.sum()).astype(np.uint8) + base).tobytes()


# ---------------------------------------------------------------------------------- sequences sections (RFC 8878 3.1.1.3)
# A compressed block's sequences are three FSE-coded symbol streams (literal-length, offset and match-length codes) over
# tables of up to 512 entries whose NCount headers libzstd writes with FSE_normalizeCount(.., useLowProbCount = nbSeq >= 2048):
# larger tables than the Huffman weights reach, and -1 counts written by libzstd itself.  The three states share one
# backward bit stream with the sequences' extra bits, so decoding it needs exactly right tables: the check is that the
# stream ends on its last bit and that literal lengths + match lengths + trailing literals make the block size.
LL_BITS = [0] * 16 + [1, 1, 1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16]
LL_BASE = list(range(16)) + [16, 18, 20, 22, 24, 28, 32, 40, 48, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536]
ML_BITS = [0] * 32 + [1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16]
ML_BASE = [i + 3 for i in range(32)] + [35, 37, 39, 41, 43, 47, 51, 59, 67, 83, 99, 131, 259, 515, 1027, 2051, 4099, 8195,
                                        16387, 32771, 65539]
PREDEFINED = {      # RFC 8878 3.1.1.3.2.2: (normalised counts, accuracy log)
    "LL": ([4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1], 6),
    "OF": ([1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1], 5),
    "ML": ([1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
            1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1], 6),
}


def zstd_compress(src, level):
    z = zstd()
    cctx = z.ZSTD_createCCtx()
    assert not z.ZSTD_isError(z.ZSTD_CCtx_setParameter(cctx, 100, level))
    cap = z.ZSTD_compressBound(len(src))
    dst = (C.c_uint8 * cap)()
    n = z.ZSTD_compress2(cctx, dst, cap, src, len(src))
    z.ZSTD_freeCCtx(cctx)
    assert not z.ZSTD_isError(n)
    return bytes(dst[:n])


def wordy_bytes(seed, size):
    """space-separated words drawn (Zipf) from a random dictionary: thousands of matches of many lengths and offsets"""
    rng = np.random.default_rng(seed)
    words = [bytes(rng.integers(97, 123, size=int(rng.integers(2, 12))).astype(np.uint8)) for _ in range(3000)]
    idx = rng.zipf(1.3, size=size // 3 + 16) % len(words)
    return b" ".join(words[i] for i in idx)[:size]


def single_block(frame):
    """-> (frame content size, the bytes of the first block), which must be the last one and compressed"""
    assert frame[:4] == MAGIC
    fhd = frame[4]
    pos = 5
    fcs_flag, single, dictid = fhd >> 6, (fhd >> 5) & 1, fhd & 3
    if not single:
        pos += 1
    pos += (0, 1, 2, 4)[dictid]
    nb = (1 if single else 0, 2, 4, 8)[fcs_flag]
    fcs = int.from_bytes(frame[pos:pos + nb], "little") + (256 if nb == 2 else 0)
    pos += nb
    bh = int.from_bytes(frame[pos:pos + 3], "little")
    pos += 3
    assert bh & 1 and (bh >> 1) & 3 == 2
    return fcs, frame[pos:pos + (bh >> 3)]


def literals_section(blk):
    """-> (regenerated size, bytes of the whole literals section)"""
    ltype, sf = blk[0] & 3, (blk[0] >> 2) & 3
    if ltype in (0, 1):
        if sf in (0, 2):
            regen, hl = blk[0] >> 3, 1
        elif sf == 1:
            regen, hl = int.from_bytes(blk[:2], "little") >> 4, 2
        else:
            regen, hl = int.from_bytes(blk[:3], "little") >> 4, 3
        return regen, hl + (regen if ltype == 0 else 1)
    if sf in (0, 1):
        v = int.from_bytes(blk[:3], "little")
        regen, comp, hl = (v >> 4) & 0x3ff, (v >> 14) & 0x3ff, 3
    elif sf == 2:
        v = int.from_bytes(blk[:4], "little")
        regen, comp, hl = (v >> 4) & 0x3fff, (v >> 18) & 0x3fff, 4
    else:
        v = int.from_bytes(blk[:5], "little")
        regen, comp, hl = (v >> 4) & 0x3ffff, (v >> 22) & 0x3ffff, 5
    return regen, hl + comp


def decode_sequences(seq, lit_count, block_size, table_of_header, table_of_norm):
    """Walks a sequences section with decode tables supplied by the code under test:
         table_of_header(bytes) -> (entries [(new_state, symbol, num_bits)], table_log, header bytes used, counts)
         table_of_norm(counts, table_log) -> entries          (the predefined distributions)
    -> (number of sequences, [(name, mode, table_log, counts)], {name: codes in sequence order}); asserts that the bit
    stream is used up exactly and that the lengths add up to the block"""
    b0 = seq[0]
    pos = 1
    assert b0 != 0
    if b0 < 128:
        nseq = b0
    elif b0 < 255:
        nseq, pos = ((b0 - 128) << 8) + seq[1], 2
    else:
        nseq, pos = seq[1] + (seq[2] << 8) + 0x7f00, 3
    modes = seq[pos]
    pos += 1
    tabs, info = [], []
    for name, mode in (("LL", modes >> 6), ("OF", (modes >> 4) & 3), ("ML", (modes >> 2) & 3)):
        if mode == 0:
            counts, log2 = PREDEFINED[name]
            tabs.append((table_of_norm(counts, log2), log2))
            info.append((name, "predefined", log2, counts))
        elif mode == 1:
            tabs.append(([(0, seq[pos], 0)], 0))
            info.append((name, "rle", 0, None))
            pos += 1
        elif mode == 2:
            entries, log2, used, counts = table_of_header(bytes(seq[pos:pos + 600]))
            tabs.append((entries, log2))
            info.append((name, "fse", log2, counts))
            pos += used
        else:
            raise AssertionError("repeat mode in a first block")
    data = seq[pos:]
    assert data and data[-1] != 0
    v = int.from_bytes(data, "little")
    nbits = v.bit_length() - 1
    state = {"n": nbits}

    def read(n):
        state["n"] -= n
        assert state["n"] >= 0, "bit stream exhausted"
        return (v >> state["n"]) & ((1 << n) - 1)
    (llt, lll), (oft, ofl), (mlt, mll) = tabs
    sl, so, sm = read(lll), read(ofl), read(mll)
    tot_ll = tot_ml = 0
    codes = {"LL": [], "OF": [], "ML": []}
    for k in range(nseq):
        llc, ofc, mlc = llt[sl][1], oft[so][1], mlt[sm][1]
        codes["LL"].append(llc)
        codes["OF"].append(ofc)
        codes["ML"].append(mlc)
        read(ofc)
        tot_ml += ML_BASE[mlc] + read(ML_BITS[mlc])
        tot_ll += LL_BASE[llc] + read(LL_BITS[llc])
        if k + 1 < nseq:
            sl = llt[sl][0] + read(llt[sl][2])
            sm = mlt[sm][0] + read(mlt[sm][2])
            so = oft[so][0] + read(oft[so][2])
    assert state["n"] == 0, "%d bits left" % state["n"]
    assert tot_ll <= lit_count and tot_ml + lit_count == block_size, (tot_ll, tot_ml, lit_count, block_size)
    return nseq, info, codes


def sequence_code_histogram(codes, nseq):
    """what ZSTD_buildCTable hands to FSE_normalizeCount in set_compressed mode: the code counts with the last sequence's
    code taken out when it is not its only occurrence -> (counts[256], total, useLowProbCount)"""
    cnt = [0] * 256
    for c in codes:
        cnt[c] += 1
    total = nseq
    if cnt[codes[-1]] > 1:
        cnt[codes[-1]] -= 1
        total -= 1
    return cnt, total, total >= 2048
