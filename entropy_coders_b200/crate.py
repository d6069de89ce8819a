"""Host-side mirror of the reference crate's public API for the FSE path (src/lib.rs:7, :112-248),
backed by the CUDA library: same names, argument meaning and error behaviour, so the parity
tests read like the crate's own tests.

What is mirrored: `Histogram`, `NormHistogram` (+ `HistError`), `fse.EncodeTable`,
`fse.DecodeTable`, `fse_compress`, `fse_compress2`, `fse_decompress`, `fse_decompress2`.
The crate's per-symbol objects (`fse::Encoder`, `fse::Decoder`, `bitstream::*`) are fine-grained
host objects; their arithmetic lives inside the encode / decode kernels and they stay the crate's
own code on the Rust side.

Where the crate panics this raises `Panic`; where it returns None / Err it returns None / raises
`HistError`.
"""
import ctypes as C

import numpy as np

from ._capi import FseError

TABLE_LOG_MIN, TABLE_LOG_MAX, TABLE_LOG_DEFAULT = 5, 15, 11


class Panic(RuntimeError):
    """The reference would panic here (assert / unwrap / ilog2(0))."""


class HistError(Exception):
    """src/histogram.rs:538-546"""
    TableLogTooLarge, TooManySymbols, Io = "TableLogTooLarge", "TooManySymbols", "Io"

    def __init__(self, kind):
        self.kind = kind
        super().__init__(kind)


def _ctx():
    from . import default_context
    return default_context()


def _dev_u8(data):
    import torch
    ctx = _ctx()
    arr = np.frombuffer(bytes(data), dtype=np.uint8)
    if arr.size == 0:
        return torch.empty(0, dtype=torch.uint8, device=ctx.device)
    return torch.from_numpy(arr.copy()).to(ctx.device)


def _ilog2(v):
    if v <= 0:
        raise Panic("ilog2 of zero")
    return v.bit_length() - 1


class Histogram:
    """src/histogram.rs:10-91"""

    def __init__(self, data):
        data = bytes(data)
        if len(data) > 0xFFFFFFFF:
            raise Panic("Data vector is too long")           # :19
        ctx = _ctx()
        if len(data) == 0:
            self._table = [0] * 256
            self._table_len = 1
        else:
            counts, tlen = ctx.histogram_blocks(_dev_u8(data), len(data))
            self._table = [int(x) & 0xFFFFFFFF for x in counts[0].cpu().tolist()]
            self._table_len = int(tlen[0].item())
        self._size = len(data)

    def table(self):
        return list(self._table)

    def table_iter(self):
        return iter(self._table[: self._table_len])

    def symbol_count(self):                                   # :79-81 counts the zero entries (quirk Q3)
        return sum(1 for x in self._table if x == 0)

    def table_len(self):
        return self._table_len

    def size(self):
        return self._size

    def optimal_log2(self):                                   # :264-277, two scalars
        min_bits = min(_ilog2(self._size) + 1, _ilog2(self._table_len - 1) + 2)
        max_bits = _ilog2(self._size - 1) - 2
        if max_bits < 0:
            raise Panic("attempt to subtract with overflow")
        return max(TABLE_LOG_MIN, min(TABLE_LOG_MAX, max(min(TABLE_LOG_DEFAULT, max_bits), min_bits)))

    def normalize(self, log2):                                # :95-155 (+ :157-261)
        import torch
        if self._table_len <= 1 or self._size == 0:
            raise Panic("ilog2 of zero")                      # :98 / :103
        ctx = _ctx()
        c64 = torch.tensor(self._table, dtype=torch.int64, device=ctx.device).reshape(1, 256)
        log2 = max(TABLE_LOG_MIN, min(TABLE_LOG_MAX, int(log2)))
        norm, l2, tlen, st = ctx.normalize(c64, log2)
        if int(st[0].item()) < 0:
            raise Panic("normalize failed: status %d" % int(st[0].item()))
        return NormHistogram(norm[0].cpu().tolist(), int(l2[0].item()), int(tlen[0].item()))

    def normalize_optimal(self):                              # :281-284
        return self.normalize(self.optimal_log2())


class NormHistogram:
    """src/histogram.rs:289-506"""

    def __init__(self, table, log2, table_len):
        self._table, self._log2, self._table_len = [int(x) for x in table], int(log2), int(table_len)

    @staticmethod
    def new(data):                                            # :299-303
        h = Histogram(data)
        return h.normalize(h.optimal_log2())

    @staticmethod
    def try_from(table):                                      # :508-536
        s = sum(abs(int(x)) for x in table)
        log2 = _ilog2(s)
        if (1 << log2) != s:
            raise ValueError("sum of counts is not a power of two")
        table_len = 0
        for i in reversed(range(256)):
            if table[i] != 0:
                table_len = i
                break
        return NormHistogram(table, log2, table_len + 1)

    def __eq__(self, o):
        return (self._table, self._log2, self._table_len) == (o._table, o._log2, o._table_len)

    def table(self):
        return list(self._table)

    def table_iter(self):
        return iter(self._table[: self._table_len])

    def log2_sum(self):
        return self._log2

    def symbol_count(self):                                   # :321-323 (quirk Q3)
        return sum(1 for x in self._table if x == 0)

    def table_len(self):
        return self._table_len

    def write_bound(self):                                    # :330-337
        return (((self._table_len * self._log2) >> 3) + 3) if self._table_len > 1 else 512

    def _dev(self):
        import torch
        ctx = _ctx()
        norm = torch.tensor(self._table, dtype=torch.int32, device=ctx.device).reshape(1, 256)
        l2 = torch.tensor([self._log2], dtype=torch.int32, device=ctx.device)
        tl = torch.tensor([self._table_len], dtype=torch.int32, device=ctx.device)
        return ctx, norm, l2, tl

    def write(self, writer):                                  # :376-431 -> bits written
        ctx, norm, l2, tl = self._dev()
        rows, nbytes, nbits = ctx.ncount_write(norm, l2, tl)
        nb = int(nbytes[0].item())
        writer.extend(rows[0, :nb].cpu().numpy().tobytes())
        return int(nbits[0].item())

    @staticmethod
    def read(data):                                           # :436-505 -> (hist, rest)
        import torch
        data = bytes(data)
        if len(data) == 0:
            raise Panic("No bytes provided to read from")    # stream_reader.rs:17
        ctx = _ctx()
        rows = _dev_u8(data).reshape(1, -1)
        lens = torch.tensor([len(data)], dtype=torch.int32, device=ctx.device)
        norm, l2, tl, cons, st = ctx.ncount_read(rows, lens)
        rc = int(st[0].item())
        if rc == -3:
            raise HistError(HistError.TableLogTooLarge)
        if rc == -4:
            raise HistError(HistError.TooManySymbols)
        if rc == -5:
            raise HistError(HistError.Io)
        if rc < 0:
            raise Panic("status %d" % rc)
        return NormHistogram(norm[0].cpu().tolist(), int(l2[0].item()), int(tl[0].item())), data[int(cons[0].item()):]


class _Fse:
    """src/fse.rs (the table builders; Encoder / Decoder live inside the kernels)"""

    class EncodeTable:                                        # :72-194
        def __init__(self, hist):
            if not (TABLE_LOG_MIN <= hist.log2_sum() <= TABLE_LOG_MAX):
                raise Panic("FSE Table must be between 2^9 to 2^16")   # :103-106 (message as in the crate)
            ctx, norm, l2, tl = hist._dev()
            table, tt, sym, st = ctx.build_encode_tables(norm, l2, tl, hist.log2_sum())
            if int(st[0].item()) < 0:
                raise Panic("status %d" % int(st[0].item()))
            size = 1 << hist.log2_sum()
            self.table_log = hist.log2_sum()
            self.table = [int(x) & 0xFFFF for x in table[0, :size].cpu().tolist()]
            self.symbol_tt = [(int(b) & 0xFFFFFFFF, int(f)) for b, f in tt[0].cpu().tolist()]
            self.symbols = sym[0, :size].cpu().tolist()

        @staticmethod
        def compress_bound(size):                             # :191-193
            return 512 + size + (size >> 7) + 4 + 8

    class DecodeTable:                                        # :253-339
        def __init__(self, hist):
            if not (TABLE_LOG_MIN <= hist.log2_sum() <= TABLE_LOG_MAX):
                raise Panic("FSE Table must be between 2^9 to 2^16")
            ctx, norm, l2, tl = hist._dev()
            table, st = ctx.build_decode_tables(norm, l2, tl, hist.log2_sum())
            if int(st[0].item()) < 0:
                raise Panic("status %d" % int(st[0].item()))
            size = 1 << hist.log2_sum()
            self.table_log = hist.log2_sum()
            raw = [int(x) & 0xFFFFFFFF for x in table[0, :size].cpu().tolist()]
            # (new_state, symbol, num_bits), fse.rs:260-265
            self.table = [(e & 0xFFFF, (e >> 16) & 0xFF, e >> 24) for e in raw]


fse = _Fse


def _compress(src, dst, n_states):
    import torch
    src = bytes(src)
    if len(src) < n_states or len(src) == 0:
        raise Panic("called `Option::unwrap()` on a `None` value")     # lib.rs:121,154,156
    ctx = _ctx()
    d, off, st, total = ctx.compress_blocks(_dev_u8(src), len(src), 0, n_states)
    code = int(st[0].item())
    if code != 0:
        raise Panic("the reference panics on this input (block status %d)" % code)
    out = d[:total].cpu().numpy().tobytes()
    dst.extend(out)
    return out


def _payload_bits(stream, header_bytes):
    pay = stream[header_bytes:]
    return (len(pay) - 1) * 8 + pay[-1].bit_length()


def fse_compress(src, dst):
    """src/lib.rs:112-143: appends header || 1-state payload to `dst` (bytearray); returns
    (NormHistogram, payload bits incl. the marker)."""
    out = _compress(src, dst, 1)
    hist, rest = NormHistogram.read(out)
    return hist, _payload_bits(out, len(out) - len(rest))


def fse_compress2(src, dst):
    """src/lib.rs:146-183: two interleaved states; returns the payload bit count."""
    out = _compress(src, dst, 2)
    _, rest = NormHistogram.read(out)
    return _payload_bits(out, len(out) - len(rest))


EXHAUST_LIMIT = 1 << 30     # hard limit of the exhaust-mode output when the caller gives no max_len


def _decompress(src, dst, n_states, max_len):
    import torch
    src = bytes(src)
    if len(src) == 0:
        raise Panic("No bytes provided to read from")        # stream_reader.rs:17 via lib.rs:191,219
    ctx = _ctx()
    # No length is stored (lib.rs:198,228): the capacity grows geometrically until the stream fits (skewed data can expand
    # far beyond 64 x: p = 0.995 gives ~175 x).  Only the stream that never terminates (every state needs 0 bits, SURVEY.md
    # Q1: the reference loops until out of memory) ends in a Panic, at EXHAUST_LIMIT or the caller's max_len.
    cap = max_len if max_len is not None else max(4096, 64 * len(src))
    comp = _dev_u8(src)
    off = torch.tensor([0, len(src)], dtype=torch.int64, device=ctx.device)
    while True:
        out, out_len, st = ctx.decompress_exhaust(comp, len(src), off, 1, cap, 15, n_states)
        code = int(st[0].item())
        if code == -2 and max_len is None and cap < EXHAUST_LIMIT:
            cap = min(cap * 8, EXHAUST_LIMIT)
            continue
        break
    if code in (-3, -4, -5, -6):
        return None                                           # .ok()? / BitStackReader::new -> None
    if code == -7:
        raise Panic("called `Option::unwrap()` on a `None` value")     # lib.rs:197,224-225
    if code == -2:
        raise Panic("decoder does not terminate within %d bytes (SURVEY.md Q1)" % cap)
    if code < 0:
        raise Panic("status %d" % code)
    n = int(out_len[0].item())
    dst.extend(out[0, :n].cpu().numpy().tobytes())
    return n


def fse_decompress(src, dst, max_len=None):
    """src/lib.rs:187-211 -> bytes appended, or None."""
    return _decompress(src, dst, 1, max_len)


def fse_decompress2(src, dst, max_len=None):
    """src/lib.rs:215-248 -> bytes appended, or None."""
    return _decompress(src, dst, 2, max_len)
