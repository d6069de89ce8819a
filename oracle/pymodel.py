"""Independent Python model of the reference crate's *mechanics* (TEST INFRASTRUCTURE ONLY).

This is the second, independently written checker named in oracle/fse_oracle.h.  Where the C
oracle restates the *semantics* (bit k of the stream is bit k%8 of byte k/8), this model follows
the reference's actual mechanics: the 64-bit accumulator, alignment-dependent partial flushes,
Vec growth, the aligned 32-bit refills of the stack reader and the cached tail words of the
stream reader -- with pointers simulated as integer addresses so buffer alignment can be swept
the way the reference's own tests do (bitstream/mod.rs:151-155).  If both agree on every byte,
the restatement in Appendix A of SURVEY.md is right.

Only tests/ may import this module.  Pure Python: keep inputs <= ~64 KiB.

Reference citations are relative to /root/reference/src.
"""

BITS = 64          # bitstream/mod.rs:1-4 on a 64-bit host
BYTES = 8
HALF_BYTES = 4
HALF_BITS = 32
M64 = (1 << 64) - 1

TABLE_LOG_MIN, TABLE_LOG_MAX, TABLE_LOG_DEFAULT = 5, 15, 11   # lib.rs:9-12


def find_mask(n):  # lib.rs:54-57
    assert n <= 32
    return (1 << n) - 1


def ilog2(v):
    if v <= 0:
        raise ArithmeticError("ilog2 of zero")  # Rust panics
    return v.bit_length() - 1


class Vec:
    """A Vec<u8> whose heap pointer has a chosen alignment (address = base + index)."""

    def __init__(self, data=b"", cap=0, base=0x1000):
        self.base = base
        self.mem = bytearray(data) + bytearray(max(cap, len(data)) - len(data))
        self.len = len(data)

    @property
    def cap(self):
        return len(self.mem)

    def reserve(self, additional):  # alloc::raw_vec amortised growth
        if self.cap - self.len >= additional:
            return
        new_cap = max(self.cap * 2, self.len + additional, 8)
        self.mem = self.mem[: self.len] + bytearray(new_cap - self.len)

    def bytes(self):
        return bytes(self.mem[: self.len])


class BitStackWriter:  # bitstream/writer.rs
    def __init__(self, vec):  # :16-40
        vec.reserve(2 * BYTES)
        self.v = vec
        self.initial_len = vec.len
        self.ptr = vec.len
        self._set_end()
        self.storage = 0
        self.bits = 0

    def _addr(self, off):
        return self.v.base + off

    def _set_end(self):  # :24-31, :100-107
        end = self.v.cap
        align = (-self._addr(end)) % HALF_BYTES
        self.end_ptr = end + align - HALF_BYTES if align != 0 else end

    def _wr(self, off, b):
        assert off < self.v.cap, "write past the Vec's capacity"
        self.v.mem[off] = b

    def flush(self):  # :43-110
        align = (-self._addr(self.ptr)) % HALF_BYTES
        if align != 0:
            to_write = min(self.bits // 8, align)
            assert to_write < HALF_BYTES
            b = (self.storage & M64).to_bytes(8, "little")
            self._wr(self.ptr, b[0])
            self._wr(self.ptr + 1, b[1])
            self._wr(self.ptr + 2, b[2])
            self.ptr += to_write
            self.storage >>= to_write * 8
            self.bits -= to_write * 8
            if (-self._addr(self.ptr)) % HALF_BYTES != 0:
                return
        inc = 1 if (self.bits // HALF_BITS) != 0 else 0
        w = (self.storage & 0xFFFFFFFF).to_bytes(4, "little")  # raw_write :113-119
        for i in range(4):
            self._wr(self.ptr + i, w[i])
        self.storage >>= inc * HALF_BITS
        self.bits -= inc * HALF_BITS
        self.ptr += inc * HALF_BYTES
        if self.ptr == self.end_ptr:  # :93-108
            self.v.len = self.ptr
            self.v.reserve(self.v.len * 2 + HALF_BYTES)
            self.ptr = self.v.len
            self._set_end()

    def write_bits_raw(self, val, bits):  # :163-180
        assert bits <= 16
        assert (val & ~((1 << bits) - 1)) == 0
        self.storage |= val << self.bits
        self.bits += bits
        assert self.bits <= 64, "accumulator overflow: caller must flush every 32 bits"

    def write_bits_raw_unmasked(self, val, bits):  # :140-149
        self.write_bits_raw(val & find_mask(bits), bits)

    def write_bits(self, val, bits):  # :185-190
        self.write_bits_raw(val, bits)
        self.flush()

    def write_bits_unmasked(self, val, bits):  # :195-198
        self.write_bits_raw_unmasked(val, bits)
        self.flush()

    def finish(self):  # :201-222
        total_size = self.ptr
        total_bits = total_size * 8 + self.bits
        self.bits = BITS
        self.flush()
        self.bits = BITS
        self.flush()
        self.v.len = (total_bits + 7) // 8
        return total_bits - self.initial_len * 8


class BitStackReader:  # bitstream/stack_reader.rs
    def __init__(self, data, base=0x2000):
        """Raises ValueError where the reference returns None."""
        self.d = bytes(data)
        self.base = base
        n = len(self.d)
        if n == 0:
            raise ValueError("None: empty")  # :18-20
        ptr = n - 1
        align = (-(base + ptr)) % HALF_BYTES
        if ptr > (HALF_BYTES - align):  # :32-36
            ptr = ptr + align - HALF_BYTES
        else:
            ptr = 0
        to_read = n - ptr  # :40-46
        buf = 0
        for i in range(min(to_read, BYTES)):
            buf |= self.d[ptr + i] << (8 * i)
        self.buffer = buf
        self.bits = to_read * 8
        self.finished = ptr == 0
        ptr = ptr - HALF_BYTES if ptr >= HALF_BYTES else 0  # :49-55
        self.ptr = ptr
        self.reload()
        if self.buffer == 0:  # :77-83
            raise ValueError("None: no marker")
        highbit = ilog2(self.buffer)
        if self.bits - highbit > 8:
            raise ValueError("None: no marker")
        self.bits = highbit
        self.reload()

    def reload(self):  # :97-172
        if self.finished:
            return
        if self.ptr == 0:
            to_read = HALF_BYTES - ((self.base + self.ptr) & (HALF_BYTES - 1))
            self.finished = self.bits <= HALF_BITS
            if not self.finished:
                to_read = 0
            read = 0
            for i in range(to_read):
                read |= self.d[self.ptr + i] << (8 * i)
            read_bits = 8 * to_read
            self.buffer = ((self.buffer << read_bits) | read) & M64
            self.bits += read_bits
            return
        will_read = 1 if self.bits <= HALF_BITS else 0
        read_bits = will_read * HALF_BITS
        read = int.from_bytes(self.d[self.ptr : self.ptr + 4], "little") if will_read else 0
        self.buffer = ((self.buffer << read_bits) | read) & M64
        self.bits += read_bits
        base_offset = self.ptr
        if base_offset >= HALF_BYTES:
            self.ptr -= will_read * HALF_BYTES
        else:
            self.ptr -= will_read * base_offset

    def peek(self, bits):  # :176-184
        assert bits <= 16
        if bits > self.bits:
            return None
        return (self.buffer >> (self.bits - bits)) & find_mask(bits)

    def read_no_reload(self, bits):  # :193-197
        v = self.peek(bits)
        if v is None:
            return None
        self.bits -= bits
        return v

    def read(self, bits):  # :211-215
        v = self.read_no_reload(bits)
        if v is None:
            return None
        self.reload()
        return v

    def available(self):
        return self.bits

    def finish(self):  # :224-226
        return self.finished and self.bits == 0


class BitStreamReader:  # bitstream/stream_reader.rs
    def __init__(self, data, total_bits):  # :16-49
        self.d = bytes(data)
        assert len(self.d) != 0
        assert (total_bits + 7) // 8 == len(self.d)
        self.total_bits = total_bits
        self.bits_read = 0
        word_offset = ((total_bits - 1) // BITS) * BYTES
        self.last0 = self._gather(word_offset)
        if word_offset + BYTES // 2 > len(self.d):
            word_offset = max(0, word_offset - BYTES // 2)
        else:
            word_offset += BYTES // 2
        self.last1 = self._gather(word_offset)

    def _gather(self, off):
        v = 0
        for i in reversed(range(BYTES)):
            if off + i < len(self.d):
                v = (v << 8) | self.d[off + i]
        return v

    def peek(self, bits):  # :82-114 ; raises EOFError for io::ErrorKind::UnexpectedEof
        if self.bits_read + bits > self.total_bits:
            raise EOFError
        idx = (self.bits_read // (BITS // 2)) * (BYTES // 2)
        bit_offset = self.bits_read & (BITS // 2 - 1)
        if idx + BYTES > len(self.d):
            word = self.last1 if (idx // (BYTES // 2)) & 1 == 1 else self.last0
        else:
            word = int.from_bytes(self.d[idx : idx + BYTES], "little")
        return (word >> bit_offset) & find_mask(bits)

    def advance_by(self, bits):  # :67-75
        if self.bits_read + bits > self.total_bits:
            raise EOFError
        self.bits_read += bits

    def read(self, bits):  # :56-60
        v = self.peek(bits)
        self.advance_by(bits)
        return v

    def finish(self):  # :123-128
        return self.d[self.bits_read // 8 :], self.total_bits - self.bits_read, self.bits_read % 8

    def finish_byte(self):  # :132-135
        return self.d[(self.bits_read + 7) // 8 :]


# ------------------------------------------------------------------ histogram.rs

class Histogram:
    def __init__(self, data):  # :18-66
        assert len(data) <= 0xFFFFFFFF
        t = [0] * 256
        for b in data:
            t[b] += 1
        self.table = t
        table_len = 0
        for i in reversed(range(256)):
            if t[i] != 0:
                table_len = i
                break
        self.table_len = table_len + 1
        self.size = len(data)

    def optimal_log2(self):  # :264-277
        min_bits_src = ilog2(self.size) + 1
        min_bits_symbols = ilog2(self.table_len - 1) + 2
        min_bits = min(min_bits_src, min_bits_symbols)
        max_bits = ilog2(self.size - 1) - 2
        if max_bits < 0:
            raise ArithmeticError("u32 underflow")
        return max(TABLE_LOG_MIN, min(TABLE_LOG_MAX, max(min(TABLE_LOG_DEFAULT, max_bits), min_bits)))

    def normalize(self, log2):  # :95-155
        log2 = max(min(max(log2, TABLE_LOG_MIN), TABLE_LOG_MAX), ilog2(self.table_len - 1) + 2)
        RTB = [0, 473195, 504333, 520860, 550000, 700000, 750000, 830000]
        scale = 62 - log2
        step = (1 << 62) // self.size
        v_step = 1 << (scale - 20)
        low_threshold = self.size >> log2
        to_distribute = 1 << log2
        largest = 0
        largest_prob = 0
        table = [0] * 256
        for i in range(self.table_len):
            t = self.table[i]
            if t == self.size:
                table[i] = to_distribute
                return NormHistogram(table, log2, self.table_len)
            if t == 0:
                continue
            if t <= low_threshold:
                table[i] = -1
                to_distribute -= 1
                continue
            prob = (t * step) >> scale
            if prob < 8:
                rest_to_beat = v_step * RTB[prob]
                prob += 1 if (t * step - (prob << scale)) > rest_to_beat else 0
            if prob > largest_prob:
                largest_prob = prob
                largest = i
            table[i] = prob
            to_distribute -= prob
        if to_distribute != 0 and -to_distribute >= (largest_prob >> 1):
            return self.normalize_slow(log2)
        table[largest] += to_distribute
        return NormHistogram(table, log2, self.table_len)

    def normalize_slow(self, log2):  # :157-261
        UNASSIGNED = -2
        low_threshold = self.size >> log2
        low_one = ((self.size * 3) & 0xFFFFFFFF) >> (log2 + 1)
        table = [0] * 256
        to_distribute = 1 << log2
        total = self.size
        for i in range(self.table_len):
            t = self.table[i]
            if t == 0:
                continue
            elif t <= low_threshold:
                table[i] = -1
                to_distribute -= 1
                total -= t
            elif t <= low_one:
                table[i] = 1
                to_distribute -= 1
                total -= t
            else:
                table[i] = UNASSIGNED
        if to_distribute == 0:
            return NormHistogram(table, log2, self.table_len, slow=True)
        if (total // to_distribute) > low_one:
            low = (total * 3) // (to_distribute * 2)
            for i in range(self.table_len):
                if table[i] == UNASSIGNED and self.table[i] <= low:
                    table[i] = 1
                    to_distribute -= 1
                    total -= self.table[i]
        if ((1 << log2) - to_distribute) == self.table_len:
            v_max, i_max = 0, 0
            for i, v in enumerate(self.table):
                if v > v_max:
                    v_max, i_max = v, i
            table[i_max] += to_distribute
            return NormHistogram(table, log2, self.table_len, slow=True)
        elif total == 0:
            while to_distribute != 0:
                for i in range(self.table_len):
                    if table[i] > 0:
                        table[i] += 1
                        to_distribute -= 1
                        if to_distribute == 0:
                            break
        else:
            v_step_log = 62 - log2
            mid = (1 << (v_step_log - 1)) - 1
            r_step = (((1 << v_step_log) * to_distribute) + mid) // total
            tmp_total = mid
            for i in range(self.table_len):
                if table[i] == UNASSIGNED:
                    end = tmp_total + self.table[i] * r_step
                    weight = (end >> v_step_log) - (tmp_total >> v_step_log)
                    if weight < 1:
                        raise ArithmeticError("cursed distribution")
                    table[i] = weight
                    tmp_total = end
        return NormHistogram(table, log2, self.table_len, slow=True)


class NormHistogram:
    def __init__(self, table, log2, table_len, slow=False):
        self.table, self.log2, self.table_len, self.slow = list(table), log2, table_len, slow

    @staticmethod
    def new(data):  # :299-303
        h = Histogram(data)
        return h.normalize(h.optimal_log2())

    def __eq__(self, o):
        return (self.table, self.log2, self.table_len) == (o.table, o.log2, o.table_len)

    def write_bound(self):  # :330-337
        return (((self.table_len * self.log2) >> 3) + 3) if self.table_len > 1 else 512

    def write(self, vec):  # :376-431
        w = BitStackWriter(vec)
        w.write_bits(self.log2 - TABLE_LOG_MIN, 4)
        threshold = 1 << self.log2
        remaining = threshold + 1
        zero_count = 0
        num_bits = self.log2 + 1
        for s in self.table[: self.table_len]:
            if remaining <= 1:
                break
            if zero_count != 0:
                if s == 0:
                    zero_count += 1
                    continue
                zero_count -= 1
                while zero_count >= 24:
                    w.write_bits(0xFFFF, 16)
                    zero_count -= 24
                while zero_count >= 3:
                    w.write_bits(0x3, 2)
                    zero_count -= 3
                w.write_bits(zero_count, 2)
            mx = (2 * threshold - 1) - remaining
            remaining -= abs(s)
            count = s + 1
            if count >= threshold:
                count += mx
            bits_to_write = num_bits - (1 if count < mx else 0)
            w.write_bits(count, bits_to_write)
            zero_count = 1 if count == 1 else 0
            if remaining < 1:
                raise ArithmeticError("Normalized histogram was incorrect somehow")
            while remaining < threshold:
                num_bits -= 1
                threshold >>= 1
        return w.finish()

    @staticmethod
    def read(data):  # :436-505 ; returns (hist, rest) or raises
        r = BitStreamReader(data, len(data) * 8)
        log2 = r.read(4) + TABLE_LOG_MIN
        if log2 > TABLE_LOG_MAX:
            raise ValueError("TableLogTooLarge")
        table = [0] * 256
        symbol = 0
        threshold = 1 << log2
        remaining = threshold + 1
        read_bit_count = log2 + 1
        previous0 = False

        def peek_or0(n):
            try:
                return r.peek(n)
            except EOFError:
                return 0

        while remaining > 1 and symbol < 256:
            if previous0:
                while peek_or0(16) == 0xFFFF:
                    r.advance_by(16)
                    symbol += 24
                while peek_or0(2) == 3:
                    r.advance_by(2)
                    symbol += 3
                symbol += r.read(2)
            if symbol >= 256:
                break
            mx = (2 * threshold - 1) - remaining
            try:
                raw = r.peek(read_bit_count)
            except EOFError:
                raw = r.peek(read_bit_count - 1)
            if (raw & (threshold - 1)) < mx:
                r.advance_by(read_bit_count - 1)
                value = raw & (threshold - 1)
            else:
                r.advance_by(read_bit_count)
                value = raw & (2 * threshold - 1)
                if value >= threshold:
                    value -= mx
            value -= 1
            remaining -= abs(value)
            table[symbol] = value
            symbol += 1
            previous0 = value == 0
            while remaining < threshold:
                read_bit_count -= 1
                threshold >>= 1
        if remaining != 1:
            raise ValueError("TooManySymbols")
        return NormHistogram(table, log2, symbol), r.finish_byte()


# ------------------------------------------------------------------------ fse.rs

def table_step(size):  # :68-70
    return size * 5 // 8 + 3


class EncodeTable:  # :72-194
    def __init__(self, hist):
        tl = hist.log2
        assert TABLE_LOG_MIN <= tl <= TABLE_LOG_MAX
        self.table_log = tl
        size = 1 << tl
        cumul = [0] * 256
        high_threshold = size - 1
        symbols = [0] * size
        acc = 0
        for i in range(hist.table_len):
            x = hist.table[i]
            cumul[i] = acc
            if x == -1:
                acc += 1
                symbols[high_threshold] = i
                high_threshold -= 1
            else:
                acc += x
        position = 0
        mask = size - 1
        step = table_step(size)
        for i in range(hist.table_len):
            for _ in range(max(hist.table[i], 0)):
                symbols[position] = i
                position = (position + step) & mask
                while position > high_threshold:
                    position = (position + step) & mask
        assert position == 0
        table = [0] * size
        for i, x in enumerate(symbols):
            table[cumul[x]] = size + i
            cumul[x] += 1
        tt = [(0, 0)] * 256
        total = 0
        for i in range(hist.table_len):
            x = hist.table[i]
            if x == 0:
                tt[i] = ((((tl + 1) << 16) - (1 << tl)) & 0xFFFFFFFF, 0)
            elif x in (-1, 1):
                tt[i] = (((tl << 16) - (1 << tl)) & 0xFFFFFFFF, total - 1)
                total += 1
            else:
                mbo = tl - ilog2(x - 1)
                tt[i] = (((mbo << 16) - (x << mbo)) & 0xFFFFFFFF, total - x)
                total += x
        self.table, self.symbol_tt, self.symbols = table, tt, symbols

    @staticmethod
    def compress_bound(size):  # :191-193
        return 512 + size + (size >> 7) + 4 + 8


class Encoder:  # :196-251
    def __init__(self, table, first_symbol):  # new_first_symbol :210-218
        self.t = table
        bits, find_state = table.symbol_tt[first_symbol]
        bits_out = ((bits + (1 << 15)) & 0xFFFFFFFF) >> 16
        value = ((bits_out << 16) - bits) & 0xFFFFFFFF
        self.value = table.table[(value >> bits_out) + find_state]

    def encode_raw(self, w, sym):  # :227-239
        bits, find_state = self.t.symbol_tt[sym]
        bits_out = ((bits + self.value) & 0xFFFFFFFF) >> 16
        w.write_bits_raw_unmasked(self.value, bits_out)
        self.value = self.t.table[(self.value >> bits_out) + find_state]

    def encode(self, w, sym):  # :242-245
        w.flush()
        self.encode_raw(w, sym)

    def finish(self, w):  # :248-250
        w.write_bits_unmasked(self.value, self.t.table_log)


class DecodeTable:  # :253-339
    def __init__(self, hist):
        tl = hist.log2
        assert TABLE_LOG_MIN <= tl <= TABLE_LOG_MAX
        self.table_log = tl
        size = 1 << tl
        sym = [0] * size
        symbol_next = [0] * 256
        high_threshold = size - 1
        for s in range(hist.table_len):
            c = hist.table[s]
            if c <= -1:
                sym[high_threshold] = s
                high_threshold -= 1
                symbol_next[s] = 1
            else:
                symbol_next[s] = c
        position = 0
        mask = size - 1
        step = table_step(size)
        for s in range(hist.table_len):
            for _ in range(max(hist.table[s], 0)):
                sym[position] = s
                position = (position + step) & mask
                while position > high_threshold:
                    position = (position + step) & mask
        assert position == 0
        self.table = []
        for i in range(size):
            s = sym[i]
            nxt = symbol_next[s]
            symbol_next[s] += 1
            nb = tl - ilog2(nxt)
            self.table.append((((nxt << nb) - size) & 0xFFFF, s, nb))  # (new_state, symbol, num_bits)


class Decoder:  # :341-386
    def __init__(self, table, reader):
        self.t = table
        st = reader.read(table.table_log)
        if st is None:
            raise ArithmeticError("unwrap on None")  # lib.rs:197,224-225
        self.state = st

    def decode_symbol_no_reload(self, reader):  # :363-373
        new_state, sym, nb = self.t.table[self.state]
        low = reader.read_no_reload(nb)
        if low is None:
            return None
        self.state = (new_state + low) & 0xFFFF
        return sym

    def decode_symbol(self, reader):  # :376-380
        s = self.decode_symbol_no_reload(reader)
        if s is None:
            return None
        reader.reload()
        return s

    def finish(self):  # :383-385
        return self.t.table[self.state][1]


# ------------------------------------------------------------------------ lib.rs

def fse_compress(src, vec):  # :112-143
    hist = NormHistogram.new(src)
    hist.write(vec)
    w = BitStackWriter(vec)
    t = EncodeTable(hist)
    chunks = [src[i : i + 2] for i in range(0, len(src), 2)][::-1]
    first = chunks[0]
    enc = Encoder(t, first[-1])
    if len(first) > 1:
        enc.encode(w, first[0])
    for n in chunks[1:]:
        enc.encode_raw(w, n[1])
        enc.encode_raw(w, n[0])
        w.flush()
    enc.finish(w)
    w.write_bits(1, 1)
    return hist, w.finish()


def fse_compress2(src, vec):  # :146-183
    hist = NormHistogram.new(src)
    hist.write(vec)
    w = BitStackWriter(vec)
    t = EncodeTable(hist)
    chunks = [src[i : i + 2] for i in range(0, len(src), 2)][::-1]
    first = chunks[0]
    rest = 1
    if len(first) == 1:
        nxt = chunks[1]
        rest = 2
        e0 = Encoder(t, first[0])
        e1 = Encoder(t, nxt[1])
        e0.encode(w, nxt[0])
    else:
        e0 = Encoder(t, first[0])
        e1 = Encoder(t, first[1])
    for n in chunks[rest:]:
        e1.encode_raw(w, n[1])
        e0.encode_raw(w, n[0])
        w.flush()
    e1.finish(w)
    e0.finish(w)
    w.write_bits(1, 1)
    return w.finish()


def fse_decompress(src, limit=None, base=0x2000):  # :187-211 ; limit guards quirk Q1
    try:
        hist, rest = NormHistogram.read(src)
        reader = BitStackReader(rest, base=base + (len(src) - len(rest)))
    except (ValueError, EOFError):
        return None
    t = DecodeTable(hist)
    d = Decoder(t, reader)
    out = bytearray()
    while True:
        s = d.decode_symbol(reader)
        if s is None:
            break
        out.append(s)
        s = d.decode_symbol_no_reload(reader)
        if s is None:
            break
        out.append(s)
        if limit is not None and len(out) > limit:
            raise OverflowError("decoder does not terminate (Q1)")
    out.append(d.finish())
    return bytes(out)


def fse_decompress2(src, limit=None, base=0x2000):  # :215-248
    try:
        hist, rest = NormHistogram.read(src)
        reader = BitStackReader(rest, base=base + (len(src) - len(rest)))
    except (ValueError, EOFError):
        return None
    t = DecodeTable(hist)
    d0 = Decoder(t, reader)
    d1 = Decoder(t, reader)
    out = bytearray()
    while True:
        s = d0.decode_symbol_no_reload(reader)
        if s is None:
            out.append(d0.finish())
            out.append(d1.finish())
            break
        out.append(s)
        s = d1.decode_symbol(reader)
        if s is None:
            out.append(d1.finish())
            out.append(d0.finish())
            break
        out.append(s)
        if limit is not None and len(out) > limit:
            raise OverflowError("decoder does not terminate (Q1)")
    return bytes(out)


def gen_sequence_lut(prob=0.2):  # lib.rs:255-270
    lut = [0] * 4096
    prob = min(max(prob, 0.005), 0.995)
    remaining, idx, s = 4096, 0, 0
    while remaining > 0:
        n = max(int(remaining * prob), 1)
        for _ in range(n):
            lut[idx] = s
            idx += 1
        s += 1
        remaining -= n
    return lut
