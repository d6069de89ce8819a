"""entropy_coders_b200 -- B200 (sm_100a) FSE / tANS coder behind the public API of the Rust crate
Cognoscan/entropy_coders.

Layers (all arithmetic runs in libfse_b200.so; this package is host plumbing only):

* `_capi`        ctypes binding of include/fse_b200.h (the drop-in C ABI).
* `Context`      device-pointer API over torch CUDA tensors: stage entry points
                 (histogram, normalise, NCount header, table builds) and the fused block pipelines.
* crate mirror   `Histogram`, `NormHistogram`, `fse.EncodeTable`, `fse.DecodeTable`,
                 `fse_compress`, `fse_compress2`, `fse_decompress`, `fse_decompress2` with the crate's
                 names, argument meaning and error behaviour (src/lib.rs:7, :112-248), each backed by
                 the CUDA path with the whole input as one block.

There is no CPU fallback: without the built library or without a CUDA device every call raises.
"""
import ctypes as C

from . import _capi
from ._capi import FseError, Params

TABLE_LOG_MIN, TABLE_LOG_MAX, TABLE_LOG_DEFAULT = 5, 15, 11  # src/lib.rs:9-12
TABLE_PER_BLOCK, TABLE_GLOBAL = 0, 1
FLAG_RAW_IF_EXPANDS = 1
GEN_KINDS = {"geo": 0, "text": 1, "few": 2, "uniform": 3}


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("entropy_coders_b200 needs a CUDA device (no CPU fallback)")
    return torch


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class Context:
    """One fse_b200_ctx: a device, a stream and reusable device workspaces."""

    def __init__(self, device=0, stream=None):
        torch = _torch()
        self._L = _capi.lib()
        self.device = torch.device("cuda", device)
        self._h = C.c_void_p()
        if stream is None:
            # Run on torch's current stream, so that tensors produced by (asynchronous) torch operations are ordered
            # before the library's kernels.  torch's default stream is the legacy default stream, whose handle is 0;
            # the C ABI reads NULL as "create a private stream", so it is passed as cudaStreamLegacy (1).
            with torch.cuda.device(self.device):
                stream = torch.cuda.current_stream(self.device).cuda_stream or 1
        s = C.c_void_p(stream) if stream else None
        rc = self._L.fse_b200_create(device, s, C.byref(self._h))
        if rc != 0:
            raise FseError(rc, "fse_b200_create")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.fse_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise FseError(rc, self._L.fse_b200_last_error(self._h).decode())

    @property
    def launches(self):
        return int(self._L.fse_b200_launch_count(self._h))

    def sync(self):
        self._ck(self._L.fse_b200_sync(self._h))

    KERNELS = ("hist", "encode", "scan", "gather", "decode")

    def set_timing(self, enable=True):
        self._ck(self._L.fse_b200_set_timing(self._h, 1 if enable else 0))

    def get_timing(self):
        """-> {kernel: (total ms, launches)} from CUDA events around each launch"""
        ms = (C.c_double * 5)()
        cnt = (C.c_uint64 * 5)()
        self._ck(self._L.fse_b200_get_timing(self._h, ms, cnt))
        return {k: (ms[i], int(cnt[i])) for i, k in enumerate(self.KERNELS)}

    # ------------------------------------------------------------------ helpers
    def _u8(self, n):
        return _torch().empty(max(int(n), 1), dtype=_torch().uint8, device=self.device)

    def params(self, block_size, table_log=0, n_states=32, table_mode=TABLE_PER_BLOCK, segment_size=0, flags=0):
        return Params(int(block_size), int(table_log), int(n_states), int(table_mode), int(segment_size), int(flags))

    def num_blocks(self, n, block_size):
        return int(self._L.fse_b200_num_blocks(n, block_size))

    def num_streams(self, n, p):
        """entries of the stream index: blocks, or segments when p.segment_size > 0"""
        return int(self._L.fse_b200_num_streams(n, C.byref(p)))

    def bound(self, n, p):
        return int(self._L.fse_b200_compress_blocks_bound(n, C.byref(p)))

    def generate(self, kind, seed, n, first_index=0):
        """Synthetic bytes of SURVEY.md 8(d), generated on the device."""
        out = self._u8(n)
        self._ck(self._L.fse_b200_generate(self._h, GEN_KINDS[kind], seed, first_index, _ptr(out), n))
        return out[:n]

    # ------------------------------------------------------------------ stages
    def histogram_blocks(self, src, block_size):
        """Histogram::new per block (src/histogram.rs:18-66) -> (counts int32[nb,256] holding u32, table_len[nb])"""
        torch = _torch()
        nb = self.num_blocks(src.numel(), block_size)
        counts = torch.empty((nb, 256), dtype=torch.int32, device=self.device)
        tlen = torch.empty(nb, dtype=torch.int32, device=self.device)
        self._ck(self._L.fse_b200_histogram_blocks(self._h, _ptr(src), src.numel(), block_size, _ptr(counts), _ptr(tlen)))
        return counts, tlen

    def histogram_global(self, src):
        torch = _torch()
        counts = torch.empty(256, dtype=torch.int64, device=self.device)
        self._ck(self._L.fse_b200_histogram_global(self._h, _ptr(src), src.numel(), _ptr(counts)))
        return counts

    def normalize(self, counts64, table_log=0):
        """Histogram::normalize (+optimal_log2 when table_log == 0), src/histogram.rs:95-277.
        counts64: int64[nt,256] -> (norm int32[nt,256], log2[nt], table_len[nt], status[nt])"""
        torch = _torch()
        counts64 = counts64.reshape(-1, 256).contiguous()
        nt = counts64.shape[0]
        norm = torch.empty((nt, 256), dtype=torch.int32, device=self.device)
        log2 = torch.empty(nt, dtype=torch.int32, device=self.device)
        tlen = torch.empty(nt, dtype=torch.int32, device=self.device)
        st = torch.empty(nt, dtype=torch.int32, device=self.device)
        self._ck(self._L.fse_b200_normalize(self._h, _ptr(counts64), nt, table_log, _ptr(norm), _ptr(log2), _ptr(tlen), _ptr(st)))
        return norm, log2, tlen, st

    def normalize_zstd(self, counts64, table_log=0, use_low_prob_count=True):
        """libzstd's FSE_normalizeCount (not the crate's normalize): counts64 int64[nt,256] ->
        (norm int32[nt,256], log2[nt], table_len[nt], status[nt])"""
        torch = _torch()
        counts64 = counts64.reshape(-1, 256).contiguous()
        nt = counts64.shape[0]
        norm = torch.empty((nt, 256), dtype=torch.int32, device=self.device)
        log2 = torch.empty(nt, dtype=torch.int32, device=self.device)
        tlen = torch.empty(nt, dtype=torch.int32, device=self.device)
        st = torch.empty(nt, dtype=torch.int32, device=self.device)
        self._ck(self._L.fse_b200_normalize_zstd(self._h, _ptr(counts64), nt, table_log, 1 if use_low_prob_count else 0,
                                                 _ptr(norm), _ptr(log2), _ptr(tlen), _ptr(st)))
        return norm, log2, tlen, st

    def ncount_write(self, norm, log2, tlen):
        """NormHistogram::write, src/histogram.rs:376-431 -> (rows uint8[nt,512], bytes[nt], bits[nt])"""
        torch = _torch()
        nt = norm.shape[0]
        out = torch.zeros((nt, 512), dtype=torch.uint8, device=self.device)
        nbytes = torch.empty(nt, dtype=torch.int32, device=self.device)
        nbits = torch.empty(nt, dtype=torch.int32, device=self.device)
        self._ck(self._L.fse_b200_ncount_write(self._h, _ptr(norm), _ptr(log2), _ptr(tlen), nt, _ptr(out), 512, _ptr(nbytes), _ptr(nbits)))
        return out, nbytes, nbits

    def ncount_read(self, rows, lens):
        """NormHistogram::read, src/histogram.rs:436-505 -> (norm, log2, table_len, consumed, status)"""
        torch = _torch()
        nt, stride = rows.shape
        norm = torch.empty((nt, 256), dtype=torch.int32, device=self.device)
        log2 = torch.empty(nt, dtype=torch.int32, device=self.device)
        tlen = torch.empty(nt, dtype=torch.int32, device=self.device)
        cons = torch.empty(nt, dtype=torch.int32, device=self.device)
        st = torch.empty(nt, dtype=torch.int32, device=self.device)
        self._ck(self._L.fse_b200_ncount_read(self._h, _ptr(rows), stride, _ptr(lens), nt, _ptr(norm), _ptr(log2), _ptr(tlen), _ptr(cons), _ptr(st)))
        return norm, log2, tlen, cons, st

    def build_encode_tables(self, norm, log2, tlen, max_table_log):
        """EncodeTable::update, src/fse.rs:101-189 -> (table int16[nt,S] holding u16, symbol_tt int32[nt,256,2], symbols uint8[nt,S], status)"""
        torch = _torch()
        nt, S = norm.shape[0], 1 << max_table_log
        table = torch.zeros((nt, S), dtype=torch.int16, device=self.device)
        tt = torch.zeros((nt, 256, 2), dtype=torch.int32, device=self.device)
        sym = torch.zeros((nt, S), dtype=torch.uint8, device=self.device)
        st = torch.empty(nt, dtype=torch.int32, device=self.device)
        self._ck(self._L.fse_b200_build_encode_tables(self._h, _ptr(norm), _ptr(log2), _ptr(tlen), nt, max_table_log,
                                                      _ptr(table), _ptr(tt), _ptr(sym), _ptr(st)))
        return table, tt, sym, st

    def build_decode_tables(self, norm, log2, tlen, max_table_log):
        """DecodeTable::update, src/fse.rs:280-338 -> (table int32[nt,S] = new_state | symbol<<16 | num_bits<<24, status)"""
        torch = _torch()
        nt, S = norm.shape[0], 1 << max_table_log
        table = torch.zeros((nt, S), dtype=torch.int32, device=self.device)
        st = torch.empty(nt, dtype=torch.int32, device=self.device)
        self._ck(self._L.fse_b200_build_decode_tables(self._h, _ptr(norm), _ptr(log2), _ptr(tlen), nt, max_table_log, _ptr(table), _ptr(st)))
        return table, st

    # ------------------------------------------------------------------ bit I/O primitives (src/bitstream)
    def bitstack_write(self, vals, bits, mark=True):
        """BitStackWriter: fields (vals[i] masked to bits[i] <= 16 bits) LSB first in index order (+ marker) -> (bytes, bit count)"""
        torch = _torch()
        n = int(vals.numel())
        cap = (int(bits.sum().item()) + 64) // 8 + 8 if n else 16
        cap = (cap + 3) & ~3
        out = torch.zeros(cap, dtype=torch.uint8, device=self.device)
        nbits = C.c_uint64()
        self._ck(self._L.fse_b200_bitstack_write(self._h, _ptr(vals), _ptr(bits), n, 1 if mark else 0, _ptr(out), cap, C.byref(nbits)))
        return out[: (nbits.value + 7) // 8], int(nbits.value)

    def bitstack_read(self, data, bits):
        """BitStackReader: the fields back in index order (read from the end) -> (vals, status)"""
        torch = _torch()
        n = int(bits.numel())
        vals = torch.zeros(max(n, 1), dtype=torch.int32, device=self.device)
        st = C.c_int32()
        self._ck(self._L.fse_b200_bitstack_read(self._h, _ptr(data), int(data.numel()), _ptr(bits), n, _ptr(vals), C.byref(st)))
        return vals[:n], int(st.value)

    def bitstream_read(self, data, total_bits, bits):
        """BitStreamReader: forward reads under a total_bits bound -> (vals, status: bits left or a negative code)"""
        torch = _torch()
        n = int(bits.numel())
        vals = torch.zeros(max(n, 1), dtype=torch.int32, device=self.device)
        st = C.c_int32()
        self._ck(self._L.fse_b200_bitstream_read(self._h, _ptr(data), int(data.numel()), int(total_bits), _ptr(bits), n, _ptr(vals), C.byref(st)))
        return vals[:n], int(st.value)

    # ------------------------------------------------------------------ fused pipelines (device tensors)
    def compress_blocks(self, src, block_size, table_log=0, n_states=32, table_mode=TABLE_PER_BLOCK, out=None, segment_size=0, flags=0):
        """-> (dst uint8[cap] (first `total` bytes valid), offsets int64[nb+1], status int32[nb], total); nb counts
        segments when segment_size > 0"""
        torch = _torch()
        p = self.params(block_size, table_log, n_states, table_mode, segment_size, flags)
        n = src.numel()
        nb = self.num_streams(n, p)
        cap = self.bound(n, p)
        if out is None:
            dst = self._u8(cap)
            offsets = torch.empty(nb + 1, dtype=torch.int64, device=self.device)
            status = torch.empty(max(nb, 1), dtype=torch.int32, device=self.device)
        else:
            dst, offsets, status = out
        total = C.c_uint64()
        self._ck(self._L.fse_b200_compress_blocks(self._h, _ptr(src), n, C.byref(p), _ptr(dst), dst.numel(), _ptr(offsets),
                                                  _ptr(status), C.byref(total)))
        return dst, offsets, status[:nb], int(total.value)

    def compress_blocks_async(self, src, p, dst, offsets, status):
        self._ck(self._L.fse_b200_compress_blocks_async(self._h, _ptr(src), src.numel(), C.byref(p), _ptr(dst), dst.numel(),
                                                        _ptr(offsets), _ptr(status)))

    def decompress_blocks(self, comp, total, offsets, n, block_size, table_log=0, n_states=32, table_mode=TABLE_PER_BLOCK, out=None,
                          segment_size=0, flags=0):
        """-> (dst uint8[n], status int32[nb])"""
        torch = _torch()
        p = self.params(block_size, table_log, n_states, table_mode, segment_size, flags)
        nb = self.num_streams(n, p)
        if out is None:
            dst = self._u8(n)
            status = torch.empty(max(nb, 1), dtype=torch.int32, device=self.device)
        else:
            dst, status = out
        self._ck(self._L.fse_b200_decompress_blocks(self._h, _ptr(comp), total, _ptr(offsets), nb, C.byref(p), _ptr(dst), n, _ptr(status)))
        return dst[:n], status[:nb]

    def decompress_blocks_async(self, comp, total, offsets, nb, p, dst, n, status):
        self._ck(self._L.fse_b200_decompress_blocks_async(self._h, _ptr(comp), total, _ptr(offsets), nb, C.byref(p), _ptr(dst), n, _ptr(status)))

    def decompress_exhaust(self, comp, total, offsets, nblocks, capacity, max_table_log=0, n_states=2):
        """The reference's termination rule (no stored length).  -> (dst uint8[nb, capacity], out_len[nb], status[nb])"""
        torch = _torch()
        p = self.params(capacity, max_table_log, n_states, TABLE_PER_BLOCK)
        dst = torch.empty((nblocks, capacity), dtype=torch.uint8, device=self.device)
        out_len = torch.zeros(nblocks, dtype=torch.int32, device=self.device)
        status = torch.empty(nblocks, dtype=torch.int32, device=self.device)
        self._ck(self._L.fse_b200_decompress_exhaust(self._h, _ptr(comp), total, _ptr(offsets), nblocks, C.byref(p), _ptr(dst),
                                                     _ptr(out_len), _ptr(status)))
        return dst, out_len, status

    def set_global_table(self, counts64, table_log=0):
        """-> (header bytes, effective log2)"""
        hdr = (C.c_uint8 * 512)()
        hb = C.c_size_t(512)
        l2 = C.c_uint32()
        self._ck(self._L.fse_b200_set_global_table(self._h, _ptr(counts64), table_log, hdr, C.byref(hb), C.byref(l2)))
        return bytes(hdr[: hb.value]), int(l2.value)

    def set_global_table_from_header(self, header):
        buf = (C.c_uint8 * len(header)).from_buffer_copy(header)
        l2 = C.c_uint32()
        self._ck(self._L.fse_b200_set_global_table_from_header(self._h, buf, len(header), C.byref(l2)))
        return int(l2.value)

    def global_table_covers(self, src):
        """-> number of bytes of src (device uint8) whose value the installed global table has no entry for (0 = every
        block of src can be coded with it; the encode kernels themselves do not look)"""
        unknown = C.c_uint64()
        self._ck(self._L.fse_b200_global_table_covers(self._h, _ptr(src), src.numel(), C.byref(unknown)))
        return int(unknown.value)

    # ------------------------------------------------------------------ host buffers (numpy / pinned torch)
    def compress_host(self, src, block_size, table_log=0, n_states=32, table_mode=TABLE_PER_BLOCK, dst=None, segment_size=0, flags=0):
        """src, dst: host uint8 arrays (numpy or CPU torch).  -> (dst, offsets, status, total).  A block that could not
        be coded does not raise here: its status word is negative (FSE_B200_ERR_BLOCK is reported through `status`)."""
        import numpy as np
        p = self.params(block_size, table_log, n_states, table_mode, segment_size, flags)
        n = int(src.size if hasattr(src, "size") and not callable(src.size) else src.numel())
        nb = self.num_streams(n, p)
        if dst is None:
            dst = np.empty(self.bound(n, p), dtype=np.uint8)
        offsets = np.zeros(nb + 1, dtype=np.uint64)
        status = np.zeros(max(nb, 1), dtype=np.int32)
        total = C.c_uint64()
        rc = self._L.fse_b200_compress_host(self._h, _host_ptr(src), n, C.byref(p), _host_ptr(dst), _host_len(dst),
                                            _host_ptr(offsets), _host_ptr(status), C.byref(total))
        if rc not in (0, -11):
            self._ck(rc)
        return dst, offsets, status[:nb], int(total.value)

    def decompress_host(self, comp, total, offsets, n, block_size, table_log=0, n_states=32, table_mode=TABLE_PER_BLOCK, dst=None,
                        segment_size=0, flags=0):
        """-> (dst, status); a block that failed to decode has a negative status word (no exception: inspect `status`)"""
        import numpy as np
        p = self.params(block_size, table_log, n_states, table_mode, segment_size, flags)
        nb = self.num_streams(n, p)
        if dst is None:
            dst = np.empty(max(n, 1), dtype=np.uint8)
        status = np.zeros(max(nb, 1), dtype=np.int32)
        rc = self._L.fse_b200_decompress_host(self._h, _host_ptr(comp), total, _host_ptr(offsets), nb, C.byref(p),
                                              _host_ptr(dst), n, _host_ptr(status))
        if rc not in (0, -11):
            self._ck(rc)
        return dst[:n], status[:nb]


    # ------------------------------------------------------------------ self-describing frame
    def frame_compress(self, src, block_size=65536, table_log=0, n_states=128, table_mode=TABLE_PER_BLOCK, segment_size=0, flags=0):
        """host uint8 array -> numpy uint8 frame (parameters, offsets and payload in one buffer).  Raises FseError
        (ERR_BLOCK) when a block could not be coded: such a frame could never be decoded."""
        import numpy as np
        p = self.params(block_size, table_log, n_states, table_mode, segment_size, flags)
        n = int(src.size if hasattr(src, "size") and not callable(src.size) else src.numel())
        frame = np.empty(int(self._L.fse_b200_frame_bound(n, C.byref(p))), dtype=np.uint8)
        nbytes = C.c_size_t()
        rc = self._L.fse_b200_frame_compress_host(self._h, _host_ptr(src), n, C.byref(p), _host_ptr(frame), frame.size, C.byref(nbytes))
        if rc == -11:
            raise FseError(rc, "frame_compress: at least one block could not be coded")
        self._ck(rc)
        return frame[: nbytes.value]

    def frame_info(self, frame):
        p = Params()
        n = C.c_size_t()
        rc = self._L.fse_b200_frame_info(_host_ptr(frame), _host_len(frame), C.byref(p), C.byref(n))
        if rc != 0:
            raise FseError(rc, "bad frame")
        return {"block_size": p.block_size, "table_log": p.table_log, "n_states": p.n_states, "table_mode": p.table_mode,
                "segment_size": p.segment_size, "flags": p.flags, "n": n.value}

    def frame_decompress(self, frame):
        import numpy as np
        n = self.frame_info(frame)["n"]
        dst = np.empty(max(n, 1), dtype=np.uint8)
        got = C.c_size_t()
        rc = self._L.fse_b200_frame_decompress_host(self._h, _host_ptr(frame), _host_len(frame), _host_ptr(dst), dst.size, C.byref(got))
        if rc == -11:
            raise FseError(rc, "frame_decompress: at least one block failed to decode (corrupt or truncated frame)")
        self._ck(rc)
        return dst[: got.value]


def _host_ptr(a):
    if hasattr(a, "ctypes"):
        return C.c_void_p(a.ctypes.data)
    return C.c_void_p(a.data_ptr())


def _host_len(a):
    return int(a.nbytes) if hasattr(a, "nbytes") else int(a.numel() * a.element_size())


_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


from .crate import (Histogram, NormHistogram, HistError, fse, fse_compress, fse_compress2,  # noqa: E402,F401
                    fse_decompress, fse_decompress2)
