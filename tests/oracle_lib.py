"""ctypes binding to the CPU oracle (oracle/_build/libfse_oracle.so).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_build", "libfse_oracle.so")


def build(force=False):
    src = os.path.join(ROOT, "oracle", "fse_oracle.c")
    hdr = os.path.join(ROOT, "oracle", "fse_oracle.h")
    if (not force and os.path.exists(SO)
            and os.path.getmtime(SO) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return SO
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    return SO


class Hist(C.Structure):
    _fields_ = [("table", C.c_uint64 * 256), ("size", C.c_uint64), ("table_len", C.c_uint32)]


class Norm(C.Structure):
    _fields_ = [("table", C.c_int32 * 256), ("log2", C.c_uint32), ("table_len", C.c_uint32)]


class SymTT(C.Structure):
    _fields_ = [("bits", C.c_uint32), ("find_state", C.c_int32)]


class EncTable(C.Structure):
    _fields_ = [("table_log", C.c_uint32), ("table", C.c_uint16 * 32768),
                ("symbol_tt", SymTT * 256), ("symbols", C.c_uint8 * 32768)]


class DecEntry(C.Structure):
    _fields_ = [("new_state", C.c_uint16), ("symbol", C.c_uint8), ("num_bits", C.c_uint8)]


class DecTable(C.Structure):
    _fields_ = [("table_log", C.c_uint32), ("fast_mode", C.c_int), ("table", DecEntry * 32768)]


class BlockParams(C.Structure):
    _fields_ = [("block_size", C.c_uint32), ("table_log", C.c_uint32), ("n_states", C.c_uint32),
                ("threads", C.c_uint32), ("use_ref2", C.c_uint32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        L = _lib
        u8p, sz = C.c_void_p, C.c_size_t
        L.fse_or_histogram.argtypes = [u8p, sz, C.POINTER(Hist)]
        L.fse_or_histogram.restype = None
        L.fse_or_optimal_log2.argtypes = [C.POINTER(Hist), C.POINTER(C.c_uint32)]
        L.fse_or_normalize.argtypes = [C.POINTER(Hist), C.c_uint32, C.POINTER(Norm)]
        L.fse_or_normalize_zstd.argtypes = [C.POINTER(Hist), C.c_uint32, C.c_int, C.POINTER(Norm)]
        L.fse_or_norm_new.argtypes = [u8p, sz, C.POINTER(Norm)]
        L.fse_or_write_bound.argtypes = [C.POINTER(Norm)]
        L.fse_or_write_bound.restype = sz
        L.fse_or_ncount_write.argtypes = [C.POINTER(Norm), u8p, sz, C.POINTER(sz)]
        L.fse_or_ncount_write.restype = C.c_long
        L.fse_or_ncount_read.argtypes = [u8p, sz, C.POINTER(Norm), C.POINTER(sz)]
        L.fse_or_norm_try_from.argtypes = [C.POINTER(C.c_int32), C.POINTER(Norm)]
        L.fse_or_symbol_count.argtypes = [C.POINTER(C.c_int32)]
        L.fse_or_symbol_count.restype = C.c_uint32
        L.fse_or_table_step.argtypes = [sz]
        L.fse_or_table_step.restype = sz
        L.fse_or_compress_bound.argtypes = [sz]
        L.fse_or_compress_bound.restype = sz
        L.fse_or_enc_table_build.argtypes = [C.POINTER(Norm), C.POINTER(EncTable)]
        L.fse_or_dec_table_build.argtypes = [C.POINTER(Norm), C.POINTER(DecTable)]
        L.fse_or_encode_payload.argtypes = [C.POINTER(EncTable), u8p, sz, C.c_uint, u8p, sz, C.POINTER(sz)]
        L.fse_or_encode_payload.restype = C.c_long
        L.fse_or_compress_n.argtypes = [u8p, sz, C.c_uint32, C.c_uint, u8p, sz, C.POINTER(sz), C.POINTER(sz)]
        L.fse_or_compress_n.restype = C.c_long
        L.fse_or_decode_payload_exhaust.argtypes = [C.POINTER(DecTable), u8p, sz, C.c_uint, u8p, sz]
        L.fse_or_decode_payload_exhaust.restype = C.c_long
        L.fse_or_decode_payload_len.argtypes = [C.POINTER(DecTable), u8p, sz, C.c_uint, u8p, sz]
        L.fse_or_decompress_n_exhaust.argtypes = [u8p, sz, C.c_uint, u8p, sz]
        L.fse_or_decompress_n_exhaust.restype = C.c_long
        L.fse_or_decompress_n_len.argtypes = [u8p, sz, C.c_uint, u8p, sz]
        L.fse_or_ref_compress2.argtypes = [u8p, sz, u8p, sz]
        L.fse_or_ref_compress2.restype = C.c_long
        L.fse_or_ref_decompress2.argtypes = [u8p, sz, u8p, sz]
        L.fse_or_ref_decompress2.restype = C.c_long
        L.fse_or_compress_blocks.argtypes = [u8p, sz, C.POINTER(BlockParams), u8p, sz, u8p, u8p]
        L.fse_or_decompress_blocks.argtypes = [u8p, sz, u8p, sz, C.POINTER(BlockParams), u8p, sz, u8p]
        L.fse_or_generate.argtypes = [C.c_int, C.c_uint64, C.c_uint64, u8p, sz]
        L.fse_or_generate.restype = None
        L.fse_or_gen_lut.argtypes = [C.c_int, u8p]
        L.fse_or_gen_lut.restype = sz
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def as_u8(data):
    if isinstance(data, np.ndarray):
        return np.ascontiguousarray(data, dtype=np.uint8)
    return np.frombuffer(bytes(data), dtype=np.uint8).copy()


GEN = {"geo": 0, "text": 1, "few": 2, "uniform": 3}


def generate(kind, seed, n, first_index=0):
    out = np.empty(n, dtype=np.uint8)
    lib().fse_or_generate(GEN[kind], seed, first_index, _p(out), n)
    return out


def histogram(data):
    d = as_u8(data)
    h = Hist()
    lib().fse_or_histogram(_p(d), d.size, C.byref(h))
    return h


def normalize(h, log2):
    n = Norm()
    rc = lib().fse_or_normalize(C.byref(h), log2, C.byref(n))
    return rc, n


def normalize_zstd(h, log2, use_low_prob_count=True):
    n = Norm()
    rc = lib().fse_or_normalize_zstd(C.byref(h), log2, 1 if use_low_prob_count else 0, C.byref(n))
    return rc, n


def hist_from_counts(counts):
    h = Hist()
    nz = [i for i, c in enumerate(counts) if c]
    for i, c in enumerate(counts):
        h.table[i] = int(c)
    h.size = int(sum(counts))
    h.table_len = (nz[-1] + 1) if nz else 1
    return h


def optimal_log2(h):
    v = C.c_uint32()
    rc = lib().fse_or_optimal_log2(C.byref(h), C.byref(v))
    return rc, v.value


def norm_from_table(table, log2=None):
    arr = (C.c_int32 * 256)(*list(table))
    n = Norm()
    rc = lib().fse_or_norm_try_from(arr, C.byref(n))
    assert rc == 0, rc
    return n


def ncount_write(n):
    cap = 1024
    out = np.zeros(cap, dtype=np.uint8)
    bits = C.c_size_t()
    ln = lib().fse_or_ncount_write(C.byref(n), _p(out), cap, C.byref(bits))
    assert ln >= 0, ln
    return out[:ln].tobytes(), bits.value


def ncount_read(data):
    d = as_u8(data)
    n = Norm()
    consumed = C.c_size_t()
    rc = lib().fse_or_ncount_read(_p(d), d.size, C.byref(n), C.byref(consumed))
    return rc, n, consumed.value


def enc_table(n):
    t = EncTable()
    rc = lib().fse_or_enc_table_build(C.byref(n), C.byref(t))
    assert rc == 0, rc
    return t


def dec_table(n):
    t = DecTable()
    rc = lib().fse_or_dec_table_build(C.byref(n), C.byref(t))
    assert rc == 0, rc
    return t


def compress_n(data, table_log=0, n_states=2):
    """-> (bytes, header_bytes, payload_bits) or raises ValueError(status)."""
    d = as_u8(data)
    cap = lib().fse_or_compress_bound(d.size) + 64 * n_states
    out = np.zeros(cap, dtype=np.uint8)
    hb, pb = C.c_size_t(), C.c_size_t()
    ln = lib().fse_or_compress_n(_p(d), d.size, table_log, n_states, _p(out), cap, C.byref(hb), C.byref(pb))
    if ln < 0:
        raise ValueError(ln)
    return out[:ln].tobytes(), hb.value, pb.value


def encode_payload(t, data, n_states):
    d = as_u8(data)
    cap = lib().fse_or_compress_bound(d.size) + 64 * n_states
    out = np.zeros(cap, dtype=np.uint8)
    pb = C.c_size_t()
    ln = lib().fse_or_encode_payload(C.byref(t), _p(d), d.size, n_states, _p(out), cap, C.byref(pb))
    if ln < 0:
        raise ValueError(ln)
    return out[:ln].tobytes(), pb.value


def decompress_n_exhaust(comp, n_states, cap):
    c = as_u8(comp)
    out = np.zeros(cap, dtype=np.uint8)
    ln = lib().fse_or_decompress_n_exhaust(_p(c), c.size, n_states, _p(out), cap)
    if ln < 0:
        raise ValueError(ln)
    return out[:ln].tobytes()


def decompress_n_len(comp, n_states, n_out):
    c = as_u8(comp)
    out = np.zeros(n_out, dtype=np.uint8)
    rc = lib().fse_or_decompress_n_len(_p(c), c.size, n_states, _p(out), n_out)
    if rc < 0:
        raise ValueError(rc)
    return out.tobytes()


def ref_compress2(data):
    d = as_u8(data)
    cap = lib().fse_or_compress_bound(d.size)
    out = np.zeros(cap, dtype=np.uint8)
    ln = lib().fse_or_ref_compress2(_p(d), d.size, _p(out), cap)
    if ln < 0:
        raise ValueError(ln)
    return out[:ln].tobytes()


def ref_decompress2(comp, cap):
    c = as_u8(comp)
    out = np.zeros(cap, dtype=np.uint8)
    ln = lib().fse_or_ref_decompress2(_p(c), c.size, _p(out), cap)
    if ln < 0:
        raise ValueError(ln)
    return out[:ln].tobytes()


def compress_blocks(data, block_size, table_log=0, n_states=32, threads=1, use_ref2=0):
    """-> (scratch[nblocks, stride] u8, sizes u64[nblocks], status i32[nblocks])"""
    d = as_u8(data)
    nb = (d.size + block_size - 1) // block_size
    stride = lib().fse_or_compress_bound(block_size) + 64 * n_states
    out = np.zeros((nb, stride), dtype=np.uint8)
    sizes = np.zeros(nb, dtype=np.uint64)
    status = np.zeros(nb, dtype=np.int32)
    p = BlockParams(block_size, table_log, n_states, threads, use_ref2)
    lib().fse_or_compress_blocks(_p(d), d.size, C.byref(p), _p(out), stride, _p(sizes), _p(status))
    return out, sizes, status


def decompress_blocks(scratch, sizes, n, block_size, n_states=32, threads=1, use_ref2=0):
    nb, stride = scratch.shape
    out = np.zeros(n, dtype=np.uint8)
    status = np.zeros(nb, dtype=np.int32)
    p = BlockParams(block_size, 0, n_states, threads, use_ref2)
    lib().fse_or_decompress_blocks(_p(scratch), stride, _p(sizes), nb, C.byref(p), _p(out), n, _p(status))
    return out, status
