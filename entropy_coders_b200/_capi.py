"""ctypes binding of libfse_b200.so (include/fse_b200.h).  No CPU fallback: importing works
anywhere, but every call needs the CUDA library and a device and raises otherwise."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("FSE_B200_LIB") or os.path.join(HERE, "libfse_b200.so")   # env: development builds

STATUS = {0: "OK", -1: "ERR_ARG", -2: "ERR_CAPACITY", -3: "ERR_TABLE_LOG", -4: "ERR_TOO_MANY", -5: "ERR_IO",
          -6: "ERR_NO_MARKER", -7: "ERR_LENGTH", -8: "ERR_PANIC", -9: "ERR_UNSUPPORTED", -10: "ERR_CUDA",
          -11: "ERR_BLOCK"}

# every symbol include/fse_b200.h declares
SYMBOLS = [
    "fse_b200_create", "fse_b200_destroy", "fse_b200_last_error", "fse_b200_version", "fse_b200_launch_count",
    "fse_b200_sync", "fse_b200_set_timing", "fse_b200_get_timing", "fse_b200_compress_bound", "fse_b200_compress_blocks_bound", "fse_b200_num_blocks", "fse_b200_num_streams",
    "fse_b200_histogram_blocks", "fse_b200_histogram_global", "fse_b200_normalize", "fse_b200_normalize_zstd", "fse_b200_ncount_write",
    "fse_b200_ncount_read", "fse_b200_build_encode_tables", "fse_b200_build_decode_tables",
    "fse_b200_compress_blocks", "fse_b200_compress_blocks_async", "fse_b200_decompress_blocks",
    "fse_b200_decompress_blocks_async", "fse_b200_decompress_exhaust", "fse_b200_set_global_table", "fse_b200_set_global_table_from_header",
    "fse_b200_global_table_covers",
    "fse_b200_compress_host", "fse_b200_decompress_host", "fse_b200_generate",
    "fse_b200_bitstack_write", "fse_b200_bitstack_read", "fse_b200_bitstream_read",
    "fse_b200_frame_bound", "fse_b200_frame_compress_host", "fse_b200_frame_info", "fse_b200_frame_decompress_host",
]


class Params(C.Structure):
    _fields_ = [("block_size", C.c_uint32), ("table_log", C.c_uint32), ("n_states", C.c_uint32),
                ("table_mode", C.c_uint32), ("segment_size", C.c_uint32), ("flags", C.c_uint32)]


class FseError(RuntimeError):
    def __init__(self, code, msg=""):
        self.code = code
        super().__init__("%s (%d)%s" % (STATUS.get(code, "?"), code, (": " + msg) if msg else ""))


_lib = None


def lib():
    """Loads the CUDA library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a); there is no CPU fallback" % SO_PATH)
    L = C.CDLL(SO_PATH)
    vp, sz, u32, u64, i32 = C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint64, C.c_int
    PP = C.POINTER(Params)
    L.fse_b200_create.argtypes = [i32, vp, C.POINTER(vp)]
    L.fse_b200_destroy.argtypes = [vp]
    L.fse_b200_destroy.restype = None
    L.fse_b200_last_error.argtypes = [vp]
    L.fse_b200_last_error.restype = C.c_char_p
    L.fse_b200_version.restype = C.c_char_p
    L.fse_b200_launch_count.argtypes = [vp]
    L.fse_b200_launch_count.restype = u64
    L.fse_b200_sync.argtypes = [vp]
    L.fse_b200_set_timing.argtypes = [vp, i32]
    L.fse_b200_get_timing.argtypes = [vp, vp, vp]
    L.fse_b200_compress_bound.argtypes = [sz]
    L.fse_b200_compress_bound.restype = sz
    L.fse_b200_compress_blocks_bound.argtypes = [sz, PP]
    L.fse_b200_compress_blocks_bound.restype = sz
    L.fse_b200_num_blocks.argtypes = [sz, u32]
    L.fse_b200_num_blocks.restype = sz
    L.fse_b200_num_streams.argtypes = [sz, PP]
    L.fse_b200_num_streams.restype = sz
    L.fse_b200_histogram_blocks.argtypes = [vp, vp, sz, u32, vp, vp]
    L.fse_b200_histogram_global.argtypes = [vp, vp, sz, vp]
    L.fse_b200_normalize.argtypes = [vp, vp, sz, u32, vp, vp, vp, vp]
    L.fse_b200_normalize_zstd.argtypes = [vp, vp, sz, u32, i32, vp, vp, vp, vp]
    L.fse_b200_ncount_write.argtypes = [vp, vp, vp, vp, sz, vp, sz, vp, vp]
    L.fse_b200_ncount_read.argtypes = [vp, vp, sz, vp, sz, vp, vp, vp, vp, vp]
    L.fse_b200_build_encode_tables.argtypes = [vp, vp, vp, vp, sz, u32, vp, vp, vp, vp]
    L.fse_b200_build_decode_tables.argtypes = [vp, vp, vp, vp, sz, u32, vp, vp]
    L.fse_b200_compress_blocks.argtypes = [vp, vp, sz, PP, vp, sz, vp, vp, C.POINTER(u64)]
    L.fse_b200_compress_blocks_async.argtypes = [vp, vp, sz, PP, vp, sz, vp, vp]
    L.fse_b200_decompress_blocks.argtypes = [vp, vp, sz, vp, sz, PP, vp, sz, vp]
    L.fse_b200_decompress_blocks_async.argtypes = [vp, vp, sz, vp, sz, PP, vp, sz, vp]
    L.fse_b200_decompress_exhaust.argtypes = [vp, vp, sz, vp, sz, PP, vp, vp, vp]
    L.fse_b200_set_global_table.argtypes = [vp, vp, u32, vp, C.POINTER(sz), C.POINTER(u32)]
    L.fse_b200_set_global_table_from_header.argtypes = [vp, vp, sz, C.POINTER(u32)]
    L.fse_b200_global_table_covers.argtypes = [vp, vp, sz, C.POINTER(u64)]
    L.fse_b200_compress_host.argtypes = [vp, vp, sz, PP, vp, sz, vp, vp, C.POINTER(u64)]
    L.fse_b200_decompress_host.argtypes = [vp, vp, sz, vp, sz, PP, vp, sz, vp]
    L.fse_b200_generate.argtypes = [vp, i32, u64, u64, vp, sz]
    L.fse_b200_bitstack_write.argtypes = [vp, vp, vp, sz, i32, vp, sz, C.POINTER(u64)]
    L.fse_b200_bitstack_read.argtypes = [vp, vp, sz, vp, sz, vp, C.POINTER(C.c_int32)]
    L.fse_b200_bitstream_read.argtypes = [vp, vp, sz, u64, vp, sz, vp, C.POINTER(C.c_int32)]
    L.fse_b200_frame_bound.argtypes = [sz, PP]
    L.fse_b200_frame_bound.restype = sz
    L.fse_b200_frame_compress_host.argtypes = [vp, vp, sz, PP, vp, sz, C.POINTER(sz)]
    L.fse_b200_frame_info.argtypes = [vp, sz, PP, C.POINTER(sz)]
    L.fse_b200_frame_decompress_host.argtypes = [vp, vp, sz, vp, sz, C.POINTER(sz)]
    _lib = L
    return L
