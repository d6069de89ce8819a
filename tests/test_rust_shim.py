"""The Rust shim (rust/, source only: no Rust toolchain in this image) is kept honest mechanically: every entry point
of include/fse_b200.h is declared in rust/src/ffi.rs with the same name and argument count, the parameter struct has the
same fields in the same order, and the status constants agree."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def c_functions():
    txt = open(os.path.join(ROOT, "include", "fse_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|size_t|void|uint64_t|const char \*)\s*(fse_b200_\w+)\s*\(([^;{]*?)\)\s*;", txt, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
    return out, txt


def rust_functions():
    txt = open(os.path.join(ROOT, "rust", "src", "ffi.rs")).read()
    out = {}
    for m in re.finditer(r"pub fn (fse_b200_\w+)\s*\((.*?)\)\s*(?:->\s*[^;]+)?;", txt, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if not args else len([a for a in args.split(",") if a.strip()])
    return out, txt


def test_every_entry_point_is_declared_with_the_same_arity():
    c, _ = c_functions()
    r, _ = rust_functions()
    assert len(c) >= 37
    assert set(c) == set(r), (sorted(set(c) - set(r)), sorted(set(r) - set(c)))
    for name, n in c.items():
        assert r[name] == n, (name, n, r[name])


def test_params_struct_and_status_codes_agree():
    _, ctxt = c_functions()
    _, rtxt = rust_functions()
    cf = re.search(r"typedef struct \{(.*?)\} fse_b200_params;", ctxt, flags=re.S).group(1)
    c_fields = re.findall(r"uint32_t\s+(\w+)\s*;", cf)
    rf = re.search(r"pub struct fse_b200_params \{(.*?)\}", rtxt, flags=re.S).group(1)
    r_fields = re.findall(r"pub (\w+): u32", rf)
    assert c_fields == r_fields == ["block_size", "table_log", "n_states", "table_mode", "segment_size", "flags"]
    for name, val in re.findall(r"(FSE_B200_(?:OK|ERR_\w+)) = (-?\d+)", ctxt):
        m = re.search(r"pub const %s: c_int = (-?\d+);" % name, rtxt)
        assert m and int(m.group(1)) == int(val), name


def test_shim_mirrors_the_crate_surface():
    """the names a crate user imports exist in the shim with the crate's signatures (src/lib.rs:7, :112-248)"""
    lib = open(os.path.join(ROOT, "rust", "src", "lib.rs")).read()
    for sig in ("pub fn fse_compress(src: &[u8], dst: &mut Vec<u8>) -> (entropy_coders::NormHistogram, usize)",
                "pub fn fse_compress2(src: &[u8], dst: &mut Vec<u8>) -> usize",
                "pub fn fse_decompress(src: &[u8], dst: &mut Vec<u8>) -> Option<usize>",
                "pub fn fse_decompress2(src: &[u8], dst: &mut Vec<u8>) -> Option<usize>",
                "pub struct Histogram", "pub fn normalize(self, log2: u32)", "pub fn optimal_log2(&self) -> u32",
                "pub struct EncodeTable", "pub struct DecodeTable", "pub use entropy_coders::fse::{Decoder, Encoder}",
                "pub use entropy_coders::bitstream", "HistError"):
        assert sig in lib, sig
    assert "set_len" not in lib          # no uninitialised Vec memory is ever exposed
