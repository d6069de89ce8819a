//! Raw bindings of include/fse_b200.h: every entry point, in the order of the header (what `bindgen` emits for it).
//! tests/test_rust_shim.py checks names and argument counts against the header on every CPU test run.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct fse_b200_params {
    pub block_size: u32,
    pub table_log: u32,
    pub n_states: u32,
    pub table_mode: u32,
    pub segment_size: u32,
    pub flags: u32,
}
/// src/fse.rs:80-84
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct fse_b200_symbol_transform {
    pub bits: u32,
    pub find_state: i32,
}
/// src/fse.rs:260-265
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct fse_b200_decode_transform {
    pub new_state: u16,
    pub symbol: u8,
    pub num_bits: u8,
}
#[repr(C)]
pub struct fse_b200_ctx {
    _private: [u8; 0],
}

pub const FSE_B200_OK: c_int = 0;
pub const FSE_B200_ERR_ARG: c_int = -1;
pub const FSE_B200_ERR_CAPACITY: c_int = -2;
pub const FSE_B200_ERR_TABLE_LOG: c_int = -3;
pub const FSE_B200_ERR_TOO_MANY: c_int = -4;
pub const FSE_B200_ERR_IO: c_int = -5;
pub const FSE_B200_ERR_NO_MARKER: c_int = -6;
pub const FSE_B200_ERR_LENGTH: c_int = -7;
pub const FSE_B200_ERR_PANIC: c_int = -8;
pub const FSE_B200_ERR_UNSUPPORTED: c_int = -9;
pub const FSE_B200_ERR_CUDA: c_int = -10;
pub const FSE_B200_ERR_BLOCK: c_int = -11;
pub const FSE_B200_TABLE_PER_BLOCK: u32 = 0;
pub const FSE_B200_TABLE_GLOBAL: u32 = 1;
pub const FSE_B200_FLAG_RAW_IF_EXPANDS: u32 = 1;
pub const FSE_B200_NUM_KERNELS: usize = 5;

extern "C" {
    pub fn fse_b200_create(device: c_int, stream: *mut c_void, out: *mut *mut fse_b200_ctx) -> c_int;
    pub fn fse_b200_destroy(ctx: *mut fse_b200_ctx);
    pub fn fse_b200_last_error(ctx: *const fse_b200_ctx) -> *const c_char;
    pub fn fse_b200_version() -> *const c_char;
    pub fn fse_b200_launch_count(ctx: *const fse_b200_ctx) -> u64;
    pub fn fse_b200_sync(ctx: *mut fse_b200_ctx) -> c_int;
    pub fn fse_b200_set_timing(ctx: *mut fse_b200_ctx, enable: c_int) -> c_int;
    pub fn fse_b200_get_timing(ctx: *mut fse_b200_ctx, ms_total: *mut f64, count: *mut u64) -> c_int;

    pub fn fse_b200_compress_bound(size: usize) -> usize;
    pub fn fse_b200_compress_blocks_bound(n: usize, p: *const fse_b200_params) -> usize;
    pub fn fse_b200_num_blocks(n: usize, block_size: u32) -> usize;
    pub fn fse_b200_num_streams(n: usize, p: *const fse_b200_params) -> usize;

    pub fn fse_b200_histogram_blocks(
        ctx: *mut fse_b200_ctx, d_src: *const u8, n: usize, block_size: u32, d_counts: *mut u32, d_table_len: *mut u32,
    ) -> c_int;
    pub fn fse_b200_histogram_global(ctx: *mut fse_b200_ctx, d_src: *const u8, n: usize, d_counts64: *mut u64) -> c_int;
    pub fn fse_b200_normalize(
        ctx: *mut fse_b200_ctx, d_counts64: *const u64, ntables: usize, table_log: u32, d_norm: *mut i32, d_log2: *mut u32,
        d_table_len: *mut u32, d_status: *mut i32,
    ) -> c_int;
    pub fn fse_b200_normalize_zstd(
        ctx: *mut fse_b200_ctx, d_counts64: *const u64, ntables: usize, table_log: u32, use_low_prob_count: c_int,
        d_norm: *mut i32, d_log2: *mut u32, d_table_len: *mut u32, d_status: *mut i32,
    ) -> c_int;
    pub fn fse_b200_ncount_write(
        ctx: *mut fse_b200_ctx, d_norm: *const i32, d_log2: *const u32, d_table_len: *const u32, ntables: usize,
        d_out: *mut u8, stride: usize, d_bytes: *mut u32, d_bits: *mut u32,
    ) -> c_int;
    pub fn fse_b200_ncount_read(
        ctx: *mut fse_b200_ctx, d_in: *const u8, stride: usize, d_len: *const u32, ntables: usize, d_norm: *mut i32,
        d_log2: *mut u32, d_table_len: *mut u32, d_consumed: *mut u32, d_status: *mut i32,
    ) -> c_int;
    pub fn fse_b200_build_encode_tables(
        ctx: *mut fse_b200_ctx, d_norm: *const i32, d_log2: *const u32, d_table_len: *const u32, ntables: usize,
        max_table_log: u32, d_table: *mut u16, d_symbol_tt: *mut fse_b200_symbol_transform, d_symbols: *mut u8,
        d_status: *mut i32,
    ) -> c_int;
    pub fn fse_b200_build_decode_tables(
        ctx: *mut fse_b200_ctx, d_norm: *const i32, d_log2: *const u32, d_table_len: *const u32, ntables: usize,
        max_table_log: u32, d_table: *mut fse_b200_decode_transform, d_status: *mut i32,
    ) -> c_int;

    pub fn fse_b200_bitstack_write(
        ctx: *mut fse_b200_ctx, d_vals: *const u32, d_bits: *const u8, n: usize, mark: c_int, d_out: *mut u8, out_cap: usize,
        h_nbits: *mut u64,
    ) -> c_int;
    pub fn fse_b200_bitstack_read(
        ctx: *mut fse_b200_ctx, d_in: *const u8, nbytes: usize, d_bits: *const u8, n: usize, d_vals: *mut u32,
        h_status: *mut i32,
    ) -> c_int;
    pub fn fse_b200_bitstream_read(
        ctx: *mut fse_b200_ctx, d_in: *const u8, nbytes: usize, total_bits: u64, d_bits: *const u8, n: usize,
        d_vals: *mut u32, h_status: *mut i32,
    ) -> c_int;

    pub fn fse_b200_compress_blocks(
        ctx: *mut fse_b200_ctx, d_src: *const u8, n: usize, p: *const fse_b200_params, d_dst: *mut u8, dst_cap: usize,
        d_offsets: *mut u64, d_status: *mut i32, h_total: *mut u64,
    ) -> c_int;
    pub fn fse_b200_compress_blocks_async(
        ctx: *mut fse_b200_ctx, d_src: *const u8, n: usize, p: *const fse_b200_params, d_dst: *mut u8, dst_cap: usize,
        d_offsets: *mut u64, d_status: *mut i32,
    ) -> c_int;
    pub fn fse_b200_decompress_blocks(
        ctx: *mut fse_b200_ctx, d_comp: *const u8, comp_bytes: usize, d_offsets: *const u64, nblocks: usize,
        p: *const fse_b200_params, d_dst: *mut u8, n: usize, d_status: *mut i32,
    ) -> c_int;
    pub fn fse_b200_decompress_blocks_async(
        ctx: *mut fse_b200_ctx, d_comp: *const u8, comp_bytes: usize, d_offsets: *const u64, nblocks: usize,
        p: *const fse_b200_params, d_dst: *mut u8, n: usize, d_status: *mut i32,
    ) -> c_int;
    pub fn fse_b200_decompress_exhaust(
        ctx: *mut fse_b200_ctx, d_comp: *const u8, comp_bytes: usize, d_offsets: *const u64, nblocks: usize,
        p: *const fse_b200_params, d_dst: *mut u8, d_out_len: *mut u32, d_status: *mut i32,
    ) -> c_int;
    pub fn fse_b200_set_global_table(
        ctx: *mut fse_b200_ctx, d_counts64: *const u64, table_log: u32, h_header: *mut u8, h_header_bytes: *mut usize,
        h_log2: *mut u32,
    ) -> c_int;
    pub fn fse_b200_set_global_table_from_header(
        ctx: *mut fse_b200_ctx, h_header: *const u8, header_bytes: usize, h_log2: *mut u32,
    ) -> c_int;
    pub fn fse_b200_global_table_covers(
        ctx: *mut fse_b200_ctx, d_src: *const u8, n: usize, h_unknown: *mut u64,
    ) -> c_int;

    pub fn fse_b200_compress_host(
        ctx: *mut fse_b200_ctx, h_src: *const u8, n: usize, p: *const fse_b200_params, h_dst: *mut u8, dst_cap: usize,
        h_offsets: *mut u64, h_status: *mut i32, h_total: *mut u64,
    ) -> c_int;
    pub fn fse_b200_decompress_host(
        ctx: *mut fse_b200_ctx, h_comp: *const u8, comp_bytes: usize, h_offsets: *const u64, nblocks: usize,
        p: *const fse_b200_params, h_dst: *mut u8, n: usize, h_status: *mut i32,
    ) -> c_int;

    pub fn fse_b200_frame_bound(n: usize, p: *const fse_b200_params) -> usize;
    pub fn fse_b200_frame_compress_host(
        ctx: *mut fse_b200_ctx, h_src: *const u8, n: usize, p: *const fse_b200_params, h_frame: *mut u8, frame_cap: usize,
        h_frame_bytes: *mut usize,
    ) -> c_int;
    pub fn fse_b200_frame_info(h_frame: *const u8, frame_bytes: usize, p_out: *mut fse_b200_params, n_out: *mut usize) -> c_int;
    pub fn fse_b200_frame_decompress_host(
        ctx: *mut fse_b200_ctx, h_frame: *const u8, frame_bytes: usize, h_dst: *mut u8, dst_cap: usize, h_n: *mut usize,
    ) -> c_int;

    pub fn fse_b200_generate(ctx: *mut fse_b200_ctx, kind: c_int, seed: u64, first_index: u64, d_dst: *mut u8, n: usize) -> c_int;
}

// the few CUDA runtime calls the device-pointer wrappers need (libcudart)
extern "C" {
    pub fn cudaMalloc(p: *mut *mut c_void, bytes: usize) -> c_int;
    pub fn cudaFree(p: *mut c_void) -> c_int;
    pub fn cudaMemcpy(dst: *mut c_void, src: *const c_void, bytes: usize, kind: c_int) -> c_int;
}
pub const CUDA_MEMCPY_H2D: c_int = 1;
pub const CUDA_MEMCPY_D2H: c_int = 2;
