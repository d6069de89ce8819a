//! Raw bindings of include/fse_b200.h (what `bindgen` would emit for the entry points used below).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
#[derive(Clone, Copy)]
pub struct fse_b200_params {
    pub block_size: u32,
    pub table_log: u32,
    pub n_states: u32,
    pub table_mode: u32,
}
#[repr(C)]
pub struct fse_b200_ctx {
    _private: [u8; 0],
}

pub const FSE_B200_OK: c_int = 0;
pub const FSE_B200_ERR_TABLE_LOG: c_int = -3;
pub const FSE_B200_ERR_TOO_MANY: c_int = -4;
pub const FSE_B200_ERR_IO: c_int = -5;
pub const FSE_B200_ERR_NO_MARKER: c_int = -6;
pub const FSE_B200_ERR_LENGTH: c_int = -7;
pub const FSE_B200_ERR_BLOCK: c_int = -11;

extern "C" {
    pub fn fse_b200_create(device: c_int, stream: *mut c_void, out: *mut *mut fse_b200_ctx) -> c_int;
    pub fn fse_b200_destroy(ctx: *mut fse_b200_ctx);
    pub fn fse_b200_last_error(ctx: *const fse_b200_ctx) -> *const c_char;
    pub fn fse_b200_compress_bound(size: usize) -> usize;
    pub fn fse_b200_compress_blocks_bound(n: usize, p: *const fse_b200_params) -> usize;
    pub fn fse_b200_num_blocks(n: usize, block_size: u32) -> usize;
    pub fn fse_b200_compress_host(
        ctx: *mut fse_b200_ctx, h_src: *const u8, n: usize, p: *const fse_b200_params, h_dst: *mut u8,
        dst_cap: usize, h_offsets: *mut u64, h_status: *mut i32, h_total: *mut u64,
    ) -> c_int;
    pub fn fse_b200_decompress_host(
        ctx: *mut fse_b200_ctx, h_comp: *const u8, comp_bytes: usize, h_offsets: *const u64, nblocks: usize,
        p: *const fse_b200_params, h_dst: *mut u8, n: usize, h_status: *mut i32,
    ) -> c_int;
    // device-pointer entry points (histogram_blocks, normalize, ncount_write/read, build_*_tables,
    // compress_blocks, decompress_blocks, decompress_exhaust, set_global_table ...) bind the same way.
}
