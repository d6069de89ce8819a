//! Drop-in front for the FSE path of `entropy_coders` (src/lib.rs:112-248) backed by libfse_b200.so.
//!
//! `fse_compress` / `fse_compress2` / `fse_decompress*` keep the crate's signatures: they append to a
//! caller `Vec<u8>`, return bit / byte counts, map "the reference returns None" to `None` and "the
//! reference panics" to a panic.  One call = one block (block_size = src.len()), n_states 1 or 2, so the
//! bytes are the reference's bytes.  `compress_blocks` / `decompress_blocks` are the new block API.
pub mod ffi;
use ffi::*;
use std::ptr;

pub struct Gpu {
    ctx: *mut fse_b200_ctx,
}
// the context is single-owner; it may move between threads but not be shared
unsafe impl Send for Gpu {}

impl Gpu {
    pub fn new(device: i32) -> Option<Self> {
        let mut ctx = ptr::null_mut();
        let rc = unsafe { fse_b200_create(device, ptr::null_mut(), &mut ctx) };
        if rc == FSE_B200_OK { Some(Self { ctx }) } else { None }
    }

    /// New API: independent blocks, dense output + offsets.  Returns (offsets, status).
    pub fn compress_blocks(&mut self, src: &[u8], p: fse_b200_params, dst: &mut Vec<u8>) -> (Vec<u64>, Vec<i32>) {
        let nb = unsafe { fse_b200_num_blocks(src.len(), p.block_size) };
        let cap = unsafe { fse_b200_compress_blocks_bound(src.len(), &p) };
        let start = dst.len();
        dst.reserve(cap);
        let (mut off, mut st, mut total) = (vec![0u64; nb + 1], vec![0i32; nb], 0u64);
        let rc = unsafe {
            fse_b200_compress_host(self.ctx, src.as_ptr(), src.len(), &p, dst.as_mut_ptr().add(start), cap,
                                   off.as_mut_ptr(), st.as_mut_ptr(), &mut total)
        };
        assert!(rc == FSE_B200_OK || rc == FSE_B200_ERR_BLOCK, "fse_b200_compress_host failed: {}", rc);
        unsafe { dst.set_len(start + total as usize) };
        (off, st)
    }

    pub fn decompress_blocks(&mut self, comp: &[u8], offsets: &[u64], n: usize, p: fse_b200_params, dst: &mut Vec<u8>) -> Vec<i32> {
        let nb = offsets.len() - 1;
        let start = dst.len();
        dst.reserve(n);
        let mut st = vec![0i32; nb];
        let rc = unsafe {
            fse_b200_decompress_host(self.ctx, comp.as_ptr(), comp.len(), offsets.as_ptr(), nb, &p,
                                     dst.as_mut_ptr().add(start), n, st.as_mut_ptr())
        };
        assert!(rc == FSE_B200_OK || rc == FSE_B200_ERR_BLOCK, "fse_b200_decompress_host failed: {}", rc);
        unsafe { dst.set_len(start + n) };
        st
    }

    fn compress_one(&mut self, src: &[u8], dst: &mut Vec<u8>, n_states: u32) -> usize {
        // `src_iter.next().unwrap()` on an empty / too short slice: lib.rs:121,154,156
        assert!(src.len() >= n_states as usize && !src.is_empty(), "called `Option::unwrap()` on a `None` value");
        let p = fse_b200_params { block_size: src.len() as u32, table_log: 0, n_states, table_mode: 0 };
        let before = dst.len();
        let (_, st) = self.compress_blocks(src, p, dst);
        assert!(st[0] == 0, "the reference panics on this input (status {})", st[0]);
        // the crate returns the payload BIT count (writer.rs:220-221): recover it from the marker
        let last = *dst.last().unwrap();
        let header = header_len(&dst[before..]);
        (dst.len() - before - header - 1) * 8 + (8 - last.leading_zeros() as usize)
    }

    /// src/lib.rs:146-183
    pub fn fse_compress2(&mut self, src: &[u8], dst: &mut Vec<u8>) -> usize { self.compress_one(src, dst, 2) }
    /// src/lib.rs:112-143 (the NormHistogram of the return tuple is re-read from the header by the caller)
    pub fn fse_compress(&mut self, src: &[u8], dst: &mut Vec<u8>) -> usize { self.compress_one(src, dst, 1) }
    // fse_decompress / fse_decompress2 (src/lib.rs:187-248) bind fse_b200_decompress_exhaust the same way:
    // status TABLE_LOG / TOO_MANY / IO / NO_MARKER -> None, LENGTH -> the unwrap panic of lib.rs:197,224.
}

impl Drop for Gpu {
    fn drop(&mut self) { unsafe { fse_b200_destroy(self.ctx) } }
}

/// Length of the NCount header at the start of a stream: `NormHistogram::read` on the host (the crate's
/// own src/histogram.rs:436-505 stays available for this).
fn header_len(stream: &[u8]) -> usize {
    let (_, rest) = entropy_coders::NormHistogram::read(stream).expect("valid header");
    stream.len() - rest.len()
}
