//! Drop-in front for the FSE path of `entropy_coders` backed by libfse_b200.so (include/fse_b200.h).
//!
//! The crate's `pub` items are the boundary (src/lib.rs:7, :112-248; src/histogram.rs; src/fse.rs; src/bitstream):
//!
//! * free functions `fse_compress`, `fse_compress2`, `fse_decompress`, `fse_decompress2` with the crate's exact
//!   signatures: they append to a caller `Vec<u8>`, return bit / byte counts, map "the reference returns None" to
//!   `None` and "the reference panics" to a panic.  One call = one block, 1 or 2 states: the bytes are the crate's;
//! * `Histogram`, `NormHistogram`, `HistError`, `fse::{EncodeTable, DecodeTable}`: same names and methods, the
//!   arithmetic runs in the CUDA kernels;
//! * `fse::{Encoder, Decoder}` and `bitstream::*` are per-symbol host objects: they are re-exported from the crate
//!   itself (their arithmetic is what the kernels run per lane);
//! * `Gpu::compress_blocks` / `decompress_blocks`: the block API the crate does not have.
//!
//! There is no CPU fallback: `Gpu::new` fails without a CUDA device.  Not compiled where it was written (no Rust
//! toolchain in that image); the C++ mirror include/entropy_coders.hpp is the compiled and tested twin.
pub mod ffi;
use ffi::*;
use std::cell::RefCell;
use std::os::raw::c_void;
use std::ptr;

pub use entropy_coders::bitstream;
pub use entropy_coders::histogram::HistError;

pub const TABLE_LOG_MIN: u32 = 5; // src/lib.rs:9
pub const TABLE_LOG_MAX: u32 = 15; // src/lib.rs:10
pub const TABLE_LOG_DEFAULT: u32 = 11; // src/lib.rs:12

pub struct Gpu {
    ctx: *mut fse_b200_ctx,
}
// the context is single-owner; it may move between threads but not be shared
unsafe impl Send for Gpu {}

thread_local! {
    static DEFAULT: RefCell<Option<Gpu>> = RefCell::new(None);
}
/// The free functions and the table types run on a per-thread default context (device 0).
fn with_gpu<R>(f: impl FnOnce(&mut Gpu) -> R) -> R {
    DEFAULT.with(|g| {
        let mut g = g.borrow_mut();
        if g.is_none() {
            *g = Some(Gpu::new(0).expect("fse_b200_create failed: a CUDA device is required (no CPU fallback)"));
        }
        f(g.as_mut().unwrap())
    })
}

/// A device array with host copies in and out.
struct Dev<T: Copy + Default> {
    p: *mut T,
    n: usize,
}
impl<T: Copy + Default> Dev<T> {
    fn new(n: usize) -> Self {
        let mut p: *mut c_void = ptr::null_mut();
        let rc = unsafe { cudaMalloc(&mut p, n.max(1) * std::mem::size_of::<T>()) };
        assert!(rc == 0, "cudaMalloc failed: {}", rc);
        Self { p: p as *mut T, n }
    }
    fn from(h: &[T]) -> Self {
        let d = Self::new(h.len());
        if !h.is_empty() {
            let rc = unsafe { cudaMemcpy(d.p as *mut c_void, h.as_ptr() as *const c_void, h.len() * std::mem::size_of::<T>(), CUDA_MEMCPY_H2D) };
            assert!(rc == 0, "cudaMemcpy failed: {}", rc);
        }
        d
    }
    fn host(&self) -> Vec<T> {
        let mut v = vec![T::default(); self.n];
        if self.n > 0 {
            let rc = unsafe { cudaMemcpy(v.as_mut_ptr() as *mut c_void, self.p as *const c_void, self.n * std::mem::size_of::<T>(), CUDA_MEMCPY_D2H) };
            assert!(rc == 0, "cudaMemcpy failed: {}", rc);
        }
        v
    }
}
impl<T: Copy + Default> Drop for Dev<T> {
    fn drop(&mut self) {
        unsafe { cudaFree(self.p as *mut c_void) };
    }
}

fn ck(rc: i32, what: &str) {
    assert!(rc == FSE_B200_OK, "{} failed: status {}", what, rc);
}

impl Gpu {
    pub fn new(device: i32) -> Option<Self> {
        let mut ctx = ptr::null_mut();
        let rc = unsafe { fse_b200_create(device, ptr::null_mut(), &mut ctx) };
        if rc == FSE_B200_OK { Some(Self { ctx }) } else { None }
    }

    /// New API: independent blocks, dense output + offsets.  Returns (offsets, status) or the library status.
    /// Blocks that could not be coded have a negative status word; `dst` then holds no bytes for them.
    pub fn compress_blocks(&mut self, src: &[u8], p: fse_b200_params, dst: &mut Vec<u8>) -> Result<(Vec<u64>, Vec<i32>), i32> {
        let ns = unsafe { fse_b200_num_streams(src.len(), &p) };
        let cap = unsafe { fse_b200_compress_blocks_bound(src.len(), &p) };
        let start = dst.len();
        dst.resize(start + cap, 0);
        let (mut off, mut st, mut total) = (vec![0u64; ns + 1], vec![0i32; ns], 0u64);
        let rc = unsafe {
            fse_b200_compress_host(self.ctx, src.as_ptr(), src.len(), &p, dst.as_mut_ptr().add(start), cap, off.as_mut_ptr(),
                                   st.as_mut_ptr(), &mut total)
        };
        if rc != FSE_B200_OK && rc != FSE_B200_ERR_BLOCK {
            dst.truncate(start);
            return Err(rc);
        }
        dst.truncate(start + total as usize);
        Ok((off, st))
    }

    /// Returns the per-stream status words.  `dst` grows by `n` zero-initialised bytes before the call, so a block that
    /// fails to decode (negative status) leaves zeros, never uninitialised memory.
    pub fn decompress_blocks(&mut self, comp: &[u8], offsets: &[u64], n: usize, p: fse_b200_params, dst: &mut Vec<u8>) -> Result<Vec<i32>, i32> {
        if offsets.is_empty() {
            return Err(FSE_B200_ERR_ARG);
        }
        let ns = offsets.len() - 1;
        let start = dst.len();
        dst.resize(start + n, 0);
        let mut st = vec![0i32; ns];
        let rc = unsafe {
            fse_b200_decompress_host(self.ctx, comp.as_ptr(), comp.len(), offsets.as_ptr(), ns, &p, dst.as_mut_ptr().add(start), n,
                                     st.as_mut_ptr())
        };
        if rc != FSE_B200_OK && rc != FSE_B200_ERR_BLOCK {
            dst.truncate(start);
            return Err(rc);
        }
        Ok(st)
    }

    /// header || payload of one slice with `n_states` states; returns the payload bit count (writer.rs:220-221)
    fn compress_one(&mut self, src: &[u8], dst: &mut Vec<u8>, n_states: u32) -> usize {
        // `src_iter.next().unwrap()` on an empty / too short slice: lib.rs:121,154,156
        assert!(src.len() >= n_states as usize && !src.is_empty(), "called `Option::unwrap()` on a `None` value");
        assert!(src.len() <= u32::MAX as usize, "Data vector is too long"); // histogram.rs:19
        let p = fse_b200_params { block_size: src.len() as u32, table_log: 0, n_states, ..Default::default() };
        let before = dst.len();
        let (_, st) = self.compress_blocks(src, p, dst).expect("fse_b200_compress_host");
        assert!(st[0] == 0, "the reference panics on this input (status {})", st[0]);
        let stream = &dst[before..];
        let (_, rest) = entropy_coders::NormHistogram::read(stream).expect("valid header");
        let last = *stream.last().unwrap();
        (rest.len() - 1) * 8 + (8 - last.leading_zeros() as usize) // payload bits including the marker
    }

    /// The reference's termination rule (no stored length): decode until the bit stack cannot supply num_bits.
    /// The capacity grows geometrically until the stream fits (a stream of p = 0.995 data expands ~175 x); only the
    /// stream that can never terminate (every state needs 0 bits: SURVEY quirk Q1) ends in a panic, at 1 GiB.
    fn decompress_one(&mut self, src: &[u8], dst: &mut Vec<u8>, n_states: u32) -> Option<usize> {
        assert!(!src.is_empty(), "No bytes provided to read from"); // stream_reader.rs:17 via lib.rs:191,219
        let comp = Dev::from(src);
        let off = Dev::from(&[0u64, src.len() as u64]);
        let mut cap = (64 * src.len()).max(4096);
        loop {
            let out: Dev<u8> = Dev::new(cap);
            let (len, st): (Dev<u32>, Dev<i32>) = (Dev::new(1), Dev::new(1));
            let p = fse_b200_params { block_size: cap as u32, table_log: 15, n_states, ..Default::default() };
            ck(unsafe { fse_b200_decompress_exhaust(self.ctx, comp.p, src.len(), off.p, 1, &p, out.p, len.p, st.p) }, "fse_b200_decompress_exhaust");
            match st.host()[0] {
                FSE_B200_ERR_TABLE_LOG | FSE_B200_ERR_TOO_MANY | FSE_B200_ERR_IO | FSE_B200_ERR_NO_MARKER => return None, // .ok()? / BitStackReader::new
                FSE_B200_ERR_LENGTH => panic!("called `Option::unwrap()` on a `None` value"), // lib.rs:197,224-225
                FSE_B200_ERR_CAPACITY if cap < (1 << 30) => cap *= 8,
                FSE_B200_ERR_CAPACITY => panic!("the decoder never terminates on this stream (every state needs 0 bits)"),
                s if s < 0 => panic!("fse_decompress: status {}", s),
                _ => {
                    let n = len.host()[0] as usize;
                    dst.extend_from_slice(&out.host()[..n]);
                    return Some(n);
                }
            }
        }
    }
}

impl Drop for Gpu {
    fn drop(&mut self) {
        unsafe { fse_b200_destroy(self.ctx) }
    }
}

/// src/lib.rs:112-143
pub fn fse_compress(src: &[u8], dst: &mut Vec<u8>) -> (entropy_coders::NormHistogram, usize) {
    let before = dst.len();
    let bits = with_gpu(|g| g.compress_one(src, dst, 1));
    let (hist, _) = entropy_coders::NormHistogram::read(&dst[before..]).expect("valid header");
    (hist, bits)
}
/// src/lib.rs:146-183
pub fn fse_compress2(src: &[u8], dst: &mut Vec<u8>) -> usize {
    with_gpu(|g| g.compress_one(src, dst, 2))
}
/// src/lib.rs:187-211
pub fn fse_decompress(src: &[u8], dst: &mut Vec<u8>) -> Option<usize> {
    with_gpu(|g| g.decompress_one(src, dst, 1))
}
/// src/lib.rs:215-248
pub fn fse_decompress2(src: &[u8], dst: &mut Vec<u8>) -> Option<usize> {
    with_gpu(|g| g.decompress_one(src, dst, 2))
}

/// src/histogram.rs:10-91, :264-284
#[derive(Clone, Debug)]
pub struct Histogram {
    table: [u32; 256],
    size: u32,
    table_len: usize,
}
impl Histogram {
    /// Histogram::new, :18-66
    pub fn new(data: &[u8]) -> Self {
        assert!(data.len() <= u32::MAX as usize, "Data vector is too long");
        let mut h = Self { table: [0; 256], size: data.len() as u32, table_len: 1 };
        if data.is_empty() {
            return h;
        }
        with_gpu(|g| {
            let d = Dev::from(data);
            let (counts, tlen): (Dev<u32>, Dev<u32>) = (Dev::new(256), Dev::new(1));
            ck(unsafe { fse_b200_histogram_blocks(g.ctx, d.p, data.len(), data.len() as u32, counts.p, tlen.p) }, "fse_b200_histogram_blocks");
            h.table.copy_from_slice(&counts.host());
            h.table_len = tlen.host()[0] as usize;
        });
        h
    }
    pub fn table(&self) -> &[u32; 256] { &self.table }
    pub fn table_len(&self) -> usize { self.table_len }
    pub fn size(&self) -> u32 { self.size }
    /// counts the ZERO entries, like the crate (:79-81)
    pub fn symbol_count(&self) -> usize { self.table.iter().filter(|&&x| x == 0).count() }
    /// :264-277 (the same integer expression; panics where the crate's ilog2 / subtraction does)
    pub fn optimal_log2(&self) -> u32 {
        let min_bits_src = self.size.ilog2() + 1;
        let min_bits_symbols = ((self.table_len - 1) as u32).ilog2() + 2;
        let max_bits = (self.size - 1).ilog2() - 2;
        TABLE_LOG_DEFAULT.min(max_bits).max(min_bits_src.min(min_bits_symbols)).clamp(TABLE_LOG_MIN, TABLE_LOG_MAX)
    }
    /// :95-155 (+ normalize_slow :157-261) on the GPU; the result is the crate's own NormHistogram type
    pub fn normalize(self, log2: u32) -> entropy_coders::NormHistogram {
        assert!(self.table_len > 1 && self.size > 0, "attempt to calculate ilog2 of zero"); // :98
        with_gpu(|g| {
            let mut c64 = [0u64; 256];
            for (d, s) in c64.iter_mut().zip(self.table.iter()) {
                *d = *s as u64;
            }
            let counts = Dev::from(&c64[..]);
            let (norm, st): (Dev<i32>, Dev<i32>) = (Dev::new(256), Dev::new(1));
            let (l2, tl): (Dev<u32>, Dev<u32>) = (Dev::new(1), Dev::new(1));
            ck(unsafe { fse_b200_normalize(g.ctx, counts.p, 1, log2.clamp(TABLE_LOG_MIN, TABLE_LOG_MAX), norm.p, l2.p, tl.p, st.p) }, "fse_b200_normalize");
            assert!(st.host()[0] >= 0, "Histogram::normalize: the reference panics on this input");
            let mut t = [0i32; 256];
            t.copy_from_slice(&norm.host());
            entropy_coders::NormHistogram::try_from(t).expect("normalised counts sum to a power of two")
        })
    }
    /// :281-284
    pub fn normalize_optimal(self) -> entropy_coders::NormHistogram {
        let log2 = self.optimal_log2();
        self.normalize(log2)
    }
}

/// NormHistogram::new (src/histogram.rs:299-303) and header I/O (:376-505) on the GPU, producing / consuming the
/// crate's own NormHistogram type.
pub mod norm_histogram {
    use super::*;
    pub fn new(data: &[u8]) -> entropy_coders::NormHistogram {
        Histogram::new(data).normalize_optimal()
    }
    /// NormHistogram::write, :376-431: appends the header, returns the bits written
    pub fn write(hist: &entropy_coders::NormHistogram, writer: &mut Vec<u8>) -> usize {
        with_gpu(|g| {
            let norm = Dev::from(&hist.table()[..]);
            let (l2, tl) = (Dev::from(&[hist.log2_sum()]), Dev::from(&[hist.table_len() as u32]));
            let (nbytes, nbits): (Dev<u32>, Dev<u32>) = (Dev::new(1), Dev::new(1));
            let out: Dev<u8> = Dev::new(512);
            ck(unsafe { fse_b200_ncount_write(g.ctx, norm.p, l2.p, tl.p, 1, out.p, 512, nbytes.p, nbits.p) }, "fse_b200_ncount_write");
            writer.extend_from_slice(&out.host()[..nbytes.host()[0] as usize]);
            nbits.host()[0] as usize
        })
    }
    /// NormHistogram::read, :436-505: (histogram, the bytes after the header)
    pub fn read(data: &[u8]) -> Result<(entropy_coders::NormHistogram, &[u8]), HistError> {
        assert!(!data.is_empty(), "No bytes provided to read from"); // stream_reader.rs:17
        with_gpu(|g| {
            let d = Dev::from(data);
            let len = Dev::from(&[data.len() as u32]);
            let (norm, st): (Dev<i32>, Dev<i32>) = (Dev::new(256), Dev::new(1));
            let (l2, tl, cons): (Dev<u32>, Dev<u32>, Dev<u32>) = (Dev::new(1), Dev::new(1), Dev::new(1));
            ck(unsafe { fse_b200_ncount_read(g.ctx, d.p, data.len(), len.p, 1, norm.p, l2.p, tl.p, cons.p, st.p) }, "fse_b200_ncount_read");
            match st.host()[0] {
                FSE_B200_ERR_TABLE_LOG => Err(HistError::TableLogTooLarge((data[0] & 0x0f) as u32 + TABLE_LOG_MIN)), // :439-441
                FSE_B200_ERR_TOO_MANY => Err(HistError::TooManySymbols),
                FSE_B200_ERR_IO => Err(HistError::Io(std::io::ErrorKind::UnexpectedEof.into())),
                s if s < 0 => panic!("NormHistogram::read: status {}", s),
                _ => {
                    let mut t = [0i32; 256];
                    t.copy_from_slice(&norm.host());
                    let h = entropy_coders::NormHistogram::try_from(t).expect("a parsed header sums to a power of two");
                    Ok((h, &data[cons.host()[0] as usize..]))
                }
            }
        })
    }
}

/// src/fse.rs: tables built on the GPU; Encoder / Decoder are the crate's own host objects.
pub mod fse {
    use super::*;
    pub use entropy_coders::fse::{Decoder, Encoder};

    #[derive(Clone, Copy, Debug, Default)]
    pub struct SymbolTransform { pub bits: u32, pub find_state: i32 } // :80-84
    #[derive(Clone, Copy, Debug, Default)]
    pub struct DecodeTransform { pub new_state: u16, pub symbol: u8, pub num_bits: u8 } // :260-265

    /// EncodeTable::new / update, :88-189
    pub struct EncodeTable {
        pub table_log: u32,
        pub table: Vec<u16>,
        pub symbol_tt: Vec<SymbolTransform>,
        pub symbols: Vec<u8>,
    }
    impl EncodeTable {
        pub fn new(hist: &entropy_coders::NormHistogram) -> Self {
            let table_log = hist.log2_sum();
            assert!((TABLE_LOG_MIN..=TABLE_LOG_MAX).contains(&table_log), "FSE Table must be between 2^9 to 2^16"); // :103-106
            let size = 1usize << table_log;
            with_gpu(|g| {
                let norm = Dev::from(&hist.table()[..]);
                let (l2, tl) = (Dev::from(&[table_log]), Dev::from(&[hist.table_len() as u32]));
                let (t, sym): (Dev<u16>, Dev<u8>) = (Dev::new(size), Dev::new(size));
                let tt: Dev<fse_b200_symbol_transform> = Dev::new(256);
                let st: Dev<i32> = Dev::new(1);
                ck(unsafe { fse_b200_build_encode_tables(g.ctx, norm.p, l2.p, tl.p, 1, table_log, t.p, tt.p, sym.p, st.p) }, "fse_b200_build_encode_tables");
                assert!(st.host()[0] >= 0, "EncodeTable::update");
                Self {
                    table_log,
                    table: t.host(),
                    symbol_tt: tt.host().iter().map(|x| SymbolTransform { bits: x.bits, find_state: x.find_state }).collect(),
                    symbols: sym.host(),
                }
            })
        }
        /// :191-193
        pub fn compress_bound(size: usize) -> usize { unsafe { fse_b200_compress_bound(size) } }
    }

    /// DecodeTable::new / update, :269-338
    pub struct DecodeTable {
        pub table_log: u32,
        pub table: Vec<DecodeTransform>,
    }
    impl DecodeTable {
        pub fn new(hist: &entropy_coders::NormHistogram) -> Self {
            let table_log = hist.log2_sum();
            assert!((TABLE_LOG_MIN..=TABLE_LOG_MAX).contains(&table_log), "FSE Table must be between 2^9 to 2^16");
            let size = 1usize << table_log;
            with_gpu(|g| {
                let norm = Dev::from(&hist.table()[..]);
                let (l2, tl) = (Dev::from(&[table_log]), Dev::from(&[hist.table_len() as u32]));
                let t: Dev<fse_b200_decode_transform> = Dev::new(size);
                let st: Dev<i32> = Dev::new(1);
                ck(unsafe { fse_b200_build_decode_tables(g.ctx, norm.p, l2.p, tl.p, 1, table_log, t.p, st.p) }, "fse_b200_build_decode_tables");
                assert!(st.host()[0] >= 0, "DecodeTable::update");
                Self {
                    table_log,
                    table: t.host().iter().map(|x| DecodeTransform { new_state: x.new_state, symbol: x.symbol, num_bits: x.num_bits }).collect(),
                }
            })
        }
    }
}
