//! Pins the CPU oracle against an execution of the reference crate.
//! NOT compiled in this repository's environment (no cargo / rustc in the image).
//!
//!     cargo run --release --example dump_reference_vectors > ../tests/golden/reference_dump.json
//!
//! For every known-answer input of tests/golden/kat.json (read from stdin or regenerated from its seed) the
//! program prints the bytes the reference's own `fse_compress` and `fse_compress2` append to an empty Vec.
//! tests/test_reference_vectors.py compares the oracle with that file when it is present.
use entropy_coders::{fse_compress, fse_compress2};

/// r16(seed, i) = (splitmix64(seed + (i >> 2)) >> (16 * (i & 3))) & 0xFFFF
fn splitmix64(x: u64) -> u64 {
    let mut z = x.wrapping_add(0x9E37_79B9_7F4A_7C15);
    z = (z ^ (z >> 30)).wrapping_mul(0xBF58_476D_1CE4_E5B9);
    z = (z ^ (z >> 27)).wrapping_mul(0x94D0_49BB_1331_11EB);
    z ^ (z >> 31)
}
fn r16(seed: u64, i: u64) -> u16 {
    ((splitmix64(seed.wrapping_add(i >> 2)) >> (16 * (i & 3))) & 0xFFFF) as u16
}
/// G_geo(0.2): the reference generator's LUT (lib.rs:255-270) driven by r16 instead of thread_rng
fn geo_lut() -> Vec<u8> {
    let mut lut = Vec::with_capacity(4096);
    let (mut remaining, mut sym) = (4096usize, 0u8);
    while remaining > 0 {
        let n = std::cmp::max(1, (remaining as f64 * 0.2) as usize).min(remaining);
        lut.extend(std::iter::repeat(sym).take(n));
        remaining -= n;
        sym += 1;
    }
    lut
}
fn geo(seed: u64, n: usize) -> Vec<u8> {
    let lut = geo_lut();
    (0..n as u64).map(|i| lut[(r16(seed, i) & 4095) as usize]).collect()
}
fn hex(b: &[u8]) -> String {
    b.iter().map(|x| format!("{:02x}", x)).collect()
}
fn dump(name: &str, src: &[u8]) {
    let mut a = Vec::new();
    let (_, bits1) = fse_compress(src, &mut a);
    let mut b = Vec::new();
    let bits2 = fse_compress2(src, &mut b);
    println!(
        "{{\"name\": \"{}\", \"src_len\": {}, \"fse_compress_hex\": \"{}\", \"fse_compress_bits\": {}, \
         \"fse_compress2_hex\": \"{}\", \"fse_compress2_bits\": {}}}",
        name, src.len(), hex(&a), bits1, hex(&b), bits2
    );
}
fn main() {
    // the same inputs as tests/golden/make_golden.py
    dump("geo-c1-4096", &geo(0xC0FFEE01, 4096));
    dump("geo-c1-65536", &geo(0xC0FFEE01, 65536));
    dump("geo-c4-131072", &geo(0xC0FFEE04, 131072));
    let abab: Vec<u8> = (0..64).map(|i| if i % 2 == 0 { b'a' } else { b'b' }).collect();
    dump("abab-64", &abab);
    let ramp: Vec<u8> = (0..4096usize).map(|i| (i % 251) as u8).collect();
    dump("ramp-251", &ramp);
}
