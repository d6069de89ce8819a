"""Property tests (hypothesis) of the CPU oracle, SURVEY.md 8(c)(iii): the structural pins the reference's own tests
hold (histogram.rs:566-586, lib.rs:280-302, fse.rs:479-506, bitstream/mod.rs:44-110) over generated inputs instead of
the crate's unseeded RNG, and agreement of the two independently written models.  The GPU differential twin of this
file is tests/test_gpu_properties.py."""
import os
import sys

import numpy as np
import pytest
from hypothesis import HealthCheck, assume, given, settings, strategies as st

import oracle_lib as O

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import pymodel as M  # noqa: E402

SCALE = int(os.environ.get("FSE_HYP_SCALE", "1"))       # FSE_HYP_SCALE=20 for a long soak
COMMON = dict(deadline=None, derandomize=SCALE == 1, suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large])


@st.composite
def byte_strings(draw, min_size=0, max_size=6000):
    """alphabets from 1 to 256 symbols, flat to very skewed, with runs"""
    n = draw(st.integers(min_size, max_size))
    nsym = draw(st.sampled_from([1, 2, 3, 5, 17, 60, 130, 256]))
    skew = draw(st.sampled_from([0.0, 0.5, 1.0, 2.0, 4.0]))
    seed = draw(st.integers(0, 2 ** 32 - 1))
    rng = np.random.default_rng(seed)
    alphabet = rng.permutation(256)[:nsym]
    w = 1.0 / (np.arange(nsym) + 1.0) ** skew
    out = alphabet[rng.choice(nsym, size=n, p=w / w.sum())].astype(np.uint8)
    if n > 8 and draw(st.booleans()):                        # a run, as real data has them
        a = int(rng.integers(0, n - 4))
        out[a:a + int(rng.integers(1, n - a))] = out[a]
    return out


@settings(max_examples=150 * SCALE, **COMMON)
@given(byte_strings(min_size=1), st.sampled_from([0, 5, 7, 9, 11, 12, 13, 15]))
def test_normalize_sums_to_table_size_and_keeps_zeros(data, tl):
    """histogram.rs:566-577: sum |norm| == 1 << log2; norm[i] == 0 exactly where count[i] == 0"""
    h = O.histogram(data)
    assume(h.table_len > 1)                                   # histogram.rs:98 panics on a single-symbol alphabet
    if tl == 0:
        rc, tl = O.optimal_log2(h)
        assume(rc == 0)
    rc, n = O.normalize(h, tl)
    assume(rc == 0)
    norm = np.array(n.table[:256])
    counts = np.array(h.table[:256])
    assert np.abs(norm).sum() == 1 << n.log2
    assert ((norm == 0) == (counts == 0)).all()
    assert n.log2 >= tl and n.table_len == h.table_len


@settings(max_examples=150 * SCALE, **COMMON)
@given(byte_strings(min_size=2), st.sampled_from([0, 6, 9, 11, 14]), st.binary(max_size=9))
def test_header_write_read_identity_with_trailing_bytes(data, tl, trailer):
    """histogram.rs:580-586: read(write(n) ++ tail) == (n, tail)"""
    h = O.histogram(data)
    assume(h.table_len > 1)
    if tl == 0:
        rc, tl = O.optimal_log2(h)
        assume(rc == 0)
    rc, n = O.normalize(h, tl)
    assume(rc == 0)
    hdr, bits = O.ncount_write(n)
    assert len(hdr) == (bits + 7) // 8
    rc, back, consumed = O.ncount_read(hdr + trailer)
    assert rc == 0 and consumed == len(hdr)
    assert back.log2 == n.log2 and back.table_len == n.table_len
    assert list(back.table[:256]) == list(n.table[:256])


@settings(max_examples=120 * SCALE, **COMMON)
@given(byte_strings(min_size=1, max_size=5000), st.sampled_from([1, 2, 4, 8, 32, 64, 128]), st.sampled_from([0, 0, 9, 11, 12]))
def test_round_trip_any_state_count(data, n_states, tl):
    """lib.rs:280-302 / fse.rs:479-506 generalised to N states: decode(encode(x)) == x, the stream is consumed exactly"""
    try:
        comp, hb, pb = O.compress_n(data, tl, n_states)
    except ValueError:
        return                                                # inputs the reference panics on (escape blocks on the GPU)
    assert len(comp) == hb + (pb + 7) // 8                    # payload bits (marker included, lib.rs:141-142), padded
    assert O.decompress_n_len(comp, n_states, data.size) == data.tobytes()


@settings(max_examples=60 * SCALE, **COMMON)
@given(byte_strings(min_size=2, max_size=700), st.integers(0, 7))
def test_c_oracle_equals_mechanics_model(data, align):
    """the C restatement and the Python model of the reference's mechanics (accumulator, Vec growth, pointer
    alignment) emit the same fse_compress / fse_compress2 bytes and decode them alike"""
    for n_states, fc in ((1, M.fse_compress), (2, M.fse_compress2)):
        try:
            comp, _, _ = O.compress_n(data, 0, n_states)
        except ValueError:
            continue
        vec = M.Vec(base=0x1000 + align)
        fc(bytes(data.tobytes()), vec)
        assert vec.bytes() == comp, n_states


@settings(max_examples=60 * SCALE, **COMMON)
@given(byte_strings(min_size=2, max_size=3000))
def test_reference_loop_structure_equals_generic_two_state_codec(data):
    """oracle: the literal lib.rs:146-183 / :215-248 loops (CPU baseline) == the N-state codec at N = 2"""
    try:
        comp, _, _ = O.compress_n(data, 0, 2)
    except ValueError:
        return
    assert O.ref_compress2(data) == comp


@pytest.mark.parametrize("n_states", [1, 2, 32, 128])
def test_single_bit_flips_never_crash_the_decoder(n_states):
    """every outcome of a corrupted stream is a status or some bytes, never an out-of-bounds access (the GPU kernels
    follow the same guards; tests/test_gpu_parity.py::test_decode_oracle_streams_and_errors)"""
    data = O.generate("text", 3, 3000)
    comp, _, _ = O.compress_n(data, 0, n_states)
    rng = np.random.default_rng(n_states)
    for _ in range(200):
        bad = bytearray(comp)
        bad[int(rng.integers(0, len(bad)))] ^= 1 << int(rng.integers(0, 8))
        try:
            O.decompress_n_len(bytes(bad), n_states, data.size)
        except ValueError:
            pass
