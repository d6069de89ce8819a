"""The crate's bitstream tests (src/bitstream/mod.rs:112-224: stack_tests{,_0,_1}, stream_tests{,_0,_1}) driven through the
GPU's own bit I/O code (fse_b200_bitstack_write / _read, fse_b200_bitstream_read = BitRowS + warp_place, the marker search
and prefix-sum reads, FwdBits) with the oracle's mechanics model of BitStackWriter / BitStackReader / BitStreamReader as
the checker."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import pymodel as M  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import entropy_coders_b200 as E
    c = E.Context(0)
    yield c
    c.close()


def dev(ctx, arr, dtype):
    import torch
    return torch.from_numpy(np.ascontiguousarray(arr, dtype=dtype)).to(ctx.device)


def model_write(fields, mark, offset=0):
    v = M.Vec(bytes(offset), base=0x1000 + 3)                # an unaligned Vec: the bytes must not depend on it (mod.rs:151-155)
    w = M.BitStackWriter(v)
    for val, nb in fields:
        w.write_bits(val, nb)
    if mark:
        w.write_bits(1, 1)
    bits = w.finish()
    return v.bytes()[offset:], bits


def fields_of(rng, n, one_bit):
    bits = np.ones(n, dtype=np.uint8) if one_bit else rng.integers(1, 17, size=n).astype(np.uint8)
    vals = (rng.integers(0, 1 << 16, size=n) & ((1 << bits.astype(np.int64)) - 1)).astype(np.uint32)
    return vals, bits


@pytest.mark.parametrize("one_bit", [True, False])
@pytest.mark.parametrize("n", [1, 2, 7, 8, 31, 32, 33, 255, 256, 257, 1000, 5000])
def test_stack_writer_and_reader(ctx, n, one_bit):
    """stack_tests: the writer reports the exact bit count and ceil((bits + 1) / 8) bytes (mod.rs:44-59); the reader
    returns the fields in reverse, ends with nothing available and finish() true (mod.rs:68-91)"""
    rng = np.random.default_rng(n * 2 + one_bit)
    vals, bits = fields_of(rng, n, one_bit)
    out, nbits = ctx.bitstack_write(dev(ctx, vals, np.uint32), dev(ctx, bits, np.uint8), mark=True)
    exp, ebits = model_write(list(zip(vals.tolist(), bits.tolist())), True)
    total = int(bits.sum())
    assert nbits == ebits == total + 1 and out.numel() == (total + 1 + 7) // 8
    assert out.cpu().numpy().tobytes() == exp
    got, st = ctx.bitstack_read(out, dev(ctx, bits, np.uint8))
    assert st == 0 and np.array_equal(got.cpu().numpy().view(np.uint32), vals)
    # the model's reader pops the same values (field n-1 first) from the GPU's bytes and ends finished
    r = M.BitStackReader(out.cpu().numpy().tobytes())
    for i in range(n - 1, -1, -1):
        assert r.read(int(bits[i])) == int(vals[i])
    assert r.available() == 0 and r.finish()
    # one field too many: read -> None; one too few: finish() false
    more = np.concatenate([np.array([3], dtype=np.uint8), bits])
    assert ctx.bitstack_read(out, dev(ctx, more, np.uint8))[1] == -7
    if n > 1:
        assert ctx.bitstack_read(out, dev(ctx, bits[1:], np.uint8))[1] == -7


@pytest.mark.parametrize("one_bit", [True, False])
@pytest.mark.parametrize("n", [1, 5, 32, 33, 300, 2000])
def test_stream_reader(ctx, n, one_bit):
    """stream_tests: forward reads of the unmarked stream; finish() leaves no bits (mod.rs:93-110); one more read is UnexpectedEof"""
    rng = np.random.default_rng(1000 + n * 2 + one_bit)
    vals, bits = fields_of(rng, n, one_bit)
    out, nbits = ctx.bitstack_write(dev(ctx, vals, np.uint32), dev(ctx, bits, np.uint8), mark=False)
    exp, ebits = model_write(list(zip(vals.tolist(), bits.tolist())), False)
    assert nbits == ebits == int(bits.sum()) and out.cpu().numpy().tobytes() == exp
    got, st = ctx.bitstream_read(out, nbits, dev(ctx, bits, np.uint8))
    assert st == 0 and np.array_equal(got.cpu().numpy().view(np.uint32), vals)
    m = M.BitStreamReader(exp, nbits)
    for i in range(n):
        assert m.read(int(bits[i])) == int(vals[i])
    more = np.concatenate([bits, np.array([1], dtype=np.uint8)])
    assert ctx.bitstream_read(out, nbits, dev(ctx, more, np.uint8))[1] == -5          # UnexpectedEof -> HistError::Io
    assert ctx.bitstream_read(out, nbits + 8, dev(ctx, bits, np.uint8))[1] == -8      # constructor assert (stream_reader.rs:18-21)


def test_stack_reader_none_cases(ctx):
    """BitStackReader::new -> None: the last byte must hold the marker (stack_reader.rs:18-20, :77-83)"""
    z = dev(ctx, np.array([0xFF, 0x00], dtype=np.uint8), np.uint8)
    b = dev(ctx, np.array([4], dtype=np.uint8), np.uint8)
    assert ctx.bitstack_read(z, b)[1] == -6
    assert ctx.bitstack_read(z[:0], b)[1] == -6
