// fse_bitio.cuh -- the bit I/O building blocks of the coders behind their own entry points, so that the crate's
// bitstream tests (src/bitstream/mod.rs:112-224: stack_tests / stream_tests) can drive them directly:
//   BitStackWriter (writer.rs:140-222)        -> BitRowS::put + warp_place (what the encode kernels pack with)
//   BitStackReader (stack_reader.rs:17-226)   -> marker search + funnel reads from the top of the stack at positions
//                                                 given by a warp prefix sum (what the decode kernels read with)
//   BitStreamReader (stream_reader.rs:16-135) -> FwdBits (what the header parse reads with)
// One warp per call; fields of 0..16 bits.
#pragma once
#include "fse_encode128.cuh"

namespace fsed {

constexpr int BITIO_K = 8;            // fields per lane per chunk: 8 x 16 bits + 31 carried bits < ROW_STRIDE64 words

// vals[i] (masked to bits[i]) are appended in index order, then an optional marker bit; out: word aligned
__global__ void __launch_bounds__(32) k_bitstack_write(const uint32_t *__restrict__ vals, const uint8_t *__restrict__ bits, uint32_t n,
                                                       int mark, uint32_t *out, uint32_t cap_words, unsigned long long *nbits_out, int *status)
{
    __shared__ uint32_t rows[32 * ROW_STRIDE64];
    const int lane = threadIdx.x;
    uint32_t *myrow = rows + lane * ROW_STRIDE64;
    uint32_t cw = 0, cb = 0, wdone = 0;
    bool ovf = false;
    const uint32_t total = n + (mark ? 1u : 0u);
    for (uint32_t c0 = 0; c0 < total; c0 += 32 * BITIO_K) {
        BitRowS br;
        br.init(myrow, lane == 0 ? cw : 0u, lane == 0 ? cb : 0u);
#pragma unroll
        for (int k = 0; k < BITIO_K; k++) {
            const uint32_t i = c0 + lane * BITIO_K + k;
            if (i < n) {
                const uint32_t nb = bits[i];
                br.put(vals[i] & ((1u << nb) - 1u), nb);             // write_bits_unmasked, writer.rs:195-198
            } else if (i == n && mark) br.put(1, 1);
        }
        const uint32_t tot = br.finish();
        __syncwarp();
        wdone += warp_place(myrow, tot, out + wdone, cap_words > wdone ? cap_words - wdone : 0, lane, cw, cb, ovf);
        __syncwarp();
    }
    if (cb) {
        if (wdone < cap_words) { if (lane == 0) out[wdone] = cw; }
        else ovf = true;
    }
    if (lane == 0) { *nbits_out = (unsigned long long)wdone * 32 + cb; *status = ovf ? ST_CAPACITY : ST_OK; }
}

// up to 16 bits at bit position q of a byte array (any alignment), little endian, LSB first
__device__ __forceinline__ uint32_t bits_at(const uint8_t *__restrict__ p, size_t nbytes, unsigned long long q, uint32_t nb)
{
    const size_t byte = (size_t)(q >> 3);
    uint32_t v = 0;
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (byte + k < nbytes) v |= (uint32_t)p[byte + k] << (8 * k);
    return (v >> (uint32_t)(q & 7)) & ((1u << nb) - 1u);
}

// The fields written by k_bitstack_write with a marker are read back from the END: field n-1 first (the stack order of
// BitStackReader::read, stack_reader.rs:211-215), each lane at the offset a warp prefix sum of the widths gives it.
// status: ST_NO_MARKER (empty input or last byte zero, :18-20, :77-83), ST_LENGTH when the stack cannot supply a field
// (read -> None) or bits are left over (finish() false, :224-226).
__global__ void __launch_bounds__(32) k_bitstack_read(const uint8_t *__restrict__ in, size_t nbytes, const uint8_t *__restrict__ bits, uint32_t n,
                                                      uint32_t *vals, int *status)
{
    const int lane = threadIdx.x;
    if (nbytes == 0 || in[nbytes - 1] == 0) { if (lane == 0) *status = ST_NO_MARKER; return; }
    unsigned long long cur = (unsigned long long)(nbytes - 1) * 8 + ilog2u(in[nbytes - 1]);     // the marker's position = bits below it
    bool bad = false;
    for (uint32_t c0 = 0; c0 < n; c0 += 32) {                 // 32 fields per step, taken from the top: field n-1-c0-lane
        const uint32_t j = c0 + lane;
        const bool on = j < n;
        const uint32_t i = on ? n - 1 - j : 0;
        const uint32_t nb = on ? bits[i] : 0u;
        const uint32_t incl = warp_incl_add(nb, lane);
        const uint32_t tot = __shfl_sync(FULL, incl, 31);
        if (tot > cur) { bad = true; break; }
        if (on) vals[i] = bits_at(in, nbytes, cur - incl, nb);
        cur -= tot;
    }
    if (lane == 0) *status = (bad || cur != 0) ? ST_LENGTH : ST_OK;
}

// BitStreamReader: fields read forward under a total_bits bound; status ST_IO on UnexpectedEof (stream_reader.rs:70-72,
// :85-87), else the number of bits left (finish(), :123-128)
__global__ void __launch_bounds__(32) k_bitstream_read(const uint8_t *__restrict__ in, size_t nbytes, unsigned long long total_bits,
                                                       const uint8_t *__restrict__ bits, uint32_t n, uint32_t *vals, int *status)
{
    if (threadIdx.x != 0) return;
    if (nbytes == 0 || (total_bits + 7) / 8 != nbytes) { *status = ST_PANIC; return; }          // :17-21
    FwdBits r{in, (uint32_t)nbytes, 0};
    for (uint32_t i = 0; i < n; i++) {
        uint32_t v = 0;
        const uint32_t nb = bits[i];
        if ((unsigned long long)r.pos + nb > total_bits || !r.peek(nb, v)) { *status = ST_IO; return; }
        vals[i] = v;
        r.pos += nb;
    }
    *status = (int)(total_bits - r.pos);
}

}  // namespace fsed
