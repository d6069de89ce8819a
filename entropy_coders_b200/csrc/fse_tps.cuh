// fse_tps.cuh -- the reference's OWN stream formats (one or two states per stream: fse_compress / fse_compress2,
// src/lib.rs:112-248): one THREAD per stream.
//
// A stream with one or two states is one or two serial chains (fse.rs:227-239, :363-373): a warp per block has one or two
// lanes at work (c4, 65 536 blocks: 26 GB/s encode, 14 GB/s decode at two states).  The parallelism is across streams, so
// here every lane runs its own block exactly as the CPU does -- states, the bit window of BitStackWriter (writer.rs:140-149)
// / BitStackReader (stack_reader.rs:97-172) in registers -- against that block's tables in SHARED memory.  The tables are
// built by one warp per block with the code of the warp-per-block kernels (k_tps_prepare_*), which also settles every block
// that needs no coder (escapes, errors), and pass through global memory to the CTA that codes the block.  Same bytes as
// k_encode_blocks / k_decode_blocks (tests/test_gpu_parity.py::test_many_streams_* and every one- / two-state test).
// History: the first form read the tables from global memory (c4 two states 75 / 45 GB/s: every look-up an
// L2 / DRAM round trip); in shared memory 148 / 125 GB/s.
#pragma once
#include "fse_kernels.cuh"
#include "fse_kernels64.cuh"
#include "fse_decode128c.cuh"

namespace fsed {

// With the tables in shared memory one thread per stream beats one warp per stream at every block count measured (256 blocks
// of 128 KiB, two states: 3.7 / 3.4 ms against 8.6 / 16.2 ms; a lone stream takes the same time as 256): the one or two
// lanes a warp-per-block kernel keeps busy run the same chain with more instructions around it.
constexpr uint32_t TPS_MIN_BLOCKS = 1;
// the host-buffer paths hand over chunks of at least this many blocks: a round of CTAs holds 148 x 28..37 streams
constexpr uint32_t TPS_PIPE_BLOCKS = 8192;
// The encoder takes the blocks in waves of this many (two states; four times as many with one): the table scratch in global
// memory (6 KiB per block) stays in the L2 between the kernel that builds a wave's tables and the one that copies them in.
constexpr uint32_t TPS_ENC_WAVE = 16384;

// per block: x = table_log, y = header bytes (encode) / header bytes consumed (decode), z = 1 when the coder has work
struct TpsTables {
    uint16_t *enc_tab;      // [nblocks << tlmax]
    uint2 *enc_tt;          // [nblocks * 256]
    uint32_t *dec_tab;      // [nblocks << tlmax]
    uint4 *meta;            // [nblocks]
    uint32_t first, count;  // this launch covers blocks [first, first + count); tables and meta are indexed by b - first
};

// ---------------------------------------------------------------------------------- encode
// one warp per block: k_encode_blocks up to the tables (same escapes, same header bytes), tables to global memory
__global__ void __launch_bounds__(512) k_tps_prepare_enc(EncArgs a, TpsTables g)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    const EncLayout lay = enc_layout(a.tlmax);
    uint8_t *my = smem_raw + (size_t)warp * lay.total;
    uint16_t *tab = reinterpret_cast<uint16_t *>(my + lay.tab);
    uint2 *tt = reinterpret_cast<uint2 *>(my + lay.tt);
    uint32_t *cnt = reinterpret_cast<uint32_t *>(my + lay.work);
    int32_t *norm = reinterpret_cast<int32_t *>(my + lay.work + 1024);
    uint32_t *cum = reinterpret_cast<uint32_t *>(my + lay.work + 2048);
    uint8_t *spread = my + lay.work + 3072;
    uint32_t *rows = reinterpret_cast<uint32_t *>(my + lay.rows);
    const uint32_t N = a.n_states;
    for (uint32_t r = blockIdx.x * wpc + warp; r < g.count; r += gridDim.x * wpc) {
        const uint32_t b = g.first + r;
        const size_t off = (size_t)b * a.block_size;
        const uint32_t bn = (uint32_t)min((size_t)a.block_size, a.n - off);
        const uint8_t *bsrc = a.src + off;
        uint8_t *bs = a.scratch + (size_t)b * a.stride;
        uint32_t *hdr_words = reinterpret_cast<uint32_t *>(bs);
        uint32_t log2 = 0;
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 8; k++) cnt[k * 32 + lane] = a.counts[(size_t)b * 256 + k * 32 + lane];
        __syncwarp();
        uint32_t table_len;
        int rc = warp_normalize(cnt, (uint64_t)bn, a.req_log2, norm, lane, log2, table_len);
        uint32_t hl = 0, ready = 0;
        int st = ST_OK;
        if (rc < 0) {                                    // blocks the reference panics on: escapes (include/fse_b200.h)
            if (table_len <= 1) { if (lane == 0) { bs[0] = 0x0E; bs[1] = 0x00; } hl = 2; st = 2; }
            else if (bn <= 4) { if ((uint32_t)lane < bn) bs[1 + lane] = bsrc[lane]; if (lane == 0) bs[0] = 0x0F; hl = 1 + bn; st = 1; }
            else st = rc;
        } else if (bn < N) {                             // fewer symbols than states: lib.rs:121,154,156
            if ((uint32_t)lane < bn) bs[1 + lane] = bsrc[lane];
            if (lane == 0) bs[0] = 0x0F;
            hl = 1 + bn; st = 1;
        } else if (log2 > a.tlmax) {
            st = ST_UNSUPPORTED;
        } else {
            const uint32_t hbits = warp_ncount_write(norm, log2, table_len, rows, hdr_words, lane);
            hl = (hbits + 7) >> 3;
            warp_spread(norm, log2, table_len, spread, cum, tab, lane);
            warp_build_encode(norm, log2, table_len, spread, cum, tab, tt, lane);
            __syncwarp();
            uint16_t *gt = g.enc_tab + ((size_t)r << a.tlmax);
            uint2 *gs = g.enc_tt + (size_t)r * 256;
            for (uint32_t i = lane; i < (1u << log2) / 2; i += 32)
                reinterpret_cast<uint32_t *>(gt)[i] = reinterpret_cast<const uint32_t *>(tab)[i];
#pragma unroll
            for (int k = 0; k < 8; k++) gs[k * 32 + lane] = tt[k * 32 + lane];
            ready = 1;
        }
        if (lane == 0) {
            g.meta[r] = make_uint4(log2, hl, ready, 0u);
            a.hlen[b] = hl;
            a.plen[b] = 0;
            a.status[b] = st;
        }
    }
}

// predicated global load / L1 prefetch (no branch: the lanes of a warp run different streams, and a divergent branch with
// its reconvergence costs more than the handful of instructions it would skip)
__device__ __forceinline__ void tps_ld_if(uint32_t &v, const uint32_t *p, bool c)
{
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %2, 0; @q ld.global.nc.u32 %0, [%1]; }" : "+r"(v) : "l"(p), "r"((uint32_t)c));
}
__device__ __forceinline__ void tps_prefetch_if(const void *p, bool c)
{
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %1, 0; @q prefetch.global.L1 [%0]; }" ::"l"(p), "r"((uint32_t)c));
}

// ---- the shared-memory form of the same encoder (k_tps_encode_smem).  The state chain is s -> nb -> s >> nb -> look-up
// (fse.rs:227-239); everything else is arranged to stay off it and to cost few instructions:
//  * the symbol transforms of the shared copy hold the ADDRESS of their first next-state cell (tab_s + 2 * find_state), so
//    the look-up address is one LEA;
//  * BitStackWriter (writer.rs:140-149) as a 64-bit window {hi, lo} filled from the TOP: emitting the low nb bits of the
//    state is two funnel shifts (lo = {hi, lo} >> nb, hi = {s, hi} >> nb), nothing is masked or OR-ed; the valid bits are
//    the top `cnt` bits, a full word is cut out with one more funnel shift.  Two symbols emit <= 24 bits (table_log <= 12),
//    so the window is flushed once per pair;
//  * the source is read as aligned 32-bit words (four symbols), one word ahead, from sectors prefetched into the L1.
// ~13 instructions per symbol instead of 46.
__device__ __forceinline__ void tps_encode_stream_smem(const EncArgs &a, uint32_t b, uint4 m, uint32_t tt_s)
{
    const uint32_t log2 = m.x, N = a.n_states;
    const size_t off = (size_t)b * a.block_size;
    const uint32_t bn = (uint32_t)min((size_t)a.block_size, a.n - off);
    const uint8_t *__restrict__ src = a.src + off;
    uint32_t *pay = reinterpret_cast<uint32_t *>(a.scratch + (size_t)b * a.stride + HDR_RESERVE);
    const uint32_t cap = a.pay_cap_words;
    uint32_t lo = 0, hi = 0, cnt = 0, wp = 0;
    auto push = [&](uint32_t v, uint32_t n) {                 // the low n & 31 bits of v on top of the window
        lo = __funnelshift_r(lo, hi, n);
        hi = __funnelshift_r(hi, v, n);
    };
    auto flush = [&]() {                                      // cnt < 64; a full word leaves the window (no branch: see tps_ld_if)
        const uint32_t w = __funnelshift_rc(lo, hi, 64 - cnt);
        asm volatile("{ .reg .pred q; setp.ne.u32 q, %2, 0; @q st.global.cs.u32 [%0], %1; }" ::"l"(pay + wp), "r"(w),
                     "r"((uint32_t)(cnt >= 32 && wp < cap)) : "memory");
        wp += cnt >> 5;
        cnt &= 31u;
    };
    auto ld_t = [&](uint32_t sym) -> uint2 { return lds_v2(tt_s + sym * 8); };
    auto first = [&](uint32_t sym) -> uint32_t {              // Encoder::new_first_symbol, fse.rs:210-218
        const uint2 t = ld_t(sym);
        const uint32_t bo = (t.x + (1u << 15)) >> 16;
        const uint32_t value = (bo << 16) - t.x;
        return lds_u16(t.y + ((value >> bo) << 1));
    };
    auto enc = [&](uint32_t &st, uint2 t) -> uint32_t {       // fse.rs:227-239; returns the bits emitted
        const uint32_t nb = (t.x + st) >> 16;
        push(st, nb);
        st = lds_u16(t.y + ((st >> nb) << 1));
        return nb;
    };
    auto word_at = [&](uint32_t &w, int32_t i, bool c) {      // if c: symbols i - 3 .. i (src + i - 3 is aligned); no branch
        const uint8_t *q = src + i - 3;
        tps_ld_if(w, reinterpret_cast<const uint32_t *>(q), c);
        tps_prefetch_if(q - 96, c && (i & 31) < 4 && i >= 99);
    };
    // inside the loop (i >= 3): the word below the current one, or the current one again when there is none (then it is not
    // used): the address is arithmetic, the load unconditional (a predicate set by ISETP is usable 13 cycles later)
    auto next_word = [&](uint32_t &w, int32_t i) {
        const uint32_t down = min(4u, (uint32_t)i - 3u) & 4u;                   // 4 when i >= 7
        const uint8_t *q = src + i - 3 - down;
        w = __ldg(reinterpret_cast<const uint32_t *>(q));
        tps_prefetch_if(q - 96, (i & 31) >= 4 && (i & 31) < 8 && i >= 103);
    };
    auto sym = [&](uint32_t c, uint32_t sel) -> uint32_t { return __byte_perm(c, 0u, sel); };    // one byte of the word
    int32_t i = (int32_t)bn - 1;
    if (N == 2) {
        // sA is the state of the parity of i (the symbol coded next), sB the other one
        uint32_t sA = first(__ldg(src + i)), sB = first(__ldg(src + i - 1));
        i -= 2;
        while (i >= 0 && (((uintptr_t)(src + i + 1)) & 3)) {   // down to a word boundary
            cnt += enc(sA, ld_t(__ldg(src + i)));
            flush();
            const uint32_t x = sA; sA = sB; sB = x;
            i--;
        }
        uint32_t w = 0u;
        word_at(w, i, i >= 3);
        for (; i >= 3; i -= 4) {
            const uint32_t c = w;
            next_word(w, i);
            const uint2 t3 = ld_t(sym(c, 0x4443)), t2 = ld_t(sym(c, 0x4442)), t1 = ld_t(sym(c, 0x4441)), t0 = ld_t(sym(c, 0x4440));
            const uint32_t n3 = enc(sA, t3);
            const uint32_t n2 = enc(sB, t2);
            cnt += n3 + n2;
            flush();
            const uint32_t n1 = enc(sA, t1);
            const uint32_t n0 = enc(sB, t0);
            cnt += n1 + n0;
            flush();
        }
        for (; i >= 0; i--) {
            cnt += enc(sA, ld_t(__ldg(src + i)));
            flush();
            const uint32_t x = sA; sA = sB; sB = x;
        }
        // i = -1: sA is the state of the odd indices.  Final states 1, 0 (fse.rs:241-250), then the marker (lib.rs:181)
        push(sA, log2); cnt += log2; flush();
        push(sB, log2); cnt += log2; flush();
    } else {
        uint32_t st = first(__ldg(src + i));
        i--;
        while (i >= 0 && (((uintptr_t)(src + i + 1)) & 3)) {
            cnt += enc(st, ld_t(__ldg(src + i)));
            flush();
            i--;
        }
        uint32_t w = 0u;
        word_at(w, i, i >= 3);
        for (; i >= 3; i -= 4) {
            const uint32_t c = w;
            next_word(w, i);
            const uint2 t3 = ld_t(sym(c, 0x4443)), t2 = ld_t(sym(c, 0x4442)), t1 = ld_t(sym(c, 0x4441)), t0 = ld_t(sym(c, 0x4440));
            const uint32_t n3 = enc(st, t3);
            const uint32_t n2 = enc(st, t2);
            cnt += n3 + n2;
            flush();
            const uint32_t n1 = enc(st, t1);
            const uint32_t n0 = enc(st, t0);
            cnt += n1 + n0;
            flush();
        }
        for (; i >= 0; i--) {
            cnt += enc(st, ld_t(__ldg(src + i)));
            flush();
        }
        push(st, log2); cnt += log2; flush();
    }
    push(1u, 1u); cnt += 1; flush();
    const uint32_t bits = wp * 32 + cnt;
    if (cnt) { if (wp < cap) pay[wp] = hi >> (32 - cnt); wp++; }
    if (wp > cap) { a.hlen[b] = 0; a.plen[b] = 0; a.status[b] = ST_CAPACITY; return; }
    a.plen[b] = (bits + 7) >> 3;
}

// The same streams with their tables in SHARED memory: a CTA takes `per_cta` blocks (as many as table sets fit: 37 at
// table_log 11: 4 KiB of next states + 2 KiB of symbol transforms each), copies their tables in, and lanes
// 0 .. lanes_per_warp - 1 of its warps run one stream each.
__global__ void __launch_bounds__(1024) k_tps_encode_smem(EncArgs a, TpsTables g, uint32_t per_cta, uint32_t lanes_per_warp)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t first = blockIdx.x * per_cta, tab_bytes = 2u << a.tlmax, set_bytes = tab_bytes + 2048u;
    const uint32_t sm0 = (uint32_t)__cvta_generic_to_shared(smem_raw);
    for (uint32_t j = 0; j < per_cta; j++) {
        const uint32_t r = first + j;
        if (r >= g.count) break;
        const uint4 *st = reinterpret_cast<const uint4 *>(g.enc_tab + ((size_t)r << a.tlmax));
        const uint4 *ss = reinterpret_cast<const uint4 *>(g.enc_tt + (size_t)r * 256);
        uint4 *dt = reinterpret_cast<uint4 *>(smem_raw + (size_t)j * set_bytes);
        const uint32_t tab_s = sm0 + j * set_bytes;
        for (uint32_t i = threadIdx.x; i < tab_bytes / 16; i += blockDim.x) dt[i] = st[i];
        for (uint32_t i = threadIdx.x; i < 128; i += blockDim.x) {       // two transforms: find_state -> address of its cell
            const uint4 v = ss[i];
            dt[tab_bytes / 16 + i] = make_uint4(v.x, tab_s + 2u * v.y, v.z, tab_s + 2u * v.w);
        }
    }
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane >= lanes_per_warp) return;
    const uint32_t j = warp * lanes_per_warp + lane, r = first + j;
    if (j >= per_cta || r >= g.count) return;
    const uint4 m = g.meta[r];
    if (!m.z) return;
    tps_encode_stream_smem(a, g.first + r, m, sm0 + j * set_bytes + tab_bytes);
}

// ---------------------------------------------------------------------------------- decode
__global__ void __launch_bounds__(512) k_tps_prepare_dec(DecArgs a, TpsTables g)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    const DecLayout lay = dec_layout(a.tlmax);
    uint8_t *my = smem_raw + (size_t)warp * lay.total;
    uint32_t *tab = reinterpret_cast<uint32_t *>(my + lay.tab);
    int32_t *norm = reinterpret_cast<int32_t *>(my + lay.norm);
    uint32_t *ctr = reinterpret_cast<uint32_t *>(my + lay.ctr);
    uint8_t *spread = my + lay.spread;
    const uint32_t N = a.n_states;
    for (uint32_t b = blockIdx.x * wpc + warp; b < a.nblocks; b += gridDim.x * wpc) {
        const size_t off = (size_t)b * a.block_size;
        const uint32_t bn = (uint32_t)min((size_t)a.block_size, a.n - off);
        uint8_t *out = a.dst + off;
        int st = ST_OK;
        const uint8_t *cs;
        uint32_t clen, log2 = 0, consumed = 0, ready = 0;
        __syncwarp();
        if (!dec_block_prologue(a, b, bn, N, out, lane, cs, clen, st)) {      // else: bad offsets, escape blocks: settled
#pragma unroll
            for (int k = 0; k < 8; k++) norm[k * 32 + lane] = 0;
            __syncwarp();
            uint32_t table_len = 0;
            int rc = 0;
            if (lane == 0) rc = ncount_read_serial(cs, clen, norm, log2, table_len, consumed);
            rc = __shfl_sync(FULL, rc, 0);
            log2 = __shfl_sync(FULL, log2, 0);
            table_len = __shfl_sync(FULL, table_len, 0);
            consumed = __shfl_sync(FULL, consumed, 0);
            __syncwarp();
            if (rc < 0) st = rc;
            else if (log2 > a.tlmax) st = ST_UNSUPPORTED;
            else if (bn < N) st = ST_LENGTH;
            else {
                warp_spread(norm, log2, table_len, spread, ctr, reinterpret_cast<uint16_t *>(tab), lane);
                warp_build_decode(norm, log2, table_len, spread, ctr, tab, lane);
                __syncwarp();
                uint32_t *gt = g.dec_tab + ((size_t)b << a.tlmax);
                for (uint32_t i = lane; i < (1u << log2); i += 32) gt[i] = tab[i];
                ready = 1;
            }
        }
        if (lane == 0) {
            g.meta[b] = make_uint4(log2, consumed, ready, 0u);
            a.status[b] = st;
        }
    }
}

// ---- the shared-memory form of the same decoder.  A stream is a serial chain (entry -> bits -> next entry), so what
// counts is the length of that chain and the instructions around it:
//  * entry of the shared copy: num_bits | symbol << 8 | (4 * new_state base) << 16 (tps_smem_entry): the shift amounts of a
//    read come straight from the entry (funnel shifts use the low five bits of their shift operand), the look-up address
//    is one LEA;
//  * the stack is read through a LEFT-aligned 64-bit window {wh, wl} (the next bits of the stack are wh's top bits): a read
//    of nb bits is wh >> (32 - nb) and the window moves up by nb, three funnel shifts, nothing masked;
//  * up to table_log 12 two reads never need more than 24 bits: the window is refilled (one 32-bit word, loaded one
//    refill ahead from sectors prefetched into the L1) once per pair when 32 bits or fewer are left;
//  * no test of the stack's depth on the chain: the bits used are counted and compared with the stream's once, at the end
//    (a stream that runs dry reads zeros -- every look-up stays inside its table -- and is reported as ST_LENGTH as before).
// Chain per symbol: LDS, SHF, LEA (and SHF + IADD beside the SHF); ~15 instructions per symbol instead of 48.
__device__ __forceinline__ uint32_t tps_smem_entry(uint32_t e) { return (e >> 24) | ((e >> 8) & 0xff00u) | ((e & 0xffffu) << 18); }

// COMPACT (table_log <= 11): 16-bit entries `num_bits | new_state base << 5` and the symbols as bytes beside them (3 bytes per
// cell instead of 4: 37 streams per SM instead of 28 at table_log 11); the state is the entry's shared ADDRESS, the symbol
// of a state is loaded with its entry from sym_c + address / 2.
template <bool COMPACT>
__device__ __forceinline__ void tps_decode_stream_smem(const DecArgs &a, uint32_t b, uint4 m, uint32_t tab_s, uint32_t sym_c)
{
    const uint32_t log2 = m.x, consumed = m.y, N = a.n_states;
    const size_t off = (size_t)b * a.block_size;
    const uint32_t bn = (uint32_t)min((size_t)a.block_size, a.n - off);
    uint8_t *out = a.dst + off;
    const unsigned long long o0 = a.offsets[b], o1 = a.offsets[b + 1];
    const uint8_t *pay = a.comp + o0 + consumed;
    const uint32_t plen = (uint32_t)(o1 - o0) - consumed;
    if (plen == 0 || pay[plen - 1] == 0) { a.status[b] = ST_NO_MARKER; return; }       // stack_reader.rs:17-92
    const uint32_t bias = (uint32_t)((uintptr_t)pay & 3);
    const uint32_t *__restrict__ origin = reinterpret_cast<const uint32_t *>(pay - bias);
    const uint32_t cur = (plen - 1) * 8 + ilog2u(pay[plen - 1]) + 8 * bias;             // marker position
    const uint32_t floor_bits = 8 * bias;
    if (cur - floor_bits < N * log2) { a.status[b] = ST_LENGTH; return; }               // lib.rs:197,224-225
    const uint32_t words = (cur + 31) >> 5;                   // origin[0 .. words) hold the stack
    const uint32_t top = words ? __ldg(origin + words - 1) : 0u;
    int32_t k = (int32_t)words - 2;                           // nx = origin[k], zeros below the stack
    uint32_t nx = k >= 0 ? __ldg(origin + k) : 0u;
    const uint32_t r = cur & 31;
    uint32_t wh = r ? top << (32 - r) : top, wl = 0u;
    uint32_t cnt = r ? r : 32u;                               // bits in the window (what lies below the stack counts too)
    const uint32_t entered0 = cnt + 32u * (words - 2u);       // bits used = entered0 - 32 k - cnt: every refill takes k down by one
    auto refill = [&]() {                                     // when 32 bits or fewer are left: wl is empty, nx goes in
        const bool need = cnt <= 32;
        wh |= __funnelshift_rc(nx, 0u, cnt);                  // nx >> cnt: nothing when cnt > 32 (the shift is clamped)
        const uint32_t t = __funnelshift_rc(0u, nx, cnt);
        wl = need ? t : wl;
        cnt = need ? cnt + 32 : cnt;
        k = need ? k - 1 : k;
        // below the stack the window is fed whatever origin[0] holds (a valid stream never uses those bits, a damaged one is
        // caught by the count of bits used); (a second word in flight was measured: no gain)
        const uint32_t *q = origin + max(k, 0);
        // wide form: loaded unconditionally (the same word again when nothing was taken: no predicate to wait for; 8 192 blocks
        // 9.6 -> 9.0 ms); compact form: predicated (c4: 68.7 against 77.0 ms unconditional, more lanes per warp at the LSU)
        if (COMPACT) tps_ld_if(nx, q, need);
        else nx = __ldg(q);
        // the third sector below: once per sector in the compact form, with every refill in the wide one (measured both ways
        // in both: 68.7 / 71.7 ms compact on c4, 10.5 / 9.6 ms wide on 8 192 blocks)
        tps_prefetch_if(q - 24, need && (!COMPACT || (k & 7) == 7) && k >= 24);
    };
    refill();
    auto take = [&](uint32_t e) -> uint32_t {                 // e & 31 bits off the top of the window
        const uint32_t bits = __funnelshift_l(wh, 0u, e);
        wh = __funnelshift_l(wl, wh, e);
        wl = __funnelshift_l(0u, wl, e);
        return bits;
    };
    auto look = [&](uint32_t &e, uint32_t &y, uint32_t state) {               // entry (and symbol) of a state
        if (COMPACT) { const uint32_t ad = tab_s + state * 2; e = lds_u16(ad); y = lds_u8((ad >> 1) + sym_c); }
        else { e = lds_u32(tab_s + state * 4); y = e; }
    };
    auto step = [&](uint32_t &e, uint32_t &y) {               // fse.rs:363-373: new_state + bits, then its entry
        const uint32_t bits = take(e);
        if (COMPACT) {
            const uint32_t ad = (e >> 4) + tab_s + bits * 2;  // bit 4 of an entry is 0: e >> 4 = 2 * base
            e = lds_u16(ad);
            y = lds_u8((ad >> 1) + sym_c);
        } else {
            e = lds_u32((e >> 16) + tab_s + bits * 4);
            y = e;
        }
    };
    auto sym_of = [&](uint32_t y) -> uint8_t { return (uint8_t)(COMPACT ? y : y >> 8); };
    // four symbols -> one word: the symbol is byte 1 of a wide entry, byte 0 of y in the compact form
    auto word_of = [&](uint32_t y0, uint32_t y1, uint32_t y2, uint32_t y3) -> uint32_t {
        return COMPACT ? __byte_perm(__byte_perm(y0, y1, 0x0040), __byte_perm(y2, y3, 0x0040), 0x5410)
                       : __byte_perm(__byte_perm(y0, y1, 0x0051), __byte_perm(y2, y3, 0x0051), 0x5410);
    };
    auto account = [&](uint32_t e0_, uint32_t e1_) {          // num_bits are the low five bits of an entry, two of them < 32
        cnt -= (e0_ + e1_) & 31u;
    };
    const uint32_t body = bn - N;
    uint32_t i = 0;
    if (N == 2) {
        // eA: entry of the state whose turn it is (state i & 1 decodes symbol i: Decoder::new reads state 0 first, fse.rs:349-352)
        const uint32_t s0 = take(log2), s1 = take(log2);
        cnt -= 2 * log2;
        refill();
        uint32_t eA, eB, yA, yB;
        look(eA, yA, s0);
        look(eB, yB, s1);
        auto single = [&]() {                                 // one symbol, then the other state's turn
            out[i] = sym_of(yA);
            const uint32_t kb = eA & 31u;
            step(eA, yA);
            cnt -= kb;
            refill();
            uint32_t x = eA; eA = eB; eB = x;
            x = yA; yA = yB; yB = x;
            i++;
        };
        while (i < body && (((uintptr_t)(out + i)) & 3)) single();      // up to an aligned output word
        auto quad = [&]() -> uint32_t {
            const uint32_t a0 = eA, a1 = eB, y0 = yA, y1 = yB;
            step(eA, yA); step(eB, yB);
            account(a0, a1);
            refill();
            const uint32_t a2 = eA, a3 = eB, y2 = yA, y3 = yB;
            step(eA, yA); step(eB, yB);
            account(a2, a3);
            refill();
            return word_of(y0, y1, y2, y3);
        };
        for (; i + 8 <= body; i += 8) {
            const uint32_t w0 = quad();
            const uint32_t w1 = quad();
            __stcs(reinterpret_cast<uint32_t *>(out + i), w0);
            __stcs(reinterpret_cast<uint32_t *>(out + i + 4), w1);
        }
        while (i < body) single();
        out[i] = sym_of(yA);                                 // Decoder::finish: symbols body, body + 1 from the states in turn
        out[i + 1] = sym_of(yB);
    } else {
        const uint32_t s = take(log2);
        cnt -= log2;
        refill();
        uint32_t e, y;
        look(e, y, s);
        auto single = [&]() {
            out[i] = sym_of(y);
            const uint32_t kb = e & 31u;
            step(e, y);
            cnt -= kb;
            refill();
            i++;
        };
        while (i < body && (((uintptr_t)(out + i)) & 3)) single();
        auto quad = [&]() -> uint32_t {
            const uint32_t a0 = e, y0 = y; step(e, y);
            const uint32_t a1 = e, y1 = y; step(e, y);
            account(a0, a1);
            refill();
            const uint32_t a2 = e, y2 = y; step(e, y);
            const uint32_t a3 = e, y3 = y; step(e, y);
            account(a2, a3);
            refill();
            return word_of(y0, y1, y2, y3);
        };
        for (; i + 8 <= body; i += 8) {
            const uint32_t w0 = quad();
            const uint32_t w1 = quad();
            __stcs(reinterpret_cast<uint32_t *>(out + i), w0);
            __stcs(reinterpret_cast<uint32_t *>(out + i + 4), w1);
        }
        while (i < body) single();
        out[i] = sym_of(y);
    }
    const uint32_t used = entered0 - 32u * (uint32_t)k - cnt;
    a.status[b] = (used != cur - floor_bits) ? ST_LENGTH : ST_OK;   // ran dry, or bits left over (lib.rs:205,245)
}

// The same streams with their tables in SHARED memory: a CTA takes `per_cta` blocks (as many as tables fit: 28 at
// table_log 11, 37 in the compact form), copies their tables in, and lanes 0 .. lanes_per_warp - 1 of its warps run one
// stream each.
template <bool COMPACT>
__global__ void __launch_bounds__(1024) k_tps_decode_smem(DecArgs a, TpsTables g, uint32_t per_cta, uint32_t lanes_per_warp)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t first = blockIdx.x * per_cta, size = 1u << a.tlmax;
    const uint32_t set_bytes = COMPACT ? 3u * size : 4u * size;
    for (uint32_t j = 0; j < per_cta; j++) {
        const uint32_t b = first + j;
        if (b >= a.nblocks) break;
        const uint4 *src = reinterpret_cast<const uint4 *>(g.dec_tab + ((size_t)b << a.tlmax));
        uint8_t *set = smem_raw + (size_t)j * set_bytes;
        for (uint32_t i = threadIdx.x; i < size / 4; i += blockDim.x) {
            const uint4 v = src[i];                           // DecodeTransform: new_state base | symbol << 16 | num_bits << 24
            if (COMPACT) {
                auto c16 = [](uint32_t e) -> uint32_t { return (e >> 24) | ((e & 0xffffu) << 5); };
                reinterpret_cast<uint2 *>(set)[i] = make_uint2(c16(v.x) | c16(v.y) << 16, c16(v.z) | c16(v.w) << 16);
                reinterpret_cast<uint32_t *>(set + 2u * size)[i] =
                    ((v.x >> 16) & 0xffu) | ((v.y >> 8) & 0xff00u) | (v.z & 0xff0000u) | ((v.w << 8) & 0xff000000u);
            } else {
                reinterpret_cast<uint4 *>(set)[i] = make_uint4(tps_smem_entry(v.x), tps_smem_entry(v.y), tps_smem_entry(v.z), tps_smem_entry(v.w));
            }
        }
    }
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane >= lanes_per_warp) return;
    const uint32_t j = warp * lanes_per_warp + lane, b = first + j;
    if (j >= per_cta || b >= a.nblocks) return;
    const uint4 m = g.meta[b];
    if (!m.z) return;
    const uint32_t tab_s = (uint32_t)__cvta_generic_to_shared(smem_raw + (size_t)j * set_bytes);
    tps_decode_stream_smem<COMPACT>(a, b, m, tab_s, COMPACT ? tab_s + 2u * size - (tab_s >> 1) : 0u);
}

}  // namespace fsed
