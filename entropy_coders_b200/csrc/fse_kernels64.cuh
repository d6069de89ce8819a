// fse_kernels64.cuh -- the 64-state fast path: one warp per block, two ADJACENT states per lane.
//
// Lane l owns states 2l and 2l+1, i.e. the symbol pairs (2m+1, 2m) with m == l (mod 32).  The two
// fields of a pair are neighbours in the bit stream, so they are merged in registers and every
// later stage (staging, bit-string building, prefix scan, bit fetch, output store) handles one pair
// per lane per step: half the shuffles, shared-memory traffic and bookkeeping per symbol of the
// 32-state path, and two independent state chains per lane to hide shared-memory latency.
// Requires table_log <= 13 (a merged pair is at most 26 bits + 5 bits of length in one word).
#pragma once
#include "fse_kernels.cuh"

namespace fsed {

constexpr int ROW_STRIDE64 = 29;                 // 832 + 31 bits = 27 words, odd stride
constexpr int ROWS64_WORDS = 32 * ROW_STRIDE64;
constexpr uint32_t PAIR_LEN_SHIFT = 27;
constexpr uint32_t PAIR_VAL_MASK = (1u << PAIR_LEN_SHIFT) - 1u;

struct Enc64Layout {
    uint32_t tab, tt, work, rows, total;
};
__host__ __device__ inline Enc64Layout enc64_layout(uint32_t tlmax)
{
    Enc64Layout l;
    uint32_t size = 1u << tlmax;
    l.tab = 0;
    l.tt = l.tab + size * 2;
    l.work = l.tt + 2048;
    uint32_t build = 3072 + size, enc = 4096;
    l.rows = l.work + (build > enc ? build : enc);
    l.total = (l.rows + ROWS64_WORDS * 4 + 15u) & ~15u;
    return l;
}

__device__ __forceinline__ uint32_t lds_u16(uint32_t saddr)
{
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(saddr));
    return v;
}
__device__ __forceinline__ uint2 lds_v2(uint32_t saddr)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(saddr));
    return v;
}

// One transition (fse.rs:227-239) with the symbol transform's find_state pre-scaled to a byte
// address: tt.y = tab_saddr + 2 * find_state.
__device__ __forceinline__ void enc_step(uint32_t tt_saddr, uint32_t sym, uint32_t &state, uint32_t &v, uint32_t &bo)
{
    uint2 t = lds_v2(tt_saddr + sym * 8);
    bo = (t.x + state) >> 16;
    v = state & ((1u << bo) - 1u);
    state = lds_u16(t.y + ((state >> bo) << 1));
}
// Encoder::new_first_symbol, fse.rs:210-218
__device__ __forceinline__ uint32_t enc_first64(uint32_t tt_saddr, uint32_t sym)
{
    uint2 t = lds_v2(tt_saddr + sym * 8);
    uint32_t bo = (t.x + (1u << 15)) >> 16;
    uint32_t value = (bo << 16) - t.x;
    return lds_u16(t.y + ((value >> bo) << 1));
}

// element classes of a symbol index
//   i >= bn          : does not exist
//   bn-64 <= i < bn  : initialises its state at no cost (lib.rs:123,155-165)
//   0 <= i < bn-64   : transition
__device__ __forceinline__ void enc_element_checked(const uint8_t *__restrict__ bsrc, int32_t i, int32_t bn, uint32_t tt_saddr,
                                                    uint32_t &state, uint32_t &v, uint32_t &bo)
{
    v = 0; bo = 0;
    if (i < 0 || i >= bn) return;
    uint32_t sym = __ldg(bsrc + i);
    if (i >= bn - 64) state = enc_first64(tt_saddr, sym);
    else enc_step(tt_saddr, sym, state, v, bo);
}

__device__ void encode64_payload_warp(const uint8_t *__restrict__ bsrc, uint32_t bn, uint32_t log2,
                                      uint32_t tab_saddr_unused, uint32_t tt_saddr, uint32_t *fld, uint32_t *rows,
                                      uint32_t *pay, uint32_t cap_words, int lane, uint32_t &bits_out, bool &overflow)
{
    (void)tab_saddr_unused;
    const int32_t M = (int32_t)((bn + 1) >> 1);                    // pairs (2m+1, 2m)
    const uint32_t kcol = (uint32_t)(M - 1 - lane) & 31;           // my pair's stream position in a round
    const int32_t mtop = M - 1 - (int32_t)kcol;                    // my pair in round 0
    const uint32_t G = (uint32_t)(M + 31) >> 5;                    // rounds of 32 pairs
    const bool aligned2 = (((uintptr_t)bsrc) & 1) == 0;
    uint32_t st0 = 0, st1 = 0;                                     // states 2*lane, 2*lane+1
    uint32_t cw = 0, cb = 0, wdone = 0;
    uint32_t *myrow = rows + lane * ROW_STRIDE64;
    uint32_t *obuf = fld;
    overflow = false;

    uint32_t sy[32];                                               // my pair of symbols for each round of a chunk
    // A chunk is "plain" when all of its 64*32 elements are existing transitions (warp uniform).
    auto plain = [&](uint32_t g0) -> bool { return g0 >= 32 && (uint32_t)M >= (g0 + 32) * 32; };
    auto fetch = [&](uint32_t g0) {
        if (plain(g0) && aligned2) {                               // one 16-bit load per pair, immediate offsets
            const uint16_t *p16 = reinterpret_cast<const uint16_t *>(bsrc) + (mtop - (int32_t)(g0 << 5));
#pragma unroll
            for (int r = 0; r < 32; r++) sy[r] = (uint32_t)__ldg(p16 - 32 * r);
        } else {
#pragma unroll
            for (int r = 0; r < 32; r++) {
                int32_t m = mtop - (int32_t)((g0 + r) << 5);
                uint32_t s = 0;
                if (m >= 0) {
                    s = (uint32_t)__ldg(bsrc + 2 * m);
                    if (2 * m + 1 < (int32_t)bn) s |= (uint32_t)__ldg(bsrc + 2 * m + 1) << 8;
                }
                sy[r] = s;
            }
        }
    };
    fetch(0);
    for (uint32_t g0 = 0; g0 < G; g0 += 32) {
        if (plain(g0)) {
            // pass 1, middle chunk: 32 rounds, two transitions per lane per round
#pragma unroll
            for (int r = 0; r < 32; r++) {
                uint32_t v1, b1, v0, b0;
                enc_step(tt_saddr, sy[r] >> 8, st1, v1, b1);       // index 2m+1 first (decreasing index order)
                enc_step(tt_saddr, sy[r] & 0xff, st0, v0, b0);
                fld[r * 32 + (kcol ^ ((r & 7) << 2))] = (v1 | (v0 << b1)) | ((b1 + b0) << PAIR_LEN_SHIFT);
            }
        } else {
            // first and last chunk: elements may be missing or be initial symbols
            for (int r = 0; r < 32; r++) {
                int32_t m = mtop - (int32_t)((g0 + r) << 5);
                uint32_t v1, b1, v0, b0;
                enc_element_checked(bsrc, m < 0 ? -1 : 2 * m + 1, (int32_t)bn, tt_saddr, st1, v1, b1);
                enc_element_checked(bsrc, m < 0 ? -1 : 2 * m, (int32_t)bn, tt_saddr, st0, v0, b0);
                fld[r * 32 + (kcol ^ ((r & 7) << 2))] = (v1 | (v0 << b1)) | ((b1 + b0) << PAIR_LEN_SHIFT);
            }
        }
        if (g0 + 32 < G) fetch(g0 + 32);
        __syncwarp();
        // pass 2: lane L serialises round L: 32 merged pairs that are consecutive in the stream
        BitRow br;
        br.init(myrow, lane == 0 ? cw : 0u, lane == 0 ? cb : 0u);
#pragma unroll
        for (int q = 0; q < 8; q++) {
            uint4 x = *reinterpret_cast<const uint4 *>(fld + lane * 32 + ((q ^ (lane & 7)) << 2));
            br.put(x.x & PAIR_VAL_MASK, x.x >> PAIR_LEN_SHIFT);
            br.put(x.y & PAIR_VAL_MASK, x.y >> PAIR_LEN_SHIFT);
            br.put(x.z & PAIR_VAL_MASK, x.z >> PAIR_LEN_SHIFT);
            br.put(x.w & PAIR_VAL_MASK, x.w >> PAIR_LEN_SHIFT);
        }
        uint32_t tot = br.finish();
        __syncwarp();
        uint32_t nw = warp_place(myrow, tot, obuf, 1024, lane, cw, cb, overflow);
        __syncwarp();
        if (wdone + nw > cap_words) { overflow = true; nw = 0; }
        for (uint32_t j = lane; j < nw; j += 32) pay[wdone + j] = obuf[j];
        wdone += nw;
        __syncwarp();
    }
    // final states 63 .. 0 (fse.rs:248-250), then the marker bit (lib.rs:141,181):
    // stream position L holds states 63-2L, 62-2L, both owned by lane 31-L
    {
        uint32_t s1 = __shfl_sync(FULL, st1, 31 - lane), s0 = __shfl_sync(FULL, st0, 31 - lane);
        const uint32_t mask = (1u << log2) - 1u;
        BitRow br;
        br.init(myrow, lane == 0 ? cw : 0u, lane == 0 ? cb : 0u);
        br.put(s1 & mask, log2);
        br.put(s0 & mask, log2);
        if (lane == 31) br.put(1, 1);
        uint32_t tot = br.finish();
        __syncwarp();
        wdone += warp_place(myrow, tot, pay + wdone, cap_words > wdone ? cap_words - wdone : 0, lane, cw, cb, overflow);
        __syncwarp();
    }
    if (cb) {
        if (wdone < cap_words) { if (lane == 0) pay[wdone] = cw; }
        else overflow = true;
    }
    bits_out = wdone * 32 + cb;
}

__global__ void __launch_bounds__(512) k_encode64_blocks(EncArgs a)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    const Enc64Layout lay = enc64_layout(a.tlmax);
    uint8_t *my = smem_raw + (size_t)warp * lay.total;
    uint16_t *tab = reinterpret_cast<uint16_t *>(my + lay.tab);
    uint2 *tt = reinterpret_cast<uint2 *>(my + lay.tt);
    uint32_t *cnt = reinterpret_cast<uint32_t *>(my + lay.work);
    int32_t *norm = reinterpret_cast<int32_t *>(my + lay.work + 1024);
    uint32_t *cum = reinterpret_cast<uint32_t *>(my + lay.work + 2048);
    uint8_t *spread = my + lay.work + 3072;
    uint32_t *fld = reinterpret_cast<uint32_t *>(my + lay.work);
    uint32_t *rows = reinterpret_cast<uint32_t *>(my + lay.rows);
    const uint32_t tab_saddr = (uint32_t)__cvta_generic_to_shared(tab);
    const uint32_t tt_saddr = (uint32_t)__cvta_generic_to_shared(tt);

    uint32_t glog2 = 0;
    if (a.global_mode) {
        glog2 = a.g.log2;
        for (uint32_t i = lane; i < (1u << glog2); i += 32) tab[i] = a.g.enc_table[i];
        for (uint32_t i = lane; i < 256; i += 32) {
            uint2 t = a.g.enc_tt[i];
            t.y = tab_saddr + 2u * t.y;
            tt[i] = t;
        }
        __syncwarp();
    }

    for (uint32_t b = blockIdx.x * wpc + warp; b < a.nblocks; b += gridDim.x * wpc) {
        const size_t off = (size_t)b * a.block_size;
        const uint32_t bn = (uint32_t)min((size_t)a.block_size, a.n - off);
        const uint8_t *bsrc = a.src + off;
        uint8_t *bs = a.scratch + (size_t)b * a.stride;
        uint32_t *hdr_words = reinterpret_cast<uint32_t *>(bs);
        uint32_t *pay = reinterpret_cast<uint32_t *>(bs + HDR_RESERVE);
        uint32_t log2 = glog2, hbytes = 0;
        int st = ST_OK;

        if (!a.global_mode) {
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 8; k++) cnt[k * 32 + lane] = a.counts[(size_t)b * 256 + k * 32 + lane];
            __syncwarp();
            uint32_t table_len;
            int rc = warp_normalize(cnt, (uint64_t)bn, a.req_log2, norm, lane, log2, table_len);
            if (rc < 0) {
                if (table_len <= 1) {
                    if (lane == 0) { bs[0] = 0x0E; bs[1] = 0x00; a.hlen[b] = 2; a.plen[b] = 0; a.status[b] = 2; }
                } else if (bn <= 4) {
                    if (lane == 0) {
                        bs[0] = 0x0F;
                        for (uint32_t i = 0; i < bn; i++) bs[1 + i] = bsrc[i];
                        a.hlen[b] = 1 + bn; a.plen[b] = 0; a.status[b] = 1;
                    }
                } else if (lane == 0) { a.hlen[b] = 0; a.plen[b] = 0; a.status[b] = rc; }
                continue;
            }
            if (bn < 64) {                       // fewer symbols than states: stored raw
                for (uint32_t i = lane; i < bn; i += 32) bs[1 + i] = bsrc[i];
                if (lane == 0) { bs[0] = 0x0F; a.hlen[b] = 1 + bn; a.plen[b] = 0; a.status[b] = 1; }
                continue;
            }
            if (log2 > a.tlmax || log2 > 13) {
                if (lane == 0) { a.hlen[b] = 0; a.plen[b] = 0; a.status[b] = ST_UNSUPPORTED; }
                continue;
            }
            uint32_t hbits = warp_ncount_write(norm, log2, table_len, rows, hdr_words, lane);
            hbytes = (hbits + 7) >> 3;
            warp_spread(norm, log2, table_len, spread, cum, tab, lane);
            warp_build_encode(norm, log2, table_len, spread, cum, tab, tt, lane);
#pragma unroll
            for (int k = 0; k < 8; k++) {        // pre-scale find_state to a shared-memory byte address
                uint2 t = tt[k * 32 + lane];
                t.y = tab_saddr + 2u * t.y;
                tt[k * 32 + lane] = t;
            }
            __syncwarp();
        } else if (bn < 64) {
            for (uint32_t i = lane; i < bn; i += 32) bs[i] = bsrc[i];
            if (lane == 0) { a.hlen[b] = bn; a.plen[b] = 0; a.status[b] = 1; }
            continue;
        }
        uint32_t pbits;
        bool ovf;
        encode64_payload_warp(bsrc, bn, log2, tab_saddr, tt_saddr, fld, rows, pay, a.pay_cap_words, lane, pbits, ovf);
        if (ovf) st = ST_CAPACITY;
        uint32_t hl = ovf ? 0 : hbytes, pl = ovf ? 0 : (pbits + 7) >> 3;
        if (!ovf && !a.global_mode && warp_raw_if_expands(a.flags, hl, pl, bs, bsrc, bn, lane)) { hl = 1; pl = bn; st = 1; }
        if (lane == 0) {
            a.hlen[b] = hl;
            a.plen[b] = pl;
            a.status[b] = st;
        }
    }
}

// ------------------------------------------------------------------------------------------
// decode, 64 states: lane l owns states 2l and 2l+1 = output bytes 64g+2l and 64g+2l+1 of round g
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) k_decode64_blocks(DecArgs a)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    const DecLayout lay = dec_layout(a.tlmax);
    uint8_t *my = smem_raw + (size_t)warp * lay.total;
    uint32_t *tab = reinterpret_cast<uint32_t *>(my + lay.tab);
    int32_t *norm = reinterpret_cast<int32_t *>(my + lay.norm);
    uint32_t *ctr = reinterpret_cast<uint32_t *>(my + lay.ctr);
    uint8_t *spread = my + lay.spread;
    uint32_t *ring = reinterpret_cast<uint32_t *>(my + lay.ring);
    const uint32_t N = 64;

    uint32_t glog2 = 0;
    if (a.global_mode) {
        glog2 = a.g.log2;
        for (uint32_t i = lane; i < (1u << glog2); i += 32) tab[i] = a.g.dec_table[i];
        __syncwarp();
    }

    for (uint32_t b = blockIdx.x * wpc + warp; b < a.nblocks; b += gridDim.x * wpc) {
        const size_t off = (size_t)b * a.block_size;
        const uint32_t bn = (uint32_t)min((size_t)a.block_size, a.n - off);
        uint8_t *out = a.dst + off;
        int st = ST_OK;
        const uint8_t *cs;
        uint32_t clen;
        if (dec_block_prologue(a, b, bn, N, out, lane, cs, clen, st)) {      // bad offsets, raw tail, escape blocks
            if (lane == 0) { a.status[b] = st; if (a.exhaust) a.out_len[b] = 0; }
            continue;
        }
        uint32_t log2 = glog2, consumed = 0;
        __syncwarp();
        if (!a.global_mode) {
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
            if (rc < 0) { if (lane == 0) a.status[b] = rc; continue; }
            if (log2 > a.tlmax || log2 > 13) { if (lane == 0) a.status[b] = ST_UNSUPPORTED; continue; }
            warp_spread(norm, log2, table_len, spread, ctr, reinterpret_cast<uint16_t *>(tab), lane);
            warp_build_decode(norm, log2, table_len, spread, ctr, tab, lane);
        }
        if (bn < N) { if (lane == 0) a.status[b] = ST_LENGTH; continue; }
        const uint8_t *pay = cs + consumed;
        const uint32_t plen = clen - consumed;
        if (plen == 0 || pay[plen - 1] == 0) { if (lane == 0) a.status[b] = ST_NO_MARKER; continue; }
        const uint32_t bias = (uint32_t)((uintptr_t)pay & 3);
        const uint32_t *origin = reinterpret_cast<const uint32_t *>(pay - bias);
        uint32_t cur = (plen - 1) * 8 + ilog2u(pay[plen - 1]) + 8 * bias;
        const uint32_t floor_bits = 8 * bias;
        if (cur - floor_bits < N * log2) { if (lane == 0) a.status[b] = ST_LENGTH; continue; }
        const uint32_t topq = cur >> 5;
        uint32_t lowq = (topq & ~127u) >= 128 ? (topq & ~127u) - 128 : 0;
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 8; k++) {
            uint32_t w = lowq + lane + 32 * k;
            if (w <= topq) ring[w & 255] = __ldg(origin + w);
        }
        uint32_t pre[4];
#pragma unroll
        for (int k = 0; k < 4; k++) pre[k] = (lowq >= 128) ? __ldg(origin + lowq - 128 + lane + 32 * k) : 0u;
        __syncwarp();
        auto ring_bits = [&](uint32_t q, uint32_t nb) -> uint32_t {
            uint32_t w = q >> 5;
            return __funnelshift_r(ring[w & 255], ring[(w + 1) & 255], q & 31) & ((1u << nb) - 1u);
        };
        // Decoder::new, fse.rs:349-352: states are read 0, 1, 2, ... from the top of the stack
        uint32_t st0, st1;
        {
            uint32_t w = ring_bits(cur - (2 * lane + 2) * log2, 2 * log2);
            st0 = w >> log2;
            st1 = w & ((1u << log2) - 1u);
        }
        cur -= N * log2;
        const uint32_t body = bn - N;
        const bool out_aligned = (((uintptr_t)out) & 1) == 0;
        bool bad = false;
        uint32_t i0 = 0;
        // full rounds: 64 transitions
        for (; i0 + 64 <= body; i0 += 64) {
            if ((cur >> 5) < lowq + 28 && lowq) {           // a round takes at most 26 words
                __syncwarp();
#pragma unroll
                for (int k = 0; k < 4; k++) ring[(lowq - 128 + lane + 32 * k) & 255] = pre[k];
                lowq -= 128;
#pragma unroll
                for (int k = 0; k < 4; k++) pre[k] = (lowq >= 128) ? __ldg(origin + lowq - 128 + lane + 32 * k) : 0u;
                __syncwarp();
            }
            uint32_t e0 = tab[st0], e1 = tab[st1];          // fse.rs:363-373, two independent chains
            uint32_t nb0 = e0 >> 24, nb1 = e1 >> 24;
            uint32_t nbs = nb0 + nb1;
            uint32_t incl = warp_incl_add5(nbs, lane);      // nbs <= 26
            uint32_t w = ring_bits(cur - incl, nbs);        // state 2l's bits are the upper part
            uint32_t tot = __shfl_sync(FULL, incl, 31);     // off the critical path: a bad stream only reads stale ring words
            if (tot > cur - floor_bits) { bad = true; break; }
            st0 = (e0 & 0xffffu) + (w >> nb1);
            st1 = (e1 & 0xffffu) + (w & ((1u << nb1) - 1u));
            uint32_t sy = ((e0 >> 16) & 0xffu) | ((e1 >> 8) & 0xff00u);
            if (out_aligned) *reinterpret_cast<uint16_t *>(out + i0 + 2 * lane) = (uint16_t)sy;
            else { out[i0 + 2 * lane] = (uint8_t)sy; out[i0 + 2 * lane + 1] = (uint8_t)(sy >> 8); }
            cur -= tot;
        }
        // last partial round
        if (!bad && i0 < body) {
            if ((cur >> 5) < lowq + 28 && lowq) {
                __syncwarp();
#pragma unroll
                for (int k = 0; k < 4; k++) ring[(lowq - 128 + lane + 32 * k) & 255] = pre[k];
                lowq -= 128;
                __syncwarp();
            }
            uint32_t ia = i0 + 2 * lane, ib = ia + 1;
            uint32_t e0 = tab[st0], e1 = tab[st1];
            uint32_t nb0 = ia < body ? (e0 >> 24) : 0u, nb1 = ib < body ? (e1 >> 24) : 0u;
            uint32_t nbs = nb0 + nb1;
            uint32_t incl = warp_incl_add(nbs, lane);
            uint32_t tot = __shfl_sync(FULL, incl, 31);
            if (tot > cur - floor_bits) bad = true;
            else {
                uint32_t w = ring_bits(cur - incl, nbs);
                if (ia < body) { out[ia] = (uint8_t)(e0 >> 16); st0 = (e0 & 0xffffu) + (w >> nb1); }
                if (ib < body) { out[ib] = (uint8_t)(e1 >> 16); st1 = (e1 & 0xffffu) + (w & ((1u << nb1) - 1u)); }
                cur -= tot;
            }
        }
        if (!bad) {                                         // Decoder::finish, fse.rs:383-385: i in [body, bn), state i % 64
            uint32_t ia = body + ((2 * lane - body) & 63), ib = body + ((2 * lane + 1 - body) & 63);
            out[ia] = (uint8_t)(tab[st0] >> 16);
            out[ib] = (uint8_t)(tab[st1] >> 16);
        }
        cur -= floor_bits;
        if (bad || cur != 0) st = ST_LENGTH;
        if (lane == 0) a.status[b] = st;
    }
}

}  // namespace fsed
